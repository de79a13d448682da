"""Store-order probe 3 (tools/probe/store_probe3.cu): free-running STG strips vs. shared-memory output tiles written by bulk copies.
Build here (nvcc cross-compiles), run on a GPU box:  python tools/probe/store_probe3.py [--build-only]"""
import ctypes as C
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, 'store_probe3.so')

if __name__ == '__main__':
    if not os.path.exists(SO) or '--build-only' in sys.argv:
        subprocess.run(['nvcc', '-shared', '-Xcompiler', '-fPIC', '-O3', '-std=c++17', '-gencode', 'arch=compute_100a,code=sm_100a',
                        os.path.join(HERE, 'store_probe3.cu'), '-o', SO], check=True)
    if '--build-only' in sys.argv:
        sys.exit(0)
    import torch
    lib = C.CDLL(SO)
    lib.probe_store3.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int]
    R = 16384
    counter = torch.zeros(1, dtype=torch.int32, device='cuda')
    for T in [int(v) for v in os.environ.get('TS', '224').split(',')]:
        out = torch.empty((R, 3, T, T), dtype=torch.float32, device='cuda')
        for occ in [int(v) for v in os.environ.get('OCC', '3').split(',')]:
            for work in [int(v) for v in os.environ.get('WORK', '0,32,64').split(',')]:
                line = f'T {T} ctas/SM {occ} work {work:3d}:'
                for mode, kr in ((0, 4), (5, 4), (1, 8), (6, 8), (4, 4), (4, 8), (4, 16)):
                    rc = 0
                    for _ in range(2):
                        rc = lib.probe_store3(out.data_ptr(), R, T, work, mode, kr, counter.data_ptr(), occ)
                    if rc != 0:
                        line += f'  m{mode}/{kr} rc={rc}'
                        continue
                    torch.cuda.synchronize()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    for _ in range(5):
                        lib.probe_store3(out.data_ptr(), R, T, work, mode, kr, counter.data_ptr(), occ)
                    e1.record()
                    torch.cuda.synchronize()
                    ms = e0.elapsed_time(e1) / 5
                    line += f'  m{mode}/{kr} {ms:5.3f}ms {R * 3 * T * T * 4 / ms / 1e6:5.0f}'
                print(line, flush=True)
        del out
