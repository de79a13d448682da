"""Store-order probe 2 (tools/probe/store_probe2.cu): free-running strips vs. the strips of one crop written in lockstep by one CTA.
Build here (nvcc cross-compiles), run on a GPU box:  python tools/probe/store_probe2.py [--build-only]"""
import ctypes as C
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, 'store_probe2.so')

if __name__ == '__main__':
    if not os.path.exists(SO) or '--build-only' in sys.argv:
        subprocess.run(['nvcc', '-shared', '-Xcompiler', '-fPIC', '-O3', '-std=c++17', '-gencode', 'arch=compute_100a,code=sm_100a',
                        os.path.join(HERE, 'store_probe2.cu'), '-o', SO], check=True)
    if '--build-only' in sys.argv:
        sys.exit(0)
    import torch
    lib = C.CDLL(SO)
    lib.probe_store2.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int]
    R = 16384
    counter = torch.zeros(1, dtype=torch.int32, device='cuda')
    for T in [int(v) for v in os.environ.get('TS', '224').split(',')]:
        out = torch.empty((R, 3, T, T), dtype=torch.float32, device='cuda')
        for occ in [int(v) for v in os.environ.get('OCC', '3').split(',')]:
            for work in [int(v) for v in os.environ.get('WORK', '0,32,64').split(',')]:
                line = f'T {T} ctas/SM {occ} work {work:3d}:'
                for mode, sr in ((0, 0), (2, 0), (1, 4), (1, 8), (1, 16), (1, 32), (3, 8), (3, 16), (3, 32)):
                    for _ in range(2):
                        rc = lib.probe_store2(out.data_ptr(), R, T, work, mode, sr, counter.data_ptr(), occ)
                    assert rc == 0, rc
                    torch.cuda.synchronize()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    for _ in range(5):
                        lib.probe_store2(out.data_ptr(), R, T, work, mode, sr, counter.data_ptr(), occ)
                    e1.record()
                    torch.cuda.synchronize()
                    ms = e0.elapsed_time(e1) / 5
                    line += f'  m{mode}/{sr} {ms:5.3f}ms {R * 3 * T * T * 4 / ms / 1e6:5.0f}'
                print(line, flush=True)
        del out
