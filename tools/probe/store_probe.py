"""Store-path probe (tools/probe/store_probe.cu): GB/s of the crop kernel's output pattern written with STG.32, with 4-D TMA
tensor stores from shared memory, and with STG.128 from shared memory, for several amounts of ALU work per row.
Build here (nvcc cross-compiles), run on a GPU box:  python tools/probe/store_probe.py [--build-only]"""
import ctypes as C
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, 'store_probe.so')


def build():
    subprocess.run(['nvcc', '-shared', '-Xcompiler', '-fPIC', '-O3', '-std=c++17', '-gencode', 'arch=compute_100a,code=sm_100a',
                    os.path.join(HERE, 'store_probe.cu'), '-o', SO], check=True)


if __name__ == '__main__':
    if not os.path.exists(SO) or '--build-only' in sys.argv:
        build()
    if '--build-only' in sys.argv:
        sys.exit(0)
    import torch
    lib = C.CDLL(SO)
    lib.probe_store.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    R, T = 16384, 224
    out = torch.empty((R, 3, T, T), dtype=torch.float32, device='cuda')
    counter = torch.zeros(1, dtype=torch.int32, device='cuda')
    clk = torch.zeros(2, dtype=torch.int64, device='cuda')
    ref = None
    occs = [int(v) for v in os.environ.get('OCC', '3').split(',')]
    for occ in occs:
      for rows in (224,):
        for work in [int(v) for v in os.environ.get('WORK', '0,32,64').split(',')]:
            line = f'ctas/SM {occ} rows {rows} work {work:3d}:'
            for mode in [int(v) for v in os.environ.get('MODES', '0,5,1').split(',')]:
                out.zero_()
                for _ in range(2):
                    rc = lib.probe_store(out.data_ptr(), R, T, rows, work, mode, counter.data_ptr(), None, occ, clk.data_ptr())
                assert rc == 0, rc
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(5):
                    lib.probe_store(out.data_ptr(), R, T, rows, work, mode, counter.data_ptr(), None, occ, clk.data_ptr())
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / 5
                nbytes = R * 3 * rows * T * 4
                chk = float(out[:64].double().sum())
                if mode == 0:
                    ref = chk
                line += f'  mode{mode} {ms:6.3f} ms {nbytes / ms / 1e6:7.0f} GB/s{"" if (chk == ref or mode > 1) else " MISMATCH"} {float(clk[0]) / max(float(clk[1]), 1) * 1e3:5.0f} MHz'
            print(line)
