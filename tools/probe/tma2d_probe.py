"""Validates the 2-D tensor-map staging (tools/probe/tma2d_probe.cu) against NumPy slicing, incl. out-of-bounds rows.
Build here (nvcc cross-compiles), run on a GPU box:  python tools/probe/tma2d_probe.py [--build-only]"""
import ctypes as C
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, 'tma2d_probe.so')


def build():
    subprocess.run(['nvcc', '-shared', '-Xcompiler', '-fPIC', '-O3', '-std=c++17', '-gencode', 'arch=compute_100a,code=sm_100a',
                    os.path.join(HERE, 'tma2d_probe.cu'), '-o', SO], check=True)


if __name__ == '__main__':
    if not os.path.exists(SO) or '--build-only' in sys.argv:
        build()
    if '--build-only' in sys.argv:
        sys.exit(0)
    import numpy as np
    import torch
    lib = C.CDLL(SO)
    lib.probe_run.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    B, H, W = 2, 64, 128                      # row pitch 384 bytes
    img = np.random.default_rng(0).integers(0, 256, (B, H, W, 3), dtype=np.uint8)
    d = torch.as_tensor(img).cuda()
    flat = img.reshape(B * H, W * 3)
    ok = True
    for (xb, row, pitch, nb) in ((0, 0, 64, 1), (16, 5, 160, 3), (128, 60, 256, 2), (240, 120, 224, 3), (320, 126, 96, 2)):
        out = torch.zeros(nb * 4 * pitch, dtype=torch.uint8, device='cuda')
        rc = lib.probe_run(d.data_ptr(), B, H, W, xb, row, pitch, nb, out.data_ptr(), None)
        torch.cuda.synchronize()
        got = out.cpu().numpy().reshape(nb * 4, pitch)
        want = np.zeros_like(got)
        for r in range(nb * 4):
            if row + r < B * H:
                seg = flat[row + r, xb:xb + pitch]
                want[r, :len(seg)] = seg
        same = rc == 0 and np.array_equal(got, want)
        print((xb, row, pitch, nb), 'rc', rc, 'OK' if same else 'MISMATCH')
        ok = ok and same
    print('TMA2D_OK' if ok else 'TMA2D_FAILED')
