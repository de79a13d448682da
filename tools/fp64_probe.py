"""Measured fp64 issue peak of this GPU (the geometry path is fp64-issue / latency bound and
MEASURED_PEAKS.json has no fp64 figure).  Compiles a DFMA microbenchmark with nvcc for sm_100a, runs it, prints
one JSON line: DFMA warp-instructions/s, TFLOP/s, and the same per SM and clock.

    python tools/fp64_probe.py            # on a GPU box
"""
import ctypes as C
import json
import os
import subprocess
import tempfile

import torch

SRC = r'''
#include <cuda_runtime.h>
template <int CHAINS>
__global__ void dfma_kernel(double* out, int iters, double a, double b) {
    double x[CHAINS];
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) x[c] = threadIdx.x * 1e-9 + c;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int c = 0; c < CHAINS; ++c) x[c] = fma(x[c], a, b);
    }
    double s = 0;
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) s += x[c];
    if (s == 123.456) out[0] = s;            // never true; keeps the chains alive
}
extern "C" int probe_launch(double* out, int grid, int block, int iters, void* stream) {
    dfma_kernel<8><<<grid, block, 0, (cudaStream_t)stream>>>(out, iters, 1.0000001, 1e-9);
    return (int)cudaGetLastError();
}
'''


def main():
    tmp = tempfile.mkdtemp(prefix='fp64probe')
    cu, so = os.path.join(tmp, 'p.cu'), os.path.join(tmp, 'p.so')
    with open(cu, 'w') as f:
        f.write(SRC)
    subprocess.run(['nvcc', '-shared', '-Xcompiler', '-fPIC', '-O3', '-gencode', 'arch=compute_100a,code=sm_100a', cu, '-o', so],
                   check=True)
    lib = C.CDLL(so)
    lib.probe_launch.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
    out = torch.zeros(8, dtype=torch.float64, device='cuda')
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    grid, block, iters, chains = sms * 8, 256, 20000, 8
    st = torch.cuda.current_stream().cuda_stream
    for _ in range(3):
        assert lib.probe_launch(out.data_ptr(), grid, block, iters, st) == 0
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 5
    e0.record()
    for _ in range(reps):
        lib.probe_launch(out.data_ptr(), grid, block, iters, st)
    e1.record()
    torch.cuda.synchronize()
    sec = e0.elapsed_time(e1) * 1e-3 / reps
    dfma = grid * block * iters * chains
    mhz = float(subprocess.run(['nvidia-smi', '--query-gpu=clocks.sm', '--format=csv,noheader,nounits', '-i', '0'],
                               capture_output=True, text=True).stdout.strip() or 0)
    print(json.dumps({'what': 'fp64 DFMA issue peak (8 independent chains per thread, 8 CTAs x 256 threads per SM)',
                      'dfma_thread_per_s': dfma / sec, 'dfma_warp_instr_per_s': dfma / 32 / sec,
                      'tflops_fp64': 2 * dfma / sec / 1e12, 'sms': sms, 'sm_mhz_after': mhz,
                      'dfma_per_sm_per_clk_at_1900MHz': dfma / sec / sms / 1.9e9}))


if __name__ == '__main__':
    main()
