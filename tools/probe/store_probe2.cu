// Store-order probe, second round: does the crop kernel's output reach DRAM faster when the strips of one crop are written
// in (loose) lockstep by the warps of ONE CTA?  R crops of [3][T][T] float32; `work` FMAs per row in four independent chains
// stand in for the resize arithmetic.
//   mode 0  free-running warps, item = (crop, 32-column strip) from a global counter (what bpc_crop_warp_kernel did in round 1)
//   mode 1  CTA = crop, warp w = strip w, __syncthreads every `sync_rows` rows
//   mode 2  CTA = crop, warp w = strip w, no barrier inside a crop (only between crops)
//   mode 3  as mode 1, but a warp may run ahead of the slowest warp by up to `sync_rows` rows (counter in shared memory)
#include <cuda_runtime.h>
#include <stdint.h>

template <int MODE>
__global__ void __launch_bounds__(256)
probe2_kernel(float* __restrict__ out, int R, int T, int work, int sync_rows, int* __restrict__ counter) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ int s_item;
    __shared__ volatile int s_row[8];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int nstrip = T / 32;
    const size_t plane = (size_t)T * T;
    float a0 = (float)lane, a1 = a0 + 1.f, a2 = a0 + 2.f, a3 = a0 + 3.f;
    const int w4 = work / 4;
    if (MODE == 0) {
        for (;;) {
            int item = 0;
            if (lane == 0) item = atomicAdd(counter, 1);
            item = __shfl_sync(0xffffffffu, item, 0);
            if (item >= R * nstrip) break;
            const int roi = item / nstrip, strip = item - roi * nstrip;
            float* o = out + (size_t)roi * 3 * plane + strip * 32 + lane;
            for (int y = 0; y < T; ++y) {
                for (int k = 0; k < w4; ++k) { a0 = fmaf(a0, 1.0001f, 0.5f); a1 = fmaf(a1, 1.0001f, 0.5f); a2 = fmaf(a2, 1.0001f, 0.5f); a3 = fmaf(a3, 1.0001f, 0.5f); }
                o[0] = a0 + a3; o[plane] = a1; o[2 * plane] = a2;
                o += T;
            }
        }
        return;
    }
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) s_item = atomicAdd(counter, 1);
        if (threadIdx.x < 8) s_row[threadIdx.x] = (threadIdx.x < nstrip) ? 0 : (1 << 30);
        __syncthreads();
        const int roi = s_item;
        if (roi >= R) break;
        float* o = out + (size_t)roi * 3 * plane + wid * 32 + lane;
        for (int y = 0; y < T; ++y) {
            if (wid < nstrip) {
                for (int k = 0; k < w4; ++k) { a0 = fmaf(a0, 1.0001f, 0.5f); a1 = fmaf(a1, 1.0001f, 0.5f); a2 = fmaf(a2, 1.0001f, 0.5f); a3 = fmaf(a3, 1.0001f, 0.5f); }
                o[0] = a0 + a3; o[plane] = a1; o[2 * plane] = a2;
                o += T;
            }
            if (MODE == 1 && (y % sync_rows) == sync_rows - 1) __syncthreads();
            if (MODE == 3 && wid < nstrip && (y & 3) == 3) {
                if (lane == 0) s_row[wid] = y;
                // wait until the slowest warp is within sync_rows rows
                for (;;) {
                    int m = s_row[lane & 7];
                    m = min(m, __shfl_xor_sync(0xffffffffu, m, 1)); m = min(m, __shfl_xor_sync(0xffffffffu, m, 2)); m = min(m, __shfl_xor_sync(0xffffffffu, m, 4));
                    if (y - m <= sync_rows) break;
                    __nanosleep(20);
                }
            }
        }
    }
}

extern "C" int probe_store2(float* out, int R, int T, int work, int mode, int sync_rows, int* counter, int ctas_per_sm) {
    cudaMemsetAsync(counter, 0, 4, 0);
    const int smem = ((227 * 1024 / ctas_per_sm - 1024) / 128) * 128 - 256;
    void (*fn)(float*, int, int, int, int, int*) = mode == 0 ? probe2_kernel<0> : (mode == 1 ? probe2_kernel<1> : (mode == 2 ? probe2_kernel<2> : probe2_kernel<3>));
    cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    fn<<<148 * ctas_per_sm, 256, smem, 0>>>(out, R, T, work, sync_rows, counter);
    return (int)cudaGetLastError();
}
