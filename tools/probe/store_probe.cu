// Store-path probe for the crop kernel's output pattern: R crops of [3][T][T] float32, written by persistent warps that own
// (crop, 32-column strip) items and walk down the rows, with `work` dependent FMAs per row standing in for the resize.
//   mode 0  one STG.32 per lane, plane and row (what bpc_crop_warp_kernel does): 128 B per warp instruction
//   mode 1  rows staged in shared memory, one 4-D TMA tensor store ({32 cols, 4 rows, 3 planes, 1 crop}) per 4 rows
//   mode 2  rows staged in shared memory, lanes re-read 16 B each and write STG.128 (512 B per warp instruction)
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

constexpr int WARPS = 8;
constexpr int WSMEM = 2 * 3 * 4 * 128;   // two buffers of [3 planes][4 rows][32 floats]

template <int MODE>
__global__ void __launch_bounds__(256)
probe_kernel(float* __restrict__ out, int R, int T, int rows, int work, int* __restrict__ counter, const __grid_constant__ CUtensorMap map,
             unsigned long long* __restrict__ clk) {
    extern __shared__ __align__(128) unsigned char smem[];
    unsigned long long c0 = clock64(), g0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g0));
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    unsigned char* wb = smem + wid * WSMEM;
    const unsigned wb_s = (unsigned)__cvta_generic_to_shared(wb);
    const int nstrip = T / 32;
    const size_t plane = (size_t)T * T;
    const int y0 = (T - rows) / 2;
    float a0 = (float)lane;
    for (;;) {
        int item = 0;
        if (lane == 0) item = atomicAdd(counter, 1);
        item = __shfl_sync(0xffffffffu, item, 0);
        if (item >= R * nstrip) break;
        if (MODE == 6) {
            // what-if: a warp owns 32 full-width rows of one crop: 896 contiguous bytes per (row, plane), rows consecutive
            const int roi6 = item / nstrip, blk = item - roi6 * nstrip;           // nstrip = 7 blocks of 32 rows at T = 224
            for (int y = blk * 32; y < blk * 32 + 32 && y < rows; ++y) {
                for (int k = 0; k < work; ++k) a0 = fmaf(a0, 1.0001f, 0.5f);
                const float4 v = make_float4(a0, a0 + 1.f, a0 + 2.f, a0 + 3.f);
                for (int p = 0; p < 3; ++p) {
                    float4* q = reinterpret_cast<float4*>(out + ((size_t)(roi6 * 3 + p) * T + y0 + y) * T);
                    q[lane] = v;
                    if (lane + 32 < T / 4) q[lane + 32] = v;
                }
            }
            continue;
        }
        if (MODE == 9 || MODE == 10) {
            // what-if: the strips of one crop run in ONE CTA (warp w = strip w), MODE 9 in lockstep (barrier every 8 rows)
            __shared__ int s_item;
            for (;;) {
                __syncthreads();
                if (threadIdx.x == 0) s_item = atomicAdd(counter, 1);
                __syncthreads();
                const int roi9 = s_item;
                if (roi9 >= R) break;
                float a9 = (float)lane;
                float* o9 = out + ((size_t)roi9 * 3 * T + y0) * T + wid * 32 + lane;
                for (int y = 0; y < rows; ++y) {
                    if (wid < nstrip) {
                        for (int k = 0; k < work; ++k) a9 = fmaf(a9, 1.0001f, 0.5f);
                        o9[0] = a9; o9[plane] = a9 + 1.f; o9[2 * plane] = a9 + 2.f;
                        o9 += T;
                    }
                    if (MODE == 9 && (y & 7) == 7) __syncthreads();
                }
            }
            break;
        }
        if (MODE == 11) {
            // what-if: a warp owns a half-width band (4 or 3 strips = 512 / 384 contiguous bytes per plane and row)
            const int roi11 = item / nstrip, sub = item - roi11 * nstrip;           // sub 0 / 1 = halves; others: nothing
            if (sub < 2) {
                const int s0 = sub * 4, ns = sub ? nstrip - 4 : 4;
                float* o11 = out + ((size_t)roi11 * 3 * T + y0) * T + s0 * 32 + lane;
                for (int y = 0; y < rows; ++y) {
                    for (int k = 0; k < work; ++k) a0 = fmaf(a0, 1.0001f, 0.5f);
                    for (int p = 0; p < 3; ++p)
                        for (int q = 0; q < ns; ++q) o11[p * plane + q * 32] = a0 + (float)q;
                    o11 += T;
                }
            }
            continue;
        }
        if (MODE == 7) {
            // reference: plain contiguous fill, one item = the bytes of one (crop, strip) item written as one contiguous run
            const size_t per = (size_t)3 * rows * 32;                               // floats per item
            float4* q = reinterpret_cast<float4*>(out + (size_t)item * per);
            const float4 v = make_float4(a0, a0 + 1.f, a0 + 2.f, a0 + 3.f);
            for (int e = lane; e < (int)(per / 4); e += 32) q[e] = v;
            continue;
        }
        const int roi = MODE == 8 ? item % R : item / nstrip, strip = MODE == 8 ? item / R : item - roi * nstrip;
        float* o = out + ((size_t)roi * 3 * T + y0) * T + strip * 32 + lane;
        float a = (float)lane, b = 1.0001f;
        int buf = 0;
        for (int y = 0; y < rows; ++y) {
            for (int k = 0; k < work; ++k) a = fmaf(a, b, 0.5f);
            if (MODE == 0 || MODE == 8) {
                o[0] = a; o[plane] = a + 1.f; o[2 * plane] = a + 2.f;
                o += T;
            } else if (MODE == 5) {
                if (a == 123.456f) o[0] = a;          // no stores: compute only
            } else if (MODE == 3) {
                // what-if: the same bytes with STG.128 -- lane l writes 16 B of row (y & ~3) + l / 8 (values are arbitrary here)
                if ((y & 3) == 3) {
                    float4* q = reinterpret_cast<float4*>(out + ((size_t)roi * 3 * T + y0 + y - 3 + (lane >> 3)) * T + strip * 32) + (lane & 7);
                    const float4 v = make_float4(a, a + 1.f, a + 2.f, a + 3.f);
                    q[0] = v; q[plane / 4] = v; q[plane / 2] = v;
                }
            } else if (MODE == 4) {
                // what-if: STG.64 -- lane l writes 8 B of row (y & ~1) + l / 16
                if ((y & 1) == 1) {
                    float2* q = reinterpret_cast<float2*>(out + ((size_t)roi * 3 * T + y0 + y - 1 + (lane >> 4)) * T + strip * 32) + (lane & 15);
                    const float2 v = make_float2(a, a + 1.f);
                    q[0] = v; q[plane / 2] = v; q[plane] = v;
                }
            } else {
                float* s = reinterpret_cast<float*>(wb + buf * (WSMEM / 2)) + (y & 3) * 32 + lane;
                s[0] = a; s[4 * 32] = a + 1.f; s[8 * 32] = a + 2.f;
                if ((y & 3) == 3) {
                    if (MODE == 1) {
                        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                        __syncwarp();
                        if (lane == 0) {
                            asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%1, %2, %3, %4}], [%5];"
                                         :: "l"(&map), "r"(strip * 32), "r"(y0 + y - 3), "r"(0), "r"(roi), "r"(wb_s + buf * (WSMEM / 2)) : "memory");
                            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                            asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                        }
                        __syncwarp();
                    } else {
                        __syncwarp();
                        // 12 (plane, row) lines of 128 B = 96 float4: three per lane
                        const float4* sb = reinterpret_cast<const float4*>(wb + buf * (WSMEM / 2));
                        for (int q = lane; q < 96; q += 32) {
                            const int line = q >> 3, p = line >> 2, r = line & 3;
                            float4* dst = reinterpret_cast<float4*>(out + ((size_t)(roi * 3 + p) * T + y0 + y - 3 + r) * T + strip * 32) + (q & 7);
                            *dst = sb[q];
                        }
                        __syncwarp();
                    }
                    buf ^= 1;
                }
            }
        }
        if (MODE == 1) {
            if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            __syncwarp();
        }
    }
    if (MODE == 1 && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        unsigned long long g1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g1));
        clk[0] = clock64() - c0; clk[1] = g1 - g0;
    }
}

extern "C" int probe_store(float* out, int R, int T, int rows, int work, int mode, int* counter, void* stream, int ctas_per_sm, unsigned long long* clk) {
    static EncodeFn encode = nullptr;
    if (!encode) {
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&encode, cudaEnableDefault, &q) != cudaSuccess || !encode) return -100;
    }
    CUtensorMap map;
    const cuuint64_t gdim[4] = {(cuuint64_t)T, (cuuint64_t)T, 3, (cuuint64_t)R};
    const cuuint64_t gstride[3] = {(cuuint64_t)T * 4, (cuuint64_t)T * T * 4, (cuuint64_t)3 * T * T * 4};
    const cuuint32_t box[4] = {32, 4, 3, 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = encode(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void*)out, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return -200 - (int)r;
    cudaMemsetAsync(counter, 0, 4, (cudaStream_t)stream);
    const int smem = ((227 * 1024 / ctas_per_sm - 1024) / 128) * 128;          // pads shared memory to pin the number of resident CTAs
    void (*fn)(float*, int, int, int, int, int*, const CUtensorMap, unsigned long long*) = mode == 0 ? probe_kernel<0> : (mode == 1 ? probe_kernel<1> : (mode == 2 ? probe_kernel<2> : (mode == 3 ? probe_kernel<3> : (mode == 4 ? probe_kernel<4> : (mode == 5 ? probe_kernel<5> : (mode == 6 ? probe_kernel<6> : (mode == 7 ? probe_kernel<7> : (mode == 8 ? probe_kernel<8> : (mode == 9 ? probe_kernel<9> : (mode == 10 ? probe_kernel<10> : probe_kernel<11>))))))))));
    cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    fn<<<148 * ctas_per_sm, 256, smem, (cudaStream_t)stream>>>(out, R, T, rows, work, counter, map, clk);
    return (int)cudaGetLastError();
}
