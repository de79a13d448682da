// Stand-alone check of the 2-D tensor-map copy used by the crop kernels: a [B*H rows][W*3/4 u32] view of the image
// pool, boxes of {pitch/4, 4} elements, several boxes per mbarrier, zero fill beyond the last row.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__global__ void probe_kernel(const __grid_constant__ CUtensorMap map, int x_words, int row, int pitch, int nboxes, uint8_t* out) {
    extern __shared__ __align__(1024) unsigned char smem[];
    const unsigned base = (unsigned)__cvta_generic_to_shared(smem);
    const unsigned bar = base + 16384;
    const int lane = threadIdx.x;
    if (lane == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    if (lane == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(nboxes * 4 * pitch) : "memory");
    __syncwarp();
    if (lane < nboxes) {
        const unsigned dst = base + lane * 4 * pitch;
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                     ::"r"(dst), "l"(&map), "r"(x_words), "r"(row + 4 * lane), "r"(bar) : "memory");
    }
    asm volatile("{\n.reg .pred p;\nW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra D;\nbra W;\nD:\n}" ::"r"(bar) : "memory");
    __syncwarp();
    for (int i = lane; i < nboxes * 4 * pitch; i += 32) out[i] = smem[i];
}

extern "C" int probe_run(const uint8_t* images, int B, int H, int W, int x_bytes, int row, int pitch, int nboxes, uint8_t* out, void* stream) {
    EncodeFn encode = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&encode, cudaEnableDefault, &q);
    if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || !encode) return -100;
    CUtensorMap map;
    const cuuint64_t gdim[2] = {(cuuint64_t)W * 3 / 4, (cuuint64_t)B * H};
    const cuuint64_t gstride[1] = {(cuuint64_t)W * 3};
    const cuuint32_t box[2] = {(cuuint32_t)pitch / 4, 4};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = encode(&map, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, (void*)images, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return -200 - (int)r;
    cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384 + 64);
    probe_kernel<<<1, 32, 16384 + 64, (cudaStream_t)stream>>>(map, x_bytes / 4, row, pitch, nboxes, out);
    return (int)cudaGetLastError();
}
