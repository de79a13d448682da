// Receiver side of the crop gather: uint8 letterboxed crops (what bpc_roi_crop_u8 writes, 147 KB at
// T=224) from one or several source buffers -> the float32 [n][3][T][T] SimplePoseNet input.
//
// This is the tail of process_pose.py:206-209 (cv2.COLOR_BGR2RGB, to_tensor, normalize) applied after
// the crops have crossed NVLink as bytes: 4x fewer bytes on the wire than the float32 planes.  The
// source pointers may be PEER memory (another GPU's buffer mapped into this process): the kernel then
// pulls the bytes over NVLink itself while it converts, so transfer and arithmetic overlap tile by tile
// and no staging copy of the gathered bytes is ever written to the receiver's HBM.
//
// Bound: the receiver's HBM write stream (602 KB per crop at T=224) when the sources are local,
// the NVLink ingest (147 KB per crop) when they are remote.
#include "common.cuh"

namespace bpc {

constexpr int GATHER_MAX_SRC = 16;
constexpr int TILE_PIX = 512;                 // pixels per warp tile
constexpr int TILE_BYTES = TILE_PIX * 3;      // 1536 = 3 x (32 lanes x 16 B)
constexpr int GATHER_WARPS = 8;

struct GatherSources {
    const uint8_t* ptr[GATHER_MAX_SRC];
    int first[GATHER_MAX_SRC + 1];            // first[s] = output slot of source s's crop 0; first[n_src] = total
};

__device__ __forceinline__ uint4 ld_stream16(const uint8_t* p) {
    uint4 v;
    asm volatile("ld.global.cs.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ void st_stream16(float* p, float4 v) {
    asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

struct TileRef {
    const uint8_t* src;     // first byte of the tile
    float* dst;             // plane 0, first pixel of the tile
    int bytes;              // valid bytes in the tile (<= TILE_BYTES)
};

// `per_src` > 0: every source holds per_src crops and the crops are WALKED round-robin over the sources (crop k of source 0, of
// source 1, ...), so that tiles of the local source (bound by this GPU's HBM writes) and of the remote ones (bound by the NVLink
// ingest) are in flight together instead of one phase after the other; the output position of a crop does not change.
__device__ __forceinline__ TileRef tile_ref(const GatherSources& g, int n_src, int per_src, long long item, int tiles, int P, float* out) {
    int crop = (int)(item / tiles);
    const int tile = (int)(item - (long long)crop * tiles);
    int s = 0;
    if (per_src > 0) {
        const int k = crop / n_src;
        s = crop - k * n_src;
        crop = g.first[s] + k;
    } else {
#pragma unroll 1
        while (s + 1 < n_src && crop >= g.first[s + 1]) ++s;
    }
    TileRef r;
    r.src = g.ptr[s] + ((size_t)(crop - g.first[s]) * P + (size_t)tile * TILE_PIX) * 3;
    r.dst = out + (size_t)crop * 3 * P + (size_t)tile * TILE_PIX;
    const int left = (P - tile * TILE_PIX) * 3;
    r.bytes = left < TILE_BYTES ? left : TILE_BYTES;
    return r;
}

// One warp per 512-pixel tile: 3 fully coalesced 16-byte loads per lane into a per-warp shared slice
// (the only way to turn interleaved BGR bytes into per-plane float4 runs without partial-line
// stores), then each lane converts 4 x 4 pixels through the 3x256 LUT and writes 3 planes x 4
// float4 -- every store instruction is 512 contiguous bytes.  The next tile's loads are issued before
// the current tile is converted, so a warp always has 1.5 KB in flight (what hides NVLink latency).
template <bool SWAP>
__global__ void __launch_bounds__(GATHER_WARPS * 32)
bpc_crops_normalise_kernel(GatherSources g, int n_src, int per_src, int T, const float* __restrict__ lut_g, float* __restrict__ out) {
    __shared__ float lut[768];
    __shared__ __align__(16) uint8_t stage[GATHER_WARPS][TILE_BYTES];
    for (int e = threadIdx.x; e < 768; e += GATHER_WARPS * 32) lut[e] = lut_g[e];
    __syncthreads();
    const int lane = lane_id(), warp = warp_id();
    const int P = T * T;
    const int tiles = (P + TILE_PIX - 1) / TILE_PIX;
    const long long nitems = (long long)g.first[n_src] * tiles;
    const long long stride = (long long)gridDim.x * GATHER_WARPS;
    long long item = (long long)blockIdx.x * GATHER_WARPS + warp;
    if (item >= nitems) return;
    uint8_t* mine = stage[warp];

    TileRef cur = tile_ref(g, n_src, per_src, item, tiles, P, out);
    uint4 v[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const int off = (k * 32 + lane) * 16;
        v[k] = off < cur.bytes ? ld_stream16(cur.src + off) : make_uint4(0, 0, 0, 0);
    }
    while (true) {
#pragma unroll
        for (int k = 0; k < 3; ++k) *reinterpret_cast<uint4*>(mine + (k * 32 + lane) * 16) = v[k];
        __syncwarp();
        const long long next = item + stride;
        const bool more = next < nitems;
        TileRef nxt = cur;
        if (more) {
            nxt = tile_ref(g, n_src, per_src, next, tiles, P, out);
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const int off = (k * 32 + lane) * 16;
                v[k] = off < nxt.bytes ? ld_stream16(nxt.src + off) : make_uint4(0, 0, 0, 0);
            }
        }
        const int npix = cur.bytes / 3;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int q = k * 128 + lane * 4;                 // 4 consecutive pixels = 12 bytes = 3 words
            if (q < npix) {
                const uint32_t* w = reinterpret_cast<const uint32_t*>(mine + q * 3);
                const uint32_t w0 = w[0], w1 = w[1], w2 = w[2];
                // byte 3j + c of the 12: channel c of pixel j
                const uint32_t c0[4] = {w0 & 255u, (w0 >> 24), (w1 >> 16) & 255u, (w2 >> 8) & 255u};
                const uint32_t c1[4] = {(w0 >> 8) & 255u, w1 & 255u, (w1 >> 24), (w2 >> 16) & 255u};
                const uint32_t c2[4] = {(w0 >> 16) & 255u, (w1 >> 8) & 255u, w2 & 255u, (w2 >> 24)};
                const uint32_t* a = SWAP ? c2 : c0;
                const uint32_t* b = SWAP ? c0 : c2;
                st_stream16(cur.dst + q, make_float4(lut[a[0]], lut[a[1]], lut[a[2]], lut[a[3]]));
                st_stream16(cur.dst + P + q, make_float4(lut[256 + c1[0]], lut[256 + c1[1]], lut[256 + c1[2]], lut[256 + c1[3]]));
                st_stream16(cur.dst + 2 * (size_t)P + q, make_float4(lut[512 + b[0]], lut[512 + b[1]], lut[512 + b[2]], lut[512 + b[3]]));
            }
        }
        if (!more) break;
        __syncwarp();
        item = next;
        cur = nxt;
    }
}

// Any T (no 16-byte structure to rely on): one thread per output element.
__global__ void bpc_crops_normalise_any_kernel(GatherSources g, int n_src, int T, int swap_rb, const float* __restrict__ lut_g,
                                               float* __restrict__ out) {
    __shared__ float lut[768];
    for (int e = threadIdx.x; e < 768; e += blockDim.x) lut[e] = lut_g[e];
    __syncthreads();
    const int P = T * T;
    const long long total = (long long)g.first[n_src] * 3 * P;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const int crop = (int)(e / (3 * P));
        const int rem = (int)(e - (long long)crop * 3 * P);
        const int plane = rem / P, pix = rem - plane * P;
        int s = 0;
        while (s + 1 < n_src && crop >= g.first[s + 1]) ++s;
        const uint8_t b = g.ptr[s][((size_t)(crop - g.first[s]) * P + pix) * 3 + (swap_rb ? 2 - plane : plane)];
        out[e] = lut[plane * 256 + b];
    }
}

}  // namespace bpc

using namespace bpc;

extern "C" int bpc_crops_normalise(const uint8_t* const* srcs, const int32_t* counts, int n_src, int T, int swap_rb,
                                   const float* lut, float* out, void* stream) {
    if (n_src < 0 || n_src > GATHER_MAX_SRC || T < 1 || T > 256) return BPC_EINVAL;
    if (n_src == 0) return BPC_OK;
    if (!srcs || !counts || !lut || !out) return BPC_EINVAL;
    GatherSources g;
    long long total = 0;
    for (int s = 0; s < n_src; ++s) {
        if (counts[s] < 0 || (counts[s] > 0 && !srcs[s])) return BPC_EINVAL;
        g.ptr[s] = srcs[s];
        g.first[s] = (int)total;
        total += counts[s];
    }
    if (total > (1LL << 24)) return BPC_ETOOBIG;          // item index and slot arithmetic stay far inside 64 / 31 bits
    for (int s = n_src; s <= GATHER_MAX_SRC; ++s) g.first[s] = (int)total;
    for (int s = n_src; s < GATHER_MAX_SRC; ++s) g.ptr[s] = nullptr;
    if (total == 0) return BPC_OK;
    cudaStream_t st = (cudaStream_t)stream;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    bool fast = (T % 4 == 0) && (((uintptr_t)out & 15) == 0);
    for (int s = 0; s < n_src; ++s) fast = fast && (((uintptr_t)srcs[s] & 15) == 0);
    if (fast) {
        const long long tiles = ((long long)T * T + TILE_PIX - 1) / TILE_PIX;
        const long long want = (total * tiles + GATHER_WARPS - 1) / GATHER_WARPS;
        auto fn = swap_rb ? bpc_crops_normalise_kernel<true> : bpc_crops_normalise_kernel<false>;
        int per_sm = 4;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, GATHER_WARPS * 32, 0) != cudaSuccess || per_sm < 1) per_sm = 1;
        const long long slots = (long long)sms * per_sm;
        const int grid = (int)(want < slots ? want : slots);
        int per_src = n_src > 1 ? counts[0] : 0;                 // equal counts: walk the sources round-robin
        for (int s2 = 1; s2 < n_src; ++s2)
            if (counts[s2] != counts[0]) per_src = 0;
        fn<<<grid, GATHER_WARPS * 32, 0, st>>>(g, n_src, per_src, T, lut, out);
    } else {
        bpc_crops_normalise_any_kernel<<<sms * 8, 256, 0, st>>>(g, n_src, T, swap_rb, lut, out);
    }
    BPC_LAUNCH_CHECK();
    return BPC_OK;
}
