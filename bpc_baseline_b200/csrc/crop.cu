// ROI crop -> aspect-preserving letterbox (cv2.resize INTER_AREA) -> colour order -> normalise.
//
// Replaces letterbox_preserving_aspect_ratio (bpc/utils/data_utils.py:34-44) and the inline transform
// of PoseEstimator._estimate_rotation (bpc/inference/process_pose.py:199-209).  The uint8 resize is
// bit-exact with OpenCV's INTER_AREA for 8UC3:
//   regime 1  both axes shrink           : float32 area taps, sequential mul/add (no FMA), cvRound
//   regime 2  both shrink, integer ratio : integer box sum; 2x2 -> (sum+2)>>2
//   regime 3  either axis grows          : 11-bit fixed-point bilinear with area-mode coordinates
//
// Three kernels per call:
//   bpc_crop_prep     one CTA per ROI: letterbox geometry (float64, as Python / OpenCV compute it), regime
//                     and class of the ROI, and -- for the fast classes -- the per-column and per-row tap
//                     descriptors (start index + three float32 weights, absent taps = +0.0f), so that no
//                     float64 arithmetic is left in the hot kernel;
//   bpc_crop_warp     persistent CTAs; the unit of work is (ROI, strip of 32 output columns), taken by ONE
//                     WARP from a global atomic counter, plus one "padding" item per ROI.  A warp stages the
//                     ~130-byte source segments its strip needs into its own slice of shared memory with TMA
//                     (2-D tensor-map boxes of four rows when the image pitch is a multiple of 16 bytes, else one
//                     1-D bulk copy per row; completion on per-warp mbarriers) and streams down its strip, one
//                     lane per column: the horizontal pass of a source row is computed once and reused by the
//                     output rows that tap it.  No CTA barrier, no idle warps behind a narrow letterbox.
//                     Classes: 1 = regime 1 with scale < 2 on both axes (<= 3 taps per axis, double-buffered
//                     batches of output rows) and scale exactly 1; 3 = regime 3 (same batches); 4 = regime 1
//                     with 4..6 taps per axis (2 <= scale <= 5; source rows stream through a four-slot ring);
//   bpc_crop_generic  persistent CTAs over the (rare) remaining ROIs (class 2): integer ratios >= 2, scale > 5,
//                     source rows streamed through shared memory in chunks (boxes as large as the image).
// The dominant traffic is the float32 output (3*T*T*4 B per ROI): each warp stores 128 contiguous bytes per
// plane and row; padding rows are written with 16-byte stores.
#include <cuda.h>

#include <stdlib.h>
#include <string.h>

#include <mutex>
#include <type_traits>

#include "common.cuh"

namespace bpc {

constexpr int CROP_BAND = 8;                 // generic kernel: output rows per work item
constexpr int CROP_RAW_BYTES = 40 * 1024;    // generic kernel: staged source bytes
constexpr int WARP_BUF = 3584;               // warp kernel: bytes of one staging buffer (two per warp)
constexpr int WARP_DESC = 32 * 16;           // warp kernel: 32 row descriptors (two rings per warp)
constexpr int WARP_SMEM = 8320;              // 2 staging buffers + 2 descriptor rings + eight mbarriers, padded to a multiple of 128
static_assert(WARP_SMEM % 128 == 0 && WARP_SMEM >= 2 * WARP_BUF + 2 * WARP_DESC + 64 && WARP_BUF % 128 == 0, "TMA box destinations are 128-byte aligned");
constexpr int WARPK_WARPS = 8;
constexpr int LUT_STRIDE = 257;              // shared-memory LUT: 256 entries + the normalised fill value per output channel
constexpr int LUT_SMEM = 3200;               // 3 * 257 floats, padded to a multiple of 128 (TMA destinations follow)
constexpr int WARPK_SMEM = WARPK_WARPS * WARP_SMEM + LUT_SMEM;
constexpr int STREAM_ROWS = 8;               // streaming class-1 path: source rows per ring slot (= one TMA box)
constexpr int STREAM_MAX_PITCH = 288;        // widest staging pitch of a class-1 strip (32 columns at scale < 2 need <= 256)
constexpr int YSRC_PAD = 16;                 // per-source-row table: zero entries behind the last row (whole slots run to completion)
static_assert(LUT_SMEM % 128 == 0 && LUT_SMEM >= 3 * LUT_STRIDE * 4, "LUT block");
// descriptors per ROI: x axis ds float4, y axis ds + 8 float4 (class 1 reads the y block as 2*ds + 16 float2 source-row records),
// ds = T rounded up to 32
__host__ __device__ __forceinline__ int desc_stride(int T) { return (T + 31) & ~31; }
// float4 records per ROI on the y axis: ds + 8 (class 1 reads the block as 2 ds + 16 float2 source-row records); up to T = 256 the
// block also holds the source-row records of a class-4 crop for the CTA kernel (h <= 5 T rows + one padded slot + 16)
__host__ __device__ __forceinline__ int ydesc_stride(int T) {
    const int base = desc_stride(T) + 8, rows4 = (5 * T + 32 + 1) / 2;
    return (T <= 256 && rows4 > base) ? rows4 : base;
}

struct RoiGeom {                             // 88 bytes, workspace
    double scale_x, scale_y, inv_x, inv_y;
    unsigned long long src;                  // byte address of (y1, x1) in its image
    int w, h, new_w, new_h, dx, dy;
    int regime;                              // 0 rejected, 1 / 2 / 3
    int cls;                                 // -1 beyond the device count, 0 rejected (all fill), 1 fast area, 3 fast linear, 2 generic
    int isx, isy;
    int pitch, pad_;
};
static_assert(sizeof(RoiGeom) == 88, "RoiGeom layout");

// 2-D tensor maps over the image pool seen as [B*H rows][W*3/4 uint32] (only when W*3 is a multiple of 16): one map
// per staging pitch, box = {pitch / 4 words, 4 rows} (2 rows for the two widest), so one TMA instruction stages four
// source rows of a strip instead of one bulk copy per row; rows / columns beyond the pool are zero-filled by the TMA
// unit, which removes the guarded tail path.
#ifndef BPC_L2_PROMO
#define BPC_L2_PROMO CU_TENSOR_MAP_L2_PROMOTION_L2_128B
#endif
constexpr int N_TMAPS = 15;
constexpr int N_TMAPS8 = 8;                  // 8-row boxes for the streaming class-1 path: pitches 64 .. 288
struct TmapSet { CUtensorMap m[N_TMAPS]; CUtensorMap m8[N_TMAPS8]; };
__host__ __device__ __forceinline__ int tmap_pitch(int need) { return need <= 448 ? (need < 64 ? 64 : ((need + 31) & ~31)) : ((need + 63) & ~63); }
__host__ __device__ __forceinline__ int tmap_index(int pitch) { return pitch <= 448 ? (pitch - 64) / 32 : 13 + (pitch - 512) / 64; }
__host__ __device__ __forceinline__ int tmap_rows(int pitch) { return pitch <= 448 ? 4 : 2; }
constexpr int TMAP_MAX_PITCH = 576;

struct YDesc {                               // generic kernel, per output row of the band
    int start;
    int n;                                   // taps (regime 1/2) ; regime 3: second source row
    float bf, bm, bl;                        // regime 1 weights ; regime 3: b0, b1 as int bits
    int flags;                               // bit0 has_first, bit1 has_last
};

__device__ __forceinline__ uint4 ld_nc_v4(const void* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// computeResizeAreaTab for one destination index d (OpenCV resize.cpp), all in float64.
__device__ __forceinline__ void area_taps(int d, double scale, int ssize, int& start, int& n, float& af, float& am, float& al, int& flags) {
    const double fsx1 = dmul((double)d, scale);
    const double fsx2 = dadd(fsx1, scale);
    const double cell = fmin(scale, dsub((double)ssize, fsx1));
    int sx1 = (int)ceil(fsx1), sx2 = (int)floor(fsx2);
    sx2 = min(sx2, ssize - 1);
    sx1 = min(sx1, sx2);
    flags = 0;
    af = 0.f; al = 0.f;
    am = __double2float_rn(ddiv(1.0, cell));
    start = sx1;
    n = sx2 - sx1;
    if (dsub((double)sx1, fsx1) > 1e-3) {
        flags |= 1;
        af = __double2float_rn(ddiv(dsub((double)sx1, fsx1), cell));
        start = sx1 - 1;
        ++n;
    }
    if (dsub(fsx2, (double)sx2) > 1e-3) {
        flags |= 2;
        al = __double2float_rn(ddiv(fmin(fmin(dsub(fsx2, (double)sx2), 1.0), cell), cell));
        ++n;
    }
}

// index of the last source tap of destination index d: start + n - 1 of area_taps() = sx2 - 1 + (the last-tap test), without
// the weights (no divisions)
__device__ __forceinline__ int area_last_tap(int d, double scale, int ssize) {
    const double fsx2 = dadd(dmul((double)d, scale), scale);
    const int sx2 = min((int)floor(fsx2), ssize - 1);
    return sx2 - 1 + (dsub(fsx2, (double)sx2) > 1e-3 ? 1 : 0);
}

// the same taps as a start index and three weights (absent taps = +0.0f); valid when n <= 3 (scale < 2)
__device__ __forceinline__ void area_taps3(int d, double scale, int ssize, int& start, int& n, float& w0, float& w1, float& w2) {
    float af, am, al;
    int flags;
    area_taps(d, scale, ssize, start, n, af, am, al, flags);
    float w[3] = {0.f, 0.f, 0.f};
    int k = 0;
    if (flags & 1) w[k++] = af;
    const int m = n - (flags & 1) - ((flags >> 1) & 1);
    for (int t = 0; t < m && k < 3; ++t) w[k++] = am;
    if ((flags & 2) && k < 3) w[k++] = al;
    w0 = w[0]; w1 = w[1]; w2 = w[2];
}

// area-mode coordinates of the generic linear resize for one destination index d.
__device__ __forceinline__ void linear_coef(int d, double scale, double inv, int ssize, int& s0, int& w0, int& w1, int& edge) {
    int s = (int)floor(dmul((double)d, scale));
    float f = __double2float_rn(dsub((double)(d + 1), dmul((double)(s + 1), inv)));
    f = (f <= 0.f) ? 0.f : __fsub_rn(f, floorf(f));
    if (s < 0) { f = 0.f; s = 0; }
    edge = 0;
    if (s + 1 >= ssize) {
        edge = 1;
        if (s >= ssize - 1) { f = 0.f; s = ssize - 1; }
    }
    s0 = s;
    w0 = __float2int_rn(__fmul_rn(__fsub_rn(1.f, f), 2048.f));
    w1 = __float2int_rn(__fmul_rn(f, 2048.f));
}


// Per-SOURCE-row records of an area resize with at most `maxtaps` taps per output row (collective over the 256 threads of a prep
// CTA): record s = (ba, bb); |ba| = weight of source row s in the output row being accumulated, sign bit of ba set = that output
// row is complete after row s; bb = weight of row s as the first tap of the next output row (+0 if it has none there).  Rows
// [h, hpad) stay zero (no contribution, no completed row) so that whole ring slots run to completion.  *bad is set when the taps
// do not have this shape (then the ROI takes the generic path).
__device__ void src_row_records(float2* ysrc, int cap, const RoiGeom& g, int maxtaps, int tid, int* bad) {
    const int hpad = min(cap, ((g.h + STREAM_ROWS - 1) / STREAM_ROWS) * STREAM_ROWS + STREAM_ROWS);
    if (g.h + STREAM_ROWS > cap) *bad = 1;
    for (int s = tid; s < hpad; s += 256) ysrc[s] = make_float2(0.f, 0.f);
    __syncthreads();
    for (int d = tid; d < g.new_h && !*bad; d += 256) {
        int ys, yn, fl; float bf, bm, bl;
        area_taps(d, g.scale_y, g.h, ys, yn, bf, bm, bl, fl);
        const int prev_last = d > 0 ? area_last_tap(d - 1, g.scale_y, g.h) : -1;
        if (yn < 1 || yn > maxtaps || ys < prev_last || ys + yn > g.h) { *bad = 1; break; }
        for (int t = 0; t < yn; ++t) {
            const int s = ys + t;
            const bool shared_first = (t == 0 && s == prev_last);       // row s also closes output row d - 1
            if (shared_first && yn == 1) { *bad = 1; break; }           // one source row closing two output rows: generic path
            float wv = (t == 0 && (fl & 1)) ? bf : ((t == yn - 1 && (fl & 2)) ? bl : bm);
            if (t == yn - 1) wv = __int_as_float(__float_as_int(wv) | (int)0x80000000u);
            if (shared_first) ysrc[s].y = wv; else ysrc[s].x = wv;
        }
    }
    __syncthreads();
}

// ------------------------------------------------------------------------------------------------------
// prep: geometry, classification and tap descriptors, one CTA per ROI
// ------------------------------------------------------------------------------------------------------
// Descriptor formats (float4; desc_stride(T) per ROI on the x axis, desc_stride(T) + 8 on the y axis):
//   class 1 (area, <= 3 taps)  x: (w0, w1, w2, bits(first source column))
//                              y: (b0, b1, b2, bits(first source row | taps << 24))          [image pitch not a multiple of 16]
//                              y: float2 per SOURCE row s (the streaming path): (ba, bb).  |ba| = weight of row s in the
//                                 output row being accumulated; sign bit of ba set = that output row is complete after row s;
//                                 bb = weight of row s as the first tap of the next output row (+0 if it has none there).
//                                 An output row then is acc = (((0 + b0 h0) + b1 h1) + b2 h2), bit-identical to OpenCV's
//                                 sum = b0 h0; sum += b1 h1; ... for non-negative terms.
//   class 3 (fixed-point)      x: (bits(w0), bits(w1), 0, bits(source column))      w = 2048, 0 beyond xmax
//                              y: (bits(b0), bits(b1), bits(second source row), bits(first source row))
__global__ void __launch_bounds__(256)
bpc_crop_prep_kernel(const uint8_t* __restrict__ images, int B, int H, int W, const int32_t* __restrict__ rois, int R,
                     const int32_t* __restrict__ n_rois_dev, int roi_first, int T, int stream_ok,
                     RoiGeom* __restrict__ geom, float4* __restrict__ xdesc, float4* __restrict__ ydesc,
                     int32_t* __restrict__ glist, int32_t* __restrict__ gcount, int32_t* __restrict__ status, int32_t* __restrict__ list1) {
    __shared__ RoiGeom g;
    __shared__ int s_bad;
    const int roi = blockIdx.x, tid = threadIdx.x;
    const int ds = desc_stride(T);
    if (tid == 0) {
        s_bad = 0;
        g.scale_x = g.scale_y = g.inv_x = g.inv_y = 0.0;
        g.src = 0ull;
        g.new_w = g.new_h = g.dx = g.dy = 0;
        g.regime = 0; g.cls = 0; g.isx = g.isy = 0; g.pitch = 16; g.pad_ = 0;
        const int32_t* r = rois + (size_t)roi * 5;
        const int img = r[0], x1 = r[1], y1 = r[2], x2 = r[3], y2 = r[4];
        const int w = x2 - x1, h = y2 - y1;
        g.w = w; g.h = h;
        if (n_rois_dev != nullptr && roi_first + roi >= *n_rois_dev) {
            g.cls = -1;
        } else {
            if (img >= 0 && img < B && x1 >= 0 && y1 >= 0 && x2 <= W && y2 <= H && w > 0 && h > 0 && w <= BPC_MAX_ROI_WIDTH) {
                // letterbox geometry, data_utils.py:35-38,41-42 (Python round = half-to-even on the f64 product)
                const double scale = ddiv((double)T, (double)max(h, w));
                const int new_w = (int)__double2ll_rn(dmul((double)w, scale));
                const int new_h = (int)__double2ll_rn(dmul((double)h, scale));
                if (new_w >= 1 && new_h >= 1 && new_w <= T && new_h <= T) {
                    g.new_w = new_w; g.new_h = new_h;
                    g.dx = (T - new_w) / 2; g.dy = (T - new_h) / 2;
                    g.inv_x = ddiv((double)new_w, (double)w);      // cv2.resize: inv_scale = dsize / ssize
                    g.inv_y = ddiv((double)new_h, (double)h);
                    g.scale_x = ddiv(1.0, g.inv_x);
                    g.scale_y = ddiv(1.0, g.inv_y);
                    if (g.scale_x >= 1.0 && g.scale_y >= 1.0) {
                        g.isx = __double2int_rn(g.scale_x);
                        g.isy = __double2int_rn(g.scale_y);
                        const bool fast = fabs(dsub(g.scale_x, (double)g.isx)) < 2.220446049250313e-16 &&
                                          fabs(dsub(g.scale_y, (double)g.isy)) < 2.220446049250313e-16;
                        g.regime = fast ? 2 : 1;
                    } else {
                        g.regime = 3;
                    }
                    g.src = (unsigned long long)(uintptr_t)images + (((unsigned long long)img * H + y1) * W + x1) * 3ull;
                    // scale exactly 1 (long side == T) is OpenCV's integer-ratio regime, but a 1 x 1 box sum is the byte itself and
                    // so is the area pass with its single tap of weight 1.0f: keep it on the fast path
                    if ((g.regime == 1 && g.scale_x < 2.0 && g.scale_y < 2.0) || (g.regime == 2 && g.isx == 1 && g.isy == 1)) g.cls = 1;
                    else if (g.regime == 3) g.cls = 3;
                    else if (g.regime == 1) {
                        // general tap counts in the warp kernel if a strip's rows fit its staging buffer:
                        // 32 columns span at most 31*scale_x + ceil(scale_x) + 2 source pixels
                        const int seg_px = (int)(31.0 * g.scale_x) + (int)ceil(g.scale_x) + 3;
                        const int pitch_max = ((3 * seg_px + 46) >> 4) << 4;
                        const bool taps6 = (int)ceil(g.scale_x) + 1 <= 6 && (int)ceil(g.scale_y) + 1 <= 6;
                        // per-strip kernel: four ring slots of >= 2 rows fit 2 * WARP_BUF; CTA kernel (full-width rows): any width
                        g.cls = (taps6 && ((stream_ok & 2) || pitch_max <= TMAP_MAX_PITCH)) ? 4 : 2;
                    } else g.cls = 2;
                }
            }
            if (status != nullptr) status[roi] = (g.regime == 0) ? 1 : 0;
        }
    }
    __syncthreads();
    const int cls = g.cls;
    float4* xd = xdesc + (size_t)roi * ds;
    float4* yd = ydesc + (size_t)roi * ydesc_stride(T);
    if (cls == 4) {
        // (w_first, w_middle, w_last, bits(start | taps << 24)); a missing first / last tap takes the middle weight
        for (int axis = 0; axis < ((stream_ok & 2) ? 1 : 2); ++axis) {
            const int nd = axis ? g.new_h : g.new_w;
            for (int d = tid; d < nd; d += 256) {
                int st, n, flags; float af, am, al;
                area_taps(d, axis ? g.scale_y : g.scale_x, axis ? g.h : g.w, st, n, af, am, al, flags);
                const float w0 = (flags & 1) ? af : ((n == 1 && (flags & 2)) ? al : am);
                const float wl = (flags & 2) ? al : am;
                (axis ? yd : xd)[d] = make_float4(w0, am, wl, __int_as_float(st | (n << 24)));
            }
        }
        if (stream_ok & 2) {
            src_row_records(reinterpret_cast<float2*>(yd), 2 * ydesc_stride(T), g, 6, tid, &s_bad);
            if (s_bad && tid == 0) { g.cls = 2; }
        }
    } else if (cls == 1) {
        for (int d = tid; d < g.new_w; d += 256) {
            int xs, xn; float w0, w1, w2;
            area_taps3(d, g.scale_x, g.w, xs, xn, w0, w1, w2);
            xd[d] = make_float4(w0, w1, w2, __int_as_float(xs));
        }
        if (!(stream_ok & 1)) {
            for (int d = tid; d < g.new_h; d += 256) {
                int ys, yn; float b0, b1, b2;
                area_taps3(d, g.scale_y, g.h, ys, yn, b0, b1, b2);
                yd[d] = make_float4(b0, b1, b2, __int_as_float(ys | (yn << 24)));
            }
        } else {
            src_row_records(reinterpret_cast<float2*>(yd), 2 * ds + YSRC_PAD, g, 3, tid, &s_bad);
            if (s_bad && tid == 0) { g.cls = 2; }
        }
    } else if (cls == 3) {
        for (int d = tid; d < g.new_w; d += 256) {
            int xs, w0, w1, edge;
            linear_coef(d, g.scale_x, g.inv_x, g.w, xs, w0, w1, edge);
            if (edge) { w0 = 2048; w1 = 0; }                        // D = S[sx] * ONE beyond xmax
            xd[d] = make_float4(__int_as_float(w0), __int_as_float(w1), 0.f, __int_as_float(xs));
        }
        for (int d = tid; d < g.new_h; d += 256) {
            int s0, b0, b1, edge;
            linear_coef(d, g.scale_y, g.inv_y, g.h, s0, b0, b1, edge);
            const int s1 = min(s0 + 1, g.h - 1);
            yd[d] = make_float4(__int_as_float(b0), __int_as_float(b1), __int_as_float(s1), __int_as_float(s0));
        }
    }
    __syncthreads();
    if (tid == 0) {
        if (g.cls == 2) glist[atomicAdd(gcount, 1)] = roi;
        if (list1 != nullptr) {
            // crops streamed by bpc_crop_cta_kernel from the front of the list, those of bpc_crop_warp_kernel from its back
            if (g.cls == 0 || g.cls == 1 || g.cls == 3 || g.cls == 4) list1[atomicAdd(gcount + 8, 1)] = roi;
            else if (g.cls != 2 && g.cls != -1) list1[R - 1 - atomicAdd(gcount + 9, 1)] = roi;      // (none today)
        }
        geom[roi] = g;
    }
}

// ------------------------------------------------------------------------------------------------------
// shared output helpers
// ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned short bf16_bits(float v) {       // round-to-nearest-even float32 -> bfloat16
    unsigned short r;
    asm("cvt.rn.bf16.f32 %0, %1;" : "=h"(r) : "f"(v));
    return r;
}

// BF16: the float path's values rounded to bfloat16 and stored channels-last ([R][T][T][3], what a bf16 tensor-core network reads)
template <bool OUT_U8, bool BF16 = false>
struct Out {
    float* outf; uint8_t* outb;
    const float* lut;           // shared memory [3][LUT_STRIDE]
    int T, swap_rb;
    uint8_t fillc[3];
    float padf[3];

    __device__ __forceinline__ void px(int roi, int y, int x, int b0, int b1, int b2) const {   // source channel order
        if (OUT_U8) {
            uint8_t* o = outb + (((size_t)roi * T + y) * T + x) * 3;
            o[0] = (uint8_t)b0; o[1] = (uint8_t)b1; o[2] = (uint8_t)b2;
        } else {
            const int s0 = swap_rb ? b2 : b0, s2 = swap_rb ? b0 : b2;
            if (BF16) {
                unsigned short* o = reinterpret_cast<unsigned short*>(outf) + (((size_t)roi * T + y) * T + x) * 3;
                o[0] = bf16_bits(lut[s0]); o[1] = bf16_bits(lut[LUT_STRIDE + b1]); o[2] = bf16_bits(lut[2 * LUT_STRIDE + s2]);
            } else {
                float* o = outf + ((size_t)roi * 3 * T + y) * T + x;
                o[0] = lut[s0];
                o[(size_t)T * T] = lut[LUT_STRIDE + b1];
                o[(size_t)2 * T * T] = lut[2 * LUT_STRIDE + s2];
            }
        }
    }
    __device__ __forceinline__ void pad(int roi, int y, int x) const {
        if (OUT_U8) {
            uint8_t* o = outb + (((size_t)roi * T + y) * T + x) * 3;
            o[0] = fillc[0]; o[1] = fillc[1]; o[2] = fillc[2];
        } else if (BF16) {
            unsigned short* o = reinterpret_cast<unsigned short*>(outf) + (((size_t)roi * T + y) * T + x) * 3;
            o[0] = bf16_bits(padf[0]); o[1] = bf16_bits(padf[1]); o[2] = bf16_bits(padf[2]);
        } else {
            float* o = outf + ((size_t)roi * 3 * T + y) * T + x;
            o[0] = padf[0]; o[(size_t)T * T] = padf[1]; o[(size_t)2 * T * T] = padf[2];
        }
    }
    // rows [ra, rb) of all three planes = fill; collective over nth threads
    __device__ void pad_rows(int roi, int ra, int rb, int tid, int nth) const {
        if (rb <= ra) return;
        if (OUT_U8) {
            uint8_t* o = outb + ((size_t)roi * T + ra) * T * 3;
            const int nbytes = (rb - ra) * T * 3;
            for (int e = tid; e < nbytes; e += nth) o[e] = fillc[e % 3];
        } else if (BF16) {
            unsigned short* o = reinterpret_cast<unsigned short*>(outf) + ((size_t)roi * T + ra) * T * 3;
            const int n = (rb - ra) * T * 3;
            const unsigned short f0 = bf16_bits(padf[0]), f1 = bf16_bits(padf[1]), f2 = bf16_bits(padf[2]);
            if ((((size_t)(uintptr_t)o) & 3) == 0 && (n & 1) == 0) {          // 32-bit stores of the period-3 pattern
                unsigned* o2 = reinterpret_cast<unsigned*>(o);
                const unsigned p0 = f0 | ((unsigned)f1 << 16), p1 = f2 | ((unsigned)f0 << 16), p2 = f1 | ((unsigned)f2 << 16);
                for (int e = tid; e < (n >> 1); e += nth) { const int m = e % 3; o2[e] = m == 0 ? p0 : (m == 1 ? p1 : p2); }
            } else {
                for (int e = tid; e < n; e += nth) { const int m = e % 3; o[e] = m == 0 ? f0 : (m == 1 ? f1 : f2); }
            }
        } else if ((T & 3) == 0) {
            const int n4 = (rb - ra) * (T >> 2);
            for (int p = 0; p < 3; ++p) {
                float4* o = reinterpret_cast<float4*>(outf + ((size_t)(roi * 3 + p) * T + ra) * T);
                const float4 v = make_float4(padf[p], padf[p], padf[p], padf[p]);
                for (int e = tid; e < n4; e += nth) o[e] = v;
            }
        } else {
            const int n1 = (rb - ra) * T;
            for (int p = 0; p < 3; ++p) {
                float* o = outf + ((size_t)(roi * 3 + p) * T + ra) * T;
                for (int e = tid; e < n1; e += nth) o[e] = padf[p];
            }
        }
    }
};

template <bool OUT_U8, bool BF16>
__device__ __forceinline__ void out_init(Out<OUT_U8, BF16>& o, float* outf, uint8_t* outb, const float* lut_s, int T, int swap_rb, uchar4 fill) {
    o.outf = outf; o.outb = outb; o.lut = lut_s; o.T = T; o.swap_rb = swap_rb;
    o.fillc[0] = fill.x; o.fillc[1] = fill.y; o.fillc[2] = fill.z;
    if (!OUT_U8)
        for (int p = 0; p < 3; ++p) o.padf[p] = lut_s[p * LUT_STRIDE + o.fillc[swap_rb ? 2 - p : p]];
}


// ------------------------------------------------------------------------------------------------------
// warp kernel: (ROI, 32-column strip) items, one warp each
// ------------------------------------------------------------------------------------------------------
// ---- packed float32x2 arithmetic (sm_100a FFMA2 / FADD2), every lane IEEE round-to-nearest ----
typedef unsigned long long u64;
__device__ __forceinline__ u64 pack2(float lo, float hi) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void unpack2(u64 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ u64 ffma2(u64 a, u64 b, u64 c) { u64 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ u64 fadd2(u64 a, u64 b) { u64 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
// ptxas 12.9 contracts a mul.rn.f32x2 feeding an add.rn.f32x2 into one FFMA2 (even with --fmad=false and
// volatile asm; the scalar forms are left alone): one rounding where OpenCV rounds twice, seen as 1-LSB
// errors in ~1e-4 of the pixels.  fma(a, b, -0.0) with a literal -0 is simplified back to a multiply and
// contracted again.  The vertical pass therefore forms its rounded products as fma(a, b, z) with
// z = (-0.0f, -0.0f) held in a register whose value the compiler cannot prove: the same value as a * b for
// non-negative operands, and an FFMA2 cannot be merged with the add that follows.
__device__ __forceinline__ u64 fprod2(u64 a, u64 b, u64 negzero2) { return ffma2(a, b, negzero2); }

// float(2^23 + byte K of v): the byte dropped into the mantissa of 8388608.0f (one PRMT, no I2F)
template <int K>
__device__ __forceinline__ float magic_byte(unsigned v) { return __uint_as_float(__byte_perm(v, 0x4B000000u, 0x7540 + K)); }

// Column weights of the 3-slot horizontal pass.  fl(b * w) is computed as fma(2^23 + b, w, -(2^23 * w)):
// the product 2^23 * w is exact, so the fused result is the correctly rounded b * w, bit-identical to
// OpenCV's separately rounded multiply.
struct ColW {
    float w[3], c[3];       // weight and -(2^23 * weight) per slot
    __device__ __forceinline__ void set(float a, float b, float d) {
        w[0] = a; w[1] = b; w[2] = d;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            c[k] = __fmul_rn(w[k], -8388608.0f);
            asm volatile("" : "+f"(c[k]));      // opaque: keep it in a register instead of re-multiplying in the inner loop
        }
    }
};

// shared-memory accesses through 32-bit shared-window addresses (no generic-address arithmetic in the hot loop)
__device__ __forceinline__ unsigned lds_u32(unsigned addr) { unsigned v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr)); return v; }
__device__ __forceinline__ unsigned lds_u8(unsigned addr) { unsigned v; asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(addr)); return v; }
__device__ __forceinline__ float lds_f32(unsigned addr) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr)); return v; }
__device__ __forceinline__ float4 lds_f4(unsigned addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}

// output store of the fast paths: the crop buffer is written once and never re-read by this kernel
#if defined(BPC_WHATIF)
__device__ float* g_whatif_base;
#endif
template <int K = 0>
__device__ __forceinline__ void stg_out(float* p, float v) {
#if defined(BPC_WHATIF) && BPC_WHATIF == 1          // only plane 0 is stored
    if (K == 0) *p = v; else asm volatile("" :: "f"(v));
#elif defined(BPC_WHATIF) && BPC_WHATIF == 2        // LSU traffic without L2 traffic: the value goes to shared memory
    asm volatile("st.shared.f32 [%0], %1;" :: "r"(((unsigned)(size_t)p) & 0x7cu), "f"(v));
#elif defined(BPC_WHATIF) && BPC_WHATIF == 4        // stores fold into a 128 KB window per CTA (57 MB in all: stays in L2)
    float* q = (float*)((size_t)g_whatif_base + ((size_t)blockIdx.x << 17) + (((size_t)p) & 0x1fffc));
    *q = v;
#elif defined(BPC_WHATIF) && BPC_WHATIF == 3        // no store, LUT value still loaded
    asm volatile("" :: "f"(v));
#elif defined(BPC_STG_CS)
    __stcs(p, v);
#else
    *p = v;
#endif
}

// horizontal pass of one source row for one output column: 9 bytes starting at shared address a4 + sh/8
__device__ __forceinline__ void h_area3(unsigned a4, int sh, const ColW& cw, u64& h01, float& h2) {
    const unsigned q0 = lds_u32(a4), q1 = lds_u32(a4 + 4), q2 = lds_u32(a4 + 8);
    const unsigned v0 = __funnelshift_r(q0, q1, sh), v1 = __funnelshift_r(q1, q2, sh), v2 = q2 >> sh;
    // pixel k = bytes 3k .. 3k+2; the (w, w) / (c, c) pairs become scalar-broadcast operands of FFMA2
    const u64 p0 = ffma2(pack2(magic_byte<0>(v0), magic_byte<1>(v0)), pack2(cw.w[0], cw.w[0]), pack2(cw.c[0], cw.c[0]));
    const u64 p1 = ffma2(pack2(magic_byte<3>(v0), magic_byte<0>(v1)), pack2(cw.w[1], cw.w[1]), pack2(cw.c[1], cw.c[1]));
    const u64 p2 = ffma2(pack2(magic_byte<2>(v1), magic_byte<3>(v1)), pack2(cw.w[2], cw.w[2]), pack2(cw.c[2], cw.c[2]));
    const float r0 = __fmaf_rn(magic_byte<2>(v0), cw.w[0], cw.c[0]);
    const float r1 = __fmaf_rn(magic_byte<1>(v1), cw.w[1], cw.c[1]);
    const float r2 = __fmaf_rn(magic_byte<0>(v2), cw.w[2], cw.c[2]);
    h01 = fadd2(fadd2(p0, p1), p2);
    h2 = __fadd_rn(__fadd_rn(r0, r1), r2);
}

// horizontal pass with NT taps evaluated (4..6): 3*NT bytes starting at shared address a4 + sh/8.  Taps beyond a
// lane's own count carry weight +0.0f (c = -0.0f): fma(2^23 + b, +0, -0) = +0 and s + (+0) = s, so padding the
// tap list to the warp's maximum leaves every sum bit-identical to the sequential ((S0*w0 + S1*w1) + ...) order.
template <int NT>
__device__ __forceinline__ void h_area_n(unsigned a4, int sh, const float (&w)[6], const float (&c)[6], u64& h01, float& h2) {
    constexpr int NV = (3 * NT + 3) / 4;            // aligned words holding 3*NT bytes; NV + 1 raw words cover any shift
    unsigned q[NV + 1], v[NV];
#pragma unroll
    for (int i = 0; i <= NV; ++i) q[i] = lds_u32(a4 + 4 * i);
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = __funnelshift_r(q[i], q[i + 1], sh);
#pragma unroll
    for (int k = 0; k < NT; ++k) {
        const int j = 3 * k;
        const float b0 = __uint_as_float(__byte_perm(v[j >> 2], 0x4B000000u, 0x7540 + (j & 3)));
        const float b1 = __uint_as_float(__byte_perm(v[(j + 1) >> 2], 0x4B000000u, 0x7540 + ((j + 1) & 3)));
        const float b2 = __uint_as_float(__byte_perm(v[(j + 2) >> 2], 0x4B000000u, 0x7540 + ((j + 2) & 3)));
        const u64 p = ffma2(pack2(b0, b1), pack2(w[k], w[k]), pack2(c[k], c[k]));
        const float r = __fmaf_rn(b2, w[k], c[k]);
        h01 = (k == 0) ? p : fadd2(h01, p);
        h2 = (k == 0) ? r : __fadd_rn(h2, r);
    }
}

// horizontal pass of the fixed-point bilinear: 6 bytes at shared address a4 + sh/8, result pre-shifted by 4
__device__ __forceinline__ void h_lin(unsigned a4, int sh, int w0, int w1, int* h) {
    const unsigned q0 = lds_u32(a4), q1 = lds_u32(a4 + 4), q2 = lds_u32(a4 + 8);
    const unsigned v0 = __funnelshift_r(q0, q1, sh), v1 = __funnelshift_r(q1, q2, sh);
    h[0] = (int)((v0 & 0xffu) * w0 + (v0 >> 24) * w1) >> 4;
    h[1] = (int)(((v0 >> 8) & 0xffu) * w0 + (v1 & 0xffu) * w1) >> 4;
    h[2] = (int)(((v0 >> 16) & 0xffu) * w0 + ((v1 >> 8) & 0xffu) * w1) >> 4;
}

// cvRound(v) for 0 <= v < 2^22 without F2I: adding 2^23 leaves round-half-even(v) in the low mantissa bits
// (no clamp: the tap weights of each axis sum to 1 within a few ulp, so v <= 255.001 and the result is <= 255)
__device__ __forceinline__ int round_u8(float v) { return __float_as_int(__fadd_rn(v, 8388608.0f)) & 0xff; }
// shared address of LUT[cvRound(v)]: lut_m = lut_base - 4 * 0x4B000000 (mod 2^32)
__device__ __forceinline__ unsigned lut_addr(float v, unsigned lut_m) { return (unsigned)__float_as_int(__fadd_rn(v, 8388608.0f)) * 4u + lut_m; }

// ---- TMA bulk copy (cp.async.bulk, SASS UBLKCP) completing on an mbarrier in shared memory ----
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(unsigned dst, unsigned long long src, unsigned bytes, unsigned bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
#ifndef BPC_MBAR_HINT_NS
#define BPC_MBAR_HINT_NS 2000
#endif
// try_wait with a suspend-time hint: a waiting warp sleeps in hardware until the phase completes (or the hint elapses) instead of
// spinning through try_wait / yield / branch -- spinning consumer and producer warps executed 30 % of the CTA kernel's
// instructions before the hint was added
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
    asm volatile("{\n"
                 ".reg .pred p;\n"
                 "BPC_WAIT:\n"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"
                 "@p bra BPC_DONE;\n"
                 "bra BPC_WAIT;\n"
                 "BPC_DONE:\n"
                 "}" :: "r"(bar), "r"(parity), "r"((unsigned)BPC_MBAR_HINT_NS) : "memory");
}

// Stage rows [s_lo, s_lo + count) of a strip into a warp buffer.  Normal case: one TMA bulk copy of `pitch` bytes
// per row (lane r issues row r), completion counted in bytes on the warp's mbarrier -> returns 1.  If the
// 16-byte-aligned over-read of the last row would leave the image pool (last rows of the last image) the
// rows are copied with guarded 16-byte cp.async instead -> returns 0 (wait with cp.async.wait_group).
__device__ __forceinline__ int warp_stage(unsigned char* buf, unsigned buf_s, unsigned bar_s, unsigned long long src_seg,
                                          unsigned long long rowstride, unsigned long long img_end, int s_lo, int count,
                                          int pitch, int lane, int lr, int lv, int rpp) {
    const unsigned long long first = src_seg + (unsigned long long)s_lo * rowstride;
    const unsigned long long last_end = ((first + (unsigned long long)(count - 1) * rowstride) & ~15ull) + (unsigned long long)pitch;
    if (last_end <= img_end) {
        if (lane == 0) mbar_expect_tx(bar_s, (unsigned)(count * pitch));
        __syncwarp();
        unsigned long long ga = first + (unsigned long long)lane * rowstride;
        unsigned dst = buf_s + lane * pitch;
        for (int r = lane; r < count; r += 32) {
            bulk_g2s(dst, ga & ~15ull, (unsigned)pitch, bar_s);
            ga += 32ull * rowstride;
            dst += 32 * pitch;
        }
        return 1;
    }
    if (lr < rpp) {
        for (int r = lr; r < count; r += rpp) {
            const unsigned long long ga = first + (unsigned long long)r * rowstride;
            const unsigned long long a = (ga & ~15ull) + (unsigned long long)lv * 16ull;
            unsigned char* dst = buf + r * pitch + lv * 16;
            if (a + 16ull <= img_end) {
                cp_async16(dst, (const void*)(uintptr_t)a);
            } else {                                               // last bytes of the image pool
                unsigned int tmp[4] = {0u, 0u, 0u, 0u};
                for (int b = 0; b < 16; ++b)
                    if (a + b < img_end) tmp[b >> 2] |= (unsigned int)(*(const uint8_t*)(uintptr_t)(a + b)) << (8 * (b & 3));
                *reinterpret_cast<uint4*>(dst) = make_uint4(tmp[0], tmp[1], tmp[2], tmp[3]);
            }
        }
    }
    cp_async_commit();
    return 0;
}

// The same through a 2-D tensor map: box i = rows [row + i*rb, +rb) x pitch bytes from word column xw, issued by lane i.
__device__ __forceinline__ void warp_stage_2d(unsigned buf_s, unsigned bar_s, const CUtensorMap* map, int xw, int row, int count,
                                              int pitch, int rb, int lane) {
    const int nboxes = (count + rb - 1) / rb;
    if (lane == 0) mbar_expect_tx(bar_s, (unsigned)(nboxes * rb * pitch));
    __syncwarp();
    if (lane < nboxes)
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                     :: "r"(buf_s + (unsigned)(lane * rb * pitch)), "l"(map), "r"(xw), "r"(row + lane * rb), "r"(bar_s) : "memory");
}

__device__ __forceinline__ int warp_max_i32(int v) {
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, m));
    return v;
}

// TT = compile-time target size (0: run-time T); ALIGNED = the image row pitch W*3 is a multiple of 16 bytes
// SWAP = write the planes in R, G, B order from B, G, R sources (cv2.COLOR_BGR2RGB, process_pose.py:206)
template <bool OUT_U8, int TT, bool ALIGNED, bool SWAP>
__global__ void __launch_bounds__(256, 3)
bpc_crop_warp_kernel(const uint8_t* __restrict__ images, int B, int H, int W, const RoiGeom* __restrict__ geom,
                     const float4* __restrict__ xdesc, const float4* __restrict__ ydesc, int32_t* __restrict__ wcount,
                     int R, int Trt, int nslot, uchar4 fill, int swap_rb, const float* __restrict__ lut_g,
                     float* __restrict__ outf, uint8_t* __restrict__ outb, const __grid_constant__ TmapSet tm, const int32_t* __restrict__ wlist) {
    extern __shared__ __align__(128) unsigned char smem[];
    float* lut = reinterpret_cast<float*>(smem);                                 // [3][LUT_STRIDE]: 256 values + the fill value
    // wlist: the crops of this kernel, stored from the back of the list (the others belong to bpc_crop_cta_kernel); null = all of them
    const int nmine = wlist ? wcount[5] : R;
    if (nmine == 0) return;
    const int T = TT ? TT : Trt;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    unsigned char* wbase = smem + LUT_SMEM + wid * WARP_SMEM;         // [2][WARP_BUF] staging, then [2][32] float4 row descriptors
    const unsigned wbase_s = (unsigned)__cvta_generic_to_shared(wbase), lut_s = (unsigned)__cvta_generic_to_shared(smem);

    // see lut_addr(); passed through shared memory so that it stays ONE register (ptxas otherwise re-derives it as window
    // base + constant with an extra add per use)
    if (tid == 0) *reinterpret_cast<volatile unsigned*>(smem + 3 * LUT_STRIDE * 4) = lut_s - 4u * 0x4B000000u;
    const unsigned bar_s = wbase_s + 2 * WARP_BUF + 2 * WARP_DESC;     // eight mbarriers: one per staging buffer / ring slot
    if (lane < 8) mbar_init(bar_s + 8 * lane, 1);
    unsigned ph = 0;                                                      // bit j: parity of the next completion of barrier j
    if (!OUT_U8)
        for (int e = tid; e < 768; e += (int)blockDim.x) lut[(e >> 8) * LUT_STRIDE + (e & 255)] = lut_g[e];
    __syncthreads();
    Out<OUT_U8> out;
    out_init(out, outf, outb, lut, T, swap_rb, fill);
    if (!OUT_U8 && tid < 3) lut[tid * LUT_STRIDE + 256] = out.padf[tid];  // entry 256 = fill: what a lane beside the image looks up
    __syncthreads();
    const unsigned lut_m = *reinterpret_cast<volatile unsigned*>(smem + 3 * LUT_STRIDE * 4);
    const int ds = desc_stride(T), ystride = ydesc_stride(T);
    constexpr bool swap = SWAP;
    const size_t plane = (size_t)T * T;
    const unsigned long long rowstride = (unsigned long long)W * 3ull;
    const unsigned long long img_end = (unsigned long long)(uintptr_t)images + (unsigned long long)B * H * rowstride;
    const float nz = __int_as_float((int)(0x80000000u | (unsigned)fill.w));     // -0.0f: fill.w is 0 at run time
    const u64 nz2 = pack2(nz, nz);
#if defined(BPC_WHATIF)
    if (tid == 0) g_whatif_base = outf;
    __syncthreads();
#endif
    const long long nitems = (long long)nmine * nslot;

    for (;;) {
        int item = 0;
        if (lane == 0) item = atomicAdd(wcount, 1);
        item = __shfl_sync(0xffffffffu, item, 0);
        if (item >= nitems) break;
        // ROI-major order: a ROI's strips and its (store-only) padding item run at about the same time, which keeps
        // whole output rows together in DRAM and blends store-bound with issue-bound work (padding items last: -8 %)
        const int ri = item / nslot, slot = item - ri * nslot;
        const int roi = wlist ? wlist[R - 1 - ri] : ri;
        const RoiGeom* gp = geom + roi;
        const int cls = gp->cls;
        if (cls == -1 || cls == 2) continue;
        const int new_w = gp->new_w, new_h = gp->new_h, dx0 = gp->dx, dy0 = gp->dy;

        if (slot == nslot - 1) {
            // ---------------- padding item: rows above / below and column strips left / right ----------------
            if (cls == 0) { out.pad_rows(roi, 0, T, lane, 32); continue; }
            out.pad_rows(roi, 0, dy0, lane, 32);
            out.pad_rows(roi, dy0 + new_h, T, lane, 32);
            // 32-column strips that do not touch the resized image (the others pad their own lanes)
            for (int c = 0; c * 32 < T; ++c) {
                if (c * 32 < dx0 + new_w && c * 32 + 32 > dx0) continue;
                const int x = c * 32 + lane;
                if (x < T)
                    for (int r = 0; r < new_h; ++r) out.pad(roi, dy0 + r, x);
            }
            continue;
        }
        if (cls == 0 || slot * 32 >= dx0 + new_w || slot * 32 + 32 <= dx0) continue;

        // ---------------- output columns [32 slot, 32 slot + 32): full 128-byte lines per plane and row ----------------
        const int x = slot * 32 + lane;
        const int xr = x - dx0;
        const bool active = xr >= 0 && xr < new_w;
        const bool padlane = !active && x < T;
        const float4 xd = xdesc[(size_t)roi * ds + min(max(xr, 0), new_w - 1)];
        const int xs = __float_as_int(xd.w) & 0xffffff;
        const int xn = (cls == 4) ? (__float_as_int(xd.w) >> 24) : 3;     // source pixels read from xs on
        const int xs_min = -warp_max_i32(-xs);
        const int xe_max = warp_max_i32(xs + xn);
        const int seg_bytes = 3 * (xe_max - xs_min);
        const int pitch_need = ((15 + seg_bytes + 8 + 15) >> 4) << 4;
        const int pitch = ALIGNED ? tmap_pitch(pitch_need) : pitch_need;
        const int rb = tmap_rows(pitch);                                  // rows per TMA box (ALIGNED)
        const CUtensorMap* map = &tm.m[ALIGNED ? tmap_index(pitch) : 0];
        const int nv = pitch >> 4;
        const int rpp = 32 / nv, lr = lane / nv, lv = lane - lr * nv;
        const unsigned long long src_seg = gp->src + 3ull * (unsigned long long)xs_min;
        const int mis0 = (int)(src_seg & 15ull), misstep = ALIGNED ? 0 : (int)(rowstride & 15ull);
        const int colc = 3 * (xs - xs_min) + (ALIGNED ? mis0 : 0);
        const int colc4 = colc & ~3, shc = (colc & 3) * 8;
        const int rows_fit = ALIGNED ? (WARP_BUF / pitch) / rb * rb : WARP_BUF / pitch;
        // tensor coordinates of the strip's first staged byte: flattened image row and uint32 column
        const unsigned long long seg_off = (src_seg & ~15ull) - (unsigned long long)(uintptr_t)images;
        const int row0 = ALIGNED ? (int)(seg_off / rowstride) : 0;
        const int xw = ALIGNED ? (int)((seg_off - (unsigned long long)row0 * rowstride) >> 2) : 0;
        const double scale_y = gp->scale_y;
        // bh output rows tap at most bh*scale + 2 source rows; whole TMA boxes round that up by rb - 1 more
        const int bh = max(1, min(32, (int)((double)(rows_fit - (ALIGNED ? 5 : 3)) / (scale_y < 1.0 ? 1.0 : scale_y))));
        const int nb = (new_h + bh - 1) / bh;
        const float4* ydr = ydesc + (size_t)roi * ystride;
        float* optr = outf + ((size_t)roi * 3 * T + dy0) * T + x;     // (plane 0, current row, column x)
        const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);

        // descriptors of batch b (rows b*bh ..) -> ring k, source rows -> buffer k; returns the first source row
        int bulk_cur = 0, bulk_next = 0;
        auto stage = [&](int b, int k, const float4& yd) -> int {
            const int cnt = min(bh, new_h - b * bh);
            if (lane < cnt) reinterpret_cast<float4*>(wbase + 2 * WARP_BUF + k * WARP_DESC)[lane] = yd;
            int lo, hi;
            if (cls != 3) {
                const int ysn = __float_as_int(yd.w);
                lo = ysn & 0xffffff; hi = lo + (ysn >> 24) - 1;
            } else {
                lo = __float_as_int(yd.w); hi = __float_as_int(yd.z);
            }
            const int s_lo = __shfl_sync(0xffffffffu, lo, 0);
            const int s_hi = __shfl_sync(0xffffffffu, hi, cnt - 1);
            if (ALIGNED) {
                warp_stage_2d(wbase_s + k * WARP_BUF, bar_s + 8 * k, map, xw, row0 + s_lo, s_hi - s_lo + 1, pitch, rb, lane);
                bulk_next = 1;
            } else {
                bulk_next = warp_stage(wbase + k * WARP_BUF, wbase_s + k * WARP_BUF, bar_s + 8 * k, src_seg, rowstride, img_end, s_lo,
                                       s_hi - s_lo + 1, pitch, lane, lr, lv, rpp);
            }
            return s_lo;
        };
        auto load_desc = [&](int b) -> float4 {
            const int y = b * bh + lane;
            return (b < nb && y < new_h && lane < bh) ? ydr[y] : zero4;
        };

        if (cls == 4) {
            // ---------------- regime 1, up to 6 taps per axis (2 <= scale <= 5) ----------------
            // Source rows stream through a ring of four slots of G rows (one mbarrier each, three slots in flight while
            // one is read); every row is staged once and its horizontal pass lives in registers, so output rows simply
            // consume rows in order -- no per-output-row batches, which at scale 4 held a single row each.
            const int nt = warp_max_i32(xn);                      // uniform: taps evaluated per source row
            float w[6], c[6];
#pragma unroll
            for (int k = 0; k < 6; ++k) {
                w[k] = (k == 0) ? xd.x : ((k < xn - 1) ? xd.y : ((k == xn - 1) ? xd.z : 0.f));
                c[k] = __fmul_rn(w[k], -8388608.0f);
                asm volatile("" : "+f"(c[k]));
            }
            const int s_first = __float_as_int(ydr[0].w) & 0xffffff;
            const int ysn_last = __float_as_int(ydr[new_h - 1].w);
            const int s_end = (ysn_last & 0xffffff) + (ysn_last >> 24);          // one past the last source row
            const int G = ALIGNED ? ((2 * WARP_BUF / pitch) >> 2) / rb * rb       // rows per slot: whole TMA boxes
                                  : min(32, (2 * WARP_BUF / pitch) >> 2);       // (1-D copies: lane r issues row r)
            const int ngroups = (s_end - s_first + G - 1) / G;
            const unsigned slot_bytes = (unsigned)(G * pitch);
            unsigned bulkmask = 0;
            auto issue = [&](int g) {
                const int j = g & 3, lo = s_first + g * G;
                int bulk = 1;
                if (ALIGNED) warp_stage_2d(wbase_s + j * slot_bytes, bar_s + 8 * j, map, xw, row0 + lo, min(G, s_end - lo), pitch, rb, lane);
                else bulk = warp_stage(wbase + j * slot_bytes, wbase_s + j * slot_bytes, bar_s + 8 * j, src_seg, rowstride, img_end,
                                       lo, min(G, s_end - lo), pitch, lane, lr, lv, rpp);
                bulkmask = (bulkmask & ~(1u << j)) | ((unsigned)bulk << j);
            };
            auto wait = [&](int g) {
                const int j = g & 3;
                if ((bulkmask >> j) & 1u) { mbar_wait(bar_s + 8 * j, (ph >> j) & 1u); ph ^= 1u << j; }
                else cp_async_wait_all();
                __syncwarp();
            };
            __syncwarp();                                   // previous item finished with the buffers
            for (int g = 0; g < min(4, ngroups); ++g) issue(g);
            wait(0);
            int gi = 0, gend = s_first + G;                 // current group and one past its last row
            int crow = -1;
            u64 hc01 = 0ull;
            float hc2 = 0.f;
            auto strip = [&](auto nt_c) {
                constexpr int NT = decltype(nt_c)::value;
                unsigned gbase = wbase_s + colc4 - (unsigned)(s_first * pitch);          // ALIGNED: row r of the group at gbase + r * pitch
                auto hrow = [&](int row) {
                    if (row >= gend) {
                        do {
                            __syncwarp();                   // every lane is done with slot gi & 3
                            if (gi + 4 < ngroups) issue(gi + 4);
                            ++gi;
                            wait(gi);
                            gend += G;
                        } while (row >= gend);
                        gbase = wbase_s + (unsigned)(gi & 3) * slot_bytes + colc4 - (unsigned)((gend - G) * pitch);
                    }
                    if (ALIGNED) {
                        h_area_n<NT>(gbase + (unsigned)(row * pitch), shc, w, c, hc01, hc2);
                    } else {
                        const int a = (row - (gend - G)) * pitch + colc + ((mis0 + row * misstep) & 15);
                        h_area_n<NT>(wbase_s + (unsigned)(gi & 3) * slot_bytes + (a & ~3), (a & 3) * 8, w, c, hc01, hc2);
                    }
                };
                float4 d = ydr[0];
                for (int y = 0; y < new_h; ++y) {
                    const float4 dn = ydr[min(y + 1, new_h - 1)];
                    const int ysn = __float_as_int(d.w);
                    const int ys = ysn & 0xffffff, n = ysn >> 24;
                    if (ys != crow) hrow(ys);                // else: the previous output row ended on this source row
                    u64 acc01 = fprod2(pack2(d.x, d.x), hc01, nz2);
                    float acc2 = __fmul_rn(d.x, hc2);
                    for (int t = 1; t < n; ++t) {
                        hrow(ys + t);
                        const float beta = (t == n - 1) ? d.z : d.y;
                        acc01 = fadd2(acc01, fprod2(pack2(beta, beta), hc01, nz2));
                        acc2 = __fadd_rn(acc2, __fmul_rn(beta, hc2));
                    }
                    crow = ys + n - 1;
                    if (active) {
                        float a0f, a1f;
                        unpack2(acc01, a0f, a1f);
                        if (OUT_U8) {
                            out.px(roi, dy0 + y, x, round_u8(a0f), round_u8(a1f), round_u8(acc2));
                        } else {
                            const unsigned l0 = lut_addr(a0f, lut_m), l1 = lut_addr(a1f, lut_m), l2 = lut_addr(acc2, lut_m);
                            optr[0] = lds_f32(swap ? l2 : l0);
                            optr[plane] = lds_f32(l1 + 4 * LUT_STRIDE);
                            optr[2 * plane] = lds_f32((swap ? l0 : l2) + 8 * LUT_STRIDE);
                            optr += T;
                        }
                    } else if (padlane) {
                        out.pad(roi, dy0 + y, x);
                    }
                    d = dn;
                }
            };
            if (nt <= 4) strip(std::integral_constant<int, 4>{});
            else if (nt == 5) strip(std::integral_constant<int, 5>{});
            else strip(std::integral_constant<int, 6>{});
            continue;
        }

        if (ALIGNED && cls == 1) {
            // ---------------- regime 1, <= 3 taps per axis (scale < 2): source rows stream in order ----------------
            // Ring of 8-row slots (one 2-D TMA box + the 64 bytes of the eight rows' (ba, bb) records per slot, one mbarrier
            // each, every slot but the one being read in flight); the horizontal pass of four rows is computed back to back
            // (12 independent loads), then every row is added to the open output row with weight |ba|; a row whose record
            // has the sign of ba set closes that output row (LUT, three 128-byte stores) and opens the next one with
            // weight bb.  No per-output-row tap loop, no tap-count branches, every source row staged exactly once.
            ColW cw;
            cw.set(xd.x, xd.y, xd.z);
            if (!active) {                               // beside the image: every row sums to 256 -> LUT entry 256 = fill
                cw.w[0] = cw.w[1] = cw.w[2] = 0.f;
                cw.c[0] = 256.f; cw.c[1] = cw.c[2] = 0.f;
            }
            const int nchunks = (gp->h + STREAM_ROWS - 1) / STREAM_ROWS;
            const unsigned slot_bytes = (unsigned)(STREAM_ROWS * pitch);
            const int nsl = min(4, (2 * WARP_BUF) / (int)slot_bytes);                   // >= 3 (pitch <= STREAM_MAX_PITCH)
            const unsigned dring = wbase_s + 2 * WARP_BUF;                              // [4 slots][8 rows] float2
            const unsigned long long ysrc = (unsigned long long)(uintptr_t)(ydesc + (size_t)roi * ystride);
            const CUtensorMap* map8 = &tm.m8[min(tmap_index(pitch), N_TMAPS8 - 1)];
            auto issue = [&](int c, int j) {
                if (lane == 0) {
                    mbar_expect_tx(bar_s + 8 * j, slot_bytes + 8u * STREAM_ROWS);
                    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                                 :: "r"(wbase_s + (unsigned)j * slot_bytes), "l"(map8), "r"(xw), "r"(row0 + c * STREAM_ROWS), "r"(bar_s + 8 * j) : "memory");
                    bulk_g2s(dring + 8u * STREAM_ROWS * j, ysrc + 8ull * STREAM_ROWS * c, 8u * STREAM_ROWS, bar_s + 8 * j);
                }
            };
            __syncwarp();                                   // previous item finished with the buffers
            for (int c = 0; c < min(nsl, nchunks); ++c) issue(c, c);
            u64 acc01 = 0ull;
            float acc2 = 0.f;
            unsigned roff = 0;                                // element offset of the open output row from optr
            int yout = 0;
            const bool store_ok = (TT && TT % 32 == 0) || x < T;
            int j = 0;
            for (int c = 0; c < nchunks; ++c) {
                mbar_wait(bar_s + 8 * j, (ph >> j) & 1u); ph ^= 1u << j;
                const unsigned rbase = wbase_s + (unsigned)j * slot_bytes + colc4;
#pragma unroll
                for (int half = 0; half < STREAM_ROWS / 4; ++half) {
                    const float4 dA = lds_f4(dring + 8u * STREAM_ROWS * j + 32u * half), dB = lds_f4(dring + 8u * STREAM_ROWS * j + 32u * half + 16);
                    u64 h01[4];
                    float h2[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) h_area3(rbase + (unsigned)((4 * half + k) * pitch), shc, cw, h01[k], h2[k]);
                    if (half == STREAM_ROWS / 4 - 1) {
                        __syncwarp();                           // every lane has read slot j: refill it
                        if (c + nsl < nchunks) issue(c + nsl, j);
                    }
                    const float ba[4] = {dA.x, dA.z, dB.x, dB.z}, bb[4] = {dA.y, dA.w, dB.y, dB.w};
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const float wa = fabsf(ba[k]);
                        acc01 = fadd2(acc01, fprod2(pack2(wa, wa), h01[k], nz2));
                        acc2 = __fadd_rn(acc2, __fmul_rn(wa, h2[k]));
                        if (__float_as_int(ba[k]) < 0) {            // output row complete
                            float a0f, a1f;
                            unpack2(acc01, a0f, a1f);
                            if (OUT_U8) {
                                if (active) out.px(roi, dy0 + yout, x, round_u8(a0f), round_u8(a1f), round_u8(acc2));
                                else if (padlane) out.pad(roi, dy0 + yout, x);
                                ++yout;
                            } else {
                                const unsigned l0 = lut_addr(a0f, lut_m), l1 = lut_addr(a1f, lut_m), l2 = lut_addr(acc2, lut_m);
                                if (store_ok) {
                                    float* o = optr + roff;
                                    stg_out<0>(o, lds_f32(swap ? l2 : l0));
                                    stg_out<1>(o + plane, lds_f32(l1 + 4 * LUT_STRIDE));
                                    stg_out<2>(o + 2 * plane, lds_f32((swap ? l0 : l2) + 8 * LUT_STRIDE));
                                }
                                roff += (unsigned)T;
                            }
                            acc01 = fprod2(pack2(bb[k], bb[k]), h01[k], nz2);
                            acc2 = __fmul_rn(bb[k], h2[k]);
                        }
                    }
                }
                j = (j + 1 == nsl) ? 0 : j + 1;
            }
            continue;
        }

        __syncwarp();                                   // previous item finished with the buffers
        int s_lo_cur = stage(0, 0, load_desc(0));
        bulk_cur = bulk_next;
        float4 ydn = load_desc(1);
        int s_lo_next = 0;

        if (!ALIGNED && cls == 1) {
            ColW cw;
            cw.set(xd.x, xd.y, xd.z);
            int crow = -1;
            u64 hc01 = 0ull;
            float hc2 = 0.f;
            unsigned roff = 0;                                // element offset of the current output row from optr
            for (int b = 0; b < nb; ++b) {
                const int k = b & 1;
                if (bulk_cur) {
                    mbar_wait(bar_s + 8 * k, (ph >> k) & 1u); ph ^= 1u << k;
                } else {
                    cp_async_wait_all();
                }
                __syncwarp();
                if (b + 1 < nb) s_lo_next = stage(b + 1, k ^ 1, ydn);
                ydn = load_desc(b + 2);
                const unsigned cur = wbase_s + k * WARP_BUF, ring = wbase_s + 2 * WARP_BUF + k * WARP_DESC;
                const int y0 = b * bh, cnt = min(bh, new_h - y0);
                if (active) {
                    float4 dnx = lds_f4(ring);
                    for (int r = 0; r < cnt; ++r) {
                        const float4 d = dnx;
                        dnx = lds_f4(ring + (r + 1) * 16);          // next row's descriptor in flight behind this row's arithmetic
                        const int ysn = __float_as_int(d.w);
                        const int ys = ysn & 0xffffff, n = ysn >> 24;
                        // byte address of the lane's first tap in row ys; with a 16-byte-multiple image pitch the
                        // word offset and the funnel shift are per-item constants (pitch is a multiple of 16)
                        int a = (ys - s_lo_cur) * pitch + colc;
                        if (!ALIGNED) a += (mis0 + ys * misstep) & 15;
                        const unsigned a4 = ALIGNED ? cur + (ys - s_lo_cur) * pitch + colc4 : cur + (a & ~3);
                        if (ys != crow) h_area3(a4, ALIGNED ? shc : (a & 3) * 8, cw, hc01, hc2);
                        u64 acc01 = fprod2(pack2(d.x, d.x), hc01, nz2);
                        float acc2 = __fmul_rn(d.x, hc2);
                        if (n > 1) {
                            int a1 = a + pitch;
                            if (!ALIGNED) a1 = (ys + 1 - s_lo_cur) * pitch + colc + ((mis0 + (ys + 1) * misstep) & 15);
                            h_area3(ALIGNED ? a4 + pitch : cur + (a1 & ~3), ALIGNED ? shc : (a1 & 3) * 8, cw, hc01, hc2);
                            acc01 = fadd2(acc01, fprod2(pack2(d.y, d.y), hc01, nz2));
                            acc2 = __fadd_rn(acc2, __fmul_rn(d.y, hc2));
                        }
                        if (n > 2) {
                            int a2 = a + 2 * pitch;
                            if (!ALIGNED) a2 = (ys + 2 - s_lo_cur) * pitch + colc + ((mis0 + (ys + 2) * misstep) & 15);
                            h_area3(ALIGNED ? a4 + 2 * pitch : cur + (a2 & ~3), ALIGNED ? shc : (a2 & 3) * 8, cw, hc01, hc2);
                            acc01 = fadd2(acc01, fprod2(pack2(d.z, d.z), hc01, nz2));
                            acc2 = __fadd_rn(acc2, __fmul_rn(d.z, hc2));
                        }
                        crow = ys + n - 1;
                        float a0f, a1f;
                        unpack2(acc01, a0f, a1f);
                        if (OUT_U8) {
                            out.px(roi, dy0 + y0 + r, x, round_u8(a0f), round_u8(a1f), round_u8(acc2));
                        } else {
                            // bits(v + 2^23) = 0x4B000000 + cvRound(v): the LUT address is one multiply-add away
                            const unsigned l0 = lut_addr(a0f, lut_m), l1 = lut_addr(a1f, lut_m), l2 = lut_addr(acc2, lut_m);
                            float* o = optr + roff;               // fresh address registers per row: no wait on the previous row's stores
                            o[0] = lds_f32(swap ? l2 : l0);
                            o[plane] = lds_f32(l1 + 4 * LUT_STRIDE);
                            o[2 * plane] = lds_f32((swap ? l0 : l2) + 8 * LUT_STRIDE);
                            roff += (unsigned)T;
                        }
                    }
                } else if (padlane) {
                    for (int r = 0; r < cnt; ++r) out.pad(roi, dy0 + y0 + r, x);
                }
                s_lo_cur = s_lo_next;
                bulk_cur = bulk_next;
            }
        } else {
            const int xw0 = __float_as_int(xd.x), xw1 = __float_as_int(xd.y);
            int rowA = -1, rowB = -1;
            int HA[3] = {0, 0, 0}, HB[3] = {0, 0, 0};
            for (int b = 0; b < nb; ++b) {
                const int k = b & 1;
                if (bulk_cur) {
                    mbar_wait(bar_s + 8 * k, (ph >> k) & 1u); ph ^= 1u << k;
                } else {
                    cp_async_wait_all();
                }
                __syncwarp();
                if (b + 1 < nb) s_lo_next = stage(b + 1, k ^ 1, ydn);
                ydn = load_desc(b + 2);
                const unsigned cur = wbase_s + k * WARP_BUF, ring = wbase_s + 2 * WARP_BUF + k * WARP_DESC;
                const int y0 = b * bh, cnt = min(bh, new_h - y0);
                if (active) {
                    for (int r = 0; r < cnt; ++r) {
                        const float4 d = lds_f4(ring + r * 16);
                        const int b0 = __float_as_int(d.x), b1 = __float_as_int(d.y);
                        const int sy0 = __float_as_int(d.w), sy1 = __float_as_int(d.z);
                        if (sy0 != rowA) {
                            if (sy0 == rowB) { HA[0] = HB[0]; HA[1] = HB[1]; HA[2] = HB[2]; }
                            else {
                                const int a = (sy0 - s_lo_cur) * pitch + colc + (ALIGNED ? 0 : ((mis0 + sy0 * misstep) & 15));
                                h_lin(cur + (a & ~3), (a & 3) * 8, xw0, xw1, HA);
                            }
                            rowA = sy0;
                        }
                        if (sy1 != rowB) {
                            if (sy1 == rowA) { HB[0] = HA[0]; HB[1] = HA[1]; HB[2] = HA[2]; }
                            else {
                                const int a = (sy1 - s_lo_cur) * pitch + colc + (ALIGNED ? 0 : ((mis0 + sy1 * misstep) & 15));
                                h_lin(cur + (a & ~3), (a & 3) * 8, xw0, xw1, HB);
                            }
                            rowB = sy1;
                        }
                        int o[3];
#pragma unroll
                        for (int c = 0; c < 3; ++c) o[c] = ((((b0 * HA[c]) >> 16) + ((b1 * HB[c]) >> 16) + 2) >> 2) & 255;
                        if (OUT_U8) {
                            out.px(roi, dy0 + y0 + r, x, o[0], o[1], o[2]);
                        } else {
                            optr[0] = lds_f32(lut_s + 4 * (swap ? o[2] : o[0]));
                            optr[plane] = lds_f32(lut_s + 4 * LUT_STRIDE + 4 * o[1]);
                            optr[2 * plane] = lds_f32(lut_s + 8 * LUT_STRIDE + 4 * (swap ? o[0] : o[2]));
                            optr += T;
                        }
                    }
                } else if (padlane) {
                    for (int r = 0; r < cnt; ++r) out.pad(roi, dy0 + y0 + r, x);
                }
                s_lo_cur = s_lo_next;
                bulk_cur = bulk_next;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------------
// CTA kernel: classes 1, 3 and 4, one crop per CTA, warp-specialised
// ------------------------------------------------------------------------------------------------------
// Warps 0 .. NS-1 (NS = T / 32 strips) are consumers: one lane per output column, source rows in order.  The LAST warp is a
// producer whose lane 0 walks the list of class-1 / class-3 crops (atomic counter) and, per crop, copies the crop's block of
// row descriptors into shared memory (one bulk copy) and then, per ring slot, issues ONE 2-D tensor copy of eight FULL-WIDTH
// source rows -- seven times fewer TMA operations than one box per strip, every source byte fetched once per crop, and the
// ~300-cycle scoreboard wait behind each TMA issue (measured: 13 % of the per-strip kernel's warp time) sits in a warp that has
// nothing else to do.  The producer runs ahead across crops (ring of four slots, two descriptor blocks), which also hides
// the per-crop set-up round trips.  Consumers meet on the slots' full / empty mbarriers, so the strips of one crop stay within
// four slots of each other: whole 896-byte output rows of a plane reach L2 close together, and the rows above / below the
// resized image are written as contiguous runs by all consumer threads.
//   class 1 (area, <= 3 taps): per SOURCE row a (ba, bb) record (see bpc_crop_prep_kernel); a row whose ba has the sign bit
//            set closes the open output row (LUT, three 128-byte stores) and opens the next with weight bb;
//   class 3 (fixed-point bilinear, the box grows): per OUTPUT row (b0, b1, second source row, first source row); an output row
//            is emitted as soon as the slot holding its second source row has landed -- its first row is the previous output
//            row's first or second row, whose horizontal pass is still in registers.
constexpr int CTA_ROWS = 8;                  // source rows per ring slot (= one TMA box)
// ring slots and resident CTAs per SM: the T = 224 float instantiation (the benchmark's) runs 4 CTAs/SM (8 warps, 63 registers,
// three slots: 55 KB of shared memory), which buys classes 3 and 4 latency hiding (0.66 -> 0.70, 0.67 -> 0.69) and leaves class 1
// where it was; the run-time-T and T = 256 instantiations (up to 9 warps) keep four slots and 3 CTAs/SM
__host__ __device__ constexpr int cta_nslot(int TT) { return TT == 224 ? 3 : 4; }
constexpr int CTA_NMAPS = 25;                // box widths 64, 128, ... 1600 bytes (8-byte elements)
constexpr int CTA_MAX_T = 256;               // 8 consumer warps
struct CtaMaps { CUtensorMap m[CTA_NMAPS]; };
__host__ __device__ __forceinline__ int cta_pitch_max(int T) { return ((15 + 3 * (2 * T - 1) + 12 + 63) >> 6) << 6; }
// staged bytes per source row of a crop: misalignment of its first byte + 3 w + what the widest horizontal pass over-reads
// (class 1: three words from the aligned tap address; class 4: up to six taps padded to the warp maximum)
__host__ __device__ __forceinline__ int cta_pitch(int mis0, int w, int cls) { return ((mis0 + 3 * w + (cls == 4 ? 40 : 12) + 63) >> 6) << 6; }
__host__ __device__ __forceinline__ int cta_desc_bytes(int T) { return ydesc_stride(T) * 16; }
__host__ __device__ __forceinline__ int cta_smem_bytes(int T, int nslot) { return LUT_SMEM + nslot * CTA_ROWS * cta_pitch_max(T) + 2 * cta_desc_bytes(T) + 256; }

// Lockstep of a crop's strips (a named barrier over the consumer warps): the seven 128-byte pieces of an output row then reach
// L2 within a short window and are written back together -- DRAM sees whole rows instead of scattered lines (measured on
// stores alone: 5.7 -> 7.2 TB/s; on the 60-400 px mix 0.79 -> 0.83 of roofline).  Class 1 only, once per ring slot: a barrier every
// eight output rows made class 3 slower (0.64 -> 0.57: many idle strips, latency-bound) and class 4 is issue-bound.
#define CTA_LOCKSTEP() do { if (cls == 1) asm volatile("bar.sync 1, %0;" :: "r"(NS * 32) : "memory"); } while (0)
__device__ __forceinline__ void mbar_arrive(unsigned bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(bar) : "memory"); }

template <bool OUT_U8, int TT, bool SWAP, bool BF16 = false>
__global__ void __launch_bounds__(TT == 224 ? 256 : 288, TT == 224 ? 4 : 3)
bpc_crop_cta_kernel(const uint8_t* __restrict__ images, int B, int H, int W, const RoiGeom* __restrict__ geom,
                    const float4* __restrict__ xdesc, const float4* __restrict__ ydesc, const int32_t* __restrict__ list1,
                    int32_t* __restrict__ counters, int Trt, uchar4 fill, int swap_rb, const float* __restrict__ lut_g,
                    float* __restrict__ outf, uint8_t* __restrict__ outb, const __grid_constant__ CtaMaps tm) {
    extern __shared__ __align__(128) unsigned char smem[];
    constexpr int CTA_NSLOT = cta_nslot(TT);
    float* lut = reinterpret_cast<float*>(smem);
    const int T = TT ? TT : Trt;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int NS = (T + 31) >> 5;                                   // consumer warps; warp NS is the producer
    const int pitch_max = cta_pitch_max(T);
    const int slot_bytes = CTA_ROWS * pitch_max;
    const int desc_bytes = cta_desc_bytes(T);
    const unsigned smem_s = (unsigned)__cvta_generic_to_shared(smem);
    const unsigned ring_s = smem_s + LUT_SMEM;
    const unsigned desc_s = ring_s + CTA_NSLOT * slot_bytes;       // [2][ystride] float4: the current and the next crop's row descriptors
    const unsigned misc_s = desc_s + 2 * desc_bytes;
    const unsigned full_s = misc_s, empty_s = misc_s + 8 * CTA_NSLOT, hfull_s = misc_s + 16 * CTA_NSLOT, hempty_s = hfull_s + 16;
    volatile int* hdr = reinterpret_cast<volatile int*>(smem + (misc_s - smem_s) + 16 * CTA_NSLOT + 32);       // [2] crop index
    if (tid == 0) {
        for (int i = 0; i < CTA_NSLOT; ++i) { mbar_init(full_s + 8 * i, 1); mbar_init(empty_s + 8 * i, NS); }
        for (int i = 0; i < 2; ++i) { mbar_init(hfull_s + 8 * i, 1); mbar_init(hempty_s + 8 * i, NS); }
        *reinterpret_cast<volatile unsigned*>(smem + 3 * LUT_STRIDE * 4) = smem_s - 4u * 0x4B000000u;              // see lut_addr()
    }
    if (!OUT_U8)
        for (int e = tid; e < 768; e += (int)blockDim.x) lut[(e >> 8) * LUT_STRIDE + (e & 255)] = lut_g[e];
    __syncthreads();
    Out<OUT_U8, BF16> out;
    out_init(out, outf, outb, lut, T, swap_rb, fill);
    if (!OUT_U8 && tid < 3) lut[tid * LUT_STRIDE + 256] = out.padf[tid];
    __syncthreads();
    const unsigned lut_m = *reinterpret_cast<volatile unsigned*>(smem + 3 * LUT_STRIDE * 4);
    const int ds = desc_stride(T), ystride = ydesc_stride(T);
    const unsigned long long rowstride = (unsigned long long)W * 3ull;
    const unsigned long long img_end = (unsigned long long)(uintptr_t)images + (unsigned long long)B * H * rowstride;
    const int n1 = counters[8];
    int32_t* work = counters + 12;

    if (wid == NS) {
        // ------------------------------------ producer ------------------------------------
        if (lane != 0) return;
        int idx = atomicAdd(work, 1);
        unsigned cg = 0;                                            // slots filled so far
        for (int hi = 0;; ++hi) {
            const int hb = hi & 1;
            if (hi >= 2) mbar_wait(hempty_s + 8 * hb, ((hi >> 1) - 1) & 1);
            if (idx >= n1) {
                hdr[hb] = -1;
                mbar_arrive(hfull_s + 8 * hb);
                break;
            }
            const int idx_next = atomicAdd(work, 1);                // round trip hidden behind this crop's copies
            const int roi = list1[idx];
            const RoiGeom* gp = geom + roi;
            const unsigned long long src = gp->src;
            const int w = gp->w, h = gp->cls == 0 ? 0 : gp->h;     // a rejected box (class 0) has no source rows: header only
            hdr[hb] = roi;
            // the part of the crop's row-descriptor block that will be read: per output row (class 3) or per source row, padded slot included
            const unsigned dbytes = gp->cls == 0 ? 0u : (unsigned)min(desc_bytes, gp->cls == 3 ? 16 * (gp->new_h + 1) : 8 * (((h + CTA_ROWS - 1) & ~(CTA_ROWS - 1)) + CTA_ROWS));
            mbar_expect_tx(hfull_s + 8 * hb, dbytes);
            if (dbytes) bulk_g2s(desc_s + hb * desc_bytes, (unsigned long long)(uintptr_t)(ydesc + (size_t)roi * ystride), dbytes, hfull_s + 8 * hb);
            const int mis0 = (int)(src & 15ull);
            const int pitch = cta_pitch(mis0, w, gp->cls);
            const unsigned long long off = (src & ~15ull) - (unsigned long long)(uintptr_t)images;
            const int row0 = (int)(off / rowstride);
            const int x8 = (int)((off - (unsigned long long)row0 * rowstride) >> 3);
            if (gp->cls == 0) {
                // nothing to stage
            } else if (pitch <= pitch_max) {
                // one 2-D tensor copy of eight full-width rows per slot (rows / columns beyond the pool are zero-filled)
                const CUtensorMap* map = &tm.m[(pitch >> 6) - 1];
                const int nchunks = (h + CTA_ROWS - 1) / CTA_ROWS;
                for (int c = 0; c < nchunks; ++c, ++cg) {
                    const unsigned j = cg % CTA_NSLOT;
                    if (cg >= CTA_NSLOT) mbar_wait(empty_s + 8 * j, ((cg / CTA_NSLOT) - 1) & 1);
                    mbar_expect_tx(full_s + 8 * j, (unsigned)(CTA_ROWS * pitch));
                    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                                 :: "r"(ring_s + j * slot_bytes), "l"(map), "r"(x8), "r"(row0 + c * CTA_ROWS), "r"(full_s + 8 * j) : "memory");
                }
            } else {
                // wide class-4 crops (rows of up to 5 T pixels): rps rows per slot, one 1-D bulk copy each, clipped to the pool
                const int rps = slot_bytes / pitch;
                const int nchunks = (h + rps - 1) / rps;
                const unsigned long long a0 = src & ~15ull;
                for (int c = 0; c < nchunks; ++c, ++cg) {
                    const unsigned j = cg % CTA_NSLOT;
                    if (cg >= CTA_NSLOT) mbar_wait(empty_s + 8 * j, ((cg / CTA_NSLOT) - 1) & 1);
                    unsigned total = 0;
                    for (int r = 0; r < rps; ++r) {
                        const unsigned long long a = a0 + (unsigned long long)(c * rps + r) * rowstride;
                        if (a < img_end) total += (unsigned)min((unsigned long long)pitch, (img_end - a) & ~15ull);
                    }
                    mbar_expect_tx(full_s + 8 * j, total);
                    for (int r = 0; r < rps; ++r) {
                        const unsigned long long a = a0 + (unsigned long long)(c * rps + r) * rowstride;
                        if (a < img_end) {
                            const unsigned nb = (unsigned)min((unsigned long long)pitch, (img_end - a) & ~15ull);
                            if (nb) bulk_g2s(ring_s + j * slot_bytes + (unsigned)(r * pitch), a, nb, full_s + 8 * j);
                        }
                    }
                }
            }
            idx = idx_next;
        }
        return;
    }

    // ------------------------------------ consumers: warp = strip ------------------------------------
    constexpr bool swap = SWAP;
    const size_t plane = (size_t)T * T;
    const float nz = __int_as_float((int)(0x80000000u | (unsigned)fill.w));     // -0.0f: fill.w is 0 at run time
    const u64 nz2 = pack2(nz, nz);
    const int nthc = NS * 32;
    const int x = wid * 32 + lane;
    unsigned cg = 0;
    for (int hi = 0;; ++hi) {
        const int hb = hi & 1;
        mbar_wait(hfull_s + 8 * hb, (hi >> 1) & 1);
        const int roi = hdr[hb];
        if (roi < 0) break;
        const RoiGeom* gp = geom + roi;
        const int cls = gp->cls;
        const int new_w = gp->new_w, new_h = gp->new_h, dx0 = gp->dx, dy0 = gp->dy, h = gp->h;
        const int mis0 = (int)(gp->src & 15ull);
        const int pitch = cta_pitch(mis0, gp->w, cls);
        const int rps = pitch <= pitch_max ? CTA_ROWS : slot_bytes / pitch;       // source rows per ring slot
        const int nchunks = (h + rps - 1) / rps;
        const unsigned dsc = desc_s + hb * desc_bytes;
        if (cls == 0) {                                             // rejected box: the whole canvas is fill (no slots were issued)
            out.pad_rows(roi, 0, T, tid, nthc);
            __syncwarp();
            if (lane == 0) mbar_arrive(hempty_s + 8 * hb);
            continue;
        }
        out.pad_rows(roi, 0, dy0, tid, nthc);                       // whole rows above / below: contiguous runs
        out.pad_rows(roi, dy0 + new_h, T, tid, nthc);
        const int xr = x - dx0;
        const bool active = xr >= 0 && xr < new_w;
        const bool touches = wid * 32 < dx0 + new_w && wid * 32 + 32 > dx0;
        const bool store_ok = (TT && TT % 32 == 0) || x < T;
        float* optr = outf + ((size_t)roi * 3 * T + dy0) * T + x;     // (plane 0, first image row, column x)
        unsigned short* optr16 = reinterpret_cast<unsigned short*>(outf) + (((size_t)roi * T + dy0) * T + x) * 3;   // BF16: pixel (row, x), channels-last
        const unsigned short padh0 = bf16_bits(out.padf[0]), padh1 = bf16_bits(out.padf[1]), padh2 = bf16_bits(out.padf[2]);
        if (!touches) {
            // a strip beside the resized image: the fill value, at the pace of the neighbours
            int ydone = 0;
            for (int c = 0; c < nchunks; ++c, ++cg) {
                const unsigned j = cg % CTA_NSLOT;
                CTA_LOCKSTEP(); mbar_wait(full_s + 8 * j, (cg / CTA_NSLOT) & 1);
                __syncwarp();
                if (lane == 0) mbar_arrive(empty_s + 8 * j);
                int ndone;
                if (cls != 3) {
                    const float ba = (lane < rps) ? lds_f32(dsc + 8u * (unsigned)(c * rps + lane)) : 0.f;
                    ndone = __popc(__ballot_sync(0xffffffffu, __float_as_int(ba) < 0));
                } else {
                    // output rows whose second source row lies in this slot (monotone in y)
                    const int last_row = c * CTA_ROWS + CTA_ROWS - 1;
                    ndone = 0;
                    for (int y0 = ydone; y0 < new_h; y0 += 32) {
                        const int yy = y0 + lane;
                        const bool ok = yy < new_h && __float_as_int(lds_f4(dsc + 16u * (unsigned)min(yy, new_h - 1)).z) <= last_row;
                        const int n = __popc(__ballot_sync(0xffffffffu, ok));
                        ndone += n;
                        if (n < 32) break;
                    }
                }
                for (int r = 0; r < ndone; ++r) {
                    if (store_ok) {
                        if (OUT_U8) out.pad(roi, dy0 + ydone + r, x);
                        else if (BF16) { optr16[0] = padh0; optr16[1] = padh1; optr16[2] = padh2; optr16 += 3 * T; }
                        else { optr[0] = out.padf[0]; optr[plane] = out.padf[1]; optr[2 * plane] = out.padf[2]; optr += T; }
                    }
                }
                ydone += ndone;
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(hempty_s + 8 * hb);         // done with this crop's descriptor block
            continue;
        }
        const float4 xd = xdesc[(size_t)roi * ds + min(max(xr, 0), new_w - 1)];
        const int xs = __float_as_int(xd.w) & 0xffffff;
        const int colc = 3 * xs + mis0;
        const unsigned colc4 = (unsigned)(colc & ~3);
        const int shc = (colc & 3) * 8;
        if (cls == 1) {
            ColW cw;
            cw.set(xd.x, xd.y, xd.z);
            if (!active) {                               // beside the image: every row sums to 256 -> LUT entry 256 = fill
                cw.w[0] = cw.w[1] = cw.w[2] = 0.f;
                cw.c[0] = 256.f; cw.c[1] = cw.c[2] = 0.f;
            }
            u64 acc01 = 0ull;
            float acc2 = 0.f;
            unsigned roff = 0;                                // element offset of the open output row from optr
            int yout = 0;
            for (int c = 0; c < nchunks; ++c, ++cg) {
                const unsigned j = cg % CTA_NSLOT;
                CTA_LOCKSTEP(); mbar_wait(full_s + 8 * j, (cg / CTA_NSLOT) & 1);
                const unsigned rbase = ring_s + j * slot_bytes + colc4;
                const unsigned rec = dsc + 8u * CTA_ROWS * (unsigned)c;
#pragma unroll
                for (int half = 0; half < CTA_ROWS / 4; ++half) {
                    const float4 dA = lds_f4(rec + 32u * half), dB = lds_f4(rec + 32u * half + 16);
                    u64 h01[4];
                    float h2[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) h_area3(rbase + (unsigned)((4 * half + k) * pitch), shc, cw, h01[k], h2[k]);
                    if (half == CTA_ROWS / 4 - 1) {
                        __syncwarp();                           // every lane has read slot j
                        if (lane == 0) mbar_arrive(empty_s + 8 * j);
                    }
                    const float ba[4] = {dA.x, dA.z, dB.x, dB.z}, bb[4] = {dA.y, dA.w, dB.y, dB.w};
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const float wa = fabsf(ba[k]);
                        acc01 = fadd2(acc01, fprod2(pack2(wa, wa), h01[k], nz2));
                        acc2 = __fadd_rn(acc2, __fmul_rn(wa, h2[k]));
                        if (__float_as_int(ba[k]) < 0) {            // output row complete
                            float a0f, a1f;
                            unpack2(acc01, a0f, a1f);
                            if (OUT_U8) {
                                if (active) out.px(roi, dy0 + yout, x, round_u8(a0f), round_u8(a1f), round_u8(acc2));
                                else if (x < T) out.pad(roi, dy0 + yout, x);
                                ++yout;
                            } else {
                                const unsigned l0 = lut_addr(a0f, lut_m), l1 = lut_addr(a1f, lut_m), l2 = lut_addr(acc2, lut_m);
                                if (store_ok) {
                                    const float v0 = lds_f32(swap ? l2 : l0), v1 = lds_f32(l1 + 4 * LUT_STRIDE), v2 = lds_f32((swap ? l0 : l2) + 8 * LUT_STRIDE);
                                    if (BF16) {
                                        unsigned short* o = optr16 + roff;
                                        o[0] = bf16_bits(v0); o[1] = bf16_bits(v1); o[2] = bf16_bits(v2);
                                    } else {
                                        float* o = optr + roff;
                                        stg_out<0>(o, v0); stg_out<1>(o + plane, v1); stg_out<2>(o + 2 * plane, v2);
                                    }
                                }
                                roff += BF16 ? 3u * (unsigned)T : (unsigned)T;
                            }
                            acc01 = fprod2(pack2(bb[k], bb[k]), h01[k], nz2);
                            acc2 = __fmul_rn(bb[k], h2[k]);
                        }
                    }
                }
            }
        } else if (cls == 4) {
            // ---------------- class 4: area, 4 .. 6 taps per axis, source rows in order ----------------
            const int xn = __float_as_int(xd.w) >> 24;
            const int nt = warp_max_i32(xn);                      // uniform: taps evaluated per source row
            float w[6], cc[6];
#pragma unroll
            for (int k = 0; k < 6; ++k) {
                w[k] = (k == 0) ? xd.x : ((k < xn - 1) ? xd.y : ((k == xn - 1) ? xd.z : 0.f));
                if (!active) w[k] = 0.f;
                cc[k] = __fmul_rn(w[k], -8388608.0f);
                if (!active && k == 0) cc[k] = 256.f;             // beside the image: every row sums to 256 -> LUT entry 256 = fill
                asm volatile("" : "+f"(cc[k]));
            }
            u64 acc01 = 0ull;
            float acc2 = 0.f;
            unsigned roff = 0;
            int yout = 0;
            auto vstep = [&](float ba, float bb, u64 h01, float h2) {
                const float wa = fabsf(ba);
                acc01 = fadd2(acc01, fprod2(pack2(wa, wa), h01, nz2));
                acc2 = __fadd_rn(acc2, __fmul_rn(wa, h2));
                if (__float_as_int(ba) < 0) {                       // output row complete
                    float a0f, a1f;
                    unpack2(acc01, a0f, a1f);
                    if (OUT_U8) {
                        if (active) out.px(roi, dy0 + yout, x, round_u8(a0f), round_u8(a1f), round_u8(acc2));
                        else if (x < T) out.pad(roi, dy0 + yout, x);
                        ++yout;
                    } else {
                        const unsigned l0 = lut_addr(a0f, lut_m), l1 = lut_addr(a1f, lut_m), l2 = lut_addr(acc2, lut_m);
                        if (store_ok) {
                            const float v0 = lds_f32(swap ? l2 : l0), v1 = lds_f32(l1 + 4 * LUT_STRIDE), v2 = lds_f32((swap ? l0 : l2) + 8 * LUT_STRIDE);
                            if (BF16) {
                                unsigned short* o = optr16 + roff;
                                o[0] = bf16_bits(v0); o[1] = bf16_bits(v1); o[2] = bf16_bits(v2);
                            } else {
                                float* o = optr + roff;
                                stg_out<0>(o, v0); stg_out<1>(o + plane, v1); stg_out<2>(o + 2 * plane, v2);
                            }
                        }
                        roff += BF16 ? 3u * (unsigned)T : (unsigned)T;
                    }
                    acc01 = fprod2(pack2(bb, bb), h01, nz2);
                    acc2 = __fmul_rn(bb, h2);
                }
            };
            auto strip = [&](auto nt_c) {
                constexpr int NT = decltype(nt_c)::value;
                for (int c = 0; c < nchunks; ++c, ++cg) {
                    const unsigned j = cg % CTA_NSLOT;
                    CTA_LOCKSTEP(); mbar_wait(full_s + 8 * j, (cg / CTA_NSLOT) & 1);
                    const unsigned rbase = ring_s + j * slot_bytes + colc4;
                    const unsigned rec = dsc + 8u * (unsigned)(c * rps);
                    int r = 0;
                    for (; r + 1 < rps; r += 2) {                   // two rows at a time: their loads and conversions interleave
                        const float4 d = make_float4(lds_f32(rec + 8u * r), lds_f32(rec + 8u * r + 4), lds_f32(rec + 8u * r + 8), lds_f32(rec + 8u * r + 12));
                        u64 ha01, hb01;
                        float ha2, hb2;
                        h_area_n<NT>(rbase + (unsigned)(r * pitch), shc, w, cc, ha01, ha2);
                        h_area_n<NT>(rbase + (unsigned)((r + 1) * pitch), shc, w, cc, hb01, hb2);
                        vstep(d.x, d.y, ha01, ha2);
                        vstep(d.z, d.w, hb01, hb2);
                    }
                    if (r < rps) {
                        const float ba = lds_f32(rec + 8u * r), bb = lds_f32(rec + 8u * r + 4);
                        u64 ha01;
                        float ha2;
                        h_area_n<NT>(rbase + (unsigned)(r * pitch), shc, w, cc, ha01, ha2);
                        vstep(ba, bb, ha01, ha2);
                    }
                    __syncwarp();                               // every lane has read slot j
                    if (lane == 0) mbar_arrive(empty_s + 8 * j);
                }
            };
            if (nt <= 4) strip(std::integral_constant<int, 4>{});
            else if (nt == 5) strip(std::integral_constant<int, 5>{});
            else strip(std::integral_constant<int, 6>{});
        } else {
            // ---------------- class 3: fixed-point bilinear ----------------
            const int xw0 = __float_as_int(xd.x), xw1 = __float_as_int(xd.y);
            const unsigned lut_s = smem_s;
            int rowA = -1, rowB = -1;
            int HA[3] = {0, 0, 0}, HB[3] = {0, 0, 0};
            int y = 0;
            float4 d = lds_f4(dsc);
            for (int c = 0; c < nchunks; ++c, ++cg) {
                const unsigned j = cg % CTA_NSLOT;
                CTA_LOCKSTEP(); mbar_wait(full_s + 8 * j, (cg / CTA_NSLOT) & 1);
                const int last_row = c * CTA_ROWS + CTA_ROWS - 1;
                const unsigned sbase = ring_s + j * slot_bytes + colc4 - (unsigned)(c * CTA_ROWS * pitch);     // source row r at sbase + r * pitch
                while (y < new_h) {
                    // (b * H) >> 16 as the high word of (b << 16) * H: one IMAD.HI instead of a multiply and a shift (0 <= b <= 2048, 0 <= H < 2^15)
                    const unsigned b0 = (unsigned)__float_as_int(d.x) << 16, b1 = (unsigned)__float_as_int(d.y) << 16;
                    const int sy0 = __float_as_int(d.w), sy1 = __float_as_int(d.z);
                    if (sy1 > last_row) break;
                    ++y;
                    d = lds_f4(dsc + 16u * (unsigned)y);              // next row's descriptor behind this row's arithmetic
                    if (active) {
                        if (sy0 != rowA) {
                            if (sy0 == rowB) { HA[0] = HB[0]; HA[1] = HB[1]; HA[2] = HB[2]; }
                            else h_lin(sbase + (unsigned)(sy0 * pitch), shc, xw0, xw1, HA);
                            rowA = sy0;
                        }
                        if (sy1 != rowB) {
                            if (sy1 == rowA) { HB[0] = HA[0]; HB[1] = HA[1]; HB[2] = HA[2]; }
                            else h_lin(sbase + (unsigned)(sy1 * pitch), shc, xw0, xw1, HB);
                            rowB = sy1;
                        }
                        unsigned o[3];                      // 4 * value + 2 low bits
#pragma unroll
                        for (int k = 0; k < 3; ++k) o[k] = __umulhi(b1, (unsigned)HB[k]) + (__umulhi(b0, (unsigned)HA[k]) + 2u);
                        if (OUT_U8) {
                            out.px(roi, dy0 + y - 1, x, (int)(o[0] >> 2) & 255, (int)(o[1] >> 2) & 255, (int)(o[2] >> 2) & 255);
                        } else {
                            const float v0 = lds_f32(lut_s + ((swap ? o[2] : o[0]) & 0x3fcu)), v1 = lds_f32(lut_s + 4 * LUT_STRIDE + (o[1] & 0x3fcu)),
                                        v2 = lds_f32(lut_s + 8 * LUT_STRIDE + ((swap ? o[0] : o[2]) & 0x3fcu));
                            if (BF16) {
                                optr16[0] = bf16_bits(v0); optr16[1] = bf16_bits(v1); optr16[2] = bf16_bits(v2);
                                optr16 += 3 * T;
                            } else {
                                stg_out<0>(optr, v0); stg_out<1>(optr + plane, v1); stg_out<2>(optr + 2 * plane, v2);
                                optr += T;
                            }
                        }
                    } else if (x < T) {
                        if (OUT_U8) out.pad(roi, dy0 + y - 1, x);
                        else if (BF16) { optr16[0] = padh0; optr16[1] = padh1; optr16[2] = padh2; optr16 += 3 * T; }
                        else { optr[0] = out.padf[0]; optr[plane] = out.padf[1]; optr[2 * plane] = out.padf[2]; optr += T; }
                    }
                }
                __syncwarp();                               // every lane is done with slot j
                if (lane == 0) mbar_arrive(empty_s + 8 * j);
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(hempty_s + 8 * hb);             // done with this crop's descriptor block
    }
}

// ------------------------------------------------------------------------------------------------------
// generic kernel: any regime / scale, persistent CTAs over (ROI, band) items of the generic list
// ------------------------------------------------------------------------------------------------------
template <bool OUT_U8, bool BF16>
__device__ void crop_generic_band(unsigned char* raw, YDesc* yd, const Out<OUT_U8, BF16>& out, const RoiGeom& g, int roi, int band,
                                  const uint8_t* __restrict__ images, int B, int H, int W, int T, int tid, int nth) {
    const int x = tid;
    const bool incol = x < T;
    const int row0 = band * CROP_BAND, row1 = min(T, row0 + CROP_BAND);
    const int regime = g.regime;
    const int dy0 = g.dy, new_h = g.new_h, new_w = g.new_w, dx0 = g.dx;
    const int ya = max(row0, dy0) - dy0, yb = min(row1, dy0 + new_h) - dy0;
    out.pad_rows(roi, row0, min(row1, dy0), tid, nth);
    out.pad_rows(roi, max(row0, dy0 + new_h), row1, tid, nth);
    if (ya >= yb) return;                                            // uniform

    const int w = g.w, h = g.h;
    if (tid < yb - ya) {
        YDesc d;
        const int yr = ya + tid;
        if (regime == 1) {
            area_taps(yr, g.scale_y, h, d.start, d.n, d.bf, d.bm, d.bl, d.flags);
        } else if (regime == 2) {
            d.start = yr * g.isy; d.n = g.isy; d.bf = d.bm = d.bl = 1.f; d.flags = 0;
        } else {
            int s0, b0, b1, edge;
            linear_coef(yr, g.scale_y, g.inv_y, h, s0, b0, b1, edge);
            d.start = s0; d.n = min(s0 + 1, h - 1);
            d.bf = __int_as_float(b0); d.bm = __int_as_float(b1); d.bl = 0.f; d.flags = 0;
        }
        yd[tid] = d;
    }
    const int xr = x - dx0;
    const bool active = incol && xr >= 0 && xr < new_w;
    int xs = 0, xn = 0, xflags = 0, xw0 = 0, xw1 = 0, xedge = 0;
    float af = 0.f, am = 0.f, al = 0.f;
    if (active) {
        if (regime == 1) area_taps(xr, g.scale_x, w, xs, xn, af, am, al, xflags);
        else if (regime == 2) { xs = xr * g.isx; xn = g.isx; }
        else linear_coef(xr, g.scale_x, g.inv_x, w, xs, xw0, xw1, xedge);
    }
    __syncthreads();

    const int s_lo = yd[0].start;
    const int s_hi = (regime == 3) ? yd[yb - ya - 1].n : (yd[yb - ya - 1].start + yd[yb - ya - 1].n - 1);
    const int pitch = ((3 * w + 15 + 15) / 16) * 16;
    const int rows_fit = CROP_RAW_BYTES / pitch;
    const unsigned long long src0 = g.src;
    const unsigned long long rowstride = (unsigned long long)W * 3ull;
    const unsigned long long img_end = (unsigned long long)(uintptr_t)images + (unsigned long long)B * H * rowstride;
    const int nvec = pitch >> 4;
    const int lane = tid & 31, wid = tid >> 5, nwarps = nth >> 5;

    int yr = ya, t = 0;                      // streaming state, uniform across the CTA except for x
    float acc[3] = {0.f, 0.f, 0.f};
    int iacc[3] = {0, 0, 0};
    int cacheA = -1, cacheB = -1;
    float hA[3] = {0.f, 0.f, 0.f};
    int HA[3] = {0, 0, 0}, HB[3] = {0, 0, 0};

    int chunk = s_lo;
    while (yr < yb) {
        const int rows = min(rows_fit, s_hi - chunk + 1);
        __syncthreads();
        for (int r = wid; r < rows; r += nwarps) {
            const unsigned long long ga = src0 + (unsigned long long)(chunk + r) * rowstride;
            const unsigned long long al16 = ga & ~15ull;
            const int need = (int)(ga - al16) + 3 * w;
            for (int v = lane; v < nvec; v += 32) {
                if (v * 16 >= need) break;
                const unsigned long long a = al16 + (unsigned long long)v * 16ull;
                uint4 q;
                if (a + 16ull <= img_end) {
                    q = ld_nc_v4((const void*)(uintptr_t)a);
                } else {
                    unsigned int tmp[4] = {0u, 0u, 0u, 0u};
                    for (int b = 0; b < 16; ++b)
                        if (a + b < img_end) tmp[b >> 2] |= (unsigned int)(*(const uint8_t*)(uintptr_t)(a + b)) << (8 * (b & 3));
                    q = make_uint4(tmp[0], tmp[1], tmp[2], tmp[3]);
                }
                *reinterpret_cast<uint4*>(raw + (size_t)r * pitch + (size_t)v * 16) = q;
            }
        }
        __syncthreads();
        const int chunk_end = chunk + rows;
        auto rowptr = [&](int sy) -> const uint8_t* {
            const unsigned long long ga = src0 + (unsigned long long)sy * rowstride;
            return raw + (size_t)(sy - chunk) * pitch + (int)(ga & 15ull);
        };

        if (regime == 3) {
            while (yr < yb) {
                const YDesc d = yd[yr - ya];
                const int sy0 = d.start, sy1 = d.n;
                if (sy1 >= chunk_end) break;
                if (active) {
                    if (cacheA != sy0) {
                        if (cacheB == sy0) { HA[0] = HB[0]; HA[1] = HB[1]; HA[2] = HB[2]; }
                        else {
                            const uint8_t* p = rowptr(sy0) + 3 * xs;
                            for (int c = 0; c < 3; ++c) HA[c] = xedge ? (int)p[c] * 2048 : (int)p[c] * xw0 + (int)p[3 + c] * xw1;
                        }
                        cacheA = sy0;
                    }
                    if (cacheB != sy1) {
                        if (sy1 == sy0) { HB[0] = HA[0]; HB[1] = HA[1]; HB[2] = HA[2]; }
                        else {
                            const uint8_t* p = rowptr(sy1) + 3 * xs;
                            for (int c = 0; c < 3; ++c) HB[c] = xedge ? (int)p[c] * 2048 : (int)p[c] * xw0 + (int)p[3 + c] * xw1;
                        }
                        cacheB = sy1;
                    }
                    const int b0 = __float_as_int(d.bf), b1 = __float_as_int(d.bm);
                    int o[3];
                    for (int c = 0; c < 3; ++c)
                        o[c] = ((((b0 * (HA[c] >> 4)) >> 16) + ((b1 * (HB[c] >> 4)) >> 16) + 2) >> 2) & 255;
                    out.px(roi, dy0 + yr, x, o[0], o[1], o[2]);
                } else if (incol) {
                    out.pad(roi, dy0 + yr, x);
                }
                ++yr;
            }
            if (yr < yb) chunk = yd[yr - ya].start;
        } else {
            while (yr < yb) {
                const YDesc d = yd[yr - ya];
                const int sy = d.start + t;
                if (sy >= chunk_end) break;
                if (active) {
                    const uint8_t* p = rowptr(sy) + 3 * xs;
                    if (regime == 1) {
                        if (cacheA != sy) {
                            float s0 = 0.f, s1 = 0.f, s2 = 0.f;
                            for (int k = 0; k < xn; ++k) {
                                const float a = (k == 0 && (xflags & 1)) ? af : ((k == xn - 1 && (xflags & 2)) ? al : am);
                                s0 = __fadd_rn(s0, __fmul_rn((float)p[3 * k + 0], a));
                                s1 = __fadd_rn(s1, __fmul_rn((float)p[3 * k + 1], a));
                                s2 = __fadd_rn(s2, __fmul_rn((float)p[3 * k + 2], a));
                            }
                            hA[0] = s0; hA[1] = s1; hA[2] = s2;
                            cacheA = sy;
                        }
                        const float beta = (t == 0 && (d.flags & 1)) ? d.bf : ((t == d.n - 1 && (d.flags & 2)) ? d.bl : d.bm);
                        for (int c = 0; c < 3; ++c) {
                            const float term = __fmul_rn(beta, hA[c]);
                            acc[c] = (t == 0) ? term : __fadd_rn(acc[c], term);
                        }
                    } else {
                        if (t == 0) { iacc[0] = 0; iacc[1] = 0; iacc[2] = 0; }
                        for (int k = 0; k < xn; ++k) {
                            iacc[0] += p[3 * k + 0]; iacc[1] += p[3 * k + 1]; iacc[2] += p[3 * k + 2];
                        }
                    }
                }
                ++t;
                if (t == d.n) {
                    if (active) {
                        int o[3];
                        if (regime == 1) {
                            for (int c = 0; c < 3; ++c) o[c] = min(255, max(0, __float2int_rn(acc[c])));
                        } else if (g.isx == 2 && g.isy == 2) {
                            for (int c = 0; c < 3; ++c) o[c] = (iacc[c] + 2) >> 2;
                        } else {
                            const float sc = __fdiv_rn(1.f, (float)(g.isx * g.isy));
                            for (int c = 0; c < 3; ++c) o[c] = min(255, max(0, __float2int_rn(__fmul_rn(__int2float_rn(iacc[c]), sc))));
                        }
                        out.px(roi, dy0 + yr, x, o[0], o[1], o[2]);
                    } else if (incol) {
                        out.pad(roi, dy0 + yr, x);
                    }
                    ++yr; t = 0;
                }
            }
            if (yr < yb) chunk = yd[yr - ya].start + t;
        }
    }
}

template <bool OUT_U8, int NTH, bool BF16 = false>
__global__ void __launch_bounds__(NTH)
bpc_crop_generic_kernel(const uint8_t* __restrict__ images, int B, int H, int W, const RoiGeom* __restrict__ geom,
                        const int32_t* __restrict__ glist, const int32_t* __restrict__ gcount, int T, int nbands,
                        uchar4 fill, int swap_rb, const float* __restrict__ lut_g, float* __restrict__ outf, uint8_t* __restrict__ outb) {
    extern __shared__ __align__(16) unsigned char raw[];
    __shared__ RoiGeom g;
    __shared__ YDesc yd[CROP_BAND];
    __shared__ float lut[OUT_U8 ? 1 : 3 * LUT_STRIDE];
    const int tid = threadIdx.x, nth = blockDim.x;
    const long long items = (long long)(*gcount) * nbands;
    if (blockIdx.x >= items) return;
    if (!OUT_U8)
        for (int e = tid; e < 768; e += nth) lut[(e >> 8) * LUT_STRIDE + (e & 255)] = lut_g[e];
    __syncthreads();
    Out<OUT_U8, BF16> out;
    out_init(out, outf, outb, lut, T, swap_rb, fill);
    for (long long item = blockIdx.x; item < items; item += gridDim.x) {
        const int roi = glist[item / nbands], band = (int)(item % nbands);
        __syncthreads();
        if (tid == 0) g = geom[roi];
        __syncthreads();
        crop_generic_band<OUT_U8, BF16>(raw, yd, out, g, roi, band, images, B, H, W, T, tid, nth);
    }
}

__global__ void bpc_lut_kernel(float m0, float m1, float m2, float s0, float s1, float s2, float* __restrict__ lut) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= 768) return;
    const int c = t >> 8, v = t & 255;
    const float mean = c == 0 ? m0 : (c == 1 ? m1 : m2);
    const float sd = c == 0 ? s0 : (c == 1 ? s1 : s2);
    // to_tensor: float(v) / 255 ; normalize: (x - mean) / std -- float32, true divisions
    lut[t] = __fdiv_rn(__fsub_rn(__fdiv_rn((float)v, 255.f), mean), sd);
}

// Tensor maps of the image pool for every staging pitch (host side, cached for the last pool seen).  The driver entry
// point is fetched through the runtime, so the library still links against cudart only.
static int tensor_maps(const uint8_t* images, int B, int H, int W, bool aligned, TmapSet* out) {
    typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                 const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static std::mutex mu;
    static TmapSet cached;
    static const uint8_t* k_images = nullptr;
    static int k_B = 0, k_H = 0, k_W = 0, k_dev = -1;
    static EncodeFn encode = nullptr;
    std::lock_guard<std::mutex> lock(mu);
    if (!aligned) { memset(out, 0, sizeof(TmapSet)); return BPC_OK; }      // the 1-D bulk-copy variants ignore the maps
    int dev = 0;
    cudaGetDevice(&dev);
    if (images == k_images && B == k_B && H == k_H && W == k_W && dev == k_dev) { *out = cached; return BPC_OK; }   // copied under the lock
    if (!encode) {
        cudaDriverEntryPointQueryResult q;
        void* fnp = nullptr;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &q);
        if (e != cudaSuccess) return (int)e;
        if (q != cudaDriverEntryPointSuccess || !fnp) return (int)cudaErrorNotSupported;
        encode = (EncodeFn)fnp;
    }
    const cuuint64_t gdim[2] = {(cuuint64_t)W * 3 / 4, (cuuint64_t)B * H};
    const cuuint64_t gstride[1] = {(cuuint64_t)W * 3};
    const cuuint32_t estr[2] = {1, 1};
    for (int i = 0; i < N_TMAPS; ++i) {
        const int pitch = i < 13 ? 64 + 32 * i : 512 + 64 * (i - 13);
        const cuuint32_t box[2] = {(cuuint32_t)pitch / 4, (cuuint32_t)tmap_rows(pitch)};
        const CUresult r = encode(&cached.m[i], CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, (void*)images, gdim, gstride, box, estr,
                                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, BPC_L2_PROMO,
                                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { k_images = nullptr; return (int)cudaErrorInvalidValue; }
    }
    for (int i = 0; i < N_TMAPS8; ++i) {
        const int pitch = 64 + 32 * i;
        const cuuint32_t box[2] = {(cuuint32_t)pitch / 4, (cuuint32_t)STREAM_ROWS};
        const CUresult r = encode(&cached.m8[i], CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, (void*)images, gdim, gstride, box, estr,
                                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, BPC_L2_PROMO,
                                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { k_images = nullptr; return (int)cudaErrorInvalidValue; }
    }
    k_images = images; k_B = B; k_H = H; k_W = W; k_dev = dev;
    *out = cached;
    return BPC_OK;
}

// Tensor maps of the CTA kernel: the pool as [B*H rows][W*3/8 uint64], box = {64 i bytes, CTA_ROWS rows}, i = 1 .. CTA_NMAPS.
static int cta_tensor_maps(const uint8_t* images, int B, int H, int W, CtaMaps* out) {
    typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                 const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static std::mutex mu;
    static CtaMaps cached;
    static const uint8_t* k_images = nullptr;
    static int k_B = 0, k_H = 0, k_W = 0, k_dev = -1;
    static EncodeFn encode = nullptr;
    std::lock_guard<std::mutex> lock(mu);
    int dev = 0;
    cudaGetDevice(&dev);
    if (images == k_images && B == k_B && H == k_H && W == k_W && dev == k_dev) { *out = cached; return BPC_OK; }
    if (!encode) {
        cudaDriverEntryPointQueryResult q;
        void* fnp = nullptr;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &q);
        if (e != cudaSuccess) return (int)e;
        if (q != cudaDriverEntryPointSuccess || !fnp) return (int)cudaErrorNotSupported;
        encode = (EncodeFn)fnp;
    }
    const cuuint64_t gdim[2] = {(cuuint64_t)W * 3 / 8, (cuuint64_t)B * H};
    const cuuint64_t gstride[1] = {(cuuint64_t)W * 3};
    const cuuint32_t estr[2] = {1, 1};
    for (int i = 0; i < CTA_NMAPS; ++i) {
        const cuuint32_t box[2] = {(cuuint32_t)(8 * (i + 1)), (cuuint32_t)CTA_ROWS};
        const CUresult r = encode(&cached.m[i], CU_TENSOR_MAP_DATA_TYPE_UINT64, 2, (void*)images, gdim, gstride, box, estr,
                                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, BPC_L2_PROMO,
                                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { k_images = nullptr; return (int)cudaErrorInvalidValue; }
    }
    k_images = images; k_B = B; k_H = H; k_W = W; k_dev = dev;
    *out = cached;
    return BPC_OK;
}

// workspace: geom[R] | xdesc[R][ds] | ydesc[R][ydesc_stride(T)] | counters[16] | glist[R] | list1[R]    (ds = desc_stride(T), float4 records)
static size_t ws_off_xdesc(int R) { return (((size_t)R * sizeof(RoiGeom)) + 15) & ~(size_t)15; }
static size_t ws_off_ydesc(int R, int T) { return ws_off_xdesc(R) + (size_t)R * desc_stride(T) * sizeof(float4); }
static size_t ws_off_count(int R, int T) { return ws_off_ydesc(R, T) + (size_t)R * ydesc_stride(T) * sizeof(float4); }
static size_t crop_workspace_bytes(int R, int T) { return ws_off_count(R, T) + 64 + 2 * (size_t)R * sizeof(int32_t) + 64; }

template <bool OUT_U8, bool BF16 = false>
static int launch_crop(const uint8_t* images, int B, int H, int W, const int32_t* rois, int R, const int32_t* n_rois_dev,
                       int roi_first, int T, const uint8_t* fill, int swap_rb, const float* lut, float* outf, uint8_t* outb,
                       int32_t* status, void* workspace, size_t workspace_bytes, void* stream) {
    if (R < 0 || B < 1 || H < 1 || W < 1 || T < 1 || T > BPC_MAX_TARGET || !fill) return BPC_EINVAL;
    if (R > 0 && (!images || !rois || !workspace || (!OUT_U8 && (!lut || !outf)) || (OUT_U8 && !outb))) return BPC_EINVAL;
    if (((uintptr_t)images & 15) != 0 || ((uintptr_t)workspace & 15) != 0) return BPC_EALIGN;
    if (R == 0) return BPC_OK;
    if (workspace_bytes < crop_workspace_bytes(R, T)) return BPC_EWORKSPACE;
    const int nslot = (T + 31) / 32 + 1;
    if ((long long)R * nslot > 0x7fffffffLL) return BPC_ETOOBIG;
    cudaStream_t st = (cudaStream_t)stream;
    unsigned char* wsb = (unsigned char*)workspace;
    RoiGeom* geom = (RoiGeom*)wsb;
    float4* xdesc = (float4*)(wsb + ws_off_xdesc(R));
    float4* ydesc = (float4*)(wsb + ws_off_ydesc(R, T));
    int32_t* gcount = (int32_t*)(wsb + ws_off_count(R, T));   // [0] generic-list length, [4] warp-item counter, [8] class-1 list length, [12] its work counter
    int32_t* wcount = gcount + 4;
    int32_t* glist = gcount + 16;
    int32_t* list1 = glist + R;
    cudaError_t e = cudaMemsetAsync(gcount, 0, 64, st);
    if (e != cudaSuccess) return (int)e;
    // 2-D TMA staging needs a 16-byte image pitch; rows narrower than the widest box keep the 1-D path
    const bool aligned = ((long long)W * 3) % 16 == 0 && (long long)W * 3 >= TMAP_MAX_PITCH;
    // class 1 through the warp-specialised CTA kernel: 2-D TMA staging, at most eight strips, full-width boxes inside the pool rows
    const bool use_cta = aligned && T <= CTA_MAX_T && (long long)W * 3 >= cta_pitch_max(T);
    if (BF16 && !use_cta) return BPC_EUNSUPPORTED;       // the bfloat16 output exists on the CTA kernel's path only
    bpc_crop_prep_kernel<<<R, 256, 0, st>>>(images, B, H, W, rois, R, n_rois_dev, roi_first, T, (aligned ? 1 : 0) | (use_cta ? 2 : 0), geom, xdesc, ydesc,
                                            glist, gcount, status, use_cta ? list1 : nullptr);
    BPC_LAUNCH_CHECK();
    const uchar4 f4 = make_uchar4(fill[0], fill[1], fill[2], 0);
    if (use_cta) {
        typedef void (*CtaFn)(const uint8_t*, int, int, int, const RoiGeom*, const float4*, const float4*, const int32_t*, int32_t*, int,
                              uchar4, int, const float*, float*, uint8_t*, const CtaMaps);
        CtaMaps cmaps;
        const int terr = cta_tensor_maps(images, B, H, W, &cmaps);
        if (terr != BPC_OK) return terr;
        const bool sw = swap_rb != 0;
        CtaFn fn;
        if (OUT_U8) fn = bpc_crop_cta_kernel<OUT_U8, 0, false>;
        else if (BF16) fn = sw ? bpc_crop_cta_kernel<OUT_U8, 0, true, BF16> : bpc_crop_cta_kernel<OUT_U8, 0, false, BF16>;
        else if (T == 224) fn = sw ? bpc_crop_cta_kernel<OUT_U8, 224, true> : bpc_crop_cta_kernel<OUT_U8, 224, false>;
        else if (T == 256) fn = sw ? bpc_crop_cta_kernel<OUT_U8, 256, true> : bpc_crop_cta_kernel<OUT_U8, 256, false>;
        else fn = sw ? bpc_crop_cta_kernel<OUT_U8, 0, true> : bpc_crop_cta_kernel<OUT_U8, 0, false>;
        const int nslot = cta_nslot((!OUT_U8 && !BF16 && T == 224) ? 224 : 0);      // as the instantiation picked above
        const int smem_bytes = cta_smem_bytes(T, nslot);
        e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, cta_smem_bytes(T == 224 ? 224 : CTA_MAX_T, nslot));
        if (e != cudaSuccess) return (int)e;
        const int threads = 32 * ((T + 31) / 32 + 1);
        int dev = 0, sms = 148, per_sm = 3;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, threads, smem_bytes) != cudaSuccess || per_sm < 1) per_sm = 1;
        const long long slots = (long long)sms * per_sm;
        const int grid = (int)((long long)R < slots ? R : slots);
        fn<<<grid, threads, smem_bytes, st>>>(images, B, H, W, geom, xdesc, ydesc, list1, gcount, T, f4, swap_rb, lut, outf, outb, cmaps);
        BPC_LAUNCH_CHECK();
    }
    if (!BF16) {
        typedef void (*WarpFn)(const uint8_t*, int, int, int, const RoiGeom*, const float4*, const float4*, int32_t*, int, int, int,
                               uchar4, int, const float*, float*, uint8_t*, const TmapSet, const int32_t*);
        TmapSet tmaps;
        const int terr = tensor_maps(images, B, H, W, aligned, &tmaps);
        if (terr != BPC_OK) return terr;
        WarpFn fn;
        const bool sw = swap_rb != 0;
#define BPC_PICK(TTV)                                                                                                  \
    (aligned ? (sw ? bpc_crop_warp_kernel<OUT_U8, TTV, true, true> : bpc_crop_warp_kernel<OUT_U8, TTV, true, false>)   \
             : (sw ? bpc_crop_warp_kernel<OUT_U8, TTV, false, true> : bpc_crop_warp_kernel<OUT_U8, TTV, false, false>))
        if (OUT_U8) fn = aligned ? bpc_crop_warp_kernel<OUT_U8, 0, true, false> : bpc_crop_warp_kernel<OUT_U8, 0, false, false>;
        else if (T == 224) fn = BPC_PICK(224);
        else if (T == 256) fn = BPC_PICK(256);
        else fn = BPC_PICK(0);
#undef BPC_PICK
        // the same constant on every call: idempotent, so concurrent callers cannot interleave set(small) / launch(large)
        e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, WARPK_SMEM);
        if (e != cudaSuccess) return (int)e;
        const int cta_threads = 256;
        const long long nitems = (long long)R * nslot;
        const long long want = (nitems + WARPK_WARPS - 1) / WARPK_WARPS;
        int dev = 0, sms = 148, per_sm = 3;                     // persistent grid: every resident CTA slot, no more
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, cta_threads, WARPK_SMEM) != cudaSuccess || per_sm < 1) per_sm = 1;
        const long long slots = (long long)sms * per_sm;
        const int grid = (int)(want < slots ? want : slots);
        fn<<<grid, cta_threads, WARPK_SMEM, st>>>(images, B, H, W, geom, xdesc, ydesc, wcount, R, T, nslot, f4, swap_rb, lut, outf, outb, tmaps, use_cta ? list1 : nullptr);
        BPC_LAUNCH_CHECK();
    }
    const int nbands = (T + CROP_BAND - 1) / CROP_BAND;
    const long long max_items = (long long)R * nbands;
    const int grid = (int)(max_items < 148 * 2 ? max_items : 148 * 2);
    const int threads = ((T + 31) / 32) * 32;                   // one thread per output column
    if (threads <= 256) {
        e = cudaFuncSetAttribute(bpc_crop_generic_kernel<OUT_U8, 256, BF16>, cudaFuncAttributeMaxDynamicSharedMemorySize, CROP_RAW_BYTES);
        if (e != cudaSuccess) return (int)e;
        bpc_crop_generic_kernel<OUT_U8, 256, BF16><<<grid, threads, CROP_RAW_BYTES, st>>>(images, B, H, W, geom, glist, gcount, T, nbands, f4,
                                                                                   swap_rb, lut, outf, outb);
    } else {
        e = cudaFuncSetAttribute(bpc_crop_generic_kernel<OUT_U8, 1024, BF16>, cudaFuncAttributeMaxDynamicSharedMemorySize, CROP_RAW_BYTES);
        if (e != cudaSuccess) return (int)e;
        bpc_crop_generic_kernel<OUT_U8, 1024, BF16><<<grid, threads, CROP_RAW_BYTES, st>>>(images, B, H, W, geom, glist, gcount, T, nbands, f4,
                                                                                    swap_rb, lut, outf, outb);
    }
    BPC_LAUNCH_CHECK();
    return BPC_OK;
}

}  // namespace bpc

using namespace bpc;

extern "C" size_t bpc_roi_crop_workspace_bytes(int R, int T) { return (R < 0 || T < 1 || T > BPC_MAX_TARGET) ? 0 : crop_workspace_bytes(R, T); }

extern "C" int bpc_roi_crop(const uint8_t* images, int B, int H, int W, const int32_t* rois, int R,
                            const int32_t* n_rois_dev, int roi_first, int T, const uint8_t* fill, int swap_rb,
                            const float* lut, float* out, int32_t* status, void* workspace, size_t workspace_bytes,
                            void* stream) {
    if (((uintptr_t)out & 15) != 0) return BPC_EALIGN;
    return launch_crop<false>(images, B, H, W, rois, R, n_rois_dev, roi_first, T, fill, swap_rb, lut, out, nullptr, status,
                              workspace, workspace_bytes, stream);
}

extern "C" int bpc_roi_crop_bf16(const uint8_t* images, int B, int H, int W, const int32_t* rois, int R,
                                 const int32_t* n_rois_dev, int roi_first, int T, const uint8_t* fill, int swap_rb,
                                 const float* lut, void* out, int32_t* status, void* workspace, size_t workspace_bytes,
                                 void* stream) {
    if (((uintptr_t)out & 15) != 0) return BPC_EALIGN;
    return launch_crop<false, true>(images, B, H, W, rois, R, n_rois_dev, roi_first, T, fill, swap_rb, lut, (float*)out, nullptr, status,
                                    workspace, workspace_bytes, stream);
}

extern "C" int bpc_roi_crop_u8(const uint8_t* images, int B, int H, int W, const int32_t* rois, int R,
                               const int32_t* n_rois_dev, int roi_first, int T, const uint8_t* fill, uint8_t* out,
                               int32_t* status, void* workspace, size_t workspace_bytes, void* stream) {
    return launch_crop<true>(images, B, H, W, rois, R, n_rois_dev, roi_first, T, fill, 0, nullptr, nullptr, out, status,
                             workspace, workspace_bytes, stream);
}

extern "C" int bpc_normalise_lut(const float* mean, const float* std_, float* lut, void* stream) {
    if (!mean || !std_ || !lut) return BPC_EINVAL;
    bpc_lut_kernel<<<3, 256, 0, (cudaStream_t)stream>>>(mean[0], mean[1], mean[2], std_[0], std_[1], std_[2], lut);
    BPC_LAUNCH_CHECK();
    return BPC_OK;
}
