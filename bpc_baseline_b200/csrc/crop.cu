// ROI crop -> aspect-preserving letterbox (cv2.resize INTER_AREA) -> colour order -> normalise.
//
// Replaces letterbox_preserving_aspect_ratio (bpc/utils/data_utils.py:34-44) and the inline transform
// of PoseEstimator._estimate_rotation (bpc/inference/process_pose.py:199-209).  The uint8 resize is
// bit-exact with OpenCV's INTER_AREA for 8UC3:
//   regime 1  both axes shrink           : float32 area taps, sequential mul/add (no FMA), cvRound
//   regime 2  both shrink, integer ratio : integer box sum; 2x2 -> (sum+2)>>2
//   regime 3  either axis grows          : 11-bit fixed-point bilinear with area-mode coordinates
//
// Mapping: one CTA per (ROI, band of BAND output rows); one thread per output column x, all three
// channels.  The source rows a band needs are staged in shared memory with 128-bit loads of the
// 16-byte-aligned superset of each row; each thread then streams down its column: the horizontal pass
// of a source row is computed once and reused by the (at most two) output rows that tap it.
// The dominant traffic is the float32 output (3*T*T*4 B per ROI): each warp stores 128 contiguous
// bytes per plane and row.
#include "common.cuh"

namespace bpc {

constexpr int CROP_BAND = 8;            // output rows per CTA
constexpr int CROP_RAW_BYTES = 40 * 1024;

struct RoiGeom {
    double scale_x, scale_y, inv_x, inv_y;
    unsigned long long src;            // byte address of (y1, x1) in its image
    int w, h, new_w, new_h, dx, dy;
    int regime;                        // 0 = rejected, 1 / 2 / 3 as above
    int isx, isy;
    int pitch;                         // shared-memory bytes per staged source row (multiple of 16)
    int rows_fit;
};

struct YDesc {                         // per output row of the band
    int start;                         // first source row tapped
    int n;                             // taps (regime 1/2) ; regime 3: second source row
    float bf, bm, bl;                  // regime 1 weights ; regime 3: b0, b1 as ints in bf/bm bits
    int flags;                         // bit0 has_first, bit1 has_last
};

__device__ __forceinline__ uint4 ld_nc_v4(const void* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}

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

template <bool OUT_U8>
__global__ void __launch_bounds__(256)
bpc_crop_kernel(const uint8_t* __restrict__ images, int B, int H, int W, const int32_t* __restrict__ rois, int R,
                const int32_t* __restrict__ n_rois_dev, int roi_first, int T, int nbands, uchar4 fill, int swap_rb,
                const float* __restrict__ lut_g, float* __restrict__ outf, uint8_t* __restrict__ outb,
                int32_t* __restrict__ status) {
    extern __shared__ __align__(16) unsigned char raw[];
    __shared__ RoiGeom g;
    __shared__ YDesc yd[CROP_BAND];
    __shared__ float lut[OUT_U8 ? 1 : 768];

    const int tid = threadIdx.x, nth = blockDim.x;
    const int roi = blockIdx.x / nbands, band = blockIdx.x - roi * nbands;
    if (roi >= R) return;
    if (n_rois_dev != nullptr && roi_first + roi >= *n_rois_dev) return;

    if (tid == 0) {
        const int32_t* r = rois + (size_t)roi * 5;
        const int img = r[0], x1 = r[1], y1 = r[2], x2 = r[3], y2 = r[4];
        const int w = x2 - x1, h = y2 - y1;
        g.regime = 0;
        g.w = w; g.h = h;
        if (img >= 0 && img < B && x1 >= 0 && y1 >= 0 && x2 <= W && y2 <= H && w > 0 && h > 0 && w <= BPC_MAX_ROI_WIDTH) {
            // letterbox geometry, data_utils.py:35-38,41-42 (Python round = half-to-even on the f64 product)
            const double scale = ddiv((double)T, (double)max(h, w));
            const int new_w = (int)__double2ll_rn(dmul((double)w, scale));
            const int new_h = (int)__double2ll_rn(dmul((double)h, scale));
            if (new_w >= 1 && new_h >= 1 && new_w <= T && new_h <= T) {
                g.new_w = new_w; g.new_h = new_h;
                g.dx = (T - new_w) / 2; g.dy = (T - new_h) / 2;
                g.inv_x = ddiv((double)new_w, (double)w);
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
                g.pitch = ((3 * w + 15 + 15) / 16) * 16;
                g.rows_fit = CROP_RAW_BYTES / g.pitch;
            }
        }
        if (band == 0 && status != nullptr) status[roi] = (g.regime == 0) ? 1 : 0;
    }
    if (!OUT_U8)
        for (int e = tid; e < 768; e += nth) lut[e] = lut_g[e];
    __syncthreads();

    const int x = tid;                                  // output column
    const int row0 = band * CROP_BAND, row1 = min(T, row0 + CROP_BAND);
    const int regime = g.regime;
    // per-plane padding values
    const uint8_t fillc[3] = {fill.x, fill.y, fill.z};
    float padf[3];
    if (!OUT_U8)
        for (int p = 0; p < 3; ++p) padf[p] = lut[p * 256 + fillc[swap_rb ? 2 - p : p]];

    auto store_px = [&](int y, int b0, int b1, int b2) {    // b* in source channel order
        if (x >= T) return;
        if (OUT_U8) {
            uint8_t* o = outb + (((size_t)roi * T + y) * T + x) * 3;
            o[0] = (uint8_t)b0; o[1] = (uint8_t)b1; o[2] = (uint8_t)b2;
        } else {
            const int s0 = swap_rb ? b2 : b0, s2 = swap_rb ? b0 : b2;
            float* o = outf + ((size_t)roi * 3 * T + y) * T + x;
            o[0] = lut[s0];
            o[(size_t)T * T] = lut[256 + b1];
            o[(size_t)2 * T * T] = lut[512 + s2];
        }
    };
    auto store_pad = [&](int y) {
        if (x >= T) return;
        if (OUT_U8) {
            uint8_t* o = outb + (((size_t)roi * T + y) * T + x) * 3;
            o[0] = fillc[0]; o[1] = fillc[1]; o[2] = fillc[2];
        } else {
            float* o = outf + ((size_t)roi * 3 * T + y) * T + x;
            o[0] = padf[0]; o[(size_t)T * T] = padf[1]; o[(size_t)2 * T * T] = padf[2];
        }
    };

    if (regime == 0) {
        for (int y = row0; y < row1; ++y) store_pad(y);
        return;
    }
    const int dy0 = g.dy, new_h = g.new_h, new_w = g.new_w, dx0 = g.dx;
    // rows of the band inside the resized image: [ya, yb) in resized coordinates
    const int ya = max(row0, dy0) - dy0, yb = min(row1, dy0 + new_h) - dy0;
    for (int y = row0; y < row1; ++y)
        if (y < dy0 || y >= dy0 + new_h) store_pad(y);
    if (ya >= yb) return;

    const int w = g.w, h = g.h;
    // ---- per-row descriptors ----------------------------------------------------------------------------
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
    // ---- per-thread column descriptor -------------------------------------------------------------------
    const int xr = x - dx0;
    const bool active = (x < T) && xr >= 0 && xr < new_w;
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
    const int pitch = g.pitch, rows_fit = g.rows_fit;
    const unsigned long long src0 = g.src;
    const unsigned long long rowstride = (unsigned long long)W * 3ull;
    const unsigned long long img_end = (unsigned long long)(uintptr_t)images + (unsigned long long)B * H * rowstride;
    const int nvec = pitch >> 4;
    const int lane = tid & 31, wid = tid >> 5, nwarps = nth >> 5;

    // streaming state (uniform across the CTA except for x)
    int yr = ya, t = 0;
    float acc[3] = {0.f, 0.f, 0.f};
    int iacc[3] = {0, 0, 0};
    int cacheA = -1, cacheB = -1;          // source rows held in hA / hB
    float hA[3] = {0.f, 0.f, 0.f};
    int HA[3] = {0, 0, 0}, HB[3] = {0, 0, 0};

    int chunk = s_lo;
    while (yr < yb) {
        const int rows = min(rows_fit, s_hi - chunk + 1);
        __syncthreads();
        for (int r = wid; r < rows; r += nwarps) {
            const unsigned long long ga = src0 + (unsigned long long)(chunk + r) * rowstride;
            const unsigned long long al16 = ga & ~15ull;
            const int need = (int)(ga - al16) + 3 * w;                 // bytes from the aligned start
            for (int v = lane; v < nvec; v += 32) {
                if (v * 16 >= need) break;
                const unsigned long long a = al16 + (unsigned long long)v * 16ull;
                uint4 q;
                if (a + 16ull <= img_end) {
                    q = ld_nc_v4((const void*)(uintptr_t)a);
                } else {                                                // last bytes of the image pool
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
                    store_px(dy0 + yr, o[0], o[1], o[2]);
                } else {
                    store_pad(dy0 + yr);
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
                        store_px(dy0 + yr, o[0], o[1], o[2]);
                    } else {
                        store_pad(dy0 + yr);
                    }
                    ++yr; t = 0;
                }
            }
            if (yr < yb) chunk = yd[yr - ya].start + t;
        }
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

template <bool OUT_U8>
static int launch_crop(const uint8_t* images, int B, int H, int W, const int32_t* rois, int R, const int32_t* n_rois_dev,
                       int roi_first, int T, const uint8_t* fill, int swap_rb, const float* lut, float* outf, uint8_t* outb,
                       int32_t* status, void* stream) {
    if (R < 0 || B < 1 || H < 1 || W < 1 || T < 1 || T > 256 || !fill) return BPC_EINVAL;
    if (R > 0 && (!images || !rois || (!OUT_U8 && (!lut || !outf)) || (OUT_U8 && !outb))) return BPC_EINVAL;
    if (((uintptr_t)images & 15) != 0) return BPC_EALIGN;
    if (R == 0) return BPC_OK;
    const int nbands = (T + CROP_BAND - 1) / CROP_BAND;
    if ((long long)R * nbands > 0x7fffffffLL) return BPC_ETOOBIG;
    const int threads = ((T + 31) / 32) * 32;
    static bool attr_set[2] = {false, false};
    if (!attr_set[OUT_U8]) {
        cudaError_t e = cudaFuncSetAttribute(bpc_crop_kernel<OUT_U8>, cudaFuncAttributeMaxDynamicSharedMemorySize, CROP_RAW_BYTES);
        if (e != cudaSuccess) return (int)e;
        attr_set[OUT_U8] = true;
    }
    bpc_crop_kernel<OUT_U8><<<R * nbands, threads, CROP_RAW_BYTES, (cudaStream_t)stream>>>(
        images, B, H, W, rois, R, n_rois_dev, roi_first, T, nbands, make_uchar4(fill[0], fill[1], fill[2], 0), swap_rb, lut, outf, outb, status);
    BPC_LAUNCH_CHECK();
    return BPC_OK;
}

}  // namespace bpc

using namespace bpc;

extern "C" int bpc_roi_crop(const uint8_t* images, int B, int H, int W, const int32_t* rois, int R,
                            const int32_t* n_rois_dev, int roi_first, int T, const uint8_t* fill, int swap_rb,
                            const float* lut, float* out, int32_t* status, void* stream) {
    if (((uintptr_t)out & 15) != 0) return BPC_EALIGN;
    return launch_crop<false>(images, B, H, W, rois, R, n_rois_dev, roi_first, T, fill, swap_rb, lut, out, nullptr, status, stream);
}

extern "C" int bpc_roi_crop_u8(const uint8_t* images, int B, int H, int W, const int32_t* rois, int R,
                               const int32_t* n_rois_dev, int roi_first, int T, const uint8_t* fill, uint8_t* out,
                               int32_t* status, void* stream) {
    return launch_crop<true>(images, B, H, W, rois, R, n_rois_dev, roi_first, T, fill, 0, nullptr, nullptr, out, status, stream);
}

extern "C" int bpc_normalise_lut(const float* mean, const float* std_, float* lut, void* stream) {
    if (!mean || !std_ || !lut) return BPC_EINVAL;
    bpc_lut_kernel<<<3, 256, 0, (cudaStream_t)stream>>>(mean[0], mean[1], mean[2], std_[0], std_[1], std_[2], lut);
    BPC_LAUNCH_CHECK();
    return BPC_OK;
}
