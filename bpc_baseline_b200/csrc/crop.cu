// ROI crop -> aspect-preserving letterbox (cv2.resize INTER_AREA) -> colour order -> normalise.
//
// Replaces letterbox_preserving_aspect_ratio (bpc/utils/data_utils.py:34-44) and the inline transform
// of PoseEstimator._estimate_rotation (bpc/inference/process_pose.py:199-209).  The uint8 resize is
// bit-exact with OpenCV's INTER_AREA for 8UC3:
//   regime 1  both axes shrink           : float32 area taps, sequential mul/add (no FMA), cvRound
//   regime 2  both shrink, integer ratio : integer box sum; 2x2 -> (sum+2)>>2
//   regime 3  either axis grows          : 11-bit fixed-point bilinear with area-mode coordinates
//
// Kernels of one call (shared pieces in crop_common.cuh):
//   bpc_crop_prep     one CTA per ROI: letterbox geometry (float64, as Python / OpenCV compute it), regime
//                     and class of the ROI, and -- for the fast classes -- the per-column and per-row tap
//                     descriptors (start index + float32 weights, absent taps = +0.0f; per-SOURCE-row records for
//                     the area classes), so that no float64 arithmetic is left in the hot kernels; sorts the ROIs
//                     into the lists of the kernels below;
//   bpc_crop_cta      (crop_cta.cu) THE hot kernel: one crop per CTA, a producer warp issuing one full-width 2-D TMA
//                     box of eight source rows per ring slot, consumer warps = 32-column strips.  Classes 1 (regime
//                     1, scale < 2, <= 3 taps per axis; scale exactly 1), 3 (regime 3) and 4 (regime 1 with 4..6 taps
//                     per axis, 2 <= scale <= 5) and rejected boxes, whenever the image pitch is a multiple of 16
//                     bytes and T <= 256;
//   bpc_crop_warp     round 1's kernel, the path for other pitches / narrow images / T > 256: persistent CTAs, the
//                     unit of work is (ROI, strip of 32 output columns), taken by ONE WARP from a global atomic
//                     counter, plus one "padding" item per ROI; per-warp TMA staging (2-D boxes of four / eight rows,
//                     or one 1-D bulk copy per row); walks its own list and exits at once when that is empty;
//   bpc_crop_generic  persistent CTAs over the (rare) remaining ROIs (class 2): integer ratios >= 2, scale > 5,
//                     source rows streamed through shared memory in chunks (boxes as large as the image).
// The dominant traffic is the float32 output (3*T*T*4 B per ROI): each warp stores 128 contiguous bytes per
// plane and row; padding rows are written with 16-byte stores.
#include <mutex>
#include <type_traits>

#include "crop_common.cuh"

namespace bpc {

#ifndef BPC_PREP_THREADS
#define BPC_PREP_THREADS 128
#endif
constexpr int PREP_THREADS = BPC_PREP_THREADS;      // threads of a prep CTA (one ROI each); 128: twice the CTAs in flight of 256 behind the serial geometry section (-33 us per 16 384 ROIs)

// Per-SOURCE-row records of an area resize with at most `maxtaps` taps per output row (collective over the PREP_THREADS threads of a prep
// CTA): record s = (ba, bb); |ba| = weight of source row s in the output row being accumulated, sign bit of ba set = that output
// row is complete after row s; bb = weight of row s as the first tap of the next output row (+0 if it has none there).  Rows
// [h, hpad) stay zero (no contribution, no completed row) so that whole ring slots run to completion.  *bad is set when the taps
// do not have this shape (then the ROI takes the generic path).
__device__ void src_row_records(float2* ysrc, int cap, const RoiGeom& g, int maxtaps, int tid, int* bad) {
    const int hpad = min(cap, ((g.h + STREAM_ROWS - 1) / STREAM_ROWS) * STREAM_ROWS + STREAM_ROWS);
    if (g.h + STREAM_ROWS > cap) *bad = 1;
    for (int s = tid; s < hpad; s += PREP_THREADS) ysrc[s] = make_float2(0.f, 0.f);
    __syncthreads();
    for (int d = tid; d < g.new_h && !*bad; d += PREP_THREADS) {
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
__global__ void __launch_bounds__(PREP_THREADS)
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
            for (int d = tid; d < nd; d += PREP_THREADS) {
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
        for (int d = tid; d < g.new_w; d += PREP_THREADS) {
            int xs, xn; float w0, w1, w2;
            area_taps3(d, g.scale_x, g.w, xs, xn, w0, w1, w2);
            xd[d] = make_float4(w0, w1, w2, __int_as_float(xs));
        }
        if (!(stream_ok & 1)) {
            for (int d = tid; d < g.new_h; d += PREP_THREADS) {
                int ys, yn; float b0, b1, b2;
                area_taps3(d, g.scale_y, g.h, ys, yn, b0, b1, b2);
                yd[d] = make_float4(b0, b1, b2, __int_as_float(ys | (yn << 24)));
            }
        } else {
            src_row_records(reinterpret_cast<float2*>(yd), 2 * ds + YSRC_PAD, g, 3, tid, &s_bad);
            if (s_bad && tid == 0) { g.cls = 2; }
        }
    } else if (cls == 3) {
        for (int d = tid; d < g.new_w; d += PREP_THREADS) {
            int xs, w0, w1, edge;
            linear_coef(d, g.scale_x, g.inv_x, g.w, xs, w0, w1, edge);
            if (edge) { w0 = 2048; w1 = 0; }                        // D = S[sx] * ONE beyond xmax
            xd[d] = make_float4(__int_as_float(w0), __int_as_float(w1), 0.f, __int_as_float(xs));
        }
        for (int d = tid; d < g.new_h; d += PREP_THREADS) {
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
// warp kernel: (ROI, 32-column strip) items, one warp each
// ------------------------------------------------------------------------------------------------------
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
    bpc_crop_prep_kernel<<<R, PREP_THREADS, 0, st>>>(images, B, H, W, rois, R, n_rois_dev, roi_first, T, (aligned ? 1 : 0) | (use_cta ? 2 : 0), geom, xdesc, ydesc,
                                            glist, gcount, status, use_cta ? list1 : nullptr);
    BPC_LAUNCH_CHECK();
    const uchar4 f4 = make_uchar4(fill[0], fill[1], fill[2], 0);
    if (use_cta) {
        const int rc = crop_cta_launch(OUT_U8 ? 1 : (BF16 ? 2 : 0), images, B, H, W, geom, xdesc, ydesc, list1, gcount, R, T, f4, swap_rb, lut,
                                       outf, outb, st);
        if (rc != BPC_OK) return rc;
    }
    if (!use_cta) {                                       // with the CTA kernel in use every fast-class crop is on ITS list
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
