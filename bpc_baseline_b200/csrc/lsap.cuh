// Exact rectangular linear-sum assignment, one CTA per problem, state O(rows) + one bit per column.
//
// Replaces scipy.optimize.linear_sum_assignment as called by match_objects,
// bpc/inference/epipolar_matching.py:104-116, and reproduces WHICH optimum SciPy returns when optima
// are not unique (Crouse's shortest augmenting path as implemented in SciPy's rectangular_lsap):
//   * rows are augmented in ascending order (after the transposition SciPy applies when there are more
//     rows than columns, rows = min(N*M, P) side);
//   * reduced cost r = ((minVal + C[i,j]) - u[i]) - v[j], float64, left to right;
//   * among equal-lowest columns an unassigned one wins -- the last such in scan order -- otherwise the
//     first in scan order; scan order = SciPy's `remaining` array (nc-1 .. 0, permuted by swap-removal).
//
// What is different from the textbook layout: nothing of size nc (up to 40 000 columns in the dense bin)
// is stored except one "assigned" bit per column.
//   * v[j] != 0 only for assigned columns, and a column never becomes unassigned again, so v lives with
//     the row that owns the column (vcol[row]);
//   * the shortest-path labels spc[j] of an augmentation are recomputed from the (short) chain of rows
//     visited so far instead of being stored: spc[j] = min_t r_t(j), first minimum wins (= SciPy's
//     strict `<` update);
//   * `remaining` is represented by "base order + a few overrides".
// The cost accessor is a template parameter: explicit matrix (match_objects drop-in) or the virtual
// epipolar cost tensor (never materialised).
#pragma once
#include "common.cuh"

namespace bpc {

struct Cand {
    double val;
    int key;   // tie-break: larger wins.  unassigned: 2^30 + pos ; assigned: 2^30 - 1 - pos
    int col;
    int tau;   // chain index of the row that produced val (SciPy's path[col])
    int row;   // owner row for an assigned column, -1 otherwise
};

__device__ __forceinline__ bool cand_better(const Cand& a, const Cand& b) {
    return a.val < b.val || (a.val == b.val && a.key > b.key);
}

__device__ __forceinline__ Cand cand_shfl_xor(const Cand& c, int m) {
    Cand o;
    o.val = shfl_xor_d(c.val, m);
    o.key = __shfl_xor_sync(0xffffffffu, c.key, m);
    o.col = __shfl_xor_sync(0xffffffffu, c.col, m);
    o.tau = __shfl_xor_sync(0xffffffffu, c.tau, m);
    o.row = __shfl_xor_sync(0xffffffffu, c.row, m);
    return o;
}

// Shared-memory state of one assignment problem.  Sizes: nr_cap rows, nc_cap columns.
struct LsapState {
    double* u;        // [nr]
    double* vcol;     // [nr]  v of the column owned by the row
    int* col4row;     // [nr]
    uint32_t* asg;    // [ceil(nc/32)] assigned-column bitmap
    // chain of the current augmentation (length <= nr + 1)
    int* crow;        // [nr+1] rows visited, crow[0] = current row
    double* cm;       // [nr+1] minVal when the row was scanned
    double* cu;       // [nr+1] u[row]
    int* scol;        // [nr+1] column selected at each step
    double* sspc;     // [nr+1] its label (= minVal after the step)
    int* spath;       // [nr+1] chain index of path[scol]
    double* newv;     // [nr+1]
    int* ovpos;       // [nr+1] `remaining` overrides: position -> column
    int* ovcol;       // [nr+1]
    uint8_t* inchain; // [nr]   SR flag
    Cand* red;        // [32]   cross-warp reduction scratch
    int* ctl;         // [4]    0: next row, 1: done flag, 2: status, 3: nov
    int nr, nc;

    __device__ __forceinline__ bool assigned(int col) const { return (asg[col >> 5] >> (col & 31)) & 1u; }
    __device__ __forceinline__ void set_assigned(int col) { asg[col >> 5] |= 1u << (col & 31); }
};

__host__ __device__ inline size_t lsap_state_bytes(int nr_cap, int nc_cap) {
    size_t b = 0;
    b += (size_t)nr_cap * 8 * 2;                 // u, vcol
    b += (size_t)(nr_cap + 1) * 8 * 4;           // cm, cu, sspc, newv
    b += sizeof(Cand) * 32;                      // red (8-byte aligned region ends here)
    b += (size_t)nr_cap * 4;                     // col4row
    b += (size_t)(nr_cap + 1) * 4 * 5;           // crow, scol, spath, ovpos, ovcol
    b += (size_t)((nc_cap + 31) / 32) * 4;       // asg
    b += 16;                                     // ctl
    b += (size_t)((nr_cap + 7) / 8) * 8;         // inchain
    return (b + 15) & ~(size_t)15;
}

// Carve the state out of a shared-memory block (8-byte aligned).  Returns the first free byte.
__device__ inline unsigned char* lsap_state_carve(LsapState& st, unsigned char* p, int nr_cap, int nc_cap) {
    st.u = (double*)p; p += (size_t)nr_cap * 8;
    st.vcol = (double*)p; p += (size_t)nr_cap * 8;
    st.cm = (double*)p; p += (size_t)(nr_cap + 1) * 8;
    st.cu = (double*)p; p += (size_t)(nr_cap + 1) * 8;
    st.sspc = (double*)p; p += (size_t)(nr_cap + 1) * 8;
    st.newv = (double*)p; p += (size_t)(nr_cap + 1) * 8;
    st.red = (Cand*)p; p += sizeof(Cand) * 32;
    st.col4row = (int*)p; p += (size_t)nr_cap * 4;
    st.crow = (int*)p; p += (size_t)(nr_cap + 1) * 4;
    st.scol = (int*)p; p += (size_t)(nr_cap + 1) * 4;
    st.spath = (int*)p; p += (size_t)(nr_cap + 1) * 4;
    st.ovpos = (int*)p; p += (size_t)(nr_cap + 1) * 4;
    st.ovcol = (int*)p; p += (size_t)(nr_cap + 1) * 4;
    st.asg = (uint32_t*)p; p += (size_t)((nc_cap + 31) / 32) * 4;
    st.ctl = (int*)p; p += 16;
    st.inchain = (uint8_t*)p; p += (size_t)((nr_cap + 7) / 8) * 8;
    return (unsigned char*)(((uintptr_t)p + 15) & ~(uintptr_t)15);
}

// All threads: reset for a problem with nr rows, nc columns.  Caller syncs afterwards.
__device__ inline void lsap_reset(LsapState& st, int nr, int nc, int nthreads, int tid) {
    st.nr = nr; st.nc = nc;
    for (int i = tid; i < nr; i += nthreads) { st.u[i] = 0.0; st.vcol[i] = 0.0; st.col4row[i] = -1; st.inchain[i] = 0; }
    for (int w = tid; w < (nc + 31) / 32; w += nthreads) st.asg[w] = 0u;
    if (tid < 4) st.ctl[tid] = 0;
}

__device__ inline Cand block_best(LsapState& st, Cand c, int nthreads, int tid) {
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) {
        const Cand o = cand_shfl_xor(c, m);
        if (cand_better(o, c)) c = o;
    }
    if (nthreads <= 32) return c;
    const int nw = (nthreads + 31) >> 5;
    __syncthreads();                      // red[] free
    if ((tid & 31) == 0) st.red[tid >> 5] = c;
    __syncthreads();
    Cand b = st.red[0];
    for (int w = 1; w < nw; ++w) {
        const Cand o = st.red[w];
        if (cand_better(o, b)) b = o;
    }
    return b;
}

// Position of a not-yet-selected column in SciPy's `remaining` array after `removed` swap-removals.
__device__ __forceinline__ int lsap_pos(const LsapState& st, int col, int removed, int nov) {
    if (col >= removed) return st.nc - 1 - col;     // still at its base position
    for (int e = 0; e < nov; ++e)
        if (st.ovcol[e] == col) return st.ovpos[e];
    return st.nc - 1 - col;                          // unreachable for a live column
}

// Label of one unassigned column over the chain 0..t (first minimum wins = SciPy's strict `<` update) as a candidate.
template <class Acc>
__device__ __forceinline__ Cand lsap_label(const LsapState& st, const Acc& acc, int col, int t, int nov) {
    double spc = __longlong_as_double(0x7ff0000000000000LL);
    int tau = -1;
    for (int a = 0; a <= t; ++a) {
        const double r = dsub(dadd(st.cm[a], acc.cost(st.crow[a], col)), st.cu[a]);
        if (r < spc) { spc = r; tau = a; }
    }
    Cand c;
    c.val = spc; c.col = col; c.tau = tau; c.row = -1;
    c.key = (1 << 30) + lsap_pos(st, col, t, nov);
    return c;
}

// Exhaustive scan of the unassigned columns (what SciPy does); every thread folds its columns into `best`.
template <class Acc>
__device__ inline void lsap_scan_unassigned_full(const LsapState& st, const Acc& acc, int t, int nov, Cand& best, int nthreads, int tid) {
    for (int col = tid; col < st.nc; col += nthreads) {
        if (st.assigned(col)) continue;
        const Cand c = lsap_label(st, acc, col, t, nov);
        if (cand_better(c, best)) best = c;
    }
}

// One shortest-augmenting-path search + dual update + augmentation for row `cur`.
// Collective over the CTA.  Acc::cost(row, col) returns the float32 cost widened to double.
// Returns with st.ctl[2] != 0 if the problem is infeasible (all-infinite / NaN row).
template <class Acc>
__device__ void lsap_augment(LsapState& st, const Acc& acc, int cur, int nthreads, int tid) {
    const int nr = st.nr, nc = st.nc;
    __syncthreads();                      // every warp has read ctl[] of the previous search before it is reset
    if (tid == 0) {
        st.crow[0] = cur; st.cm[0] = 0.0; st.cu[0] = st.u[cur];
        st.ctl[1] = 0; st.ctl[3] = 0;
    }
    __syncthreads();
    for (int t = 0;; ++t) {
        const int nov = st.ctl[3];
        Cand best;
        best.val = __longlong_as_double(0x7ff0000000000000LL); best.key = -1; best.col = -1; best.tau = -1; best.row = -1;
        // unassigned columns: v == 0 (the accessor may enumerate only columns that can attain the minimum)
        acc.scan_unassigned(st, t, nov, best, nthreads, tid);
        // assigned columns not yet in the tree, one per owning row
        for (int i = tid; i < nr; i += nthreads) {
            const int col = st.col4row[i];
            if (col < 0 || st.inchain[i]) continue;
            const double vj = st.vcol[i];
            double spc = __longlong_as_double(0x7ff0000000000000LL);
            int tau = -1;
            for (int a = 0; a <= t; ++a) {
                const double r = dsub(dsub(dadd(st.cm[a], acc.cost(st.crow[a], col)), st.cu[a]), vj);
                if (r < spc) { spc = r; tau = a; }
            }
            Cand c;
            c.val = spc; c.col = col; c.tau = tau; c.row = i;
            c.key = (1 << 30) - 1 - lsap_pos(st, col, t, nov);
            if (cand_better(c, best)) best = c;
        }
        best = block_best(st, best, nthreads, tid);
        if (tid == 0) {
            if (best.col < 0 || !(best.val < __longlong_as_double(0x7ff0000000000000LL))) {
                st.ctl[2] = 1; st.ctl[1] = 1;              // infeasible
            } else {
                st.scol[t] = best.col; st.sspc[t] = best.val; st.spath[t] = best.tau;
                // swap-remove the selected position from `remaining`
                const bool is_assigned = best.key < (1 << 30);
                const int pos = is_assigned ? ((1 << 30) - 1 - best.key) : (best.key - (1 << 30));
                const int lastpos = nc - 1 - t;
                int n_ov = nov;
                if (pos != lastpos) {
                    int clast = nc - 1 - lastpos;            // base column of the last position
                    for (int e = 0; e < n_ov; ++e)
                        if (st.ovpos[e] == lastpos) clast = st.ovcol[e];
                    int slot = -1;
                    for (int e = 0; e < n_ov; ++e)
                        if (st.ovpos[e] == pos) slot = e;
                    if (slot < 0) slot = n_ov++;
                    st.ovpos[slot] = pos; st.ovcol[slot] = clast;
                }
                for (int e = 0; e < n_ov; ++e)               // the last position no longer exists
                    if (st.ovpos[e] == lastpos) { st.ovpos[e] = st.ovpos[n_ov - 1]; st.ovcol[e] = st.ovcol[n_ov - 1]; --n_ov; break; }
                st.ctl[3] = n_ov;
                if (is_assigned) {
                    const int i = best.row;
                    st.crow[t + 1] = i; st.cm[t + 1] = best.val; st.cu[t + 1] = st.u[i];
                    st.inchain[i] = 1;
                } else {
                    // sink reached: dual update (SciPy: u[cur] += minVal; visited rows / columns)
                    const int L = t;
                    const double minVal = best.val;
                    st.u[cur] = dadd(st.u[cur], minVal);
                    for (int a = 1; a <= L; ++a) {
                        const int i = st.crow[a];                  // owns scol[a-1]
                        st.u[i] = dadd(st.u[i], dsub(minVal, st.sspc[a - 1]));
                        st.newv[a - 1] = dsub(st.vcol[i], dsub(minVal, st.sspc[a - 1]));
                        st.vcol[i] = st.newv[a - 1];
                        st.inchain[i] = 0;
                    }
                    st.newv[L] = 0.0;                              // the sink was unassigned: v == 0
                    // augment along path[]
                    int a = L;
                    for (;;) {
                        const int pa = st.spath[a];                // chain index of path[scol[a]]
                        const int i = st.crow[pa];
                        st.col4row[i] = st.scol[a];
                        st.vcol[i] = st.newv[a];
                        if (pa == 0) break;
                        a = pa - 1;                                // the column row i owned before
                    }
                    st.set_assigned(st.scol[L]);
                    st.ctl[1] = 1;
                }
            }
        }
        __syncthreads();
        if (st.ctl[1]) break;
    }
}

}  // namespace bpc
