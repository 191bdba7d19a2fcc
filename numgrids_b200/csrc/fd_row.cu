// fd_row.cu -- finite-difference stencil apply along the CONTIGUOUS axis (second translation unit of the
// finite-difference path; fd.cu holds the gather and march kernels and the dispatch).  Same arithmetic
// contract as fd.cu: acc = 0.0; acc = acc + w_k * x_k in stored tap order, DMUL + DADD, never FMA.
#include "fd_fuse.cuh"

namespace {

// (b) derivative axis IS the contiguous one: the [rows][n] view is cut into chunks of 64 doubles (lane l
// holds the aligned pair 2l, 2l+1: one 512-byte load per chunk) and a warp takes a group of U consecutive
// chunks of one row.  All U loads (plus the two edge pairs the group's first / last lanes need from the
// neighbouring groups) are issued before any arithmetic.  The taps of neighbouring positions come from
// the neighbouring lanes by warp shuffles, rotated through the previous / next chunk's registers at the
// chunk edges.  Slab halos (HALO) enter as the positions -P..-1 and n..n+P-1 of the row; the one-sided
// schemes overwrite the central result on the (at most P) boundary positions of a row end.
// Per-warp context of the row kernel and the body for one group of U chunks.  HEAD / TAIL say whether
// (and for TAIL, how) the group touches the low / high end of its row: only then can a position fall outside [0, n) (slab
// halo or nothing there), a chunk be partial, or a one-sided scheme apply; interior groups run without
// any of those tests.
template <int P, int MODE, bool HALO>
struct RowCtx {
    static constexpr bool FUSED = MODE != 0;
    const double* __restrict__ src;
    double* __restrict__ dst;
    const double* __restrict__ xdiv;
    const double* lo_row;                                    // this row's P halo values, or nullptr
    const double* hi_row;
    i64 base, pre0, pre_a, post0, post_a, div0, div_a;
    int n;

    __device__ __forceinline__ double value_at(const FuseDev& F, int q) const {
        if (q < 0) return (HALO && lo_row && q >= -P) ? lo_row[P + q] : 0.0;
        if (q >= n) return (HALO && hi_row && q - n < P) ? hi_row[q - n] : 0.0;
        double x = src[q];
        if (FUSED && fm_pre<MODE>(F)) x = __dmul_rn(F.pre.ptr[pre0 + q * pre_a], x);
        return x;
    }
    template <bool CHECK>
    __device__ __forceinline__ double2 pair_at(const FuseDev& F, int pos) const {   // aligned pair (n is even)
        double2 v = make_double2(0.0, 0.0);
        if (!CHECK || (pos >= 0 && pos < n)) {
            v = ldg128(src + pos);
            if (FUSED && fm_pre<MODE>(F)) {      // (a row-constant table is the same address twice: one L1 line, no branch)
                v.x = __dmul_rn(F.pre.ptr[pre0 + pos * pre_a], v.x);
                v.y = __dmul_rn(F.pre.ptr[pre0 + (pos + 1) * pre_a], v.y);
            }
        } else if (HALO) {
            v.x = value_at(F, pos);
            v.y = value_at(F, pos + 1);
        }
        return v;
    }
};

// One-sided scheme at position q of the row (global boundary): sum_k w[k] * x(q + off[k]) from 0.0 in stored order.
// Out of line on purpose: it runs for at most P positions per row end, but inlined into every head / tail variant of
// the group body (four copies each) it made the kernels 12-14 thousand instructions long -- 200 KB of code that no
// instruction cache holds (stall_no_instruction 2.1 per issue in profiles/r01_fd_row_fused_ncu_summary.txt).
template <int P, int MODE, bool HALO>
__device__ __noinline__ double fd_row_one_sided(const double* __restrict__ src, const double* lo_row, const double* hi_row,
                                                const double* pre_ptr, i64 pre0, i64 pre_a, int n, int q,
                                                const double* w, const int* off, int n_side) {
    double a = 0.0;
    for (int k = 0; k < n_side; ++k) {
        const int t = q + off[k];
        double x;
        if (t < 0) x = (HALO && lo_row && t >= -P) ? lo_row[P + t] : 0.0;
        else if (t >= n) x = (HALO && hi_row && t - n < P) ? hi_row[t - n] : 0.0;
        else {
            x = src[t];
            if (MODE != 0 && pre_ptr) x = __dmul_rn(pre_ptr[pre0 + t * pre_a], x);
        }
        a = __dadd_rn(a, __dmul_rn(w[k], x));
    }
    return a;
}

template <int P, int MODE, bool HALO, int U, bool HEAD, int TAIL>
__device__ __forceinline__ void fd_row_group(const RowCtx<P, MODE, HALO>& R, const FdTables& T,
                                             const FuseDev& F, int c0, int lane) {
    constexpr int H = (P + 1) / 2;                           // neighbour lanes needed on each side
    constexpr int C = 2 * H;                                 // window index of this lane's .x
    const int n = R.n;
    const int p0 = c0 * 64 + 2 * lane;                       // this lane's pair in the group's first chunk
    double2 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) v[u] = R.template pair_at<TAIL == 2>(F, p0 + 64 * u);
    // virtual previous / next chunk: only the lanes next to the group are ever read
    double2 pcl = make_double2(0.0, 0.0), ncr = make_double2(0.0, 0.0);
    if (lane >= 32 - H) pcl = R.template pair_at<HEAD>(F, p0 - 64);
    if (lane < H) ncr = R.template pair_at<TAIL != 0>(F, p0 + 64 * U);
    const bool one_lo = HEAD && !(HALO && R.lo_row), one_hi = TAIL != 0 && !(HALO && R.hi_row);
    // compile-time modes: coefficients that do not vary along the row are read once per group
    double post_c = 0.0, div_c = 0.0;
    if (MODE > 0) {
        if ((MODE & FM_POST) && R.post_a == 0) post_c = __ldg(F.post.ptr + R.post0);
        if ((MODE & FM_DIV) && R.div_a == 0) div_c = __ldg(F.div.ptr + R.div0);
    }

#pragma unroll
    for (int u = 0; u < U; ++u) {
        const int pos = p0 + 64 * u;
        if (TAIL == 2 && (c0 + u) * 64 >= n) break;          // warp-uniform: chunk past the row end
        const double2 prev = u == 0 ? pcl : v[u > 0 ? u - 1 : 0];
        const double2 next = u == U - 1 ? ncr : v[u < U - 1 ? u + 1 : 0];
        double win[4 * H + 2];                               // positions pos-2H .. pos+2H+1 of the row
        win[C] = v[u].x;
        win[C + 1] = v[u].y;
#pragma unroll
        for (int j = 1; j <= H; ++j) {
            const double2 s = (lane >= 32 - j) ? prev : v[u];   // what lane (l+j)&31 needs from this lane
            win[C - 2 * j] = __shfl_sync(0xffffffffu, s.x, (lane - j) & 31);
            win[C - 2 * j + 1] = __shfl_sync(0xffffffffu, s.y, (lane - j) & 31);
            const double2 t = (lane < j) ? next : v[u];
            win[C + 2 * j] = __shfl_sync(0xffffffffu, t.x, (lane + j) & 31);
            win[C + 2 * j + 1] = __shfl_sync(0xffffffffu, t.y, (lane + j) & 31);
        }
        double ax = 0.0, ay = 0.0;
#pragma unroll
        for (int kk = 0; kk < 2 * P + 1; ++kk) {
            ax = __dadd_rn(ax, __dmul_rn(T.wc[kk], win[C - P + kk]));
            ay = __dadd_rn(ay, __dmul_rn(T.wc[kk], win[C + 1 - P + kk]));
        }
        // one-sided schemes on the global boundary: only the first / last chunk of a row can hold them
        // (TAIL == 1: the group ends exactly at the row end, so only its last chunk; HEAD: only its first)
        if ((one_lo && u == 0) || (one_hi && (TAIL == 1 ? u == U - 1 : (c0 + u) * 64 + 64 + P > n))) {
            const double* pre_ptr = (MODE != 0 && fm_pre<MODE>(F)) ? F.pre.ptr : nullptr;
            auto one_sided = [&](int q, const double* w, const int* off) -> double {
                return fd_row_one_sided<P, MODE, HALO>(R.src, R.lo_row, R.hi_row, pre_ptr, R.pre0, R.pre_a, n, q, w, off, T.n_side);
            };
            if (TAIL != 2 || pos < n) {
                if (one_lo && pos < P) {
                    ax = one_sided(pos, T.wf, T.of);
                    if (pos + 1 < P) ay = one_sided(pos + 1, T.wf, T.of);
                }
                if (one_hi && pos + 1 >= n - P) {
                    ay = one_sided(pos + 1, T.wb, T.ob);
                    if (pos >= n - P) ax = one_sided(pos, T.wb, T.ob);
                }
            }
        }
        if (TAIL != 2 || pos < n) {
            double2 r;
            r.x = __dmul_rn(ax, T.inv_h_pow);
            r.y = __dmul_rn(ay, T.inv_h_pow);
            if (R.xdiv) {
                const double2 xd = ldg128(R.xdiv + pos);
                r.x = __ddiv_rn(r.x, xd.x);
                r.y = __ddiv_rn(r.y, xd.y);
            }
            if (MODE > 0) {
                double2 mn = make_double2(0.0, 0.0), ad = mn, pc = mn, dc = mn;
                if (MODE & FM_MINUEND) mn = *reinterpret_cast<const double2*>(F.minuend + R.base + pos);   // may alias out: plain load
                if (MODE & FM_ADDEND) ad = *reinterpret_cast<const double2*>(F.addend + R.base + pos);
                if (MODE & FM_POST) pc = coef_pair(F.post.ptr, R.post0, pos, R.post_a, post_c);
                if (MODE & FM_DIV) dc = coef_pair(F.div.ptr, R.div0, pos, R.div_a, div_c);
                r.x = fm_tail_vals<MODE>(r.x, mn.x, pc.x, ad.x, dc.x);
                r.y = fm_tail_vals<MODE>(r.y, mn.y, pc.y, ad.y, dc.y);
            } else if (MODE != 0) {
                const i64 po = fm_post<MODE>(F) ? R.post0 + pos * R.post_a : 0;
                const i64 dv = fm_div<MODE>(F) ? R.div0 + pos * R.div_a : 0;
                r.x = fm_tail<MODE>(F, r.x, R.base + pos, po, dv);
                r.y = fm_tail<MODE>(F, r.y, R.base + pos + 1, po + R.post_a, dv + R.div_a);
            }
            *reinterpret_cast<double2*>(R.dst + pos) = r;
        }
    }
}

template <int P, int MODE, bool HALO, int U>
__global__ void __launch_bounds__(256)
fd_row_kernel(const double* __restrict__ in, double* __restrict__ out, i64 rows, int n, int axis,
              const __grid_constant__ FdTables T, const double* __restrict__ xdiv, const __grid_constant__ FuseDev F,
              const double* __restrict__ lo,
              const double* __restrict__ hi, int gpr /* groups per row */) {
    const int lane = threadIdx.x & 31;
    const i64 g = (i64)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (g >= rows * gpr) return;                             // whole warps leave together
    i64 row;
    int grp;
    if (g < 0x7fffffffLL) {                                  // 32-bit divide (a 64-bit one costs ~100 instructions)
        const unsigned r32 = (unsigned)g / (unsigned)gpr;
        row = r32;
        grp = (int)((unsigned)g - r32 * (unsigned)gpr);
    } else {
        row = g / gpr;
        grp = (int)(g - row * gpr);
    }
    constexpr bool FUSED = MODE != 0;
    RowCtx<P, MODE, HALO> R;
    R.n = n;
    R.base = row * (i64)n;
    R.src = in + R.base;
    R.dst = out + R.base;
    R.xdiv = xdiv;
    R.lo_row = (HALO && lo) ? lo + row * P : nullptr;
    R.hi_row = (HALO && hi) ? hi + row * P : nullptr;
    R.pre0 = R.pre_a = R.post0 = R.post_a = R.div0 = R.div_a = 0;
    if (FUSED) {
        const Idx4 ix = unravel4(R.base, F.shape);           // coordinate along `axis` (the last one) is 0
        if (fm_pre<MODE>(F)) { R.pre0 = coef_off(F.pre, ix); R.pre_a = F.pre.s[axis]; }
        if (fm_post<MODE>(F)) { R.post0 = coef_off(F.post, ix); R.post_a = F.post.s[axis]; }
        if (fm_div<MODE>(F)) { R.div0 = coef_off(F.div, ix); R.div_a = F.div.s[axis]; }
    }
    const int c0 = grp * U;                                  // first chunk of the group
    const bool head = c0 == 0;
    // 0: the group and its right neighbour pair are inside the row; 1: the group ends exactly at the row
    // end (only the neighbour pair is outside); 2: ragged (partial or missing chunks)
    const int tail = (c0 + U) * 64 + 64 <= n ? 0 : ((c0 + U) * 64 == n ? 1 : 2);
    if (!head) {
        if (tail == 0) fd_row_group<P, MODE, HALO, U, false, 0>(R, T, F, c0, lane);
        else if (tail == 1) fd_row_group<P, MODE, HALO, U, false, 1>(R, T, F, c0, lane);
        else fd_row_group<P, MODE, HALO, U, false, 2>(R, T, F, c0, lane);
    } else {
        if (tail == 0) fd_row_group<P, MODE, HALO, U, true, 0>(R, T, F, c0, lane);
        else if (tail == 1) fd_row_group<P, MODE, HALO, U, true, 1>(R, T, F, c0, lane);
        else fd_row_group<P, MODE, HALO, U, true, 2>(R, T, F, c0, lane);
    }
}

template <int P>
int launch_row(const ng_plan* p, const double* in, double* out, const FuseDev& F, const double* lo,
               const double* hi, cudaStream_t st) {
    const int n = (int)p->n;
    {
        const i64 rows = p->outer;
        const int nchunk = (n + 63) / 64;
        const bool halo = lo || hi;
        // U consecutive chunks per warp (U x 512 B in flight); small problems keep U = 1 so that every SM gets CTAs
        const bool deep = rows * nchunk >= 148LL * 8 * 4 * 4 && nchunk >= 3;
#define NG_FD_ROW(FU, HA, UU)                                                                            \
    do {                                                                                                 \
        const int gpr = (nchunk + UU - 1) / UU;                                                          \
        fd_row_kernel<P, FU, HA, UU><<<(unsigned)((rows * gpr + 7) / 8), 256, 0, st>>>(                    \
            in, out, rows, n, p->axis, p->fd, p->d_xdiv, F, lo, hi, gpr);                                \
    } while (0)
        const int ue = (int)NG_ENV_INT("NG_FD_ROW_U", 0);     // tuning: force the chunks per warp
        // measured at 1024^3 / 8192^2 (profiles/r01_tune_fd.txt): U = 1 0.60 / 0.65, U = 2 0.63 / 0.76,
        // U = 4 0.80 / 0.92, U = 8 0.69 / 0.58 of the HBM peak
        const int uu = ue ? ue : (!deep ? 1 : 4);
#define NG_FD_ROW2(FU, HA) do { if (uu >= 4) NG_FD_ROW(FU, HA, 4); else NG_FD_ROW(FU, HA, 1); } while (0)
        if (F.any) { if (halo) NG_FD_ROW2(-1, true); else NG_FD_ROW2(-1, false); }
        else       { if (halo) NG_FD_ROW2(0, true); else NG_FD_ROW2(0, false); }
#undef NG_FD_ROW2
#undef NG_FD_ROW
        NG_TRACE(uu >= 4 ? "fd_row<u4>" : "fd_row<u1>");
    }
    NG_LAUNCH_CHECK();
    return NG_OK;
}

}  // namespace

// This file is compiled twice (numgrids_b200/build.py) so that the two halves build in parallel:
// NG_ROW_PSET = 0 instantiates the half widths 1 and 2 and the dispatcher, NG_ROW_PSET = 1 the half widths 3 and 4.
// (Fused applies of 2^18 points and more run the shared-memory row tiles of fd_tile.cu with compile-time fuse patterns;
// what reaches the MODE = -1 run-time form here is small, launch-bound arrays and slab halos.)
#ifndef NG_ROW_PSET
#define NG_ROW_PSET 0
#endif
int ng_fd_row_launch_wide(int P, const ng_plan* p, const double* in, double* out, const FuseDev& F,
                          const double* lo, const double* hi, cudaStream_t st);
int ng_fd_tile_try_launch(int P, const ng_plan* p, const double* in, double* out, const FuseDev& F, cudaStream_t st,
                          bool* handled);
#if NG_ROW_PSET == 0
int ng_fd_row_launch(int P, const ng_plan* p, const double* in, double* out, const FuseDev& F,
                     const double* lo, const double* hi, cudaStream_t st) {
    if (!lo && !hi) {                                        // shared-memory row tiles (fd_tile.cu)
        bool handled = false;
        const int rc = ng_fd_tile_try_launch(P, p, in, out, F, st, &handled);
        if (rc || handled) return rc;
    }
    switch (P) {
        case 1: return launch_row<1>(p, in, out, F, lo, hi, st);
        case 2: return launch_row<2>(p, in, out, F, lo, hi, st);
        case 3:
        case 4: return ng_fd_row_launch_wide(P, p, in, out, F, lo, hi, st);
    }
    ng_set_error("ng_fd_row_launch: unsupported half width %d", P);
    return NG_EINVAL;
}
#else
int ng_fd_row_launch_wide(int P, const ng_plan* p, const double* in, double* out, const FuseDev& F,
                          const double* lo, const double* hi, cudaStream_t st) {
    if (P == 3) return launch_row<3>(p, in, out, F, lo, hi, st);
    return launch_row<4>(p, in, out, F, lo, hi, st);
}
#endif
