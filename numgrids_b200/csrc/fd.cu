// fd.cu -- finite-difference stencil apply (FiniteDifferenceDiff / LogDiff).
//
// Replaces findiff.FinDiff(axis, h, order, acc)(f) behind reference numgrids/diff.py:88 and
// :259-262.  Arithmetic contract (bit-exact with the reference):
//   acc = 0.0; for k in stored tap order: acc = acc + (w_k * f[i+o_k])   (DMUL, DADD -- never FMA)
//   out = acc * inv_h_pow            [LogDiff: then out / x[i]]
// central taps on [nb, n-nb), forward on [0, nb), backward on [n-nb, n).
//
// Layout: the array is viewed as [outer][n][inner] around the derivative axis.  Three kernels:
//   * fd_march_kernel (this file): axis not contiguous, >= 2^21 points -- a thread owns a pair of adjacent
//     columns and marches along n with a (2P+1)-deep register window; every element is loaded once, every
//     access is 128 bits wide (0.96-0.98 of the measured HBM peak at 1024^3).
//   * fd_row_kernel (fd_row.cu): axis contiguous (also the slab-sharded axis, with halos) -- a warp owns a
//     group of 64-double chunks of one row, the taps of neighbouring positions come by warp shuffles.
//   * fd_apply_kernel (this file): any shape / alignment / tap layout and small arrays -- one thread per
//     output pair, taps gathered through L1/L2 (one-load-deep threads: lowest latency at cfg1's 512 x 512).
// Algorithmic traffic 16 B/point (read f once, write out once).  All three produce the same bits
// (tools/check_fd_stream.py, tests/test_gpu_ab.py).
#include "common.cuh"

namespace {

template <int VEC>
struct Vec;
template <>
struct Vec<1> {
    typedef double T;
    __device__ static __forceinline__ double get(const T& v, int) { return v; }
    __device__ static __forceinline__ void set(T& v, int, double x) { v = x; }
};
template <>
struct Vec<2> {
    typedef double2 T;
    __device__ static __forceinline__ double get(const T& v, int j) { return j ? v.y : v.x; }
    __device__ static __forceinline__ void set(T& v, int j, double x) {
        if (j) v.y = x; else v.x = x;
    }
};

// FUSED: ng_fuse epilogue/prologue present.  HALO: slab mode (lo/hi neighbour planes).
template <int VEC, bool FUSED, bool HALO>
__global__ void __launch_bounds__(256)
fd_apply_kernel(const double* __restrict__ in, double* __restrict__ out, i64 outer, i64 n,
                i64 inner, int axis, FdTables T, const double* __restrict__ xdiv, FuseDev F,
                const double* __restrict__ lo, const double* __restrict__ hi) {
    typedef typename Vec<VEC>::T VT;
    const i64 inner_v = inner / VEC;
    const i64 total = outer * n * inner_v;
    for (i64 idx = (i64)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (i64)gridDim.x * blockDim.x) {
        i64 c, i, o;
        if (idx < 0x7fffffffLL) {
            unsigned t = (unsigned)idx, q = t / (unsigned)inner_v;
            c = (t - q * (unsigned)inner_v) * VEC;
            unsigned q2 = q / (unsigned)n;
            i = q - q2 * (unsigned)n;
            o = q2;
        } else {
            c = (idx % inner_v) * VEC;
            i64 t = idx / inner_v;
            i = t % n;
            o = t / n;
        }
        const bool has_lo = HALO && lo != nullptr, has_hi = HALO && hi != nullptr;
        const double* w;
        const int* off;
        int nt;
        if (i < T.nb && !has_lo) { w = T.wf; off = T.of; nt = T.n_side; }
        else if (i >= n - T.nb && !has_hi) { w = T.wb; off = T.ob; nt = T.n_side; }
        else { w = T.wc; off = T.oc; nt = T.n_center; }

        const i64 lin = (o * n + i) * inner + c;
        Idx4 ix;
        i64 pre_base = 0;
        if (FUSED) {
            ix = unravel4(lin, F.shape);
            if (F.pre.ptr) pre_base = coef_off(F.pre, ix) - ix.i[axis] * F.pre.s[axis];
        }
        double acc[VEC];
#pragma unroll
        for (int j = 0; j < VEC; ++j) acc[j] = 0.0;
        for (int k = 0; k < nt; ++k) {
            const i64 ii = i + off[k];
            VT v;
            if (HALO && ii < 0) {
#pragma unroll
                for (int j = 0; j < VEC; ++j)
                    Vec<VEC>::set(v, j, lo[(o * inner + c + j) * T.nb + (T.nb + ii)]);
            } else if (HALO && ii >= n) {
#pragma unroll
                for (int j = 0; j < VEC; ++j)
                    Vec<VEC>::set(v, j, hi[(o * inner + c + j) * T.nb + (ii - n)]);
            } else {
                v = *reinterpret_cast<const VT*>(in + (o * n + ii) * inner + c);
            }
            const double wk = w[k];
#pragma unroll
            for (int j = 0; j < VEC; ++j) {
                double x = Vec<VEC>::get(v, j);
                if (FUSED && F.pre.ptr && !(HALO && (ii < 0 || ii >= n))) {   // halo planes arrive pre-multiplied
                    // coefficient at the tap location; the last (contiguous) grid axis advances with j
                    i64 po = pre_base + ii * F.pre.s[axis];
                    if (VEC > 1) po += j * F.pre.s[F.vec_axis];   // c+j stays inside the innermost row
                    x = __dmul_rn(F.pre.ptr[po], x);
                }
                acc[j] = __dadd_rn(acc[j], __dmul_rn(wk, x));
            }
        }
        VT r;
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
            double d = __dmul_rn(acc[j], T.inv_h_pow);
            if (xdiv) d = __ddiv_rn(d, xdiv[i]);
            if (FUSED) {
                Idx4 jx = ix;
                if (VEC > 1 && j) jx.i[F.vec_axis] += j;
                d = fuse_tail(F, d, lin + j, jx);
            }
            Vec<VEC>::set(r, j, d);
        }
        *reinterpret_cast<VT*>(out + lin) = r;
    }
}

// -------------------------------------------------------------------------------------------------
// Streaming kernels for the central scheme with offsets -P..P (every findiff central stencil), P <= 4.
// Each input element is loaded from global memory exactly once (plus a (2P)-element window refill
// per segment) and every access is 128 bits wide, so the kernels run at the HBM roofline (16 B/pt);
// the gather kernel above re-reads every tap through L1/L2 and stays the any-shape fallback.
// Same arithmetic as the gather kernel, bit for bit: acc = 0.0; acc = acc + w_k * x_k in stored order.
// -------------------------------------------------------------------------------------------------
__device__ __forceinline__ double2 ldg128(const double* p) {
    return __ldg(reinterpret_cast<const double2*>(p));
}

// (a) derivative axis is NOT the contiguous one: thread = one pair of adjacent columns of the
// [outer][n][inner] view, marching along n with a (2P+1)-deep window of double2 in registers (the loop
// is unrolled by the window depth, so the rotation costs no moves).  Consecutive threads own
// consecutive column pairs: every load/store of a warp is one contiguous 512-byte row piece.  n is cut
// into `nseg` segments when there are too few columns to fill the machine.
#ifndef NG_FD_MINB
#define NG_FD_MINB 1
#endif
// XD: LogDiff's `/ x[i]` (two correctly rounded divides per pair).  The divide sequences are latency-bound,
// so that variant is capped at 64 registers (4 CTAs per SM: 0.80 -> 0.87 of the HBM peak at 1024^3); the
// plain stencil is faster uncapped (0.97 vs 0.91).
template <int P, bool FUSED, bool XD>
__global__ void __launch_bounds__(256, FUSED ? 2 : (XD ? 4 : NG_FD_MINB))
fd_march_kernel(const double* __restrict__ in, double* __restrict__ out, i64 outer, int n, i64 inner,
                int axis, FdTables T, const double* __restrict__ xdiv, FuseDev F, int seg_len, int nseg) {
    constexpr int W = 2 * P + 1;
    const i64 inner2 = inner >> 1, ncol = outer * inner2;
    const i64 idx = (i64)blockIdx.x * 256 + threadIdx.x;
    if (idx >= ncol * nseg) return;
    const int seg = (int)(idx / ncol);
    const i64 col = idx - (i64)seg * ncol;
    const i64 o = col / inner2, c = (col - o * inner2) * 2;
    const int i0 = seg * seg_len, i1 = min(n, i0 + seg_len);
    const i64 base = o * (i64)n * inner + c;                 // element i of this column pair: base + i*inner
    const double* __restrict__ src = in + base;
    double* __restrict__ dst = out + base;

    i64 pre0 = 0, pre_a = 0, pre_v = 0, post0 = 0, post_a = 0, post_v = 0, div0 = 0, div_a = 0, div_v = 0;
    if (FUSED) {
        const Idx4 ix = unravel4(base, F.shape);             // coordinate along `axis` is 0 here
        if (F.pre.ptr) { pre0 = coef_off(F.pre, ix); pre_a = F.pre.s[axis]; pre_v = F.pre.s[F.vec_axis]; }
        if (F.post.ptr) { post0 = coef_off(F.post, ix); post_a = F.post.s[axis]; post_v = F.post.s[F.vec_axis]; }
        if (F.div.ptr) { div0 = coef_off(F.div, ix); div_a = F.div.s[axis]; div_v = F.div.s[F.vec_axis]; }
    }
    auto ld = [&](int i) -> double2 {
        double2 v = ldg128(src + (i64)i * inner);
        if (FUSED && F.pre.ptr) {                            // `coeff * v[i]` once, as the value enters the window
            const i64 po = pre0 + i * pre_a;
            v.x = __dmul_rn(F.pre.ptr[po], v.x);
            v.y = __dmul_rn(F.pre.ptr[po + pre_v], v.y);
        }
        return v;
    };
    // full-shape tail operands of row i (plain loads: they may alias `out`); the launch checks alignment
    const double2 zero2 = make_double2(0.0, 0.0);
    auto fetch_m = [&](int i) -> double2 {
        return (FUSED && F.minuend) ? *reinterpret_cast<const double2*>(F.minuend + base + (i64)i * inner) : zero2;
    };
    auto fetch_a = [&](int i) -> double2 {
        return (FUSED && F.addend) ? *reinterpret_cast<const double2*>(F.addend + base + (i64)i * inner) : zero2;
    };
    auto finish = [&](int i, double ax, double ay, const double2& m, const double2& a) {
        double2 r;
        r.x = __dmul_rn(ax, T.inv_h_pow);
        r.y = __dmul_rn(ay, T.inv_h_pow);
        if (XD) { const double xd = xdiv[i]; r.x = __ddiv_rn(r.x, xd); r.y = __ddiv_rn(r.y, xd); }
        if (FUSED) {
            r.x = fuse_tail_vals(F, r.x, m.x, a.x, post0 + i * post_a, div0 + i * div_a);
            r.y = fuse_tail_vals(F, r.y, m.y, a.y, post0 + i * post_a + post_v, div0 + i * div_a + div_v);
        }
        *reinterpret_cast<double2*>(dst + (i64)i * inner) = r;
    };
    auto one_sided = [&](int i, const double* w, const int* off) {
        double ax = 0.0, ay = 0.0;
        for (int k = 0; k < T.n_side; ++k) {
            const double2 v = ld(i + off[k]);
            ax = __dadd_rn(ax, __dmul_rn(w[k], v.x));
            ay = __dadd_rn(ay, __dmul_rn(w[k], v.y));
        }
        finish(i, ax, ay, fetch_m(i), fetch_a(i));
    };
    for (int i = i0; i < min(P, i1); ++i) one_sided(i, T.wf, T.of);
    for (int i = max(i0, n - P); i < i1; ++i) one_sided(i, T.wb, T.ob);

    const int ia = max(i0, P), ib = min(i1, n - P);
    if (ia >= ib) return;
    // window: old[] = f[i-P .. i+P-1].  Full groups of W outputs issue their W loads up front (W x 16 B
    // in flight per thread), then compute; the last (< W) outputs go one at a time.
    double2 old[W - 1];
#pragma unroll
    for (int u = 0; u < W - 1; ++u) old[u] = ld(ia - P + u);
    int i = ia;
    for (; i + W <= ib; i += W) {
        double2 nw[W], tm[W], ta[W];                         // f[i+P .. i+P+W-1]; tail operands of rows i .. i+W-1
#pragma unroll
        for (int u = 0; u < W; ++u) nw[u] = ld(i + P + u);
        if (FUSED) {
#pragma unroll
            for (int u = 0; u < W; ++u) { tm[u] = fetch_m(i + u); ta[u] = fetch_a(i + u); }
        }
#pragma unroll
        for (int u = 0; u < W; ++u) {
            double ax = 0.0, ay = 0.0;
#pragma unroll
            for (int k = 0; k < W; ++k) {                    // x_k = f[i+u-P+k]
                const double2 x = (u + k < W - 1) ? old[u + k < W - 1 ? u + k : 0] : nw[u + k >= W - 1 ? u + k - (W - 1) : 0];
                ax = __dadd_rn(ax, __dmul_rn(T.wc[k], x.x));
                ay = __dadd_rn(ay, __dmul_rn(T.wc[k], x.y));
            }
            finish(i + u, ax, ay, tm[u], ta[u]);
        }
#pragma unroll
        for (int u = 0; u < W - 1; ++u) old[u] = nw[u + 1];
    }
    for (; i < ib; ++i) {
        const double2 v = ld(i + P);
        double ax = 0.0, ay = 0.0;
#pragma unroll
        for (int k = 0; k < W - 1; ++k) {
            ax = __dadd_rn(ax, __dmul_rn(T.wc[k], old[k].x));
            ay = __dadd_rn(ay, __dmul_rn(T.wc[k], old[k].y));
        }
        ax = __dadd_rn(ax, __dmul_rn(T.wc[W - 1], v.x));
        ay = __dadd_rn(ay, __dmul_rn(T.wc[W - 1], v.y));
        finish(i, ax, ay, fetch_m(i), fetch_a(i));
#pragma unroll
        for (int u = 0; u < W - 2; ++u) old[u] = old[u + 1];
        old[W - 2] = v;
    }
}

// first / last nb planes along the axis -> [rows][nb]
__global__ void fd_pack_halo_kernel(const double* __restrict__ in, double* __restrict__ lo_planes,
                                    double* __restrict__ hi_planes, i64 outer, i64 n, i64 inner,
                                    int nb, FuseDev F) {
    const i64 rows = outer * inner;
    const i64 total = rows * nb;
    for (i64 idx = (i64)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (i64)gridDim.x * blockDim.x) {
        const i64 row = idx / nb;
        const int k = (int)(idx - row * nb);
        const i64 o = row / inner, c = row - o * inner;
        const i64 a = (o * n + k) * inner + c, b = (o * n + (n - nb + k)) * inner + c;
        double va = in[a], vb = in[b];
        if (F.pre.ptr) {                                  // ng_fd_pack_halo_fused: planes leave as pre * value
            va = __dmul_rn(F.pre.ptr[coef_off(F.pre, unravel4(a, F.shape))], va);
            vb = __dmul_rn(F.pre.ptr[coef_off(F.pre, unravel4(b, F.shape))], vb);
        }
        if (lo_planes) lo_planes[idx] = va;
        if (hi_planes) hi_planes[idx] = vb;
    }
}

// last-axis, nb == 2 fast path: one thread per row moves both 16-byte ends with 128-bit accesses
__global__ void fd_pack_halo_last2_kernel(const double* __restrict__ in, double2* __restrict__ lo_planes,
                                          double2* __restrict__ hi_planes, i64 rows, i64 n) {
    for (i64 row = (i64)blockIdx.x * blockDim.x + threadIdx.x; row < rows;
         row += (i64)gridDim.x * blockDim.x) {
        const double* p = in + row * n;
        if (lo_planes) lo_planes[row] = *reinterpret_cast<const double2*>(p);
        if (hi_planes) hi_planes[row] = *reinterpret_cast<const double2*>(p + n - 2);
    }
}

inline int grid_for(i64 work_items, int block) {
    i64 g = (work_items + block - 1) / block;
    const i64 cap = 148LL * 64;   // grid-stride beyond 64 CTAs per SM
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (int)g;
}

}  // namespace

// central offsets -P..P in ascending stored order, P = nb <= 4: the streaming kernels apply
static int central_half_width(const FdTables& T) {
    if (T.n_center % 2 == 0 || T.n_center < 3 || T.n_center > 9) return 0;
    const int P = T.n_center / 2;
    if (T.nb != P) return 0;
    for (int k = 0; k < T.n_center; ++k) if (T.oc[k] != k - P) return 0;
    return P;
}

template <int P>
static int launch_stream(const ng_plan* p, const double* in, double* out, const FuseDev& F,
                         const double* lo, const double* hi, cudaStream_t st) {
    const i64 target = 148LL * 2048 * 2;                     // threads (two full waves) worth aiming for
    const int n = (int)p->n;
    if (p->inner > 1) {
        const i64 ncol = p->outer * (p->inner / 2);
        i64 nseg = (target + ncol - 1) / ncol;
        const i64 max_seg = n / 16 > 0 ? n / 16 : 1;          // keep segments >= 16 rows (window refill <= 2P/16, L2 hits)
        if (nseg > max_seg) nseg = max_seg;
        if (nseg < 1) nseg = 1;
        const int seg_len = (int)((n + nseg - 1) / nseg);
        nseg = (n + seg_len - 1) / seg_len;
        const i64 threads = ncol * nseg;
        const unsigned grid = (unsigned)((threads + 255) / 256);
#define NG_FD_MARCH(FU, XD) fd_march_kernel<P, FU, XD><<<grid, 256, 0, st>>>(in, out, p->outer, n, p->inner, p->axis, p->fd, p->d_xdiv, F, seg_len, (int)nseg)
        if (F.any) { if (p->d_xdiv) NG_FD_MARCH(true, true); else NG_FD_MARCH(true, false); }
        else       { if (p->d_xdiv) NG_FD_MARCH(false, true); else NG_FD_MARCH(false, false); }
#undef NG_FD_MARCH
        NG_TRACE("fd_march");
    } else {
        return ng_fd_row_launch(P, p, in, out, F, lo, hi, st);
    }
    NG_LAUNCH_CHECK();
    return NG_OK;
}

int ng_fd_launch(const ng_plan* p, const double* in, double* out, const FuseDev& F,
                 const double* lo, const double* hi, cudaStream_t st) {
    const bool halo = (lo != nullptr) || (hi != nullptr);
    const bool aligned = (((uintptr_t)in | (uintptr_t)out) & 15) == 0;
    {
        // streaming kernels: 128-bit accesses throughout, so pairs along the contiguous direction must
        // be aligned and must not straddle a row of the innermost grid axis
        const bool off = NG_ENV_INT("NG_FD_STREAM", 1) == 0;   // tuning / A-B checks: 0 = gather kernel only
        const int P = off ? 0 : central_half_width(p->fd);
        const i64 contiguous = p->inner > 1 ? p->inner : p->n;
        // tiny arrays are launch/latency-bound: the gather kernel's one-load-deep threads finish sooner
        // than a march (NG_FD_STREAM_MIN: points from which the march kernel takes over)
        const i64 march_min = NG_ENV_INT("NG_FD_STREAM_MIN", 1LL << 21);
        const bool ok = P > 0 && aligned && contiguous % 2 == 0 && p->n >= 2 * P + 1 && p->n >= p->fd.n_side &&
                        p->n < (1LL << 30) &&
                        (p->inner == 1 || p->size >= march_min) &&
                        p->size > 0 && !(halo && p->inner > 1) &&
                        (!F.any || F.shape[F.vec_axis] % 2 == 0) &&
                        (!halo || ((((uintptr_t)lo | (uintptr_t)hi) & 7) == 0)) &&
                        ((((uintptr_t)F.minuend | (uintptr_t)F.addend) & 15) == 0);   // fetched as aligned pairs
        if (ok) {
            switch (P) {
                case 1: return launch_stream<1>(p, in, out, F, lo, hi, st);
                case 2: return launch_stream<2>(p, in, out, F, lo, hi, st);
                case 3: return launch_stream<3>(p, in, out, F, lo, hi, st);
                case 4: return launch_stream<4>(p, in, out, F, lo, hi, st);
            }
        }
    }
    // 128-bit path: pairs must not straddle a row of the innermost (non-unit) grid axis when
    // coefficients are indexed per axis
    const bool vec2 = (p->inner % 2 == 0) && aligned && (!F.any || F.shape[F.vec_axis] % 2 == 0);
    const i64 items = p->size / (vec2 ? 2 : 1);
    if (items == 0) return NG_OK;
    const int block = 256;
    const int grid = grid_for(items, block);
#define NG_FD_GO(V, FU, HA)                                                                      \
    fd_apply_kernel<V, FU, HA><<<grid, block, 0, st>>>(in, out, p->outer, p->n, p->inner,        \
                                                       p->axis, p->fd, p->d_xdiv, F, lo, hi)
    if (vec2) {
        if (F.any) { if (halo) NG_FD_GO(2, true, true); else NG_FD_GO(2, true, false); }
        else       { if (halo) NG_FD_GO(2, false, true); else NG_FD_GO(2, false, false); }
    } else {
        if (F.any) { if (halo) NG_FD_GO(1, true, true); else NG_FD_GO(1, true, false); }
        else       { if (halo) NG_FD_GO(1, false, true); else NG_FD_GO(1, false, false); }
    }
#undef NG_FD_GO
    NG_TRACE("fd_apply");
    NG_LAUNCH_CHECK();
    return NG_OK;
}

extern "C" int ng_fd_pack_halo_fused(const ng_plan* p, const double* d_in, const ng_coef* pre,
                                     double* d_lo_planes, double* d_hi_planes, void* stream) {
    NG_REQUIRE(p && p->kind == PK_FD, "ng_fd_pack_halo_fused: FD plan required");
    NG_REQUIRE(d_in, "ng_fd_pack_halo_fused: null input");
    ng_fuse f;
    memset(&f, 0, sizeof(f));
    if (pre) f.pre = *pre;
    FuseDev F;
    ng_fuse_to_dev(p, &f, &F);
    const i64 total = p->outer * p->inner * p->fd.nb;
    if (total == 0) return NG_OK;
    fd_pack_halo_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(
        d_in, d_lo_planes, d_hi_planes, p->outer, p->n, p->inner, p->fd.nb, F);
    NG_LAUNCH_CHECK();
    return NG_OK;
}

// Planes [x_begin, x_end) of axis 0 only: rows x_begin*R .. x_end*R of the [rows][nb] halo buffers
// (R = rows per axis-0 plane).  The chunked host pipeline exchanges halos chunk by chunk with it.
extern "C" int ng_fdlap_pack_halo_range(const ng_plan* p, const double* d_in, double* d_lo_planes,
                                        double* d_hi_planes, int64_t x_begin, int64_t x_end, void* stream) {
    NG_REQUIRE(p && p->kind == PK_FDLAP && p->ndim >= 2, "ng_fdlap_pack_halo_range: Laplacian plan (ndim >= 2) required");
    NG_REQUIRE(d_in, "ng_fdlap_pack_halo_range: null input");
    NG_REQUIRE(0 <= x_begin && x_begin <= x_end && x_end <= p->shape[0],
               "ng_fdlap_pack_halo_range: [%lld, %lld) is not inside axis 0", (long long)x_begin, (long long)x_end);
    const int nb = p->lap[p->ndim - 1].nb;
    const i64 n = p->shape[p->ndim - 1];
    const i64 per_plane = p->size / n / p->shape[0];           // rows per axis-0 plane
    const i64 r0 = x_begin * per_plane, rows = (x_end - x_begin) * per_plane;
    if (rows == 0) return NG_OK;
    const double* src = d_in + r0 * n;
    double* lo = d_lo_planes ? d_lo_planes + r0 * nb : nullptr;
    double* hi = d_hi_planes ? d_hi_planes + r0 * nb : nullptr;
    const bool al = (((uintptr_t)src | (uintptr_t)lo | (uintptr_t)hi) & 15) == 0;
    if (nb == 2 && n % 2 == 0 && al) {
        fd_pack_halo_last2_kernel<<<grid_for(rows, 256), 256, 0, (cudaStream_t)stream>>>(
            src, (double2*)lo, (double2*)hi, rows, n);
    } else {
        FuseDev F;
        memset(&F, 0, sizeof(F));
        fd_pack_halo_kernel<<<grid_for(rows * nb, 256), 256, 0, (cudaStream_t)stream>>>(src, lo, hi, rows, n, 1, nb, F);
    }
    NG_LAUNCH_CHECK();
    return NG_OK;
}

extern "C" int ng_fd_pack_halo(const ng_plan* p, const double* d_in, double* d_lo_planes,
                               double* d_hi_planes, void* stream) {
    NG_REQUIRE(p && (p->kind == PK_FD || p->kind == PK_FDLAP), "ng_fd_pack_halo: FD/FDLAP plan required");
    NG_REQUIRE(d_in, "ng_fd_pack_halo: null input");
    const int nb = p->kind == PK_FD ? p->fd.nb : p->lap[p->ndim - 1].nb;
    i64 outer = p->outer, n = p->n, inner = p->inner;
    if (p->kind == PK_FDLAP) { n = p->shape[p->ndim - 1]; outer = p->size / n; inner = 1; }
    const i64 total = outer * inner * nb;
    if (total == 0) return NG_OK;
    const bool al = (((uintptr_t)d_in | (uintptr_t)d_lo_planes | (uintptr_t)d_hi_planes) & 15) == 0;
    if (inner == 1 && nb == 2 && n % 2 == 0 && al) {
        fd_pack_halo_last2_kernel<<<grid_for(outer, 256), 256, 0, (cudaStream_t)stream>>>(
            d_in, (double2*)d_lo_planes, (double2*)d_hi_planes, outer, n);
        NG_LAUNCH_CHECK();
        return NG_OK;
    }
    FuseDev F;
    memset(&F, 0, sizeof(F));
    fd_pack_halo_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(
        d_in, d_lo_planes, d_hi_planes, outer, n, inner, nb, F);
    NG_LAUNCH_CHECK();
    return NG_OK;
}
