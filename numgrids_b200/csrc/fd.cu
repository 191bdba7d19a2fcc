// fd.cu -- finite-difference stencil apply (FiniteDifferenceDiff / LogDiff).
//
// Replaces findiff.FinDiff(axis, h, order, acc)(f) behind reference numgrids/diff.py:88 and
// :259-262.  Arithmetic contract (bit-exact with the reference):
//   acc = 0.0; for k in stored tap order: acc = acc + (w_k * f[i+o_k])   (DMUL, DADD -- never FMA)
//   out = acc * inv_h_pow            [LogDiff: then out / x[i]]
// central taps on [nb, n-nb), forward on [0, nb), backward on [n-nb, n).
//
// Layout: the array is viewed as [outer][n][inner] around the derivative axis.
//   * inner >= 2 (axis is not last): one thread owns VEC (=2 when inner is even, 128-bit
//     LDG/STG) consecutive `inner` elements of one output row; the taps of a warp are
//     full coalesced row segments n*inner apart that stay L1/L2 resident between rows.
//   * inner == 1 (last axis): lanes run along the pencil; the +-p neighbours come from the
//     same and the two adjacent 128-byte lines (L1 hits).
// Algorithmic traffic 16 B/point (read f once, write out once).
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

int ng_fd_launch(const ng_plan* p, const double* in, double* out, const FuseDev& F,
                 const double* lo, const double* hi, cudaStream_t st) {
    const bool halo = (lo != nullptr) || (hi != nullptr);
    const bool aligned = (((uintptr_t)in | (uintptr_t)out) & 15) == 0;
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
