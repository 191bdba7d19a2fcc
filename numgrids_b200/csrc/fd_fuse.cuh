// fd_fuse.cuh -- the fused elementwise work (ng_fuse) of the contiguous-axis finite-difference kernels (fd_row.cu,
// fd_tile.cu) as a template parameter MODE, shared by both translation units.
#pragma once
#include "common.cuh"

namespace {

__device__ __forceinline__ double2 ldg128(const double* p) {
    return __ldg(reinterpret_cast<const double2*>(p));
}

// Fused elementwise work (ng_fuse) as a template parameter MODE: 0 = none; -1 = every step tested at run time
// (any combination); > 0 = the steps present, as a compile-time bit mask, for the combinations the curvilinear
// composites of curv.cu issue -- no flag tests, no offset arithmetic for absent operands (the run-time form cost
// 134 thread-instructions per point against 57 for the plain kernel, profiles/r01_fd_row_fused_ncu_summary.txt).
enum : int { FM_PRE = 1, FM_MINUEND = 2, FM_POST = 4, FM_ADDEND = 8, FM_ADDZERO = 16, FM_DIV = 32, FM_FINITE = 64 };
template <int MODE> __device__ __forceinline__ bool fm_pre(const FuseDev& F) { return MODE < 0 ? F.pre.ptr != nullptr : (MODE & FM_PRE) != 0; }
template <int MODE> __device__ __forceinline__ bool fm_post(const FuseDev& F) { return MODE < 0 ? F.post.ptr != nullptr : (MODE & FM_POST) != 0; }
template <int MODE> __device__ __forceinline__ bool fm_div(const FuseDev& F) { return MODE < 0 ? F.div.ptr != nullptr : (MODE & FM_DIV) != 0; }
// correctly rounded divide, out of line: ~40 instructions that would otherwise be inlined 8 x 6 times per kernel
__device__ __noinline__ double fd_row_div(double a, double b) { return __ddiv_rn(a, b); }
// the tail of ng_fuse (same steps and order as fuse_tail_off in common.cuh)
template <int MODE>
__device__ __forceinline__ double fm_tail(const FuseDev& F, double d, i64 lin, i64 post_off, i64 div_off) {
    if (MODE < 0) return fuse_tail_off(F, d, lin, post_off, div_off);
    if (MODE & FM_MINUEND) d = __dsub_rn(F.minuend[lin], d);
    if (MODE & FM_POST) d = __dmul_rn(F.post.ptr[post_off], d);
    if (MODE & FM_ADDEND) d = __dadd_rn(F.addend[lin], d);
    else if (MODE & FM_ADDZERO) d = __dadd_rn(0.0, d);
    if (MODE & FM_DIV) d = fd_row_div(d, F.div.ptr[div_off]);
    if (MODE & FM_FINITE) d = ng_nonfinite(d) ? 0.0 : d;
    return d;
}

// the same with the operands of the pair already in registers (compile-time modes): coefficient pairs come from one
// hoisted scalar when the table is constant along the row (spherical / cylindrical metrics along z), full-shape
// operands (minuend / addend) from one aligned 128-bit load per pair
template <int MODE>
__device__ __forceinline__ double fm_tail_vals(double d, double mn, double pc, double ad, double dc) {
    if (MODE & FM_MINUEND) d = __dsub_rn(mn, d);
    if (MODE & FM_POST) d = __dmul_rn(pc, d);
    if (MODE & FM_ADDEND) d = __dadd_rn(ad, d);
    else if (MODE & FM_ADDZERO) d = __dadd_rn(0.0, d);
    if (MODE & FM_DIV) d = fd_row_div(d, dc);
    if (MODE & FM_FINITE) d = ng_nonfinite(d) ? 0.0 : d;
    return d;
}
__device__ __forceinline__ double2 coef_pair(const double* base, i64 off0, int pos, i64 stride, double hoisted) {
    if (stride == 0) return make_double2(hoisted, hoisted);          // (no offset arithmetic on this path)
    const i64 off = off0 + pos * stride;
    return make_double2(__ldg(base + off), __ldg(base + off + stride));
}

// the compile-time fuse patterns of the curvilinear composites (curv.cu)
inline int fuse_mask(const FuseDev& F) {
    return (F.pre.ptr ? FM_PRE : 0) | (F.minuend ? FM_MINUEND : 0) | (F.post.ptr ? FM_POST : 0) | (F.addend ? FM_ADDEND : 0) |
           (!F.addend && F.add_zero ? FM_ADDZERO : 0) | (F.div.ptr ? FM_DIV : 0) | (F.finite ? FM_FINITE : 0);
}
// the twelve patterns curv.cu issues: X(mask)
#define NG_FD_FUSE_MODES(X)                                                                                      \
    X(FM_POST)                                   /* laplacian: work = (J/h^2) * D f                         */   \
    X(FM_ADDZERO)                                /*            out = 0.0 + D work                           */   \
    X(FM_ADDEND)                                 /*            out = out + D work                           */   \
    X(FM_ADDZERO | FM_DIV | FM_FINITE)           /*            ... / J, finite (1-D grids)                  */   \
    X(FM_ADDEND | FM_DIV | FM_FINITE)                                                                            \
    X(FM_POST | FM_FINITE)                       /* gradient                                                */   \
    X(FM_PRE | FM_ADDZERO)                       /* divergence                                              */   \
    X(FM_PRE | FM_ADDEND)                                                                                        \
    X(FM_PRE | FM_ADDZERO | FM_DIV | FM_FINITE)                                                                  \
    X(FM_PRE | FM_ADDEND | FM_DIV | FM_FINITE)                                                                   \
    X(FM_PRE)                                    /* curl: out = D(h v)                                      */   \
    X(FM_PRE | FM_MINUEND | FM_POST | FM_FINITE) /*       out = finite(c * (out - D(h v)))                  */

}  // namespace
