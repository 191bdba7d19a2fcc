// bc.cu -- array-level boundary conditions applied in place on a device array.
//
// Replaces the NumPy slicing of reference numgrids/boundary.py:
//   DirichletBC.apply (:157-166)   u[face] = g
//   NeumannBC.apply   (:216-245)   u[face] = u[adjacent] + (h * normal_sign) * g     (first order)
// One thread per face point; the face is the array with `axis` removed (C order), so face point
// (o, c) of the [outer][n][inner] view sits at (o*n + i_face)*inner + c.  HBM-bound and tiny:
// 8 B/face point read (+8 for Neumann) and 8 B written -- it exists so that a time-stepping loop
// can keep its field on the device between derivative applies.
#include "common.cuh"

namespace {

template <bool NEUMANN>
__global__ void __launch_bounds__(256)
bc_apply_kernel(double* __restrict__ u, const double* __restrict__ g, double g_scalar, double hs,
                i64 outer, i64 n, i64 inner, i64 i_face, i64 i_adj) {
    const i64 total = outer * inner;
    for (i64 idx = (i64)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (i64)gridDim.x * blockDim.x) {
        const i64 o = idx / inner, c = idx - o * inner;
        const double gv = g ? g[idx] : g_scalar;
        double v = gv;
        if (NEUMANN) v = __dadd_rn(u[(o * n + i_adj) * inner + c], __dmul_rn(hs, gv));
        u[(o * n + i_face) * inner + c] = v;
    }
}

}  // namespace

extern "C" int ng_bc_apply(int ndim, const int64_t* shape, int axis, int side, int kind, double hs,
                           const double* d_g, double g_scalar, double* d_u, void* stream) {
    NG_REQUIRE(ndim >= 1 && ndim <= NG_MAX_DIM && shape, "ng_bc_apply: bad shape");
    NG_REQUIRE(axis >= 0 && axis < ndim, "ng_bc_apply: axis %d out of range", axis);
    NG_REQUIRE(side == 0 || side == 1, "ng_bc_apply: side must be 0 (low) or 1 (high)");
    NG_REQUIRE(kind == NG_BC_DIRICHLET || kind == NG_BC_NEUMANN, "ng_bc_apply: unknown kind %d", kind);
    NG_REQUIRE(d_u, "ng_bc_apply: NULL array");
    i64 outer = 1, inner = 1;
    for (int a = 0; a < ndim; ++a) NG_REQUIRE(shape[a] >= 1, "ng_bc_apply: shape[%d] must be positive", a);
    for (int a = 0; a < axis; ++a) outer *= shape[a];
    for (int a = axis + 1; a < ndim; ++a) inner *= shape[a];
    const i64 n = shape[axis];
    NG_REQUIRE(kind == NG_BC_DIRICHLET || n >= 2, "ng_bc_apply: a Neumann face needs an adjacent layer");
    const i64 total = outer * inner;
    if (total == 0) return NG_OK;
    const i64 i_face = side == 0 ? 0 : n - 1, i_adj = side == 0 ? 1 : n - 2;
    i64 gx = (total + 255) / 256;
    if (gx > 148 * 32) gx = 148 * 32;
    cudaStream_t st = (cudaStream_t)stream;
    if (kind == NG_BC_NEUMANN)
        bc_apply_kernel<true><<<(unsigned)gx, 256, 0, st>>>(d_u, d_g, g_scalar, hs, outer, n, inner, i_face, i_adj);
    else
        bc_apply_kernel<false><<<(unsigned)gx, 256, 0, st>>>(d_u, d_g, g_scalar, hs, outer, n, inner, i_face, i_adj);
    NG_LAUNCH_CHECK();
    return NG_OK;
}
