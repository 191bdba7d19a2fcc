// interp.cu -- multilinear interpolation of a meshed array at a tensor-product target grid or at a list
// of points (Interpolator / MultiGrid.transfer with method="linear").
//
// Replaces what reference numgrids/interpol.py:34-46, :60-75 delegates to SciPy for method="linear":
//   1-D  scipy.interpolate.interp1d(kind="linear") -> numpy.interp:
//            slope = (v[j+1]-v[j]) / (x[j+1]-x[j]);  out = slope*(xq-x[j]) + v[j]
//   2-D  RegularGridInterpolator's compiled path:
//            v00*(1-y0)*(1-y1) + v01*(1-y0)*y1 + v10*y0*(1-y1) + v11*y0*y1      (left to right)
//   N-D  RegularGridInterpolator._evaluate_linear: value = 0; for every hypercube corner (first axis
//            slowest, lower index first): value = value + v[corner] * (((1*w_0)*w_1)*...)
// with j / (i, y) from the interval search, y = (xq-x[i])/(x[i+1]-x[i]).  The host builds the per-axis
// tables (index, y, 1-y; 1-D: xq-x[j] and x[j+1]-x[j]) with the reference's NumPy expressions and
// passes them in; every product and sum here is separately rounded, in the order above, so the result
// is bit-identical to SciPy's (tests/test_interp.py).
// One thread per output point; `scattered` = the tables are indexed by the point number on every axis
// (a list of points) instead of by the point's multi-index (tensor-product target).
#include "common.cuh"

namespace {

struct InterpArgs {
    const int* idx[NG_MAX_DIM];        // lower interval index per target coordinate
    const double* y[NG_MAX_DIM];       // normalised distance          (1-D: xq - x[j])
    const double* omy[NG_MAX_DIM];     // 1 - y                        (1-D: x[j+1] - x[j])
    i64 sstride[NG_MAX_DIM];           // source strides (elements)
    i64 dshape[NG_MAX_DIM];            // target shape (tensor mode)
    i64 total;
    int ndim, scattered;
};

template <int ND>
__global__ void __launch_bounds__(256) interp_linear_kernel(const double* __restrict__ v, double* __restrict__ out, InterpArgs A) {
    for (i64 t = (i64)blockIdx.x * blockDim.x + threadIdx.x; t < A.total; t += (i64)gridDim.x * blockDim.x) {
        i64 ti[ND];
        if (A.scattered) {
#pragma unroll
            for (int a = 0; a < ND; ++a) ti[a] = t;
        } else {
            i64 r = t;
#pragma unroll
            for (int a = ND - 1; a > 0; --a) { ti[a] = r % A.dshape[a]; r /= A.dshape[a]; }
            ti[0] = r;
        }
        i64 off = 0;
#pragma unroll
        for (int a = 0; a < ND; ++a) off += (i64)A.idx[a][ti[a]] * A.sstride[a];
        double res;
        if (ND == 1) {
            const double lo = v[off], hi = v[off + A.sstride[0]];
            const double dx = A.y[0][ti[0]], den = A.omy[0][ti[0]];
            if (den == 0.0) res = hi;                        // xq is the last knot (host flag): v[-1] exactly
            else if (dx == 0.0) res = lo;                    // xq is a knot: numpy.interp returns v[j] exactly
            else res = __dadd_rn(__dmul_rn(__ddiv_rn(__dsub_rn(hi, lo), den), dx), lo);
        } else if (ND == 2) {
            const double y0 = A.y[0][ti[0]], y1 = A.y[1][ti[1]], z0 = A.omy[0][ti[0]], z1 = A.omy[1][ti[1]];
            const double v00 = v[off], v01 = v[off + A.sstride[1]], v10 = v[off + A.sstride[0]],
                         v11 = v[off + A.sstride[0] + A.sstride[1]];
            res = __dmul_rn(__dmul_rn(v00, z0), z1);
            res = __dadd_rn(res, __dmul_rn(__dmul_rn(v01, z0), y1));
            res = __dadd_rn(res, __dmul_rn(__dmul_rn(v10, y0), z1));
            res = __dadd_rn(res, __dmul_rn(__dmul_rn(v11, y0), y1));
        } else {
            double w[ND][2];
#pragma unroll
            for (int a = 0; a < ND; ++a) { w[a][0] = A.omy[a][ti[a]]; w[a][1] = A.y[a][ti[a]]; }
            res = 0.0;
#pragma unroll
            for (int c = 0; c < (1 << ND); ++c) {            // corner bits: axis 0 is the most significant
                double weight = 1.0;
                i64 o = off;
#pragma unroll
                for (int a = 0; a < ND; ++a) {
                    const int b = (c >> (ND - 1 - a)) & 1;
                    weight = __dmul_rn(weight, w[a][b]);
                    if (b) o += A.sstride[a];
                }
                res = __dadd_rn(res, __dmul_rn(v[o], weight));
            }
        }
        out[t] = res;
    }
}

}  // namespace

extern "C" int ng_interp_linear(int ndim, const int64_t* src_shape, const int64_t* dst_shape, int scattered,
                                const int32_t* const* d_idx, const double* const* d_y,
                                const double* const* d_omy, const double* d_values, double* d_out,
                                void* stream) {
    NG_REQUIRE(ndim >= 1 && ndim <= NG_MAX_DIM && src_shape && dst_shape, "ng_interp_linear: bad shape");
    NG_REQUIRE(d_idx && d_y && d_omy && d_values && d_out, "ng_interp_linear: NULL buffer");
    InterpArgs A;
    memset(&A, 0, sizeof(A));
    A.ndim = ndim;
    A.scattered = scattered ? 1 : 0;
    i64 st = 1, total = 1;
    for (int a = ndim - 1; a >= 0; --a) {
        NG_REQUIRE(src_shape[a] >= 2, "ng_interp_linear: source axis %d needs at least 2 points", a);
        A.sstride[a] = st;
        st *= src_shape[a];
    }
    for (int a = 0; a < ndim; ++a) {
        NG_REQUIRE(d_idx[a] && d_y[a] && d_omy[a], "ng_interp_linear: table %d is NULL", a);
        NG_REQUIRE(dst_shape[a] >= 0, "ng_interp_linear: negative target extent");
        A.idx[a] = d_idx[a];
        A.y[a] = d_y[a];
        A.omy[a] = d_omy[a];
        A.dshape[a] = dst_shape[a];
        if (!scattered) total *= dst_shape[a];
    }
    if (scattered) total = dst_shape[0];
    A.total = total;
    if (total == 0) return NG_OK;
    i64 grid = (total + 255) / 256;
    if (grid > 148LL * 32) grid = 148LL * 32;
    cudaStream_t s = (cudaStream_t)stream;
    switch (ndim) {
        case 1: interp_linear_kernel<1><<<(unsigned)grid, 256, 0, s>>>(d_values, d_out, A); break;
        case 2: interp_linear_kernel<2><<<(unsigned)grid, 256, 0, s>>>(d_values, d_out, A); break;
        case 3: interp_linear_kernel<3><<<(unsigned)grid, 256, 0, s>>>(d_values, d_out, A); break;
        default: interp_linear_kernel<4><<<(unsigned)grid, 256, 0, s>>>(d_values, d_out, A); break;
    }
    NG_LAUNCH_CHECK();
    return NG_OK;
}

// -------------------------------------------------------------------------------------------------
// Tensor-product B-spline evaluation: Interpolator with method "cubic" (N-D: SciPy's NdBSpline behind
// RegularGridInterpolator) and "cubic" / "quadratic" in 1-D (interp1d -> make_interp_spline -> BSpline),
// reference numgrids/interpol.py:34-46.  The spline coefficients are the host's (SciPy's own solve, as in
// the reference); per target coordinate and axis the host passes the first coefficient index and the k+1
// non-zero basis values (SciPy's de Boor recursion).  Each point sums its (k+1)^d terms in SciPy's order:
// last axis fastest, factor = ((1*b_0)*b_1)*..., value += c * factor -- bit-identical to SciPy's evaluation.
// -------------------------------------------------------------------------------------------------
namespace {

struct SplineArgs {
    const int* first[NG_MAX_DIM];      // first coefficient index per target coordinate
    const double* basis[NG_MAX_DIM];   // [ntargets][nb] basis values
    i64 cstride[NG_MAX_DIM];           // coefficient strides (elements)
    i64 dshape[NG_MAX_DIM];
    int nb[NG_MAX_DIM];                // k + 1 per axis
    i64 total;
    int ndim, scattered;
};

__global__ void __launch_bounds__(256) interp_bspline_kernel(const double* __restrict__ c, double* __restrict__ out, SplineArgs A) {
    for (i64 t = (i64)blockIdx.x * blockDim.x + threadIdx.x; t < A.total; t += (i64)gridDim.x * blockDim.x) {
        i64 ti[NG_MAX_DIM] = {0, 0, 0, 0};
        if (A.scattered) {
            for (int a = 0; a < A.ndim; ++a) ti[a] = t;
        } else {
            i64 r = t;
            for (int a = A.ndim - 1; a > 0; --a) { ti[a] = r % A.dshape[a]; r /= A.dshape[a]; }
            ti[0] = r;
        }
        i64 base = 0;
        const double* b[NG_MAX_DIM] = {nullptr, nullptr, nullptr, nullptr};
        int terms = 1;
        for (int a = 0; a < A.ndim; ++a) {
            base += (i64)A.first[a][ti[a]] * A.cstride[a];
            b[a] = A.basis[a] + ti[a] * A.nb[a];
            terms *= A.nb[a];
        }
        double res = 0.0;
        for (int e = 0; e < terms; ++e) {                    // flat tap index, last axis fastest
            int j[NG_MAX_DIM] = {0, 0, 0, 0};
            int r = e;
            for (int a = A.ndim - 1; a > 0; --a) { j[a] = r % A.nb[a]; r /= A.nb[a]; }
            j[0] = r;
            double factor = 1.0;
            i64 o = base;
            for (int a = 0; a < A.ndim; ++a) {
                factor = __dmul_rn(factor, b[a][j[a]]);
                o += (i64)j[a] * A.cstride[a];
            }
            res = __dadd_rn(res, __dmul_rn(c[o], factor));
        }
        out[t] = res;
    }
}

}  // namespace

extern "C" int ng_interp_bspline(int ndim, const int64_t* coef_shape, const int64_t* dst_shape, int scattered,
                                 const int32_t* nbasis, const int32_t* const* d_first,
                                 const double* const* d_basis, const double* d_coef, double* d_out,
                                 void* stream) {
    NG_REQUIRE(ndim >= 1 && ndim <= NG_MAX_DIM && coef_shape && dst_shape && nbasis, "ng_interp_bspline: bad shape");
    NG_REQUIRE(d_first && d_basis && d_coef && d_out, "ng_interp_bspline: NULL buffer");
    SplineArgs A;
    memset(&A, 0, sizeof(A));
    A.ndim = ndim;
    A.scattered = scattered ? 1 : 0;
    i64 st = 1, total = 1;
    for (int a = ndim - 1; a >= 0; --a) {
        NG_REQUIRE(nbasis[a] >= 1 && nbasis[a] <= 8, "ng_interp_bspline: degree + 1 must be in 1..8 (axis %d)", a);
        NG_REQUIRE(coef_shape[a] >= nbasis[a], "ng_interp_bspline: axis %d has fewer coefficients than basis functions", a);
        A.cstride[a] = st;
        st *= coef_shape[a];
    }
    for (int a = 0; a < ndim; ++a) {
        NG_REQUIRE(d_first[a] && d_basis[a], "ng_interp_bspline: table %d is NULL", a);
        NG_REQUIRE(dst_shape[a] >= 0, "ng_interp_bspline: negative target extent");
        A.first[a] = d_first[a];
        A.basis[a] = d_basis[a];
        A.nb[a] = nbasis[a];
        A.dshape[a] = dst_shape[a];
        if (!scattered) total *= dst_shape[a];
    }
    if (scattered) total = dst_shape[0];
    A.total = total;
    if (total == 0) return NG_OK;
    i64 grid = (total + 255) / 256;
    if (grid > 148LL * 32) grid = 148LL * 32;
    interp_bspline_kernel<<<(unsigned)grid, 256, 0, (cudaStream_t)stream>>>(d_coef, d_out, A);
    NG_LAUNCH_CHECK();
    return NG_OK;
}
