// quad.cu -- tensor-product quadrature: the integral of a meshed array over the whole grid
// (SURVEY.md 8f, N3).
//
// Replaces reference numgrids/integration.py:44-72 (Integral.__call__).  There, per axis, the array
// is either summed and scaled by the spacing (periodic axis) or multiplied by the Kronecker-lifted
// inverse of the first-derivative matrix with its first row/column removed, of which only the LAST
// row is kept -- i.e. a fixed weight vector per axis (w[0] = 0, w[1:] = last row of inv(D[1:,1:])).
// The whole operator is therefore  sum_{i0..} w0[i0] w1[i1] ... f[i0, i1, ...]:  ONE streaming pass
// over f, 8 B/point, HBM-bound.  The weight vectors are computed by the host exactly as the
// reference does (scipy.sparse.linalg.inv of the same matrix) and passed in.
//
// Mapping: the array is [rows][n_last]; a warp takes one (row, 1024-point segment) tile, unravels the
// row index once (product of the outer weights), lanes read 128-bit pairs along the row (coalesced)
// and accumulate f * w_last with FMA; tile sum * outer weight goes to the lane's running total.
// Reduction: warp shuffles -> shared memory -> one partial per CTA -> a second one-CTA kernel sums
// the partials in a fixed order, so the result is deterministic for a given shape (no atomics).
// The summation order differs from the reference's axis-by-axis sequential sums; the parity bar for
// this path is therefore a tolerance (tests/test_integral.py), not bit-exactness.
#include "common.cuh"

namespace {

constexpr int QUAD_THREADS = 256;
constexpr int QUAD_SEG = 1024;          // points per warp tile along the last axis
constexpr int QUAD_MAX_CTAS = 148 * 8;

struct QuadArgs {
    const double* w[NG_MAX_DIM];        // per-axis weights, padded on the left with unit axes
    i64 shape[NG_MAX_DIM];
    i64 rows, n_last, segs;
};

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    return v;
}

template <bool VEC2>
__global__ void __launch_bounds__(QUAD_THREADS)
quad_partial_kernel(const double* __restrict__ f, QuadArgs A, double* __restrict__ partials) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const i64 warps_total = (i64)gridDim.x * (QUAD_THREADS / 32);
    const i64 tiles = A.rows * A.segs;
    const double* __restrict__ wl = A.w[NG_MAX_DIM - 1];
    double total = 0.0;
    for (i64 t = (i64)blockIdx.x * (QUAD_THREADS / 32) + warp; t < tiles; t += warps_total) {
        const i64 r = t / A.segs, s = t - r * A.segs;
        // outer weight of this row (axes 0..2 of the padded 4-D view)
        i64 q = r;
        const i64 i2 = q % A.shape[2]; q /= A.shape[2];
        const i64 i1 = q % A.shape[1]; q /= A.shape[1];
        double wrow = A.w[2] ? A.w[2][i2] : 1.0;            // absent (padded) axes weigh 1
        if (A.w[1]) wrow *= A.w[1][i1];
        if (A.w[0]) wrow *= A.w[0][q];
        const i64 c0 = s * QUAD_SEG;
        const i64 c1 = (c0 + QUAD_SEG < A.n_last) ? c0 + QUAD_SEG : A.n_last;
        const double* __restrict__ row = f + r * A.n_last;
        double acc0 = 0.0, acc1 = 0.0;
        if (VEC2) {
#pragma unroll 4
            for (i64 c = c0 + 2 * lane; c < c1; c += 64) {
                const double2 v = __ldg(reinterpret_cast<const double2*>(row + c));
                const double2 w = __ldg(reinterpret_cast<const double2*>(wl + c));
                acc0 = fma(v.x, w.x, acc0);
                acc1 = fma(v.y, w.y, acc1);
            }
        } else {
#pragma unroll 4
            for (i64 c = c0 + lane; c < c1; c += 32) acc0 = fma(__ldg(row + c), __ldg(wl + c), acc0);
        }
        total = fma(acc0 + acc1, wrow, total);
    }
    __shared__ double sm[QUAD_THREADS / 32];
    total = warp_sum(total);
    if (lane == 0) sm[warp] = total;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < QUAD_THREADS / 32; ++k) s += sm[k];
        partials[blockIdx.x] = s;
    }
}

__global__ void __launch_bounds__(1024) quad_final_kernel(const double* __restrict__ partials, int n,
                                                          double* __restrict__ out) {
    __shared__ double sm[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double v = 0.0;
    for (int k = threadIdx.x; k < n; k += 1024) v += partials[k];
    v = warp_sum(v);
    if (lane == 0) sm[warp] = v;
    __syncthreads();
    if (warp == 0) {
        v = warp_sum(sm[lane]);
        if (lane == 0) out[0] = v;
    }
}

}  // namespace

extern "C" int64_t ng_quad_workspace_doubles(void) { return QUAD_MAX_CTAS; }

extern "C" int ng_quad_apply(int ndim, const int64_t* shape, const double* const* d_w,
                             const double* d_f, double* d_work, double* d_out, void* stream) {
    NG_REQUIRE(ndim >= 1 && ndim <= NG_MAX_DIM && shape, "ng_quad_apply: bad shape");
    NG_REQUIRE(d_w && d_f && d_work && d_out, "ng_quad_apply: NULL buffer");
    QuadArgs A;
    const int pad = NG_MAX_DIM - ndim;
    i64 rows = 1;
    for (int a = 0; a < NG_MAX_DIM; ++a) {
        if (a < pad) {
            A.shape[a] = 1;
            A.w[a] = nullptr;
        } else {
            NG_REQUIRE(shape[a - pad] >= 1, "ng_quad_apply: shape[%d] must be positive", a - pad);
            NG_REQUIRE(d_w[a - pad], "ng_quad_apply: weight table %d is NULL", a - pad);
            A.shape[a] = shape[a - pad];
            A.w[a] = d_w[a - pad];
        }
        if (a < NG_MAX_DIM - 1) rows *= A.shape[a];
    }
    A.rows = rows;
    A.n_last = A.shape[NG_MAX_DIM - 1];
    A.segs = (A.n_last + QUAD_SEG - 1) / QUAD_SEG;
    const i64 tiles = A.rows * A.segs;
    i64 ctas = (tiles + QUAD_THREADS / 32 - 1) / (QUAD_THREADS / 32);
    if (ctas > QUAD_MAX_CTAS) ctas = QUAD_MAX_CTAS;
    if (ctas < 1) ctas = 1;
    cudaStream_t st = (cudaStream_t)stream;
    // 128-bit path: every row start and the last-axis weight table must be 16-byte aligned
    const bool vec2 = (A.n_last % 2 == 0) && (((uintptr_t)d_f) % 16 == 0) && (((uintptr_t)A.w[NG_MAX_DIM - 1]) % 16 == 0);
    if (vec2)
        quad_partial_kernel<true><<<(unsigned)ctas, QUAD_THREADS, 0, st>>>(d_f, A, d_work);
    else
        quad_partial_kernel<false><<<(unsigned)ctas, QUAD_THREADS, 0, st>>>(d_f, A, d_work);
    NG_LAUNCH_CHECK();
    quad_final_kernel<<<1, 1024, 0, st>>>(d_work, (int)ctas, d_out);
    NG_LAUNCH_CHECK();
    return NG_OK;
}

// -------------------------------------------------------------------------------------------------
// Norm of the difference of two arrays in one pass (16 B/point): the scalar the reference's
// ErrorEstimator._compute_norm takes of `f_current - f_interpolated` (numgrids/amr.py:96-103):
//   kind 0  max(|a - b|)          kind 1  sqrt(mean((a - b)**2))          kind 2  mean(|a - b|)
// The difference and the square are rounded like NumPy's (one DADD, one DMUL); the sums use the same
// fixed-order two-stage reduction as the quadrature above (max is exact, the sums agree with NumPy's
// pairwise summation to rounding).
// -------------------------------------------------------------------------------------------------
namespace {

template <int KIND>
__device__ __forceinline__ double norm_term(double a, double b) {
    const double d = __dsub_rn(a, b);
    return KIND == 1 ? __dmul_rn(d, d) : fabs(d);
}
template <int KIND>
__device__ __forceinline__ double norm_join(double x, double y) {
    if (KIND == 0) return (x > y || isnan(x)) ? x : y;       // np.max propagates NaN
    return x + y;
}

template <int KIND, bool VEC2>
__global__ void __launch_bounds__(QUAD_THREADS)
diff_norm_partial_kernel(const double* __restrict__ a, const double* __restrict__ b, i64 n,
                         double* __restrict__ partials) {
    const i64 stride = (i64)gridDim.x * QUAD_THREADS;
    double acc = 0.0;
    if (VEC2) {
        const i64 n2 = n >> 1;
        for (i64 i = (i64)blockIdx.x * QUAD_THREADS + threadIdx.x; i < n2; i += stride) {
            const double2 x = __ldg(reinterpret_cast<const double2*>(a) + i);
            const double2 y = __ldg(reinterpret_cast<const double2*>(b) + i);
            acc = norm_join<KIND>(acc, norm_term<KIND>(x.x, y.x));
            acc = norm_join<KIND>(acc, norm_term<KIND>(x.y, y.y));
        }
    } else {
        for (i64 i = (i64)blockIdx.x * QUAD_THREADS + threadIdx.x; i < n; i += stride)
            acc = norm_join<KIND>(acc, norm_term<KIND>(__ldg(a + i), __ldg(b + i)));
    }
    __shared__ double sm[QUAD_THREADS / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc = norm_join<KIND>(acc, __shfl_down_sync(0xffffffffu, acc, o));
    if (lane == 0) sm[warp] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = sm[0];
#pragma unroll
        for (int k = 1; k < QUAD_THREADS / 32; ++k) s = norm_join<KIND>(s, sm[k]);
        partials[blockIdx.x] = s;
    }
}

template <int KIND>
__global__ void __launch_bounds__(1024) diff_norm_final_kernel(const double* __restrict__ partials, int np,
                                                               i64 n, double* __restrict__ out) {
    __shared__ double sm[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double v = 0.0;
    for (int k = threadIdx.x; k < np; k += 1024) v = norm_join<KIND>(v, partials[k]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = norm_join<KIND>(v, __shfl_down_sync(0xffffffffu, v, o));
    if (lane == 0) sm[warp] = v;
    __syncthreads();
    if (warp == 0) {
        v = sm[lane];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v = norm_join<KIND>(v, __shfl_down_sync(0xffffffffu, v, o));
        if (lane == 0) {
            if (KIND == 1) v = sqrt(__ddiv_rn(v, (double)n));
            if (KIND == 2) v = __ddiv_rn(v, (double)n);
            out[0] = v;
        }
    }
}

template <int KIND>
int diff_norm_launch(const double* a, const double* b, i64 n, double* work, double* out, cudaStream_t st) {
    i64 ctas = (n / 2 + QUAD_THREADS - 1) / QUAD_THREADS;
    if (ctas > QUAD_MAX_CTAS) ctas = QUAD_MAX_CTAS;
    if (ctas < 1) ctas = 1;
    const bool vec2 = n % 2 == 0 && ((((uintptr_t)a) | ((uintptr_t)b)) % 16) == 0;
    if (vec2) diff_norm_partial_kernel<KIND, true><<<(unsigned)ctas, QUAD_THREADS, 0, st>>>(a, b, n, work);
    else diff_norm_partial_kernel<KIND, false><<<(unsigned)ctas, QUAD_THREADS, 0, st>>>(a, b, n, work);
    NG_LAUNCH_CHECK();
    diff_norm_final_kernel<KIND><<<1, 1024, 0, st>>>(work, (int)ctas, n, out);
    NG_LAUNCH_CHECK();
    return NG_OK;
}

}  // namespace

extern "C" int ng_diff_norm(int64_t n, const double* d_a, const double* d_b, int kind, double* d_work,
                            double* d_out, void* stream) {
    NG_REQUIRE(n >= 1, "ng_diff_norm: empty arrays");
    NG_REQUIRE(d_a && d_b && d_work && d_out, "ng_diff_norm: NULL buffer");
    cudaStream_t st = (cudaStream_t)stream;
    switch (kind) {
        case NG_NORM_MAX: return diff_norm_launch<0>(d_a, d_b, n, d_work, d_out, st);
        case NG_NORM_L2: return diff_norm_launch<1>(d_a, d_b, n, d_work, d_out, st);
        case NG_NORM_MEAN: return diff_norm_launch<2>(d_a, d_b, n, d_work, d_out, st);
    }
    ng_set_error("ng_diff_norm: unknown norm kind %d", kind);
    return NG_EINVAL;
}
