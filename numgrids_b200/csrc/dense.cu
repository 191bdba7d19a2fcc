// dense.cu -- ordered dense contraction along one axis (ChebyshevDiff; non-power-of-two FFTDiff).
//
// Replaces `self._diff_matrix * f.reshape(-1)` (reference numgrids/diff.py:194-197): the
// Kronecker-embedded sparse matvec is, per pencil, out_i = sum_j M[i][j] in_j.  Parity at 1e-12
// needs the reference's exact summation order: accumulate from 0.0, ONE j
// at a time, ascending or descending, product and sum rounded separately (DMUL + DADD, no FMA, no
// split-K, no tree/warp reduction, no tensor cores).  That makes this an FP64-pipe-bound ordered
// DGEMM: 2n FP64 instructions per point (roofline = 148 SM x 64 lanes x clk), HBM traffic 16 B/pt.
//
// Mapping: Out[n x C] = M[n x n] . F[n x C] with C = outer*inner flattened columns; column
// (o,c) element k lives at ((o*n + k)*inner + c).  CTA tile BM x BN outputs, K panels of BK
// staged in shared memory (M panel from the k-major copy Mt so that the row values a thread
// needs are contiguous), TM x TN register tile per thread; the panel loop and the in-panel loop
// run in the plan's direction.  LAST (inner == 1): pencils are contiguous, so the F panel is
// loaded coalesced along k and transposed in shared memory, and lanes run along i for the store.
// The FMA variant (use_fma) serves the dense form of the periodic spectral matrix
// (numgrids/diff.py:145-157) whose tolerance (1e-11) allows contraction.
#include "common.cuh"

namespace {

constexpr int BM = 64, BN = 64, BK = 16, TM = 4, TN = 4, NT = 256;

template <bool LAST, bool DESC, bool FMA, bool FUSED>
__global__ void __launch_bounds__(NT)
dense_apply_kernel(const double* __restrict__ in, double* __restrict__ out,
                   const double* __restrict__ Mt, i64 outer, int n, i64 inner, int axis,
                   FuseDev F) {
    __shared__ double Ms[BK][BM + 2];
    __shared__ double Fs[BK][BN + 2];
    const i64 ncols = outer * inner;
    const i64 col0 = (i64)blockIdx.x * BN;
    const int row0 = blockIdx.y * BM;
    const int tid = threadIdx.x;
    const int tx = tid % 16, ty = tid / 16;
    const int ti = (LAST ? tx : ty) * TM;   // first local output row (i) of this thread
    const int tc = (LAST ? ty : tx) * TN;   // first local column

    double acc[TM][TN];
#pragma unroll
    for (int a = 0; a < TM; ++a)
#pragma unroll
        for (int b = 0; b < TN; ++b) acc[a][b] = 0.0;

    const int npanel = (n + BK - 1) / BK;
    for (int pp = 0; pp < npanel; ++pp) {
        const int p = DESC ? (npanel - 1 - pp) : pp;
        const int k0 = p * BK;
        const int kmax = min(BK, n - k0);
        // ---- stage M panel: Ms[kk][i] = M[row0+i][k0+kk] = Mt[(k0+kk)*n + row0+i]
        for (int e = tid; e < BK * BM; e += NT) {
            const int kk = e / BM, i = e % BM;
            double v = 0.0;
            if (kk < kmax && row0 + i < n) v = Mt[(i64)(k0 + kk) * n + row0 + i];
            Ms[kk][i] = v;
        }
        // ---- stage F panel: Fs[kk][c] = pre * in[col0+c, k0+kk]
        for (int e = tid; e < BK * BN; e += NT) {
            int kk, c;
            if (LAST) { kk = e % BK; c = e / BK; }   // coalesced along k (contiguous pencils)
            else      { c = e % BN; kk = e / BN; }   // coalesced along the inner axis
            double v = 0.0;
            const i64 col = col0 + c;
            if (kk < kmax && col < ncols) {
                const i64 o = LAST ? col : col / inner;
                const i64 cc = LAST ? 0 : col - o * inner;
                const i64 lin = (o * n + k0 + kk) * inner + cc;
                v = in[lin];
                if (FUSED && F.pre.ptr) {
                    const Idx4 ix = unravel4(lin, F.shape);
                    v = __dmul_rn(F.pre.ptr[coef_off(F.pre, ix)], v);
                }
            }
            Fs[kk][c] = v;
        }
        __syncthreads();
        for (int q = 0; q < kmax; ++q) {
            const int kk = DESC ? (kmax - 1 - q) : q;
            double a[TM], b[TN];
#pragma unroll
            for (int x = 0; x < TM; ++x) a[x] = Ms[kk][ti + x];
#pragma unroll
            for (int y = 0; y < TN; ++y) b[y] = Fs[kk][tc + y];
#pragma unroll
            for (int x = 0; x < TM; ++x)
#pragma unroll
                for (int y = 0; y < TN; ++y) {
                    if (FMA) acc[x][y] = fma(a[x], b[y], acc[x][y]);
                    else acc[x][y] = __dadd_rn(acc[x][y], __dmul_rn(a[x], b[y]));
                }
        }
        __syncthreads();
    }
    // ---- epilogue
#pragma unroll
    for (int y = 0; y < TN; ++y) {
        const i64 col = col0 + tc + y;
        if (col >= ncols) continue;
        const i64 o = LAST ? col : col / inner;
        const i64 cc = LAST ? 0 : col - o * inner;
#pragma unroll
        for (int x = 0; x < TM; ++x) {
            const int i = row0 + ti + x;
            if (i >= n) continue;
            const i64 lin = (o * n + i) * inner + cc;
            double d = acc[x][y];
            if (FUSED) {
                const Idx4 ix = unravel4(lin, F.shape);
                d = fuse_tail(F, d, lin, ix);
            }
            out[lin] = d;
        }
    }
}

template <bool LAST, bool DESC, bool FMA>
int launch3(const ng_plan* p, const double* in, double* out, const FuseDev& F, cudaStream_t st) {
    const i64 ncols = p->outer * p->inner;
    dim3 grid((unsigned)((ncols + BN - 1) / BN), (unsigned)((p->n + BM - 1) / BM));
    if (F.any)
        dense_apply_kernel<LAST, DESC, FMA, true><<<grid, NT, 0, st>>>(
            in, out, p->d_Mt, p->outer, (int)p->n, p->inner, p->axis, F);
    else
        dense_apply_kernel<LAST, DESC, FMA, false><<<grid, NT, 0, st>>>(
            in, out, p->d_Mt, p->outer, (int)p->n, p->inner, p->axis, F);
    NG_LAUNCH_CHECK();
    return NG_OK;
}

template <bool LAST>
int launch1(const ng_plan* p, const double* in, double* out, const FuseDev& F, cudaStream_t st) {
    if (p->use_fma) return launch3<LAST, false, true>(p, in, out, F, st);
    if (p->descending) return launch3<LAST, true, false>(p, in, out, F, st);
    return launch3<LAST, false, false>(p, in, out, F, st);
}

}  // namespace

int ng_dense_launch(const ng_plan* p, const double* in, double* out, const FuseDev& F,
                    cudaStream_t st) {
    if (p->size == 0) return NG_OK;
    return p->inner == 1 ? launch1<true>(p, in, out, F, st) : launch1<false>(p, in, out, F, st);
}
