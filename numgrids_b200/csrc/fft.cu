// fft.cu -- periodic spectral derivative (FFTDiff): real(ifft(W * fft(f, axis), axis)).
//
// Replaces the closure of reference numgrids/diff.py:133-143 (numpy.fft / pocketfft c2c).
// Tolerance for this path is rel. L-inf <= 1e-11 (BASELINE.json), so FMA is allowed here.
//
// Power-of-two n: one CTA keeps P pencil PAIRS in shared memory.  Two real pencils a, b are packed
// as z = a + i b; because W is Hermitian ((ik)^order with the Nyquist bin zeroed for odd order,
// diff.py:117-125) the operator F^-1 diag(W) F is real, so Re/Im of the result are D a and D b:
// one complex transform pair serves two pencils.  Forward = decimation-in-frequency (natural in,
// bit-reversed out), the multiplier table is pre-permuted to bit-reversed order and pre-scaled by
// 1/n on upload (exact: n is a power of two), inverse = decimation-in-time (bit-reversed in,
// natural out): no reordering pass.  Twiddles come from a host-computed fp64 table.
// HBM traffic 16 B/point (read f, write the real part).
//
// Other n: the dense real matrix D = Re(F^-1 diag(W) F) (what diff.py:145-157 builds) is applied by
// the ordered contraction kernel in dense.cu with FMA enabled.
#include "common.cuh"

namespace {

constexpr int FFT_NT = 256;

// W_br : [n][2] multipliers in bit-reversed order, pre-divided by n.  tw : [n/2][2] = exp(-2 pi i j/n).
template <bool LAST, bool FUSED>
__global__ void __launch_bounds__(FFT_NT)
fft_diff_kernel(const double* __restrict__ in, double* __restrict__ out,
                const double2* __restrict__ W_br, const double2* __restrict__ tw, i64 outer, int n,
                int log2n, i64 inner, int P, int nv /* samples per pencil in memory (<= n; the rest is zero padding) */,
                FuseDev F) {
    extern __shared__ double smem[];
    double* sre = smem;                // [P][n]
    double* sim = smem + (i64)P * n;   // [P][n]
    // A pencil with a non-finite sample (the reference's inf/nan rows on coordinate singularities,
    // numgrids/curvilinear.py:236-244) must not leak into the pencil it shares a complex transform
    // with: it is transformed as zeros and its whole output is NaN, as numpy.fft would make it.
    int* sbad = reinterpret_cast<int*>(smem + (i64)2 * P * n);   // [2P]
    for (int e = threadIdx.x; e < 2 * P; e += FFT_NT) sbad[e] = 0;
    __syncthreads();
    const i64 npencil = outer * inner;
    const i64 q0 = (i64)blockIdx.x * 2 * P;   // first pencil of this CTA
    const int tid = threadIdx.x;
    const int total = 2 * P * n;

    // ---- load: pencil q0+2p -> re, q0+2p+1 -> im
    for (int e = tid; e < total; e += FFT_NT) {
        int pl, k;
        if (LAST) { pl = e / n; k = e - pl * n; }
        else      { k = e / (2 * P); pl = e - k * (2 * P); }
        const i64 q = q0 + pl;
        double v = 0.0;
        if (q < npencil && k < nv) {
            const i64 o = LAST ? q : q / inner;
            const i64 c = LAST ? 0 : q - o * inner;
            const i64 lin = (o * nv + k) * inner + c;
            v = in[lin];
            if (FUSED && F.pre.ptr) {
                const Idx4 ix = unravel4(lin, F.shape);
                v = __dmul_rn(F.pre.ptr[coef_off(F.pre, ix)], v);
            }
        }
        if (ng_nonfinite(v)) { sbad[pl] = 1; v = 0.0; }
        ((pl & 1) ? sim : sre)[(pl >> 1) * n + k] = v;
    }
    __syncthreads();

    const int half = n >> 1;
    const int nbf = P * half;
    // ---- forward DIF: spans n/2 .. 1
    for (int s = log2n - 1; s >= 0; --s) {
        const int m = 1 << s;
        for (int e = tid; e < nbf; e += FFT_NT) {
            const int p = e / half, b = e - p * half;
            const int j = b & (m - 1);
            const int i0 = ((b >> s) << (s + 1)) + j + p * n, i1 = i0 + m;
            const double2 w = tw[j << (log2n - 1 - s)];
            const double ur = sre[i0], ui = sim[i0], vr = sre[i1], vi = sim[i1];
            sre[i0] = ur + vr;
            sim[i0] = ui + vi;
            const double dr = ur - vr, di = ui - vi;
            sre[i1] = fma(dr, w.x, -(di * w.y));
            sim[i1] = fma(dr, w.y, di * w.x);
        }
        __syncthreads();
    }
    // ---- multiply by (ik)^order / n, bit-reversed order
    for (int e = tid; e < P * n; e += FFT_NT) {
        const double2 w = W_br[e & (n - 1)];
        const double xr = sre[e], xi = sim[e];
        sre[e] = fma(xr, w.x, -(xi * w.y));
        sim[e] = fma(xr, w.y, xi * w.x);
    }
    __syncthreads();
    // ---- inverse DIT: spans 1 .. n/2, conjugate twiddles
    for (int s = 0; s < log2n; ++s) {
        const int m = 1 << s;
        for (int e = tid; e < nbf; e += FFT_NT) {
            const int p = e / half, b = e - p * half;
            const int j = b & (m - 1);
            const int i0 = ((b >> s) << (s + 1)) + j + p * n, i1 = i0 + m;
            const double2 w = tw[j << (log2n - 1 - s)];
            const double ur = sre[i0], ui = sim[i0], xr = sre[i1], xi = sim[i1];
            const double vr = fma(xr, w.x, xi * w.y);      // x * conj(w)
            const double vi = fma(xi, w.x, -(xr * w.y));
            sre[i0] = ur + vr;
            sim[i0] = ui + vi;
            sre[i1] = ur - vr;
            sim[i1] = ui - vi;
        }
        __syncthreads();
    }
    // ---- store real part -> pencil a, imaginary part -> pencil b
    for (int e = tid; e < total; e += FFT_NT) {
        int pl, k;
        if (LAST) { pl = e / n; k = e - pl * n; }
        else      { k = e / (2 * P); pl = e - k * (2 * P); }
        const i64 q = q0 + pl;
        if (q >= npencil || k >= nv) continue;
        const i64 o = LAST ? q : q / inner;
        const i64 c = LAST ? 0 : q - o * inner;
        const i64 lin = (o * nv + k) * inner + c;
        double d = ((pl & 1) ? sim : sre)[(pl >> 1) * n + k];
        if (sbad[pl]) d = ng_nan();
        if (FUSED) {
            const Idx4 ix = unravel4(lin, F.shape);
            d = fuse_tail(F, d, lin, ix);
        }
        out[lin] = d;
    }
}

}  // namespace

int ng_fft_launch(const ng_plan* p, const double* in, double* out, const FuseDev& F,
                  cudaStream_t st) {
    if (p->size == 0) return NG_OK;
    if (!p->fft_pow2) {
        NG_REQUIRE(!F.mid.ptr, "double apply requested from a kernel that does not implement it (ng_fft_twice_supported)");
        return ng_dense_launch(p, in, out, F, st);
    }
    {   // contiguous axis, 64 <= n <= 1024: two-pass register kernel (fft2.cu)
        bool handled = false;
        const int rc = ng_fft2_try_launch(p, in, out, F, st, &handled);
        if (rc != NG_OK || handled) return rc;
    }
    NG_REQUIRE(!F.mid.ptr, "double apply requested from a kernel that does not implement it (ng_fft_twice_supported)");
    const int n = p->fft_len;                            // transform length (== p->n unless the plan is zero-padded)
    const i64 npencil = p->outer * p->inner;
    // pencil pairs per CTA: enough butterflies for 256 threads, bounded by shared memory
    int P = 1;
    while (P * n < 1024 && (i64)(2 * P) * 2 <= npencil) P *= 2;
    const size_t smem = (size_t)P * n * 2 * sizeof(double) + (size_t)2 * P * sizeof(int);
    const i64 nblk = (npencil + 2 * P - 1) / (2 * P);
    const bool last = p->inner == 1;
#define NG_FFT_GO(LA, FU)                                                                        \
    do {                                                                                         \
        NG_FUNC_SMEM(200 * 1024, fft_diff_kernel<LA, FU>);                                       \
        fft_diff_kernel<LA, FU><<<(unsigned)nblk, FFT_NT, smem, st>>>(                           \
            in, out, (const double2*)p->d_W, (const double2*)p->d_tw, p->outer, n, p->log2n,     \
            p->inner, P, (int)p->n, F);                                                          \
    } while (0)
    if (last) { if (F.any) NG_FFT_GO(true, true); else NG_FFT_GO(true, false); }
    else      { if (F.any) NG_FFT_GO(false, true); else NG_FFT_GO(false, false); }
#undef NG_FFT_GO
    NG_TRACE(p->fft_len != p->n ? "fft_radix2<padded>" : "fft_radix2");
    NG_LAUNCH_CHECK();
    return NG_OK;
}
