// fft.cu -- periodic spectral derivative (FFTDiff): real(ifft(W * fft(f, axis), axis)).
//
// Replaces the closure of reference numgrids/diff.py:133-143 (numpy.fft / pocketfft c2c).
// Tolerance for this path is rel. L-inf <= 1e-11 (BASELINE.json), so FMA is allowed here.
//
// Power-of-two n outside the register kernels' range (n < 64, n > 1024; fft2.cu / fft_tile.cu serve 64 ... 1024) and long
// zero-padded plans: one CTA keeps P pencil PAIRS in shared memory.  Two real pencils a, b are packed as z = a + i b;
// because W is Hermitian ((ik)^order with the Nyquist bin zeroed for odd order, diff.py:117-125) the operator
// F^-1 diag(W) F is real, so Re/Im of the result are D a and D b: one complex transform pair serves two pencils.
// Forward = decimation-in-frequency (natural in, bit-reversed out) in radix-16 register passes, the multiplier table is
// pre-permuted to bit-reversed order and pre-scaled by 1/n on upload (exact: n is a power of two), inverse =
// decimation-in-time (bit-reversed in, natural out): no reordering pass.  The last forward pass, the multiplication and
// the first inverse pass are ONE shared-memory round trip; the factors of a pass are powers of one value of the
// host-computed fp64 table.  HBM traffic 16 B/point (read f, write the real part).
//
// Other n: the native mixed-radix kernel (fft_mixed.cu), or -- plans created with the dense real matrix
// D = Re(F^-1 diag(W) F) (what diff.py:145-157 builds) -- the ordered contraction kernel in dense.cu with FMA enabled.
#include "fft_net.cuh"

namespace {

constexpr int FFT_NT = 256;

// shared-memory index with one double of padding per 16: the radix-16 passes below gather at strides of 1, 16, 256 ...
// elements, and a stride of 16 doubles would put a warp's 32 accesses on one bank
__device__ __forceinline__ int fpad(int i) { return i + (i >> 4); }

// One radix-R pass (R = 2^LR) of the in-place transform over the stages s_lo .. s_lo+LR-1 (spans m_lo .. m_lo R/2): the R
// elements base + q m_lo of a sub-network go through the register network of fft_net.cuh -- the same butterflies as LR
// radix-2 stages in place, with one shared-memory round trip and one barrier instead of LR -- and the part of the
// radix-2 twiddles that depends on the position j_lo inside the span collapses into ONE factor per element,
// exp(-+2 pi i j_lo k / (R m_lo)), k = the element's (bit-reversed) output index.  FWD: DIF (network, then factors);
// inverse: the conjugate transpose (conjugate factors, then the DIT network).
template <int LR, bool FWD>
__device__ __forceinline__ void fft_pass(double* sre, double* sim, const double2* __restrict__ tw, int n, int log2n, int P,
                                         int s_lo) {
    constexpr int R = 1 << LR;
    const int m_lo = 1 << s_lo, per = n >> LR, items = P * per;
    const int sh = log2n - LR - s_lo;                 // table index of angle a / (R m_lo) on the n-point circle: a << sh
    for (int e = threadIdx.x; e < items; e += FFT_NT) {
        const int p = e / per, b = e - p * per;
        const int j_lo = b & (m_lo - 1);
        const int base = ((b >> s_lo) << (s_lo + LR)) + j_lo + p * n;
        double2 x[R];
#pragma unroll
        for (int q = 0; q < R; ++q) { const int i = fpad(base + q * m_lo); x[q] = make_double2(sre[i], sim[i]); }
        if (FWD) dif_forward<R>(x);
        if (j_lo != 0) {                              // (j_lo = 0: the whole sub-network has trivial factors)
            const double2 w = tw[j_lo << sh];         // j_lo < m_lo: the index is below n / R
            double2 pw[R];
            twiddle_powers<R>(make_double2(w.x, FWD ? w.y : -w.y), pw);
#pragma unroll
            for (int q = 1; q < R; ++q) { const double2 t = pw[bitrev(q, LR)]; x[q] = cmul(x[q], t.x, t.y); }
        }
        if (!FWD) dit_inverse<R>(x);
#pragma unroll
        for (int q = 0; q < R; ++q) { const int i = fpad(base + q * m_lo); sre[i] = x[q].x; sim[i] = x[q].y; }
    }
}
template <bool FWD>
__device__ __forceinline__ void fft_pass_any(int lr, double* sre, double* sim, const double2* tw, int n, int log2n, int P, int s_lo) {
    switch (lr) {
        case 4: fft_pass<4, FWD>(sre, sim, tw, n, log2n, P, s_lo); break;
        case 3: fft_pass<3, FWD>(sre, sim, tw, n, log2n, P, s_lo); break;
        case 2: fft_pass<2, FWD>(sre, sim, tw, n, log2n, P, s_lo); break;
        default: fft_pass<1, FWD>(sre, sim, tw, n, log2n, P, s_lo); break;
    }
}

// The last forward pass (stages 0 .. LR-1: trivial factors), the multiplication by the bit-reversed multiplier table and the
// first inverse pass work on the same R contiguous elements: one shared-memory round trip and one barrier instead of three.
template <int LR>
__device__ __forceinline__ void fft_mid(double* sre, double* sim, const double2* __restrict__ W_br, int n, int P) {
    constexpr int R = 1 << LR;
    const int items = (P * n) >> LR;
    for (int e = threadIdx.x; e < items; e += FFT_NT) {
        const int base = e << LR;
        double2 x[R];
#pragma unroll
        for (int q = 0; q < R; ++q) { const int i = fpad(base + q); x[q] = make_double2(sre[i], sim[i]); }
        dif_forward<R>(x);
#pragma unroll
        for (int q = 0; q < R; ++q) {
            const double2 w = W_br[(base + q) & (n - 1)];
            x[q] = cmul(x[q], w.x, w.y);
        }
        dit_inverse<R>(x);
#pragma unroll
        for (int q = 0; q < R; ++q) { const int i = fpad(base + q); sre[i] = x[q].x; sim[i] = x[q].y; }
    }
}

// W_br : [n][2] multipliers in bit-reversed order, pre-divided by n.  tw : [n/2][2] = exp(-2 pi i j/n).
template <bool LAST, bool FUSED>
__global__ void __launch_bounds__(FFT_NT, 2)
fft_diff_kernel(const double* __restrict__ in, double* __restrict__ out,
                const double2* __restrict__ W_br, const double2* __restrict__ tw, i64 outer, int n,
                int log2n, i64 inner, int P, int nv /* samples per pencil in memory (<= n; the rest is zero padding) */,
                FuseDev F) {
    extern __shared__ double smem[];
    const int plane = fpad(P * n) + 1;
    double* sre = smem;                // [P][n], padded index fpad()
    double* sim = smem + plane;
    // A pencil with a non-finite sample (the reference's inf/nan rows on coordinate singularities,
    // numgrids/curvilinear.py:236-244) must not leak into the pencil it shares a complex transform
    // with: it is transformed as zeros and its whole output is NaN, as numpy.fft would make it.
    int* sbad = reinterpret_cast<int*>(smem + (i64)2 * plane);   // [2P]
    for (int e = threadIdx.x; e < 2 * P; e += FFT_NT) sbad[e] = 0;
    __syncthreads();
    const i64 npencil = outer * inner;
    const i64 q0 = (i64)blockIdx.x * 2 * P;   // first pencil of this CTA
    const int tid = threadIdx.x;
    const int total = 2 * P * n;

    // ---- load: pencil q0+2p -> re, q0+2p+1 -> im.  n and 2P are powers of two (2P divides the block size): the element ->
    // (pencil, sample) map is shifts and masks, and off the contiguous axis a thread keeps ONE pencil for the whole loop,
    // so the 64-bit division that locates it runs once per thread instead of once per element
    const int lp = 31 - __clz(2 * P);                      // log2(2P)
    i64 pbase = 0;                                         // !LAST: offset of sample 0 of this thread's pencil
    bool pvalid = false;
    if (!LAST) {
        const i64 q = q0 + (tid & (2 * P - 1));
        pvalid = q < npencil;
        if (pvalid) { const i64 o = q / inner; pbase = o * nv * inner + (q - o * inner); }
    }
    // (LU loads are issued back to back before the first is used: one load in flight per thread leaves the kernel waiting
    // on HBM latency -- measured 62 against 76 Gpts/s at n = 2048)
    constexpr int LU = 8;
    for (int e0 = tid; e0 < total; e0 += FFT_NT * LU) {
        double v[LU];
        i64 lin[LU];
#pragma unroll
        for (int u = 0; u < LU; ++u) {
            const int e = e0 + u * FFT_NT;
            int pl, k;
            if (LAST) { pl = e >> log2n; k = e & (n - 1); }
            else      { k = e >> lp; pl = e & (2 * P - 1); }
            const bool ok = e < total && k < nv && (LAST ? q0 + pl < npencil : pvalid);
            lin[u] = ok ? (LAST ? (q0 + pl) * nv + k : pbase + (i64)k * inner) : -1;
            v[u] = ok ? in[lin[u]] : 0.0;
        }
        if (FUSED && F.pre.ptr) {
#pragma unroll
            for (int u = 0; u < LU; ++u)
                if (lin[u] >= 0) v[u] = __dmul_rn(F.pre.ptr[coef_off(F.pre, unravel4(lin[u], F.shape))], v[u]);
        }
#pragma unroll
        for (int u = 0; u < LU; ++u) {
            const int e = e0 + u * FFT_NT;
            if (e >= total) break;
            int pl, k;
            if (LAST) { pl = e >> log2n; k = e & (n - 1); }
            else      { k = e >> lp; pl = e & (2 * P - 1); }
            double x = v[u];
            if (ng_nonfinite(x)) { sbad[pl] = 1; x = 0.0; }
            ((pl & 1) ? sim : sre)[fpad((pl >> 1) * n + k)] = x;
        }
    }
    __syncthreads();

    // ---- forward DIF: radix-16 passes from the widest spans down to the last one (radix lr0 = 16 / 8 / 4 / 2), which
    // runs fused with the multiplication by (ik)^order / n (bit-reversed order) and the first inverse pass
    const int lr0 = (log2n & 3) ? (log2n & 3) : 4;
    for (int s_hi = log2n; s_hi > lr0; s_hi -= 4) {
        fft_pass_any<true>(4, sre, sim, tw, n, log2n, P, s_hi - 4);
        __syncthreads();
    }
    switch (lr0) {
        case 4: fft_mid<4>(sre, sim, W_br, n, P); break;
        case 3: fft_mid<3>(sre, sim, W_br, n, P); break;
        case 2: fft_mid<2>(sre, sim, W_br, n, P); break;
        default: fft_mid<1>(sre, sim, W_br, n, P); break;
    }
    __syncthreads();
    // ---- inverse DIT: the remaining passes in reverse order, conjugated
    for (int s_lo = lr0; s_lo < log2n; s_lo += 4) {
        fft_pass_any<false>(4, sre, sim, tw, n, log2n, P, s_lo);
        __syncthreads();
    }
    // ---- store real part -> pencil a, imaginary part -> pencil b
    for (int e = tid; e < total; e += FFT_NT) {
        int pl, k;
        if (LAST) { pl = e >> log2n; k = e & (n - 1); }
        else      { k = e >> lp; pl = e & (2 * P - 1); }
        if (k >= nv || !(LAST ? q0 + pl < npencil : pvalid)) continue;
        const i64 lin = LAST ? (q0 + pl) * nv + k : pbase + (i64)k * inner;
        double d = ((pl & 1) ? sim : sre)[fpad((pl >> 1) * n + k)];
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
    if (p->fft_mixed) return ng_fft_mixed_launch(p, in, out, F, st);
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
    // (a radix-16 pass has P n / 16 work items: at least one per thread, within ~70 KB so that three CTAs share an SM)
    int P = 1;
    const bool last = p->inner == 1;
    while (P * n < 4096 && (last || P < FFT_NT / 2) && (i64)(2 * P) * 2 <= npencil) P *= 2;    // (off the contiguous axis 2P divides the block size: see the kernel's load)
    const size_t smem = (size_t)((P * n + (P * n >> 4) + 1) * 2) * sizeof(double) + (size_t)2 * P * sizeof(int);   // fpad()
    const i64 nblk = (npencil + 2 * P - 1) / (2 * P);
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
