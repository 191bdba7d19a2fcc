// fft_mixed.cu -- periodic spectral derivative real(ifft(W * fft(f, axis), axis)) (reference numgrids/diff.py:133-143) on an
// axis whose length n is NOT a power of two, as a native n-point mixed-radix transform: n = 2^a 3^b 5^c 7^d 11^e 13^f.
//
// The reference's numpy.fft (pocketfft) is O(n log n) with native-length rounding for any such n; the zero-padded
// power-of-two embedding (ng_fft_plan_create_padded) is faster but carries a circular convolution's rounding (1e-13
// instead of 1e-15 of a first derivative).  This kernel is the n-point transform itself:
//   * one CTA keeps P pencil PAIRS in shared memory (two real pencils a, b as z = a + i b: W is Hermitian, so Re / Im of
//     the result are D a and D b -- fft.cu's packing);
//   * forward = in-place decimation in frequency, one pass per factor r of n: the r elements base + q m of a block of
//     length L = r m go through an r-point DFT held in registers (radix 2/4/8/16: the networks of fft_net.cuh; radix
//     3/5/7/9/11/13: the symmetric direct form below; radix 6/10/12/14/15: prime-factor butterflies of the two), output
//     p is multiplied by exp(-2 pi i j p / L) -- a power of ONE table value -- and stored back in place, so the spectrum
//     ends up in digit-reversed order; the host picks the factorisation with the fewest passes;
//   * the multiplier table is uploaded in that order, pre-divided by n;
//   * inverse = the conjugate transpose of the forward passes in reverse order: digit-reversed in, natural out -- no
//     reordering pass anywhere; the last forward pass, the multiplication and the first inverse pass share one
//     shared-memory round trip.
// Power-of-two radices run first (their strides are multiples of the odd part: conflict-free), the others after them.
// Measured: profiles/r02b_fft_mixed.txt (54-147 Gpts/s for n = 6 ... 3000).
// FMA is allowed here (tolerance 1e-11, BASELINE.json); HBM traffic 16 B per point.
#include "fft_net.cuh"

#include <type_traits>
#include <vector>

namespace {

constexpr int MX_NT = 256;
constexpr int MX_MAX_STAGES = 12;
struct MixStages { int nst; int radix[MX_MAX_STAGES]; };

__device__ __forceinline__ int mpad(int i, int psh) { return i + (i >> psh); }   // psh = 4: fft.cu's padding (strides of 16 doubles would share a bank); 31: none

// a / d for 0 <= a < 2^22 through a float reciprocal, corrected: the passes below divide by run-time block lengths once
// per work item, and a 32-bit integer division is ~20 instructions next to a radix-3 butterfly's 30
__device__ __forceinline__ int qdiv(int a, int d, float rd) {
    int q = __float2int_rz(((float)a + 0.5f) * rd);
    const int r = a - q * d;
    q += (r >= d) - (r < 0);
    return q;
}

template <int R> __device__ __forceinline__ double odd_cos(int k);
template <int R> __device__ __forceinline__ double odd_sin(int k);
template <> __device__ __forceinline__ double odd_cos<3>(int k) {   // cos(2 pi k / 3), k = 0..2
    switch (k) {
        case 0: return 1.0;
        case 1: return -0.5;
        case 2: return -0.5;
        default: return 0.0;
    }
}
template <> __device__ __forceinline__ double odd_sin<3>(int k) {   // sin(2 pi k / 3)
    switch (k) {
        case 0: return 0.0;
        case 1: return 0.8660254037844386;
        case 2: return -0.8660254037844386;
        default: return 0.0;
    }
}
template <> __device__ __forceinline__ double odd_cos<5>(int k) {   // cos(2 pi k / 5), k = 0..4
    switch (k) {
        case 0: return 1.0;
        case 1: return 0.30901699437494745;
        case 2: return -0.8090169943749475;
        case 3: return -0.8090169943749475;
        case 4: return 0.30901699437494745;
        default: return 0.0;
    }
}
template <> __device__ __forceinline__ double odd_sin<5>(int k) {   // sin(2 pi k / 5)
    switch (k) {
        case 0: return 0.0;
        case 1: return 0.9510565162951535;
        case 2: return 0.5877852522924731;
        case 3: return -0.5877852522924731;
        case 4: return -0.9510565162951535;
        default: return 0.0;
    }
}
template <> __device__ __forceinline__ double odd_cos<7>(int k) {   // cos(2 pi k / 7), k = 0..6
    switch (k) {
        case 0: return 1.0;
        case 1: return 0.6234898018587335;
        case 2: return -0.2225209339563144;
        case 3: return -0.9009688679024191;
        case 4: return -0.9009688679024191;
        case 5: return -0.2225209339563144;
        case 6: return 0.6234898018587335;
        default: return 0.0;
    }
}
template <> __device__ __forceinline__ double odd_sin<7>(int k) {   // sin(2 pi k / 7)
    switch (k) {
        case 0: return 0.0;
        case 1: return 0.7818314824680298;
        case 2: return 0.9749279121818236;
        case 3: return 0.4338837391175581;
        case 4: return -0.4338837391175581;
        case 5: return -0.9749279121818236;
        case 6: return -0.7818314824680298;
        default: return 0.0;
    }
}
template <> __device__ __forceinline__ double odd_cos<9>(int k) {   // cos(2 pi k / 9), k = 0..8
    switch (k) {
        case 0: return 1.0;
        case 1: return 0.766044443118978;
        case 2: return 0.17364817766693036;
        case 3: return -0.5;
        case 4: return -0.9396926207859084;
        case 5: return -0.9396926207859084;
        case 6: return -0.5;
        case 7: return 0.17364817766693036;
        case 8: return 0.766044443118978;
        default: return 0.0;
    }
}
template <> __device__ __forceinline__ double odd_sin<9>(int k) {   // sin(2 pi k / 9)
    switch (k) {
        case 0: return 0.0;
        case 1: return 0.6427876096865394;
        case 2: return 0.984807753012208;
        case 3: return 0.8660254037844386;
        case 4: return 0.3420201433256687;
        case 5: return -0.3420201433256687;
        case 6: return -0.8660254037844386;
        case 7: return -0.984807753012208;
        case 8: return -0.6427876096865394;
        default: return 0.0;
    }
}
template <> __device__ __forceinline__ double odd_cos<11>(int k) {   // cos(2 pi k / 11), k = 0..10
    switch (k) {
        case 0: return 1.0;
        case 1: return 0.8412535328311812;
        case 2: return 0.41541501300188644;
        case 3: return -0.14231483827328514;
        case 4: return -0.6548607339452851;
        case 5: return -0.9594929736144974;
        case 6: return -0.9594929736144974;
        case 7: return -0.6548607339452851;
        case 8: return -0.14231483827328514;
        case 9: return 0.41541501300188644;
        case 10: return 0.8412535328311812;
        default: return 0.0;
    }
}
template <> __device__ __forceinline__ double odd_sin<11>(int k) {   // sin(2 pi k / 11)
    switch (k) {
        case 0: return 0.0;
        case 1: return 0.5406408174555976;
        case 2: return 0.9096319953545183;
        case 3: return 0.9898214418809327;
        case 4: return 0.7557495743542583;
        case 5: return 0.28173255684142967;
        case 6: return -0.28173255684142967;
        case 7: return -0.7557495743542583;
        case 8: return -0.9898214418809327;
        case 9: return -0.9096319953545183;
        case 10: return -0.5406408174555976;
        default: return 0.0;
    }
}
template <> __device__ __forceinline__ double odd_cos<13>(int k) {   // cos(2 pi k / 13), k = 0..12
    switch (k) {
        case 0: return 1.0;
        case 1: return 0.8854560256532099;
        case 2: return 0.5680647467311558;
        case 3: return 0.12053668025532305;
        case 4: return -0.3546048870425356;
        case 5: return -0.7485107481711011;
        case 6: return -0.970941817426052;
        case 7: return -0.970941817426052;
        case 8: return -0.7485107481711011;
        case 9: return -0.3546048870425356;
        case 10: return 0.12053668025532305;
        case 11: return 0.5680647467311558;
        case 12: return 0.8854560256532099;
        default: return 0.0;
    }
}
template <> __device__ __forceinline__ double odd_sin<13>(int k) {   // sin(2 pi k / 13)
    switch (k) {
        case 0: return 0.0;
        case 1: return 0.46472317204376856;
        case 2: return 0.8229838658936564;
        case 3: return 0.992708874098054;
        case 4: return 0.9350162426854148;
        case 5: return 0.6631226582407952;
        case 6: return 0.23931566428755777;
        case 7: return -0.23931566428755777;
        case 8: return -0.6631226582407952;
        case 9: return -0.9350162426854148;
        case 10: return -0.992708874098054;
        case 11: return -0.8229838658936564;
        case 12: return -0.46472317204376856;
        default: return 0.0;
    }
}

// R-point DFT, R odd, natural order in and out, in the symmetric form: with a_q = x_q + x_{R-q}, b_q = x_q - x_{R-q}
//   y_p, y_{R-p} = (x_0 + sum_q a_q cos(2 pi p q / R))  -/+  i (sum_q b_q sin(2 pi p q / R))     (forward; inverse: +/-)
template <int R, bool FWD>
__device__ __forceinline__ void odd_dft(double2* x) {
    constexpr int H = R / 2;
    double2 a[H], b[H];
#pragma unroll
    for (int q = 1; q <= H; ++q) {
        a[q - 1] = make_double2(x[q].x + x[R - q].x, x[q].y + x[R - q].y);
        b[q - 1] = make_double2(x[q].x - x[R - q].x, x[q].y - x[R - q].y);
    }
    double2 y0 = x[0];
#pragma unroll
    for (int q = 0; q < H; ++q) { y0.x += a[q].x; y0.y += a[q].y; }
#pragma unroll
    for (int p = 1; p <= H; ++p) {
        double2 A = x[0], B = make_double2(0.0, 0.0);
#pragma unroll
        for (int q = 1; q <= H; ++q) {
            const int k = (p * q) % R;
            const double c = odd_cos<R>(k), s = odd_sin<R>(k);
            A.x = fma(a[q - 1].x, c, A.x); A.y = fma(a[q - 1].y, c, A.y);
            B.x = fma(b[q - 1].x, s, B.x); B.y = fma(b[q - 1].y, s, B.y);
        }
        // A - i B = (A.x + B.y, A.y - B.x),  A + i B = (A.x - B.y, A.y + B.x)
        const double2 lo = make_double2(A.x + B.y, A.y - B.x), hi = make_double2(A.x - B.y, A.y + B.x);
        x[p] = FWD ? lo : hi;
        x[R - p] = FWD ? hi : lo;
    }
    x[0] = y0;
}

template <int R> struct IsPow2 { static constexpr bool v = (R & (R - 1)) == 0; };

// Composite radices 6 10 12 14 15 = A * B with coprime factors: a prime-factor (Good-Thomas) butterfly in registers -- input
// index (B n1 + A n2) mod R, output index the CRT combination of (k1, k2) -- needs NO factors between its two stages, so
// a length like 1000 = 10 * 10 * 10 takes three shared-memory passes instead of the four of 8 * 5 * 5 * 5.
__host__ __device__ constexpr int mx_split_a(int r) { return r == 6 || r == 10 || r == 14 ? 2 : r == 12 ? 4 : r == 15 ? 3 : 0; }
__host__ __device__ constexpr int mx_modinv(int a, int m) {
    for (int x = 1; x <= m; ++x)
        if ((a % m) * x % m == 1 % m) return x;
    return 1;
}
__host__ __device__ constexpr int mx_slot(int ra, int rb, int n1, int n2) { return (rb * n1 + ra * n2) % (ra * rb); }
// frequency digit held by register q after the forward r-point butterfly (power-of-two networks leave bit-reversed order)
__host__ __device__ constexpr int mx_base_freq(int r, int q) {
    if ((r & (r - 1)) != 0) return q;
    int l = 0;
    while ((1 << l) < r) ++l;
    return bitrev(q, l);
}
__host__ __device__ constexpr int mx_freq(int r, int q) {
    const int ra = mx_split_a(r);
    if (!ra) return mx_base_freq(r, q);
    const int rb = r / ra;
    for (int n1 = 0; n1 < ra; ++n1)
        for (int n2 = 0; n2 < rb; ++n2)
            if (mx_slot(ra, rb, n1, n2) == q)
                return (mx_base_freq(ra, n1) * rb * mx_modinv(rb, ra) + mx_base_freq(rb, n2) * ra * mx_modinv(ra, rb)) % r;
    return 0;
}

template <int B, int E, class F>
__device__ __forceinline__ void static_for(F&& f) {
    if constexpr (B < E) {
        f(std::integral_constant<int, B>{});
        static_for<B + 1, E>(f);
    }
}

template <int R, bool FWD>
__device__ __forceinline__ void small_dft(double2* x) {
    constexpr int RA = mx_split_a(R);
    if constexpr (RA != 0) {
        constexpr int RB = R / RA;
        auto stage_a = [&]() {
            static_for<0, RB>([&](auto n2c) {
                constexpr int n2 = decltype(n2c)::value;
                double2 t[RA];
                static_for<0, RA>([&](auto n1c) { constexpr int n1 = decltype(n1c)::value; t[n1] = x[mx_slot(RA, RB, n1, n2)]; });
                small_dft<RA, FWD>(t);
                static_for<0, RA>([&](auto n1c) { constexpr int n1 = decltype(n1c)::value; x[mx_slot(RA, RB, n1, n2)] = t[n1]; });
            });
        };
        auto stage_b = [&]() {
            static_for<0, RA>([&](auto n1c) {
                constexpr int n1 = decltype(n1c)::value;
                double2 t[RB];
                static_for<0, RB>([&](auto n2c) { constexpr int n2 = decltype(n2c)::value; t[n2] = x[mx_slot(RA, RB, n1, n2)]; });
                small_dft<RB, FWD>(t);
                static_for<0, RB>([&](auto n2c) { constexpr int n2 = decltype(n2c)::value; x[mx_slot(RA, RB, n1, n2)] = t[n2]; });
            });
        };
        if (FWD) { stage_a(); stage_b(); } else { stage_b(); stage_a(); }
    } else if constexpr (IsPow2<R>::v) {
        if (FWD) dif_forward<R>(x); else dit_inverse<R>(x);
    } else {
        odd_dft<R, FWD>(x);
    }
}

// One pass over the blocks of length L (L = R m): see the header.  tw = exp(-2 pi i t / n), t < n.
template <int R, bool FWD>
__device__ __noinline__ void mix_pass(double* sre, double* sim, const double2* __restrict__ tw, int n, int P, int L, int psh) {
    const int m = L / R, per = n / R, items = P * per, step = n / L;
    const float rper = 1.0f / (float)per, rm = 1.0f / (float)m;
    for (int e = threadIdx.x; e < items; e += MX_NT) {
        const int p = qdiv(e, per, rper), b = e - p * per;
        const int blk = qdiv(b, m, rm), j = b - blk * m;
        const int base = p * n + blk * L + j;
        double2 x[R];
#pragma unroll
        for (int q = 0; q < R; ++q) { const int i = mpad(base + q * m, psh); x[q] = make_double2(sre[i], sim[i]); }
        if (FWD) small_dft<R, true>(x);
        if (j != 0) {                               // (j = 0: trivial factors)
            const double2 w = tw[j * step];         // exp(-2 pi i j / L); the other outputs' factors are its powers
            double2 pw[R];
            twiddle_powers<R>(make_double2(w.x, FWD ? w.y : -w.y), pw);
            static_for<1, R>([&](auto qc) {
                constexpr int q = decltype(qc)::value;
                constexpr int f = mx_freq(R, q);
                x[q] = cmul(x[q], pw[f].x, pw[f].y);
            });
        }
        if (!FWD) small_dft<R, false>(x);
#pragma unroll
        for (int q = 0; q < R; ++q) { const int i = mpad(base + q * m, psh); sre[i] = x[q].x; sim[i] = x[q].y; }
    }
}

// The last forward pass (blocks of R contiguous elements: trivial factors), the multiplication by the digit-reversed
// multiplier table and the first inverse pass touch the same R elements: one shared-memory round trip instead of three.
template <int R>
__device__ __noinline__ void mix_mid(double* sre, double* sim, const double2* __restrict__ W_dr, int n, int P, int psh) {
    const int per = n / R, items = P * per;
    const float rper = 1.0f / (float)per;
    for (int e = threadIdx.x; e < items; e += MX_NT) {
        const int p = qdiv(e, per, rper);
        const int base = e * R, k0 = base - p * n;
        double2 x[R];
#pragma unroll
        for (int q = 0; q < R; ++q) { const int i = mpad(base + q, psh); x[q] = make_double2(sre[i], sim[i]); }
        small_dft<R, true>(x);
#pragma unroll
        for (int q = 0; q < R; ++q) {
            const double2 w = W_dr[k0 + q];
            x[q] = cmul(x[q], w.x, w.y);
        }
        small_dft<R, false>(x);
#pragma unroll
        for (int q = 0; q < R; ++q) { const int i = mpad(base + q, psh); sre[i] = x[q].x; sim[i] = x[q].y; }
    }
}
__device__ __forceinline__ void mix_mid_any(int r, double* sre, double* sim, const double2* W_dr, int n, int P, int psh) {
    switch (r) {
        case 2: mix_mid<2>(sre, sim, W_dr, n, P, psh); break;
        case 4: mix_mid<4>(sre, sim, W_dr, n, P, psh); break;
        case 8: mix_mid<8>(sre, sim, W_dr, n, P, psh); break;
        case 16: mix_mid<16>(sre, sim, W_dr, n, P, psh); break;
        case 3: mix_mid<3>(sre, sim, W_dr, n, P, psh); break;
        case 5: mix_mid<5>(sre, sim, W_dr, n, P, psh); break;
        case 6: mix_mid<6>(sre, sim, W_dr, n, P, psh); break;
        case 10: mix_mid<10>(sre, sim, W_dr, n, P, psh); break;
        case 12: mix_mid<12>(sre, sim, W_dr, n, P, psh); break;
        case 14: mix_mid<14>(sre, sim, W_dr, n, P, psh); break;
        case 15: mix_mid<15>(sre, sim, W_dr, n, P, psh); break;
        case 7: mix_mid<7>(sre, sim, W_dr, n, P, psh); break;
        case 9: mix_mid<9>(sre, sim, W_dr, n, P, psh); break;
        case 11: mix_mid<11>(sre, sim, W_dr, n, P, psh); break;
        default: mix_mid<13>(sre, sim, W_dr, n, P, psh); break;
    }
}

template <bool FWD>
__device__ __forceinline__ void mix_pass_any(int r, double* sre, double* sim, const double2* tw, int n, int P, int L, int psh) {
    switch (r) {
        case 2: mix_pass<2, FWD>(sre, sim, tw, n, P, L, psh); break;
        case 4: mix_pass<4, FWD>(sre, sim, tw, n, P, L, psh); break;
        case 8: mix_pass<8, FWD>(sre, sim, tw, n, P, L, psh); break;
        case 16: mix_pass<16, FWD>(sre, sim, tw, n, P, L, psh); break;
        case 3: mix_pass<3, FWD>(sre, sim, tw, n, P, L, psh); break;
        case 5: mix_pass<5, FWD>(sre, sim, tw, n, P, L, psh); break;
        case 6: mix_pass<6, FWD>(sre, sim, tw, n, P, L, psh); break;
        case 10: mix_pass<10, FWD>(sre, sim, tw, n, P, L, psh); break;
        case 12: mix_pass<12, FWD>(sre, sim, tw, n, P, L, psh); break;
        case 14: mix_pass<14, FWD>(sre, sim, tw, n, P, L, psh); break;
        case 15: mix_pass<15, FWD>(sre, sim, tw, n, P, L, psh); break;
        case 7: mix_pass<7, FWD>(sre, sim, tw, n, P, L, psh); break;
        case 9: mix_pass<9, FWD>(sre, sim, tw, n, P, L, psh); break;
        case 11: mix_pass<11, FWD>(sre, sim, tw, n, P, L, psh); break;
        default: mix_pass<13, FWD>(sre, sim, tw, n, P, L, psh); break;
    }
}

// W_dr : [n] multipliers / n in the digit-reversed order of the forward passes.  tw : [n] = exp(-2 pi i t / n).
template <bool LAST, bool FUSED>
__global__ void __launch_bounds__(MX_NT, 2)
fft_mixed_kernel(const double* __restrict__ in, double* __restrict__ out, const double2* __restrict__ W_dr,
                 const double2* __restrict__ tw, i64 outer, int n, i64 inner, int P, int psh, const __grid_constant__ MixStages S, FuseDev F) {
    extern __shared__ double smem[];
    const int plane = mpad(P * n, psh) + 1;
    double* sre = smem;                // [P][n], padded index mpad()
    double* sim = smem + plane;
    // a pencil with a non-finite sample is transformed as zeros and comes out all-NaN, as numpy.fft makes it, without
    // touching the pencil that shares its complex transform (fft.cu; numgrids/curvilinear.py:236-244 relies on it)
    int* sbad = reinterpret_cast<int*>(smem + (i64)2 * plane);   // [2P]
    for (int e = threadIdx.x; e < 2 * P; e += MX_NT) sbad[e] = 0;
    __syncthreads();
    const i64 npencil = outer * inner;
    const i64 q0 = (i64)blockIdx.x * 2 * P;   // first pencil of this CTA
    const int tid = threadIdx.x;
    const int total = 2 * P * n;
    const float rn = 1.0f / (float)n;
    const int lp = 31 - __clz(2 * P);                      // 2P is a power of two that divides the block size
    i64 pbase = 0;                                         // !LAST: offset of sample 0 of this thread's pencil
    bool pvalid = false;
    if (!LAST) {
        const i64 q = q0 + (tid & (2 * P - 1));
        pvalid = q < npencil;
        if (pvalid) { const i64 o = q / inner; pbase = o * n * inner + (q - o * inner); }
    }
    // ---- load: pencil q0+2p -> re, q0+2p+1 -> im; on the contiguous axis the CTA's pencils are ONE contiguous range
    constexpr int LU = 8;
    for (int e0 = tid; e0 < total; e0 += MX_NT * LU) {
        double v[LU];
        i64 lin[LU];
#pragma unroll
        for (int u = 0; u < LU; ++u) {
            const int e = e0 + u * MX_NT;
            bool ok = e < total;
            if (LAST) {
                lin[u] = q0 * n + e;
                ok = ok && lin[u] < npencil * n;
            } else {
                lin[u] = pbase + (i64)(e >> lp) * inner;
                ok = ok && pvalid;
            }
            if (!ok) lin[u] = -1;
            v[u] = ok ? in[lin[u]] : 0.0;
        }
        if (FUSED && F.pre.ptr) {
#pragma unroll
            for (int u = 0; u < LU; ++u)
                if (lin[u] >= 0) v[u] = __dmul_rn(F.pre.ptr[coef_off(F.pre, unravel4(lin[u], F.shape))], v[u]);
        }
#pragma unroll
        for (int u = 0; u < LU; ++u) {
            const int e = e0 + u * MX_NT;
            if (e >= total) break;
            int pl, k;
            if (LAST) { pl = qdiv(e, n, rn); k = e - pl * n; }
            else      { k = e >> lp; pl = e & (2 * P - 1); }
            double x = v[u];
            if (ng_nonfinite(x)) { sbad[pl] = 1; x = 0.0; }
            ((pl & 1) ? sim : sre)[mpad((pl >> 1) * n + k, psh)] = x;
        }
    }
    __syncthreads();

    // ---- forward passes; the last one runs fused with the multiplication by (ik)^order / n and the first inverse pass
    {
        int L = n;
        for (int s = 0; s + 1 < S.nst; ++s) {
            mix_pass_any<true>(S.radix[s], sre, sim, tw, n, P, L, psh);
            L /= S.radix[s];
            __syncthreads();
        }
        mix_mid_any(S.radix[S.nst - 1], sre, sim, W_dr, n, P, psh);
        __syncthreads();
    }
    // ---- inverse: the adjoint passes in reverse order
    {
        int L = S.radix[S.nst - 1];
        for (int s = S.nst - 2; s >= 0; --s) {
            L *= S.radix[s];
            mix_pass_any<false>(S.radix[s], sre, sim, tw, n, P, L, psh);
            __syncthreads();
        }
    }
    // ---- store real part -> pencil a, imaginary part -> pencil b
    for (int e = tid; e < total; e += MX_NT) {
        int pl, k;
        i64 lin;
        if (LAST) {
            pl = qdiv(e, n, rn); k = e - pl * n;
            lin = q0 * n + e;
            if (lin >= npencil * n) break;
        } else {
            k = e >> lp; pl = e & (2 * P - 1);
            if (!pvalid) break;
            lin = pbase + (i64)k * inner;
        }
        double d = ((pl & 1) ? sim : sre)[mpad((pl >> 1) * n + k, psh)];
        if (sbad[pl]) d = ng_nan();
        if (FUSED) {
            const Idx4 ix = unravel4(lin, F.shape);
            d = fuse_tail(F, d, lin, ix);
        }
        out[lin] = d;
    }
}

// n as a product of radices from {16 15 14 13 12 11 10 9 8 7 6 5 4 3 2} with the fewest passes (ties: the smallest largest
// radix, then the largest smallest one), powers of two first, the others ascending; false if n has a prime factor > 13
struct MixBest { int cnt, mx, mn; int r[MX_MAX_STAGES]; };
bool mix_search(i64 n, int depth, MixBest& cur, MixBest& best) {
    if (n == 1) {
        if (best.cnt < 0 || cur.cnt < best.cnt || (cur.cnt == best.cnt && (cur.mx < best.mx || (cur.mx == best.mx && cur.mn > best.mn)))) best = cur;
        return true;
    }
    if (depth >= MX_MAX_STAGES || (best.cnt >= 0 && depth + 1 > best.cnt)) return false;
    bool any = false;
    const int last = depth ? cur.r[depth - 1] : 16;
    for (int r = last; r >= 2; --r) {                       // non-increasing radices: every multiset once
        if (n % r) continue;
        MixBest nxt = cur;
        nxt.r[depth] = r;
        nxt.cnt = depth + 1;
        nxt.mx = depth ? cur.mx : r;
        nxt.mn = r;
        any |= mix_search(n / r, depth + 1, nxt, best);
    }
    return any;
}
bool mix_factor(i64 n, MixStages* S) {
    S->nst = 0;
    if (n < 2) return false;
    i64 m = n;
    const int primes[6] = {2, 3, 5, 7, 11, 13};
    for (int pi = 0; pi < 6; ++pi)
        while (m % primes[pi] == 0) m /= primes[pi];
    if (m != 1) return false;
    MixBest cur{}, best{};
    best.cnt = -1;
    if (!mix_search(n, 0, cur, best) || best.cnt < 1) return false;
    int rest[MX_MAX_STAGES], nr = 0;
    for (int i = 0; i < best.cnt; ++i) {                    // (found in non-increasing order)
        const int r = best.r[i];
        if ((r & (r - 1)) == 0) S->radix[S->nst++] = r;
        else rest[nr++] = r;
    }
    for (int i = nr - 1; i >= 0; --i) S->radix[S->nst++] = rest[i];
    return true;
}

// ---- plan-time choices: pencil pairs per CTA and the shared-memory padding -------------------------------------------

// Pencil pairs per CTA: a power of two (off the contiguous axis 2P divides the block size) within 64 KB of shared memory
// (three CTAs per SM), chosen for the fewest block-wide iterations per pencil: a pass of radix r has P n / r work items
// for 256 threads, and e.g. n = 360 = 8 * 5 * 9 leaves 40 % of the last iteration idle at P = 4.
int mix_choose_pairs(int n, const MixStages& S, i64 npencil, bool last) {
    int P = 1;
    double best = 1e300;
    for (int c = 1; c <= 1024; c *= 2) {
        if (c > 1 && ((i64)(2 * c) * 2 > npencil || (size_t)c * n * 18 > 64 * 1024 || (!last && c > MX_NT / 2))) break;
        double cost = 6.0;                                  // per-CTA fixed part (launch, table init, barriers)
        for (int s = 0; s < S.nst; ++s) {
            const int r = S.radix[s];
            const int it = (int)(((i64)c * n / r + MX_NT - 1) / MX_NT);
            cost += (double)it * r * (s + 1 == S.nst ? 1.5 : 2.0);   // (the fused middle pass runs once, the others twice)
        }
        cost += 2.0 * (((i64)2 * c * n + MX_NT - 1) / MX_NT);              // load + store
        cost /= c;
        if (cost < best * 0.97) { best = cost; P = c; }
    }
    return P;
}

// Shared-memory wavefronts of one 64-bit access of a warp: the two half-warps are served separately, each in as many
// passes as its most contended 8-byte bank has distinct words.
int mix_wavefronts(const int* word, int lanes) {
    int total = 0;
    for (int h = 0; h < lanes; h += 16) {
        int worst = 0;
        for (int b = 0; b < 16; ++b) {
            int distinct = 0, seen[16];
            for (int l = h; l < h + 16 && l < lanes; ++l) {
                if ((word[l] & 15) != b) continue;
                bool dup = false;
                for (int k = 0; k < distinct; ++k) dup |= seen[k] == word[l];
                if (!dup) seen[distinct++] = word[l];
            }
            if (distinct > worst) worst = distinct;
        }
        total += worst;
    }
    return total;
}

// Bank-conflict model of one CTA (the kernel's own index maps) for the padding i + (i >> psh).  Lengths that are not
// powers of two put the passes' bases anywhere, and fft.cu's i + (i >> 4) -- made for strides of 16 -- then costs every
// unaligned unit-stride access one conflict per half-warp (ncu at n = 1000: 231 M conflicts, 1.6 x the wavefronts of a
// conflict-free kernel); which shift is best depends on n, its factors and P, so the plan asks this model.
double mix_conflict_score(int n, const MixStages& S, int P, bool last, int psh) {
    auto pad = [&](int i) { return i + (i >> psh); };
    const int plane = pad(P * n) + 1;
    double score = 0.0;
    int word[32];
    // load / store phases: lane -> (pencil, sample)
    const int total = 2 * P * n;
    int lp = 0;
    while ((1 << lp) < 2 * P) ++lp;
    for (int e0 = 0; e0 < total; e0 += 32) {
        int lanes = 0;
        for (int e = e0; e < e0 + 32 && e < total; ++e) {
            const int pl = last ? e / n : (e & (2 * P - 1)), k = last ? e - pl * n : (e >> lp);
            word[lanes++] = ((pl & 1) ? plane : 0) + pad((pl >> 1) * n + k);
        }
        score += 2.0 * mix_wavefronts(word, lanes);
    }
    // passes: 2 arrays x (load + store) = 4 accesses of the same pattern per element; every pass but the last runs twice
    int L = n;
    for (int s = 0; s < S.nst; ++s) {
        const int r = S.radix[s], m = L / r, per = n / r, items = P * per;
        const double weight = s + 1 == S.nst ? 4.0 : 8.0;
        // (work items are spread over the block's 8 warps in rounds of 256; every warp sees 32 consecutive items)
        for (int e0 = 0; e0 < items; e0 += 32)
            for (int q = 0; q < r; ++q) {
                int lanes = 0;
                for (int e = e0; e < e0 + 32 && e < items; ++e) {
                    const int pp = e / per, b = e - pp * per, blk = b / m, j = b - blk * m;
                    word[lanes++] = pad(pp * n + blk * L + j + q * m);
                }
                score += weight * mix_wavefronts(word, lanes);
            }
        L = m;
    }
    return score;
}

}  // namespace

bool ng_fft_mixed_supported(i64 n) {
    MixStages S;
    return n >= 3 && n <= NG_FFT_MIXED_MAX && (n & (n - 1)) != 0 && mix_factor(n, &S);
}

extern "C" int ng_fft_mixed_radices(int64_t n, int* radices12, int* count) {
    NG_REQUIRE(radices12 != nullptr && count != nullptr, "output pointers are NULL");
    *count = 0;
    MixStages S;
    if (!ng_fft_mixed_supported(n) || !mix_factor(n, &S)) {
        ng_set_error("n = %lld has no mixed-radix plan (a power of two, a prime factor above 13, or n > %d)", (long long)n,
                     NG_FFT_MIXED_MAX);
        return NG_EINVAL;
    }
    for (int s = 0; s < S.nst; ++s) radices12[s] = S.radix[s];
    *count = S.nst;
    return NG_OK;
}

// Host tables of the native n-point plan (api.cu uploads them): multipliers / n in digit-reversed order, the full twiddle
// circle; the stage list goes into the plan.
int ng_fft_mixed_tables(ng_plan* p, const double* h_W_re_im, std::vector<double>& Wdr, std::vector<double>& tw) {
    const i64 n = p->n;
    MixStages S;
    if (!ng_fft_mixed_supported(n) || !mix_factor(n, &S)) {
        ng_set_error("n = %lld has no mixed-radix plan (prime factors above 13, or n > %d)", (long long)n, NG_FFT_MIXED_MAX);
        return NG_EINVAL;
    }
    p->fft_mixed = 1;
    p->fft_len = (int)n;
    p->mix_nst = S.nst;
    for (int s = 0; s < S.nst; ++s) p->mix_radix[s] = S.radix[s];
    {
        const bool last = p->inner == 1;
        p->mix_pairs = mix_choose_pairs((int)n, S, p->outer * p->inner, last);
        double best = 1e300;
        const int shifts[5] = {31, 3, 4, 5, 6};              // 31: no padding
        for (int k = 0; k < 5; ++k) {
            const double sc = mix_conflict_score((int)n, S, p->mix_pairs, last, shifts[k]);
            if (sc < best * 0.98) { best = sc; p->mix_pad_shift = shifts[k]; }
        }
    }
    Wdr.assign((size_t)(2 * n), 0.0);
    tw.assign((size_t)(2 * n), 0.0);
    for (i64 i = 0; i < n; ++i) {
        // storage position i = sum_s q_s m_s holds bin k = f_1 + r_1 (f_2 + r_2 (...)), f_s = freq_digit(q_s)
        i64 L = n, rem = i, mult = 1, k = 0;
        for (int s = 0; s < S.nst; ++s) {
            const i64 m = L / S.radix[s];
            const i64 q = rem / m;
            rem -= q * m;
            k += (i64)mx_freq(S.radix[s], (int)q) * mult;
            mult *= S.radix[s];
            L = m;
        }
        Wdr[(size_t)(2 * i)] = h_W_re_im[2 * k] / (double)n;
        Wdr[(size_t)(2 * i + 1)] = h_W_re_im[2 * k + 1] / (double)n;
    }
    const long double two_pi = 6.283185307179586476925286766559005768L;
    for (i64 t = 0; t < n; ++t) {
        const long double a = two_pi * (long double)t / (long double)n;
        tw[(size_t)(2 * t)] = (double)cosl(a);
        tw[(size_t)(2 * t + 1)] = (double)(-sinl(a));
    }
    return NG_OK;
}

int ng_fft_mixed_launch(const ng_plan* p, const double* in, double* out, const FuseDev& F, cudaStream_t st) {
    NG_REQUIRE(!F.mid.ptr, "double apply requested from a kernel that does not implement it (ng_fft_twice_supported)");
    const int n = (int)p->n;
    const i64 npencil = p->outer * p->inner;
    const bool last = p->inner == 1;
    const int P = p->mix_pairs;                              // pencil pairs per CTA and padding: chosen with the plan
    const int psh_env = (int)NG_ENV_INT("NG_FFT_MIXED_PAD", 0);
    const int psh = psh_env >= 3 ? psh_env : p->mix_pad_shift;
    const size_t smem = (size_t)((P * n + (P * n >> 3) + 1) * 2) * sizeof(double) + (size_t)2 * P * sizeof(int);   // (room for any shift >= 3)
    const i64 nblk = (npencil + 2 * P - 1) / (2 * P);
    MixStages S;
    S.nst = p->mix_nst;
    for (int s = 0; s < MX_MAX_STAGES; ++s) S.radix[s] = s < S.nst ? p->mix_radix[s] : 1;
#define NG_MIX_GO(LA, FU)                                                                                   \
    do {                                                                                                    \
        NG_FUNC_SMEM(200 * 1024, fft_mixed_kernel<LA, FU>);                                                 \
        fft_mixed_kernel<LA, FU><<<(unsigned)nblk, MX_NT, smem, st>>>(                                      \
            in, out, (const double2*)p->d_W, (const double2*)p->d_twn, p->outer, n, p->inner, P, psh, S, F);     \
    } while (0)
    if (last) { if (F.any) NG_MIX_GO(true, true); else NG_MIX_GO(true, false); }
    else      { if (F.any) NG_MIX_GO(false, true); else NG_MIX_GO(false, false); }
#undef NG_MIX_GO
    NG_TRACE("fft_mixed");
    NG_LAUNCH_CHECK();
    return NG_OK;
}
