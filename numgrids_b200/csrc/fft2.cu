// fft2.cu -- production kernel of the periodic spectral derivative for a contiguous (last) axis,
// n = 64 .. 1024 a power of two:  out = real(ifft(W * fft(f))) (reference numgrids/diff.py:133-143).
//
// Two-pass ("four-step") FFT held in registers, one transform per T = n/RT lanes:
//   pass 1   each lane loads RT strided samples of TWO real pencils straight from global memory
//            (coalesced across lanes) as one complex value a + i b, runs a radix-RT DIF network in
//            registers (compile-time twiddles), applies the inter-pass twiddles exp(-2 pi i k t / n)
//            from a host-built table and parks the row in shared memory (padded: conflict-free);
//   pass 2   each lane owns G = RT/R2 sub-transforms of length R2: DIF forward, multiply by the
//            reference's (ik)^order table (already divided by n), DIT inverse -- all in registers;
//   pass 1'  rows come back from shared memory, conjugate twiddles, radix-RT DIT inverse, and the
//            real / imaginary parts are stored as the two output pencils.
// Because W is Hermitian the operator is real, so Re/Im of the complex result are D a and D b.
// Per complex sample: 2 shared-memory round trips (64 B, i.e. 32 B per real point) instead of the
// 2*log2(n) of the radix-2 kernel in fft.cu; only __syncwarp() is needed.  HBM traffic 16 B/point.
// FMA is used freely: this path's tolerance is 1e-11 (BASELINE.json).
#include "common.cuh"

namespace {

__device__ __forceinline__ double cos32(int k) {      // cos(2 pi k / 32), k = 0..15, correctly rounded
    switch (k) {
        case 0: return 1.0;
        case 1: return 0.9807852804032304;
        case 2: return 0.9238795325112867;
        case 3: return 0.8314696123025452;
        case 4: return 0.7071067811865476;
        case 5: return 0.5555702330196022;
        case 6: return 0.3826834323650898;
        case 7: return 0.19509032201612828;
        case 8: return 0.0;
        case 9: return -0.19509032201612828;
        case 10: return -0.3826834323650898;
        case 11: return -0.5555702330196022;
        case 12: return -0.7071067811865476;
        case 13: return -0.8314696123025452;
        case 14: return -0.9238795325112867;
        default: return -0.9807852804032304;
    }
}
__device__ __forceinline__ double sin32(int k) { return cos32(k < 8 ? 8 - k : k - 8); }
// sin(2 pi k/32) = cos(2 pi (8-k)/32) for k <= 8 and cos(2 pi (k-8)/32) for k >= 8 (both non-negative)

__device__ __forceinline__ double2 cmul(double2 a, double c, double s) {     // a * (c + i s)
    return make_double2(fma(a.x, c, -(a.y * s)), fma(a.x, s, a.y * c));
}

// multiply by exp(SIGN * 2 pi i k / 32), k in [0, 16), with the trivial cases folded at compile time
template <int SIGN>
__device__ __forceinline__ double2 rot32(double2 a, int k) {
    if (k == 0) return a;
    if (k == 8) return SIGN < 0 ? make_double2(a.y, -a.x) : make_double2(-a.y, a.x);
    return cmul(a, cos32(k), SIGN < 0 ? -sin32(k) : sin32(k));
}

__host__ __device__ __forceinline__ constexpr int bitrev(int v, int bits) {
    int r = 0;
    for (int b = 0; b < bits; ++b) if (v & (1 << b)) r |= 1 << (bits - 1 - b);
    return r;
}
template <int R> struct Log2 { static constexpr int v = 1 + Log2<R / 2>::v; };
template <> struct Log2<1> { static constexpr int v = 0; };

// Forward DIF network on x[0..R): natural order in, X[k] ends in x[bitrev(k)].
template <int R>
__device__ __forceinline__ void dif_forward(double2* x) {
#pragma unroll
    for (int s = R / 2; s >= 1; s >>= 1) {
#pragma unroll
        for (int i = 0; i < R; ++i) {
            if ((i & s) == 0) {
                const double2 u = x[i], v = x[i + s];
                x[i] = make_double2(u.x + v.x, u.y + v.y);
                const double2 d = make_double2(u.x - v.x, u.y - v.y);
                x[i + s] = rot32<-1>(d, (i & (s - 1)) * (16 / s));       // exp(-2 pi i j / (2s))
            }
        }
    }
}
// Inverse DIT network: x[bitrev(k)] = X[k] in, natural order out (unnormalised).
template <int R>
__device__ __forceinline__ void dit_inverse(double2* x) {
#pragma unroll
    for (int s = 1; s <= R / 2; s <<= 1) {
#pragma unroll
        for (int i = 0; i < R; ++i) {
            if ((i & s) == 0) {
                const double2 u = x[i];
                const double2 v = rot32<+1>(x[i + s], (i & (s - 1)) * (16 / s));
                x[i] = make_double2(u.x + v.x, u.y + v.y);
                x[i + s] = make_double2(u.x - v.x, u.y - v.y);
            }
        }
    }
}

// Non-finite samples (the reference's inf/nan rows on coordinate singularities before its final
// isfinite select, numgrids/curvilinear.py:236-244).  Two real pencils share one complex transform,
// so one bad pencil would poison its partner.  inf/nan survive every +, -, * of the transform and
// every output depends on every input of both pencils, so ONE test of a finished output per lane
// detects a poisoned pair at no cost to the clean path; the pair is then redone one pencil at a
// time by this warp-cooperative radix-2 routine (rare, slow): a pencil with a non-finite sample
// comes out all-NaN (what numpy.fft makes of it), a clean one gets its derivative.
template <bool FUSED>
__device__ __noinline__ void fft_slow_pencil(const double* __restrict__ in, double* __restrict__ out, i64 base,
                                             i64 step, int ax, int N, int log2n, double2* S,
                                             const double2* __restrict__ Wn, const double2* __restrict__ tw,
                                             const FuseDev& F, int nv /* samples in memory (zero padding above) */) {
    // sample k of the pencil sits at base + k*step and has coordinate k along grid axis `ax`
    const int lane = threadIdx.x & 31;
    Idx4 ix;
    if (FUSED) ix = unravel4(base, F.shape);
    bool bad = false;
    for (int k = lane; k < N; k += 32) {
        double v = k < nv ? in[base + k * step] : 0.0;
        if (FUSED && F.pre.ptr) {
            ix.i[ax] = k;
            v = __dmul_rn(F.pre.ptr[coef_off(F.pre, ix)], v);
        }
        if (ng_nonfinite(v)) { bad = true; v = 0.0; }
        S[k] = make_double2(v, 0.0);
    }
    bad = __any_sync(0xffffffffu, bad);
    if (!bad) {
        const int half = N >> 1;
        for (int st = log2n - 1; st >= 0; --st) {            // forward DIF: natural in, bit-reversed out
            const int m = 1 << st;
            for (int e = lane; e < half; e += 32) {
                const int j = e & (m - 1);
                const int i0 = ((e >> st) << (st + 1)) + j, i1 = i0 + m;
                const double2 w = tw[j << (log2n - 1 - st)];
                const double2 u = S[i0], v = S[i1];
                S[i0] = make_double2(u.x + v.x, u.y + v.y);
                S[i1] = cmul(make_double2(u.x - v.x, u.y - v.y), w.x, w.y);
            }
            __syncwarp();
        }
        for (int k = lane; k < N; k += 32) {
            const double2 w = Wn[__brev((unsigned)k) >> (32 - log2n)];
            S[k] = cmul(S[k], w.x, w.y);
        }
        __syncwarp();
        for (int st = 0; st < log2n; ++st) {                 // inverse DIT: bit-reversed in, natural out
            const int m = 1 << st;
            for (int e = lane; e < half; e += 32) {
                const int j = e & (m - 1);
                const int i0 = ((e >> st) << (st + 1)) + j, i1 = i0 + m;
                const double2 w = tw[j << (log2n - 1 - st)];
                const double2 u = S[i0], v = cmul(S[i1], w.x, -w.y);
                S[i0] = make_double2(u.x + v.x, u.y + v.y);
                S[i1] = make_double2(u.x - v.x, u.y - v.y);
            }
            __syncwarp();
        }
    }
    for (int k = lane; k < nv; k += 32) {
        double d = bad ? ng_nan() : S[k].x;
        if (FUSED) {
            ix.i[ax] = k;
            d = fuse_tail(F, d, base + k * step, ix);
        }
        out[base + k * step] = d;
    }
    __syncwarp();
}

// ng_fuse along a contiguous pencil, for a lane that owns samples idx = t + r*R2 (r compile-time):
// the coefficient offsets of the pencil are found once (one unravel per pencil), every operand
// becomes a per-lane base pointer plus r times a per-pencil step, and a coefficient that does
// not vary along the pencil (spherical J/h^2 tables depend on (r, theta) only) is loaded once --
// instead of an unravel + 4-term 64-bit offset per sample, which cost more instructions than the
// transform itself.
struct PencilCoef {
    const double* p;      // table entry of this lane's sample r = 0 (nullptr = absent)
    i64 step;             // elements between this lane's consecutive samples (0: constant along the pencil)
    double c;             // the value when step == 0
    __device__ __forceinline__ void init(const CoefDev& C, const Idx4& ix0, int vec_axis, int t, int R2) {
        const i64 s = C.s[vec_axis];
        p = C.ptr ? C.ptr + coef_off(C, ix0) + t * s : nullptr;
        step = s * R2;
        c = (p && step == 0) ? *p : 0.0;
    }
    __device__ __forceinline__ double at(int r) const { return step == 0 ? c : p[r * step]; }
};
struct PencilTail {
    PencilCoef post, div;
    const double *minuend, *addend;
    i64 mstep;            // elements between this lane's consecutive samples in the full-shape operands
    // the pencil's sample k sits at lin0 + k*step and has coordinate k along grid axis `ax`
    // (contiguous pencils: step 1, ax = the innermost axis)
    __device__ __forceinline__ void init(const FuseDev& F, i64 lin0, i64 step, int ax, int t, int R2) {
        Idx4 ix = unravel4(lin0, F.shape);
        ix.i[ax] = 0;
        post.init(F.post, ix, ax, t, R2);
        div.init(F.div, ix, ax, t, R2);
        minuend = F.minuend ? F.minuend + lin0 + t * step : nullptr;
        addend = F.addend ? F.addend + lin0 + t * step : nullptr;
        mstep = step * R2;
    }
    // same steps and order as fuse_tail (common.cuh); sample r of this lane sits at offset r*mstep
    __device__ __forceinline__ double apply(const FuseDev& F, double d, int r, int R2) const {
        if (minuend) d = __dsub_rn(minuend[r * mstep], d);
        if (post.p) d = __dmul_rn(post.at(r), d);
        if (addend) d = __dadd_rn(addend[r * mstep], d);
        else if (F.add_zero) d = __dadd_rn(0.0, d);
        if (div.p) d = __ddiv_rn(d, div.at(r));
        if (F.finite) d = ng_nonfinite(d) ? 0.0 : d;
        return d;
    }
};
__device__ __forceinline__ PencilCoef pencil_pre(const FuseDev& F, i64 lin0, int ax, int t, int R2) {
    PencilCoef c;
    Idx4 ix = unravel4(lin0, F.shape);
    ix.i[ax] = 0;
    c.init(F.pre, ix, ax, t, R2);
    return c;
}

// RT samples per lane in pass 1, R2 = n / RT = lanes per transform, G = RT / R2 sub-transforms per lane
// STR: the transform axis is not the contiguous one -- pencil p = o*inner + c of the [outer][n][inner] view
// starts at o*n*inner + c and its samples are `inner` elements apart.  The pair (2*pair, 2*pair+1) and the
// 2*TPW*8 pencils of a CTA are adjacent columns, so the 8-byte accesses of the lanes fall into sectors /
// lines that the neighbouring thread groups and warps of the CTA touch at the same time (L1 / L2 reuse).
// PAD: zero-padded plan (ng_fft_plan_create_padded): a pencil has `nv` <= N samples in memory, the rest of the
// N-point transform is zero; only the first nv outputs are stored.  Plain applies only (FUSED must be false).
template <int RT, int R2, bool FUSED, int MINB, bool STR, bool PAD = false>
__global__ void __launch_bounds__(256, MINB)
fft2pass_kernel(const double* __restrict__ in, double* __restrict__ out,
                const double2* __restrict__ Wn, const double2* __restrict__ tw, i64 npencil, i64 inner, int axis,
                int pair16,   // STR: the two pencils of a pair are one aligned 16-byte element of every row
                     const __grid_constant__ FuseDev F, int nv) {   // (address taken for the rare slow path: no local copy)
    static_assert(!(PAD && FUSED), "padded transforms are plain applies");
    constexpr int N = RT * R2, T = R2, G = RT / R2, TPW = 32 / T;
    const int NV = PAD ? nv : N;                       // samples per pencil in memory
    constexpr int PITCH = R2 + 1, UNITS = RT * PITCH;
    constexpr int LRT = Log2<RT>::v, LR2 = Log2<R2>::v;
    extern __shared__ double2 fsm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int tr = lane / T, t = lane - tr * T;
    const i64 pair = ((i64)blockIdx.x * (blockDim.x >> 5) + warp) * TPW + tr;
    const i64 pa = 2 * pair, pb = pa + 1;
    const bool va = pa < npencil, vb = pb < npencil;
    double2* S = fsm + (size_t)(warp * TPW + tr) * UNITS;
    // element k of pencil a / b: in[ba + k*step]; coordinate k along grid axis `ax`
    const i64 step = STR ? inner : 1;
    const int ax = STR ? axis : F.vec_axis;
    auto pencil_base = [&](i64 p) -> i64 {
        if (!STR) return p * NV;
        const i64 o = p / inner;
        return o * NV * inner + (p - o * inner);
    };
    const i64 ba = va ? pencil_base(pa) : 0, bb = vb ? pencil_base(pb) : 0;

    // multi-index of sample 0 of each pencil (one unravel per pencil; the FFT axis is the innermost
    // non-unit axis, so sample idx only changes that coordinate)
    double2 x[RT];
    // ---- pass 1: strided samples r*R2 + t of both pencils
    {
        PencilCoef ca, cb;
        const bool pre = FUSED && F.pre.ptr;
        if (pre) { ca = pencil_pre(F, ba, ax, t, R2); cb = pencil_pre(F, bb, ax, t, R2); }
#pragma unroll
        for (int r = 0; r < RT; ++r) {
            const int idx = r * R2 + t;
            double a = 0.0, b = 0.0;
            if (!PAD || idx < NV) {
                if (STR && pair16) {
                    if (va) { const double2 v = *reinterpret_cast<const double2*>(in + ba + idx * step); a = v.x; b = v.y; }
                } else {
                    if (va) a = in[ba + idx * step];
                    if (vb) b = in[bb + idx * step];
                }
            }
            if (pre) {                               // (an absent pencil must stay exactly zero)
                if (va) a = __dmul_rn(ca.at(r), a);
                if (vb) b = __dmul_rn(cb.at(r), b);
            }
            x[r] = make_double2(a, b);
        }
    }
    dif_forward<RT>(x);
#pragma unroll
    for (int q = 0; q < RT; ++q) {
        const int kr = bitrev(q, LRT);
        double2 v = x[q];
        if (kr != 0) {
            const double2 w = tw[kr * t];              // exp(-2 pi i kr t / n)
            v = cmul(v, w.x, w.y);
        }
        S[kr * PITCH + t] = v;
    }
    __syncwarp();
    // ---- pass 2: G sub-transforms of length R2, rows kr = t + T*g
#pragma unroll
    for (int g = 0; g < G; ++g) {
        const int kr = t + T * g;
        double2* y = x + g * R2;
#pragma unroll
        for (int j = 0; j < R2; ++j) y[j] = S[kr * PITCH + j];
        dif_forward<R2>(y);
#pragma unroll
        for (int q = 0; q < R2; ++q) {
            const int k2 = bitrev(q, LR2);
            const double2 w = Wn[kr + RT * k2];        // (ik)^order / n at frequency kr + RT*k2
            y[q] = cmul(y[q], w.x, w.y);
        }
        dit_inverse<R2>(y);
#pragma unroll
        for (int j = 0; j < R2; ++j) S[kr * PITCH + j] = y[j];
    }
    __syncwarp();
    // ---- pass 1': rows back, conjugate twiddles, inverse radix-RT
#pragma unroll
    for (int q = 0; q < RT; ++q) {
        const int kr = bitrev(q, LRT);
        double2 v = S[kr * PITCH + t];
        if (kr != 0) {
            const double2 w = tw[kr * t];
            v = cmul(v, w.x, -w.y);
        }
        x[q] = v;
    }
    dit_inverse<RT>(x);
    if (__any_sync(0xffffffffu, ng_nonfinite(x[0].x) || ng_nonfinite(x[0].y))) {   // poisoned pair: redo
        __syncwarp();
        const i64 p0 = ((i64)blockIdx.x * (blockDim.x >> 5) + warp) * TPW * 2;
        for (int q = 0; q < 2 * TPW; ++q)
            if (p0 + q < npencil)
                fft_slow_pencil<FUSED>(in, out, pencil_base(p0 + q), step, ax, N, LRT + LR2,
                                       fsm + (size_t)(warp * TPW + (q >> 1)) * UNITS, Wn, tw, F, NV);
        return;
    }
    PencilTail ta, tb;
    if (FUSED) { ta.init(F, ba, step, ax, t, R2); tb.init(F, bb, step, ax, t, R2); }
#pragma unroll
    for (int r = 0; r < RT; ++r) {
        const int idx = r * R2 + t;
        double a = x[r].x, b = x[r].y;
        if (FUSED) {
            if (va) a = ta.apply(F, a, r, R2);
            if (vb) b = tb.apply(F, b, r, R2);
        }
        if (PAD && idx >= NV) continue;
        if (STR && pair16) {
            if (va) *reinterpret_cast<double2*>(out + ba + idx * step) = make_double2(a, b);
        } else {
            if (va) out[ba + idx * step] = a;
            if (vb) out[bb + idx * step] = b;
        }
    }
}

// mbarrier / bulk-copy helpers of the ring kernel
__device__ __forceinline__ unsigned pf_smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void pf_mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(pf_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void pf_mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(pf_smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void pf_mbar_wait(unsigned long long* bar, unsigned parity) {
    const unsigned addr = pf_smem_u32(bar);
    unsigned done, spins = 0;
    do {
        if (++spins > (1u << 24)) __trap();          // bounded: a protocol bug must not hang the GPU
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(addr), "r"(parity) : "memory");
    } while (!done);
}
__device__ __forceinline__ void pf_bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
        ::"r"(pf_smem_u32(dst)), "l"(src), "r"(bytes), "r"(pf_smem_u32(bar)) : "memory");
}

// -------------------------------------------------------------------------------------------------
// Ring variant (n = 512, 1024) -- the production kernel of cfg3.  Differences from the persistent
// kernel above, all aimed at the two things ncu showed (profiles/r01_fft_cfg3_ncu_summary.txt):
// half of the twiddle / multiplier loads missed the ~28 KB of L1 left beside 197 KB of shared
// memory (long-scoreboard stalls), and 6 warps leave two of the four schedulers with one warp.
//   * both tables live in shared memory (twiddles transposed to [kr][t]: conflict-free);
//   * the exchange buffer is XOR-swizzled instead of padded (exactly 16 B per complex sample);
//   * NW = 8 warps (two per scheduler) share NSTAGE = 4 raw-input stages: the warp that has just
//     pulled unit j out of a stage refills it with unit j + NSTAGE, which its partner warp
//     (same scheduler, half a transform out of phase) consumes; one 16 KB cp.async.bulk per unit,
//     completing on the CONSUMER's own mbarrier so that a fast warp can never observe a stale phase.
// Shared memory: 2 tables + NW exchange buffers + NSTAGE stages = (2 + 8 + 4) * 16 KB = 224 KB.
// -------------------------------------------------------------------------------------------------
template <int RT, int R2, bool FUSED, int NW, int NSTAGE>
__global__ void __launch_bounds__(32 * NW, 1)
fft2pass_ring_kernel(const double* __restrict__ in, double* __restrict__ out,
                     const double2* __restrict__ Wn, const double2* __restrict__ tw, i64 npencil,
                     const __grid_constant__ FuseDev F) {   // (address taken for the rare slow path: no local copy)
    constexpr int N = RT * R2, T = R2, G = RT / R2, TPW = 32 / T;
    constexpr int LRT = Log2<RT>::v, LR2 = Log2<R2>::v;
    constexpr int PPU = 2 * TPW;                 // pencils per warp unit
    constexpr int UNIT = PPU * N;                // doubles per unit (16 KB)
    constexpr int ROWB = R2 * 16;                // bytes per exchange row
    static_assert(NW % NSTAGE == 0, "a stage is refilled by the warp that drained it");
    extern __shared__ __align__(1024) unsigned char rsm[];
    double2* TWS = reinterpret_cast<double2*>(rsm);                     // [RT][T]  exp(-2 pi i kr t / n)
    double2* WNS = TWS + N;                                             // [N]      (ik)^order / n
    unsigned char* Sall = reinterpret_cast<unsigned char*>(WNS + N);    // [NW*TPW][N] double2, swizzled
    double* rawall = reinterpret_cast<double*>(Sall + (size_t)NW * TPW * N * 16);   // [NSTAGE][UNIT]
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(rawall + (size_t)NSTAGE * UNIT);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int tr = lane / T, t = lane - tr * T;
    unsigned char* Sb = Sall + (size_t)(warp * TPW + tr) * N * 16;
    const unsigned t16 = (unsigned)t << 4;

    for (int e = threadIdx.x; e < N; e += 32 * NW) {
        TWS[e] = tw[(e / T) * (e % T)];
        WNS[e] = Wn[e];
    }
    if (threadIdx.x < NW) pf_mbar_init(bars + threadIdx.x, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();

    const i64 nunits = (npencil + PPU - 1) / PPU;
    const i64 njl = (nunits - (i64)blockIdx.x + gridDim.x - 1) / gridDim.x;   // units of this CTA
    auto issue = [&](i64 jl) {                   // one lane: unit jl of this CTA -> stage jl % NSTAGE
        const i64 p0 = (jl * gridDim.x + blockIdx.x) * PPU;
        const i64 left = npencil - p0;
        const unsigned bytes = (unsigned)((left < PPU ? left : PPU) * N * 8);
        unsigned long long* bar = bars + (int)(jl % NW);
        pf_mbar_expect_tx(bar, bytes);
        pf_bulk_g2s(rawall + (size_t)(jl % NSTAGE) * UNIT, in + p0 * N, bytes, bar);
    };
    if (warp < NSTAGE && lane == 0 && warp < njl) issue(warp);
    const double* raw = rawall + (size_t)(warp % NSTAGE) * UNIT + (size_t)(2 * tr) * N;
    unsigned long long* mybar = bars + warp;
    unsigned phase = 0;
    for (i64 jl = warp; jl < njl; jl += NW) {
        const i64 pa = (jl * gridDim.x + blockIdx.x) * PPU + 2 * tr, pb = pa + 1;
        const bool va = pa < npencil, vb = pb < npencil;
        PencilCoef ca, cb;
        const bool pre = FUSED && F.pre.ptr;
        if (pre) { ca = pencil_pre(F, va ? pa * N : 0, F.vec_axis, t, R2); cb = pencil_pre(F, vb ? pb * N : 0, F.vec_axis, t, R2); }
        pf_mbar_wait(mybar, phase);
        phase ^= 1;
        double2 x[RT];
#pragma unroll
        for (int r = 0; r < RT; ++r) {
            const int idx = r * R2 + t;
            double a = va ? raw[idx] : 0.0, b = vb ? raw[N + idx] : 0.0;
            if (pre) {                               // (an absent pencil must stay exactly zero)
                if (va) a = __dmul_rn(ca.at(r), a);
                if (vb) b = __dmul_rn(cb.at(r), b);
            }
            x[r] = make_double2(a, b);
        }
        __syncwarp();                              // the stage is drained
        if (lane == 0 && jl + NSTAGE < njl) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            issue(jl + NSTAGE);                    // refill it for the partner warp
        }
        dif_forward<RT>(x);
#pragma unroll
        for (int q = 0; q < RT; ++q) {
            const int kr = bitrev(q, LRT);
            double2 v = x[q];
            if (kr != 0) {
                const double2 w = TWS[kr * T + t];
                v = cmul(v, w.x, w.y);
            }
            *reinterpret_cast<double2*>(Sb + (t16 ^ ((kr & (R2 - 1)) << 4)) + kr * ROWB) = v;
        }
        __syncwarp();
#pragma unroll
        for (int g = 0; g < G; ++g) {
            const int kr = t + T * g;
            unsigned char* row = Sb + (size_t)kr * ROWB;
            double2* y = x + g * R2;
#pragma unroll
            for (int j = 0; j < R2; ++j) y[j] = *reinterpret_cast<const double2*>(row + (t16 ^ (j << 4)));
            dif_forward<R2>(y);
#pragma unroll
            for (int q = 0; q < R2; ++q) {
                const int k2 = bitrev(q, LR2);
                const double2 w = WNS[kr + RT * k2];
                y[q] = cmul(y[q], w.x, w.y);
            }
            dit_inverse<R2>(y);
#pragma unroll
            for (int j = 0; j < R2; ++j) *reinterpret_cast<double2*>(row + (t16 ^ (j << 4))) = y[j];
        }
        __syncwarp();
#pragma unroll
        for (int q = 0; q < RT; ++q) {
            const int kr = bitrev(q, LRT);
            double2 v = *reinterpret_cast<const double2*>(Sb + (t16 ^ ((kr & (R2 - 1)) << 4)) + kr * ROWB);
            if (kr != 0) {
                const double2 w = TWS[kr * T + t];
                v = cmul(v, w.x, -w.y);
            }
            x[q] = v;
        }
        dit_inverse<RT>(x);
        __syncwarp();                              // the exchange buffer is rewritten by the next unit
        if (__any_sync(0xffffffffu, ng_nonfinite(x[0].x) || ng_nonfinite(x[0].y))) {   // poisoned pair: redo
            const i64 p0 = (jl * gridDim.x + blockIdx.x) * PPU;
            for (int q = 0; q < PPU; ++q)
                if (p0 + q < npencil)
                    fft_slow_pencil<FUSED>(in, out, (p0 + q) * N, 1, F.vec_axis, N, LRT + LR2,
                                           reinterpret_cast<double2*>(Sall + (size_t)(warp * TPW + (q >> 1)) * N * 16),
                                           Wn, tw, F, N);
            continue;
        }
        PencilTail ta, tb;
        if (FUSED) { ta.init(F, va ? pa * N : 0, 1, F.vec_axis, t, R2); tb.init(F, vb ? pb * N : 0, 1, F.vec_axis, t, R2); }
        double* const oa = out + pa * N;
        double* const ob = oa + N;
#pragma unroll
        for (int r = 0; r < RT; ++r) {
            const int idx = r * R2 + t;
            double a = x[r].x, b = x[r].y;
            if (FUSED) {
                if (va) a = ta.apply(F, a, r, R2);
                if (vb) b = tb.apply(F, b, r, R2);
            }
            if (va) oa[idx] = a;
            if (vb) ob[idx] = b;
        }
    }
}

template <int RT, int R2, int NW, int NSTAGE>
int launch_ring(const ng_plan* p, const double* in, double* out, const FuseDev& F, cudaStream_t st) {
    constexpr int T = R2, TPW = 32 / T, N = RT * R2, PPU = 2 * TPW;
    const i64 npencil = p->outer;
    const i64 nunits = (npencil + PPU - 1) / PPU;
    i64 g = (nunits + NW - 1) / NW;
    const size_t smem = (size_t)2 * N * 16 + (size_t)NW * TPW * N * 16 + (size_t)NSTAGE * PPU * N * 8 + NW * 8;
#define NG_FFT2_RING(FU)                                                                          \
    do {                                                                                          \
        static int per_sm = 0;                           /* persistent: resident CTAs per SM */   \
        NG_FUNC_SMEM(smem, fft2pass_ring_kernel<RT, R2, FU, NW, NSTAGE>);                         \
        if (!per_sm) {                                                                            \
            int occ = 0;                                                                          \
            NG_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(                             \
                &occ, fft2pass_ring_kernel<RT, R2, FU, NW, NSTAGE>, 32 * NW, smem));              \
            per_sm = occ > 0 ? occ : 1;                                                           \
        }                                                                                         \
        if (g > (i64)148 * per_sm) g = (i64)148 * per_sm;                                         \
        fft2pass_ring_kernel<RT, R2, FU, NW, NSTAGE><<<(unsigned)g, 32 * NW, smem, st>>>(         \
            in, out, (const double2*)p->d_Wn, (const double2*)p->d_twn, npencil, F);              \
    } while (0)
    if (F.any) NG_FFT2_RING(true); else NG_FFT2_RING(false);
#undef NG_FFT2_RING
    NG_TRACE("fft2pass_ring");
    NG_LAUNCH_CHECK();
    return NG_OK;
}

template <int RT, int R2, int MINB, bool STR = false>
int launch(const ng_plan* p, const double* in, double* out, const FuseDev& F, cudaStream_t st) {
    constexpr int T = R2, TPW = 32 / T, WARPS = 8;
    const i64 npencil = p->outer * p->inner;            // last axis: one pencil per outer index
    const i64 pairs = (npencil + 1) / 2;
    const i64 per_block = (i64)WARPS * TPW;
    const unsigned grid = (unsigned)((pairs + per_block - 1) / per_block);
    const size_t smem = (size_t)WARPS * TPW * RT * (R2 + 1) * sizeof(double2);
#define NG_FFT2_GO(FU)                                                                            \
    do {                                                                                          \
        NG_FUNC_SMEM(smem, fft2pass_kernel<RT, R2, FU, MINB, STR>);                               \
        fft2pass_kernel<RT, R2, FU, MINB, STR><<<grid, 32 * WARPS, smem, st>>>(                         \
            in, out, (const double2*)p->d_Wn, (const double2*)p->d_twn, npencil, p->inner, p->axis,    \
            (int)(STR && p->inner % 2 == 0 && ((((uintptr_t)in | (uintptr_t)out) & 15) == 0)), F, RT * R2);   \
    } while (0)
    if (F.any) NG_FFT2_GO(true); else NG_FFT2_GO(false);
#undef NG_FFT2_GO
    NG_TRACE(STR ? "fft2pass<strided>" : "fft2pass");
    NG_LAUNCH_CHECK();
    return NG_OK;
}

// zero-padded plan, plain apply: the (RT*R2)-point transform on pencils of p->n samples
template <int RT, int R2, bool STR>
int launch_padded(const ng_plan* p, const double* in, double* out, const FuseDev& F, cudaStream_t st) {
    constexpr int T = R2, TPW = 32 / T, WARPS = 8;
    const i64 npencil = p->outer * p->inner;
    const i64 pairs = (npencil + 1) / 2;
    const i64 per_block = (i64)WARPS * TPW;
    const unsigned grid = (unsigned)((pairs + per_block - 1) / per_block);
    const size_t smem = (size_t)WARPS * TPW * RT * (R2 + 1) * sizeof(double2);
    NG_FUNC_SMEM(smem, fft2pass_kernel<RT, R2, false, 1, STR, true>);
    fft2pass_kernel<RT, R2, false, 1, STR, true><<<grid, 32 * WARPS, smem, st>>>(
        in, out, (const double2*)p->d_Wn, (const double2*)p->d_twn, npencil, p->inner, p->axis,
        (int)(STR && p->inner % 2 == 0 && ((((uintptr_t)in | (uintptr_t)out) & 15) == 0)), F, (int)p->n);
    NG_TRACE(STR ? "fft2pass<strided,padded>" : "fft2pass<padded>");
    NG_LAUNCH_CHECK();
    return NG_OK;
}

}  // namespace

// Returns NG_OK and sets *handled when this kernel family covers the plan.
int ng_fft2_try_launch(const ng_plan* p, const double* in, double* out, const FuseDev& F,
                       cudaStream_t st, bool* handled) {
    *handled = false;
    if (!p->d_Wn || !p->d_twn) return NG_OK;
    if (NG_ENV_INT("NG_FFT_TWOPASS", 1) == 0) return NG_OK;
    int rc = NG_OK;
    if (p->fft_len != p->n) {
        // zero-padded plan (any axis length): plain applies on the two-pass kernel, fused ones on the radix-2 kernel
        if (F.any) return NG_OK;
        const bool str = p->inner != 1;
        switch (p->fft_len) {
            case 1024: rc = str ? launch_padded<32, 32, true>(p, in, out, F, st) : launch_padded<32, 32, false>(p, in, out, F, st); break;
            case 512: rc = str ? launch_padded<32, 16, true>(p, in, out, F, st) : launch_padded<32, 16, false>(p, in, out, F, st); break;
            case 256: rc = str ? launch_padded<16, 16, true>(p, in, out, F, st) : launch_padded<16, 16, false>(p, in, out, F, st); break;
            default: return NG_OK;
        }
        *handled = (rc == NG_OK);
        return rc;
    }
    if (p->inner != 1) {
        // transform axis not contiguous: the one-shot kernel with strided pencils (adjacent columns per CTA)
        if (NG_ENV_INT("NG_FFT_STRIDED", 1) == 0) return NG_OK;   // A/B: 0 = radix-2 shared-memory fallback (fft.cu)
        switch (p->n) {
            case 1024: rc = launch<32, 32, 1, true>(p, in, out, F, st); break;
            case 512: rc = launch<32, 16, 1, true>(p, in, out, F, st); break;
            case 256: rc = launch<16, 16, 1, true>(p, in, out, F, st); break;
            case 128: rc = launch<16, 8, 1, true>(p, in, out, F, st); break;
            case 64: rc = launch<8, 8, 1, true>(p, in, out, F, st); break;
            default: return NG_OK;
        }
        *handled = (rc == NG_OK);
        return rc;
    }
    // the persistent ring kernel needs enough pencils to fill 148 CTAs and a 16-byte aligned input
    // (NG_FFT_PREFETCH, tuning: 0 = one-shot kernel, no input staging)
    const bool ring = NG_ENV_INT("NG_FFT_PREFETCH", 1) != 0 && (((uintptr_t)in & 15) == 0) && p->outer >= 2 * 148 * 6;
    const int ve = (int)NG_ENV_INT("NG_FFT_VARIANT", 0);          // tuning: warps / input stages per CTA
    const int variant = ve ? ve : (p->n >= 512 ? 2 : 3);   // measured: profiles/r01_tune_fft.txt
    switch (p->n) {
        case 1024:
            rc = !ring ? launch<32, 32, 1>(p, in, out, F, st)
                 : variant == 1 ? launch_ring<32, 32, 6, 6>(p, in, out, F, st)
                                : launch_ring<32, 32, 8, 4>(p, in, out, F, st);
            break;
        case 512:
            rc = !ring ? launch<32, 16, 1>(p, in, out, F, st)
                 : variant == 1 ? launch_ring<32, 16, 6, 6>(p, in, out, F, st)
                                : launch_ring<32, 16, 8, 4>(p, in, out, F, st);
            break;
        case 256:
            rc = !ring ? launch<16, 16, 1>(p, in, out, F, st)
                 : variant == 3 ? launch_ring<16, 16, 16, 8>(p, in, out, F, st)
                                : launch_ring<16, 16, 8, 4>(p, in, out, F, st);
            break;
        case 128:
            rc = !ring ? launch<16, 8, 1>(p, in, out, F, st)
                 : variant == 3 ? launch_ring<16, 8, 16, 8>(p, in, out, F, st)
                                : launch_ring<16, 8, 8, 4>(p, in, out, F, st);
            break;
        case 64: rc = launch<8, 8, 1>(p, in, out, F, st); break;
        default: return NG_OK;
    }
    *handled = (rc == NG_OK);
    return rc;
}
