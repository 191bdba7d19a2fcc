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
#include "fft_net.cuh"

namespace {

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
    const int rv = PAD ? (NV - t + R2 - 1) / R2 : RT;      // this lane's samples r*R2 + t < NV (zero-padded plans)

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
            x[r] = make_double2(a, b);
        }
        if (pre) pencil_mul_all<RT>(x, ca, cb, va, vb, rv);
    }
    // double apply (FuseDev::mid): the transform runs twice with the pointwise multiply between, the pair stays
    // in registers; a run-time loop, so the unrolled body exists once
    const int nrep = (FUSED && F.mid.ptr) ? 2 : 1;
#pragma unroll 1
    for (int rep = 0; rep < nrep; ++rep) {
    if (FUSED && rep == 1) {
        const PencilCoef ma = pencil_coef(F, F.mid, ba, ax, t, R2), mb = pencil_coef(F, F.mid, bb, ax, t, R2);
        pencil_mul_all<RT>(x, ma, mb, va, vb, rv);            // (also zeroes the padding again: `work` has NV samples)
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
    __syncwarp();                                  // the exchange buffer is rewritten by the second transform
    }
    if (__any_sync(0xffffffffu, ng_nonfinite(x[0].x) || ng_nonfinite(x[0].y))) {   // poisoned pair: redo
        __syncwarp();
        const i64 p0 = ((i64)blockIdx.x * (blockDim.x >> 5) + warp) * TPW * 2;
        for (int q = 0; q < 2 * TPW; ++q)
            if (p0 + q < npencil)
                fft_slow_pencil<FUSED>(in, out, pencil_base(p0 + q), step, ax, N, LRT + LR2,
                                       fsm + (size_t)(warp * TPW + (q >> 1)) * UNITS, Wn, tw, F, NV);
        return;
    }
    if (FUSED) {
        PencilTail ta, tb;
        ta.init(F, ba, step, ax, t, R2);
        tb.init(F, bb, step, ax, t, R2);
        pencil_tail_all<RT>(F, ta, tb, x, rv);
    }
#pragma unroll
    for (int r = 0; r < RT; ++r) {
        const int idx = r * R2 + t;
        const double a = x[r].x, b = x[r].y;
        if (PAD && idx >= NV) continue;
        if (STR && pair16) {
            if (va) *reinterpret_cast<double2*>(out + ba + idx * step) = make_double2(a, b);
        } else {
            if (va) out[ba + idx * step] = a;
            if (vb) out[bb + idx * step] = b;
        }
    }
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
// PAD: zero-padded plan (ng_fft_plan_create_padded): a pencil has nv <= N samples in memory (nv even), the rest of the
// N-point transform is zero; a unit's pencils are still ONE contiguous run (one bulk copy); plain applies only.
template <int RT, int R2, bool FUSED, int NW, int NSTAGE, bool PAD = false>
__global__ void __launch_bounds__(32 * NW, 1)
fft2pass_ring_kernel(const double* __restrict__ in, double* __restrict__ out,
                     const double2* __restrict__ Wn, const double2* __restrict__ tw, i64 npencil,
                     const __grid_constant__ FuseDev F, int nv) {   // (F: address taken for the rare slow path: no local copy)
    constexpr int N = RT * R2, T = R2, G = RT / R2, TPW = 32 / T;
    const int NV = PAD ? nv : N;                 // samples per pencil in memory
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
        const unsigned bytes = (unsigned)((left < PPU ? left : PPU) * NV * 8);
        unsigned long long* bar = bars + (int)(jl % NW);
        pf_mbar_expect_tx(bar, bytes);
        pf_bulk_g2s(rawall + (size_t)(jl % NSTAGE) * UNIT, in + p0 * NV, bytes, bar);
    };
    if (warp < NSTAGE && lane == 0 && warp < njl) issue(warp);
    const int rv = PAD ? (NV - t + R2 - 1) / R2 : RT;       // this lane's samples r*R2 + t < NV (zero-padded plans)
    const double* raw = rawall + (size_t)(warp % NSTAGE) * UNIT + (size_t)(2 * tr) * NV;
    unsigned long long* mybar = bars + warp;
    unsigned phase = 0;
    for (i64 jl = warp; jl < njl; jl += NW) {
        const i64 pa = (jl * gridDim.x + blockIdx.x) * PPU + 2 * tr, pb = pa + 1;
        const bool va = pa < npencil, vb = pb < npencil;
        PencilCoef ca, cb;
        const bool pre = FUSED && F.pre.ptr;
        if (pre) { ca = pencil_pre(F, va ? pa * NV : 0, F.vec_axis, t, R2); cb = pencil_pre(F, vb ? pb * NV : 0, F.vec_axis, t, R2); }
        pf_mbar_wait(mybar, phase);
        phase ^= 1;
        double2 x[RT];
#pragma unroll
        for (int r = 0; r < RT; ++r) {
            const int idx = r * R2 + t;
            const bool in_mem = !PAD || idx < NV;
            x[r] = make_double2(va && in_mem ? raw[idx] : 0.0, vb && in_mem ? raw[NV + idx] : 0.0);
        }
        if (pre) pencil_mul_all<RT>(x, ca, cb, va, vb, rv);
        __syncwarp();                              // the stage is drained
        if (lane == 0 && jl + NSTAGE < njl) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            issue(jl + NSTAGE);                    // refill it for the partner warp
        }
        const int nrep = (FUSED && F.mid.ptr) ? 2 : 1;     // double apply (FuseDev::mid), see fft2pass_kernel
#pragma unroll 1
        for (int rep = 0; rep < nrep; ++rep) {
        if (FUSED && rep == 1) {
            const PencilCoef ma = pencil_coef(F, F.mid, va ? pa * NV : 0, F.vec_axis, t, R2),
                             mb = pencil_coef(F, F.mid, vb ? pb * NV : 0, F.vec_axis, t, R2);
            pencil_mul_all<RT>(x, ma, mb, va, vb, rv);
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
        __syncwarp();                              // the exchange buffer is rewritten by the next unit / transform
        }
        if (__any_sync(0xffffffffu, ng_nonfinite(x[0].x) || ng_nonfinite(x[0].y))) {   // poisoned pair: redo
            const i64 p0 = (jl * gridDim.x + blockIdx.x) * PPU;
            for (int q = 0; q < PPU; ++q)
                if (p0 + q < npencil)
                    fft_slow_pencil<FUSED>(in, out, (p0 + q) * NV, 1, F.vec_axis, N, LRT + LR2,
                                           reinterpret_cast<double2*>(Sall + (size_t)(warp * TPW + (q >> 1)) * N * 16),
                                           Wn, tw, F, NV);
            continue;
        }
        if (FUSED) {
            PencilTail ta, tb;
            ta.init(F, va ? pa * NV : 0, 1, F.vec_axis, t, R2);
            tb.init(F, vb ? pb * NV : 0, 1, F.vec_axis, t, R2);
            pencil_tail_all<RT>(F, ta, tb, x, rv);
        }
        double* const oa = out + pa * NV;
        double* const ob = oa + NV;
#pragma unroll
        for (int r = 0; r < RT; ++r) {
            const int idx = r * R2 + t;
            if (PAD && idx >= NV) continue;
            if (va) oa[idx] = x[r].x;
            if (vb) ob[idx] = x[r].y;
        }
    }
}

template <int RT, int R2, int NW, int NSTAGE, bool PAD = false>
int launch_ring(const ng_plan* p, const double* in, double* out, const FuseDev& F, cudaStream_t st) {
    constexpr int T = R2, TPW = 32 / T, N = RT * R2, PPU = 2 * TPW;
    const i64 npencil = p->outer;
    const i64 nunits = (npencil + PPU - 1) / PPU;
    i64 g = (nunits + NW - 1) / NW;
    const size_t smem = (size_t)2 * N * 16 + (size_t)NW * TPW * N * 16 + (size_t)NSTAGE * PPU * N * 8 + NW * 8;
#define NG_FFT2_RING(FU)                                                                          \
    do {                                                                                          \
        static int per_sm = 0;                           /* persistent: resident CTAs per SM */   \
        NG_FUNC_SMEM(smem, fft2pass_ring_kernel<RT, R2, FU, NW, NSTAGE, PAD>);                    \
        if (!per_sm) {                                                                            \
            int occ = 0;                                                                          \
            NG_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(                             \
                &occ, fft2pass_ring_kernel<RT, R2, FU, NW, NSTAGE, PAD>, 32 * NW, smem));         \
            per_sm = occ > 0 ? occ : 1;                                                           \
        }                                                                                         \
        if (g > (i64)148 * per_sm) g = (i64)148 * per_sm;                                         \
        fft2pass_ring_kernel<RT, R2, FU, NW, NSTAGE, PAD><<<(unsigned)g, 32 * NW, smem, st>>>(    \
            in, out, (const double2*)p->d_Wn, (const double2*)p->d_twn, npencil, F, (int)p->n);   \
    } while (0)
    if (F.any) NG_FFT2_RING(true); else NG_FFT2_RING(false);
#undef NG_FFT2_RING
    NG_TRACE(PAD ? "fft2pass_ring<padded>" : "fft2pass_ring");
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
#define NG_FFT2_PAD(FU)                                                                           \
    do {                                                                                          \
        NG_FUNC_SMEM(smem, fft2pass_kernel<RT, R2, FU, 1, STR, true>);                            \
        fft2pass_kernel<RT, R2, FU, 1, STR, true><<<grid, 32 * WARPS, smem, st>>>(                \
            in, out, (const double2*)p->d_Wn, (const double2*)p->d_twn, npencil, p->inner, p->axis,    \
            (int)(STR && p->inner % 2 == 0 && ((((uintptr_t)in | (uintptr_t)out) & 15) == 0)), F, (int)p->n);   \
    } while (0)
    if (F.any) NG_FFT2_PAD(true); else NG_FFT2_PAD(false);
#undef NG_FFT2_PAD
    NG_TRACE(STR ? "fft2pass<strided,padded>" : "fft2pass<padded>");
    NG_LAUNCH_CHECK();
    return NG_OK;
}

}  // namespace

// A fused apply with FuseDev::mid (double apply) is honoured exactly when ng_fft2_try_launch takes the plan.
bool ng_fft_twice_supported(const ng_plan* p) {
    if (p->kind != PK_FFT || !p->fft_pow2 || !p->d_Wn || !p->d_twn) return false;
    if (NG_ENV_INT("NG_FFT_TWOPASS", 1) == 0 || NG_ENV_INT("NG_FFT_TWICE", 1) == 0) return false;
    if (p->fft_len != p->n)                                  // zero-padded plans: the 256 / 512 / 1024-point kernels
        return p->fft_len == 256 || p->fft_len == 512 || p->fft_len == 1024;
    if (p->inner != 1 && NG_ENV_INT("NG_FFT_STRIDED", 1) == 0) return false;
    return p->n == 64 || p->n == 128 || p->n == 256 || p->n == 512 || p->n == 1024;
}

// Returns NG_OK and sets *handled when this kernel family covers the plan.
int ng_fft2_try_launch(const ng_plan* p, const double* in, double* out, const FuseDev& F,
                       cudaStream_t st, bool* handled) {
    *handled = false;
    if (!p->d_Wn || !p->d_twn) return NG_OK;
    if (NG_ENV_INT("NG_FFT_TWOPASS", 1) == 0) return NG_OK;
    int rc = NG_OK;
    if (p->fft_len != p->n) {
        // zero-padded plan (any axis length): the two-pass kernels up to a 1024-point transform, the radix-2 kernel above
        const bool str = p->inner != 1;
        if (str) {                                          // strided pencils: TMA tiles (fft_tile.cu), rows past n zero-filled
            rc = ng_fft_tile_try_launch(p, in, out, F, st, handled);
            if (rc || *handled) return rc;
        }
        // contiguous pencils, enough of them, an even length: the persistent ring kernel (one bulk copy per unit)
        if (!str && NG_ENV_INT("NG_FFT_PREFETCH", 1) != 0 && (((uintptr_t)in & 15) == 0) && p->outer >= 2 * 148 * 6 && p->n % 2 == 0) {
            switch (p->fft_len) {
                case 1024: rc = launch_ring<32, 32, 8, 4, true>(p, in, out, F, st); *handled = (rc == NG_OK); return rc;
                case 512: rc = launch_ring<32, 16, 8, 4, true>(p, in, out, F, st); *handled = (rc == NG_OK); return rc;
                case 256: rc = launch_ring<16, 16, 16, 8, true>(p, in, out, F, st); *handled = (rc == NG_OK); return rc;
                default: break;
            }
        }
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
        rc = ng_fft_tile_try_launch(p, in, out, F, st, handled);  // plain applies: TMA tiles (fft_tile.cu)
        if (rc || *handled) return rc;
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
