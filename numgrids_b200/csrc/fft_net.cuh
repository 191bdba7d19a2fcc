// fft_net.cuh -- pieces shared by the periodic-derivative kernels (fft2.cu: contiguous axis, fft_tile.cu: strided
// axis through shared-memory tiles): compile-time radix networks held in registers, the per-pencil coefficient
// helpers of the fused applies, the rare non-finite slow path, and the mbarrier / bulk-copy wrappers.
// Reference: numgrids/diff.py:133-143 (out = real(ifft(W * fft(f)))).
#pragma once
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
// p[f] = w^f, f = 1 .. R-1, from ONE table value: p[f] = p[floor(f/2)] p[ceil(f/2)], at most ceil(log2 f) products deep
// (a few 1e-16; the periodic path's tolerance is 1e-11).  The shared-memory FFT passes need w^(j f) for every output f
// of a radix-R butterfly: R-1 scattered 16-byte table loads per work item were 3-4 x the load/store pipe traffic of the
// butterfly's own shared-memory accesses (ncu: l1tex 74 % busy, FP64 pipe 18 %).
template <int R>
__device__ __forceinline__ void twiddle_powers(double2 w, double2* p) {
    p[0] = make_double2(1.0, 0.0);
    p[1] = w;
#pragma unroll
    for (int f = 2; f < R; ++f) {
        const double2 a = p[f / 2], b = p[f - f / 2];
        p[f] = make_double2(fma(a.x, b.x, -(a.y * b.y)), fma(a.x, b.y, a.y * b.x));
    }
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
    const int nrep = (FUSED && F.mid.ptr) ? 2 : 1;           // double apply: D(mid * D(.)), see FuseDev
    for (int rep = 0; rep < nrep && !bad; ++rep) {
        if (rep == 1) {
            for (int k = lane; k < N; k += 32) {
                ix.i[ax] = k;
                double v = 0.0;                              // (zero-padded plans: `work` has nv samples, the rest is padding)
                if (k < nv) v = __dmul_rn(F.mid.ptr[coef_off(F.mid, ix)], S[k].x);
                if (ng_nonfinite(v)) { bad = true; v = 0.0; }
                S[k] = make_double2(v, 0.0);
            }
            bad = __any_sync(0xffffffffu, bad);
            if (bad) break;
            __syncwarp();
        }
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
// The elementwise work of a fused apply on a lane's RT samples of a pencil pair, restructured as straight-line
// passes over the register array with the (warp-uniform) "is this operand present / constant along the pencil"
// branches OUTSIDE the unrolled loops: per element the same operations in the same order as fuse_tail, but ~10
// branches per pair instead of ~10 per sample (ncu on the cfg4 gradient: 6.8 BRA + 7.6 BSSY/BSYNC + 15 ISETP per
// point and 5.4 instruction-fetch stalls per issue with the per-sample form).
// rv (zero-padded plans): only this lane's samples r < rv exist in memory -- operands are fetched for those only (the
// rest of the register array is the transform's padding: forced to zero by pencil_mul_all, never stored).
template <int RT>
__device__ __forceinline__ void pencil_mul_all(double2* x, const PencilCoef& ca, const PencilCoef& cb, bool va, bool vb, int rv = RT) {
    // x[r] *= coefficient; an absent pencil stays exactly zero (its table entries are pencil 0's and may be non-finite)
    if (ca.step == 0) {
        const double a = ca.c, b = cb.c;
#pragma unroll
        for (int r = 0; r < RT; ++r) { x[r].x = __dmul_rn(a, x[r].x); x[r].y = __dmul_rn(b, x[r].y); }
    } else {
#pragma unroll
        for (int r = 0; r < RT; ++r)
            if (r < rv) { x[r].x = __dmul_rn(ca.p[r * ca.step], x[r].x); x[r].y = __dmul_rn(cb.p[r * cb.step], x[r].y); }
    }
    if (!va || !vb || rv < RT) {
#pragma unroll
        for (int r = 0; r < RT; ++r) { if (!va || r >= rv) x[r].x = 0.0; if (!vb || r >= rv) x[r].y = 0.0; }
    }
}
template <int RT>
__device__ __forceinline__ void pencil_tail_all(const FuseDev& F, const PencilTail& ta, const PencilTail& tb, double2* x, int rv = RT) {
    if (ta.minuend) {
#pragma unroll
        for (int r = 0; r < RT; ++r)
            if (r < rv) { x[r].x = __dsub_rn(ta.minuend[r * ta.mstep], x[r].x); x[r].y = __dsub_rn(tb.minuend[r * tb.mstep], x[r].y); }
    }
    if (ta.post.p) {
        if (ta.post.step == 0) {
            const double a = ta.post.c, b = tb.post.c;
#pragma unroll
            for (int r = 0; r < RT; ++r) { x[r].x = __dmul_rn(a, x[r].x); x[r].y = __dmul_rn(b, x[r].y); }
        } else {
#pragma unroll
            for (int r = 0; r < RT; ++r)
                if (r < rv) { x[r].x = __dmul_rn(ta.post.p[r * ta.post.step], x[r].x); x[r].y = __dmul_rn(tb.post.p[r * tb.post.step], x[r].y); }
        }
    }
    if (ta.addend) {
#pragma unroll
        for (int r = 0; r < RT; ++r)
            if (r < rv) { x[r].x = __dadd_rn(ta.addend[r * ta.mstep], x[r].x); x[r].y = __dadd_rn(tb.addend[r * tb.mstep], x[r].y); }
    } else if (F.add_zero) {
#pragma unroll
        for (int r = 0; r < RT; ++r) { x[r].x = __dadd_rn(0.0, x[r].x); x[r].y = __dadd_rn(0.0, x[r].y); }
    }
    if (ta.div.p) {
        if (ta.div.step == 0) {
            const double a = ta.div.c, b = tb.div.c;
#pragma unroll
            for (int r = 0; r < RT; ++r) { x[r].x = __ddiv_rn(x[r].x, a); x[r].y = __ddiv_rn(x[r].y, b); }
        } else {
#pragma unroll
            for (int r = 0; r < RT; ++r)
                if (r < rv) { x[r].x = __ddiv_rn(x[r].x, ta.div.p[r * ta.div.step]); x[r].y = __ddiv_rn(x[r].y, tb.div.p[r * tb.div.step]); }
        }
    }
    if (F.finite) {
#pragma unroll
        for (int r = 0; r < RT; ++r) { x[r].x = ng_nonfinite(x[r].x) ? 0.0 : x[r].x; x[r].y = ng_nonfinite(x[r].y) ? 0.0 : x[r].y; }
    }
}

__device__ __forceinline__ PencilCoef pencil_coef(const FuseDev& F, const CoefDev& C, i64 lin0, int ax, int t, int R2) {
    PencilCoef c;
    Idx4 ix = unravel4(lin0, F.shape);
    ix.i[ax] = 0;
    c.init(C, ix, ax, t, R2);
    return c;
}
__device__ __forceinline__ PencilCoef pencil_pre(const FuseDev& F, i64 lin0, int ax, int t, int R2) {
    return pencil_coef(F, F.pre, lin0, ax, t, R2);
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

}  // namespace
