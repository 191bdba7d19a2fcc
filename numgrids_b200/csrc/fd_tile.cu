// fd_tile.cu -- finite-difference stencil apply along the CONTIGUOUS axis through shared-memory row tiles
// (FinDiff(...)(f) / LogDiff along the last axis, reference numgrids/diff.py:88,259-262, and the fused passes of
// the curvilinear composites, numgrids/curvilinear.py:153-317).  Same arithmetic contract as fd.cu / fd_row.cu:
// acc = 0.0; acc = acc + w_k * x_k in stored tap order, DMUL + DADD, never FMA; then * inv_h_pow, / x (LogDiff),
// then the ng_fuse tail.
//
// Why a second contiguous-axis kernel: fd_row.cu takes its taps from neighbouring lanes by warp shuffles with
// register rotation at chunk edges and six head / tail body variants -- 57 thread-instructions per point plain and
// 100+ with fused elementwise work, i.e. issue-bound long before HBM (0.41-0.61 of the peak for the fused passes,
// profiles/r02_tune_fd_fused.txt).  Here the [rows][n] view is cut into tiles of RB rows x 128 columns with a halo
// of PH = 2 or 4 columns; a producer lane brings a tile in with ONE 2-D tensor-map TMA (out-of-bounds columns and
// rows zero-filled) through an S-slot ring with full / empty mbarriers; a consumer lane owns an aligned pair of
// outputs and reads its whole window with PH + 1 conflict-free LDS.128 -- no shuffles, no edge variants, no
// per-chunk range tests.  The (at most P) one-sided positions at a row end are left to a small second kernel
// (fd_tile_ends_kernel).  Full-shape tail operands (addend / minuend) are fetched as aligned
// pairs BEFORE the wait for the tile.  Slab halos stay with fd_row.cu.
#include <cuda.h>

#include "fd_fuse.cuh"

namespace {

constexpr int TW = 128;                       // output columns per tile

__device__ __forceinline__ unsigned ft_smem(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool ft_try(unsigned bar, unsigned parity) {
    unsigned done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    return done != 0;
}
__device__ __forceinline__ void ft_wait(unsigned bar, unsigned parity) {
    unsigned spins = 0;
    while (!ft_try(bar, parity))
        if (++spins > (1u << 24)) __trap();               // a protocol bug traps instead of hanging the GPU
}

// The (at most 2 P) positions at the two ends of every row take the one-sided schemes: sum_k w[k] * x(q + off[k]) from
// 0.0 in stored order.  They are NOT computed by the tile kernel (a divergent out-of-line routine on two of every
// 2 n / 64 warp passes cost 44 % of all instructions, ncu) -- it leaves those outputs unwritten, and this small kernel,
// launched behind it, computes them from global memory with the run-time form of the fused work: thread <-> output.
__global__ void __launch_bounds__(256)
fd_tile_ends_kernel(const double* __restrict__ in, double* __restrict__ out, i64 rows, int n, int axis, int P,
                    const __grid_constant__ FdTables T, const double* __restrict__ xdiv, const __grid_constant__ FuseDev F) {
    const i64 idx = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= rows * 2 * P) return;
    const i64 row = idx / (2 * P);
    const int e = (int)(idx - row * 2 * P);
    const bool fwd = e < P;
    const int q = fwd ? e : n - 2 * P + e;
    const double* w = fwd ? T.wf : T.wb;
    const int* off = fwd ? T.of : T.ob;
    const i64 base = row * (i64)n;
    Idx4 ix;
    if (F.any) ix = unravel4(base, F.shape);               // coordinate along `axis` (the last one) is 0
    // all taps in flight before the sum (a run-time loop would serialise their DRAM latencies)
    double x[NG_MAX_TAPS];
#pragma unroll
    for (int k = 0; k < NG_MAX_TAPS; ++k) {
        x[k] = 0.0;
        if (k < T.n_side) {
            const int t = q + off[k];
            x[k] = __ldg(in + base + t);
            if (F.any && F.pre.ptr) {
                ix.i[axis] = t;
                x[k] = __dmul_rn(__ldg(F.pre.ptr + coef_off(F.pre, ix)), x[k]);
            }
        }
    }
    double a = 0.0;
#pragma unroll
    for (int k = 0; k < NG_MAX_TAPS; ++k)
        if (k < T.n_side) a = __dadd_rn(a, __dmul_rn(w[k], x[k]));
    double r = __dmul_rn(a, T.inv_h_pow);
    if (xdiv) r = __ddiv_rn(r, xdiv[q]);
    if (F.any) {
        ix.i[axis] = q;
        r = fuse_tail(F, r, base + q, ix);
    }
    out[base + q] = r;
}

// P: half width of the central scheme; MODE: fused elementwise work (fd_fuse.cuh); RB rows per tile; S ring slots.
// 8 consumer warps + 1 producer warp.
template <int P, int MODE, int RB, int S>
__global__ void __launch_bounds__(288, 3)
fd_tile_kernel(const double* __restrict__ in, double* __restrict__ out, i64 rows, int n, int axis,
               const __grid_constant__ FdTables T, const double* __restrict__ xdiv, const __grid_constant__ FuseDev F,
               i64 ntiles, int nchunk, const __grid_constant__ CUtensorMap tm) {
    constexpr int PH = 2 * ((P + 1) / 2);                  // halo columns on each side (whole pairs)
    constexpr int W = TW + 2 * PH;                         // doubles per staged row
    static_assert(RB == 8, "one row of a tile per consumer warp");
    constexpr int NP = TW / 64;                            // 64-column passes of the warp's row
    constexpr unsigned TILE_B = RB * W * 8;
    constexpr bool FUSED = MODE != 0;
    extern __shared__ __align__(128) unsigned char tsm[];
    const unsigned ring = (ft_smem(tsm) + 127u) & ~127u;
    const unsigned full = ring + S * TILE_B, empty = full + 8 * S;
    // (the shuffle tells the compiler that the warp index -- and with it the row, its multi-index and its coefficient
    // offsets -- is warp-uniform: that arithmetic moves to the uniform datapath)
    const int lane = threadIdx.x & 31, warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);

    if (threadIdx.x == 0) {
        for (int s = 0; s < S; ++s) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(full + 8 * s), "r"(1) : "memory");
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(empty + 8 * s), "r"(8) : "memory");
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const i64 nmine = (ntiles - (i64)blockIdx.x + gridDim.x - 1) / gridDim.x;       // tiles of this CTA

    if (warp == 8) {                                       // ---- producer: one lane
        if (lane != 0) return;
        for (i64 t = 0; t < nmine; ++t) {
            const int s = (int)(t % S);
            if (t >= S) ft_wait(empty + 8 * s, (unsigned)(((t / S) - 1) & 1));
            const unsigned id = (unsigned)(t * gridDim.x + blockIdx.x);          // ntiles < 2^31 (launcher)
            const unsigned rb = id / (unsigned)nchunk;
            const int cc = (int)(id - rb * (unsigned)nchunk);
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(full + 8 * s), "r"(TILE_B) : "memory");
            asm volatile(
                "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                ::"r"(ring + s * TILE_B), "l"(&tm), "r"(cc * TW - PH), "r"((int)(rb * RB)), "r"(full + 8 * s) : "memory");
        }
        return;
    }

    // ---- consumers
    for (i64 t = 0; t < nmine; ++t) {
        const int s = (int)(t % S);
        const unsigned id = (unsigned)(t * gridDim.x + blockIdx.x);
        const unsigned rb = id / (unsigned)nchunk;
        const int c0 = (int)(id - rb * (unsigned)nchunk) * TW;
        // this warp's row of the tile, its coefficient offsets (once per tile), its NP pairs of the row
        const i64 row = (i64)rb * RB + warp;
        const bool row_ok = row < rows;
        const i64 base = (row_ok ? row : 0) * (i64)n;
        i64 pre0 = 0, pre_a = 0, post0 = 0, post_a = 0, div0 = 0, div_a = 0;
        if (FUSED) {
            const Idx4 ix = unravel4(base, F.shape);       // coordinate along `axis` (the last one) is 0
            if (fm_pre<MODE>(F)) { pre0 = coef_off(F.pre, ix); pre_a = F.pre.s[axis]; }
            if (fm_post<MODE>(F)) { post0 = coef_off(F.post, ix); post_a = F.post.s[axis]; }
            if (fm_div<MODE>(F)) { div0 = coef_off(F.div, ix); div_a = F.div.s[axis]; }
        }
        // coefficients that do not vary along the row: one load per tile, in flight while the tile lands
        double pre_c = 0.0, post_c = 0.0, div_c = 0.0;
        if (FUSED) {
            if (fm_pre<MODE>(F) && pre_a == 0) pre_c = __ldg(F.pre.ptr + pre0);
            if (fm_post<MODE>(F) && post_a == 0) post_c = __ldg(F.post.ptr + post0);
            if (fm_div<MODE>(F) && div_a == 0) div_c = __ldg(F.div.ptr + div0);
        }
        int pos[NP];
        bool live[NP];
        double2 mn[NP], ad[NP];
#pragma unroll
        for (int j = 0; j < NP; ++j) {
            pos[j] = c0 + j * 64 + 2 * lane;
            live[j] = row_ok && pos[j] < n;                // n is even: a pair is inside or outside
            if (!live[j]) pos[j] = 0;
            mn[j] = ad[j] = make_double2(0.0, 0.0);
            // full-shape tail operands: in flight while the tile lands (they may alias `out`: plain loads)
            if (MODE > 0 ? (MODE & FM_MINUEND) != 0 : (MODE < 0 && F.minuend != nullptr))
                mn[j] = *reinterpret_cast<const double2*>(F.minuend + base + pos[j]);
            if (MODE > 0 ? (MODE & FM_ADDEND) != 0 : (MODE < 0 && F.addend != nullptr))
                ad[j] = *reinterpret_cast<const double2*>(F.addend + base + pos[j]);
        }
        ft_wait(full + 8 * s, (unsigned)((t / S) & 1));
#pragma unroll
        for (int j = 0; j < NP; ++j) {
            // window: positions q0 .. q0 + 2 PH + 1 of the row, q0 = pos - PH (tile column 0 is c0 - PH)
            const unsigned a = ring + s * TILE_B + (unsigned)((warp * W + j * 64 + 2 * lane) * 8);
            double win[2 * PH + 2];
#pragma unroll
            for (int k = 0; k <= PH; ++k) {
                double2 v;
                asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(a + 16 * k));
                win[2 * k] = v.x;
                win[2 * k + 1] = v.y;
            }
            if (!live[j]) continue;
            const int q = pos[j];
            if (FUSED && fm_pre<MODE>(F)) {
                // `coeff * v` on the way in.  Window entries outside [0, n) are zero-filled and only feed central sums
                // that the one-sided schemes replace: their coefficient index is clamped, their value does not matter
                if (pre_a == 0) {
#pragma unroll
                    for (int k = 0; k < 2 * PH + 2; ++k) win[k] = __dmul_rn(pre_c, win[k]);
                } else {
#pragma unroll
                    for (int k = 0; k < 2 * PH + 2; ++k) {
                        int t_ = q - PH + k;
                        t_ = t_ < 0 ? 0 : (t_ >= n ? n - 1 : t_);
                        win[k] = __dmul_rn(__ldg(F.pre.ptr + pre0 + t_ * pre_a), win[k]);
                    }
                }
            }
            double ax = 0.0, ay = 0.0;
#pragma unroll
            for (int kk = 0; kk < 2 * P + 1; ++kk) {
                ax = __dadd_rn(ax, __dmul_rn(T.wc[kk], win[PH - P + kk]));
                ay = __dadd_rn(ay, __dmul_rn(T.wc[kk], win[PH - P + 1 + kk]));
            }
            double2 r;
            r.x = __dmul_rn(ax, T.inv_h_pow);
            r.y = __dmul_rn(ay, T.inv_h_pow);
            if (xdiv) {
                const double2 xd = ldg128(xdiv + q);
                r.x = __ddiv_rn(r.x, xd.x);
                r.y = __ddiv_rn(r.y, xd.y);
            }
            if (MODE > 0) {
                double2 pc = make_double2(0.0, 0.0), dc = pc;
                if (MODE & FM_POST) pc = coef_pair(F.post.ptr, post0, q, post_a, post_c);
                if (MODE & FM_DIV) dc = coef_pair(F.div.ptr, div0, q, div_a, div_c);
                r.x = fm_tail_vals<MODE>(r.x, mn[j].x, pc.x, ad[j].x, dc.x);
                r.y = fm_tail_vals<MODE>(r.y, mn[j].y, pc.y, ad[j].y, dc.y);
            } else if (MODE < 0) {
                // run-time form: same steps and order as fuse_tail_off, the full-shape operands already fetched
                const i64 po = F.post.ptr ? post0 + q * post_a : 0, dv = F.div.ptr ? div0 + q * div_a : 0;
                r.x = fuse_tail_vals(F, r.x, mn[j].x, ad[j].x, po, dv);
                r.y = fuse_tail_vals(F, r.y, mn[j].y, ad[j].y, po + post_a, dv + div_a);
            }
            // the P positions at either row end belong to fd_tile_ends_kernel
            const bool sx = q >= P && q < n - P, sy = q + 1 >= P && q + 1 < n - P;
            if (sx && sy) *reinterpret_cast<double2*>(out + base + q) = r;
            else if (sx) out[base + q] = r.x;
            else if (sy) out[base + q + 1] = r.y;
        }
        __syncwarp();
        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(empty + 8 * s) : "memory");
    }
}

typedef CUresult (*ft_encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                 const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                 CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// 2-D map over the [rows][n] view (n fastest), box = cols x box_rows, no swizzle, zero OOB fill; cached per thread
int ft_make_tmap(CUtensorMap* tm, const double* base, i64 rows, i64 n, int cols, int box_rows) {
    static std::atomic<ft_encode_fn> encode{nullptr};
    ft_encode_fn e = encode.load(std::memory_order_acquire);
    if (!e) {
        void* fp = nullptr;
        cudaDriverEntryPointQueryResult qres;
        NG_CUDA_OK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &qres));
        if (!fp || qres != cudaDriverEntryPointSuccess) {
            ng_set_error("cuTensorMapEncodeTiled is not available from this driver");
            return NG_ECUDA;
        }
        e = (ft_encode_fn)fp;
        encode.store(e, std::memory_order_release);
    }
    struct Key { const void* p; i64 r, n; int c, b; };
    struct Cache { Key k[8]; CUtensorMap m[8]; int used = 0, next = 0; };
    thread_local Cache cache;
    for (int i = 0; i < cache.used; ++i) {
        const Key& k = cache.k[i];
        if (k.p == base && k.r == rows && k.n == n && k.c == cols && k.b == box_rows) { *tm = cache.m[i]; return NG_OK; }
    }
    const cuuint64_t gdim[2] = {(cuuint64_t)n, (cuuint64_t)rows};
    const cuuint64_t gstr[1] = {(cuuint64_t)n * 8};
    const cuuint32_t box[2] = {(cuuint32_t)cols, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = e(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double*>(base), gdim, gstr, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        ng_set_error("cuTensorMapEncodeTiled (fd tile) failed with CUresult %d", (int)r);
        return NG_ECUDA;
    }
    const int slot = cache.used < 8 ? cache.used++ : (cache.next = (cache.next + 1) % 8);
    cache.k[slot] = Key{base, rows, n, cols, box_rows};
    cache.m[slot] = *tm;
    return NG_OK;
}

template <int P, int MODE>
int launch_tile(const ng_plan* p, const double* in, double* out, const FuseDev& F, cudaStream_t st) {
    constexpr int RB = 8, S = 6;
    constexpr int PH = 2 * ((P + 1) / 2), W = TW + 2 * PH;
    const i64 rows = p->outer;
    const int n = (int)p->n;
    const int nchunk = (n + TW - 1) / TW;
    const i64 ntiles = ((rows + RB - 1) / RB) * nchunk;
    CUtensorMap tm;
    int rc = ft_make_tmap(&tm, in, rows, n, W, RB);
    if (rc) return rc;
    const size_t smem = (size_t)S * RB * W * 8 + 16 * S + 128;
    NG_FUNC_SMEM(smem, fd_tile_kernel<P, MODE, RB, S>);
    i64 g = ntiles < 148 * 3 ? ntiles : 148 * 3;
    fd_tile_kernel<P, MODE, RB, S><<<(unsigned)g, 288, smem, st>>>(in, out, rows, n, p->axis, p->fd, p->d_xdiv, F, ntiles,
                                                                   nchunk, tm);
    NG_LAUNCH_CHECK();
    const i64 nend = rows * 2 * P;
    fd_tile_ends_kernel<<<(unsigned)((nend + 255) / 256), 256, 0, st>>>(in, out, rows, n, p->axis, P, p->fd, p->d_xdiv, F);
    NG_TRACE(MODE > 0 ? "fd_tile<mode>" : MODE < 0 ? "fd_tile<fused>" : "fd_tile");
    NG_LAUNCH_CHECK();
    return NG_OK;
}

}  // namespace

// Contiguous-axis applies without slab halos: plain, the twelve compile-time fuse patterns (P = 2) and the run-time
// form.  *handled stays false when the shape does not suit the tiles (fd_row.cu / fd.cu take it).
int ng_fd_tile_try_launch(int P, const ng_plan* p, const double* in, double* out, const FuseDev& F, cudaStream_t st,
                          bool* handled) {
    *handled = false;
    const int want = (int)NG_ENV_INT("NG_FD_TILE", 1);      // A/B: 0 = off, 1 = fused applies, 2 = plain applies too
    if (want == 0 || (!F.any && want < 2)) return NG_OK;
    // small arrays are launch-bound: the one-load-deep kernels finish sooner (NG_FD_TILE_MIN: points from which the tiles take over)
    if (p->inner != 1 || p->n % 2 != 0 || p->outer >= (1LL << 31) || p->size < NG_ENV_INT("NG_FD_TILE_MIN", 1LL << 18)) return NG_OK;
    if (p->fd.n_side > p->n || p->size >= (1LL << 40)) return NG_OK;     // (the kernel's tile arithmetic is 32-bit)
    int rc = NG_OK;
    if (!F.any) {
        switch (P) {
            case 1: rc = launch_tile<1, 0>(p, in, out, F, st); break;
            case 2: rc = launch_tile<2, 0>(p, in, out, F, st); break;
            case 3: rc = launch_tile<3, 0>(p, in, out, F, st); break;
            case 4: rc = launch_tile<4, 0>(p, in, out, F, st); break;
            default: return NG_OK;
        }
        *handled = (rc == NG_OK);
        return rc;
    }
    if (P == 2) {
        switch (fuse_mask(F)) {
#define NG_MODE(M) case (M): rc = launch_tile<2, (M)>(p, in, out, F, st); *handled = (rc == NG_OK); return rc;
            NG_FD_FUSE_MODES(NG_MODE)
#undef NG_MODE
            default: break;
        }
    }
    switch (P) {
        case 1: rc = launch_tile<1, -1>(p, in, out, F, st); break;
        case 2: rc = launch_tile<2, -1>(p, in, out, F, st); break;
        case 3: rc = launch_tile<3, -1>(p, in, out, F, st); break;
        case 4: rc = launch_tile<4, -1>(p, in, out, F, st); break;
        default: return NG_OK;
    }
    *handled = (rc == NG_OK);
    return rc;
}
