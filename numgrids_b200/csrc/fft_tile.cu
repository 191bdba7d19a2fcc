// fft_tile.cu -- periodic spectral derivative along an axis that is NOT the contiguous one
// (cylindrical phi in the middle, doubly periodic boxes; reference numgrids/diff.py:133-143 applies
// numpy.fft along any axis).  n = 64 .. 1024 a power of two, array viewed as [outer][n][inner].
//
// Why a second kernel: with strided pencils every lane of the register kernel (fft2.cu) touches its
// own 32-byte sector -- 32 L1 wavefronts for 512 useful bytes per load instruction -- and the LSU, not
// HBM, sets the pace (0.27-0.31 of the copy peak, profiles/r01_tune_fft.txt).  Here no lane ever
// loads from or stores to global memory:
//   * a tile = all n rows of C = 8 or 16 adjacent columns (64 / 128 contiguous bytes per row) is
//     brought in by TMA (cp.async.bulk.tensor) into a hardware-swizzled slot;
//   * a warp owns the column PAIRS of its transforms: the 16-byte chunk c of every row.  The swizzle
//     (chunk index XOR row bits) makes a quarter-warp's 8 consecutive rows hit 8 distinct 16-byte bank
//     groups, so the column loads are conflict-free;
//   * the pair's n x 16 B column is also its exchange buffer between the two register passes (same
//     XOR-by-row trick as fft2.cu's ring kernel, mapped into the column) and finally receives the
//     derivative in the input's own positions;
//   * the finished tile leaves by TMA stores (full 64 / 128-byte row segments): every warp of the
//     group stores its own quarter / half of the rows and refills it with the same rows of the slot's
//     next tile as soon as its store has read them, so no single warp carries the whole drain.
// A CTA runs 8 warps as NG groups of WPT warps (one tile per group at a time) over NSLOT slots: tile s
// of the CTA sits in slot s % NSLOT, is worked by group s % NG, and the group's leader refills the
// slot with tile s + NSLOT once the store has drained it.  Consecutive uses of a slot belong to
// DIFFERENT groups, and a parity wait cannot tell "the load before mine has not landed yet" from "mine
// has landed" -- so a slot has TWO mbarriers, for its even and its odd uses: (2 NSLOT) % NG == 0 makes
// every phase of one mbarrier a tile of the same group, which waits for them in order (the usual
// single-consumer argument), and a phase is re-armed only after that group has passed the group barrier
// behind its wait.
// Plain applies only (fused composites keep the strided register kernel); HBM traffic 16 B/point.
#include "fft_net.cuh"

#include <cuda.h>

namespace {

__device__ __forceinline__ void ft_tma_load_3d(unsigned dst, const CUtensorMap* tm, int c0, int c1, int c2, unsigned bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
        ::"r"(dst), "l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(bar) : "memory");
}
__device__ __forceinline__ void ft_tma_store_3d(const CUtensorMap* tm, int c0, int c1, int c2, unsigned src) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%1, %2, %3}], [%4];"
                 ::"l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(src) : "memory");
}
// out[box] += tile (IEEE double adds performed by the memory system): the `addend == out` accumulate of the composites
__device__ __forceinline__ void ft_tma_reduce_add_3d(const CUtensorMap* tm, int c0, int c1, int c2, unsigned src) {
    asm volatile("cp.reduce.async.bulk.tensor.3d.global.shared::cta.add.tile.bulk_group [%0, {%1, %2, %3}], [%4];"
                 ::"l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(src) : "memory");
}
__device__ __forceinline__ void ft_group_bar(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// C columns per tile (row = C*8 bytes: 64 -> SWIZZLE_64B, 128 -> SWIZZLE_128B), NSLOT slots, 8 warps
// FUSED: ng_fuse work on the register array (pre, the double apply `mid`, post, + 0.0, div, finite -- fft_net.cuh);
// red_add: the apply accumulates into its own output (`addend == out`, no divide / select behind it): the tile is
// stored with a TMA reduce-add instead of loading the addend (out + d and d + out are the same IEEE sum).
template <int RT, int R2, int C, int NSLOT, int MINB, bool FUSED>
__global__ void __launch_bounds__(256, MINB)
fft_tile_kernel(const double* __restrict__ in, double* __restrict__ out, const double2* __restrict__ Wn,
                const double2* __restrict__ tw, i64 ntiles, int tiles_per_outer, i64 inner, int axis, int red_add,
                int nv,      // samples per pencil in memory (< N: zero-padded plan -- the rows past nv are the TMA's zero fill)
                const __grid_constant__ CUtensorMap tm_in, const __grid_constant__ CUtensorMap tm_out,
                const __grid_constant__ FuseDev F) {
    constexpr int N = RT * R2, T = R2, G = RT / R2, TPW = 32 / T, NW = 8;
    constexpr int LRT = Log2<RT>::v, LR2 = Log2<R2>::v;
    constexpr int RS = C * 8;                            // bytes per tile row
    constexpr int WPT = C / (2 * TPW);                   // warps per tile
    constexpr int NG = NW / WPT;                         // groups per CTA
    constexpr int BR = N / WPT;                          // rows per TMA box: every warp of a group moves its own
    constexpr int SLOT_B = N * RS, BOX_B = BR * RS;      // row range of the tile, in and out
    static_assert(RS == 32 || RS == 64 || RS == 128, "tile rows are one swizzle span");
    static_assert(WPT >= 1 && NW % WPT == 0, "groups tile the CTA");
    static_assert((2 * NSLOT) % NG == 0, "every other use of a slot belongs to the same group (see the header)");
    extern __shared__ __align__(1024) unsigned char tsm[];
    unsigned char* slots = tsm;                                                       // [NSLOT][N][RS], swizzled
    double2* TWS = reinterpret_cast<double2*>(tsm + (size_t)NSLOT * SLOT_B);          // [RT][T]  exp(-2 pi i kr t / n)
    double2* WNS = TWS + N;                                                           // [N]      (ik)^order / n
    unsigned long long* full = reinterpret_cast<unsigned long long*>(WNS + N);        // [NSLOT][2]
    int* poison = reinterpret_cast<int*>(full + 2 * NSLOT);                               // [NG][2]: flag of a group's even / odd tiles
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int tr = lane / T, t = lane - tr * T;
    const int grp = warp / WPT, wg = warp - grp * WPT;   // group, warp within the group
    const int cp = wg * TPW + tr;                        // this lane's column pair = 16-byte chunk of every row
    const bool leader = (wg == 0 && lane == 0);

    for (int e = threadIdx.x; e < N; e += 32 * NW) {
        TWS[e] = tw[(e / T) * (e % T)];
        WNS[e] = Wn[e];
    }
    if (threadIdx.x < 2 * NSLOT) pf_mbar_init(full + threadIdx.x, WPT);
    if (threadIdx.x < 2 * NG) poison[threadIdx.x] = 0;
    if ((pf_smem_u32(tsm) & 1023u) != 0) __trap();     // the swizzle formula below assumes 1 KB-aligned slots
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();

    const i64 nmine = (ntiles - (i64)blockIdx.x + gridDim.x - 1) / gridDim.x;         // tiles of this CTA
    auto issue = [&](i64 s, int b) {                     // one lane: row box b of tile s of this CTA -> slot s % NSLOT
        const i64 id = s * gridDim.x + blockIdx.x;
        const int o = (int)(id / tiles_per_outer);
        const int c0 = (int)(id - (i64)o * tiles_per_outer) * C;
        const int sl = (int)(s % NSLOT);
        unsigned long long* bar = full + 2 * sl + (int)((s / NSLOT) & 1);
        pf_mbar_expect_tx(bar, (unsigned)BOX_B);
        ft_tma_load_3d(pf_smem_u32(slots + (size_t)sl * SLOT_B + (size_t)b * BOX_B), &tm_in, c0, b * BR, o, pf_smem_u32(bar));
    };
    if (threadIdx.x == 0)
        for (int s = 0; s < NSLOT && s < nmine; ++s)
            for (int b = 0; b < WPT; ++b) issue(s, b);

    // byte offset of row e of this lane's pair column inside a slot (hardware swizzle: 16-byte chunk ^= row bits)
    auto col = [&](int e) -> unsigned {
        const int sw = RS == 128 ? (e & 7) : RS == 64 ? ((e >> 1) & 3) : ((e >> 2) & 1);
        return (unsigned)e * RS + (unsigned)((cp ^ sw) << 4);
    };

    int par = 0;
    for (i64 s = grp; s < nmine; s += NG, par ^= 1) {
        // the flag alternates per tile: a warp that runs ahead into the group's next tile never touches the flag a
        // slower warp is about to read
        int* myflag = poison + 2 * grp + par;
        const int sl = (int)(s % NSLOT);
        unsigned char* slot = slots + (size_t)sl * SLOT_B;
        pf_mbar_wait(full + 2 * sl + (int)((s / NSLOT) & 1), (unsigned)((s / (2 * NSLOT)) & 1));
        const i64 id = s * gridDim.x + blockIdx.x;
        const int o = (int)(id / tiles_per_outer);
        const int c0 = (int)(id - (i64)o * tiles_per_outer) * C;
        const bool vab = c0 + 2 * cp < inner;      // this lane's pencil pair exists (inner is even)
        const i64 lin0 = vab ? (i64)o * nv * inner + c0 + 2 * cp : 0;     // sample 0 of pencil a (b: + 1)
        double2 x[RT];
#pragma unroll
        for (int r = 0; r < RT; ++r) x[r] = *reinterpret_cast<const double2*>(slot + col(r * R2 + t));
        if (FUSED && F.pre.ptr) {
            const PencilCoef ca = pencil_coef(F, F.pre, lin0, axis, t, R2), cb = pencil_coef(F, F.pre, lin0 + (vab ? 1 : 0), axis, t, R2);
            pencil_mul_all<RT>(x, ca, cb, vab, vab);
        }
        const int nrep = (FUSED && F.mid.ptr) ? 2 : 1;     // double apply (FuseDev::mid): a run-time loop around the one body
#pragma unroll 1
        for (int rep = 0; rep < nrep; ++rep) {
        if (FUSED && rep == 1) {
            const PencilCoef ma = pencil_coef(F, F.mid, lin0, axis, t, R2), mb = pencil_coef(F, F.mid, lin0 + (vab ? 1 : 0), axis, t, R2);
            pencil_mul_all<RT>(x, ma, mb, vab, vab);
        }
        __syncwarp();                              // the column becomes the exchange buffer
        dif_forward<RT>(x);
#pragma unroll
        for (int q = 0; q < RT; ++q) {
            const int kr = bitrev(q, LRT);
            double2 v = x[q];
            if (kr != 0) {
                const double2 w = TWS[kr * T + t];
                v = cmul(v, w.x, w.y);
            }
            *reinterpret_cast<double2*>(slot + col(kr * R2 + (t ^ (kr & (R2 - 1))))) = v;
        }
        __syncwarp();
#pragma unroll
        for (int g = 0; g < G; ++g) {
            const int kr = t + T * g;
            double2* y = x + g * R2;
#pragma unroll
            for (int j = 0; j < R2; ++j) y[j] = *reinterpret_cast<const double2*>(slot + col(kr * R2 + (j ^ t)));
            dif_forward<R2>(y);
#pragma unroll
            for (int q = 0; q < R2; ++q) {
                const int k2 = bitrev(q, LR2);
                const double2 w = WNS[kr + RT * k2];
                y[q] = cmul(y[q], w.x, w.y);
            }
            dit_inverse<R2>(y);
#pragma unroll
            for (int j = 0; j < R2; ++j) *reinterpret_cast<double2*>(slot + col(kr * R2 + (j ^ t))) = y[j];
        }
        __syncwarp();
#pragma unroll
        for (int q = 0; q < RT; ++q) {
            const int kr = bitrev(q, LRT);
            double2 v = *reinterpret_cast<const double2*>(slot + col(kr * R2 + (t ^ (kr & (R2 - 1)))));
            if (kr != 0) {
                const double2 w = TWS[kr * T + t];
                v = cmul(v, w.x, -w.y);
            }
            x[q] = v;
        }
        dit_inverse<RT>(x);
        __syncwarp();                              // every lane has its rows back: the column is free again
        }
        const bool bad = __any_sync(0xffffffffu, ng_nonfinite(x[0].x) || ng_nonfinite(x[0].y));
        if (bad && lane == 0) atomicOr(myflag, 1);
        if (FUSED) {
            PencilTail ta, tb;
            ta.init(F, lin0, inner, axis, t, R2);
            tb.init(F, lin0 + (vab ? 1 : 0), inner, axis, t, R2);
            if (red_add) {
                ta.addend = tb.addend = nullptr;   // the reduce-add store performs `out + d`
                if (bad) {                         // poisoned pairs are redone below from `out` as it was: add -0.0 (the identity)
#pragma unroll
                    for (int r = 0; r < RT; ++r) x[r] = make_double2(-0.0, -0.0);
                }
            }
            if (!(red_add && bad)) pencil_tail_all<RT>(F, ta, tb, x);
        }
#pragma unroll
        for (int r = 0; r < RT; ++r) *reinterpret_cast<double2*>(slot + col(r * R2 + t)) = x[r];
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        ft_group_bar(1 + grp, 32 * WPT);
        if (lane == 0) {                           // every warp stores (and below refills) its own rows of the tile
            if (FUSED && red_add) ft_tma_reduce_add_3d(&tm_out, c0, wg * BR, o, pf_smem_u32(slot + (size_t)wg * BOX_B));
            else ft_tma_store_3d(&tm_out, c0, wg * BR, o, pf_smem_u32(slot + (size_t)wg * BOX_B));
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        if (*reinterpret_cast<volatile int*>(myflag)) {
            // Rare: a non-finite sample poisoned a pair (see fft_slow_pencil).  Once the tile's store has landed, the
            // pencils of the poisoned warps are redone one at a time straight from / to global memory; the drained
            // slot is the scratch space.
            if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
            ft_group_bar(1 + grp, 32 * WPT);
            if (bad) {
                for (int q = 0; q < 2 * TPW; ++q) {
                    const i64 cc = c0 + (i64)(wg * TPW) * 2 + q;
                    if (cc < inner)
                        fft_slow_pencil<FUSED>(in, out, (i64)o * nv * inner + cc, inner, axis, N, LRT + LR2,
                                               reinterpret_cast<double2*>(slot + (size_t)(wg * TPW + (q >> 1)) * N * 16),
                                               Wn, tw, F, nv);
                }
            }
            ft_group_bar(1 + grp, 32 * WPT);
            if (leader) *myflag = 0;
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            ft_group_bar(1 + grp, 32 * WPT);
        }
        if (lane == 0) {
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            if (s + NSLOT < nmine) issue(s + NSLOT, wg);
        }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

typedef CUresult (*ft_encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                 const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                 CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// 3-D map over the [outer][n][inner] view (inner fastest), box = C x rows x 1, swizzle = the row size
int ft_make_tmap(CUtensorMap* tm, const double* base, i64 outer, i64 n, i64 inner, int cols, int rows) {
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
    struct Key { const void* p; i64 o, n, i; int c, r; };
    struct Cache { Key k[8]; CUtensorMap m[8]; int used = 0, next = 0; };
    thread_local Cache cache;
    for (int i = 0; i < cache.used; ++i) {
        const Key& k = cache.k[i];
        if (k.p == base && k.o == outer && k.n == n && k.i == inner && k.c == cols && k.r == rows) { *tm = cache.m[i]; return NG_OK; }
    }
    const cuuint64_t gdim[3] = {(cuuint64_t)inner, (cuuint64_t)n, (cuuint64_t)outer};
    const cuuint64_t gstr[2] = {(cuuint64_t)inner * 8, (cuuint64_t)n * inner * 8};
    const cuuint32_t box[3] = {(cuuint32_t)cols, (cuuint32_t)rows, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    const CUresult r = e(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, const_cast<double*>(base), gdim, gstr, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE,
                         cols * 8 == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : cols * 8 == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        ng_set_error("cuTensorMapEncodeTiled (fft tile) failed with CUresult %d", (int)r);
        return NG_ECUDA;
    }
    const int slot = cache.used < 8 ? cache.used++ : (cache.next = (cache.next + 1) % 8);
    cache.k[slot] = Key{base, outer, n, inner, cols, rows};
    cache.m[slot] = *tm;
    return NG_OK;
}

template <int RT, int R2, int C, int NSLOT, int MINB>
int launch_tile(const ng_plan* p, const double* in, double* out, const FuseDev& F, int red_add, cudaStream_t st) {
    constexpr int N = RT * R2, WPT = C / (2 * (32 / R2)), NG = 8 / WPT, BR = N / WPT;
    static_assert(BR <= 256, "TMA box extent");
    const i64 tpo = (p->inner + C - 1) / C;
    const i64 ntiles = p->outer * tpo;
    CUtensorMap tmi, tmo;
    int rc = ft_make_tmap(&tmi, in, p->outer, p->n, p->inner, C, BR);      // p->n rows in memory (== N unless zero-padded)
    if (!rc) rc = ft_make_tmap(&tmo, out, p->outer, p->n, p->inner, C, BR);
    if (rc) return rc;
    const size_t smem = (size_t)NSLOT * N * C * 8 + (size_t)2 * N * 16 + 2 * NSLOT * 8 + NG * 8 + 16;
    i64 g = (ntiles + NG - 1) / NG;                      // at least one tile per group
    if (g > (i64)148 * MINB) g = (i64)148 * MINB;
#define NG_FFT_TILE(FU)                                                                                             \
    do {                                                                                                            \
        NG_FUNC_SMEM(smem, fft_tile_kernel<RT, R2, C, NSLOT, MINB, FU>);                                            \
        fft_tile_kernel<RT, R2, C, NSLOT, MINB, FU><<<(unsigned)g, 256, smem, st>>>(                                \
            in, out, (const double2*)p->d_Wn, (const double2*)p->d_twn, ntiles, (int)tpo, p->inner, p->axis, red_add, \
            (int)p->n, tmi, tmo, F);                                                                                           \
    } while (0)
    if (F.any) NG_FFT_TILE(true); else NG_FFT_TILE(false);
#undef NG_FFT_TILE
    NG_TRACE(F.any ? (red_add ? "fft_tile<fused,add>" : "fft_tile<fused>") : p->fft_len != p->n ? "fft_tile<padded>" : "fft_tile");
    NG_LAUNCH_CHECK();
    return NG_OK;
}

}  // namespace

// Plain applies along a non-contiguous axis.  Needs 16-byte aligned arrays and an even inner extent (TMA strides
// are multiples of 16 bytes) of at least one tile width; everything else stays on the strided register kernel.
int ng_fft_tile_try_launch(const ng_plan* p, const double* in, double* out, const FuseDev& F, cudaStream_t st,
                           bool* handled) {
    *handled = false;
    if (p->inner == 1 || !p->d_Wn || !p->d_twn) return NG_OK;
    if (p->fft_len != p->n && (F.any || p->n > p->fft_len)) return NG_OK;      // zero-padded plans: plain applies
    if (NG_ENV_INT("NG_FFT_TILE", 1) == 0) return NG_OK;        // A/B: 0 = strided register kernel
    if ((p->inner & 1) || p->inner < 16 || ((((uintptr_t)in | (uintptr_t)out) & 15) != 0)) return NG_OK;
    if (p->outer > 0x7fffffff || p->inner > 0x7fffffff) return NG_OK;
    // fused applies: everything but a minuend; an addend only as the in-place accumulate `out = out + D(...)` with no
    // divide / select behind it (the periodic pass of a Laplacian or divergence that is not the last axis), which
    // becomes a TMA reduce-add store.  The rest stays on the strided register kernel.
    int red_add = 0;
    if (F.any) {
        if (NG_ENV_INT("NG_FFT_TILE", 1) == 2) return NG_OK;    // A/B: 2 = plain applies only
        if (F.minuend) return NG_OK;
        if (F.addend) {
            if (F.addend != out || F.div.ptr || F.finite) return NG_OK;
            red_add = 1;
        }
    }
    int rc = NG_OK;
    switch (p->fft_len) {
        case 1024: rc = launch_tile<32, 32, 8, 3, 1>(p, in, out, F, red_add, st); break;
        case 512: rc = launch_tile<32, 16, 16, 3, 1>(p, in, out, F, red_add, st); break;
        case 256: rc = launch_tile<16, 16, 16, 3, 2>(p, in, out, F, red_add, st); break;
        case 128: rc = launch_tile<16, 8, 16, 6, 2>(p, in, out, F, red_add, st); break;
        case 64: rc = launch_tile<8, 8, 16, 12, 2>(p, in, out, F, red_add, st); break;
        default: return NG_OK;
    }
    *handled = (rc == NG_OK);
    return rc;
}
