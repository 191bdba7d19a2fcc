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
#include <type_traits>

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

// -------------------------------------------------------------------------------------------------
// Production kernel for the config shapes (axis not last, n % 128 == 0, inner % 128 == 0):
// 128 x 128 output tile, 256 threads, 8 x 8 register tile per thread laid out as 2-wide pairs
// strided by 32 (conflict-free LDS.128: 'a' pairs broadcast per half-warp, 'b' pairs contiguous),
// k panels of 16 double-buffered in shared memory with cp.async (LDGSTS) so the next panel streams
// in under the 2048 FP64 instructions of the current one.  Same arithmetic contract as above.
// -------------------------------------------------------------------------------------------------
#ifndef NG_DENSE_FK
#define NG_DENSE_FK 16
#endif
#ifndef NG_DENSE_MINB
#define NG_DENSE_MINB 2
#endif
constexpr int FM = 128, FN = 128, FK = NG_DENSE_FK, FKS = FK / 16;   // FKS: staging copies scale with the panel depth

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gmem_src) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(gmem_src) : "memory");
}
// zero-filling variants (src-size 0 reads nothing and writes zeros): rows / columns past the array
__device__ __forceinline__ void cp_async16_z(void* smem_dst, const void* gmem_src, bool valid) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    const int sz = valid ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gmem_src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async8_z(void* smem_dst, const void* gmem_src, bool valid) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    const int sz = valid ? 8 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(d), "l"(gmem_src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// NCP = column pairs per thread: 4 -> 128-column tiles (8x8 outputs per thread, big problems),
// 1 -> 32-column tiles (8x2 per thread, several CTAs per SM): small grids such as cfg4 need the finer
// tiles to balance 148 SMs (256 big tiles would leave a 27 % idle tail).
// LASTAX: the contraction axis is the contiguous one (pencil r = row of a [R][n] matrix, Out = F Mt):
// the F panel is transposed on the fly by 8-byte cp.async (source coalesced along k), outputs are
// stored as (i, i+1) pairs of one pencil.
// GEN: any n and any (even) column count -- Mt is the zero-padded copy (pitch `npitch`, a multiple of 128),
// k panels past n are zero-filled (0 * 0 adds +-0 to a sum that started at +0.0: the bits do not change),
// column tiles are cut per outer index with their tail columns zero-filled, rows / columns past the array
// are not stored and row groups of 32 past n are not computed.  !GEN is the exact-shape kernel unchanged.
// THR = threads per CTA: 256 (16 threads across the columns), or 128 (8 across: with NCP = 2 the same 128 x 32 tile
// as <NCP = 1, 256> but 8 x 4 outputs per thread -- 6 LDS.128 per 64 FP64 instructions instead of 5 per 32).
template <bool DESC, bool FMA, bool FUSED, int NCP, bool LASTAX, bool GEN, int THR = 256>
__global__ void __launch_bounds__(THR, NCP == 4 ? 1 : (THR == 128 ? 3 : NG_DENSE_MINB))
dense_fast_kernel(const double* __restrict__ in, double* __restrict__ out,
                  const double* __restrict__ Mt, int n, i64 inner, int axis, FuseDev F,
                  int npitch, i64 npencil_last, int tiles_per_o) {
    constexpr int TXN = THR / 16, CS = 2 * TXN;       // threads across the columns, column stride between a thread's pairs
    constexpr int TFN = CS * NCP, HALF = TFN / 2;
    constexpr int MQ = 4 * FKS * (256 / THR);         // 16-byte M-panel chunks per thread
    // LASTAX stages F by 8-byte copies with consecutive threads walking k: rows one pitch apart.  A
    // pitch of TFN doubles puts all 16 k of a column in one bank (16-way conflict on every copy);
    // two doubles of padding spread them over 8 banks and keep rows 16-byte aligned for the LDS.128.
    constexpr int TFP = TFN + (LASTAX ? 2 : 0);
    extern __shared__ __align__(16) double dsm[];
    double (*Ms)[FK][FM] = reinterpret_cast<double (*)[FK][FM]>(dsm);                  // [2][FK][FM]
    double (*Fs)[FK][TFP] = reinterpret_cast<double (*)[FK][TFP]>(dsm + 2 * FK * FM);  // [2][FK][TFP]
    const i64 col0 = (i64)blockIdx.x * TFN;         // flattened (o, c) column; tiles never straddle o
    const int row0 = blockIdx.y * FM;
    i64 o = LASTAX ? 0 : col0 / inner, c0 = LASTAX ? 0 : col0 - o * inner;
    int ncv = TFN;                                  // valid columns (pencils) of this tile
    if (GEN) {
        if (LASTAX) {
            ncv = (int)(npencil_last - col0 < TFN ? npencil_last - col0 : TFN);
        } else {
            o = blockIdx.x / tiles_per_o;
            c0 = (i64)(blockIdx.x - o * tiles_per_o) * TFN;
            ncv = (int)(inner - c0 < TFN ? inner - c0 : TFN);
        }
    }
    const int jmax = GEN ? ((n - row0 + 31) / 32 < 4 ? (n - row0 + 31) / 32 : 4) : 4;   // 32-row groups with rows < n
    // element (k, c): fbase + k*inner + c;  LASTAX: pencil r = col0 + c, element (k, c) at fbase + c*n + k
    const double* fbase = LASTAX ? in + col0 * (i64)n : in + o * (i64)n * inner + c0;
    const int tid = threadIdx.x, tx = tid % TXN, ty = tid / TXN;

    double acc[8][2 * NCP];
#pragma unroll
    for (int a = 0; a < 8; ++a)
#pragma unroll
        for (int b = 0; b < 2 * NCP; ++b) acc[a][b] = 0.0;

    const int npanel = GEN ? (n + FK - 1) / FK : n / FK;
    Idx4 pre_ix;                                   // multi-index of (row 0, this thread's chunk column)
    if (FUSED && F.pre.ptr && !LASTAX) {
        const int chp = (tid % HALF) * 2;
        pre_ix = unravel4((o * n) * inner + c0 + ((GEN && chp >= ncv) ? 0 : chp), F.shape);
    }
    // prologue coefficients (`coeff * v[i]`) of the panel in flight, fetched with the panel so that
    // their global-load latency hides under the previous panel's arithmetic
    double pcn[2 * NCP * FKS];
#pragma unroll
    for (int q = 0; q < 2 * NCP * FKS; ++q) pcn[q] = 0.0;
    auto stage = [&](int p, int buf) {
        const int k0 = p * FK;
        if (FUSED && !LASTAX && F.pre.ptr) {
#pragma unroll
            for (int q = 0; q < NCP * FKS; ++q) {
                const int kk = (tid + q * THR) / HALF;
                if (GEN && (k0 + kk >= n || ((tid + q * THR) % HALF) * 2 >= ncv)) {   // zero-filled chunk
                    pcn[2 * q] = pcn[2 * q + 1] = 0.0;
                    continue;
                }
                Idx4 ix = pre_ix;
                ix.i[axis] = k0 + kk;
                const i64 po = coef_off(F.pre, ix);
                pcn[2 * q] = F.pre.ptr[po];
                pcn[2 * q + 1] = F.pre.ptr[po + F.pre.s[F.vec_axis]];
            }
        }
        // M panel: 16 x 128 doubles = 1024 16-byte chunks (4 per thread); F panel: 16 x TFN (NCP per thread)
#pragma unroll
        for (int q = 0; q < MQ; ++q) {
            const int e = tid + q * THR;
            const int kk = e >> 6, ch = (e & 63) * 2;
            cp_async16(&Ms[buf][kk][ch], Mt + (i64)(k0 + kk) * npitch + row0 + ch);
        }
        if (LASTAX) {
            // 16 x TFN doubles as 8-byte copies: consecutive threads walk k (128 B contiguous per pencil)
#pragma unroll
            for (int q = 0; q < 2 * NCP * FKS; ++q) {
                const int e = tid + q * THR;
                const int kk = e % FK, c = e / FK;
                if (GEN) {
                    const bool ok = k0 + kk < n && c < ncv;
                    cp_async8_z(&Fs[buf][kk][c], ok ? fbase + (i64)c * n + k0 + kk : in, ok);
                } else {
                    cp_async8(&Fs[buf][kk][c], fbase + (i64)c * n + k0 + kk);
                }
            }
        } else {
#pragma unroll
            for (int q = 0; q < NCP * FKS; ++q) {
                const int e = tid + q * THR;
                const int kk = e / HALF, ch = (e % HALF) * 2;
                if (GEN) {
                    const bool ok = k0 + kk < n && ch < ncv;
                    cp_async16_z(&Fs[buf][kk][ch], ok ? fbase + (i64)(k0 + kk) * inner + ch : in, ok);
                } else {
                    cp_async16(&Fs[buf][kk][ch], fbase + (i64)(k0 + kk) * inner + ch);
                }
            }
        }
        cp_async_commit();
    };
    // the panel loop, with the number of live 32-row groups (JM) a compile-time constant: a run-time
    // test inside the k loop would fence the scheduling of LDS and FP64 instructions across it
    auto run = [&](auto jm_c) {
    constexpr int JM = decltype(jm_c)::value;
    stage(DESC ? npanel - 1 : 0, 0);
    for (int pp = 0; pp < npanel; ++pp) {
        const int buf = pp & 1;
        double pcc[2 * NCP * FKS];                  // this panel's prologue coefficients
#pragma unroll
        for (int q = 0; q < 2 * NCP * FKS; ++q) pcc[q] = pcn[q];
        if (pp + 1 < npanel) {
            stage(DESC ? npanel - 2 - pp : pp + 1, buf ^ 1);
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        if (FUSED && F.pre.ptr) {
            // prologue multiply (`coeff * v[i]`, curvilinear.py:212): each thread scales the chunks it
            // copied itself (visible to it after wait_group), before the panel is published
            const int k0 = (DESC ? npanel - 1 - pp : pp) * FK;
            if (LASTAX) {
#pragma unroll
                for (int q = 0; q < 2 * NCP * FKS; ++q) {
                    const int e = tid + q * THR;
                    const int kk = e % FK, c = e / FK;
                    if (GEN && (k0 + kk >= n || c >= ncv)) continue;      // zero-filled element stays zero
                    const Idx4 ix = unravel4((col0 + c) * (i64)n + k0 + kk, F.shape);
                    Fs[buf][kk][c] = __dmul_rn(F.pre.ptr[coef_off(F.pre, ix)], Fs[buf][kk][c]);
                }
            } else {
#pragma unroll
                for (int q = 0; q < NCP * FKS; ++q) {
                    const int e = tid + q * THR;
                    const int kk = e / HALF, ch = (e % HALF) * 2;
                    Fs[buf][kk][ch] = __dmul_rn(pcc[2 * q], Fs[buf][kk][ch]);
                    Fs[buf][kk][ch + 1] = __dmul_rn(pcc[2 * q + 1], Fs[buf][kk][ch + 1]);
                }
            }
        }
        __syncthreads();
#pragma unroll 4
        for (int q = 0; q < FK; ++q) {
            const int kk = DESC ? FK - 1 - q : q;
            double a[8], b[2 * NCP];
#pragma unroll
            for (int j = 0; j < JM; ++j) {
                const double2 av = *reinterpret_cast<const double2*>(&Ms[buf][kk][ty * 2 + 32 * j]);
                a[2 * j] = av.x; a[2 * j + 1] = av.y;
            }
#pragma unroll
            for (int j = 0; j < NCP; ++j) {
                const double2 bv = *reinterpret_cast<const double2*>(&Fs[buf][kk][tx * 2 + CS * j]);
                b[2 * j] = bv.x; b[2 * j + 1] = bv.y;
            }
#pragma unroll
            for (int x = 0; x < 2 * JM; ++x) {
#pragma unroll
                for (int y = 0; y < 2 * NCP; ++y) {
                    if (FMA) acc[x][y] = fma(a[x], b[y], acc[x][y]);
                    else acc[x][y] = __dadd_rn(acc[x][y], __dmul_rn(a[x], b[y]));
                }
            }
        }
        __syncthreads();
    }
    };
    if (!GEN || jmax >= 4) run(std::integral_constant<int, 4>{});
    else if (jmax == 3) run(std::integral_constant<int, GEN ? 3 : 4>{});
    else if (jmax == 2) run(std::integral_constant<int, GEN ? 2 : 4>{});
    else run(std::integral_constant<int, GEN ? 1 : 4>{});
    if (LASTAX) {
        // out[(col0 + c)*n + i]: store (i, i+1) pairs of one pencil; one unravel per pencil
#pragma unroll
        for (int y = 0; y < 2 * NCP; ++y) {
            if (GEN && tx * 2 + (y & 1) + CS * (y >> 1) >= ncv) continue;
            const i64 r = col0 + tx * 2 + (y & 1) + CS * (y >> 1);
            i64 po = 0, dv = 0, pa = 0, da = 0;             // offsets of (pencil r, i = 0), strides along i
            if (FUSED) {
                const Idx4 ix0 = unravel4(r * (i64)n, F.shape);
                if (F.post.ptr) { po = coef_off(F.post, ix0); pa = F.post.s[axis]; }
                if (F.div.ptr) { dv = coef_off(F.div, ix0); da = F.div.s[axis]; }
            }
            // the pairs (i, i+1), i = row0 + ty*2 + 32 j: the elementwise tail as straight-line passes with the
            // (uniform) operand tests outside the loops -- same operations per element in the same order as fuse_tail
            double2 v[4];
            i64 lin[4];
            bool live[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int i = row0 + ty * 2 + 32 * j;
                live[j] = !(GEN && i >= n);                 // n is even: the pair (i, i+1) is inside or outside
                lin[j] = r * (i64)n + (live[j] ? i : 0);
                v[j] = make_double2(acc[2 * j][y], acc[2 * j + 1][y]);
            }
            if (FUSED) {
                if (F.minuend) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) { const double2 m = *reinterpret_cast<const double2*>(F.minuend + lin[j]); v[j].x = __dsub_rn(m.x, v[j].x); v[j].y = __dsub_rn(m.y, v[j].y); }
                }
                if (F.post.ptr) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) { const i64 i = lin[j] - r * (i64)n; v[j].x = __dmul_rn(F.post.ptr[po + i * pa], v[j].x); v[j].y = __dmul_rn(F.post.ptr[po + (i + 1) * pa], v[j].y); }
                }
                if (F.addend) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) { const double2 a = *reinterpret_cast<const double2*>(F.addend + lin[j]); v[j].x = __dadd_rn(a.x, v[j].x); v[j].y = __dadd_rn(a.y, v[j].y); }
                } else if (F.add_zero) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) { v[j].x = __dadd_rn(0.0, v[j].x); v[j].y = __dadd_rn(0.0, v[j].y); }
                }
                if (F.div.ptr) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) { const i64 i = lin[j] - r * (i64)n; v[j].x = __ddiv_rn(v[j].x, F.div.ptr[dv + i * da]); v[j].y = __ddiv_rn(v[j].y, F.div.ptr[dv + (i + 1) * da]); }
                }
                if (F.finite) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) { v[j].x = isfinite(v[j].x) ? v[j].x : 0.0; v[j].y = isfinite(v[j].y) ? v[j].y : 0.0; }
                }
            }
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (live[j]) *reinterpret_cast<double2*>(out + lin[j]) = v[j];
        }
        return;
    }
    // ---- epilogue: rows ty*2 + {0,1} + 32*jr, column pairs tx*2 + 32*jc.  One unravel per column
    // pair (the row only changes the coordinate along `axis`).
#pragma unroll
    for (int jc = 0; jc < NCP; ++jc) {
        if (GEN && tx * 2 + CS * jc >= ncv) continue;                   // inner is even: whole pairs
        const i64 lin0 = (o * n) * inner + c0 + tx * 2 + CS * jc;       // row 0 of this column pair
        i64 po = 0, dv = 0, pa = 0, da = 0, pv = 0, dvv = 0;   // offsets of row 0, strides along i / the pair
        if (FUSED) {
            const Idx4 ix0 = unravel4(lin0, F.shape);          // (row 0: the coordinate along `axis` is 0)
            if (F.post.ptr) { po = coef_off(F.post, ix0); pa = F.post.s[axis]; pv = F.post.s[F.vec_axis]; }
            if (F.div.ptr) { dv = coef_off(F.div, ix0); da = F.div.s[axis]; dvv = F.div.s[F.vec_axis]; }
        }
        double2 v[8];
        i64 lin[8];
        int ri[8];
        bool live[8];
#pragma unroll
        for (int x = 0; x < 8; ++x) {
            const int i = row0 + ty * 2 + (x & 1) + 32 * (x >> 1);
            live[x] = !(GEN && i >= n);
            ri[x] = live[x] ? i : 0;
            lin[x] = lin0 + (i64)ri[x] * inner;
            v[x] = make_double2(acc[x][2 * jc], acc[x][2 * jc + 1]);
        }
        if (FUSED) {      // straight-line passes, operand tests outside the loops (see the contiguous-axis epilogue above)
            if (F.minuend) {
#pragma unroll
                for (int x = 0; x < 8; ++x) { const double2 m = *reinterpret_cast<const double2*>(F.minuend + lin[x]); v[x].x = __dsub_rn(m.x, v[x].x); v[x].y = __dsub_rn(m.y, v[x].y); }
            }
            if (F.post.ptr) {
#pragma unroll
                for (int x = 0; x < 8; ++x) { const i64 o_ = po + ri[x] * pa; v[x].x = __dmul_rn(F.post.ptr[o_], v[x].x); v[x].y = __dmul_rn(F.post.ptr[o_ + pv], v[x].y); }
            }
            if (F.addend) {
#pragma unroll
                for (int x = 0; x < 8; ++x) { const double2 a = *reinterpret_cast<const double2*>(F.addend + lin[x]); v[x].x = __dadd_rn(a.x, v[x].x); v[x].y = __dadd_rn(a.y, v[x].y); }
            } else if (F.add_zero) {
#pragma unroll
                for (int x = 0; x < 8; ++x) { v[x].x = __dadd_rn(0.0, v[x].x); v[x].y = __dadd_rn(0.0, v[x].y); }
            }
            if (F.div.ptr) {
#pragma unroll
                for (int x = 0; x < 8; ++x) { const i64 o_ = dv + ri[x] * da; v[x].x = __ddiv_rn(v[x].x, F.div.ptr[o_]); v[x].y = __ddiv_rn(v[x].y, F.div.ptr[o_ + dvv]); }   // the pair stays inside the innermost row
            }
            if (F.finite) {
#pragma unroll
                for (int x = 0; x < 8; ++x) { v[x].x = isfinite(v[x].x) ? v[x].x : 0.0; v[x].y = isfinite(v[x].y) ? v[x].y : 0.0; }
            }
        }
#pragma unroll
        for (int x = 0; x < 8; ++x)
            if (live[x]) *reinterpret_cast<double2*>(out + lin[x]) = v[x];
    }
}

template <bool DESC, bool FMA, int NCP, bool LASTAX, bool GEN = false, int THR = 256>
int launch_fast2(const ng_plan* p, const double* in, double* out, const FuseDev& F, cudaStream_t st) {
    constexpr int TFN = (THR / 8) * NCP;
    const i64 ncols = p->outer * p->inner;
    dim3 grid((unsigned)(ncols / TFN), (unsigned)(p->n / FM));
    const int tiles_per_o = (int)((p->inner + TFN - 1) / TFN);
    if (GEN) {
        grid.x = LASTAX ? (unsigned)((p->outer + TFN - 1) / TFN) : (unsigned)(p->outer * tiles_per_o);
        grid.y = (unsigned)((p->n + FM - 1) / FM);
    }
    const double* Mt = GEN ? p->d_Mtp : p->d_Mt;
    const int npitch = GEN ? p->n_pad : (int)p->n;
    const size_t smem = (size_t)2 * FK * (FM + TFN + (LASTAX ? 2 : 0)) * sizeof(double);
#define NG_DENSE_FAST(FU)                                                                       \
    do {                                                                                        \
        NG_FUNC_SMEM(smem, dense_fast_kernel<DESC, FMA, FU, NCP, LASTAX, GEN, THR>);            \
        dense_fast_kernel<DESC, FMA, FU, NCP, LASTAX, GEN, THR><<<grid, THR, smem, st>>>(               \
            in, out, Mt, (int)p->n, p->inner, p->axis, F, npitch, p->outer, tiles_per_o);       \
    } while (0)
    if (F.any) NG_DENSE_FAST(true); else NG_DENSE_FAST(false);
#undef NG_DENSE_FAST
    NG_TRACE(THR == 128 ? (LASTAX ? "dense_fast<8x4,lastax>" : "dense_fast<8x4>")
             : GEN ? (LASTAX ? "dense_fast<gen,lastax>" : "dense_fast<gen>")
                 : NCP == 4 ? (LASTAX ? "dense_fast<ncp4,lastax>" : "dense_fast<ncp4>")
                 : NCP == 2 ? (LASTAX ? "dense_fast<ncp2,lastax>" : "dense_fast<ncp2>")
                            : (LASTAX ? "dense_fast<ncp1,lastax>" : "dense_fast<ncp1>"));
    NG_LAUNCH_CHECK();
    return NG_OK;
}

template <bool DESC, bool FMA>
int launch_fast(const ng_plan* p, const double* in, double* out, const FuseDev& F, cudaStream_t st) {
    const i64 big_tiles = (p->outer * p->inner / FN) * (p->n / FM);
    const int force = (int)NG_ENV_INT("NG_DENSE_NCP", 0);     // tuning: force the columns-per-thread variant
    if (p->inner == 1) {                       // contiguous contraction axis
        const bool wide_l = force ? force == 4 : (big_tiles >= 4 * 148 && p->outer % FN == 0);
        if (wide_l && p->outer % FN == 0) return launch_fast2<DESC, FMA, 4, true>(p, in, out, F, st);
        if (force == 2 && p->outer % 64 == 0) return launch_fast2<DESC, FMA, 2, true>(p, in, out, F, st);
        if (force != 1 && NG_ENV_INT("NG_DENSE_8X4", 1) != 0) return launch_fast2<DESC, FMA, 2, true, false, 128>(p, in, out, F, st);
        return launch_fast2<DESC, FMA, 1, true>(p, in, out, F, st);
    }
    const bool wide = force ? force == 4 : (big_tiles >= 4 * 148 && p->inner % FN == 0);
    if (wide && p->inner % FN == 0) return launch_fast2<DESC, FMA, 4, false>(p, in, out, F, st);
    if (force == 2 && p->inner % 64 == 0) return launch_fast2<DESC, FMA, 2, false>(p, in, out, F, st);
    // small grids (cfg4): 128 x 32 tiles with 8 x 4 outputs per thread, three CTAs of 128 threads per SM
    if (force != 1 && NG_ENV_INT("NG_DENSE_8X4", 1) != 0) return launch_fast2<DESC, FMA, 2, false, false, 128>(p, in, out, F, st);
    return launch_fast2<DESC, FMA, 1, false>(p, in, out, F, st);
}

// 0: fallback kernel; 1: exact shapes (n % 128 == 0, columns % 32 == 0); 2: general (GEN) kernel
int fast_mode(const ng_plan* p, const double* in, const double* out, const FuseDev& F) {
    if ((((uintptr_t)in | (uintptr_t)out | (uintptr_t)F.addend | (uintptr_t)F.minuend) & 15) != 0) return 0;   // 128-bit accesses
    if (F.any && p->inner != 1 && F.shape[F.vec_axis] % 2 != 0) return 0;
    const int force = (int)NG_ENV_INT("NG_DENSE_FAST", 1);    // tuning / A-B: 0 = fallback only, 2 = general kernel even on exact shapes
    if (force == 0) return 0;
    const bool exact = p->n % FM == 0 && (p->inner == 1 ? p->outer % 32 == 0 : p->inner % 32 == 0);
    if (exact && force != 2) return 1;
    // general: 128-bit accesses need whole pairs along the contiguous direction
    if (!p->d_Mtp || (p->inner == 1 ? p->n % 2 != 0 : p->inner % 2 != 0)) return 0;
    return 2;
}

template <bool DESC, bool FMA>
int launch_general(const ng_plan* p, const double* in, double* out, const FuseDev& F, cudaStream_t st) {
    if (p->inner == 1) return launch_fast2<DESC, FMA, 1, true, true>(p, in, out, F, st);
    return launch_fast2<DESC, FMA, 1, false, true>(p, in, out, F, st);
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
    NG_TRACE("dense_apply");
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
    const int mode = fast_mode(p, in, out, F);
    if (mode == 1) {
        if (p->use_fma) return launch_fast<false, true>(p, in, out, F, st);
        if (p->descending) return launch_fast<true, false>(p, in, out, F, st);
        return launch_fast<false, false>(p, in, out, F, st);
    }
    if (mode == 2) {
        if (p->use_fma) return launch_general<false, true>(p, in, out, F, st);
        if (p->descending) return launch_general<true, false>(p, in, out, F, st);
        return launch_general<false, false>(p, in, out, F, st);
    }
    return p->inner == 1 ? launch1<true>(p, in, out, F, st) : launch1<false>(p, in, out, F, st);
}
