// fdlap3d.cu -- 3-D production kernel of the fused Cartesian FD Laplacian (see fdlap.cu).
//
// Shape (nx, ny, nz), z contiguous; acc=4 second-derivative tables (central offsets -2..2,
// one-sided 6 taps) -- the reference default `Diff(g, 2, a)` (numgrids/api.py:31, acc=4).
//
// Mapping (HBM-bound, 16 B/point; co-limited by the FP64 pipe: 33 non-contractable DMUL/DADD per
// point because parity forbids FMA):
//   * a warp covers 64 contiguous z (one double2 = 128-bit LDG/STG per lane) x R rows of y;
//     a CTA stacks WY warps along y; grid = (z tiles, y tiles, x chunks);
//   * the CTA marches along x keeping the 5-plane x window of its own points in registers, so
//     every input point is fetched from HBM once (plus 4 planes of overlap per x chunk);
//   * y neighbours: R own rows + 4 halo rows (L1/L2 hits: they are the own rows of the
//     neighbouring warps); z neighbours: the adjacent lanes' double2 by warp shuffle, tile-edge
//     lanes load the 16 B next to the tile (or the slab halo);
//   * the first/last 2 indices of each axis use the one-sided schemes, recomputed by the owning
//     threads from direct loads in warp-uniform branches (rare).
// Arithmetic is the reference's, bit for bit: per axis acc = 0.0; acc = acc + w_k*f_k in stored
// tap order; d_a = acc * inv_h2_a; out = (d_x + d_y) + d_z.
#include <cuda.h>

#include "common.cuh"

namespace {

struct Lap3W {
    double wc[3][5], wf[3][6], wb[3][6], ih[3];
};

__device__ __forceinline__ double2 ld2(const double* p) {
    return *reinterpret_cast<const double2*>(p);
}
__device__ __forceinline__ double2 shfl_up2(double2 v) {
    return make_double2(__shfl_up_sync(0xffffffffu, v.x, 1), __shfl_up_sync(0xffffffffu, v.y, 1));
}
__device__ __forceinline__ double2 shfl_dn2(double2 v) {
    return make_double2(__shfl_down_sync(0xffffffffu, v.x, 1), __shfl_down_sync(0xffffffffu, v.y, 1));
}
__device__ __forceinline__ double mad_rn(double acc, double w, double f) {
    return __dadd_rn(acc, __dmul_rn(w, f));
}
// 5-tap central sum on two lanes of a double2 column
__device__ __forceinline__ double2 central5(const double* w, double2 a, double2 b, double2 c,
                                            double2 d, double2 e, double ih) {
    double sx = 0.0, sy = 0.0;
    sx = mad_rn(sx, w[0], a.x); sy = mad_rn(sy, w[0], a.y);
    sx = mad_rn(sx, w[1], b.x); sy = mad_rn(sy, w[1], b.y);
    sx = mad_rn(sx, w[2], c.x); sy = mad_rn(sy, w[2], c.y);
    sx = mad_rn(sx, w[3], d.x); sy = mad_rn(sy, w[3], d.y);
    sx = mad_rn(sx, w[4], e.x); sy = mad_rn(sy, w[4], e.y);
    return make_double2(__dmul_rn(sx, ih), __dmul_rn(sy, ih));
}
// same without the leading `0.0 +` (see ring2_compute for why this is still bit-exact after a final + 0.0)
__device__ __forceinline__ double2 central5_nz(const double* w, double2 a, double2 b, double2 c,
                                               double2 d, double2 e, double ih) {
    double sx = __dmul_rn(w[0], a.x), sy = __dmul_rn(w[0], a.y);
    sx = mad_rn(sx, w[1], b.x); sy = mad_rn(sy, w[1], b.y);
    sx = mad_rn(sx, w[2], c.x); sy = mad_rn(sy, w[2], c.y);
    sx = mad_rn(sx, w[3], d.x); sy = mad_rn(sy, w[3], d.y);
    sx = mad_rn(sx, w[4], e.x); sy = mad_rn(sy, w[4], e.y);
    return make_double2(__dmul_rn(sx, ih), __dmul_rn(sy, ih));
}
// 6-tap one-sided sum along a strided direction: taps at p[k*step], k = 0..5 (step may be negative)
__device__ __forceinline__ double2 onesided6(const double* w, const double* p, i64 step, double ih) {
    double sx = 0.0, sy = 0.0;
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        const double2 v = ld2(p + k * step);
        sx = mad_rn(sx, w[k], v.x);
        sy = mad_rn(sy, w[k], v.y);
    }
    return make_double2(__dmul_rn(sx, ih), __dmul_rn(sy, ih));
}
// one-sided along the contiguous axis for a single output: taps p[k*dir]
__device__ __forceinline__ double onesided6_z(const double* w, const double* p, int dir, double ih) {
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < 6; ++k) s = mad_rn(s, w[k], p[k * dir]);
    return __dmul_rn(s, ih);
}

template <int R, int WY>
__global__ void __launch_bounds__(32 * WY)
fdlap3d_march_kernel(const double* __restrict__ in, double* __restrict__ out, int nx, int ny, int nz,
                     int xchunk, int xbeg, int xend, Lap3W W, const double* __restrict__ lo,
                     const double* __restrict__ hi) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int z0 = blockIdx.x * 64;
    const int z = z0 + 2 * lane;
    const int y0 = (blockIdx.y * WY + warp) * R;
    const int xs = xbeg + blockIdx.z * xchunk;
    const int xe = min(xend, xs + xchunk);
    if (y0 >= ny) return;                       // warp-uniform
    const bool zok = z < nz;                    // nz is even: the pair is valid or not as a whole
    const i64 sy = nz, sx = (i64)ny * nz;
    const double2 zero2 = make_double2(0.0, 0.0);

    // which lane holds the last pair of the pencil (z = nz-2) in this tile, if any
    const bool tile_has_zlo = (z0 == 0);
    const bool tile_has_zhi = (z0 + 64 >= nz);
    const bool fix_zlo = tile_has_zlo && lo == nullptr;
    const bool fix_zhi = tile_has_zhi && hi == nullptr;

    // register window: w[j][r] = plane (x-2+j), row y0+r, this lane's pair
    double2 w[5][R];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int xp = xs - 2 + j;
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const bool ok = zok && xp >= 0 && xp < nx && (y0 + r) < ny;
            w[j + 1][r] = ok ? ld2(in + xp * sx + (i64)(y0 + r) * sy + z) : zero2;
        }
    }

    for (int x = xs; x < xe; ++x) {
        // ---- shift window, fetch plane x+2
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int r = 0; r < R; ++r) w[j][r] = w[j + 1][r];
        {
            const int xp = x + 2;
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const bool ok = zok && xp < nx && (y0 + r) < ny;
                w[4][r] = ok ? ld2(in + xp * sx + (i64)(y0 + r) * sy + z) : zero2;
            }
        }
        const double* plane = in + x * sx;
        // ---- y column of the centre plane: rows y0-2 .. y0+R+1
        double2 col[R + 4];
#pragma unroll
        for (int r = 0; r < R; ++r) col[r + 2] = w[2][r];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int ya = y0 - 2 + h, yb = y0 + R + h;
            col[h] = (zok && ya >= 0) ? ld2(plane + (i64)ya * sy + z) : zero2;
            col[R + 2 + h] = (zok && yb < ny) ? ld2(plane + (i64)yb * sy + z) : zero2;
        }
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int y = y0 + r;
            if (y >= ny) break;                 // warp-uniform
            const double* row = plane + (i64)y * sy;
            const double2 c = w[2][r];
            // ---- z neighbours: adjacent lanes, tile-edge lanes from memory / slab halo
            double2 L = shfl_up2(c), Rn = shfl_dn2(c);
            if (lane == 0) {
                if (z0 > 0) L = ld2(row + z0 - 2);
                else if (lo) L = ld2(lo + ((i64)x * ny + y) * 2);
                else L = zero2;
            }
            if (lane == 31) {
                if (z0 + 64 < nz) Rn = ld2(row + z0 + 64);
                else if (z0 + 64 == nz && hi) Rn = ld2(hi + ((i64)x * ny + y) * 2);
                else Rn = zero2;
            }
            if (tile_has_zhi && hi && z == nz - 2 && lane != 31)      // partial last tile in slab mode
                Rn = ld2(hi + ((i64)x * ny + y) * 2);
            // ---- central sums
            double2 dx = central5(W.wc[0], w[0][r], w[1][r], c, w[3][r], w[4][r], W.ih[0]);
            double2 dy = central5(W.wc[1], col[r], col[r + 1], c, col[r + 3], col[r + 4], W.ih[1]);
            double2 dz;
            {
                double s0 = 0.0, s1 = 0.0;
                const double* wz = W.wc[2];
                s0 = mad_rn(s0, wz[0], L.x);  s1 = mad_rn(s1, wz[0], L.y);
                s0 = mad_rn(s0, wz[1], L.y);  s1 = mad_rn(s1, wz[1], c.x);
                s0 = mad_rn(s0, wz[2], c.x);  s1 = mad_rn(s1, wz[2], c.y);
                s0 = mad_rn(s0, wz[3], c.y);  s1 = mad_rn(s1, wz[3], Rn.x);
                s0 = mad_rn(s0, wz[4], Rn.x); s1 = mad_rn(s1, wz[4], Rn.y);
                dz = make_double2(__dmul_rn(s0, W.ih[2]), __dmul_rn(s1, W.ih[2]));
            }
            // ---- one-sided fix-ups (uniform per warp / per step)
            if (zok) {
                if (x < 2) dx = onesided6(W.wf[0], row + z, sx, W.ih[0]);
                else if (x >= nx - 2) dx = onesided6(W.wb[0], row + z, -sx, W.ih[0]);
                if (y < 2) dy = onesided6(W.wf[1], row + z, sy, W.ih[1]);
                else if (y >= ny - 2) dy = onesided6(W.wb[1], row + z, -sy, W.ih[1]);
                if (fix_zlo && z == 0) {
                    dz.x = onesided6_z(W.wf[2], row, 1, W.ih[2]);
                    dz.y = onesided6_z(W.wf[2], row + 1, 1, W.ih[2]);
                }
                if (fix_zhi && z == nz - 2) {
                    dz.x = onesided6_z(W.wb[2], row + nz - 2, -1, W.ih[2]);
                    dz.y = onesided6_z(W.wb[2], row + nz - 1, -1, W.ih[2]);
                }
                double2 o;
                o.x = __dadd_rn(__dadd_rn(dx.x, dy.x), dz.x);
                o.y = __dadd_rn(__dadd_rn(dx.y, dy.y), dz.y);
                *reinterpret_cast<double2*>(out + x * sx + (i64)y * sy + z) = o;
            }
        }
    }
}

// -------------------------------------------------------------------------------------------------
// v2 fast path: full tiles only (nz % 64 == 0, ny % R == 0).  No bounds checks in the steady state:
// halo rows / tile-edge columns are CLAMPED to valid addresses (their garbage only reaches outputs
// that the one-sided fix-ups overwrite), the 6-plane register window (x-2 .. x+3, the last one
// prefetched a step ahead) rotates by renaming through a 6x unrolled march, and every row is
// finished (dx, dy, dz, combine, store) before the next starts to keep live registers low.
// -------------------------------------------------------------------------------------------------
template <int R>
struct Rows {
    double2 v[R];
};

struct Lap3C {            // central weights + scales: kernel parameter (constant bank operands)
    double wc[3][5], ih[3];
};
struct Lap3Side {         // one-sided tables: read from global memory on the rare fix-up path
    double wf[3][6], wb[3][6], ih[3];
};

// One-sided replacement of the axis sums that sit on a global boundary, then the final combine.
// Deliberately NOT inlined: it is executed by <1% of the warp-steps and must not bloat the march.
__device__ __noinline__ double2 fix_and_combine(const Lap3Side* __restrict__ S, const double* row,
                                                i64 sx, int sy, int x, int y, int nx, int ny,
                                                int zfix /*0 none, 1 first pair, 2 last pair*/,
                                                double2 dx, double2 dy, double2 dz) {
    if (x < 2) dx = onesided6(S->wf[0], row, sx, S->ih[0]);
    else if (x >= nx - 2) dx = onesided6(S->wb[0], row, -sx, S->ih[0]);
    if (y < 2) dy = onesided6(S->wf[1], row, sy, S->ih[1]);
    else if (y >= ny - 2) dy = onesided6(S->wb[1], row, -(i64)sy, S->ih[1]);
    if (zfix == 1) {
        dz.x = onesided6_z(S->wf[2], row, 1, S->ih[2]);
        dz.y = onesided6_z(S->wf[2], row + 1, 1, S->ih[2]);
    } else if (zfix == 2) {
        dz.x = onesided6_z(S->wb[2], row, -1, S->ih[2]);
        dz.y = onesided6_z(S->wb[2], row + 1, -1, S->ih[2]);
    }
    return make_double2(__dadd_rn(__dadd_rn(dx.x, dy.x), dz.x), __dadd_rn(__dadd_rn(dx.y, dy.y), dz.y));
}

template <int R>
struct MarchCtx {
    const double* pc;     // centre plane x, at (row y0, this lane's z)
    const double* pn;     // plane min(x+3, nx-1), same row/column
    const double* pe;     // tile-edge column source for this lane (lane 0 / 31), plane x, row y0
    double* po;           // output plane x, row y0
    const Lap3Side* side;
    i64 sx, esx;          // plane strides of pc/po/pn and of pe
    int esy, sy;          // row strides of pe and of pc
    int oh[4];            // clamped halo row offsets relative to row y0
    int y0, ny, nx, lane, zfix;
    bool is_edge_lane, rowfix;
};

template <int R>
__device__ __forceinline__ void march_step(int x, const MarchCtx<R>& c, const Lap3C& W,
                                           const Rows<R>& A, const Rows<R>& B, const Rows<R>& C,
                                           const Rows<R>& D, const Rows<R>& E, Rows<R>& F) {
    // ---- loads: prefetch plane x+3, halo rows of plane x, tile-edge columns of plane x
#pragma unroll
    for (int r = 0; r < R; ++r) F.v[r] = ld2(c.pn + r * c.sy);
    double2 col[R + 4];
    col[0] = ld2(c.pc + c.oh[0]);
    col[1] = ld2(c.pc + c.oh[1]);
    col[R + 2] = ld2(c.pc + c.oh[2]);
    col[R + 3] = ld2(c.pc + c.oh[3]);
    double2 ed[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        ed[r] = make_double2(0.0, 0.0);
        if (c.is_edge_lane) ed[r] = ld2(c.pe + r * c.esy);
    }
#pragma unroll
    for (int r = 0; r < R; ++r) col[r + 2] = C.v[r];
    const bool fix = c.rowfix || (x < 2) || (x >= c.nx - 2);     // warp-uniform
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const double2 cc = C.v[r];
        double2 L = shfl_up2(cc), Rn = shfl_dn2(cc);
        if (c.lane == 0) L = ed[r];
        if (c.lane == 31) Rn = ed[r];
        const double2 dx = central5(W.wc[0], A.v[r], B.v[r], cc, D.v[r], E.v[r], W.ih[0]);
        const double2 dy = central5(W.wc[1], col[r], col[r + 1], cc, col[r + 3], col[r + 4], W.ih[1]);
        double2 dz;
        {
            double s0 = 0.0, s1 = 0.0;
            const double* wz = W.wc[2];
            s0 = mad_rn(s0, wz[0], L.x);  s1 = mad_rn(s1, wz[0], L.y);
            s0 = mad_rn(s0, wz[1], L.y);  s1 = mad_rn(s1, wz[1], cc.x);
            s0 = mad_rn(s0, wz[2], cc.x); s1 = mad_rn(s1, wz[2], cc.y);
            s0 = mad_rn(s0, wz[3], cc.y); s1 = mad_rn(s1, wz[3], Rn.x);
            s0 = mad_rn(s0, wz[4], Rn.x); s1 = mad_rn(s1, wz[4], Rn.y);
            dz = make_double2(__dmul_rn(s0, W.ih[2]), __dmul_rn(s1, W.ih[2]));
        }
        double2 o;
        if (fix) {
            o = fix_and_combine(c.side, c.pc + r * c.sy, c.sx, c.sy, x, c.y0 + r, c.nx, c.ny, c.zfix,
                                dx, dy, dz);
        } else {
            o.x = __dadd_rn(__dadd_rn(dx.x, dy.x), dz.x);
            o.y = __dadd_rn(__dadd_rn(dx.y, dy.y), dz.y);
        }
        *reinterpret_cast<double2*>(c.po + r * c.sy) = o;
    }
}

template <int R, int WY, int MINB>
__global__ void __launch_bounds__(32 * WY, MINB)
fdlap3d_fast_kernel(const double* __restrict__ in, double* __restrict__ out, int nx, int ny, int nz,
                    int xchunk, int xbeg, int xend, Lap3C W, const Lap3Side* __restrict__ side,
                    const double* __restrict__ lo, const double* __restrict__ hi) {
    MarchCtx<R> c;
    c.lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int z0 = blockIdx.x * 64;
    const int z = z0 + 2 * c.lane;
    c.y0 = (blockIdx.y * WY + warp) * R;
    if (c.y0 >= ny) return;                       // warp-uniform; no block-level sync is used
    const int xs = xbeg + blockIdx.z * xchunk;
    const int xe = min(xend, xs + xchunk);
    c.nx = nx; c.ny = ny;
    c.sy = nz;
    c.sx = (i64)ny * nz;
    c.side = side;
    c.oh[0] = (max(c.y0 - 2, 0) - c.y0) * nz;
    c.oh[1] = (max(c.y0 - 1, 0) - c.y0) * nz;
    c.oh[2] = (min(c.y0 + R, ny - 1) - c.y0) * nz;
    c.oh[3] = (min(c.y0 + R + 1, ny - 1) - c.y0) * nz;
    c.rowfix = (c.y0 < 2) || (c.y0 + R > ny - 2) || ((z0 == 0) && (lo == nullptr)) ||
               ((z0 + 64 == nz) && (hi == nullptr));
    c.zfix = 0;
    if (z0 == 0 && lo == nullptr && c.lane == 0) c.zfix = 1;
    if (z0 + 64 == nz && hi == nullptr && c.lane == 31) c.zfix = 2;
    c.is_edge_lane = (c.lane == 0) || (c.lane == 31);
    const i64 rowbase = (i64)c.y0 * nz + z;
    // tile-edge column source: neighbour tile in the same row, the slab halo, or (global boundary)
    // this lane's own pair as a harmless stand-in
    c.pe = in + xs * c.sx + rowbase;
    c.esx = c.sx;
    c.esy = nz;
    if (c.lane == 0) {
        if (z0 > 0) c.pe -= 2;
        else if (lo) { c.pe = lo + ((i64)xs * ny + c.y0) * 2; c.esx = (i64)ny * 2; c.esy = 2; }
    } else if (c.lane == 31) {
        if (z0 + 64 < nz) c.pe += 2;
        else if (hi) { c.pe = hi + ((i64)xs * ny + c.y0) * 2; c.esx = (i64)ny * 2; c.esy = 2; }
    }
    c.pc = in + xs * c.sx + rowbase;
    c.po = out + xs * c.sx + rowbase;
    c.pn = in + (i64)min(xs + 3, nx - 1) * c.sx + rowbase;

    Rows<R> w0, w1, w2, w3, w4, w5;
    {   // planes xs-2 .. xs+2 (clamped into the array)
        const double* p;
        p = in + (i64)max(xs - 2, 0) * c.sx + rowbase;
#pragma unroll
        for (int r = 0; r < R; ++r) w0.v[r] = ld2(p + r * c.sy);
        p = in + (i64)max(xs - 1, 0) * c.sx + rowbase;
#pragma unroll
        for (int r = 0; r < R; ++r) w1.v[r] = ld2(p + r * c.sy);
        p = c.pc;
#pragma unroll
        for (int r = 0; r < R; ++r) w2.v[r] = ld2(p + r * c.sy);
        p = in + (i64)min(xs + 1, nx - 1) * c.sx + rowbase;
#pragma unroll
        for (int r = 0; r < R; ++r) w3.v[r] = ld2(p + r * c.sy);
        p = in + (i64)min(xs + 2, nx - 1) * c.sx + rowbase;
#pragma unroll
        for (int r = 0; r < R; ++r) w4.v[r] = ld2(p + r * c.sy);
    }
    int x = xs;
#define NG_STEP(A, B, C, D, E, F)                       \
    march_step<R>(x, c, W, A, B, C, D, E, F);           \
    ++x;                                                \
    if (x >= xe) break;                                 \
    c.pc += c.sx; c.po += c.sx; c.pe += c.esx;          \
    if (x + 3 <= nx - 1) c.pn += c.sx;
    for (;;) {
        NG_STEP(w0, w1, w2, w3, w4, w5)
        NG_STEP(w1, w2, w3, w4, w5, w0)
        NG_STEP(w2, w3, w4, w5, w0, w1)
        NG_STEP(w3, w4, w5, w0, w1, w2)
        NG_STEP(w4, w5, w0, w1, w2, w3)
        NG_STEP(w5, w0, w1, w2, w3, w4)
    }
#undef NG_STEP
}

// -------------------------------------------------------------------------------------------------
// v3: shared-memory plane ring fed by the bulk-copy engine (cp.async.bulk + mbarrier, SASS UBLKCP).
//
// A CTA owns a (TY = R*WY rows) x (64 z) tile and marches along x.  One producer warp streams
// whole planes of the tile WITH their +-2 halo -- (TY+4) rows x 68 doubles, one 544-byte bulk
// copy per row and lane -- into a ring of S shared-memory slots, `S-3` planes ahead of the
// consumers, completing on a per-slot "full" mbarrier (transaction bytes).  WY consumer warps keep
// the 5-plane x window of their own points in registers (one LDS.128 per row and step) and read
// the y-halo rows and z neighbours of the centre plane from the ring, so the steady state has no
// global loads, no shuffles and no bounds checks in the consumer instruction stream; a slot is
// handed back through its "empty" mbarrier two steps after it was filled.  HBM sees every input
// point once (+ tile halos, which are L2 hits between neighbouring CTAs) and every output once.
// One-sided boundary schemes: z and y from the ring (all 7 points are in the tile), x from global
// memory; each only in warp-uniform branches on the boundary tiles / steps.
// -------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) {
    return (unsigned)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
// Every wait is bounded: a protocol bug (e.g. a wrong transaction byte count) traps after ~2^24
// polls instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    const unsigned addr = smem_u32(bar);
    unsigned done, spins = 0;
    do {
        if (++spins > (1u << 24)) __trap();
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(addr), "r"(parity) : "memory");
    } while (!done);
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, unsigned bytes,
                                         unsigned long long* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
        ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

constexpr int RING_TZ = 64;            // z points per tile (one double2 per lane)
constexpr int RING_PITCH = RING_TZ + 4;   // doubles per staged row (2 halo columns each side)

// Boundary fix-up + combine for one row pair (rare; kept out of line).  `cen` points at this
// lane's pair of the centre plane inside the ring (row pitch RING_PITCH); `grow` at the same pair
// in global memory (for the x direction, whose one-sided taps are other planes).
__device__ __noinline__ double2 ring_fix_and_combine(const Lap3Side* __restrict__ S, const double* cen,
                                                     const double* grow, i64 sx, int x, int y, int nx,
                                                     int ny, int zfix, double2 dx, double2 dy,
                                                     double2 dz) {
    if (x < 2) dx = onesided6(S->wf[0], grow, sx, S->ih[0]);
    else if (x >= nx - 2) dx = onesided6(S->wb[0], grow, -sx, S->ih[0]);
    if (y < 2) dy = onesided6(S->wf[1], cen, RING_PITCH, S->ih[1]);
    else if (y >= ny - 2) dy = onesided6(S->wb[1], cen, -RING_PITCH, S->ih[1]);
    if (zfix == 1) {
        dz.x = onesided6_z(S->wf[2], cen, 1, S->ih[2]);
        dz.y = onesided6_z(S->wf[2], cen + 1, 1, S->ih[2]);
    } else if (zfix == 2) {
        dz.x = onesided6_z(S->wb[2], cen, -1, S->ih[2]);
        dz.y = onesided6_z(S->wb[2], cen + 1, -1, S->ih[2]);
    }
    return make_double2(__dadd_rn(__dadd_rn(dx.x, dy.x), dz.x), __dadd_rn(__dadd_rn(dx.y, dy.y), dz.y));
}

template <int R>
__device__ __forceinline__ void ring_compute(int x, int y0, int z, int nx, int ny, int nz, i64 sx,
                                             const Lap3C& W, const Lap3Side* side,
                                             const double* __restrict__ in, double* __restrict__ out,
                                             const double* lo_pair, const double* hi_pair,
                                             const double* cen, bool rowfix, int zfix,
                                             const Rows<R>& A, const Rows<R>& B, const Rows<R>& C,
                                             const Rows<R>& D, const Rows<R>& E) {
    double2 colv[R + 4];
    colv[0] = ld2(cen - 2 * RING_PITCH);
    colv[1] = ld2(cen - RING_PITCH);
    colv[R + 2] = ld2(cen + R * RING_PITCH);
    colv[R + 3] = ld2(cen + (R + 1) * RING_PITCH);
#pragma unroll
    for (int r = 0; r < R; ++r) colv[r + 2] = C.v[r];
    const bool fix = rowfix || (x < 2) || (x >= nx - 2);        // warp-uniform
    const i64 gbase = x * sx + (i64)y0 * nz + z;
#pragma unroll
    for (int r = 0; r < R; ++r) {
        if (y0 + r >= ny) break;                                // ragged last tile (warp-uniform)
        const double2 cc = C.v[r];
        const double* crow = cen + r * RING_PITCH;
        double2 L = ld2(crow - 2), Rn = ld2(crow + 2);
        if (lo_pair) L = ld2(lo_pair + ((i64)x * ny + y0 + r) * 2);      // slab halo (lane 0 only)
        if (hi_pair) Rn = ld2(hi_pair + ((i64)x * ny + y0 + r) * 2);     // slab halo (lane 31 only)
        const double2 dx = central5(W.wc[0], A.v[r], B.v[r], cc, D.v[r], E.v[r], W.ih[0]);
        const double2 dy = central5(W.wc[1], colv[r], colv[r + 1], cc, colv[r + 3], colv[r + 4], W.ih[1]);
        double2 dz;
        {
            double s0 = 0.0, s1 = 0.0;
            const double* wz = W.wc[2];
            s0 = mad_rn(s0, wz[0], L.x);  s1 = mad_rn(s1, wz[0], L.y);
            s0 = mad_rn(s0, wz[1], L.y);  s1 = mad_rn(s1, wz[1], cc.x);
            s0 = mad_rn(s0, wz[2], cc.x); s1 = mad_rn(s1, wz[2], cc.y);
            s0 = mad_rn(s0, wz[3], cc.y); s1 = mad_rn(s1, wz[3], Rn.x);
            s0 = mad_rn(s0, wz[4], Rn.x); s1 = mad_rn(s1, wz[4], Rn.y);
            dz = make_double2(__dmul_rn(s0, W.ih[2]), __dmul_rn(s1, W.ih[2]));
        }
        double2 o;
        if (fix) {
            o = ring_fix_and_combine(side, crow, in + gbase + (i64)r * nz, sx, x, y0 + r, nx, ny, zfix,
                                     dx, dy, dz);
        } else {
            o.x = __dadd_rn(__dadd_rn(dx.x, dy.x), dz.x);
            o.y = __dadd_rn(__dadd_rn(dx.y, dy.y), dz.y);
        }
        *reinterpret_cast<double2*>(out + gbase + (i64)r * nz) = o;
    }
}

template <int R, int WY, int S, int MINB>
__global__ void __launch_bounds__(32 * (WY + 1), MINB)
fdlap3d_ring_kernel(const double* __restrict__ in, double* __restrict__ out, int nx, int ny, int nz,
                    int xchunk, int xbeg, int xend, Lap3C W, const Lap3Side* __restrict__ side,
                    const double* __restrict__ lo, const double* __restrict__ hi) {
    constexpr int TY = R * WY;
    constexpr int ROWS = TY + 4;
    constexpr int PLANE = ROWS * RING_PITCH;          // doubles per ring slot
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* ring = reinterpret_cast<double*>(smem_raw);
    unsigned long long* full = reinterpret_cast<unsigned long long*>(ring + (size_t)S * PLANE);
    unsigned long long* empty = full + S;

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int z0 = blockIdx.x * RING_TZ;
    const int yb = blockIdx.y * TY;                   // first row of the tile
    const int xs = xbeg + blockIdx.z * xchunk;
    const int xe = min(xend, xs + xchunk);
    const i64 sx = (i64)ny * nz;
    const int p_first = xs - 2, p_last = xe + 1;      // planes that travel through the ring
    const int nplanes = p_last - p_first + 1;

    if (threadIdx.x == 0) {
        for (int s = 0; s < S; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], WY); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp == WY) {
        // ================= producer warp: one bulk copy per tile row and plane =================
        const int zlo = max(z0 - 2, 0), zhi = min(z0 + RING_TZ + 2, nz);     // [zlo, zhi) of each row
        const unsigned row_bytes = (unsigned)(zhi - zlo) * 8u;
        const int dst_col = zlo - (z0 - 2);
        const int ylo = max(yb - 2, 0), yhi = min(yb + TY + 2, ny);          // rows [ylo, yhi)
        const unsigned plane_bytes = row_bytes * (unsigned)(yhi - ylo);
        for (int q = 0; q < nplanes; ++q) {
            const int s = q % S;
            if (q >= S) mbar_wait(&empty[s], ((q / S) - 1) & 1);
            const int p = p_first + q;
            if (p < 0 || p >= nx) {                    // nothing to stage: just complete the phase
                if (lane == 0) mbar_arrive(&full[s]);
                continue;
            }
            if (lane == 0) mbar_arrive_expect_tx(&full[s], plane_bytes);
            __syncwarp();
            double* slot = ring + (size_t)s * PLANE;
            for (int y = ylo + lane; y < yhi; y += 32)
                bulk_g2s(slot + (y - (yb - 2)) * RING_PITCH + dst_col, in + p * sx + (i64)y * nz + zlo,
                         row_bytes, &full[s]);
        }
        return;
    }

    // ===================== consumer warps =====================
    const int j0 = warp * R;                           // first tile-local row of this warp
    const int y0 = yb + j0;
    if (y0 >= ny) {
        // warp entirely outside the grid (ragged last tile): still hand the slots back
        for (int q = 2; q < nplanes; ++q) {
            mbar_wait(&full[q % S], (q / S) & 1);
            if (lane == 0) mbar_arrive(&empty[(q - 2) % S]);
        }
        return;
    }
    const int z = z0 + 2 * lane;
    const int col = 2 + 2 * lane;                      // this lane's pair inside a staged row
    const bool zfix_lo = (z0 == 0) && (lo == nullptr), zfix_hi = (z0 + RING_TZ == nz) && (hi == nullptr);
    const bool rowfix = (y0 < 2) || (y0 + R > ny - 2) || zfix_lo || zfix_hi;      // warp-uniform
    const int zfix = (zfix_lo && lane == 0) ? 1 : ((zfix_hi && lane == 31) ? 2 : 0);
    const double* lo_pair = ((z0 == 0) && lane == 0) ? lo : nullptr;
    const double* hi_pair = ((z0 + RING_TZ == nz) && lane == 31) ? hi : nullptr;

    Rows<R> w0, w1, w2, w3, w4;
    int q = 0;
    // one ring plane per step: `NEW` receives plane p = p_first + q; the centre is plane p-2 in (A..E)
#define NG_RING_STEP(A, B, C, D, NEW)                                                            \
    {                                                                                            \
        const int s = q % S;                                                                     \
        mbar_wait(&full[s], (q / S) & 1);                                                        \
        const double* slot = ring + (size_t)s * PLANE + (j0 + 2) * RING_PITCH + col;             \
        _Pragma("unroll") for (int r = 0; r < R; ++r) NEW.v[r] = ld2(slot + r * RING_PITCH);     \
        const int x = p_first + q - 2;                                                           \
        if (x >= xs) {                                                                           \
            const double* cen = ring + (size_t)((q - 2) % S) * PLANE + (j0 + 2) * RING_PITCH + col; \
            ring_compute<R>(x, y0, z, nx, ny, nz, sx, W, side, in, out, lo_pair, hi_pair, cen,   \
                            rowfix, zfix, A, B, C, D, NEW);                                      \
        }                                                                                        \
        if (q >= 2) {                                                                            \
            __syncwarp();                                                                        \
            if (lane == 0) mbar_arrive(&empty[(q - 2) % S]);                                     \
        }                                                                                        \
        ++q;                                                                                     \
        if (q >= nplanes) break;                                                                 \
    }
    for (;;) {
        NG_RING_STEP(w1, w2, w3, w4, w0)
        NG_RING_STEP(w2, w3, w4, w0, w1)
        NG_RING_STEP(w3, w4, w0, w1, w2)
        NG_RING_STEP(w4, w0, w1, w2, w3)
        NG_RING_STEP(w0, w1, w2, w3, w4)
    }
#undef NG_RING_STEP
}

// -------------------------------------------------------------------------------------------------
// v4: same plane ring as v3 with a lean consumer instruction stream:
//   * one central weight set shared by the three axes (they are the same np.linalg.solve output
//     when every axis uses the same (order, acc); checked on the host) -> 8 uniform operands;
//   * 32-bit shared addresses advanced incrementally (no div/mod, no generic->shared conversion),
//     LDS.128 through inline PTX; running global output offset;
//   * the 4 prologue steps (window fill) are peeled, so the steady-state step has no prologue tests;
//   * slab-halo handling compiled in only for the SLAB instantiation.
// -------------------------------------------------------------------------------------------------
struct Lap3U {
    double w[5], ih[3];
    double wfz[6], wbz[6];     // one-sided tables of the contiguous axis (global z boundary, inline fix)
};

__device__ __forceinline__ double lds1(unsigned addr) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
    return v;
}
// one-sided 6-tap sum along z read from the ring: taps at addr + k*step bytes
__device__ __forceinline__ double onesided6_z_smem(const double* w, unsigned addr, int step, double ih) {
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < 6; ++k) s = mad_rn(s, w[k], lds1(addr + k * step));
    return __dmul_rn(s, ih);
}

__device__ __forceinline__ double2 lds2(unsigned addr) {
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr));
    return v;
}
__device__ __forceinline__ void mbar_wait_a(unsigned addr, unsigned parity) {
    unsigned done, spins = 0;
    do {
        if (++spins > (1u << 24)) __trap();
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(addr), "r"(parity) : "memory");
    } while (!done);
}
__device__ __forceinline__ void mbar_arrive_a(unsigned addr) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(addr) : "memory");
}

template <int R, bool SLAB>
__device__ __forceinline__ void ring2_compute(int x, int y0, int nx, int ny, int nz, i64 sx, i64 goff,
                                              const Lap3U& W, const Lap3Side* side,
                                              const double* __restrict__ in, double* __restrict__ out,
                                              unsigned a_hlo, unsigned a_hhi,
                                              unsigned a_cen, bool rowfix, int ztile, int zfix,
                                              const Rows<R>& A, const Rows<R>& B, const Rows<R>& C,
                                              const Rows<R>& D, const Rows<R>& E) {
    constexpr unsigned PB = RING_PITCH * 8;
    double2 colv[R + 4];
    colv[0] = lds2(a_cen - 2 * PB);
    colv[1] = lds2(a_cen - PB);
    colv[R + 2] = lds2(a_cen + R * PB);
    colv[R + 3] = lds2(a_cen + (R + 1) * PB);
#pragma unroll
    for (int r = 0; r < R; ++r) colv[r + 2] = C.v[r];
    const bool fix = rowfix || (x < 2) || (x >= nx - 2);        // warp-uniform
    // Global z boundary (first / last z tile without a slab halo): the boundary lane needs 2R
    // one-sided 6-tap sums (2 z points x R rows).  Each sum is sequential (reference order), but
    // the 2R sums are independent: lanes 0..2R-1 each compute ONE of them from the ring and the
    // boundary lane collects them by shuffle -- one sum's latency per step instead of 2R (the
    // single-lane version cost +35 % FP64 issue on every boundary tile, i.e. on half of the
    // tiles of a 256-wide slab).  ztile: 1 = first tile, 2 = last tile, 3 = both (nz == 64).
    const int lane_ = threadIdx.x & 31;
    double zsum = 0.0;
    if (ztile == 1 || ztile == 2) {                             // warp-uniform
        if (lane_ < 2 * R) {
            const int rr = lane_ >> 1, cc_ = lane_ & 1;
            const unsigned edge = ztile == 1 ? a_cen - 16u * lane_ : a_cen + 16u * (31 - lane_);   // the boundary lane's pair
            zsum = onesided6_z_smem(ztile == 1 ? W.wfz : W.wbz, edge + rr * PB + cc_ * 8, ztile == 1 ? 8 : -8, W.ih[2]);
        }
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
        if (y0 + r >= ny) break;                                // ragged last tile (warp-uniform)
        const double2 cc = C.v[r];
        double2 L = lds2(a_cen + r * PB - 16), Rn = lds2(a_cen + r * PB + 16);
        if (SLAB) {     // neighbour rank's planes, staged next to the ring by the producer's halo TMAs
            if (a_hlo) L = lds2(a_hlo + r * 16);        // lane 0 of the first z tile
            if (a_hhi) Rn = lds2(a_hhi + r * 16);       // lane 31 of the last z tile
        }
        // The reference starts every tap sum from +0.0.  `0.0 + p` only differs from `p` when p is
        // -0.0, and a sum that starts from its first product can end as -0.0 only if EVERY product
        // was -0.0; that sign survives to the output only if all three axis sums are -0.0, which the
        // single `+ 0.0` after the combine turns back into the reference's +0.0.  Every other value
        // is bit-identical, so the three leading adds per point are dropped (36 -> 34 FP64 ops/point).
        const double2 dx = central5_nz(W.w, A.v[r], B.v[r], cc, D.v[r], E.v[r], W.ih[0]);
        const double2 dy = central5_nz(W.w, colv[r], colv[r + 1], cc, colv[r + 3], colv[r + 4], W.ih[1]);
        double2 dz;
        {
            double s0 = __dmul_rn(W.w[0], L.x), s1 = __dmul_rn(W.w[0], L.y);
            s0 = mad_rn(s0, W.w[1], L.y);  s1 = mad_rn(s1, W.w[1], cc.x);
            s0 = mad_rn(s0, W.w[2], cc.x); s1 = mad_rn(s1, W.w[2], cc.y);
            s0 = mad_rn(s0, W.w[3], cc.y); s1 = mad_rn(s1, W.w[3], Rn.x);
            s0 = mad_rn(s0, W.w[4], Rn.x); s1 = mad_rn(s1, W.w[4], Rn.y);
            dz = make_double2(__dmul_rn(s0, W.ih[2]), __dmul_rn(s1, W.ih[2]));
        }
        // global z boundary (first / last z tile without a slab halo): the owning lane replaces its
        // two z sums by the one-sided schemes, all 7 points come from the ring (warp-uniform test)
        if (ztile == 3) {
            if (zfix == 1) {
                dz.x = onesided6_z_smem(W.wfz, a_cen + r * PB, 8, W.ih[2]);
                dz.y = onesided6_z_smem(W.wfz, a_cen + r * PB + 8, 8, W.ih[2]);
            } else if (zfix == 2) {
                dz.x = onesided6_z_smem(W.wbz, a_cen + r * PB, -8, W.ih[2]);
                dz.y = onesided6_z_smem(W.wbz, a_cen + r * PB + 8, -8, W.ih[2]);
            }
        } else if (ztile) {
            const double v0 = __shfl_sync(0xffffffffu, zsum, 2 * r), v1 = __shfl_sync(0xffffffffu, zsum, 2 * r + 1);
            if (zfix) { dz.x = v0; dz.y = v1; }
        }
        double2 o;
        if (fix) {
            const double* crow = (const double*)__cvta_shared_to_generic((size_t)(a_cen + r * PB));
            o = ring_fix_and_combine(side, crow, in + goff + (i64)r * nz, sx, x, y0 + r, nx, ny, 0,
                                     dx, dy, dz);
        } else {
            o.x = __dadd_rn(__dadd_rn(dx.x, dy.x), dz.x);
            o.y = __dadd_rn(__dadd_rn(dx.y, dy.y), dz.y);
        }
        o.x = __dadd_rn(o.x, 0.0);
        o.y = __dadd_rn(o.y, 0.0);
        *reinterpret_cast<double2*>(out + goff + (i64)r * nz) = o;
    }
}

__device__ __forceinline__ void tma_load_3d(void* dst_smem, const CUtensorMap* tmap, int c0, int c1, int c2,
                                            unsigned long long* bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes "
        "[%0], [%1, {%2, %3, %4}], [%5];"
        ::"r"(smem_u32(dst_smem)), "l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void tma_load_2d(void* dst_smem, const CUtensorMap* tmap, int c0, int c1,
                                            unsigned long long* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes "
        "[%0], [%1, {%2, %3}], [%4];"
        ::"r"(smem_u32(dst_smem)), "l"(tmap), "r"(c0), "r"(c1), "r"(smem_u32(bar)) : "memory");
}

// TMAP: the producer issues ONE 3-D tensor-map TMA per plane (box 68 x (TY+4) x 1, out-of-bounds
// elements zero-filled by the hardware) instead of TY+4 row copies.
template <int R, int WY, int S, int MINB, bool SLAB, bool TMAP>
__global__ void __launch_bounds__(32 * (WY + 1), MINB)
fdlap3d_ring2_kernel(const double* __restrict__ in, double* __restrict__ out, int nx, int ny, int nz,
                     int xchunk, int xbeg, int xend, int zmode, Lap3U W, const Lap3Side* __restrict__ side,
                     const double* __restrict__ lo, const double* __restrict__ hi,
                     const __grid_constant__ CUtensorMap tmap, const __grid_constant__ CUtensorMap tmap_lo,
                     const __grid_constant__ CUtensorMap tmap_hi) {
    static_assert(!SLAB || TMAP, "slab halos are staged by TMA");
    constexpr int TY = R * WY;
    constexpr int ROWS = TY + 4;
    constexpr int PLANE = ROWS * RING_PITCH;          // doubles per ring slot
    constexpr unsigned PLANE_B = PLANE * 8, RING_B = S * PLANE_B, PB = RING_PITCH * 8;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    // TMA destinations must be 128-byte aligned (the launch reserves 128 spare bytes)
    double* ring = reinterpret_cast<double*>(smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u));
    // slab mode: per slot, the nb=2 neighbour columns of the tile's rows ([ROWS][2] doubles) per side
    constexpr int HALO = SLAB ? ((ROWS * 2 + 15) / 16) * 16 : 0;   // doubles per slot and side (128-B slots)
    constexpr unsigned HALO_B = HALO * 8;
    double* hlo = ring + (size_t)S * PLANE;
    double* hhi = hlo + (size_t)S * HALO;
    unsigned long long* full = reinterpret_cast<unsigned long long*>(hhi + (size_t)S * HALO);
    unsigned long long* empty = full + S;

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // zmode 0: every z tile; 1: interior tiles only (no halo dependence); 2: the first and last tile
    // (bits 4, 5: A/B switches of tools/ab_fdlap.py -- single-lane z sums, 3-D halo boxes)
    const bool ab_zsingle = (zmode & 16) != 0, ab_halo3d = (zmode & 32) != 0;
    zmode &= 3;
    const int zt_idx = zmode == 1 ? (int)blockIdx.x + 1
                                  : (zmode == 2 ? (blockIdx.x == 0 ? 0 : nz / RING_TZ - 1) : (int)blockIdx.x);
    const int z0 = zt_idx * RING_TZ;
    const int yb = blockIdx.y * TY;
    const int xs = xbeg + blockIdx.z * xchunk;
    const bool lo_tile = SLAB && (z0 == 0) && (lo != nullptr);
    const bool hi_tile = SLAB && (z0 + RING_TZ == nz) && (hi != nullptr);
    const int xe = min(xend, xs + xchunk);
    const i64 sx = (i64)ny * nz;
    const int p_first = xs - 2;
    const int nplanes = xe - xs + 4;                  // planes xs-2 .. xe+1 travel through the ring

    if (threadIdx.x == 0) {
        for (int s = 0; s < S; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], WY); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp == WY) {
        // ================= producer warp: one bulk copy per tile row and plane =================
        const int zlo = max(z0 - 2, 0), zhi = min(z0 + RING_TZ + 2, nz);
        const unsigned row_bytes = (unsigned)(zhi - zlo) * 8u;
        const int dst_col = zlo - (z0 - 2);
        const int ylo = max(yb - 2, 0), yhi = min(yb + TY + 2, ny);
        const unsigned plane_bytes = row_bytes * (unsigned)(yhi - ylo);
        int s = 0;
        unsigned eph = 1;                              // parity of the `empty` phase to wait for next
        if (TMAP) {
            if (lane == 0) {
                for (int q = 0; q < nplanes; ++q) {
                    if (q >= S) mbar_wait(&empty[s], eph ^ 1);
                    mbar_arrive_expect_tx(&full[s], PLANE_B + (lo_tile ? ROWS * 16u : 0u) + (hi_tile ? ROWS * 16u : 0u));   // bytes the TMAs deliver (slots are padded to 128 B)
                    tma_load_3d(ring + (size_t)s * PLANE, &tmap, z0 - 2, yb - 2, p_first + q, &full[s]);
                    // halo planes are [nx][ny][2]: the tile's ROWS x 2 values of one plane are ONE contiguous
                    // 16*ROWS-byte run, fetched as a single-row 2-D box (a {2, ROWS, 1} box would cost ROWS
                    // 16-byte row requests -- as many TMA row operations as the main box)
                    if (ab_halo3d) {
                        if (lo_tile) tma_load_3d(hlo + (size_t)s * HALO, &tmap_lo, 0, yb - 2, p_first + q, &full[s]);
                        if (hi_tile) tma_load_3d(hhi + (size_t)s * HALO, &tmap_hi, 0, yb - 2, p_first + q, &full[s]);
                    } else {
                        if (lo_tile) tma_load_2d(hlo + (size_t)s * HALO, &tmap_lo, 2 * (yb - 2), p_first + q, &full[s]);
                        if (hi_tile) tma_load_2d(hhi + (size_t)s * HALO, &tmap_hi, 2 * (yb - 2), p_first + q, &full[s]);
                    }
                    if (++s == S) { s = 0; if (q >= S) eph ^= 1; }
                }
            }
            return;
        }
        const double* gplane = in + (i64)p_first * sx + zlo;
        for (int q = 0; q < nplanes; ++q) {
            if (q >= S) mbar_wait(&empty[s], eph ^ 1);
            const int p = p_first + q;
            if (p < 0 || p >= nx) {
                if (lane == 0) mbar_arrive(&full[s]);
            } else {
                if (lane == 0) mbar_arrive_expect_tx(&full[s], plane_bytes);
                __syncwarp();
                double* slot = ring + (size_t)s * PLANE + dst_col;
                for (int y = ylo + lane; y < yhi; y += 32)
                    bulk_g2s(slot + (y - (yb - 2)) * RING_PITCH, gplane + (i64)y * nz, row_bytes, &full[s]);
            }
            gplane += sx;
            if (++s == S) { s = 0; if (q >= S) eph ^= 1; }
        }
        return;
    }

    // ===================== consumer warps =====================
    const int j0 = warp * R;
    const int y0 = yb + j0;
    if (y0 >= ny) {                                    // ragged last tile: only hand the slots back
        for (int q = 2; q < nplanes; ++q) {
            mbar_wait(&full[q % S], (q / S) & 1);
            if (lane == 0) mbar_arrive(&empty[(q - 2) % S]);
        }
        return;
    }
    const int z = z0 + 2 * lane;
    const bool zfix_lo = (z0 == 0) && (lo == nullptr), zfix_hi = (z0 + RING_TZ == nz) && (hi == nullptr);
    const bool rowfix = (y0 < 2) || (y0 + R > ny - 2);                            // warp-uniform, rare
    const int ztile = ab_zsingle ? ((zfix_lo || zfix_hi) ? 3 : 0)
                                 : ((zfix_lo ? 1 : 0) | (zfix_hi ? 2 : 0));      // warp-uniform
    const int zfix = (zfix_lo && lane == 0) ? 1 : ((zfix_hi && lane == 31) ? 2 : 0);
    // this lane's halo pair address in slot 0 (0 = not an edge lane of an edge tile)
    unsigned a_hlo = (lo_tile && lane == 0) ? smem_u32(hlo) + (unsigned)(j0 + 2) * 16u : 0u;
    unsigned a_hhi = (hi_tile && lane == 31) ? smem_u32(hhi) + (unsigned)(j0 + 2) * 16u : 0u;

    const unsigned my_off = (unsigned)(((j0 + 2) * RING_PITCH + 2 + 2 * lane) * 8);
    const unsigned ring_a = smem_u32(ring) + my_off;
    unsigned a_new = ring_a, bar_new = smem_u32(full), ph = 0;
    unsigned a_cen = ring_a, bar_cen = smem_u32(empty);          // centre = plane q-2 (valid from q = 2)
    int s_new = 0, s_cen = 0;
    int x = xs - 2;                                               // centre plane index once q >= 2 (x = p_first + q - 2)
    i64 goff = (i64)(xs) * sx + (i64)y0 * nz + z;                 // offset of (x = xs, y0, z)

    Rows<R> w0, w1, w2, w3, w4;
#define NG_ADV_NEW()                                                                   \
    a_new += PLANE_B; bar_new += 8;                                                    \
    if (++s_new == S) { s_new = 0; a_new -= RING_B; bar_new -= 8 * S; ph ^= 1; }
#define NG_ADV_CEN()                                                                   \
    a_cen += PLANE_B; bar_cen += 8;                                                    \
    if (SLAB) { if (a_hlo) a_hlo += HALO_B; if (a_hhi) a_hhi += HALO_B; }              \
    if (++s_cen == S) {                                                                \
        s_cen = 0; a_cen -= RING_B; bar_cen -= 8 * S;                                  \
        if (SLAB) { if (a_hlo) a_hlo -= S * HALO_B; if (a_hhi) a_hhi -= S * HALO_B; }  \
    }
#define NG_LOAD_NEW(NEW)                                                               \
    mbar_wait_a(bar_new, ph);                                                          \
    _Pragma("unroll") for (int r = 0; r < R; ++r) NEW.v[r] = lds2(a_new + r * PB);     \
    NG_ADV_NEW()
#define NG_RELEASE_CEN()                                                               \
    __syncwarp();                                                                      \
    if (lane == 0) mbar_arrive_a(bar_cen);                                             \
    NG_ADV_CEN()
    // ---- prologue: planes xs-2 .. xs+1 fill the window; planes xs-2, xs-1 are never a centre
    NG_LOAD_NEW(w0)
    NG_LOAD_NEW(w1)
    NG_LOAD_NEW(w2)
    NG_RELEASE_CEN()
    NG_LOAD_NEW(w3)
    NG_RELEASE_CEN()
    x = xs;
    int left = xe - xs;                                           // steps to go
#define NG_RING2_STEP(A, B, C, D, NEW)                                                 \
    {                                                                                  \
        NG_LOAD_NEW(NEW)                                                               \
        ring2_compute<R, SLAB>(x, y0, nx, ny, nz, sx, goff, W, side, in, out, a_hlo, a_hhi, a_cen, \
                               rowfix, ztile, zfix, A, B, C, D, NEW);                         \
        NG_RELEASE_CEN()                                                               \
        ++x; goff += sx;                                                               \
        if (--left == 0) break;                                                        \
    }
    for (;;) {
        NG_RING2_STEP(w0, w1, w2, w3, w4)
        NG_RING2_STEP(w1, w2, w3, w4, w0)
        NG_RING2_STEP(w2, w3, w4, w0, w1)
        NG_RING2_STEP(w3, w4, w0, w1, w2)
        NG_RING2_STEP(w4, w0, w1, w2, w3)
    }
#undef NG_RING2_STEP
#undef NG_RELEASE_CEN
#undef NG_LOAD_NEW
#undef NG_ADV_CEN
#undef NG_ADV_NEW
}

bool tables_match(const FdTables& T) {
    if (T.n_center != 5 || T.n_side != 6 || T.nb != 2) return false;
    for (int k = 0; k < 5; ++k) if (T.oc[k] != k - 2) return false;
    for (int k = 0; k < 6; ++k) if (T.of[k] != k || T.ob[k] != -k) return false;
    return true;
}

}  // namespace

// 3-D tensor map over the input array (z fastest): box = 68 x rows x 1, no swizzle, zero OOB fill.
typedef CUresult (*ng_encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                 const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                 CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static int make_tmap(CUtensorMap* tm, const double* in, i64 nx, i64 ny, i64 nz, int box_z, int rows) {
    static ng_encode_fn encode = nullptr;
    if (!encode) {
        void* fp = nullptr;
        cudaDriverEntryPointQueryResult qres;
        NG_CUDA_OK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &qres));
        if (!fp || qres != cudaDriverEntryPointSuccess) {
            ng_set_error("cuTensorMapEncodeTiled is not available from this driver");
            return NG_ECUDA;
        }
        encode = (ng_encode_fn)fp;
    }
    const cuuint64_t gdim[3] = {(cuuint64_t)nz, (cuuint64_t)ny, (cuuint64_t)nx};
    const cuuint64_t gstr[2] = {(cuuint64_t)nz * 8, (cuuint64_t)ny * nz * 8};
    const cuuint32_t box[3] = {(cuuint32_t)box_z, (cuuint32_t)rows, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = encode(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, (void*)in, gdim, gstr, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        ng_set_error("cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
        return NG_ECUDA;
    }
    return NG_OK;
}

// tuning knobs (read once from the environment; defaults chosen on B200)
static int env_int(const char* name, int dflt) {
    const char* s = getenv(name);
    return s ? atoi(s) : dflt;
}

// 2-D tensor map over a halo array [nx][ny][2] viewed as nx rows of 2*ny doubles: box = 2*rows x 1.
static int make_tmap_halo(CUtensorMap* tm, const double* halo, i64 nx, i64 ny, int rows) {
    ng_encode_fn encode = nullptr;
    {
        void* fp = nullptr;
        cudaDriverEntryPointQueryResult qres;
        NG_CUDA_OK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &qres));
        if (!fp || qres != cudaDriverEntryPointSuccess) {
            ng_set_error("cuTensorMapEncodeTiled is not available from this driver");
            return NG_ECUDA;
        }
        encode = (ng_encode_fn)fp;
    }
    const cuuint64_t gdim[2] = {(cuuint64_t)(2 * ny), (cuuint64_t)nx};
    const cuuint64_t gstr[1] = {(cuuint64_t)(2 * ny) * 8};
    const cuuint32_t box[2] = {(cuuint32_t)(2 * rows), 1};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = encode(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, (void*)halo, gdim, gstr, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        ng_set_error("cuTensorMapEncodeTiled (halo) failed with CUresult %d", (int)r);
        return NG_ECUDA;
    }
    return NG_OK;
}

int ng_fdlap3d_try_launch(const ng_plan* p, const double* in, double* out, const double* lo,
                          const double* hi, int xbeg, int xend, int zmode, cudaStream_t st, bool* handled) {
    *handled = false;
    if (p->ndim != 3) return NG_OK;
    const i64 nx = p->shape[0], ny = p->shape[1], nz = p->shape[2];
    if (nz % 2 || nx < 6 || ny < 6 || nz < 8 || nx * ny * nz >= (1LL << 40)) return NG_OK;
    if ((((uintptr_t)in | (uintptr_t)out | (uintptr_t)lo | (uintptr_t)hi) & 15) != 0) return NG_OK;
    for (int a = 0; a < 3; ++a) if (!tables_match(p->lap[a])) return NG_OK;
    const int force_generic = env_int("NG_FDLAP_GENERIC", 0);
    if (force_generic) return NG_OK;

    Lap3W W;
    for (int a = 0; a < 3; ++a) {
        for (int k = 0; k < 5; ++k) W.wc[a][k] = p->lap[a].wc[k];
        for (int k = 0; k < 6; ++k) { W.wf[a][k] = p->lap[a].wf[k]; W.wb[a][k] = p->lap[a].wb[k]; }
        W.ih[a] = p->lap[a].inv_h_pow;
    }
    // Defaults from the B200 sweeps in profiles/ (tools/tune_fdlap.py): plane-ring kernel (v4) with
    // R=2 rows/warp, 4 consumer warps, 8 ring slots, 64-plane x chunks.  Environment overrides are
    // for tuning only.
    const int cfg_R = env_int("NG_FDLAP_R", 2);
    const int cfg_v = env_int("NG_FDLAP_V", 4);
    const int cfg_wy = env_int("NG_FDLAP_WY", 4);
    const unsigned gx = (unsigned)((nz + 63) / 64);
    int xchunk = env_int("NG_FDLAP_XCHUNK", 0);
    if (xchunk < 1) {
        // 64 planes per chunk on big grids; shorter chunks when that would leave SMs idle
        const i64 tiles = (i64)gx * ((ny + cfg_R * cfg_wy - 1) / (cfg_R * cfg_wy));
        const i64 want_chunks = (148 * 6 + tiles - 1) / tiles;
        xchunk = 64;
        if (want_chunks > 1 && (xend - xbeg) / want_chunks < 64) xchunk = (int)((xend - xbeg) / want_chunks);
        if (xchunk < 8) xchunk = 8;
    }
    if (xchunk > xend - xbeg) xchunk = xend - xbeg;
    const unsigned gz = (unsigned)((xend - xbeg + xchunk - 1) / xchunk);
    const bool same_w = memcmp(W.wc[0], W.wc[1], sizeof(W.wc[0])) == 0 &&
                        memcmp(W.wc[0], W.wc[2], sizeof(W.wc[0])) == 0;
    if (cfg_v == 4 && same_w && nz % 64 == 0 && ny * nz < (1LL << 28) && p->d_Mt) {
        Lap3U WU;
        for (int k = 0; k < 5; ++k) WU.w[k] = W.wc[0][k];
        for (int a = 0; a < 3; ++a) WU.ih[a] = W.ih[a];
        for (int k = 0; k < 6; ++k) { WU.wfz[k] = W.wf[2][k]; WU.wbz[k] = W.wb[2][k]; }
        const int cfg_s = env_int("NG_FDLAP_S", 8);
        const bool slab = (lo != nullptr) || (hi != nullptr);
        const bool use_tmap = env_int("NG_FDLAP_TMAP", 1) != 0;
        const bool ab_halo3d = env_int("NG_FDLAP_HALO3D", 0) != 0;          // A/B switches (tools/ab_fdlap.py)
        const int ab_flags = (env_int("NG_FDLAP_ZSINGLE", 0) ? 16 : 0) | (ab_halo3d ? 32 : 0);
        bool launched = false, skip_count = false;
#define NG_RING2_GO1(RR, WW, SS, MB, SL, TM)                                                         \
    {                                                                                              \
        constexpr int plane = (RR * WW + 4) * RING_PITCH;                                          \
        const size_t smem = (size_t)SS * plane * sizeof(double) + 2 * SS * sizeof(unsigned long long) + \
                            (SL ? (size_t)SS * 2 * ((((RR * WW + 4) * 2 + 15) / 16) * 16) * sizeof(double) : 0); \
        static bool attr_done = false;                                                             \
        if (!attr_done) {                                                                          \
            NG_CUDA_OK(cudaFuncSetAttribute(fdlap3d_ring2_kernel<RR, WW, SS, MB, SL, TM>,              \
                                            cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem + 128)); \
            attr_done = true;                                                                      \
        }                                                                                          \
        const unsigned ztiles = (unsigned)(nz / RING_TZ);                                          \
        const unsigned gxz = zmode == 1 ? (ztiles > 2 ? ztiles - 2 : 0) : (zmode == 2 ? (ztiles < 2 ? ztiles : 2) : gx); \
        if (gxz == 0) { launched = true; skip_count = true; } else {                               \
        dim3 grid(gxz, (unsigned)((ny + RR * WW - 1) / (RR * WW)), gz);                            \
        CUtensorMap tm, tm_lo, tm_hi;                                                              \
        memset(&tm, 0, sizeof(tm)); memset(&tm_lo, 0, sizeof(tm)); memset(&tm_hi, 0, sizeof(tm));  \
        if (TM) {                                                                                  \
            int trc = make_tmap(&tm, in, nx, ny, nz, RING_PITCH, RR * WW + 4);                     \
            if (!trc && SL && lo) trc = ab_halo3d ? make_tmap(&tm_lo, lo, nx, ny, 2, 2, RR * WW + 4) : make_tmap_halo(&tm_lo, lo, nx, ny, RR * WW + 4); \
            if (!trc && SL && hi) trc = ab_halo3d ? make_tmap(&tm_hi, hi, nx, ny, 2, 2, RR * WW + 4) : make_tmap_halo(&tm_hi, hi, nx, ny, RR * WW + 4); \
            if (trc) return trc;                                                                   \
        }                                                                                          \
        fdlap3d_ring2_kernel<RR, WW, SS, MB, SL, TM><<<grid, 32 * (WW + 1), smem + 128, st>>>(     \
            in, out, (int)nx, (int)ny, (int)nz, xchunk, xbeg, xend, zmode | ab_flags, WU, (const Lap3Side*)p->d_Mt, lo, hi, tm,  \
            tm_lo, tm_hi);                                                                         \
        launched = true; }                                                                         \
    }
#define NG_RING2_GO(RR, WW, SS, MB)                                                                \
    if (!launched && cfg_R == RR && cfg_wy == WW && cfg_s == SS &&                                 \
        (ny % (RR * WW) == 0 || ny % (RR * WW) >= 5)) {                                            \
        if (slab) {                                                                                \
            if (use_tmap) { NG_RING2_GO1(RR, WW, SS, MB, true, true) }                             \
        } else {                                                                                   \
            if (use_tmap) { NG_RING2_GO1(RR, WW, SS, MB, false, true) } else { NG_RING2_GO1(RR, WW, SS, MB, false, false) } \
        }                                                                                          \
    }
        NG_RING2_GO(2, 8, 6, 2)
        NG_RING2_GO(2, 8, 5, 2)
        NG_RING2_GO(2, 8, 8, 2)
        NG_RING2_GO(2, 4, 6, 3)
        NG_RING2_GO(2, 4, 8, 3)
        NG_RING2_GO(1, 8, 6, 3)
        NG_RING2_GO(4, 4, 6, 2)
        NG_RING2_GO(4, 4, 5, 2)
        NG_RING2_GO(3, 8, 6, 1)
        NG_RING2_GO(2, 6, 6, 2)
        NG_RING2_GO(2, 6, 8, 2)
#undef NG_RING2_GO
#undef NG_RING2_GO1
        if (launched && skip_count) {          // nothing to do for this z part (e.g. no interior tiles)
            *handled = true;
            return NG_OK;
        }
        if (launched) {
            NG_LAUNCH_CHECK();
            *handled = true;
            return NG_OK;
        }
    }
    if (zmode == 1) { *handled = true; return NG_OK; }     // fallback kernels: part 2 computes everything
    if (cfg_v == 3 && nz % 64 == 0 && ny * nz < (1LL << 28) && p->d_Mt) {
        Lap3C WC;
        for (int a = 0; a < 3; ++a) {
            for (int k = 0; k < 5; ++k) WC.wc[a][k] = W.wc[a][k];
            WC.ih[a] = W.ih[a];
        }
        const int cfg_s = env_int("NG_FDLAP_S", 6);
        bool launched = false;
#define NG_RING_GO(RR, WW, SS, MB)                                                                   \
    if (!launched && cfg_R == RR && cfg_wy == WW && cfg_s == SS &&                                 \
        (ny % (RR * WW) == 0 || ny % (RR * WW) >= 5)) {                                            \
        constexpr int plane = (RR * WW + 4) * RING_PITCH;                                          \
        const size_t smem = (size_t)SS * plane * sizeof(double) + 2 * SS * sizeof(unsigned long long); \
        static bool attr_done = false;                                                             \
        if (!attr_done) {                                                                          \
            NG_CUDA_OK(cudaFuncSetAttribute(fdlap3d_ring_kernel<RR, WW, SS, MB>,                       \
                                            cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
            attr_done = true;                                                                      \
        }                                                                                          \
        dim3 grid(gx, (unsigned)((ny + RR * WW - 1) / (RR * WW)), gz);                             \
        fdlap3d_ring_kernel<RR, WW, SS, MB><<<grid, 32 * (WW + 1), smem, st>>>(                        \
            in, out, (int)nx, (int)ny, (int)nz, xchunk, xbeg, xend, WC, (const Lap3Side*)p->d_Mt, lo, hi);     \
        launched = true;                                                                           \
    }
        NG_RING_GO(2, 8, 6, 2)
        NG_RING_GO(2, 8, 5, 2)
        NG_RING_GO(2, 8, 8, 2)
        NG_RING_GO(2, 4, 6, 3)
        NG_RING_GO(2, 4, 8, 3)
        NG_RING_GO(1, 8, 6, 3)
        NG_RING_GO(1, 8, 8, 2)
        NG_RING_GO(4, 4, 6, 2)
        NG_RING_GO(4, 4, 5, 2)
        NG_RING_GO(4, 8, 5, 1)
        NG_RING_GO(3, 8, 6, 1)
#undef NG_RING_GO
        if (launched) {
            NG_LAUNCH_CHECK();
            *handled = true;
            return NG_OK;
        }
    }
    if (cfg_v >= 2 && nz % 64 == 0 && ny % cfg_R == 0 && ny * nz < (1LL << 28) && p->d_Mt) {
        Lap3C WC;
        for (int a = 0; a < 3; ++a) {
            for (int k = 0; k < 5; ++k) WC.wc[a][k] = W.wc[a][k];
            WC.ih[a] = W.ih[a];
        }
#define NG_FAST_GO(RR, WW, MB)                                                                 \
    do {                                                                                       \
        dim3 grid(gx, (unsigned)((ny + RR * WW - 1) / (RR * WW)), gz);                         \
        fdlap3d_fast_kernel<RR, WW, MB><<<grid, 32 * WW, 0, st>>>(                             \
            in, out, (int)nx, (int)ny, (int)nz, xchunk, xbeg, xend, WC, (const Lap3Side*)p->d_Mt, lo, hi); \
    } while (0)
        if (cfg_R == 1) { if (cfg_wy == 4) NG_FAST_GO(1, 4, 4); else NG_FAST_GO(1, 8, 2); }
        else if (cfg_R == 2) { if (cfg_wy == 4) NG_FAST_GO(2, 4, 4); else NG_FAST_GO(2, 8, 2); }
        else if (cfg_R == 3) { if (cfg_wy == 4) NG_FAST_GO(3, 4, 3); else NG_FAST_GO(3, 8, 1); }
        else { if (cfg_wy == 4) NG_FAST_GO(4, 4, 2); else NG_FAST_GO(4, 8, 1); }
#undef NG_FAST_GO
        NG_LAUNCH_CHECK();
        *handled = true;
        return NG_OK;
    }
#define NG_LAP_GO(RR, WW)                                                                    \
    do {                                                                                       \
        dim3 grid(gx, (unsigned)((ny + RR * WW - 1) / (RR * WW)), gz);                         \
        fdlap3d_march_kernel<RR, WW><<<grid, 32 * WW, 0, st>>>(in, out, (int)nx, (int)ny,      \
                                                               (int)nz, xchunk, xbeg, xend, W, lo, hi);    \
    } while (0)
    if (cfg_R == 1) NG_LAP_GO(1, 8);
    else if (cfg_R == 2) NG_LAP_GO(2, 8);
    else if (cfg_R == 8) NG_LAP_GO(8, 4);
    else NG_LAP_GO(4, 8);
#undef NG_LAP_GO
    NG_LAUNCH_CHECK();
    *handled = true;
    return NG_OK;
}
