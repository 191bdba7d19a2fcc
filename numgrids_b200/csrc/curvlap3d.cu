// curvlap3d.cu -- single-pass fused curvilinear Laplacian on 3-D grids whose three axes are
// finite-difference axes (equidistant or logarithmic):
//
//     res = 0;  for i in (0, 1, 2):  inner = (J / h_i**2) * D_i f;  res = res + D_i(inner)
//     res = res / J;  res = where(isfinite(res), res, 0)            (numgrids/curvilinear.py:235-244)
//
// with D_i = Diff(grid, 1, i) (acc = 4: central taps -2..2, one-sided 5 taps; LogDiff divides by x).
// The generic path (curv.cu) runs this as six fused per-axis passes through a workspace array:
// 6 x 16 B/point of memory traffic.  Here `inner` never leaves the SM: one read of f and one write
// of the result per point (16 B/point + the compact coefficient tables).
//
// Mapping.  A CTA owns a tile of TY = R*WY rows (y) x 64 z and marches along x; planes of the tile with a
// +-4 halo (two chained radius-2 stencils) stream through a ring of S = 8 shared-memory slots -- one 3-D
// tensor-map TMA per plane, box 72 x (TY+8) x 1, out-of-bounds zero-filled, issued by whichever warp
// frees a slot last; the WY warps, each R rows x 64 z (a double2 per lane and row), work independently
// of each other (mbarriers and one shared counter per slot only):
//   x: gx(q) = c_x * D_x f at plane q from the own pairs of planes q-2..q+2 in the ring (5 LDS.128); a
//      5-plane register window of gx gives the outer derivative of plane q-2;
//   y: the R+8 rows y0-4..y0+R+3 of the centre plane's own column pair (LDS.128) -> gy on rows
//      y0-2..y0+R+1 -> outer derivative on the R own rows, all in registers;
//   z: gz of the own pair from the row in the ring; neighbours' gz by warp shuffles; the two pairs a
//      warp needs beyond its 64 columns are computed by lanes 0..2R-1 in one extra pass.
// One-sided schemes: z inline (the first / last z tile knows it: warp-uniform branches, all operands in
// the ring / in neighbouring lanes); x and y through an out-of-line routine that evaluates
// D_i(c_i D_i f) of one column pair straight from global memory -- on the planes x < 2, x >= nx-2 (and
// the gx of the planes q < 2, q >= nx-2) and for the first / last warp of the y range only.
//
// Arithmetic is the reference's, operation for operation (separately rounded DMUL / DADD / DDIV, no
// FMA), with one exact simplification: every tap sum starts from its first product instead of from
// 0.0.  `0.0 + p` differs from `p` only for p = -0.0; a sum can then end as -0.0 instead of +0.0 only if
// every product was -0.0, which changes nothing downstream but the sign of a zero that is then added to a
// running result that is never -0.0 (it starts as `0.0 + X`): the stored bits are the reference's.
// Coefficients that are the constant 1.0 (unit-scale grids: the second reference idiom of the Cartesian
// Laplacian) are skipped -- `1.0 * v` and `v / 1.0` are exact.
#include <cuda.h>

#include "common.cuh"

namespace {

// suspend-time hint of the mbarrier waits (see NG_MBAR_SUSPEND_NS in common.cuh); NG_MBAR_HINT_NS overrides it for A/B runs
__constant__ unsigned c_mbar_hint_ns = NG_MBAR_SUSPEND_NS;
int sync_mbar_hint(cudaStream_t st) {
    static std::atomic<long long> current{NG_MBAR_SUSPEND_NS};
    const long long want = NG_ENV_INT("NG_MBAR_HINT_NS", NG_MBAR_SUSPEND_NS);
    if (want != current.load(std::memory_order_relaxed)) {
        const unsigned v = (unsigned)want;
        NG_CUDA_OK(cudaMemcpyToSymbolAsync(c_mbar_hint_ns, &v, sizeof(v), 0, cudaMemcpyHostToDevice, st));
        current.store(want, std::memory_order_relaxed);
    }
    return NG_OK;
}


constexpr int CL_TZ = 64;                    // z points per tile (one double2 per lane)
constexpr int CL_H = 4;                      // halo on every side
constexpr int CL_PITCH = CL_TZ + 2 * CL_H;   // doubles per staged row

struct ClAxis {
    double wc[5], wf[5], wb[5], ih;
};
struct ClTabs {                  // kernel parameter (constant bank): the fast path reads wc / ih, z also wf / wb
    ClAxis a[3];
};
struct ClGen {                   // general grids only: coefficient tables (nullptr = the constant 1.0) and
    CoefDev c[3];                // J / h_i**2
    CoefDev J;
    const double* xdiv[3];       // log axes: the coordinates LogDiff divides by (nullptr otherwise)
};
struct ClParams {                // device copy for the out-of-line boundary routine
    ClTabs T;
    ClGen G;
};

__device__ __forceinline__ unsigned cl_smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ double2 cl_lds2(unsigned addr) {
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr));
    return v;
}
__device__ __forceinline__ double2 cl_ld2(const double* p) { return *reinterpret_cast<const double2*>(p); }
__device__ __forceinline__ bool cl_try(unsigned addr, unsigned parity) {
    unsigned done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done) : "r"(addr), "r"(parity), "r"(c_mbar_hint_ns) : "memory");
    return done != 0;
}
__device__ __noinline__ void cl_wait_slow(unsigned addr, unsigned parity) {
    unsigned spins = 0;
    while (!cl_try(addr, parity))
        if (++spins > (1u << 24)) __trap();          // a protocol bug traps instead of hanging the GPU
}
__device__ __forceinline__ void cl_wait(unsigned addr, unsigned parity) {
    if (!cl_try(addr, parity)) cl_wait_slow(addr, parity);
}
__device__ __forceinline__ void cl_arrive(unsigned addr) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(addr) : "memory");
}
__device__ __forceinline__ double cl_mad(double acc, double w, double f) { return __dadd_rn(acc, __dmul_rn(w, f)); }

// sum_k w[k] * v[k] in stored order starting from the first product, on both halves of a pair
__device__ __forceinline__ double2 cl_sum5(const double* w, double2 a, double2 b, double2 c, double2 d, double2 e) {
    double sx = __dmul_rn(w[0], a.x), sy = __dmul_rn(w[0], a.y);
    sx = cl_mad(sx, w[1], b.x); sy = cl_mad(sy, w[1], b.y);
    sx = cl_mad(sx, w[2], c.x); sy = cl_mad(sy, w[2], c.y);
    sx = cl_mad(sx, w[3], d.x); sy = cl_mad(sy, w[3], d.y);
    sx = cl_mad(sx, w[4], e.x); sy = cl_mad(sy, w[4], e.y);
    return make_double2(sx, sy);
}
__device__ __forceinline__ double cl_sum5s(const double* w, double a, double b, double c, double d, double e) {
    double s = __dmul_rn(w[0], a);
    s = cl_mad(s, w[1], b);
    s = cl_mad(s, w[2], c);
    s = cl_mad(s, w[3], d);
    return cl_mad(s, w[4], e);
}
__device__ __forceinline__ double2 cl_scale(double2 s, double ih) {
    return make_double2(__dmul_rn(s.x, ih), __dmul_rn(s.y, ih));
}
__device__ __forceinline__ double2 cl_coef2(const CoefDev& c, int x, int y, int z) {
    const i64 o = x * c.s[0] + y * c.s[1] + z * c.s[2];
    return make_double2(__ldg(c.ptr + o), __ldg(c.ptr + o + c.s[2]));
}
// the elementwise steps between the tap sum and the next stage: * inv_h, / x (LogDiff), * coefficient
template <bool GEN>
__device__ __forceinline__ double2 cl_finish_xy(double2 s, double ih, const ClGen& G, int axis, int idx, bool with_coef,
                                                int x, int y, int z) {
    double2 v = cl_scale(s, ih);
    if (GEN) {
        if (G.xdiv[axis]) {
            const double d = __ldg(G.xdiv[axis] + idx);
            v.x = __ddiv_rn(v.x, d);
            v.y = __ddiv_rn(v.y, d);
        }
        if (with_coef && G.c[axis].ptr) {
            const double2 c = cl_coef2(G.c[axis], x, y, z);
            v.x = __dmul_rn(c.x, v.x);
            v.y = __dmul_rn(c.y, v.y);
        }
    }
    return v;
}
template <bool GEN>
__device__ __forceinline__ double2 cl_finish_z(double2 s, double ih, const ClGen& G, bool with_coef, int x, int y, int z) {
    double2 v = cl_scale(s, ih);
    if (GEN) {
        if (G.xdiv[2]) {
            v.x = __ddiv_rn(v.x, __ldg(G.xdiv[2] + z));
            v.y = __ddiv_rn(v.y, __ldg(G.xdiv[2] + z + 1));
        }
        if (with_coef && G.c[2].ptr) {
            const double2 c = cl_coef2(G.c[2], x, y, z);
            v.x = __dmul_rn(c.x, v.x);
            v.y = __dmul_rn(c.y, v.y);
        }
    }
    return v;
}

// ---- out-of-line boundary routine (global memory) ------------------------------------------------
// g(i) = c * D f at index i of `axis` (0 or 1) for the column pair at `p` (= &f[x][y][z], axis index i),
// scheme chosen by i like the reference: forward on [0, 2), backward on [n-2, n), central between.
__device__ double2 cl_inner_global(const ClParams* P, bool gen, const double* p, i64 stride, int axis, int i, int n,
                                   int x, int y, int z) {
    const ClAxis& A = P->T.a[axis];
    double2 s;
    if (i < 2) s = cl_sum5(A.wf, cl_ld2(p), cl_ld2(p + stride), cl_ld2(p + 2 * stride), cl_ld2(p + 3 * stride), cl_ld2(p + 4 * stride));
    else if (i >= n - 2) s = cl_sum5(A.wb, cl_ld2(p), cl_ld2(p - stride), cl_ld2(p - 2 * stride), cl_ld2(p - 3 * stride), cl_ld2(p - 4 * stride));
    else s = cl_sum5(A.wc, cl_ld2(p - 2 * stride), cl_ld2(p - stride), cl_ld2(p), cl_ld2(p + stride), cl_ld2(p + 2 * stride));
    return gen ? cl_finish_xy<true>(s, A.ih, P->G, axis, i, true, x, y, z) : cl_finish_xy<false>(s, A.ih, P->G, axis, i, true, x, y, z);
}
__device__ __noinline__ double2 cl_gx_global(const ClParams* P, int gen, const double* p, i64 stride, int q, int nx, int y, int z) {
    return cl_inner_global(P, gen != 0, p, stride, 0, q, nx, q, y, z);
}
// D_axis(c D_axis f) at index i of `axis` (0 or 1), everything from global memory
__device__ __noinline__ double2 cl_term_global(const ClParams* P, int gen, const double* p, i64 stride, int axis, int i, int n,
                                               int x, int y, int z) {
    const ClAxis& A = P->T.a[axis];
    const double* w;
    int first, step;
    if (i < 2) { w = A.wf; first = 0; step = 1; }
    else if (i >= n - 2) { w = A.wb; first = 0; step = -1; }
    else { w = A.wc; first = -2; step = 1; }
    double2 g[5];
#pragma unroll
    for (int k = 0; k < 5; ++k) {
        const int o = (first + k) * step;
        g[k] = cl_inner_global(P, gen != 0, p + o * stride, stride, axis, i + o, n, axis == 0 ? x + o : x, axis == 1 ? y + o : y, z);
    }
    const double2 s = cl_sum5(w, g[0], g[1], g[2], g[3], g[4]);
    return gen ? cl_finish_xy<true>(s, A.ih, P->G, axis, i, false, x, y, z) : cl_finish_xy<false>(s, A.ih, P->G, axis, i, false, x, y, z);
}

__device__ __forceinline__ void cl_tma_3d(unsigned dst, const CUtensorMap* tmap, int c0, int c1, int c2, unsigned bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes "
        "[%0], [%1, {%2, %3, %4}], [%5];"
        ::"r"(dst), "l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(bar) : "memory");
}

template <int R>
struct ClRows {
    double2 v[R];
};

// The work of one consumer warp for output plane x (its R rows): y and z terms from the centre slot, the
// x term from the gx window, combine, store.
// YS: the y inner derivatives are SHARED through shared memory instead of recomputed per warp: every warp forms gy on
// its own R rows only (and warps 0..3 on one of the four rows just outside the tile), publishes them, the CTA meets at
// one named barrier per plane, and the outer derivative reads the R + 4 rows it needs -- (R + 0.5) / R instead of
// (R + 4) / R inner derivatives per own row.  gy_w: shared address of this plane's buffer ([TY + 4][32] pairs, row 0 =
// tile row -2); warp / WY / yb place the warp in the tile.
template <int R, bool GEN, bool YS, int WY>
__device__ __forceinline__ void cl_output(int x, int y0, int z0, int lane, int nx, int ny, int nz, const ClTabs& T, const ClGen& G,
                                          const ClParams* P, const double* __restrict__ in, double* __restrict__ out,
                                          unsigned a_cen /*smem: (row y0, own pair) of plane x*/, unsigned a_tile /*smem: (row y0, column z0)*/,
                                          const ClRows<R>& G0, const ClRows<R>& G1, const ClRows<R>& G2,
                                          const ClRows<R>& G3, const ClRows<R>& G4, unsigned gy_w, int warp, int yb) {
    constexpr unsigned PB = CL_PITCH * 8;
    const int z = z0 + 2 * lane;
    const i64 sx = (i64)ny * nz;
    const i64 goff = x * sx + (i64)y0 * nz + z;
    const bool xfix = (x < 2) || (x >= nx - 2);                   // block-uniform
    const bool yfix = (y0 < 4) || (y0 + R > ny - 4);              // warp-uniform: rows within 4 of a y face
    const bool first = (z0 == 0), last = (z0 + CL_TZ == nz);      // block-uniform

    // ---------------- y: rows y0-4 .. y0+R+3 of the own column pair
    double2 F[R + 8];
    double2 Y[R];
    if (YS) {
        constexpr int TYc = R * WY;
#pragma unroll
        for (int i = 2; i < R + 6; ++i) F[i] = cl_lds2(a_cen + (unsigned)((i - 4) * (int)PB));    // rows y0-2 .. y0+R+1
        double2 GO[R];                                   // gy on the own rows (central scheme; rows within 2 of a y face
#pragma unroll                                           // hold values nobody reads)
        for (int r = 0; r < R; ++r) {
            const int yr = y0 + r;
            GO[r] = make_double2(0.0, 0.0);
            if (yr >= 2 && yr < ny - 2)
                GO[r] = cl_finish_xy<GEN>(cl_sum5(T.a[1].wc, F[r + 2], F[r + 3], F[r + 4], F[r + 5], F[r + 6]), T.a[1].ih, G, 1,
                                          yr, true, x, yr, z);
            asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(gy_w + (unsigned)((warp * R + r + 2) * 512 + lane * 16)),
                         "d"(GO[r].x), "d"(GO[r].y) : "memory");
        }
        if (warp < 4) {                                  // one of the four rows just outside the tile: -2, -1, TY, TY+1
            const int hr = warp < 2 ? warp - 2 : TYc + warp - 2;
            const int yh = yb + hr;
            if (yh >= 2 && yh < ny - 2) {
                const unsigned a = a_tile + (unsigned)((hr - warp * R) * (int)PB) + 16u * lane;    // (row yh, own pair)
                const double2 h = cl_finish_xy<GEN>(cl_sum5(T.a[1].wc, cl_lds2(a - 2 * PB), cl_lds2(a - PB), cl_lds2(a), cl_lds2(a + PB),
                                                            cl_lds2(a + 2 * PB)), T.a[1].ih, G, 1, yh, true, x, yh, z);
                asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(gy_w + (unsigned)((hr + 2) * 512 + lane * 16)),
                             "d"(h.x), "d"(h.y) : "memory");
            }
        }
        asm volatile("bar.sync 1, %0;" ::"n"(32 * WY) : "memory");
        if (!yfix) {
            double2 GY[R + 4];
#pragma unroll
            for (int j = 0; j < R + 4; ++j)
                GY[j] = (j >= 2 && j < R + 2) ? GO[j - 2] : cl_lds2(gy_w + (unsigned)((warp * R + j) * 512 + lane * 16));
#pragma unroll
            for (int r = 0; r < R; ++r)
                Y[r] = cl_finish_xy<GEN>(cl_sum5(T.a[1].wc, GY[r], GY[r + 1], GY[r + 2], GY[r + 3], GY[r + 4]), T.a[1].ih, G, 1,
                                         y0 + r, false, x, y0 + r, z);
        } else {
#pragma unroll
            for (int r = 0; r < R; ++r) Y[r] = cl_term_global(P, GEN, in + goff + (i64)r * nz, nz, 1, y0 + r, ny, x, y0 + r, z);
        }
    } else {
#pragma unroll
    for (int i = 0; i < R + 8; ++i) F[i] = cl_lds2(a_cen + (unsigned)((i - 4) * (int)PB));
    if (!yfix) {
        double2 GY[R + 4];
#pragma unroll
        for (int j = 0; j < R + 4; ++j)
            GY[j] = cl_finish_xy<GEN>(cl_sum5(T.a[1].wc, F[j], F[j + 1], F[j + 2], F[j + 3], F[j + 4]), T.a[1].ih, G, 1,
                                      y0 - 2 + j, true, x, y0 - 2 + j, z);
#pragma unroll
        for (int r = 0; r < R; ++r)
            Y[r] = cl_finish_xy<GEN>(cl_sum5(T.a[1].wc, GY[r], GY[r + 1], GY[r + 2], GY[r + 3], GY[r + 4]), T.a[1].ih, G, 1,
                                     y0 + r, false, x, y0 + r, z);
    } else {
#pragma unroll
        for (int r = 0; r < R; ++r) Y[r] = cl_term_global(P, GEN, in + goff + (i64)r * nz, nz, 1, y0 + r, ny, x, y0 + r, z);
    }
    }

    // ---------------- z: inner derivative of the own pairs, of the 2R pairs beyond the warp's columns
    double2 GZ[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const unsigned a = a_cen + r * PB;
        const double2 A = cl_lds2(a - 16), B = F[r + 4], C = cl_lds2(a + 16);
        double2 s = make_double2(cl_sum5s(T.a[2].wc, A.x, A.y, B.x, B.y, C.x), cl_sum5s(T.a[2].wc, A.y, B.x, B.y, C.x, C.y));
        if (first && lane == 0) {                 // z = 0, 1: forward scheme on f(z .. z+4)
            const double2 D = cl_lds2(a + 32);
            s = make_double2(cl_sum5s(T.a[2].wf, B.x, B.y, C.x, C.y, D.x), cl_sum5s(T.a[2].wf, B.y, C.x, C.y, D.x, D.y));
        }
        if (last && lane == 31) {                 // z = nz-2, nz-1: backward scheme on f(z), f(z-1), .. f(z-4)
            const double2 Zp = cl_lds2(a - 32);
            s = make_double2(cl_sum5s(T.a[2].wb, B.x, A.y, A.x, Zp.y, Zp.x), cl_sum5s(T.a[2].wb, B.y, B.x, A.y, A.x, Zp.y));
        }
        GZ[r] = cl_finish_z<GEN>(s, T.a[2].ih, G, true, x, y0 + r, z);
    }
    double2 HZ = make_double2(0.0, 0.0);          // lane i < 2R: gz of row i/2 at columns z0-2, z0-1 (i even) / z0+64, z0+65 (i odd)
    if (lane < 2 * R) {
        const int hr = lane >> 1, hs = lane & 1;
        const int zc = hs ? z0 + CL_TZ : z0 - 2;
        if (zc >= 0 && zc < nz) {                 // (outside the grid on the first / last tile: never used)
            const unsigned a = a_tile + hr * PB + (hs ? CL_TZ * 8 : -16);
            const double2 A = cl_lds2(a - 16), B = cl_lds2(a), C = cl_lds2(a + 16);
            const double2 s = make_double2(cl_sum5s(T.a[2].wc, A.x, A.y, B.x, B.y, C.x), cl_sum5s(T.a[2].wc, A.y, B.x, B.y, C.x, C.y));
            HZ = cl_finish_z<GEN>(s, T.a[2].ih, G, true, x, y0 + hr, zc);
        }
    }

    // ---------------- per row: x term, z outer derivative, combine, store
#pragma unroll
    for (int r = 0; r < R; ++r) {
        double2 X;
        if (!xfix) X = cl_finish_xy<GEN>(cl_sum5(T.a[0].wc, G0.v[r], G1.v[r], G2.v[r], G3.v[r], G4.v[r]), T.a[0].ih, G, 0, x, false, x, y0 + r, z);
        else X = cl_term_global(P, GEN, in + goff + (i64)r * nz, sx, 0, x, nx, x, y0 + r, z);
        const double2 g = GZ[r];
        double2 gl = make_double2(__shfl_up_sync(0xffffffffu, g.x, 1), __shfl_up_sync(0xffffffffu, g.y, 1));
        double2 gr = make_double2(__shfl_down_sync(0xffffffffu, g.x, 1), __shfl_down_sync(0xffffffffu, g.y, 1));
        const double2 hl = make_double2(__shfl_sync(0xffffffffu, HZ.x, 2 * r), __shfl_sync(0xffffffffu, HZ.y, 2 * r));
        const double2 hh = make_double2(__shfl_sync(0xffffffffu, HZ.x, 2 * r + 1), __shfl_sync(0xffffffffu, HZ.y, 2 * r + 1));
        if (lane == 0) gl = hl;
        if (lane == 31) gr = hh;
        double2 s = make_double2(cl_sum5s(T.a[2].wc, gl.x, gl.y, g.x, g.y, gr.x), cl_sum5s(T.a[2].wc, gl.y, g.x, g.y, gr.x, gr.y));
        if (first) {                              // z = 0, 1: forward scheme on gz(z .. z+4)
            const double2 grr = make_double2(__shfl_down_sync(0xffffffffu, g.x, 2), __shfl_down_sync(0xffffffffu, g.y, 2));
            if (lane == 0)
                s = make_double2(cl_sum5s(T.a[2].wf, g.x, g.y, gr.x, gr.y, grr.x), cl_sum5s(T.a[2].wf, g.y, gr.x, gr.y, grr.x, grr.y));
        }
        if (last) {                               // z = nz-2, nz-1: backward scheme
            const double2 gll = make_double2(__shfl_up_sync(0xffffffffu, g.x, 2), __shfl_up_sync(0xffffffffu, g.y, 2));
            if (lane == 31)
                s = make_double2(cl_sum5s(T.a[2].wb, g.x, gl.y, gl.x, gll.y, gll.x), cl_sum5s(T.a[2].wb, g.y, g.x, gl.y, gl.x, gll.y));
        }
        const double2 Z = cl_finish_z<GEN>(s, T.a[2].ih, G, false, x, y0 + r, z);
        double2 o;
        o.x = __dadd_rn(__dadd_rn(__dadd_rn(0.0, X.x), Y[r].x), Z.x);
        o.y = __dadd_rn(__dadd_rn(__dadd_rn(0.0, X.y), Y[r].y), Z.y);
        if (GEN && G.J.ptr) {
            const double2 J = cl_coef2(G.J, x, y0 + r, z);
            o.x = __ddiv_rn(o.x, J.x);
            o.y = __ddiv_rn(o.y, J.y);
        }
        if (ng_nonfinite(o.x)) o.x = 0.0;
        if (ng_nonfinite(o.y)) o.y = 0.0;
        *reinterpret_cast<double2*>(out + goff + (i64)r * nz) = o;
    }
}

// No producer warp: the register file is split by scheduler, so a ninth warp on the SM would cap every
// thread at 168 registers (three warps on one scheduler) and the 5-plane gx window would spill; with
// exactly two CTAs x four warps each thread may use 255.  Slots are handed back through a shared counter:
// the LAST warp to finish an iteration re-arms the slot's mbarrier and issues the TMA for the plane that
// takes its place at once, so the ring stays S-5 planes ahead without anybody waiting for a producer.
template <int R, int WY, int S, bool GEN, bool YS = false>
__global__ void __launch_bounds__(32 * WY, (R * WY >= 32 ? 1 : 2))
curvlap3d_kernel(double* __restrict__ out, const double* __restrict__ in, int nx, int ny, int nz, int xchunk,
                 const __grid_constant__ ClTabs T, const __grid_constant__ ClGen G, const ClParams* __restrict__ P,
                 const __grid_constant__ CUtensorMap tmap) {
    constexpr int TY = R * WY;
    constexpr int ROWS = TY + 2 * CL_H;
    constexpr unsigned PB = CL_PITCH * 8, PLANE_B = ROWS * PB;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const unsigned ring = (cl_smem_u32(smem_raw) + 127u) & ~127u;        // TMA destinations: 128-byte aligned
    const unsigned full = ring + S * PLANE_B;
    int* const done = reinterpret_cast<int*>(smem_raw + (full - cl_smem_u32(smem_raw)) + 8 * S);   // warps finished with a slot
    const unsigned gy_base = (full + 12 * S + 15u) & ~15u;     // YS: two buffers of [TY + 4][32] pairs, alternating per plane
    constexpr unsigned GY_B = (TY + 4) * 512;

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int z0 = blockIdx.x * CL_TZ;
    const int yb = blockIdx.y * TY;
    const int xs = blockIdx.z * xchunk;
    const int xe = min(nx, xs + xchunk);
    const int p_first = xs - 4;                       // planes xs-4 .. xe+3 travel through the ring
    const int nplanes = xe - xs + 8;
    const int niter = xe - xs + 4;                    // iteration it: gx of plane q = xs-2+it, output plane q-2

    if (threadIdx.x == 0) {
        for (int s = 0; s < S; ++s) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(full + 8 * s), "r"(1) : "memory");
            done[s] = 0;
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        for (int n = 0; n < S && n < nplanes; ++n) {  // the first S planes: every slot is free
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(full + 8 * n), "r"(PLANE_B) : "memory");
            cl_tma_3d(ring + n * PLANE_B, &tmap, z0 - CL_H, yb - CL_H, p_first + n, full + 8 * n);
        }
    }
    __syncthreads();

    // hand slot `it % S` back; the last of the WY warps refills it with plane it + S
    auto release = [&](int it) {
        __syncwarp();
        if (lane == 0) {
            const int s = it % S;
            __threadfence_block();
            if (atomicAdd(&done[s], 1) == WY - 1) {
                done[s] = 0;
                const int n = it + S;
                if (n < nplanes) {
                    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(full + 8 * s), "r"(PLANE_B) : "memory");
                    cl_tma_3d(ring + s * PLANE_B, &tmap, z0 - CL_H, yb - CL_H, p_first + n, full + 8 * s);
                }
            }
        }
    };

    const int y0 = yb + warp * R;
    if (y0 >= ny) {                                    // warp outside the grid (ragged last y tile): keep pace, hand the slots back
        for (int it = 0; it < niter; ++it) {
            cl_wait(full + 8 * ((it + 4) % S), ((it + 4) / S) & 1);
            if (YS && it >= 4) asm volatile("bar.sync 1, %0;" ::"n"(32 * WY) : "memory");   // (the CTA's barrier of cl_output)
            release(it);
        }
        return;
    }
    const int z = z0 + 2 * lane;
    const unsigned tile_off = (unsigned)((warp * R + CL_H) * CL_PITCH + CL_H) * 8;      // (row y0, column z0) inside a slot
    const unsigned my_off = tile_off + 16u * lane;
    const i64 sx = (i64)ny * nz;

    for (int n = 0; n < 4; ++n) cl_wait(full + 8 * n, 0);           // planes xs-4 .. xs-1

    // gx window of the planes q-4 .. q: shifted by register moves (80 MOVs against ~1700 instructions of work
    // per step) rather than by unrolling the step five times, which would not fit the instruction cache
    ClRows<R> g0, g1, g2, g3, g4;
#pragma unroll
    for (int r = 0; r < R; ++r) g1.v[r] = g2.v[r] = g3.v[r] = g4.v[r] = make_double2(0.0, 0.0);
#pragma unroll 1
    for (int it = 0; it < niter; ++it) {
        // one iteration: plane q+2 lands -> gx(q) -> output plane q-2 -> slot of plane q-2 goes back
        const int q = xs - 2 + it;
        g0 = g1; g1 = g2; g2 = g3; g3 = g4;
        cl_wait(full + 8 * ((it + 4) % S), ((it + 4) / S) & 1);
        if (q >= 2 && q < nx - 2) {
            const unsigned a0 = ring + (it % S) * PLANE_B + my_off;
            const unsigned a1 = ring + ((it + 1) % S) * PLANE_B + my_off;
            const unsigned a2 = ring + ((it + 2) % S) * PLANE_B + my_off;
            const unsigned a3 = ring + ((it + 3) % S) * PLANE_B + my_off;
            const unsigned a4 = ring + ((it + 4) % S) * PLANE_B + my_off;
#pragma unroll
            for (int r = 0; r < R; ++r)
                g4.v[r] = cl_finish_xy<GEN>(cl_sum5(T.a[0].wc, cl_lds2(a0 + r * PB), cl_lds2(a1 + r * PB), cl_lds2(a2 + r * PB),
                                                    cl_lds2(a3 + r * PB), cl_lds2(a4 + r * PB)),
                                            T.a[0].ih, G, 0, q, true, q, y0 + r, z);
        } else if (q >= 0 && q < nx) {
#pragma unroll
            for (int r = 0; r < R; ++r)
                g4.v[r] = cl_gx_global(P, GEN, in + q * sx + (i64)(y0 + r) * nz + z, sx, q, nx, y0 + r, z);
        } else {
#pragma unroll
            for (int r = 0; r < R; ++r) g4.v[r] = make_double2(0.0, 0.0);
        }
        if (it >= 4) {                                  // the first four iterations only fill the window
            const unsigned slot = ring + (it % S) * PLANE_B;
            cl_output<R, GEN, YS, WY>(q - 2, y0, z0, lane, nx, ny, nz, T, G, P, in, out, slot + my_off, slot + tile_off, g0, g1, g2,
                                      g3, g4, gy_base + (it & 1) * GY_B, warp, yb);
        }
        release(it);
    }
}

bool cl_tables_ok(const FdTables& T) {
    if (T.n_center != 5 || T.n_side != 5 || T.nb != 2) return false;
    for (int k = 0; k < 5; ++k)
        if (T.oc[k] != k - 2 || T.of[k] != k || T.ob[k] != -k) return false;
    return true;
}

typedef CUresult (*cl_encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                 const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                 CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int cl_make_tmap(CUtensorMap* tm, const double* in, i64 nx, i64 ny, i64 nz, int rows) {
    static std::atomic<cl_encode_fn> cached{nullptr};
    cl_encode_fn encode = cached.load(std::memory_order_acquire);
    if (!encode) {
        void* fp = nullptr;
        cudaDriverEntryPointQueryResult qres;
        NG_CUDA_OK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &qres));
        if (!fp || qres != cudaDriverEntryPointSuccess) {
            ng_set_error("cuTensorMapEncodeTiled is not available from this driver");
            return NG_ECUDA;
        }
        encode = (cl_encode_fn)fp;
        cached.store(encode, std::memory_order_release);
    }
    const cuuint64_t gdim[3] = {(cuuint64_t)nz, (cuuint64_t)ny, (cuuint64_t)nx};
    const cuuint64_t gstr[2] = {(cuuint64_t)nz * 8, (cuuint64_t)ny * nz * 8};
    const cuuint32_t box[3] = {(cuuint32_t)CL_PITCH, (cuuint32_t)rows, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    const CUresult r = encode(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, const_cast<double*>(in), gdim, gstr, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        ng_set_error("cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
        return NG_ECUDA;
    }
    return NG_OK;
}

template <int R, int WY, int S, bool YS = false>
int cl_launch(ng_plan* p, const double* in, double* out, const ClParams& hp, bool gen, cudaStream_t st) {
    constexpr int TY = R * WY;
    const i64 nx = p->shape[0], ny = p->shape[1], nz = p->shape[2];
    CUtensorMap tm;
    int rc = cl_make_tmap(&tm, in, nx, ny, nz, TY + 2 * CL_H);
    if (!rc) rc = sync_mbar_hint(st);
    if (rc) return rc;
    // x chunks: 8 planes of every chunk are loaded for the window only, yet SHORT chunks win (B200, 512^3:
    // 16 planes 1.58 ms, 64 planes 1.83 ms, 256 planes 3.1 ms, profiles/r02_tune_curvlap.txt): a warp issues
    // one instruction per ~5 cycles (dependent FP64 chains, two warps per scheduler), so the machine only
    // fills when many CTAs at different points of their march share an SM over time
    int xchunk = (int)NG_ENV_INT("NG_CURV_XCHUNK", 0);
    if (xchunk < 1) xchunk = 16;
    if (xchunk > nx) xchunk = (int)nx;
    dim3 grid((unsigned)(nz / CL_TZ), (unsigned)((ny + TY - 1) / TY), (unsigned)((nx + xchunk - 1) / xchunk));
    const size_t smem = (size_t)S * (TY + 2 * CL_H) * CL_PITCH * 8 + 8 * S + 4 * S + 128 + (YS ? 2 * (TY + 4) * 512 + 16 : 0);
    if (gen) {
        NG_FUNC_SMEM(smem, curvlap3d_kernel<R, WY, S, true, YS>);
        curvlap3d_kernel<R, WY, S, true, YS><<<grid, 32 * WY, smem, st>>>(out, in, (int)nx, (int)ny, (int)nz, xchunk, hp.T, hp.G,
                                                                            (const ClParams*)p->d_curvlap, tm);
        NG_TRACE("curvlap3d<gen>");
    } else {
        NG_FUNC_SMEM(smem, curvlap3d_kernel<R, WY, S, false, YS>);
        curvlap3d_kernel<R, WY, S, false, YS><<<grid, 32 * WY, smem, st>>>(out, in, (int)nx, (int)ny, (int)nz, xchunk, hp.T, hp.G,
                                                                             (const ClParams*)p->d_curvlap, tm);
        NG_TRACE("curvlap3d<unit>");
    }
    NG_LAUNCH_CHECK();
    return NG_OK;
}

}  // namespace

// The single-pass kernel covers a PK_CURV plan when every axis is a first-derivative finite-difference plan
// with the acc = 4 tables and the shape fits the tiling; *handled stays false otherwise (six-pass path).
int ng_curvlap3d_try_launch(ng_plan* p, const double* in, double* out, cudaStream_t st, bool* handled) {
    *handled = false;
    if (p->ndim != 3 || NG_ENV_INT("NG_CURV_FUSED", 1) == 0) return NG_OK;
    const i64 nx = p->shape[0], ny = p->shape[1], nz = p->shape[2];
    if (nz % CL_TZ != 0 || ny % 4 != 0 || ny < 16 || nx < 16 || ny * nz >= (1LL << 28) || nx * ny * nz >= (1LL << 40)) return NG_OK;
    if ((((uintptr_t)in | (uintptr_t)out) & 15) != 0) return NG_OK;
    for (int a = 0; a < 3; ++a) {
        const ng_plan* q = p->axis_plans[a];
        if (!q || q->kind != PK_FD || !cl_tables_ok(q->fd)) return NG_OK;
    }
    ClParams hp;
    memset(&hp, 0, sizeof(hp));
    bool gen = false;
    for (int a = 0; a < 3; ++a) {
        const FdTables& F = p->axis_plans[a]->fd;
        for (int k = 0; k < 5; ++k) { hp.T.a[a].wc[k] = F.wc[k]; hp.T.a[a].wf[k] = F.wf[k]; hp.T.a[a].wb[k] = F.wb[k]; }
        hp.T.a[a].ih = F.inv_h_pow;
        hp.G.xdiv[a] = p->axis_plans[a]->d_xdiv;
        const CoefHost& c = p->cLap[a];
        if (!c.is_one) {
            hp.G.c[a].ptr = c.d_ptr;
            for (int b = 0; b < NG_MAX_DIM; ++b) hp.G.c[a].s[b] = c.stride[b];
        }
        gen = gen || hp.G.xdiv[a] || hp.G.c[a].ptr;
    }
    if (!p->cJ.is_one) {
        hp.G.J.ptr = p->cJ.d_ptr;
        for (int b = 0; b < NG_MAX_DIM; ++b) hp.G.J.s[b] = p->cJ.stride[b];
        gen = true;
    }
    if (!p->d_curvlap) {                            // device copy of the tables for the boundary routine (once per plan)
        int rc = ng_plan_bind_device(p);
        if (rc) return rc;
        NG_CUDA_OK(cudaMalloc((void**)&p->d_curvlap, sizeof(ClParams)));
        NG_CUDA_OK(cudaMemcpyAsync(p->d_curvlap, &hp, sizeof(ClParams), cudaMemcpyHostToDevice, st));
    }
    int rc;
    // tuning: tile rows / ring depth.  Two rows per warp recompute more of the y inner derivative ((R+4)/R rows per own
    // row) but put four warps on every scheduler instead of two, and the kernel is latency-bound (dependent FP64 chains):
    // 512^3 unit scale 1.45 ms against 1.64 ms (profiles/r02_tune_curvlap.txt)
    switch ((int)NG_ENV_INT("NG_CURV_CFG", 4)) {
        case 0: rc = cl_launch<4, 4, 8>(p, in, out, hp, gen, st); break;      // 16 x 64 tiles, two CTAs of 4 warps per SM
        case 1: rc = cl_launch<4, 8, 9>(p, in, out, hp, gen, st); break;      // 32 x 64 tiles, one CTA of 8 warps per SM
        case 2: rc = cl_launch<2, 8, 8>(p, in, out, hp, gen, st); break;      // 16 x 64 tiles, 2 rows per warp: two CTAs of 8 warps
        case 8: rc = cl_launch<2, 8, 6, true>(p, in, out, hp, gen, st); break;   // ... with the y inner derivatives shared through smem
        // the same with a ring of 6 planes (one ahead of the 5-plane window): 512^3 1.46 -> 1.36 ms -- the chunks are
        // 16 planes long, so a deeper ring buys nothing, while 55 KB less shared memory per SM is 55 KB more L1 for the
        // out-of-line boundary routine and the few spills (profiles/r02b_time_curvlap.txt)
        default:
            // general metrics (coefficient loads, divides in every inner derivative): the y inner derivatives shared through
            // shared memory, one CTA barrier per plane (2.48 -> 2.35 ms at 512^3); unit scale: recomputed per warp -- there
            // the barrier costs more than the 20 FP64 operations per row it saves (1.36 vs 1.70 ms)
            rc = gen ? cl_launch<2, 8, 6, true>(p, in, out, hp, gen, st) : cl_launch<2, 8, 6>(p, in, out, hp, gen, st);
            break;
    }
    if (rc) return rc;
    *handled = true;
    return NG_OK;
}
