// fdlap.cu -- fused Cartesian finite-difference Laplacian.
//
// The reference has no Cartesian `laplacian`; its idiom is
//     Diff(g,2,0)(f) + Diff(g,2,1)(f) + Diff(g,2,2)(f)
// (reference tests/test_diff.py:29-39), i.e. three findiff applies (numgrids/diff.py:88) and two
// full-array adds: 15 strided NumPy passes and ~8 temporaries.  Here it is ONE pass, 16 B/point:
//     d_a = (sum_k w_k f[i_a + o_k], stored tap order, from 0.0, DMUL + DADD) * inv_h2_a
//     out = ((d_0 + d_1) + d_2) ...
// bit-identical to the three-call idiom.  One-sided schemes on the first/last nb indices of every
// axis; in slab mode the last axis is sharded and its neighbour planes arrive as [rows][nb] halos.
//
// Kernels:
//   fdlap_generic_kernel   any ndim <= 4, one thread per point, taps through L1/L2.
//   fdlap3d_march_kernel   3-D production path: each thread owns R rows x 2 contiguous points
//                          (128-bit loads/stores), marches along axis 0 with a register window of
//                          2p+1 planes, takes the axis-1 halo rows from L1/L2 and the axis-2
//                          neighbours by warp shuffle.
#include "common.cuh"

namespace {

struct LapParams {
    FdTables t[NG_MAX_DIM];
    i64 shape[NG_MAX_DIM];
    i64 stride[NG_MAX_DIM];
    int ndim;
};

__global__ void __launch_bounds__(256)
fdlap_generic_kernel(const double* __restrict__ in, double* __restrict__ out,
                     const LapParams* __restrict__ Pp, i64 first, i64 total,
                     const double* __restrict__ lo, const double* __restrict__ hi) {
    const LapParams& P = *Pp;
    for (i64 lin = first + (i64)blockIdx.x * blockDim.x + threadIdx.x; lin < total;
         lin += (i64)gridDim.x * blockDim.x) {
        const Idx4 ix = unravel4(lin, P.shape);
        double res = 0.0;
        for (int a = 0; a < P.ndim; ++a) {
            const FdTables& T = P.t[a];
            const i64 i = ix.i[a], n = P.shape[a];
            const bool sharded = (a == P.ndim - 1);
            const bool has_lo = sharded && lo, has_hi = sharded && hi;
            const double* w; const int* off; int nt;
            if (i < T.nb && !has_lo) { w = T.wf; off = T.of; nt = T.n_side; }
            else if (i >= n - T.nb && !has_hi) { w = T.wb; off = T.ob; nt = T.n_side; }
            else { w = T.wc; off = T.oc; nt = T.n_center; }
            double acc = 0.0;
            for (int k = 0; k < nt; ++k) {
                const i64 ii = i + off[k];
                double v;
                if (ii < 0) v = lo[(lin / n) * T.nb + (T.nb + ii)];          // sharded => stride 1
                else if (ii >= n) v = hi[(lin / n) * T.nb + (ii - n)];
                else v = in[lin + (ii - i) * P.stride[a]];
                acc = __dadd_rn(acc, __dmul_rn(w[k], v));
            }
            const double d = __dmul_rn(acc, T.inv_h_pow);
            res = (a == 0) ? d : __dadd_rn(res, d);
        }
        out[lin] = res;
    }
}

}  // namespace

int ng_fdlap3d_try_launch(const ng_plan* p, const double* in, double* out, const double* lo,
                          const double* hi, int xbeg, int xend, int zmode, cudaStream_t st, bool* handled);

// Output planes [xbeg, xend) along axis 0 (the whole array when xbeg = 0, xend = shape[0]).
// zmode: 0 = all of the last axis; 1 = only the part that needs no slab halo; 2 = the rest.
int ng_fdlap_launch(const ng_plan* p, const double* in, double* out, const double* lo,
                    const double* hi, i64 xbeg, i64 xend, int zmode, cudaStream_t st) {
    if (p->size == 0 || xend <= xbeg) return NG_OK;
    int drc = ng_check_plan_device(p, "ng_fdlap_apply");
    if (drc) return drc;
    bool handled = false;
    int rc = ng_fdlap3d_try_launch(p, in, out, lo, hi, (int)xbeg, (int)xend, zmode, st, &handled);
    if (rc != NG_OK || handled) return rc;
    if (zmode == 1 || zmode == 3) return NG_OK;   // generic kernel: part 2 computes everything
    const i64 plane = p->size / p->shape[0];
    const i64 first = xbeg * plane, last = xend * plane;
    i64 g = (last - first + 255) / 256;
    if (g > 148LL * 64) g = 148LL * 64;
    fdlap_generic_kernel<<<(unsigned)g, 256, 0, st>>>(in, out, (const LapParams*)p->d_M, first, last,
                                                      lo, hi);
    NG_TRACE("fdlap_generic");
    NG_LAUNCH_CHECK();
    return NG_OK;
}

// Build the device-resident parameter block (stored in plan->d_M).
int ng_fdlap_upload_params(ng_plan* p) {
    LapParams hp;
    memset(&hp, 0, sizeof(hp));
    hp.ndim = p->ndim;
    i64 s = 1;
    for (int a = NG_MAX_DIM - 1; a >= 0; --a) {
        hp.shape[a] = p->shape[a];
        hp.stride[a] = s;
        s *= p->shape[a];
    }
    for (int a = 0; a < p->ndim; ++a) hp.t[a] = p->lap[a];
    int brc = ng_plan_bind_device(p);
    if (brc) return brc;
    NG_CUDA_OK(cudaMalloc((void**)&p->d_M, sizeof(LapParams)));
    p->device_bytes += sizeof(LapParams);
    NG_CUDA_OK(cudaMemcpy(p->d_M, &hp, sizeof(LapParams), cudaMemcpyHostToDevice));
    if (p->ndim == 3) {
        // one-sided tables of the 3-D march kernel (struct Lap3Side in fdlap3d.cu): wf[3][6], wb[3][6], ih[3]
        double side[39];
        for (int a = 0; a < 3; ++a)
            for (int k = 0; k < 6; ++k) {
                side[a * 6 + k] = p->lap[a].wf[k];
                side[18 + a * 6 + k] = p->lap[a].wb[k];
            }
        for (int a = 0; a < 3; ++a) side[36 + a] = p->lap[a].inv_h_pow;
        NG_CUDA_OK(cudaMalloc((void**)&p->d_Mt, sizeof(side)));
        p->device_bytes += sizeof(side);
        NG_CUDA_OK(cudaMemcpy(p->d_Mt, side, sizeof(side), cudaMemcpyHostToDevice));
    }
    return NG_OK;
}
