// transpose.cu -- block copies of the slab <-> pencil redistribution (multi-GPU only).
//
// The reference is single-process; this serves BASELINE.json's north_star: a Chebyshev or FFT
// derivative along the SHARDED (last) axis needs whole pencils, so a rank's slab [A][B][zl] is
// redistributed to [A_r][B][Z] (axis 0 partitioned instead), the plan is applied there, and the
// result comes back.  Both directions are sets of 2-D strided block copies whose rows are the
// contiguous z segments; the remote side of a block may be a peer GPU's memory (NVLink P2P
// stores on the way out, P2P loads on the way back), so copy and exchange are ONE kernel.
//
// The elementwise work of ng_fuse rides along on the LOCAL side of a block, indexed in the
// rank's own slab layout: `pre` multiplies values as they leave the slab (fuse_side 1), the tail
// (minuend / post / addend / divisor / finite) is applied as results arrive in it (fuse_side 2).
#include "common.cuh"

namespace {

constexpr int MAX_BLOCKS = 16;
struct CopyBlocks {
    ng_copy_block b[MAX_BLOCKS];
    int n;
};

template <int VEC, int SIDE>
__global__ void __launch_bounds__(256)
slab_block_copy_kernel(CopyBlocks B, FuseDev F) {
    const ng_copy_block& k = B.b[blockIdx.y];
    const i64 rl = k.row_len / VEC;
    const i64 total = k.rows * rl;
    const i64 lstride = SIDE == 2 ? k.dst_stride : k.src_stride;   // row stride on the local side
    for (i64 idx = (i64)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (i64)gridDim.x * blockDim.x) {
        i64 row, col;
        if (total < 0x7fffffffLL) {
            const unsigned r = (unsigned)idx / (unsigned)rl;
            row = r;
            col = ((unsigned)idx - r * (unsigned)rl) * VEC;
        } else {
            row = idx / rl;
            col = (idx - row * rl) * VEC;
        }
        const double* s = k.src + row * k.src_stride + col;
        double* d = k.dst + row * k.dst_stride + col;
        double v[VEC];
        if (VEC == 2) {
            const double2 t = *reinterpret_cast<const double2*>(s);
            v[0] = t.x;
            v[VEC - 1] = t.y;
        } else {
            v[0] = *s;
        }
        if (SIDE != 0) {
            const i64 lin = k.local_offset + row * lstride + col;
            Idx4 ix = unravel4(lin, F.shape);
#pragma unroll
            for (int j = 0; j < VEC; ++j) {
                Idx4 jx = ix;
                jx.i[F.vec_axis] += j;                       // pairs never straddle a z row
                if (SIDE == 1) { if (F.pre.ptr) v[j] = __dmul_rn(F.pre.ptr[coef_off(F.pre, jx)], v[j]); }
                else v[j] = fuse_tail(F, v[j], lin + j, jx);
            }
        }
        if (VEC == 2) *reinterpret_cast<double2*>(d) = make_double2(v[0], v[VEC - 1]);
        else *d = v[0];
    }
}

}  // namespace

extern "C" int ng_slab_block_copy(int nblocks, const ng_copy_block* blocks, int ndim,
                                  const int64_t* local_shape, const ng_fuse* fuse, int fuse_side,
                                  void* stream) {
    NG_REQUIRE(nblocks >= 0 && nblocks <= MAX_BLOCKS, "ng_slab_block_copy: at most %d blocks per call", MAX_BLOCKS);
    NG_REQUIRE(nblocks == 0 || blocks, "ng_slab_block_copy: NULL blocks");
    NG_REQUIRE(fuse_side >= 0 && fuse_side <= 2, "ng_slab_block_copy: fuse_side must be 0, 1 or 2");
    NG_REQUIRE(fuse_side == 0 || (fuse && local_shape && ndim >= 1 && ndim <= NG_MAX_DIM),
               "ng_slab_block_copy: fused copies need the fuse record and the local slab shape");
    CopyBlocks B;
    B.n = 0;
    bool vec2 = true;
    i64 most = 0;
    for (int i = 0; i < nblocks; ++i) {
        const ng_copy_block& k = blocks[i];
        NG_REQUIRE(k.rows >= 0 && k.row_len >= 0, "ng_slab_block_copy: negative extent in block %d", i);
        if (k.rows == 0 || k.row_len == 0) continue;
        NG_REQUIRE(k.src && k.dst, "ng_slab_block_copy: NULL pointer in block %d", i);
        NG_REQUIRE(k.src_stride >= k.row_len && k.dst_stride >= k.row_len,
                   "ng_slab_block_copy: rows of block %d overlap", i);
        vec2 = vec2 && ((k.row_len | k.src_stride | k.dst_stride | k.local_offset) % 2 == 0) &&
               ((((uintptr_t)k.src | (uintptr_t)k.dst) & 15) == 0);
        B.b[B.n++] = k;
        const i64 items = k.rows * k.row_len;
        if (items > most) most = items;
    }
    if (B.n == 0) return NG_OK;
    FuseDev F;
    memset(&F, 0, sizeof(F));
    for (int a = 0; a < NG_MAX_DIM; ++a) F.shape[a] = 1;
    if (fuse_side != 0) {
        ng_plan view;                                  // only the shape matters to ng_fuse_to_dev
        for (int a = 0; a < ndim; ++a) view.shape[a] = local_shape[a];
        ng_fuse_to_dev(&view, fuse, &F);
        if (F.shape[F.vec_axis] % 2 != 0) vec2 = false;
    }
    const i64 per = vec2 ? 2 : 1;
    i64 gx = (most / per + 255) / 256;
    const i64 cap = (148LL * 16 + B.n - 1) / B.n;      // ~16 CTAs per SM over all blocks, grid-stride beyond
    if (gx > cap) gx = cap;
    if (gx < 1) gx = 1;
    const dim3 grid((unsigned)gx, (unsigned)B.n);
    cudaStream_t st = (cudaStream_t)stream;
#define NG_BC_GO(V, S) slab_block_copy_kernel<V, S><<<grid, 256, 0, st>>>(B, F)
    if (vec2) { if (fuse_side == 0) NG_BC_GO(2, 0); else if (fuse_side == 1) NG_BC_GO(2, 1); else NG_BC_GO(2, 2); }
    else      { if (fuse_side == 0) NG_BC_GO(1, 0); else if (fuse_side == 1) NG_BC_GO(1, 1); else NG_BC_GO(1, 2); }
#undef NG_BC_GO
    NG_LAUNCH_CHECK();
    return NG_OK;
}
