/* numgrids_b200.h -- C ABI of libnumgrids_b200.so (sm_100a CUDA kernels for the numgrids hot path).
 *
 * The reference (maroba/numgrids, pure Python) has no FFI; its seam for this path is the
 * GridDiff protocol "subclasses set self.operator to a callable NDArray -> NDArray"
 * (reference numgrids/diff.py:26-28, :48-61) reached through Axis.create_diff_operator
 * (numgrids/axes.py:124-140), plus the four CurvilinearGrid methods
 * (numgrids/curvilinear.py:153-317).  Each entry point below replaces the arithmetic behind one
 * of those callables; the reference lines it replaces are cited per function.  INTEGRATION.md
 * shows the ctypes binding a maintainer would add.
 *
 * Conventions
 *   - all arrays are float64, C-contiguous, shape == grid shape (1 <= ndim <= NG_MAX_DIM);
 *   - "d_" pointers are device pointers, "h_" pointers are host pointers;
 *   - every function returns 0 on success, else an NG_E* code; ng_last_error() gives the text
 *     (thread-local);
 *   - plans are immutable after creation (ng_curv_set_coef excepted, before first use), own only
 *     their device tables/workspace and never allocate user-visible arrays;
 *   - device calls are asynchronous on `stream` (a cudaStream_t passed as void*; NULL = default
 *     stream); the *_host calls stage H2D/D2H themselves and return after the result is in h_out;
 *   - all coefficient tables (FD weights from numpy.linalg.solve, Chebyshev matrix entries,
 *     FFT multipliers, scale-factor tables) are computed by the host layer exactly as the
 *     reference computes them and passed in; the library never recomputes them.
 *   - no CPU fallback: without a usable CUDA device every compute call fails with NG_ECUDA.
 */
#ifndef NUMGRIDS_B200_H
#define NUMGRIDS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NG_MAX_DIM 4
#define NG_MAX_TAPS 16
#define NG_ABI_VERSION 1

enum ng_status {
    NG_OK = 0,
    NG_EINVAL = 1,       /* bad argument */
    NG_ECUDA = 2,        /* CUDA runtime error / no device */
    NG_ENOMEM = 3,
    NG_EUNSUPPORTED = 4
};

typedef struct ng_plan ng_plan; /* opaque */

/* A broadcastable coefficient array: element (i0..i3) lives at ptr[sum_a i_a*stride[a]];
 * stride 0 = constant along that axis.  ptr == NULL means "absent". */
typedef struct ng_coef {
    const double* ptr;
    int64_t stride[NG_MAX_DIM];
} ng_coef;

/* Elementwise work fused around one per-axis apply D.  Evaluation order (each step IEEE
 * round-to-nearest, never contracted), steps with absent operands skipped:
 *     x   = pre * in                      (before the derivative; `coeff * v[i]`, curvilinear.py:212-213)
 *     d   = D(x)
 *     d   = minuend - d                   (curl: D_j(..) - D_k(..), curvilinear.py:296-298, :312-314)
 *     d   = post * d                      ((1.0/h_i)*D f, (J/h_i**2)*D f, 1/(h_j h_k)*(..))
 *     d   = addend + d   |  0.0 + d       (result = result + D_i(..), curvilinear.py:213, :242)
 *     d   = d / divisor                   (result / J, curvilinear.py:214, :243)
 *     d   = isfinite(d) ? d : 0.0         (np.where(np.isfinite(..)), curvilinear.py:175, :215, :244)
 * minuend/addend are full-shape device arrays and may alias `out` (never `in`). */
typedef struct ng_fuse {
    ng_coef pre;
    const double* minuend;
    ng_coef post;
    const double* addend;
    int32_t add_zero;
    ng_coef divisor;
    int32_t finite;
} ng_fuse;

int ng_abi_version(void);
const char* ng_last_error(void);
/* number of visible CUDA devices (0 without a driver/GPU; never fails) */
int ng_device_count(void);
/* number of kernels this library launched since load (for bench.py's gpu_launches) */
int64_t ng_launch_count(void);
/* name of the kernel instance the calling thread launched last ("" before the first launch): lets a test
 * assert that the instance it means to check against the oracle is the one that ran */
const char* ng_last_kernel(void);
/* Tuning / A-B switches (NG_* environment variables) are read once per call site and cached; call this
 * after changing the environment to make the library read them again.  Not needed in production. */
void ng_tuning_reload(void);

/* ---- finite differences: FiniteDifferenceDiff / LogDiff ------------------------------------
 * Replaces findiff.FinDiff(axis, h, order, acc)(f) behind numgrids/diff.py:88 and :259-262.
 * Three schemes (central on [nb, n-nb), forward on [0, nb), backward on [n-nb, n), nb = n_center/2),
 * each out = (sum_k w[k]*in[i+off[k]], accumulated in stored order from 0.0) * inv_h_pow.
 * h_x_div (length shape[axis]) or NULL: LogDiff's `/ meshed_coords[axis]` (diff.py:261-262). */
int ng_fd_plan_create(int ndim, const int64_t* shape, int axis,
                      int n_center, const double* w_center, const int32_t* off_center,
                      int n_side, const double* w_fwd, const int32_t* off_fwd,
                      const double* w_bwd, const int32_t* off_bwd,
                      double inv_h_pow, const double* h_x_div, ng_plan** out);

/* ---- Chebyshev: ChebyshevDiff ---------------------------------------------------------------
 * Replaces `self._diff_matrix * f.reshape(-1)` (numgrids/diff.py:194-197): out_i = sum_j M[i][j] in_j
 * along `axis`, accumulated sequentially from 0.0 with separately rounded multiply and add,
 * j ascending (descending_k = 0) or descending (1) -- the order scipy's matvec uses on the
 * reference's Kronecker matrix.  h_M is the n x n row-major 1-D matrix. */
int ng_cheb_plan_create(int ndim, const int64_t* shape, int axis,
                        const double* h_M, int descending_k, ng_plan** out);

/* ---- periodic spectral derivative: FFTDiff --------------------------------------------------
 * Replaces real(ifft(W * fft(f, axis), axis)) (numgrids/diff.py:133-143).  h_W_re_im holds the n
 * complex multipliers of diff.py:117-125 interleaved (re, im).  Power-of-two n (<= 8192): in-shared-memory / register
 * FFT kernels.  Other n with h_D == NULL: the native n-point mixed-radix transform (n = 2^a 3^b 5^c 7^d 11^e 13^f <= 8192,
 * any derivative order, numpy.fft's O(n log n) and rounding class; another prime factor -> NG_EINVAL).  Other n with
 * h_D given (n x n row-major real matrix F^-1 diag(W) F, the matrix of diff.py:145-157): dense ordered contraction. */
int ng_fft_plan_create(int ndim, const int64_t* shape, int axis,
                       const double* h_W_re_im, const double* h_D, ng_plan** out);

/* The same for an axis of ANY length n in O(n log n): the operator is a circular convolution with the real kernel
 * d = ifft(W); the host lays d out on a circle of m points (m a power of two, 2n-1 <= m <= 8192: d[k] at k, d[n-k] at
 * m-k, zeros between) and passes its m-point transform h_Wm_re_im (m complex values, interleaved).  The kernels transform
 * zero-padded pencils with m points and store the first n outputs; arrays keep n samples along `axis`.  Rounding differs
 * from numpy.fft's n-point transform by a few 1e-14 for first derivatives; higher orders amplify it (the host layer
 * keeps the dense contraction there). */
int ng_fft_plan_create_padded(int ndim, const int64_t* shape, int axis, int64_t m, const double* h_Wm_re_im,
                              ng_plan** out);

/* The radices of the native mixed-radix transform of a non-power-of-two length n (at most 12, each from
 * {2..16} \ {primes above 13}, product n, fewest passes; powers of two first): what ng_fft_plan_create(..., h_D = NULL)
 * runs for this n.  NG_EINVAL if n has no such plan (a power of two, a prime factor above 13, n > 8192).  Host only:
 * needs no device. */
int ng_fft_mixed_radices(int64_t n, int* radices12, int* count);

/* ---- apply ------------------------------------------------------------------------------------
 * GridDiff.__call__ (numgrids/diff.py:48-61).  d_out must not alias d_in. */
int ng_apply(const ng_plan* plan, const double* d_in, double* d_out, void* stream);
int ng_apply_fused(const ng_plan* plan, const double* d_in, double* d_out,
                   const ng_fuse* fuse, void* stream);
/* `count` fields of the plan's shape stacked contiguously ([count][shape...], in and out): one launch for all of
 * them.  Small grids are launch-bound one field at a time; a stack streams at the large-array rate. */
int ng_apply_batch(const ng_plan* plan, const double* d_in, double* d_out, int64_t count, void* stream);
/* Same from host buffers (pinned or pageable): H2D, kernel, D2H, synchronous.  The device staging buffers (one input,
 * one output array) stay with the plan for the next call; ng_plan_release_staging frees them. */
int ng_apply_host(ng_plan* plan, const double* h_in, double* h_out);
int ng_plan_release_staging(ng_plan* plan);
/* Slab variant for a finite-difference plan whose `axis` is sharded across ranks: d_lo_halo /
 * d_hi_halo hold the nb neighbour planes [rows][nb] (rows = product of the other extents) or are
 * NULL on the global boundary, where the one-sided schemes are used. */
int ng_fd_apply_slab(const ng_plan* plan, const double* d_in, double* d_out,
                     const double* d_lo_halo, const double* d_hi_halo, void* stream);
/* Pack the first / last `nb` planes along the plan's axis into contiguous [rows][nb] buffers. */
int ng_fd_pack_halo(const ng_plan* plan, const double* d_in, double* d_lo_planes,
                    double* d_hi_planes, void* stream);
int ng_plan_halo_width(const ng_plan* plan);
/* The same with the elementwise work of ng_fuse (slab form of the curvilinear composites,
 * numgrids/curvilinear.py:169-317, when the sharded axis is a finite-difference axis).  The
 * halo planes must already hold `pre * value` (ng_fd_pack_halo_fused on the sending rank, whose
 * coefficient at a shared point is the same number); all other fuse operands are local. */
int ng_apply_fused_slab(const ng_plan* plan, const double* d_in, double* d_out, const ng_fuse* fuse,
                        const double* d_lo_halo, const double* d_hi_halo, void* stream);
int ng_fd_pack_halo_fused(const ng_plan* plan, const double* d_in, const ng_coef* pre,
                          double* d_lo_planes, double* d_hi_planes, void* stream);

/* ---- slab <-> pencil redistribution (multi-GPU; no counterpart in the single-process reference)
 * A Chebyshev / FFT derivative along the sharded last axis needs whole pencils: a rank's slab
 * [A][B][zl] is redistributed to [A_r][B][Z], the plan applied there, and the result sent back.
 * Each direction is one launch over <= 16 strided 2-D blocks (rows = contiguous z segments);
 * src or dst of a block may be peer-GPU memory (NVLink P2P), so copy and exchange are one kernel.
 * fuse_side 0: plain copy; 1: values are multiplied by fuse->pre as they are read from the local
 * slab; 2: the tail of ng_fuse (minuend .. finite) is applied as values are written into the local
 * slab.  local_offset is the linear index, in the rank's own slab array of shape local_shape, of
 * the block's first element on its LOCAL side (src for side 1, dst for side 2). */
typedef struct ng_copy_block {
    const double* src;
    double* dst;
    int64_t rows, row_len;
    int64_t src_stride, dst_stride;   /* in elements */
    int64_t local_offset;
} ng_copy_block;
int ng_slab_block_copy(int nblocks, const ng_copy_block* blocks, int ndim, const int64_t* local_shape,
                       const ng_fuse* fuse, int fuse_side, void* stream);

/* ---- fused Cartesian Laplacian ----------------------------------------------------------------
 * The reference idiom Diff(g,2,0)(f) + Diff(g,2,1)(f) + ... on equidistant axes (reference
 * tests/test_diff.py:29-39, docs/guide/differentiation.md) in ONE pass: per-axis sums as in
 * ng_fd_plan_create, combined as ((d0 + d1) + d2)...; axis_plans are ndim FD plans (one per axis,
 * same shape).  Slab variant: last axis sharded, halos as in ng_fd_apply_slab. */
int ng_fdlap_plan_create(int ndim, const int64_t* shape, ng_plan* const* axis_plans, ng_plan** out);
int ng_fdlap_apply(const ng_plan* plan, const double* d_in, double* d_out, void* stream);
int ng_fdlap_apply_slab(const ng_plan* plan, const double* d_in, double* d_out,
                        const double* d_lo_halo, const double* d_hi_halo, void* stream);
/* Output planes [x_begin, x_end) of axis 0 only (the input must be complete; halos are needed only
 * for those planes).  Lets the host overlap the halo exchange of one plane range with the stencil
 * of another. */
int ng_fdlap_apply_slab_range(const ng_plan* plan, const double* d_in, double* d_out,
                              const double* d_lo_halo, const double* d_hi_halo, int64_t x_begin,
                              int64_t x_end, void* stream);
/* Halo/compute overlap along the sharded axis: z_part 1 computes the part of the slab whose
 * stencils never touch a neighbour rank's planes (it can run while the exchange is in flight,
 * the halo pointers are not read), z_part 2 the remaining boundary part, 0 everything.  Parts 1
 * and 2 together always cover the slab exactly once. */
int ng_fdlap_apply_slab_part(const ng_plan* plan, const double* d_in, double* d_out,
                             const double* d_lo_halo, const double* d_hi_halo, int z_part, void* stream);
int ng_fdlap_apply_host(ng_plan* plan, const double* h_in, double* h_out);
/* ng_fd_pack_halo for the planes [x_begin, x_end) of axis 0 only (rows x_begin*R .. x_end*R of the [rows][nb]
 * buffers, R = rows per plane): lets a chunked host pipeline exchange halos chunk by chunk. */
int ng_fdlap_pack_halo_range(const ng_plan* plan, const double* d_in, double* d_lo_planes,
                             double* d_hi_planes, int64_t x_begin, int64_t x_end, void* stream);

/* ---- one multi-GPU step of the fused Cartesian Laplacian (one process per GPU, last axis sharded) ------------------
 * pack kernel (stores this rank's first / last nb columns straight into the neighbours' halo buffers) -> one 8-byte flag
 * per neighbour -> interior z tiles on `stream` || boundary z tiles on the plan's side stream once both neighbours'
 * flags have arrived -> join on `stream`.  The library maps no peer memory: every pointer below must be valid in THIS
 * process -- the neighbours' buffers through CUDA IPC, fabric handles, or torch symmetric memory (numgrids_b200/slab.py).
 * Halo buffers hold [rows][nb] doubles (rows = product of the unsharded local extents), two sets that alternate per
 * step.  Flags are step counters stored with release / read with acquire at system scope; they must start at 0. */
typedef struct ng_slab_comm {
    int32_t rank, nranks;
    double* my_halo[2][2];        /* [set][0 = from the lower neighbour, 1 = from the upper one]: this rank's buffers */
    double* peer_lo_halo[2];      /* [set] the LOWER neighbour's my_halo[set][1], mapped here (NULL on rank 0)         */
    double* peer_hi_halo[2];      /* [set] the UPPER neighbour's my_halo[set][0], mapped here (NULL on the last rank)  */
    uint64_t* my_flags;           /* 2 words in this rank's memory: [0] written by the lower, [1] by the upper neighbour */
    uint64_t* peer_lo_flag;       /* the lower neighbour's my_flags + 1, mapped here                                   */
    uint64_t* peer_hi_flag;       /* the upper neighbour's my_flags + 0, mapped here                                   */
} ng_slab_comm;
/* local_shape = this rank's slab; axis_plans = the ndim second-derivative FD plans of that shape (global spacings). */
int ng_slab_plan_create(const ng_slab_comm* comm, int ndim, const int64_t* local_shape, ng_plan* const* axis_plans,
                        ng_plan** out);
int ng_slab_lap_apply(ng_plan* plan, const double* d_f, double* d_out, void* stream);
/* 1 after a wait gave up (~10 s without a neighbour's flag: crashed rank, mismatched step counts); synchronises. */
int ng_slab_timed_out(ng_plan* plan);

/* ---- curvilinear composites: CurvilinearGrid.gradient/divergence/laplacian/curl ---------------
 * numgrids/curvilinear.py:153-317.  axis_plans[i] is the first-derivative plan of axis i
 * (curvilinear.py:142-149).  Coefficients are evaluated on the host exactly as the reference does
 * (J/h_i**2, J/h_i, 1.0/h_i, h_i, 1.0/(h_j*h_k), J) and registered with ng_curv_set_coef in
 * broadcast-compact form (`count` doubles at h_values, element strides per axis, 0 = broadcast). */
enum ng_coef_kind {
    NG_COEF_J = 0,        /* index ignored */
    NG_COEF_LAP = 1,      /* J / h_i**2        index i */
    NG_COEF_DIV = 2,      /* J / h_i           index i */
    NG_COEF_GRAD = 3,     /* 1.0 / h_i         index i */
    NG_COEF_H = 4,        /* h_i               index i */
    NG_COEF_CURL = 5      /* 1.0 / (h_j * h_k) index = output component (0 for 2-D) */
};
int ng_curv_plan_create(int ndim, const int64_t* shape, ng_plan* const* axis_plans, ng_plan** out);
int ng_curv_set_coef(ng_plan* plan, int kind, int index, const double* h_values, int64_t count,
                     const int64_t* stride);
/* Optional: sq = a plan of the same shape and axis whose multiplier is the SQUARE of axis_plans[axis]'s (a periodic
 * axis: D applied twice as one transform).  ng_curv_laplacian then evaluates D(c D f) as c * (D D f) on that axis
 * whenever the registered J/h_axis**2 table does not vary along it (numgrids/curvilinear.py:235-244 applies D twice;
 * for such c the two are the same operator, and the single transform carries less rounding).  The plan must outlive
 * the curvilinear plan. */
int ng_curv_set_axis_square(ng_plan* plan, int axis, const ng_plan* sq);
int ng_curv_laplacian(ng_plan* plan, const double* d_f, double* d_out, void* stream);
int ng_curv_gradient(ng_plan* plan, const double* d_f, double* const* d_out, void* stream);
int ng_curv_divergence(ng_plan* plan, const double* const* d_v, double* d_out, void* stream);
int ng_curv_curl(ng_plan* plan, const double* const* d_v, double* const* d_out, void* stream);

/* ---- array-level boundary conditions (numgrids/boundary.py:157-166, :216-245) ------------------
 * In place on a device array of the given shape, so that a time-stepping loop keeps its field on
 * the device between derivative applies.  The face is the array with `axis` removed (C order).
 *   NG_BC_DIRICHLET  u[face] = g
 *   NG_BC_NEUMANN    u[face] = u[adjacent layer] + hs * g,  hs = h * normal_sign from the host
 *                    (the reference's first-order form: h = |x[1]-x[0]| at the face, sign -1 low / +1 high)
 * d_g: device array of face values, or NULL to use g_scalar everywhere. */
enum ng_bc_kind { NG_BC_DIRICHLET = 0, NG_BC_NEUMANN = 1 };
int ng_bc_apply(int ndim, const int64_t* shape, int axis, int side /* 0 low, 1 high */, int kind,
                double hs, const double* d_g, double g_scalar, double* d_u, void* stream);

/* ---- quadrature over the whole grid (numgrids/integration.py:44-72, Integral.__call__) ----------
 * out[0] = sum_{i0,i1,..} w_0[i0] w_1[i1] ... f[i0,i1,..]: one streaming pass (8 B/point).  d_w[a] is
 * the device table of shape[a] weights the host derives exactly as the reference does (periodic axis:
 * the spacing everywhere, integration.py:61-63; otherwise 0 followed by the last row of
 * inv(D[1:,1:]), integration.py:33-40, :65-70).  d_work: >= ng_quad_workspace_doubles() doubles of
 * device scratch; d_out: one device double.  Deterministic for a given shape (no atomics). */
int64_t ng_quad_workspace_doubles(void);
int ng_quad_apply(int ndim, const int64_t* shape, const double* const* d_w, const double* d_f,
                  double* d_work, double* d_out, void* stream);

/* ---- norm of a difference (numgrids/amr.py:96-103, ErrorEstimator._compute_norm) ----------------
 * out[0] = max|a-b| (NG_NORM_MAX), sqrt(mean((a-b)^2)) (NG_NORM_L2) or mean|a-b| (NG_NORM_MEAN) of two
 * device arrays of n doubles, one pass.  d_work: >= ng_quad_workspace_doubles() doubles of scratch. */
enum ng_norm_kind { NG_NORM_MAX = 0, NG_NORM_L2 = 1, NG_NORM_MEAN = 2 };
int ng_diff_norm(int64_t n, const double* d_a, const double* d_b, int kind, double* d_work,
                 double* d_out, void* stream);

/* ---- multilinear interpolation (numgrids/interpol.py:34-46, :60-75 with method="linear"; ---------
 * numgrids/grids.py:291-327 MultiGrid.transfer) --------------------------------------------------
 * d_out[t] = interpolant of d_values (C order, src_shape) at target point t.  scattered = 0: the
 * target is the tensor-product grid dst_shape and table a is indexed by the point's coordinate along
 * axis a (Interpolator.__call__(Grid)); scattered = 1: dst_shape[0] points, every table indexed by the
 * point number.  Tables per axis (host-built with the reference's NumPy expressions): d_idx = lower
 * interval index i, d_y = (xq-x[i])/(x[i+1]-x[i]), d_omy = 1-y; for ndim == 1 (numpy.interp's form)
 * d_y = xq-x[i] and d_omy = x[i+1]-x[i].  Arithmetic order per dimensionality: see csrc/interp.cu. */
int ng_interp_linear(int ndim, const int64_t* src_shape, const int64_t* dst_shape, int scattered,
                     const int32_t* const* d_idx, const double* const* d_y, const double* const* d_omy,
                     const double* d_values, double* d_out, void* stream);

/* Tensor-product B-spline evaluation (Interpolator methods "cubic" / "quadratic": SciPy's NdBSpline / BSpline
 * behind numgrids/interpol.py:34-46).  d_coef: the spline coefficients (C order, coef_shape) the host obtained
 * from SciPy exactly as the reference does; per axis a: nbasis[a] = degree + 1, d_first[a][t] = first coefficient
 * index and d_basis[a][t*nbasis[a] + j] = j-th non-zero basis value at target coordinate t (SciPy's de Boor
 * recursion, host-built).  Target layout and `scattered` as for ng_interp_linear.  Terms summed in SciPy's order. */
int ng_interp_bspline(int ndim, const int64_t* coef_shape, const int64_t* dst_shape, int scattered,
                      const int32_t* nbasis, const int32_t* const* d_first, const double* const* d_basis,
                      const double* d_coef, double* d_out, void* stream);

/* ---- lifetime ---------------------------------------------------------------------------------*/
int ng_plan_destroy(ng_plan* plan);
/* bytes of device memory owned by the plan (tables + workspace + host-call staging) */
int64_t ng_plan_device_bytes(const ng_plan* plan);

#ifdef __cplusplus
}
#endif
#endif /* NUMGRIDS_B200_H */
