// api.cu -- the C ABI of libnumgrids_b200.so (declared in include/numgrids_b200.h): plan creation,
// table upload, dispatch and host-buffer staging.  All validation that maps to Python exceptions
// in the reference stays in the Python layer; this file rejects only what would make a kernel
// read out of bounds.
#include <stdarg.h>
#include <stdlib.h>

#include <mutex>
#include <new>
#include <vector>

#include "common.cuh"

std::atomic<long long> g_ng_launches{0};
static thread_local char t_err[512] = "";

void ng_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(t_err, sizeof(t_err), fmt, ap);
    va_end(ap);
}

thread_local const char* t_ng_last_kernel = "";
static std::atomic<int> g_env_gen{0};

long long ng_env_cached(NgEnvCache* c, const char* name, long long dflt) {
    const int gen = g_env_gen.load(std::memory_order_acquire);
    if (c->gen.load(std::memory_order_acquire) == gen) return c->val.load(std::memory_order_relaxed);
    const char* s = getenv(name);
    const long long v = s ? atoll(s) : dflt;
    c->val.store(v, std::memory_order_relaxed);
    c->gen.store(gen, std::memory_order_release);
    return v;
}

int ng_current_device() {
    int d = 0;
    if (cudaGetDevice(&d) != cudaSuccess) { cudaGetLastError(); return 0; }
    return d;
}

// Device memory of a plan lives on the device that was current when its first table was uploaded.
int ng_plan_bind_device(ng_plan* p) {
    const int d = ng_current_device();
    if (p->device < 0) p->device = d;
    NG_REQUIRE(p->device == d, "plan tables live on CUDA device %d but the current device is %d", p->device, d);
    return NG_OK;
}

int ng_check_plan_device(const ng_plan* p, const char* who) {
    if (!p || p->device < 0) return NG_OK;
    const int d = ng_current_device();
    NG_REQUIRE(p->device == d, "%s: the plan was created on CUDA device %d but the current device is %d "
               "(create one plan per device)", who, p->device, d);
    return NG_OK;
}

int ng_fdlap_upload_params(ng_plan* p);

namespace {

int set_shape(ng_plan* p, int ndim, const int64_t* shape, int axis) {
    NG_REQUIRE(ndim >= 1 && ndim <= NG_MAX_DIM, "ndim must be in 1..%d, got %d", NG_MAX_DIM, ndim);
    NG_REQUIRE(shape != nullptr, "shape is NULL");
    NG_REQUIRE(axis >= 0 && axis < ndim, "axis %d out of range for ndim %d", axis, ndim);
    p->ndim = ndim;
    p->axis = axis;
    p->size = 1;
    for (int a = 0; a < NG_MAX_DIM; ++a) p->shape[a] = 1;
    for (int a = 0; a < ndim; ++a) {
        NG_REQUIRE(shape[a] >= 1, "shape[%d] = %lld must be positive", a, (long long)shape[a]);
        p->shape[a] = shape[a];
        p->size *= shape[a];
    }
    p->outer = 1;
    p->inner = 1;
    for (int a = 0; a < axis; ++a) p->outer *= shape[a];
    for (int a = axis + 1; a < ndim; ++a) p->inner *= shape[a];
    p->n = shape[axis];
    return NG_OK;
}

int upload(ng_plan* p, double** d_dst, const double* h_src, i64 count) {
    int rc = ng_plan_bind_device(p);
    if (rc) return rc;
    NG_CUDA_OK(cudaMalloc((void**)d_dst, sizeof(double) * (size_t)(count > 0 ? count : 1)));
    p->device_bytes += sizeof(double) * count;
    if (count > 0)
        NG_CUDA_OK(cudaMemcpy(*d_dst, h_src, sizeof(double) * (size_t)count, cudaMemcpyHostToDevice));
    return NG_OK;
}

// zero-padded k-major matrix for the general dense kernel (dense.cu): [n_pad][n_pad], n_pad = ceil128(n)
int upload_padded_mt(ng_plan* p, const double* Mt) {
    const i64 n = p->n, np = (n + 127) / 128 * 128;
    if (np > 8192) return NG_OK;                      // (the general kernel is then simply not used)
    std::vector<double> pad((size_t)(np * np), 0.0);
    for (i64 k = 0; k < n; ++k) memcpy(&pad[(size_t)(k * np)], Mt + k * n, sizeof(double) * (size_t)n);
    p->n_pad = (int)np;
    return upload(p, &p->d_Mtp, pad.data(), np * np);
}

int require_device() {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        cudaGetLastError();
        ng_set_error("no usable CUDA device (%s); numgrids_b200 has no CPU fallback",
                     e != cudaSuccess ? cudaGetErrorString(e) : "0 devices");
        return NG_ECUDA;
    }
    return NG_OK;
}

int ensure_stage(ng_plan* p, int nbuf_in) {
    const i64 need = p->size * nbuf_in;
    if (p->stage_elems >= need && p->d_stage_in && p->d_stage_out) return NG_OK;
    int rc = ng_plan_bind_device(p);
    if (rc) return rc;
    if (p->d_stage_in) cudaFree(p->d_stage_in);
    if (p->d_stage_out) cudaFree(p->d_stage_out);
    p->d_stage_in = p->d_stage_out = nullptr;
    NG_CUDA_OK(cudaMalloc((void**)&p->d_stage_in, sizeof(double) * (size_t)need));
    NG_CUDA_OK(cudaMalloc((void**)&p->d_stage_out, sizeof(double) * (size_t)p->size));
    p->device_bytes += sizeof(double) * (need + p->size - (p->stage_elems > 0 ? p->stage_elems + p->size : 0));
    p->stage_elems = need;
    return NG_OK;
}

FuseDev no_fuse(const ng_plan* p) {
    FuseDev F;
    memset(&F, 0, sizeof(F));
    for (int a = 0; a < NG_MAX_DIM; ++a) F.shape[a] = p->shape[a];
    F.vec_axis = 0;
    for (int a = 0; a < NG_MAX_DIM; ++a) if (p->shape[a] > 1) F.vec_axis = a;
    return F;
}

}  // namespace

void ng_fuse_to_dev(const ng_plan* p, const ng_fuse* f, FuseDev* out) {
    *out = no_fuse(p);
    if (!f) return;
    auto cp = [](CoefDev& d, const ng_coef& s) {
        d.ptr = s.ptr;
        for (int a = 0; a < NG_MAX_DIM; ++a) d.s[a] = s.ptr ? s.stride[a] : 0;
    };
    cp(out->pre, f->pre);
    cp(out->post, f->post);
    cp(out->div, f->divisor);
    out->minuend = f->minuend;
    out->addend = f->addend;
    out->add_zero = f->add_zero;
    out->finite = f->finite;
    out->any = (f->pre.ptr || f->post.ptr || f->divisor.ptr || f->minuend || f->addend ||
                f->add_zero || f->finite) ? 1 : 0;
}

int ng_apply_any(const ng_plan* p, const double* in, double* out, const FuseDev& F,
                 cudaStream_t st) {
    int rc = ng_check_plan_device(p, "apply");
    if (rc) return rc;
    switch (p->kind) {
        case PK_FD: return ng_fd_launch(p, in, out, F, nullptr, nullptr, st);
        case PK_CHEB: return ng_dense_launch(p, in, out, F, st);
        case PK_FFT: return ng_fft_launch(p, in, out, F, st);
        default:
            ng_set_error("plan kind %d cannot be applied with ng_apply", p->kind);
            return NG_EINVAL;
    }
}

extern "C" {

int ng_abi_version(void) { return NG_ABI_VERSION; }
const char* ng_last_error(void) { return t_err; }
int64_t ng_launch_count(void) { return g_ng_launches.load(); }
const char* ng_last_kernel(void) { return t_ng_last_kernel; }
void ng_tuning_reload(void) { g_env_gen.fetch_add(1, std::memory_order_acq_rel); }

int ng_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

int ng_fd_plan_create(int ndim, const int64_t* shape, int axis, int n_center,
                      const double* w_center, const int32_t* off_center, int n_side,
                      const double* w_fwd, const int32_t* off_fwd, const double* w_bwd,
                      const int32_t* off_bwd, double inv_h_pow, const double* h_x_div,
                      ng_plan** out) {
    NG_REQUIRE(out != nullptr, "out is NULL");
    *out = nullptr;
    NG_REQUIRE(n_center >= 1 && n_center <= NG_MAX_TAPS && n_side >= 1 && n_side <= NG_MAX_TAPS,
               "tap counts (%d central, %d one-sided) must be in 1..%d", n_center, n_side, NG_MAX_TAPS);
    NG_REQUIRE(w_center && off_center && w_fwd && off_fwd && w_bwd && off_bwd, "NULL tap table");
    ng_plan* p = new (std::nothrow) ng_plan();
    if (!p) { ng_set_error("out of host memory"); return NG_ENOMEM; }
    int rc = set_shape(p, ndim, shape, axis);
    if (rc) { delete p; return rc; }
    p->kind = PK_FD;
    FdTables& T = p->fd;
    memset(&T, 0, sizeof(T));
    T.n_center = n_center;
    T.n_side = n_side;
    T.nb = n_center / 2;
    T.inv_h_pow = inv_h_pow;
    const i64 n = p->n;
    bool ok = n >= 2 * (i64)T.nb;
    for (int k = 0; k < n_center; ++k) {
        T.wc[k] = w_center[k]; T.oc[k] = off_center[k];
        if (off_center[k] < -T.nb || off_center[k] > T.nb) ok = false;
    }
    for (int k = 0; k < n_side; ++k) {
        T.wf[k] = w_fwd[k]; T.of[k] = off_fwd[k];
        T.wb[k] = w_bwd[k]; T.ob[k] = off_bwd[k];
        // forward scheme is used at i in [0, nb), backward at [n-nb, n): all taps must stay inside
        if (off_fwd[k] < 0 || (T.nb - 1) + off_fwd[k] >= n) ok = false;
        if (off_bwd[k] > 0 || (n - T.nb) + off_bwd[k] < 0) ok = false;
    }
    if (!ok) {
        ng_set_error("axis of %lld points is too short for a %d/%d-tap finite-difference stencil",
                     (long long)n, n_center, n_side);
        delete p;
        return NG_EINVAL;
    }
    if (h_x_div) {
        rc = require_device();
        if (!rc) rc = upload(p, &p->d_xdiv, h_x_div, n);
        if (rc) { ng_plan_destroy(p); return rc; }
    }
    *out = p;
    return NG_OK;
}

int ng_cheb_plan_create(int ndim, const int64_t* shape, int axis, const double* h_M,
                        int descending_k, ng_plan** out) {
    NG_REQUIRE(out != nullptr, "out is NULL");
    *out = nullptr;
    NG_REQUIRE(h_M != nullptr, "matrix is NULL");
    ng_plan* p = new (std::nothrow) ng_plan();
    if (!p) { ng_set_error("out of host memory"); return NG_ENOMEM; }
    int rc = set_shape(p, ndim, shape, axis);
    if (!rc) rc = require_device();
    if (rc) { delete p; return rc; }
    p->kind = PK_CHEB;
    p->descending = descending_k ? 1 : 0;
    const i64 n = p->n;
    std::vector<double> Mt((size_t)(n * n));
    for (i64 i = 0; i < n; ++i)
        for (i64 k = 0; k < n; ++k) Mt[(size_t)(k * n + i)] = h_M[i * n + k];
    rc = upload(p, &p->d_M, h_M, n * n);
    if (!rc) rc = upload(p, &p->d_Mt, Mt.data(), n * n);
    if (!rc) rc = upload_padded_mt(p, Mt.data());
    if (rc) { ng_plan_destroy(p); return rc; }
    *out = p;
    return NG_OK;
}

// Device tables of an m-point power-of-two transform (m = the axis length, or the padded length of
// ng_fft_plan_create_padded): multipliers / m in bit-reversed order + half twiddle circle (radix-2 kernel, fft.cu),
// multipliers / m in natural order + full twiddle circle (two-pass register kernels, fft2.cu).
static int build_fft_tables(ng_plan* p, i64 m, const double* h_W_re_im) {
    p->fft_pow2 = 1;
    p->fft_len = (int)m;
    int l = 0;
    while ((1LL << l) < m) ++l;
    p->log2n = l;
    std::vector<double> Wbr((size_t)(2 * m)), tw((size_t)m);   // tw: m/2 complex
    for (i64 k = 0; k < m; ++k) {
        i64 r = 0;
        for (int b = 0; b < l; ++b) if (k & (1LL << b)) r |= 1LL << (l - 1 - b);
        // position r of the DIF output holds bin k
        Wbr[(size_t)(2 * r)] = h_W_re_im[2 * k] / (double)m;
        Wbr[(size_t)(2 * r + 1)] = h_W_re_im[2 * k + 1] / (double)m;
    }
    // twiddles exp(-2 pi i j / m), j < m/2, evaluated in long double and rounded once
    const long double two_pi = 6.283185307179586476925286766559005768L;
    for (i64 j = 0; j < m / 2; ++j) {
        const long double a = two_pi * (long double)j / (long double)m;
        tw[(size_t)(2 * j)] = (double)cosl(a);
        tw[(size_t)(2 * j + 1)] = (double)(-sinl(a));
    }
    int rc = upload(p, &p->d_W, Wbr.data(), 2 * m);
    if (!rc) rc = upload(p, &p->d_tw, tw.data(), m);
    std::vector<double> Wn((size_t)(2 * m)), twn((size_t)(2 * m));
    for (i64 k = 0; k < m; ++k) {
        Wn[(size_t)(2 * k)] = h_W_re_im[2 * k] / (double)m;
        Wn[(size_t)(2 * k + 1)] = h_W_re_im[2 * k + 1] / (double)m;
        const long double a = two_pi * (long double)k / (long double)m;
        twn[(size_t)(2 * k)] = (double)cosl(a);
        twn[(size_t)(2 * k + 1)] = (double)(-sinl(a));
    }
    if (!rc) rc = upload(p, &p->d_Wn, Wn.data(), 2 * m);
    if (!rc) rc = upload(p, &p->d_twn, twn.data(), 2 * m);
    return rc;
}

// Periodic axis of ANY length n as a zero-padded power-of-two transform: the operator is a circular convolution with
// the real kernel d = ifft(W); laid out on a circle of m >= 2n-1 points (d[k] at k, d[n-k] at m-k) its m-point
// transform times the transform of the zero-padded pencil gives, on the first n outputs, exactly the n-point circular
// convolution.  h_Wm_re_im = fft_m(d on the circle), computed by the host; rows in memory keep n samples.
int ng_fft_plan_create_padded(int ndim, const int64_t* shape, int axis, int64_t m, const double* h_Wm_re_im,
                              ng_plan** out) {
    NG_REQUIRE(out != nullptr, "out is NULL");
    *out = nullptr;
    NG_REQUIRE(h_Wm_re_im != nullptr, "multiplier table is NULL");
    ng_plan* p = new (std::nothrow) ng_plan();
    if (!p) { ng_set_error("out of host memory"); return NG_ENOMEM; }
    int rc = set_shape(p, ndim, shape, axis);
    if (!rc) rc = require_device();
    if (rc) { delete p; return rc; }
    p->kind = PK_FFT;
    if (!(m >= 2 && (m & (m - 1)) == 0 && m <= 8192 && m >= 2 * p->n - 1)) {
        ng_set_error("padded transform length %lld must be a power of two in [2n-1, 8192] for n = %lld", (long long)m,
                     (long long)p->n);
        delete p;
        return NG_EINVAL;
    }
    rc = build_fft_tables(p, m, h_Wm_re_im);
    if (rc) { ng_plan_destroy(p); return rc; }
    *out = p;
    return NG_OK;
}

int ng_fft_plan_create(int ndim, const int64_t* shape, int axis, const double* h_W_re_im,
                       const double* h_D, ng_plan** out) {
    NG_REQUIRE(out != nullptr, "out is NULL");
    *out = nullptr;
    NG_REQUIRE(h_W_re_im != nullptr, "multiplier table is NULL");
    ng_plan* p = new (std::nothrow) ng_plan();
    if (!p) { ng_set_error("out of host memory"); return NG_ENOMEM; }
    int rc = set_shape(p, ndim, shape, axis);
    if (!rc) rc = require_device();
    if (rc) { delete p; return rc; }
    p->kind = PK_FFT;
    const i64 n = p->n;
    const bool pow2 = n >= 2 && (n & (n - 1)) == 0 && n <= 8192;
    const bool force_dense = NG_ENV_INT("NG_FFT_DENSE", 0) != 0;
    const bool mixed = !pow2 && !h_D && ng_fft_mixed_supported(n);     // (a caller that passes the dense matrix asks for it)
    if (pow2 && !(force_dense && h_D)) {
        rc = build_fft_tables(p, n, h_W_re_im);
    } else if (mixed) {
        std::vector<double> Wdr, tw;
        rc = ng_fft_mixed_tables(p, h_W_re_im, Wdr, tw);
        if (!rc) rc = upload(p, &p->d_W, Wdr.data(), 2 * n);
        if (!rc) rc = upload(p, &p->d_twn, tw.data(), 2 * n);
    } else {
        if (!h_D) {
            ng_set_error("n = %lld is neither a power of two nor a product of 2, 3, 5, 7, 11, 13 (<= 8192): the dense spectral matrix is required",
                         (long long)n);
            delete p;
            return NG_EINVAL;
        }
        p->use_fma = 1;
        std::vector<double> Mt((size_t)(n * n));
        for (i64 i = 0; i < n; ++i)
            for (i64 k = 0; k < n; ++k) Mt[(size_t)(k * n + i)] = h_D[i * n + k];
        rc = upload(p, &p->d_M, h_D, n * n);
        if (!rc) rc = upload(p, &p->d_Mt, Mt.data(), n * n);
        if (!rc) rc = upload_padded_mt(p, Mt.data());
    }
    if (rc) { ng_plan_destroy(p); return rc; }
    *out = p;
    return NG_OK;
}

int ng_apply(const ng_plan* plan, const double* d_in, double* d_out, void* stream) {
    return ng_apply_fused(plan, d_in, d_out, nullptr, stream);
}

int ng_apply_fused(const ng_plan* plan, const double* d_in, double* d_out, const ng_fuse* fuse,
                   void* stream) {
    NG_REQUIRE(plan && d_in && d_out, "ng_apply: NULL plan or buffer");
    NG_REQUIRE(d_in != d_out, "ng_apply: output must not alias input");
    FuseDev F;
    ng_fuse_to_dev(plan, fuse, &F);
    return ng_apply_any(plan, d_in, d_out, F, (cudaStream_t)stream);
}

// `count` fields of the plan's shape stacked in one array [count][shape...]: along any axis the stack is the
// same [outer][n][inner] view with `count` times the outer extent, so one launch covers all of them -- a
// 512 x 512 field is launch-bound on its own (cfg1: 0.2 of the HBM peak), eighty of them stream.
// The staging arrays of the *_host entry points (16 GiB at 1024^3) belong to the plan until released or destroyed.
int ng_plan_release_staging(ng_plan* p) {
    NG_REQUIRE(p != nullptr, "plan is NULL");
    if (!p->d_stage_in && !p->d_stage_out) return NG_OK;
    int rc = ng_check_plan_device(p, "ng_plan_release_staging");
    if (rc) return rc;
    if (p->d_stage_in) cudaFree(p->d_stage_in);
    if (p->d_stage_out) cudaFree(p->d_stage_out);
    p->device_bytes -= sizeof(double) * (p->stage_elems + p->size);
    p->d_stage_in = p->d_stage_out = nullptr;
    p->stage_elems = 0;
    return NG_OK;
}

int ng_apply_batch(const ng_plan* plan, const double* d_in, double* d_out, int64_t count, void* stream) {
    NG_REQUIRE(plan && d_in && d_out && d_in != d_out, "ng_apply_batch: bad plan or buffers");
    NG_REQUIRE(plan->kind == PK_FD || plan->kind == PK_CHEB || plan->kind == PK_FFT,
               "plan kind %d cannot be applied with ng_apply_batch", plan->kind);
    NG_REQUIRE(count >= 1 && plan->size * count < (1LL << 40), "ng_apply_batch: bad field count %lld", (long long)count);
    ng_plan view = *plan;                                 // tables are shared, nothing is owned
    view.outer = plan->outer * count;
    view.size = plan->size * count;
    view.d_stage_in = view.d_stage_out = nullptr;
    FuseDev F = no_fuse(&view);
    return ng_apply_any(&view, d_in, d_out, F, (cudaStream_t)stream);
}

// Host-buffer pipeline shared by ng_apply_host / ng_fdlap_apply_host: the array is cut in chunks of
// axis-0 planes; chunk k+1 is copied in (stream 0 of 3) while chunk k is computed (stream 1) and
// chunk k-1 is copied out (stream 2), so H2D and D2H use both PCIe directions at once and the
// kernels hide under the copies.  `launch(first_plane, last_plane, stream)` computes output planes
// [first, last); `halo` = input planes it needs on either side (2 for the Laplacian / FD along axis
// 0 would need the whole axis -> those run unchunked).  Pageable host memory degrades to serialised
// copies but stays correct.
extern "C++" {
template <class Launch>
static int host_pipeline(ng_plan* plan, const double* h_in, double* h_out, i64 halo, bool chunkable,
                         Launch launch) {
    int rc = ensure_stage(plan, 1);
    if (rc) return rc;
    const i64 n0 = plan->shape[0];
    const i64 plane = plan->size / n0;                       // elements per axis-0 plane
    // three staging streams per device, created on first use
    static std::mutex st_mutex;
    static cudaStream_t st_all[64][3] = {};
    cudaStream_t* st = st_all[ng_current_device() & 63];
    {
        std::lock_guard<std::mutex> lock(st_mutex);
        if (!st[0])
            for (int i = 0; i < 3; ++i) NG_CUDA_OK(cudaStreamCreateWithFlags(&st[i], cudaStreamNonBlocking));
    }
    // ~256 MiB chunks, at least 8 planes, at most 64 chunks
    i64 chunk_elems = 32LL << 20;                             // 256 MiB
    if (const i64 mb = NG_ENV_INT("NG_HOST_CHUNK_MB", 0)) chunk_elems = mb * 131072;   // tests
    if (chunk_elems < 1) chunk_elems = 1;
    i64 cp = chunk_elems / (plane > 0 ? plane : 1);
    if (cp < 8) cp = 8;
    if (!chunkable || plan->size <= chunk_elems || cp >= n0) cp = n0;
    const int nchunk = (int)((n0 + cp - 1) / cp);
    if (nchunk > 256) {                                       // keep the event arrays bounded
        cp = (n0 + 255) / 256;
    }
    const int nc = (int)((n0 + cp - 1) / cp);
    std::vector<cudaEvent_t> in_done((size_t)nc), out_done((size_t)nc);
    for (int k = 0; k < nc; ++k) {
        NG_CUDA_OK(cudaEventCreateWithFlags(&in_done[k], cudaEventDisableTiming));
        NG_CUDA_OK(cudaEventCreateWithFlags(&out_done[k], cudaEventDisableTiming));
    }
    rc = NG_OK;
    auto run_chunk = [&](int c, cudaEvent_t ready) -> int {       // compute + copy-out of chunk c
        const i64 ca = c * cp, cb = (ca + cp < n0) ? ca + cp : n0;
        cudaStreamWaitEvent(st[1], ready, 0);
        int r = launch(ca, cb, st[1]);
        if (r) return r;
        cudaEventRecord(out_done[c], st[1]);
        cudaStreamWaitEvent(st[2], out_done[c], 0);
        if (cudaMemcpyAsync(h_out + ca * plane, plan->d_stage_out + ca * plane,
                            sizeof(double) * (size_t)((cb - ca) * plane), cudaMemcpyDeviceToHost,
                            st[2]) != cudaSuccess) return NG_ECUDA;
        return NG_OK;
    };
    for (int k = 0; k < nc && rc == NG_OK; ++k) {
        const i64 a = k * cp, b = (a + cp < n0) ? a + cp : n0;
        if (cudaMemcpyAsync(plan->d_stage_in + a * plane, h_in + a * plane,
                            sizeof(double) * (size_t)((b - a) * plane), cudaMemcpyHostToDevice,
                            st[0]) != cudaSuccess) { rc = NG_ECUDA; break; }
        cudaEventRecord(in_done[k], st[0]);
        if (halo == 0) {
            rc = run_chunk(k, in_done[k]);
        } else {
            // chunk k-1 can run once chunk k (its upper halo) has landed; the last chunk follows at once
            if (k >= 1) rc = run_chunk(k - 1, in_done[k]);
            if (rc == NG_OK && k == nc - 1) rc = run_chunk(k, in_done[k]);
        }
    }
    cudaError_t e = cudaSuccess;
    for (int i = 0; i < 3; ++i) {
        cudaError_t ei = cudaStreamSynchronize(st[i]);
        if (ei != cudaSuccess) e = ei;
    }
    for (int k = 0; k < nc; ++k) { cudaEventDestroy(in_done[k]); cudaEventDestroy(out_done[k]); }
    if (rc) return rc;
    if (e != cudaSuccess) {
        ng_set_error("host pipeline failed: %s", cudaGetErrorString(e));
        return NG_ECUDA;
    }
    return NG_OK;
}
}  // extern "C++"

int ng_apply_host(ng_plan* plan, const double* h_in, double* h_out) {
    NG_REQUIRE(plan && h_in && h_out, "ng_apply_host: NULL plan or buffer");
    NG_REQUIRE(plan->kind == PK_FD || plan->kind == PK_CHEB || plan->kind == PK_FFT,
               "plan kind %d cannot be applied with ng_apply_host", plan->kind);
    int rc = require_device();
    if (!rc) rc = ng_check_plan_device(plan, "ng_apply_host");
    if (rc) return rc;
    // pencils along an axis other than 0 are independent across axis-0 planes: chunk a shallow view
    const bool chunkable = plan->axis != 0 && plan->ndim > 1;
    const i64 plane = plan->size / plan->shape[0];
    return host_pipeline(plan, h_in, h_out, 0, chunkable, [&](i64 a, i64 b, cudaStream_t s) -> int {
        if (a == 0 && b == plan->shape[0]) {
            FuseDev F = no_fuse(plan);
            return ng_apply_any(plan, plan->d_stage_in, plan->d_stage_out, F, s);
        }
        ng_plan view = *plan;                                // tables are shared, nothing is owned
        view.shape[0] = b - a;
        view.size = (b - a) * plane;
        view.outer = plan->outer / plan->shape[0] * (b - a);
        FuseDev F = no_fuse(&view);
        return ng_apply_any(&view, plan->d_stage_in + a * plane, plan->d_stage_out + a * plane, F, s);
    });
}

int ng_plan_halo_width(const ng_plan* plan) {
    if (!plan) return 0;
    if (plan->kind == PK_FD) return plan->fd.nb;
    if (plan->kind == PK_FDLAP) return plan->lap[plan->ndim - 1].nb;
    return 0;
}

int ng_fd_apply_slab(const ng_plan* plan, const double* d_in, double* d_out,
                     const double* d_lo_halo, const double* d_hi_halo, void* stream) {
    NG_REQUIRE(plan && plan->kind == PK_FD, "ng_fd_apply_slab: finite-difference plan required");
    NG_REQUIRE(d_in && d_out && d_in != d_out, "ng_fd_apply_slab: bad buffers");
    FuseDev F = no_fuse(plan);
    return ng_fd_launch(plan, d_in, d_out, F, d_lo_halo, d_hi_halo, (cudaStream_t)stream);
}

int ng_apply_fused_slab(const ng_plan* plan, const double* d_in, double* d_out, const ng_fuse* fuse,
                        const double* d_lo_halo, const double* d_hi_halo, void* stream) {
    NG_REQUIRE(plan && plan->kind == PK_FD, "ng_apply_fused_slab: finite-difference plan required");
    NG_REQUIRE(d_in && d_out && d_in != d_out, "ng_apply_fused_slab: bad buffers");
    FuseDev F;
    ng_fuse_to_dev(plan, fuse, &F);
    return ng_fd_launch(plan, d_in, d_out, F, d_lo_halo, d_hi_halo, (cudaStream_t)stream);
}

int ng_fdlap_plan_create(int ndim, const int64_t* shape, ng_plan* const* axis_plans,
                         ng_plan** out) {
    NG_REQUIRE(out != nullptr, "out is NULL");
    *out = nullptr;
    NG_REQUIRE(axis_plans != nullptr, "axis_plans is NULL");
    ng_plan* p = new (std::nothrow) ng_plan();
    if (!p) { ng_set_error("out of host memory"); return NG_ENOMEM; }
    int rc = set_shape(p, ndim, shape, ndim - 1 >= 0 ? ndim - 1 : 0);
    if (!rc) rc = require_device();
    if (rc) { delete p; return rc; }
    p->kind = PK_FDLAP;
    for (int a = 0; a < ndim; ++a) {
        const ng_plan* q = axis_plans[a];
        bool ok = q && q->kind == PK_FD && q->ndim == ndim && q->axis == a && !q->d_xdiv;
        for (int b = 0; ok && b < ndim; ++b) ok = q->shape[b] == p->shape[b];
        if (!ok) {
            ng_set_error("axis_plans[%d] must be an equidistant FD plan of the same shape along axis %d", a, a);
            delete p;
            return NG_EINVAL;
        }
        p->lap[a] = q->fd;
    }
    rc = ng_fdlap_upload_params(p);
    if (rc) { ng_plan_destroy(p); return rc; }
    *out = p;
    return NG_OK;
}

int ng_fdlap_apply(const ng_plan* plan, const double* d_in, double* d_out, void* stream) {
    return ng_fdlap_apply_slab(plan, d_in, d_out, nullptr, nullptr, stream);
}

int ng_fdlap_apply_slab(const ng_plan* plan, const double* d_in, double* d_out,
                        const double* d_lo_halo, const double* d_hi_halo, void* stream) {
    NG_REQUIRE(plan && plan->kind == PK_FDLAP, "ng_fdlap_apply: Laplacian plan required");
    NG_REQUIRE(d_in && d_out && d_in != d_out, "ng_fdlap_apply: bad buffers");
    return ng_fdlap_launch(plan, d_in, d_out, d_lo_halo, d_hi_halo, 0, plan->shape[0], 0, (cudaStream_t)stream);
}

int ng_fdlap_apply_slab_range(const ng_plan* plan, const double* d_in, double* d_out,
                              const double* d_lo_halo, const double* d_hi_halo, int64_t x_begin,
                              int64_t x_end, void* stream) {
    NG_REQUIRE(plan && plan->kind == PK_FDLAP, "ng_fdlap_apply_slab_range: Laplacian plan required");
    NG_REQUIRE(d_in && d_out && d_in != d_out, "ng_fdlap_apply_slab_range: bad buffers");
    NG_REQUIRE(0 <= x_begin && x_begin <= x_end && x_end <= plan->shape[0],
               "ng_fdlap_apply_slab_range: [%lld, %lld) is not inside axis 0 of %lld planes",
               (long long)x_begin, (long long)x_end, (long long)plan->shape[0]);
    return ng_fdlap_launch(plan, d_in, d_out, d_lo_halo, d_hi_halo, x_begin, x_end, 0, (cudaStream_t)stream);
}

int ng_fdlap_apply_slab_part(const ng_plan* plan, const double* d_in, double* d_out,
                             const double* d_lo_halo, const double* d_hi_halo, int z_part, void* stream) {
    NG_REQUIRE(plan && plan->kind == PK_FDLAP, "ng_fdlap_apply_slab_part: Laplacian plan required");
    NG_REQUIRE(d_in && d_out && d_in != d_out, "ng_fdlap_apply_slab_part: bad buffers");
    NG_REQUIRE(z_part >= 0 && z_part <= 3, "ng_fdlap_apply_slab_part: z_part must be 0, 1 or 2");
    return ng_fdlap_launch(plan, d_in, d_out, d_lo_halo, d_hi_halo, 0, plan->shape[0], z_part, (cudaStream_t)stream);
}

int ng_fdlap_apply_host(ng_plan* plan, const double* h_in, double* h_out) {
    NG_REQUIRE(plan && plan->kind == PK_FDLAP && h_in && h_out, "ng_fdlap_apply_host: bad arguments");
    int rc = require_device();
    if (!rc) rc = ng_check_plan_device(plan, "ng_fdlap_apply_host");
    if (rc) return rc;
    // output planes [a, b) need input planes [a-2, b+2) (and up to 5 further in on the global
    // boundary, which the >= 8-plane chunks cover): run chunk k when chunk k+1 has landed
    return host_pipeline(plan, h_in, h_out, 2, plan->ndim > 1, [&](i64 a, i64 b, cudaStream_t s) -> int {
        return ng_fdlap_launch(plan, plan->d_stage_in, plan->d_stage_out, nullptr, nullptr, a, b, 0, s);
    });
}

int64_t ng_plan_device_bytes(const ng_plan* plan) { return plan ? plan->device_bytes : 0; }

int ng_plan_destroy(ng_plan* p) {
    if (!p) return NG_OK;
    auto fr = [](double* q) { if (q) cudaFree(q); };
    fr(p->d_xdiv); fr(p->d_M); fr(p->d_Mt); fr(p->d_Mtp); fr(p->d_W); fr(p->d_tw); fr(p->d_Wn); fr(p->d_twn); fr(p->d_work);
    fr(p->d_stage_in); fr(p->d_stage_out);
    if (p->d_curvlap) cudaFree(p->d_curvlap);
    if (p->slab) ng_slab_state_destroy(p->slab);
    fr(p->cJ.d_ptr);
    for (int a = 0; a < NG_MAX_DIM; ++a) {
        fr(p->cLap[a].d_ptr); fr(p->cDiv[a].d_ptr); fr(p->cGrad[a].d_ptr);
        fr(p->cH[a].d_ptr); fr(p->cCurl[a].d_ptr);
    }
    delete p;
    return NG_OK;
}

}  // extern "C"
