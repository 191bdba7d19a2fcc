// common.cuh -- plan object, error plumbing and the fused-epilogue device helpers shared by all
// translation units of libnumgrids_b200.so.  sm_100a only; compiled with -fmad=false so that
// every `a*b+c` in the bit-exact paths stays a separately rounded DMUL + DADD (the reference's
// NumPy/SciPy arithmetic); the FFT path opts in to FMA with explicit fma().
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <atomic>
#include <string>
#include <vector>

#include "../../include/numgrids_b200.h"

typedef long long i64;

// ---------------------------------------------------------------------------------------------
// error plumbing
// ---------------------------------------------------------------------------------------------
void ng_set_error(const char* fmt, ...);
extern std::atomic<long long> g_ng_launches;

#define NG_CUDA_OK(expr)                                                                   \
    do {                                                                                   \
        cudaError_t _e = (expr);                                                           \
        if (_e != cudaSuccess) {                                                           \
            ng_set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, \
                         __LINE__);                                                        \
            return NG_ECUDA;                                                               \
        }                                                                                  \
    } while (0)

#define NG_LAUNCH_CHECK()                                                              \
    do {                                                                               \
        g_ng_launches.fetch_add(1, std::memory_order_relaxed);                         \
        cudaError_t _e = cudaGetLastError();                                           \
        if (_e != cudaSuccess) {                                                       \
            ng_set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e),   \
                         __FILE__, __LINE__);                                          \
            return NG_ECUDA;                                                           \
        }                                                                              \
    } while (0)

#define NG_REQUIRE(cond, ...)            \
    do {                                 \
        if (!(cond)) {                   \
            ng_set_error(__VA_ARGS__);   \
            return NG_EINVAL;            \
        }                                \
    } while (0)

// ---------------------------------------------------------------------------------------------
// tuning switches, kernel trace, per-device function attributes
// ---------------------------------------------------------------------------------------------
// Tuning / A-B switches come from the environment but are read ONCE per call site: the value is
// cached until ng_tuning_reload() (tests and tuning tools call it after changing the environment),
// so the hot path never calls getenv().
struct NgEnvCache {
    std::atomic<int> gen{-1};
    std::atomic<long long> val{0};
};
long long ng_env_cached(NgEnvCache* c, const char* name, long long dflt);
#define NG_ENV_INT(name, dflt)                                  \
    ([]() -> long long {                                        \
        static NgEnvCache cache_;                               \
        return ng_env_cached(&cache_, name, (long long)(dflt)); \
    }())

// name of the kernel instance the calling thread launched last (ng_last_kernel(); tests assert that
// the instance they mean to check is the one that ran)
extern thread_local const char* t_ng_last_kernel;
#define NG_TRACE(name) (t_ng_last_kernel = (name))

// cudaFuncAttributeMaxDynamicSharedMemorySize is per device: set it once per (kernel, device).
// Usage: NG_FUNC_SMEM(bytes, kernel<template, args>);
int ng_current_device();
#define NG_FUNC_SMEM(bytes, ...)                                                                  \
    do {                                                                                          \
        static std::atomic<unsigned long long> done_{0};                                          \
        const unsigned long long bit_ = 1ull << (ng_current_device() & 63);                       \
        if (!(done_.load(std::memory_order_relaxed) & bit_)) {                                    \
            NG_CUDA_OK(cudaFuncSetAttribute(__VA_ARGS__, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                            (int)(bytes)));                                       \
            done_.fetch_or(bit_, std::memory_order_relaxed);                                      \
        }                                                                                         \
    } while (0)

// ---------------------------------------------------------------------------------------------
// tables passed to kernels by value
// ---------------------------------------------------------------------------------------------
struct FdTables {
    int n_center, n_side, nb;
    int oc[NG_MAX_TAPS], of[NG_MAX_TAPS], ob[NG_MAX_TAPS];
    double wc[NG_MAX_TAPS], wf[NG_MAX_TAPS], wb[NG_MAX_TAPS];
    double inv_h_pow;
};

struct CoefDev {
    const double* ptr;
    i64 s[NG_MAX_DIM];
};

struct FuseDev {
    CoefDev pre, post, div;
    // Double apply (internal, not part of ng_fuse): out = tail(D(mid * D(pre * in))) in ONE launch -- the
    // D_i((J/h_i^2) D_i f) term of the curvilinear Laplacian with the inner array never leaving the SM.
    // Honoured by the register FFT kernels only (ng_fft_twice_supported, fft2.cu); same arithmetic in the
    // same order as the two launches it replaces (pencil pairs poisoned by a non-finite value are redone by the slow
    // routine in both cases, but then for both transforms: agreement to rounding there).
    CoefDev mid;
    const double* minuend;
    const double* addend;
    int add_zero;
    int finite;
    int any;              // 0 = plain apply
    int vec_axis;         // innermost grid axis with extent > 1 (the axis a double2 pair runs along)
    i64 shape[NG_MAX_DIM];
};

enum PlanKind { PK_FD = 1, PK_CHEB = 2, PK_FFT = 3, PK_FDLAP = 4, PK_CURV = 5, PK_SLAB = 6 };
void ng_slab_state_destroy(void* state);     // slab.cu

struct CoefHost {           // device-resident compact coefficient table owned by a curv plan
    double* d_ptr = nullptr;
    i64 count = 0;
    i64 stride[NG_MAX_DIM] = {0, 0, 0, 0};
    bool is_one = false;    // the table is the single value 1.0 (multiplying / dividing by it is exact: skipped)
};

struct ng_plan {
    int kind = 0;
    int ndim = 0;
    int axis = 0;
    i64 shape[NG_MAX_DIM] = {1, 1, 1, 1};
    i64 size = 1;
    i64 outer = 1, n = 1, inner = 1;   // [outer][n][inner] view around `axis`
    i64 device_bytes = 0;
    int device = -1;                   // CUDA device that holds the plan's tables (-1: the plan owns no device memory)

    // PK_FD
    FdTables fd;
    double* d_xdiv = nullptr;

    // PK_CHEB (and the dense fallback of PK_FFT)
    double* d_M = nullptr;    // [n][n] row-major
    double* d_Mt = nullptr;   // transposed copy (k-major) for broadcast-friendly panel loads
    double* d_Mtp = nullptr;  // the same, zero-padded to [n_pad][n_pad], n_pad = n rounded up to 128 (general kernel)
    int n_pad = 0;
    int descending = 0;
    int use_fma = 0;

    // PK_FFT
    int fft_pow2 = 0;
    int log2n = 0;
    int fft_len = 0;          // transform length: n, or the power of two >= 2n-1 of a zero-padded plan (ng_fft_plan_create_padded)
    double* d_W = nullptr;    // [n][2]
    double* d_tw = nullptr;   // [n/2][2] forward twiddles exp(-2 pi i k / n) (radix-2 kernel)
    double* d_Wn = nullptr;   // [n][2] multipliers / n in natural order (two-pass kernel, fft2.cu)
    double* d_twn = nullptr;  // [n][2] exp(-2 pi i k / n), k < n (two-pass kernel)
    int fft_mixed = 0;        // native mixed-radix plan of a non-power-of-two n (fft_mixed.cu): d_W digit-reversed, d_twn
    int mix_nst = 0;
    int mix_radix[12] = {0};
    int mix_pairs = 1;        // pencil pairs per CTA
    int mix_pad_shift = 4;    // shared-memory index padding i + (i >> shift), chosen by the plan's bank-conflict model (31: none)

    // PK_FDLAP
    FdTables lap[NG_MAX_DIM];

    // PK_CURV
    ng_plan* axis_plans[NG_MAX_DIM] = {nullptr, nullptr, nullptr, nullptr};
    CoefHost cJ, cLap[NG_MAX_DIM], cDiv[NG_MAX_DIM], cGrad[NG_MAX_DIM], cH[NG_MAX_DIM],
        cCurl[NG_MAX_DIM];
    double* d_work = nullptr;
    void* d_curvlap = nullptr;   // parameter block of the single-pass Laplacian (curvlap3d.cu)
    const ng_plan* axis_sq[NG_MAX_DIM] = {nullptr, nullptr, nullptr, nullptr};   // PK_CURV: D o D as one apply (periodic axes)

    // PK_SLAB
    void* slab = nullptr;        // SlabState (slab.cu): local Laplacian plan, peer pointers, side stream, events

    // host-call staging (ng_apply_host / ng_fdlap_apply_host)
    double* d_stage_in = nullptr;
    double* d_stage_out = nullptr;
    i64 stage_elems = 0;
};

// per-kind entry points implemented in the other translation units
int ng_fd_launch(const ng_plan* p, const double* in, double* out, const FuseDev& F,
                 const double* lo, const double* hi, cudaStream_t st);
// contiguous-axis streaming kernel (fd_row.cu); P = half width of the central stencil (1..4)
int ng_fd_row_launch(int P, const ng_plan* p, const double* in, double* out, const FuseDev& F,
                     const double* lo, const double* hi, cudaStream_t st);
int ng_dense_launch(const ng_plan* p, const double* in, double* out, const FuseDev& F,
                    cudaStream_t st);
int ng_fft_launch(const ng_plan* p, const double* in, double* out, const FuseDev& F,
                  cudaStream_t st);
int ng_fft2_try_launch(const ng_plan* p, const double* in, double* out, const FuseDev& F,
                       cudaStream_t st, bool* handled);
bool ng_fft_twice_supported(const ng_plan* p);
// native mixed-radix transform of a non-power-of-two length with prime factors <= 13 (fft_mixed.cu)
#define NG_FFT_MIXED_MAX 8192
bool ng_fft_mixed_supported(i64 n);
int ng_fft_mixed_tables(ng_plan* p, const double* h_W_re_im, std::vector<double>& Wdr, std::vector<double>& tw);
int ng_fft_mixed_launch(const ng_plan* p, const double* in, double* out, const FuseDev& F, cudaStream_t st);
int ng_fft_tile_try_launch(const ng_plan* p, const double* in, double* out, const FuseDev& F,
                           cudaStream_t st, bool* handled);
int ng_fdlap_launch(const ng_plan* p, const double* in, double* out, const double* lo,
                    const double* hi, i64 xbeg, i64 xend, int zmode, cudaStream_t st);
int ng_apply_any(const ng_plan* p, const double* in, double* out, const FuseDev& F,
                 cudaStream_t st);
// NG_OK when the calling thread's current device is the one the plan's tables live on
int ng_check_plan_device(const ng_plan* p, const char* who);
// call before the first cudaMalloc on behalf of a plan: records (or checks) the owning device
int ng_plan_bind_device(ng_plan* p);
void ng_fuse_to_dev(const ng_plan* p, const ng_fuse* f, FuseDev* out);

// mbarrier.try_wait suspend-time hint (ns): a waiting thread is parked by the hardware until the phase completes
// or the hint expires, instead of polling -- a producer lane that spins issues ~10 thread-instructions per grid
// point in the plane-ring kernels, slots the consumer warps of the same scheduler then do not get
#define NG_MBAR_SUSPEND_NS 1000u

// ---------------------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------------------
struct Idx4 {
    i64 i[NG_MAX_DIM];
};

__device__ __forceinline__ Idx4 unravel4(i64 lin, const i64* shape) {
    Idx4 r;
    if (lin < 0x7fffffffLL) {   // 32-bit fast path (all config sizes)
        unsigned t = (unsigned)lin;
        unsigned s3 = (unsigned)shape[3], s2 = (unsigned)shape[2], s1 = (unsigned)shape[1];
        unsigned q = t / s3; r.i[3] = t - q * s3; t = q;
        q = t / s2; r.i[2] = t - q * s2; t = q;
        q = t / s1; r.i[1] = t - q * s1; r.i[0] = q;
    } else {
        i64 t = lin;
        r.i[3] = t % shape[3]; t /= shape[3];
        r.i[2] = t % shape[2]; t /= shape[2];
        r.i[1] = t % shape[1]; r.i[0] = t / shape[1];
    }
    return r;
}

__device__ __forceinline__ i64 coef_off(const CoefDev& c, const Idx4& ix) {
    return ix.i[0] * c.s[0] + ix.i[1] * c.s[1] + ix.i[2] * c.s[2] + ix.i[3] * c.s[3];
}

// fuse_tail with the coefficient offsets already resolved (kernels that walk one axis resolve the
// other coordinates once and step the offset); same steps and order as fuse_tail below.
__device__ __forceinline__ double fuse_tail_off(const FuseDev& F, double d, i64 lin, i64 post_off, i64 div_off) {
    if (F.minuend) d = __dsub_rn(F.minuend[lin], d);
    if (F.post.ptr) d = __dmul_rn(F.post.ptr[post_off], d);
    if (F.addend) d = __dadd_rn(F.addend[lin], d);
    else if (F.add_zero) d = __dadd_rn(0.0, d);
    if (F.div.ptr) d = __ddiv_rn(d, F.div.ptr[div_off]);
    if (F.finite) d = isfinite(d) ? d : 0.0;
    return d;
}

// fuse_tail_off with the full-shape operands (minuend / addend) already fetched: kernels that stream
// fetch them together with the input, ahead of the arithmetic (they may alias `out`, so the compiler
// cannot move those loads over earlier stores itself).  Same steps and order.
__device__ __forceinline__ double fuse_tail_vals(const FuseDev& F, double d, double m, double a, i64 post_off,
                                                 i64 div_off) {
    if (F.minuend) d = __dsub_rn(m, d);
    if (F.post.ptr) d = __dmul_rn(F.post.ptr[post_off], d);
    if (F.addend) d = __dadd_rn(a, d);
    else if (F.add_zero) d = __dadd_rn(0.0, d);
    if (F.div.ptr) d = __ddiv_rn(d, F.div.ptr[div_off]);
    if (F.finite) d = isfinite(d) ? d : 0.0;
    return d;
}

// inf / nan test on the integer pipe (isfinite() would spend FP64-pipe slots in FP64-bound kernels)
__device__ __forceinline__ bool ng_nonfinite(double v) {
    return (__double2hiint(v) & 0x7ff00000) == 0x7ff00000;
}
__device__ __forceinline__ double ng_nan() { return __longlong_as_double(0x7ff8000000000000LL); }

// The elementwise tail of ng_fuse (see numgrids_b200.h for the order); `lin` is the linear index
// of the output element, `ix` its multi-index.
__device__ __forceinline__ double fuse_tail(const FuseDev& F, double d, i64 lin, const Idx4& ix) {
    if (F.minuend) d = __dsub_rn(F.minuend[lin], d);
    if (F.post.ptr) d = __dmul_rn(F.post.ptr[coef_off(F.post, ix)], d);
    if (F.addend) d = __dadd_rn(F.addend[lin], d);
    else if (F.add_zero) d = __dadd_rn(0.0, d);
    if (F.div.ptr) d = __ddiv_rn(d, F.div.ptr[coef_off(F.div, ix)]);
    if (F.finite) d = isfinite(d) ? d : 0.0;
    return d;
}
