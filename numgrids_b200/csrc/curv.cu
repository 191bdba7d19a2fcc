// curv.cu -- curvilinear composites (CurvilinearGrid.gradient / divergence / laplacian / curl,
// reference numgrids/curvilinear.py:153-317) as sequences of per-axis applies with the
// elementwise work fused into the apply kernels' prologue/epilogue (ng_fuse):
//
//   laplacian : for i: work = (J/h_i^2) * D_i f            [post]
//                      out  = (0.0 | out) + D_i work        [add_zero | addend]
//               last axis also: out / J, isfinite select    [divisor, finite]
//               (periodic axes: both applies in ONE launch, `work` stays in registers [mid])
//   divergence: for i: out = (0.0 | out) + D_i((J/h_i) * v_i)   [pre, add_zero | addend], last: /J, finite
//   gradient  : out_i = finite((1/h_i) * D_i f)
//   curl      : out_i = D_j(h_k v_k);  out_i = finite(1/(h_j h_k) * (out_i - D_k(h_j v_j)))
//
// The reference materialises ~4 full arrays per axis (coeff, inner, D(inner), result+...) and
// recomputes J/h_i**2 on every call; here the Laplacian touches one workspace array and the
// coefficients are broadcast-compact tables (e.g. r^2 sin(theta) as a [nr][ntheta] table).
// Every elementwise step is a separately rounded IEEE operation in the reference's order.
#include "common.cuh"

int ng_curvlap3d_try_launch(ng_plan* p, const double* in, double* out, cudaStream_t st, bool* handled);

namespace {

CoefDev to_dev(const CoefHost& c) {
    CoefDev d;
    d.ptr = c.d_ptr;
    for (int a = 0; a < NG_MAX_DIM; ++a) d.s[a] = c.stride[a];
    return d;
}

FuseDev base_fuse(const ng_plan* p) {
    FuseDev F;
    ng_fuse_to_dev(p, nullptr, &F);
    F.any = 1;
    return F;
}

int ensure_work(ng_plan* p) {
    if (p->d_work) return NG_OK;
    int rc = ng_plan_bind_device(p);
    if (rc) return rc;
    NG_CUDA_OK(cudaMalloc((void**)&p->d_work, sizeof(double) * (size_t)p->size));
    p->device_bytes += sizeof(double) * p->size;
    return NG_OK;
}

int check_curv(const ng_plan* p, const char* who) {
    if (!p || p->kind != PK_CURV) {
        ng_set_error("%s: curvilinear plan required", who);
        return NG_EINVAL;
    }
    return ng_check_plan_device(p, who);
}

#define NG_NEED(c, what)                                                           \
    do {                                                                           \
        if (!(c).d_ptr) {                                                          \
            ng_set_error("curvilinear coefficient %s not registered", what);       \
            return NG_EINVAL;                                                      \
        }                                                                          \
    } while (0)

}  // namespace

extern "C" {

int ng_curv_plan_create(int ndim, const int64_t* shape, ng_plan* const* axis_plans, ng_plan** out) {
    NG_REQUIRE(out != nullptr, "out is NULL");
    *out = nullptr;
    NG_REQUIRE(ndim >= 1 && ndim <= NG_MAX_DIM && shape && axis_plans, "bad arguments");
    ng_plan* p = new (std::nothrow) ng_plan();
    if (!p) { ng_set_error("out of host memory"); return NG_ENOMEM; }
    p->kind = PK_CURV;
    p->ndim = ndim;
    p->size = 1;
    for (int a = 0; a < ndim; ++a) { p->shape[a] = shape[a]; p->size *= shape[a]; }
    for (int a = 0; a < ndim; ++a) {
        const ng_plan* q = axis_plans[a];
        bool ok = q && (q->kind == PK_FD || q->kind == PK_CHEB || q->kind == PK_FFT) &&
                  q->ndim == ndim && q->axis == a;
        for (int b = 0; ok && b < ndim; ++b) ok = q->shape[b] == p->shape[b];
        if (!ok) {
            ng_set_error("axis_plans[%d] must be a derivative plan of the same shape along axis %d", a, a);
            delete p;
            return NG_EINVAL;
        }
        p->axis_plans[a] = axis_plans[a];
    }
    *out = p;
    return NG_OK;
}

int ng_curv_set_coef(ng_plan* p, int kind, int index, const double* h_values, int64_t count,
                     const int64_t* stride) {
    int rc = check_curv(p, "ng_curv_set_coef");
    if (rc) return rc;
    NG_REQUIRE(h_values && stride && count >= 1, "ng_curv_set_coef: bad table");
    NG_REQUIRE(index >= 0 && index < NG_MAX_DIM, "ng_curv_set_coef: bad index %d", index);
    CoefHost* c = nullptr;
    switch (kind) {
        case NG_COEF_J: c = &p->cJ; break;
        case NG_COEF_LAP: c = &p->cLap[index]; break;
        case NG_COEF_DIV: c = &p->cDiv[index]; break;
        case NG_COEF_GRAD: c = &p->cGrad[index]; break;
        case NG_COEF_H: c = &p->cH[index]; break;
        case NG_COEF_CURL: c = &p->cCurl[index]; break;
        default: ng_set_error("ng_curv_set_coef: unknown kind %d", kind); return NG_EINVAL;
    }
    // the table must cover every index the strides can reach
    i64 reach = 0;
    for (int a = 0; a < p->ndim; ++a) {
        NG_REQUIRE(stride[a] >= 0, "ng_curv_set_coef: negative stride");
        reach += (p->shape[a] - 1) * stride[a];
    }
    NG_REQUIRE(reach < count, "ng_curv_set_coef: strides reach element %lld of a %lld-element table",
               (long long)reach, (long long)count);
    if (c->d_ptr) { cudaFree(c->d_ptr); p->device_bytes -= sizeof(double) * c->count; c->d_ptr = nullptr; }
    rc = ng_plan_bind_device(p);
    if (rc) return rc;
    NG_CUDA_OK(cudaMalloc((void**)&c->d_ptr, sizeof(double) * (size_t)count));
    NG_CUDA_OK(cudaMemcpy(c->d_ptr, h_values, sizeof(double) * (size_t)count, cudaMemcpyHostToDevice));
    p->device_bytes += sizeof(double) * count;
    c->count = count;
    for (int a = 0; a < NG_MAX_DIM; ++a) c->stride[a] = a < p->ndim ? stride[a] : 0;
    c->is_one = (count == 1 && h_values[0] == 1.0);
    if (p->d_curvlap) {          // the single-pass kernel's parameter block is rebuilt from the new tables
        cudaFree(p->d_curvlap);
        p->d_curvlap = nullptr;
    }
    return NG_OK;
}

int ng_curv_set_axis_square(ng_plan* p, int axis, const ng_plan* sq) {
    int rc = check_curv(p, "ng_curv_set_axis_square");
    if (rc) return rc;
    NG_REQUIRE(axis >= 0 && axis < p->ndim, "ng_curv_set_axis_square: bad axis %d", axis);
    bool ok = sq && sq->kind == PK_FFT && sq->ndim == p->ndim && sq->axis == axis;
    for (int b = 0; ok && b < p->ndim; ++b) ok = sq->shape[b] == p->shape[b];
    NG_REQUIRE(ok, "ng_curv_set_axis_square: a periodic-axis plan of the same shape along axis %d is required", axis);
    p->axis_sq[axis] = sq;
    return NG_OK;
}

int ng_curv_laplacian(ng_plan* p, const double* d_f, double* d_out, void* stream) {
    int rc = check_curv(p, "ng_curv_laplacian");
    if (rc) return rc;
    NG_REQUIRE(d_f && d_out && d_f != d_out, "ng_curv_laplacian: bad buffers");
    NG_NEED(p->cJ, "J");
    for (int i = 0; i < p->ndim; ++i) NG_NEED(p->cLap[i], "J/h_i**2");
    cudaStream_t st = (cudaStream_t)stream;
    {   // all-finite-difference 3-D grids: one pass, `inner` never leaves the SM (curvlap3d.cu)
        bool handled = false;
        rc = ng_curvlap3d_try_launch(p, d_f, d_out, st, &handled);
        if (rc || handled) return rc;
    }
    rc = ensure_work(p);
    if (rc) return rc;
    for (int i = 0; i < p->ndim; ++i) {
        const ng_plan* D = p->axis_plans[i];
        FuseDev F2 = base_fuse(D);
        if (i == 0) F2.add_zero = 1; else F2.addend = d_out;
        if (i == p->ndim - 1) { F2.div = to_dev(p->cJ); F2.finite = 1; }
        if (p->axis_sq[i] && p->cLap[i].stride[i] == 0 && NG_ENV_INT("NG_CURV_SQUARE", 1) != 0) {
            // periodic axis, coefficient constant along it: D(c D f) = c (D D f) -- ONE transform with the squared
            // multiplier, then the multiply as the apply's `post` (a non-finite c gives the same non-finite row)
            F2.post = to_dev(p->cLap[i]);
            rc = ng_apply_any(p->axis_sq[i], d_f, d_out, F2, st);
            if (rc) return rc;
            continue;
        }
        if (ng_fft_twice_supported(D)) {
            // periodic axis: D_i, the multiply and D_i again in one launch -- the pencil pair stays in registers
            F2.mid = to_dev(p->cLap[i]);
            rc = ng_apply_any(D, d_f, d_out, F2, st);
            if (rc) return rc;
            continue;
        }
        FuseDev F1 = base_fuse(D);
        F1.post = to_dev(p->cLap[i]);
        rc = ng_apply_any(D, d_f, p->d_work, F1, st);
        if (rc) return rc;
        rc = ng_apply_any(D, p->d_work, d_out, F2, st);
        if (rc) return rc;
    }
    return NG_OK;
}

int ng_curv_gradient(ng_plan* p, const double* d_f, double* const* d_out, void* stream) {
    int rc = check_curv(p, "ng_curv_gradient");
    if (rc) return rc;
    NG_REQUIRE(d_f && d_out, "ng_curv_gradient: bad buffers");
    for (int i = 0; i < p->ndim; ++i) {
        NG_NEED(p->cGrad[i], "1/h_i");
        NG_REQUIRE(d_out[i] && d_out[i] != d_f, "ng_curv_gradient: bad output %d", i);
    }
    for (int i = 0; i < p->ndim; ++i) {
        const ng_plan* D = p->axis_plans[i];
        FuseDev F = base_fuse(D);
        F.post = to_dev(p->cGrad[i]);
        F.finite = 1;
        rc = ng_apply_any(D, d_f, d_out[i], F, (cudaStream_t)stream);
        if (rc) return rc;
    }
    return NG_OK;
}

int ng_curv_divergence(ng_plan* p, const double* const* d_v, double* d_out, void* stream) {
    int rc = check_curv(p, "ng_curv_divergence");
    if (rc) return rc;
    NG_REQUIRE(d_v && d_out, "ng_curv_divergence: bad buffers");
    NG_NEED(p->cJ, "J");
    for (int i = 0; i < p->ndim; ++i) {
        NG_NEED(p->cDiv[i], "J/h_i");
        NG_REQUIRE(d_v[i] && d_v[i] != d_out, "ng_curv_divergence: bad input %d", i);
    }
    for (int i = 0; i < p->ndim; ++i) {
        const ng_plan* D = p->axis_plans[i];
        FuseDev F = base_fuse(D);
        F.pre = to_dev(p->cDiv[i]);
        if (i == 0) F.add_zero = 1; else F.addend = d_out;
        if (i == p->ndim - 1) { F.div = to_dev(p->cJ); F.finite = 1; }
        rc = ng_apply_any(D, d_v[i], d_out, F, (cudaStream_t)stream);
        if (rc) return rc;
    }
    return NG_OK;
}

int ng_curv_curl(ng_plan* p, const double* const* d_v, double* const* d_out, void* stream) {
    int rc = check_curv(p, "ng_curv_curl");
    if (rc) return rc;
    NG_REQUIRE(d_v && d_out, "ng_curv_curl: bad buffers");
    NG_REQUIRE(p->ndim == 2 || p->ndim == 3, "Curl is only defined for 2D and 3D grids.");
    cudaStream_t st = (cudaStream_t)stream;
    const int ncomp = p->ndim == 2 ? 1 : 3;
    for (int i = 0; i < p->ndim; ++i) NG_NEED(p->cH[i], "h_i");
    for (int c = 0; c < ncomp; ++c) {
        NG_NEED(p->cCurl[c], "1/(h_j h_k)");
        NG_REQUIRE(d_out[c], "ng_curv_curl: NULL output %d", c);
        for (int i = 0; i < p->ndim; ++i)
            NG_REQUIRE(d_v[i] && d_v[i] != d_out[c], "ng_curv_curl: output %d aliases input %d", c, i);
    }
    for (int c = 0; c < ncomp; ++c) {
        // 3-D: (i,j,k) cyclic; 2-D scalar curl: j = 0, k = 1
        const int j = p->ndim == 2 ? 0 : (c + 1) % 3;
        const int k = p->ndim == 2 ? 1 : (c + 2) % 3;
        {   // out_c = D_j(h_k * v_k)
            const ng_plan* D = p->axis_plans[j];
            FuseDev F = base_fuse(D);
            F.pre = to_dev(p->cH[k]);
            rc = ng_apply_any(D, d_v[k], d_out[c], F, st);
            if (rc) return rc;
        }
        {   // out_c = finite( 1/(h_j h_k) * (out_c - D_k(h_j * v_j)) )
            const ng_plan* D = p->axis_plans[k];
            FuseDev F = base_fuse(D);
            F.pre = to_dev(p->cH[j]);
            F.minuend = d_out[c];
            F.post = to_dev(p->cCurl[c]);
            F.finite = 1;
            rc = ng_apply_any(D, d_v[j], d_out[c], F, st);
            if (rc) return rc;
        }
    }
    return NG_OK;
}

}  // extern "C"
