// slab.cu -- one step of the slab-sharded fused Cartesian Laplacian behind the C ABI (SURVEY.md 8b:
// "ng_slab_plan_create"): pack -> neighbour signals -> interior tiles || boundary tiles -> join, the
// sequence of numgrids_b200/slab.py::SlabLaplacian._call_symm, with its own side stream, events and
// signal kernels.  The reference is single-process (no counterpart); one process per GPU.
//
// What the library does NOT do is map peer memory: the caller hands in device pointers that are valid
// in this process -- this rank's halo buffers and flag words, and the neighbours' buffers and flag words
// mapped here (CUDA IPC, cuMem fabric handles, or torch symmetric memory as in slab.py).  No NCCL, no
// MPI: the only inter-GPU traffic is the pack kernel's stores into the neighbours' halo buffers and one
// 8-byte flag per neighbour and step, both over NVLink.
//
// Signals are step counters (never reset): at step s a rank stores s into the flag word it owns in each
// neighbour ("from my upper neighbour" / "from my lower neighbour") after a system-scope fence behind its
// pack kernel, and waits until both of its own flag words have reached s.  Two halo sets alternate; a
// neighbour overwrites set q again at step s+2, after it has seen my flag of step s+1, which I stored
// after my boundary launch of step s (the last reader of set q) was enqueued ahead of it.
#include "common.cuh"

namespace {

struct SlabState {
    ng_plan* lap = nullptr;              // the local fused-Laplacian plan (owned)
    int rank = 0, nranks = 1;
    double* my_halo[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};   // [set][0 = lo halo, 1 = hi halo]
    double* peer_lo_halo[2] = {nullptr, nullptr};   // [set] lower neighbour's HI halo, mapped here
    double* peer_hi_halo[2] = {nullptr, nullptr};   // [set] upper neighbour's LO halo, mapped here
    unsigned long long* my_flags = nullptr;         // [2]: 0 written by the lower neighbour, 1 by the upper one
    unsigned long long* peer_lo_flag = nullptr;     // lower neighbour's flag word 1, mapped here
    unsigned long long* peer_hi_flag = nullptr;     // upper neighbour's flag word 0, mapped here
    unsigned long long step = 0;
    cudaStream_t side = nullptr;
    cudaEvent_t packed = nullptr, done = nullptr;
    int* d_timeout = nullptr;                       // set by a wait that gave up
};

__global__ void slab_signal_kernel(unsigned long long* lo, unsigned long long* hi, unsigned long long v) {
    if (threadIdx.x == 0) {
        __threadfence_system();                     // the pack kernel's peer stores are visible before the flag
        if (lo) asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(lo), "l"(v) : "memory");
        if (hi) asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(hi), "l"(v) : "memory");
    }
}

// Bounded: a neighbour that never signals (crashed rank, mismatched step counts) makes this kernel give up after
// ~10 s and flag the timeout instead of hanging the GPU.
__global__ void slab_wait_kernel(const unsigned long long* a, const unsigned long long* b, unsigned long long v,
                                 int* timeout) {
    if (threadIdx.x != 0) return;
    const unsigned long long* w[2] = {a, b};
    for (int k = 0; k < 2; ++k) {
        if (!w[k]) continue;
        unsigned long long got = 0;
        long long t0 = clock64();
        for (;;) {
            asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(got) : "l"(w[k]) : "memory");
            if (got >= v) break;
            if (clock64() - t0 > 20000000000LL) { *timeout = 1; return; }
            __nanosleep(200);
        }
    }
}

}  // namespace

extern "C" {

int ng_slab_plan_create(const ng_slab_comm* comm, int ndim, const int64_t* local_shape, ng_plan* const* axis_plans,
                        ng_plan** out) {
    NG_REQUIRE(out != nullptr, "out is NULL");
    *out = nullptr;
    NG_REQUIRE(comm && comm->nranks >= 1 && comm->rank >= 0 && comm->rank < comm->nranks, "ng_slab_plan_create: bad communicator record");
    NG_REQUIRE(ndim == 3, "ng_slab_plan_create: 3-D grids (the slab step of the fused Laplacian)");
    const bool has_lo = comm->rank > 0, has_hi = comm->rank < comm->nranks - 1;
    for (int q = 0; q < 2; ++q) {
        NG_REQUIRE(!has_lo || (comm->my_halo[q][0] && comm->peer_lo_halo[q]), "ng_slab_plan_create: missing lower-neighbour buffers (set %d)", q);
        NG_REQUIRE(!has_hi || (comm->my_halo[q][1] && comm->peer_hi_halo[q]), "ng_slab_plan_create: missing upper-neighbour buffers (set %d)", q);
    }
    NG_REQUIRE(comm->nranks == 1 || comm->my_flags, "ng_slab_plan_create: flag words missing");
    NG_REQUIRE(!has_lo || comm->peer_lo_flag, "ng_slab_plan_create: lower neighbour's flag word missing");
    NG_REQUIRE(!has_hi || comm->peer_hi_flag, "ng_slab_plan_create: upper neighbour's flag word missing");
    ng_plan* lap = nullptr;
    int rc = ng_fdlap_plan_create(ndim, local_shape, axis_plans, &lap);
    if (rc) return rc;
    SlabState* S = new (std::nothrow) SlabState();
    ng_plan* p = new (std::nothrow) ng_plan();
    if (!S || !p) { delete S; delete p; ng_plan_destroy(lap); ng_set_error("out of host memory"); return NG_ENOMEM; }
    S->lap = lap;
    S->rank = comm->rank;
    S->nranks = comm->nranks;
    for (int q = 0; q < 2; ++q) {
        S->my_halo[q][0] = has_lo ? comm->my_halo[q][0] : nullptr;
        S->my_halo[q][1] = has_hi ? comm->my_halo[q][1] : nullptr;
        S->peer_lo_halo[q] = has_lo ? comm->peer_lo_halo[q] : nullptr;
        S->peer_hi_halo[q] = has_hi ? comm->peer_hi_halo[q] : nullptr;
    }
    S->my_flags = (unsigned long long*)comm->my_flags;
    S->peer_lo_flag = has_lo ? (unsigned long long*)comm->peer_lo_flag : nullptr;
    S->peer_hi_flag = has_hi ? (unsigned long long*)comm->peer_hi_flag : nullptr;
    p->kind = PK_SLAB;
    p->ndim = ndim;
    p->size = lap->size;
    for (int a = 0; a < ndim; ++a) p->shape[a] = local_shape[a];
    p->device = lap->device;
    p->slab = S;
    cudaError_t e = cudaStreamCreateWithFlags(&S->side, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&S->packed, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&S->done, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaMalloc((void**)&S->d_timeout, sizeof(int));
    if (e == cudaSuccess) e = cudaMemset(S->d_timeout, 0, sizeof(int));
    if (e != cudaSuccess) {
        ng_set_error("ng_slab_plan_create: %s", cudaGetErrorString(e));
        ng_plan_destroy(p);
        return NG_ECUDA;
    }
    *out = p;
    return NG_OK;
}

int ng_slab_lap_apply(ng_plan* plan, const double* d_f, double* d_out, void* stream) {
    NG_REQUIRE(plan && plan->kind == PK_SLAB && plan->slab, "ng_slab_lap_apply: slab plan required");
    NG_REQUIRE(d_f && d_out && d_f != d_out, "ng_slab_lap_apply: bad buffers");
    int rc = ng_check_plan_device(plan, "ng_slab_lap_apply");
    if (rc) return rc;
    SlabState* S = (SlabState*)plan->slab;
    cudaStream_t main = (cudaStream_t)stream;
    if (S->nranks == 1) return ng_fdlap_apply(S->lap, d_f, d_out, stream);
    const unsigned long long step = ++S->step;
    const int q = (int)(step & 1);
    // main stream: edge columns straight into the neighbours' halo buffers (before the stencil: slab.py explains why)
    rc = ng_fd_pack_halo(S->lap, d_f, S->peer_lo_halo[q], S->peer_hi_halo[q], main);
    if (rc) return rc;
    NG_CUDA_OK(cudaEventRecord(S->packed, main));
    // side stream: signals, then the first / last z tile with the received columns
    NG_CUDA_OK(cudaStreamWaitEvent(S->side, S->packed, 0));
    slab_signal_kernel<<<1, 32, 0, S->side>>>(S->peer_lo_flag, S->peer_hi_flag, step);
    NG_LAUNCH_CHECK();
    slab_wait_kernel<<<1, 32, 0, S->side>>>(S->peer_lo_flag ? S->my_flags + 0 : nullptr,
                                            S->peer_hi_flag ? S->my_flags + 1 : nullptr, step, S->d_timeout);
    NG_LAUNCH_CHECK();
    rc = ng_fdlap_apply_slab_part(S->lap, d_f, d_out, S->my_halo[q][0], S->my_halo[q][1], 2, S->side);
    if (rc) return rc;
    NG_CUDA_OK(cudaEventRecord(S->done, S->side));
    // main stream: interior z tiles (no halo needed), then join
    rc = ng_fdlap_apply_slab_part(S->lap, d_f, d_out, nullptr, nullptr, 1, main);
    if (rc) return rc;
    NG_CUDA_OK(cudaStreamWaitEvent(main, S->done, 0));
    return NG_OK;
}

int ng_slab_timed_out(ng_plan* plan) {
    if (!plan || plan->kind != PK_SLAB || !plan->slab) return 0;
    SlabState* S = (SlabState*)plan->slab;
    int h = 0;
    if (!S->d_timeout || cudaMemcpy(&h, S->d_timeout, sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess) return 0;
    return h;
}

}  // extern "C"

void ng_slab_state_destroy(void* state) {
    SlabState* S = (SlabState*)state;
    if (!S) return;
    if (S->side) { cudaStreamSynchronize(S->side); cudaStreamDestroy(S->side); }
    if (S->packed) cudaEventDestroy(S->packed);
    if (S->done) cudaEventDestroy(S->done);
    if (S->d_timeout) cudaFree(S->d_timeout);
    if (S->lap) ng_plan_destroy(S->lap);
    delete S;
}
