/*
 * picles1d.cu — kernels and C ABI (picles1d_* of include/picles_b200.h) of the reference's ONE-DIMENSIONAL
 * model, WaveGrowth1D.  Arithmetic: physics1d.h.  Three kernels per model step, as on the 2-D path:
 *   k1d_advance          advance! for every particle (one thread each): adaptive RK over DT or the off/boundary
 *                        branches, the NaN/Inf/e_max resets, and the deposit record (charge, floor node, weights)
 *   k1d_project_remesh   per node: the charges that land on it merged in the reference's order (merge! is not a
 *                        sum: the order matters), State, and NodeToParticle! for the node's particle
 *   k1d_seed             SeedParticle!
 * One-dimensional grids are small; nothing here is tuned beyond coalesced SoA planes.  No CPU fallback.
 */
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <vector>

#include "../../include/picles_b200.h"
#include "physics1d.h"

using namespace picles1d;

namespace {

__device__ __forceinline__ void load_particle(const Arrays& A, int i, Particle1& p) {
    p.u[0] = A.z0[i]; p.u[1] = A.z1[i]; p.u[2] = A.z2[i];
    p.t = A.t[i]; p.dt = A.dt[i]; p.qold = A.qold[i];
    p.iter = A.iter[i]; p.flags = A.flags[i]; p.status = A.status[i];
}
__device__ __forceinline__ void store_particle(const Arrays& A, int i, const Particle1& p) {
    A.z0[i] = p.u[0]; A.z1[i] = p.u[1]; A.z2[i] = p.u[2];
    A.t[i] = p.t; A.dt[i] = p.dt; A.qold[i] = p.qold;
    A.iter[i] = p.iter; A.flags[i] = p.flags; A.status[i] = p.status;
}
__device__ void tally_flush(const Tally1& c, Counters* dc) {
    if (c.integrated) atomicAdd(&dc->n_integrated, (unsigned long long)c.integrated);
    if (c.substeps) atomicAdd(&dc->n_substeps, (unsigned long long)c.substeps);
    if (c.rejects) atomicAdd(&dc->n_rejects, (unsigned long long)c.rejects);
    if (c.rhs) atomicAdd(&dc->n_rhs, (unsigned long long)c.rhs);
    if (c.reseed) atomicAdd(&dc->n_reseed_advance, (unsigned long long)c.reseed);
    if (c.fixups) atomicAdd(&dc->n_fixups, (unsigned long long)c.fixups);
    if (c.failed) atomicAdd(&dc->n_failed, (unsigned long long)c.failed);
    if (c.deposited) atomicAdd(&dc->n_deposited, (unsigned long long)c.deposited);
    if (c.A) atomicAdd(&dc->n_A, (unsigned long long)c.A);
    if (c.B) atomicAdd(&dc->n_B, (unsigned long long)c.B);
    if (c.D) atomicAdd(&dc->n_D, (unsigned long long)c.D);
    if (c.reach) atomicMax(&dc->reach, c.reach);
    if (c.max_attempts) atomicMax(&dc->max_attempts, c.max_attempts);
}

__global__ void __launch_bounds__(128) k1d_seed(Arrays A, picles_params_t P, const double* __restrict__ u0) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= A.Nx) return;
    Particle1 p;
    double s[3];
    p1_seed(A, P, i, u0[i], p, s);
    store_particle(A, i, p);
    A.S[i] = s[0]; A.S[i + A.Nx] = s[1]; A.S[i + 2 * (int64_t)A.Nx] = s[2];
    A.r_ifl[i] = P1_NO_DEPOSIT;
}

__global__ void __launch_bounds__(128) k1d_advance(Arrays A, picles_params_t P, double DT, Counters* dc) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    Tally1 c;
    memset(&c, 0, sizeof c);
    if (i < A.Nx) {
        Particle1 p;
        load_particle(A, i, p);
        double e, m, wf, wc;
        int64_t ifl;
        p1_advance(A, P, i, DT, p, c, e, m, wf, wc, ifl);
        store_particle(A, i, p);
        A.r_e[i] = e; A.r_m[i] = m; A.r_wf[i] = wf; A.r_wc[i] = wc; A.r_ifl[i] = ifl;
    }
    tally_flush(c, dc);
}

__global__ void __launch_bounds__(128) k1d_project_remesh(Arrays A, picles_params_t P, double DT, Counters* dc) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    Tally1 c;
    memset(&c, 0, sizeof c);
    if (i < A.Nx) {
        const int R = dc->reach; /* written by k1d_advance of this step (same stream) */
        double g[3];
        p1_gather_node(A, P.periodic_boundary, R, (int64_t)i + 1, g);
        A.S[i] = g[0]; A.S[i + A.Nx] = g[1]; A.S[i + 2 * (int64_t)A.Nx] = g[2];
        Particle1 p;
        load_particle(A, i, p);
        p1_remesh(A, P, i, DT, g, A.w0[i], p, c);
        store_particle(A, i, p);
    }
    tally_flush(c, dc);
}

} /* namespace */

/* ---- C ABI --------------------------------------------------------------------------------------------------- */
struct picles1d_handle {
    int device = -1;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[3] = {nullptr, nullptr, nullptr};
    bool have_grid = false, have_params = false, seeded = false;
    Arrays A = {};
    picles_params_t P;
    Counters* d_counters = nullptr;
    double *d_xn = nullptr, *d_w0 = nullptr, *d_w1 = nullptr;
    std::vector<void*> allocs;
    picles_counters_t last;
    char err[512];
};

static thread_local char g_err1[512] = "";
static int fail1(picles1d_t* h, int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (h) snprintf(h->err, sizeof h->err, "%s", buf);
    snprintf(g_err1, sizeof g_err1, "%s", buf);
    return code;
}
#define CK1(call)                                                                                               \
    do {                                                                                                        \
        cudaError_t e_ = (call);                                                                                \
        if (e_ != cudaSuccess)                                                                                  \
            return fail1(h, PICLES_ERR_CUDA, "%s failed at %s:%d: %s", #call, __FILE__, __LINE__, cudaGetErrorString(e_)); \
    } while (0)
template <class T>
static int dalloc1(picles1d_t* h, T** p, int64_t count) {
    void* q = nullptr;
    cudaError_t e = cudaMalloc(&q, (size_t)(count > 0 ? count : 1) * sizeof(T));
    if (e != cudaSuccess) return fail1(h, PICLES_ERR_ALLOC, "cudaMalloc: %s", cudaGetErrorString(e));
    h->allocs.push_back(q);
    *p = (T*)q;
    return 0;
}
#define DALLOC1(ptr, count)                    \
    do {                                       \
        int rc_ = dalloc1(h, &(ptr), (count)); \
        if (rc_) return rc_;                   \
    } while (0)

extern "C" {

const char* picles1d_last_error(picles1d_t* h) { return h ? h->err : g_err1; }

int picles1d_create(picles1d_t** out, int device_id) {
    if (!out) return fail1(nullptr, PICLES_ERR_ARG, "picles1d_create: out is NULL");
    *out = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) return fail1(nullptr, PICLES_ERR_CUDA, "no CUDA device: this library has no CPU fallback");
    if (device_id < 0 || device_id >= n) return fail1(nullptr, PICLES_ERR_ARG, "device %d out of range (%d devices)", device_id, n);
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device_id) != cudaSuccess || prop.major < 10)
        return fail1(nullptr, PICLES_ERR_CUDA, "device %d is not an sm_100-class GPU", device_id);
    picles1d_t* h = new picles1d_handle();
    h->device = device_id;
    h->err[0] = 0;
    memset(&h->last, 0, sizeof h->last);
    cudaError_t e = cudaSetDevice(device_id);
    if (e == cudaSuccess) e = cudaStreamCreate(&h->stream);
    for (int k = 0; k < 3 && e == cudaSuccess; k++) e = cudaEventCreate(&h->ev[k]);
    void* dc = nullptr;
    if (e == cudaSuccess) e = cudaMalloc(&dc, sizeof(Counters));
    if (e != cudaSuccess) {
        int rc = fail1(nullptr, PICLES_ERR_CUDA, "picles1d_create: %s", cudaGetErrorString(e));
        picles1d_destroy(h);
        return rc;
    }
    h->d_counters = (Counters*)dc;
    *out = h;
    return PICLES_OK;
}

int picles1d_destroy(picles1d_t* h) {
    if (!h) return PICLES_OK;
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    for (void* p : h->allocs) cudaFree(p);
    if (h->d_counters) cudaFree(h->d_counters);
    for (int k = 0; k < 3; k++) if (h->ev[k]) cudaEventDestroy(h->ev[k]);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
    return PICLES_OK;
}

int picles1d_set_grid(picles1d_t* h, int Nx, double xmin, double dx, const double* x_nodes) {
    if (!h) return fail1(nullptr, PICLES_ERR_ARG, "null handle");
    if (h->have_grid) return fail1(h, PICLES_ERR_STATE, "grid already set");
    if (Nx < 2 || !x_nodes || !(dx > 0.0)) return fail1(h, PICLES_ERR_ARG, "picles1d_set_grid: Nx >= 2, dx > 0 and node coordinates are required");
    CK1(cudaSetDevice(h->device));
    Arrays& A = h->A;
    A.Nx = Nx; A.xmin = xmin; A.dx = dx;
    DALLOC1(h->d_xn, Nx); DALLOC1(h->d_w0, Nx); DALLOC1(h->d_w1, Nx);
    A.xn = h->d_xn; A.w0 = h->d_w0; A.w1 = h->d_w1;
    DALLOC1(A.z0, Nx); DALLOC1(A.z1, Nx); DALLOC1(A.z2, Nx);
    DALLOC1(A.t, Nx); DALLOC1(A.dt, Nx); DALLOC1(A.qold, Nx);
    DALLOC1(A.iter, Nx); DALLOC1(A.flags, Nx); DALLOC1(A.status, Nx);
    DALLOC1(A.S, 3 * (int64_t)Nx);
    DALLOC1(A.r_e, Nx); DALLOC1(A.r_m, Nx); DALLOC1(A.r_wf, Nx); DALLOC1(A.r_wc, Nx); DALLOC1(A.r_ifl, Nx);
    CK1(cudaMemcpyAsync(h->d_xn, x_nodes, sizeof(double) * Nx, cudaMemcpyHostToDevice, h->stream));
    CK1(cudaStreamSynchronize(h->stream));
    h->have_grid = true;
    return PICLES_OK;
}

int picles1d_set_params(picles1d_t* h, const picles_params_t* P) {
    if (!h || !P) return fail1(h, PICLES_ERR_ARG, "null argument");
    if (!P->adaptive) return fail1(h, PICLES_ERR_ARG, "adaptive = false is not supported");
    if (P->solver != PICLES_SOLVER_TSIT5 && P->solver != PICLES_SOLVER_DP5 && P->solver != PICLES_SOLVER_AUTOTSIT5)
        return fail1(h, PICLES_ERR_ARG, "unknown solver id %d", (int)P->solver);
    if (P->has_defaults)
        return fail1(h, PICLES_ERR_ARG, "the 1-D path seeds from the wind sea only (ODEinit_type = \"wind_sea\"): ParticleDefaults would put every "
                                        "particle at defaults.x (core_1D.jl:215-218)");
    h->P = *P;
    /* AutoTsit5 on the 1-D path runs as Tsit5: the switch to Rosenbrock23 is not modelled here */
    if (h->P.solver == PICLES_SOLVER_AUTOTSIT5) h->P.solver = PICLES_SOLVER_TSIT5;
    h->have_params = true;
    return PICLES_OK;
}

int picles1d_seed(picles1d_t* h, const double* u0) {
    if (!h || !u0) return fail1(h, PICLES_ERR_ARG, "null argument");
    if (!h->have_grid || !h->have_params) return fail1(h, PICLES_ERR_STATE, "picles1d_seed: grid and parameters first");
    CK1(cudaSetDevice(h->device));
    const int Nx = h->A.Nx;
    CK1(cudaMemcpyAsync(h->d_w0, u0, sizeof(double) * Nx, cudaMemcpyHostToDevice, h->stream));
    k1d_seed<<<(Nx + 127) / 128, 128, 0, h->stream>>>(h->A, h->P, h->d_w0);
    CK1(cudaGetLastError());
    CK1(cudaStreamSynchronize(h->stream));
    memset(&h->last, 0, sizeof h->last);
    h->seeded = true;
    return PICLES_OK;
}

int picles1d_step(picles1d_t* h, double t, double dt_model, const double* u_t, const double* u_t1) {
    (void)t;
    if (!h || !u_t || !u_t1) return fail1(h, PICLES_ERR_ARG, "null argument");
    if (!h->seeded) return fail1(h, PICLES_ERR_STATE, "picles1d_step: seed first");
    if (!(dt_model > 0.0)) return fail1(h, PICLES_ERR_ARG, "dt_model must be positive");
    CK1(cudaSetDevice(h->device));
    const int Nx = h->A.Nx;
    const int g = (Nx + 127) / 128;
    CK1(cudaMemcpyAsync(h->d_w0, u_t, sizeof(double) * Nx, cudaMemcpyHostToDevice, h->stream));
    CK1(cudaMemcpyAsync(h->d_w1, u_t1, sizeof(double) * Nx, cudaMemcpyHostToDevice, h->stream));
    CK1(cudaMemsetAsync(h->d_counters, 0, sizeof(Counters), h->stream));
    CK1(cudaEventRecord(h->ev[0], h->stream));
    k1d_advance<<<g, 128, 0, h->stream>>>(h->A, h->P, dt_model, h->d_counters);
    CK1(cudaEventRecord(h->ev[1], h->stream));
    k1d_project_remesh<<<g, 128, 0, h->stream>>>(h->A, h->P, dt_model, h->d_counters);
    CK1(cudaEventRecord(h->ev[2], h->stream));
    CK1(cudaGetLastError());
    Counters c;
    CK1(cudaMemcpyAsync(&c, h->d_counters, sizeof c, cudaMemcpyDeviceToHost, h->stream));
    CK1(cudaStreamSynchronize(h->stream));
    picles_counters_t& L = h->last;
    memset(&L, 0, sizeof L);
    L.n_active = Nx;
    L.n_integrated = (int64_t)c.n_integrated; L.n_substeps = (int64_t)c.n_substeps; L.n_rejects = (int64_t)c.n_rejects;
    L.n_rhs = (int64_t)c.n_rhs; L.n_reseed_advance = (int64_t)c.n_reseed_advance; L.n_fixups = (int64_t)c.n_fixups;
    L.n_failed = (int64_t)c.n_failed; L.n_deposited = (int64_t)c.n_deposited;
    L.n_remesh_A = (int64_t)c.n_A; L.n_remesh_B = (int64_t)c.n_B; L.n_remesh_D = (int64_t)c.n_D;
    L.reach = c.reach; L.max_attempts = c.max_attempts;
    float ms = 0.f;
    CK1(cudaEventElapsedTime(&ms, h->ev[0], h->ev[1])); L.ms_advance = ms;
    CK1(cudaEventElapsedTime(&ms, h->ev[1], h->ev[2])); L.ms_project = ms;
    return PICLES_OK;
}

int picles1d_get_state(picles1d_t* h, double* S) {
    if (!h || !S) return fail1(h, PICLES_ERR_ARG, "null argument");
    if (!h->seeded) return fail1(h, PICLES_ERR_STATE, "seed first");
    CK1(cudaSetDevice(h->device));
    CK1(cudaMemcpy(S, h->A.S, sizeof(double) * 3 * h->A.Nx, cudaMemcpyDeviceToHost));
    return PICLES_OK;
}

int picles1d_get_particles(picles1d_t* h, double* z, double* t, double* dt, uint8_t* flags, int32_t* status) {
    if (!h || !z) return fail1(h, PICLES_ERR_ARG, "null argument");
    if (!h->seeded) return fail1(h, PICLES_ERR_STATE, "seed first");
    CK1(cudaSetDevice(h->device));
    const int Nx = h->A.Nx;
    CK1(cudaMemcpy(z, h->A.z0, sizeof(double) * Nx, cudaMemcpyDeviceToHost));
    CK1(cudaMemcpy(z + Nx, h->A.z1, sizeof(double) * Nx, cudaMemcpyDeviceToHost));
    CK1(cudaMemcpy(z + 2 * (int64_t)Nx, h->A.z2, sizeof(double) * Nx, cudaMemcpyDeviceToHost));
    if (t) CK1(cudaMemcpy(t, h->A.t, sizeof(double) * Nx, cudaMemcpyDeviceToHost));
    if (dt) CK1(cudaMemcpy(dt, h->A.dt, sizeof(double) * Nx, cudaMemcpyDeviceToHost));
    if (flags) CK1(cudaMemcpy(flags, h->A.flags, Nx, cudaMemcpyDeviceToHost));
    if (status) CK1(cudaMemcpy(status, h->A.status, sizeof(int32_t) * Nx, cudaMemcpyDeviceToHost));
    return PICLES_OK;
}

int picles1d_get_counters(picles1d_t* h, picles_counters_t* c) {
    if (!h || !c) return fail1(h, PICLES_ERR_ARG, "null argument");
    *c = h->last;
    return PICLES_OK;
}

} /* extern "C" */
