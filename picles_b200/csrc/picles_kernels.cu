/*
 * picles_kernels.cu — sm_100a kernels of the PiCLES particle-in-cell step.
 *
 *   k_seed      SeedParticle for every node                      (run.jl:199-247, core_2D.jl:434-488)
 *   k_advance   advance! : adaptive RK over DT + deposit record  (mapping_2D.jl:118-243)  FP64-bound
 *   k_project   ParticleToNode! as a deterministic gather         (mapping_2D.jl:59-73,
 *                                                                  ParticleInCell.jl:341-538) HBM-bound
 *   k_remesh    remesh!/NodeToParticle!                           (mapping_2D.jl:250-356)  HBM-bound
 *   k_energy    sum of State[:,:,1]                               (run.jl:23-25)
 *
 * Layout in HBM (one y-strip per GPU): every per-node quantity is its own plane of
 * ny*Nx doubles with i (x) fastest — the memory order of the reference's column-major
 * (Nx,Ny[,3]) arrays — so a warp touches 32 consecutive doubles (256 B) per plane.
 * Deposit records carry `halo` extra rows on both sides for the neighbour strips.
 *
 * Compiled with --fmad=false: physics.h spells out every fused multiply-add so the
 * results are bit-identical to the CPU oracle.
 */
#include <cuda_runtime.h>
#include <stdint.h>

#include "physics.h"
#include "picles_device.h"

namespace picles {

static_assert(PH_REACH_MAX == PH_REACH_MAX_ABI, "reach limits out of sync");

/* ---- block-level tally reduction: one atomic set per block ------------------- */
__device__ __forceinline__ int32_t warp_sum(int32_t v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ int32_t warp_max(int32_t v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_down_sync(0xffffffffu, v, o));
    return v;
}

#define TALLY_NSUM 12
__device__ void tally_flush(const Tally& c, DeviceCounters* dc) {
    __shared__ int32_t s_sum[TALLY_NSUM];
    __shared__ int32_t s_max[2];
    if (threadIdx.x < TALLY_NSUM) s_sum[threadIdx.x] = 0;
    if (threadIdx.x < 2) s_max[threadIdx.x] = 0;
    __syncthreads();
    int32_t v[TALLY_NSUM] = {c.integrated, c.substeps, c.rejects, c.rhs, c.reseed, c.fixups,
                             c.failed, c.deposited, c.A, c.B, c.C, c.D};
    int lane = threadIdx.x & 31;
#pragma unroll
    for (int k = 0; k < TALLY_NSUM; k++) {
        int32_t s = warp_sum(v[k]);
        if (lane == 0 && s) atomicAdd(&s_sum[k], s);
    }
    int32_t r = warp_max(c.reach), a = warp_max(c.max_attempts);
    if (lane == 0) {
        atomicMax(&s_max[0], r);
        atomicMax(&s_max[1], a);
    }
    __syncthreads();
    if (threadIdx.x < TALLY_NSUM && s_sum[threadIdx.x])
        atomicAdd(&dc->sums[threadIdx.x], (unsigned long long)s_sum[threadIdx.x]);
    if (threadIdx.x == 0) {
        if (s_max[0]) atomicMax(&dc->reach, s_max[0]);
        if (s_max[1]) atomicMax(&dc->max_attempts, s_max[1]);
    }
}

__device__ __forceinline__ void load_particle(const DeviceArrays& A, int64_t l, Particle& p) {
    p.u0 = A.z[0][l]; p.u1 = A.z[1][l]; p.u2 = A.z[2][l]; p.u3 = A.z[3][l]; p.u4 = A.z[4][l];
    p.t = A.t[l]; p.dt = A.dt[l]; p.qold = A.qold[l];
    p.iter = A.iter[l];
    p.flags = A.flags[l];
    p.status = A.status[l];
}
__device__ __forceinline__ void store_particle(const DeviceArrays& A, int64_t l, const Particle& p) {
    A.z[0][l] = p.u0; A.z[1][l] = p.u1; A.z[2][l] = p.u2; A.z[3][l] = p.u3; A.z[4][l] = p.u4;
    A.t[l] = p.t; A.dt[l] = p.dt; A.qold[l] = p.qold;
    A.iter[l] = p.iter;
    A.flags[l] = p.flags;
    A.status[l] = p.status;
}
__device__ __forceinline__ void store_record(const DeviceArrays& A, int64_t le, const Record& r) {
    A.rec[0][le] = r.e; A.rec[1][le] = r.mx; A.rec[2][le] = r.my; A.rec[3][le] = r.wxc; A.rec[4][le] = r.wyc;
    A.cell[le] = r.cell;
}

/* ---- seed ------------------------------------------------------------------ */
__global__ void __launch_bounds__(256) k_seed(DeviceArrays A, picles_params_t P, const double* __restrict__ u0,
                                              const double* __restrict__ v0) {
    int64_t n = (int64_t)A.Nx * A.ny;
    for (int64_t l = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; l < n; l += (int64_t)gridDim.x * blockDim.x) {
        Particle p;
        double e, mx, my;
        seed_particle(P, A.mask[l], u0[l], v0[l], p, e, mx, my);
        store_particle(A, l, p);
        A.S[0][l] = e; A.S[1][l] = mx; A.S[2][l] = my;
        Record r;
        r.e = r.mx = r.my = r.wxc = r.wyc = 0.0;
        r.cell = PH_CELL_INVALID;
        store_record(A, l + (int64_t)A.halo * A.Nx, r);
    }
}

/* ---- advance ---------------------------------------------------------------- */
/*
 * One thread per particle, grid-stride over the strip so a block owns many 32-particle
 * chunks and flushes its counters once.  The adaptive loop runs in SIMT lock-step: lanes
 * whose particle reached t+DT wait at the loop exit for the slowest lane of the warp
 * (neighbouring nodes carry near-identical states, so attempt counts are close).
 */
template <bool PER_NODE_M>
__global__ void __launch_bounds__(ADV_THREADS, ADV_MIN_BLOCKS)
k_advance(DeviceArrays A, picles_params_t P, double DT, DeviceCounters* dc) {
    /* stage derivatives k_j[0:3], j = 1..7: 21 doubles per thread, one column per thread
       (consecutive threads -> consecutive 8-byte words: conflict-free) */
    __shared__ double s_k[21 * ADV_THREADS];
    KStrided K;
    K.base = &s_k[threadIdx.x];
    K.stride = ADV_THREADS;
    Tally c;
    tally_zero(c);
    int64_t n = (int64_t)A.Nx * A.ny;
    for (int64_t l = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; l < n; l += (int64_t)gridDim.x * blockDim.x) {
        int64_t le = l + (int64_t)A.halo * A.Nx;
        uint8_t flags = A.flags[l];
        if (!(flags & PICLES_PF_ACTIVE)) continue; /* record stays invalid (set at seed) */
        Particle p;
        load_particle(A, l, p);
        double M[4];
        if (PER_NODE_M) { M[0] = A.M[0][l]; M[1] = A.M[1][l]; M[2] = A.M[2][l]; M[3] = A.M[3][l]; }
        else { M[0] = A.Mc[0]; M[1] = A.Mc[1]; M[2] = A.Mc[2]; M[3] = A.Mc[3]; }
        double pc = A.pc ? A.pc[l] : 0.0;
        Record r;
        advance_particle(P, p, A.mask[l], DT, A.u_t[l], A.v_t[l], A.u_t1[l], A.v_t1[l], M, pc, r, c, K);
        store_particle(A, l, p);
        store_record(A, le, r);
    }
    tally_flush(c, dc);
}

/* ---- projection gather --------------------------------------------------------- */
/*
 * One thread per target node; the per-target arithmetic (fast interior path and the
 * generic wrap/fold path) is gather_node() in physics.h so the CPU tests can run the
 * very same code.  Neighbouring threads read overlapping record windows, which L1/L2
 * serve; HBM sees each record once.
 */
__global__ void __launch_bounds__(PRJ_THREADS) k_project(DeviceArrays A, int n_classes, int accumulate, const DeviceCounters* __restrict__ dc) {
    RecView V;
    V.Nx = A.Nx; V.Ny = A.Ny; V.bx = A.bx; V.by = A.by; V.j0 = A.j0; V.ny = A.ny; V.halo = A.halo;
    V.e = A.rec[0]; V.mx = A.rec[1]; V.my = A.rec[2]; V.wx = A.rec[3]; V.wy = A.rec[4];
    V.cell = A.cell;
    /* deposits landing on this strip come from its own particles (reach) and from the
       neighbours' rows received into the halo (reach_halo) */
    int R = min(max(dc->reach, dc->reach_halo), PH_REACH_MAX);
    if (A.ny != A.Ny) R = min(R, A.halo); /* strips: the host rejects reach > halo (PICLES_ERR_HALO) */
    /* 2-D launch: blockIdx.y strides rows, threads run along x (no integer division) */
    for (int jr = blockIdx.y; jr < A.ny; jr += gridDim.y) {
        const int J = jr + 1 + A.j0; /* global 1-based target row */
        for (int i0 = blockIdx.x * blockDim.x + threadIdx.x; i0 < A.Nx; i0 += gridDim.x * blockDim.x) {
            const int I = i0 + 1;
            const int64_t l = (int64_t)jr * A.Nx + i0;
            double s0 = 0.0, s1 = 0.0, s2 = 0.0;
            if (accumulate) { s0 = A.S[0][l]; s1 = A.S[1][l]; s2 = A.S[2][l]; }
            gather_node(V, I, J, R, n_classes, s0, s1, s2);
            A.S[0][l] = s0;
            A.S[1][l] = s1;
            A.S[2][l] = s2;
        }
    }
}

/* ---- remesh -------------------------------------------------------------------- */
__global__ void __launch_bounds__(RMS_THREADS) k_remesh(DeviceArrays A, picles_params_t P, double DT, DeviceCounters* dc) {
    Tally c;
    tally_zero(c);
    int64_t n = (int64_t)A.Nx * A.ny;
    for (int64_t l = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; l < n; l += (int64_t)gridDim.x * blockDim.x) {
        uint8_t flags = A.flags[l];
        if (!(flags & PICLES_PF_ACTIVE)) continue;
        double e = A.S[0][l], mx = A.S[1][l], my = A.S[2][l];
        double wu = A.u_t[l], wv = A.v_t[l];
        bool boundary = (flags & PICLES_PF_BOUNDARY) != 0;
        bool enough = (e >= P.minimal_state[0]) && (mx * mx + my * my >= P.minimal_state[1]);
        bool windy = (wu * wu + wv * wv >= P.wind_min_squared);
        Particle p;
        if (!boundary && enough) {
            /* branch A touches only u, flags: skip the loads it does not need */
            p.flags = flags; p.status = 0; p.iter = 0; p.qold = 0.0; p.t = 0.0; p.dt = 0.0;
            remesh_particle(P, p, e, mx, my, wu, wv, DT, c);
            A.z[0][l] = p.u0; A.z[1][l] = p.u1; A.z[2][l] = p.u2; A.z[3][l] = p.u3; A.z[4][l] = p.u4;
            A.flags[l] = p.flags;
        } else if (windy) {
            load_particle(A, l, p);
            remesh_particle(P, p, e, mx, my, wu, wv, DT, c);
            store_particle(A, l, p);
        } else {
            p.flags = flags;
            remesh_particle(P, p, e, mx, my, wu, wv, DT, c);
            if (p.flags != flags) A.flags[l] = p.flags;
        }
    }
    tally_flush(c, dc);
}

/* ---- energy sum (deterministic two-stage) ---------------------------------------- */
__global__ void __launch_bounds__(256) k_energy(const double* __restrict__ e, int64_t n, double* __restrict__ partial) {
    __shared__ double sh[256];
    double s = 0.0;
    for (int64_t l = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; l < n; l += (int64_t)gridDim.x * blockDim.x) s += e[l];
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) partial[blockIdx.x] = sh[0];
}

/* ---- halo pack / unpack: H rows of the 5 record planes + cell plane -------------- */
__global__ void k_halo_pack(DeviceArrays A, char* __restrict__ send_lo, char* __restrict__ send_hi) {
    int64_t m = (int64_t)A.halo * A.Nx;
    for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < m; q += (int64_t)gridDim.x * blockDim.x) {
        int64_t lo = (int64_t)A.halo * A.Nx + q;                 /* first owned rows */
        int64_t hi = (int64_t)A.ny * A.Nx + q;                   /* last owned rows (ext index = ny+halo-halo) */
#pragma unroll
        for (int k = 0; k < 5; k++) {
            ((double*)send_lo)[k * m + q] = A.rec[k][lo];
            ((double*)send_hi)[k * m + q] = A.rec[k][hi];
        }
        ((int32_t*)(send_lo + 5 * m * 8))[q] = A.cell[lo];
        ((int32_t*)(send_hi + 5 * m * 8))[q] = A.cell[hi];
    }
}
__global__ void k_halo_unpack(DeviceArrays A, const char* __restrict__ recv_lo, const char* __restrict__ recv_hi,
                              DeviceCounters* dc) {
    int64_t m = (int64_t)A.halo * A.Nx;
    int32_t reach = 0;
    for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < m; q += (int64_t)gridDim.x * blockDim.x) {
        int64_t lo = q;                                           /* lower halo rows */
        int64_t hi = (int64_t)(A.ny + A.halo) * A.Nx + q;         /* upper halo rows */
#pragma unroll
        for (int k = 0; k < 5; k++) {
            A.rec[k][lo] = ((const double*)recv_lo)[k * m + q];
            A.rec[k][hi] = ((const double*)recv_hi)[k * m + q];
        }
        int32_t clo = ((const int32_t*)(recv_lo + 5 * m * 8))[q], chi = ((const int32_t*)(recv_hi + 5 * m * 8))[q];
        A.cell[lo] = clo;
        A.cell[hi] = chi;
        reach = max(reach, max(cell_reach(clo), cell_reach(chi)));
    }
    reach = warp_max(reach);
    if ((threadIdx.x & 31) == 0 && reach) atomicMax(&dc->reach_halo, reach);
}
__global__ void k_fill_i32(int32_t* p, int64_t n, int32_t v) {
    for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < n; q += (int64_t)gridDim.x * blockDim.x) p[q] = v;
}

/* ---- roofline denominators measured in place ------------------------------------ */
/* 8 independent DFMA chains per thread: the FP64 pipe's issue-rate ceiling */
__global__ void __launch_bounds__(256) k_fp64_peak(double* out, int iters) {
    double a0 = threadIdx.x * 1e-3, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double m = 0.999999, c = 1e-7;
    for (int k = 0; k < iters; k++) {
#pragma unroll
        for (int r = 0; r < 8; r++) {
            a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
            a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
        }
    }
    double s = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
    if (s == 123.456) out[0] = s; /* keep the chains alive */
}
__global__ void __launch_bounds__(256) k_copy_f64(double2* __restrict__ dst, const double2* __restrict__ src, int64_t n2) {
    for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < n2; q += (int64_t)gridDim.x * blockDim.x) dst[q] = src[q];
}

/* ---- self-test of the fast-path arithmetic (pmath.h) against the IEEE operators ---- */
__device__ __forceinline__ uint64_t st_next(uint64_t& s) {
    s ^= s << 13; s ^= s >> 7; s ^= s << 17;
    return s;
}
/* operand classes: 0 raw bit patterns (every exponent, NaN/Inf/denormals included),
   1 moderate magnitudes 2^[-40,40] (the physics range), 2 one operand from a table of
   special values */
__device__ double st_operand(uint64_t& s, int cls) {
    uint64_t r = st_next(s);
    if (cls == 0) return __longlong_as_double((long long)r);
    if (cls == 1) {
        uint64_t mant = r & 0x000fffffffffffffull;
        uint64_t e = 1023 - 40 + ((r >> 52) % 81);
        uint64_t sign = (r >> 63) << 63;
        return __longlong_as_double((long long)(sign | (e << 52) | mant));
    }
    const double tab[12] = {0.0, -0.0, 1.0, -1.0, 4.9406564584124654e-324, 2.2250738585072014e-308,
                            1.7976931348623157e308, __longlong_as_double(0x7ff0000000000000ll),
                            __longlong_as_double(0x7ff8000000000000ll), 1.9999999999999998, 0.85, 1e-300};
    return tab[(r >> 32) % 12];
}
__global__ void k_selftest_math(uint64_t seed, int iters, unsigned long long* out /* [6] */) {
    uint64_t s = seed ^ (0x9E3779B97F4A7C15ull * (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x + 1));
    unsigned long long n_div = 0, bad_div = 0, mis_div = 0, n_sqrt = 0, bad_sqrt = 0, mis_sqrt = 0;
    for (int it = 0; it < iters; it++) {
        int cls = it % 3;
        double a = st_operand(s, cls == 2 ? (it & 1 ? 2 : 1) : cls);
        double b = st_operand(s, cls == 2 ? (it & 1 ? 1 : 2) : cls);
        unsigned bad = 0;
        double q = (it & 4) ? pm_divz_fast(a, b, &bad) : pm_div_fast(a, b, &bad);
        double qi = a / b;
        n_div++;
        if (bad) bad_div++;
        else if (__double_as_longlong(q) != __double_as_longlong(qi) && !(q != q && qi != qi)) mis_div++;
        double x = fabs(a);
        if (cls == 2 && (it & 2)) x = a;
        bad = 0;
        double r = (it & 4) ? pm_sqrtz_fast(x, &bad) : pm_sqrt_fast(x, &bad);
        double ri = sqrt(x);
        n_sqrt++;
        if (bad) bad_sqrt++;
        else if (__double_as_longlong(r) != __double_as_longlong(ri) && !(r != r && ri != ri)) mis_sqrt++;
    }
    atomicAdd(&out[0], n_div); atomicAdd(&out[1], bad_div); atomicAdd(&out[2], mis_div);
    atomicAdd(&out[3], n_sqrt); atomicAdd(&out[4], bad_sqrt); atomicAdd(&out[5], mis_sqrt);
}

/* ---- launchers ------------------------------------------------------------------ */
static int grid_for(int64_t n, int threads, int sms, int blocks_per_sm) {
    int64_t need = (n + threads - 1) / threads;
    int64_t cap = (int64_t)sms * blocks_per_sm;
    if (need < 1) need = 1;
    return (int)(need < cap ? need : cap);
}

void launch_seed(const DeviceArrays& A, const picles_params_t& P, const double* u0, const double* v0, int sms,
                 cudaStream_t st) {
    int64_t n = (int64_t)A.Nx * A.ny;
    k_seed<<<grid_for(n, 256, sms, 8), 256, 0, st>>>(A, P, u0, v0);
}

void launch_advance(const DeviceArrays& A, const picles_params_t& P, double DT, DeviceCounters* dc, int sms,
                    cudaStream_t st) {
    int64_t n = (int64_t)A.Nx * A.ny;
    int g = grid_for(n, ADV_THREADS, sms, ADV_MIN_BLOCKS);
    bool pn = (A.M[0] != nullptr);
    if (pn) k_advance<true><<<g, ADV_THREADS, 0, st>>>(A, P, DT, dc);
    else k_advance<false><<<g, ADV_THREADS, 0, st>>>(A, P, DT, dc);
}

void launch_project(const DeviceArrays& A, int n_classes, int accumulate, const DeviceCounters* dc, int sms, cudaStream_t st) {
    int gx = (A.Nx + PRJ_THREADS - 1) / PRJ_THREADS;
    int gy = A.ny < 65535 ? A.ny : 65535;
    (void)sms;
    k_project<<<dim3(gx, gy), PRJ_THREADS, 0, st>>>(A, n_classes, accumulate, dc);
}

void launch_remesh(const DeviceArrays& A, const picles_params_t& P, double DT, DeviceCounters* dc, int sms,
                   cudaStream_t st) {
    int64_t n = (int64_t)A.Nx * A.ny;
    k_remesh<<<grid_for(n, RMS_THREADS, sms, 8), RMS_THREADS, 0, st>>>(A, P, DT, dc);
}

void launch_energy(const double* e, int64_t n, double* partial, int nblocks, cudaStream_t st) {
    k_energy<<<nblocks, 256, 0, st>>>(e, n, partial);
}

void launch_halo_pack(const DeviceArrays& A, char* lo, char* hi, int sms, cudaStream_t st) {
    int64_t m = (int64_t)A.halo * A.Nx;
    if (m > 0) k_halo_pack<<<grid_for(m, 256, sms, 4), 256, 0, st>>>(A, lo, hi);
}
void launch_halo_unpack(const DeviceArrays& A, const char* lo, const char* hi, DeviceCounters* dc, int sms, cudaStream_t st) {
    int64_t m = (int64_t)A.halo * A.Nx;
    if (m > 0) k_halo_unpack<<<grid_for(m, 256, sms, 4), 256, 0, st>>>(A, lo, hi, dc);
}
void launch_fill_i32(int32_t* p, int64_t n, int32_t v, int sms, cudaStream_t st) {
    if (n > 0) k_fill_i32<<<grid_for(n, 256, sms, 4), 256, 0, st>>>(p, n, v);
}

void launch_selftest_math(uint64_t seed, int iters, unsigned long long* out, int sms, cudaStream_t st) {
    k_selftest_math<<<sms * 8, 256, 0, st>>>(seed, iters, out);
}
void launch_fp64_peak(double* out, int iters, int sms, cudaStream_t st, int64_t* fmas) {
    int blocks = sms * 8;
    k_fp64_peak<<<blocks, 256, 0, st>>>(out, iters);
    *fmas = (int64_t)blocks * 256 * (int64_t)iters * 64;
}
void launch_copy_f64(double* dst, const double* src, int64_t n, int sms, cudaStream_t st) {
    k_copy_f64<<<sms * 16, 256, 0, st>>>((double2*)dst, (const double2*)src, n / 2);
}

} /* namespace picles */
