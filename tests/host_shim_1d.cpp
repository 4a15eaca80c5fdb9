/* Host build of the 1-D DEVICE arithmetic (picles_b200/csrc/physics1d.h) driven as the kernels of picles1d.cu drive
   it — one "thread" per particle for advance, one per node for gather + remesh — so the CPU suite checks the header
   the kernels are compiled from against the independent oracle (oracle/picles_oracle_1d.c), bit for bit, before any
   GPU time is spent.  Test infrastructure; never part of the product. */
#include <cstring>
#include <vector>

#include "../picles_b200/csrc/physics1d.h"

using namespace picles1d;

struct Shim1 {
    int Nx;
    picles_params_t P;
    std::vector<double> xn, z0, z1, z2, t, dt, qold, w0, w1, S, re, rm, rwf, rwc;
    std::vector<int64_t> iter, rifl;
    std::vector<uint8_t> flags;
    std::vector<int32_t> status;
    Arrays A;
    picles_counters_t last;
};

static void bind(Shim1* s) {
    Arrays& A = s->A;
    A.xn = s->xn.data(); A.z0 = s->z0.data(); A.z1 = s->z1.data(); A.z2 = s->z2.data();
    A.t = s->t.data(); A.dt = s->dt.data(); A.qold = s->qold.data(); A.iter = s->iter.data();
    A.flags = s->flags.data(); A.status = s->status.data(); A.w0 = s->w0.data(); A.w1 = s->w1.data();
    A.S = s->S.data(); A.r_e = s->re.data(); A.r_m = s->rm.data(); A.r_wf = s->rwf.data(); A.r_wc = s->rwc.data();
    A.r_ifl = s->rifl.data();
}
static void load(const Arrays& A, int i, Particle1& p) {
    p.u[0] = A.z0[i]; p.u[1] = A.z1[i]; p.u[2] = A.z2[i]; p.t = A.t[i]; p.dt = A.dt[i]; p.qold = A.qold[i];
    p.iter = A.iter[i]; p.flags = A.flags[i]; p.status = A.status[i];
}
static void store(const Arrays& A, int i, const Particle1& p) {
    A.z0[i] = p.u[0]; A.z1[i] = p.u[1]; A.z2[i] = p.u[2]; A.t[i] = p.t; A.dt[i] = p.dt; A.qold[i] = p.qold;
    A.iter[i] = p.iter; A.flags[i] = p.flags; A.status[i] = p.status;
}

extern "C" {
void* shim1_create(int Nx, double xmin, double dx, const double* x_nodes, const picles_params_t* P) {
    Shim1* s = new Shim1();
    s->Nx = Nx; s->P = *P;
    if (s->P.solver == PICLES_SOLVER_AUTOTSIT5) s->P.solver = PICLES_SOLVER_TSIT5;
    s->xn.assign(x_nodes, x_nodes + Nx);
    for (auto* v : {&s->z0, &s->z1, &s->z2, &s->t, &s->dt, &s->qold, &s->w0, &s->w1, &s->re, &s->rm, &s->rwf, &s->rwc}) v->assign(Nx, 0.0);
    s->S.assign(3 * (size_t)Nx, 0.0);
    s->iter.assign(Nx, 0); s->rifl.assign(Nx, P1_NO_DEPOSIT); s->flags.assign(Nx, 0); s->status.assign(Nx, 0);
    s->A.Nx = Nx; s->A.xmin = xmin; s->A.dx = dx;
    bind(s);
    memset(&s->last, 0, sizeof s->last);
    return s;
}
void shim1_destroy(void* h) { delete (Shim1*)h; }
void shim1_seed(void* h, const double* u0) {
    Shim1* s = (Shim1*)h;
    for (int i = 0; i < s->Nx; i++) {
        Particle1 p;
        double st[3];
        p1_seed(s->A, s->P, i, u0[i], p, st);
        store(s->A, i, p);
        s->S[i] = st[0]; s->S[i + s->Nx] = st[1]; s->S[i + 2 * (size_t)s->Nx] = st[2];
        s->rifl[i] = P1_NO_DEPOSIT;
    }
}
void shim1_step(void* h, double t, double DT, const double* u_t, const double* u_t1) {
    (void)t;
    Shim1* s = (Shim1*)h;
    s->w0.assign(u_t, u_t + s->Nx); s->w1.assign(u_t1, u_t1 + s->Nx);
    bind(s);
    Tally1 c;
    memset(&c, 0, sizeof c);
    for (int i = 0; i < s->Nx; i++) {
        Particle1 p;
        load(s->A, i, p);
        p1_advance(s->A, s->P, i, DT, p, c, s->re[i], s->rm[i], s->rwf[i], s->rwc[i], s->rifl[i]);
        store(s->A, i, p);
    }
    const int R = c.reach;
    std::vector<double> Snew(3 * (size_t)s->Nx);
    for (int i = 0; i < s->Nx; i++) { /* gather reads records only; remesh writes particle arrays only */
        double g[3];
        p1_gather_node(s->A, s->P.periodic_boundary, R, (int64_t)i + 1, g);
        Snew[i] = g[0]; Snew[i + s->Nx] = g[1]; Snew[i + 2 * (size_t)s->Nx] = g[2];
        Particle1 p;
        load(s->A, i, p);
        p1_remesh(s->A, s->P, i, DT, g, s->w0[i], p, c);
        store(s->A, i, p);
    }
    s->S = Snew;
    bind(s);
    picles_counters_t& L = s->last;
    memset(&L, 0, sizeof L);
    L.n_active = s->Nx; L.n_integrated = c.integrated; L.n_substeps = c.substeps; L.n_rejects = c.rejects; L.n_rhs = c.rhs;
    L.n_reseed_advance = c.reseed; L.n_fixups = c.fixups; L.n_failed = c.failed; L.n_deposited = c.deposited;
    L.n_remesh_A = c.A; L.n_remesh_B = c.B; L.n_remesh_D = c.D; L.reach = c.reach; L.max_attempts = c.max_attempts;
}
void shim1_get_state(void* h, double* S) { Shim1* s = (Shim1*)h; memcpy(S, s->S.data(), sizeof(double) * 3 * s->Nx); }
void shim1_get_particles(void* h, double* z, double* t, double* dt, uint8_t* flags, int32_t* status) {
    Shim1* s = (Shim1*)h;
    const int N = s->Nx;
    memcpy(z, s->z0.data(), 8 * N); memcpy(z + N, s->z1.data(), 8 * N); memcpy(z + 2 * (size_t)N, s->z2.data(), 8 * N);
    memcpy(t, s->t.data(), 8 * N); memcpy(dt, s->dt.data(), 8 * N); memcpy(flags, s->flags.data(), N);
    memcpy(status, s->status.data(), 4 * N);
}
void shim1_get_counters(void* h, picles_counters_t* c) { *c = ((Shim1*)h)->last; }
/* known-answer hooks: the device header's merge rule and right-hand side by themselves */
void shim1_merge(double* g, const double* c) { p1_merge(g, c); }
void shim1_rhs(const picles_params_t* P, const double* z, double u, double* dz) { p1_rhs(*P, z, u, dz); }
}
