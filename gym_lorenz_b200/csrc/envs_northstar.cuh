// North-star env kinds (BASELINE.json north_star; SURVEY.md section 0 "D1-D3" and 8d):
// classical RK4 with S substeps per control interval, zero-order-hold control added to the
// derivative, per-env parameters held in registers next to the state.  These kinds have no
// counterpart class in the reference (its Lorenz envs are Euler x 1, dynamic.py:70-75);
// they keep dynamic.py's observation / reward / action contract and are validated against
// adaptive integration (scipy DOP853, rtol 1e-13) per control interval at rtol 1e-9.
//
// Compiled into tu_northstar.cu with FMA contraction ON: per RHS evaluation 7 FMA-pipe
// instructions, per RK4 substep 49 (28 RHS + 9 stage + 12 combine) = 87 algorithmic flop.
#pragma once
#include "kernels_common.cuh"

namespace cl {

template <typename R>
struct LorenzPar { R sigma, rho, beta; };

// f(s) + u : dx = sigma (y - x) + u1 ; dy = x (rho - z) - y + u2 ; dz = x y - beta z + u3
// (textbook form: used for the observation's derivative half, once per control interval)
template <typename R>
__device__ __forceinline__ void lorenz_rhs_u(const LorenzPar<R>& q, R x, R y, R z, R u1, R u2, R u3,
                                             R& dx, R& dy, R& dz) {
  dx = fma(q.sigma, y - x, u1);
  dy = fma(x, q.rho - z, u2 - y);
  dz = fma(x, y, fma(-q.beta, z, u3));
}

// The integrator works on the shifted coordinate zs = z - rho for the whole control interval:
//   dx = sigma (y - x) + u1 ;  dy = -x zs + (u2 - y) ;  d(zs) = x y - beta zs + c3 ,  c3 = u3 - beta rho
// (u is held over the interval, so c3 is an interval constant).  Same vector field, one FP64
// instruction less per evaluation: 6 (2 DADD + 4 DFMA) instead of 7, i.e. 45 instead of 49 per RK4
// substep (24 RHS + 9 stage + 12 combine).  The algorithmic count stays 87 flop.
template <typename R>
__device__ __forceinline__ void lorenz_rhs_s(const R sigma, const R beta, R x, R y, R zs, R u1, R u2, R c3,
                                             R& dx, R& dy, R& dz) {
  dx = fma(sigma, y - x, u1);
  dy = fma(-x, zs, u2 - y);
  dz = fma(x, y, fma(-beta, zs, c3));
}

// Step sizes (h, h/2, h/3, h/6) and -- on the warp-uniform fast path -- the parameters arrive
// as kernel-parameter (constant-bank) operands: on sm_100 a DFMA with three distinct register
// sources issues every 3 cycles per scheduler, one with <= 2 register sources every 2
// (tools/dfma_probe.cu, DESIGN.md "FP64 cost model"), so only the 8 inherently three-register
// FMAs per substep (dy, dz of each stage) pay the slow rate.
#define CL_LORENZ_RK4_SUBSTEP()                                                                     \
  {                                                                                                 \
    R k1x, k1y, k1z, kx, ky, kz, ax, ay, az;                                                        \
    lorenz_rhs_s(sigma, beta, x, y, zs, u1, u2, c3, k1x, k1y, k1z);                                 \
    ax = fma(h6, k1x, x); ay = fma(h6, k1y, y); az = fma(h6, k1z, zs);                              \
    lorenz_rhs_s(sigma, beta, fma(hh, k1x, x), fma(hh, k1y, y), fma(hh, k1z, zs), u1, u2, c3, kx, ky, kz);   \
    ax = fma(h3, kx, ax); ay = fma(h3, ky, ay); az = fma(h3, kz, az);                               \
    lorenz_rhs_s(sigma, beta, fma(hh, kx, x), fma(hh, ky, y), fma(hh, kz, zs), u1, u2, c3, k1x, k1y, k1z);   \
    ax = fma(h3, k1x, ax); ay = fma(h3, k1y, ay); az = fma(h3, k1z, az);                            \
    lorenz_rhs_s(sigma, beta, fma(h, k1x, x), fma(h, k1y, y), fma(h, k1z, zs), u1, u2, c3, kx, ky, kz);      \
    x = fma(h6, kx, ax); y = fma(h6, ky, ay); zs = fma(h6, kz, az);                                 \
  }

// `br` = beta * rho rounded once (host-side product on the constant path, __dmul_rn on the
// per-env path: the same IEEE operation, so both paths give identical bits for equal parameters).
template <typename R>
__device__ __forceinline__ void lorenz_rk4(const R sigma, const R rho, const R beta, const R br, R& x, R& y, R& z,
                                           R u1, R u2, R u3, const R h, const R hh, const R h3, const R h6,
                                           int substeps) {
  R zs = z - rho;
  const R c3 = u3 - br;
#pragma unroll 2
  for (int k = 0; k < substeps; ++k) CL_LORENZ_RK4_SUBSTEP()
  z = zs + rho;
}

// Fully unrolled variant for the common substep counts: without an inner loop ptxas does not
// force the warp to drain its outstanding loads (the next interval's action prefetch) at a loop
// head, so that latency hides behind the S x 45 FMA-pipe instructions of the interval.
template <typename R, int S>
__device__ __forceinline__ void lorenz_rk4_fixed(const R sigma, const R rho, const R beta, const R br, R& x, R& y,
                                                 R& z, R u1, R u2, R u3, const R h, const R hh, const R h3,
                                                 const R h6) {
  R zs = z - rho;
  const R c3 = u3 - br;
#pragma unroll
  for (int k = 0; k < S; ++k) CL_LORENZ_RK4_SUBSTEP()
  z = zs + rho;
}

template <typename R>
__device__ __forceinline__ void lorenz_rk4_any(const R sigma, const R rho, const R beta, const R br, R& x, R& y,
                                               R& z, R u1, R u2, R u3, const R h, const R hh, const R h3,
                                               const R h6, int S) {
  if (S == 16) {  // the benchmark configuration: a plain compare-and-branch instead of the jump table
    lorenz_rk4_fixed<R, 16>(sigma, rho, beta, br, x, y, z, u1, u2, u3, h, hh, h3, h6);
    return;
  }
  switch (S) {  // warp-uniform
    case 8: lorenz_rk4_fixed<R, 8>(sigma, rho, beta, br, x, y, z, u1, u2, u3, h, hh, h3, h6); break;
    case 4: lorenz_rk4_fixed<R, 4>(sigma, rho, beta, br, x, y, z, u1, u2, u3, h, hh, h3, h6); break;
    case 1: lorenz_rk4_fixed<R, 1>(sigma, rho, beta, br, x, y, z, u1, u2, u3, h, hh, h3, h6); break;
    default: lorenz_rk4<R>(sigma, rho, beta, br, x, y, z, u1, u2, u3, h, hh, h3, h6, S); break;
  }
}

// CL_ENV_LORENZ_RK4 / _F32.  planes: x y z sigma rho beta
template <typename R>
struct EnvLorenzRK4 {
  typedef R real;
  enum { NSTATE = 6, NINT = 0, OBS = 6, ACT = 3, NOISE = 0 };
  struct S { R x, y, z; LorenzPar<R> q; bool uni; };
  __device__ static void load(S& s, const KParams& p, int64_t i) {
    s.x = ldp<R>(p, 0, i); s.y = ldp<R>(p, 1, i); s.z = ldp<R>(p, 2, i);
    s.q.sigma = ldp<R>(p, 3, i); s.q.rho = ldp<R>(p, 4, i); s.q.beta = ldp<R>(p, 5, i);
  }
  // all 32 lanes vote: if every live env of the warp carries the nominal parameters, the warp
  // integrates with them as constant-bank operands; otherwise with its per-env registers.
  __device__ static void prepare(S& s, const KParams& p, bool live) {
    const bool same = !live || (sizeof(R) == 8
        ? (s.q.sigma == (R)p.nom[0] && s.q.rho == (R)p.nom[1] && s.q.beta == (R)p.nom[2])
        : (s.q.sigma == (R)p.nomf[0] && s.q.rho == (R)p.nomf[1] && s.q.beta == (R)p.nomf[2]));
    s.uni = __all_sync(0xffffffffu, same);
  }
  __device__ static void store(const S& s, const KParams& p, int64_t i) {
    stp<R>(p, 0, i, s.x); stp<R>(p, 1, i, s.y); stp<R>(p, 2, i, s.z);
    stp<R>(p, 3, i, s.q.sigma); stp<R>(p, 4, i, s.q.rho); stp<R>(p, 5, i, s.q.beta);
  }
  // shared-memory planes [NSTATE][32] of one env-warp (k_rollout_sm)
  __device__ static void load_sm(S& s, const R* b, unsigned lane) {
    s.x = b[lane]; s.y = b[32 + lane]; s.z = b[64 + lane];
    s.q.sigma = b[96 + lane]; s.q.rho = b[128 + lane]; s.q.beta = b[160 + lane];
  }
  __device__ static void store_sm(const S& s, R* b, unsigned lane) {
    b[lane] = s.x; b[32 + lane] = s.y; b[64 + lane] = s.z;   // the parameter planes never change
  }
  __device__ static bool uses_noise(const KParams&) { return false; }
  __device__ static bool finite(const S& s) { return isfinite(s.x + s.y + s.z); }
  __device__ static bool time_limit(const KParams& p, int32_t n) {
    return p.max_steps > 0 && n >= p.max_steps;
  }
  // nominal (10, 28, 8/3) as dynamic.py:31-33, optionally jittered per env (new)
  __device__ static void init_persistent(S& s, const KParams& p, const Stream& rng) {
    double j[3] = {1.0, 1.0, 1.0};
    if (p.param_jitter > 0.0) draw_uniform<3>(rng, TAG_PARAM, 1.0 - p.param_jitter, 1.0 + p.param_jitter, j);
    s.q.sigma = (R)(10.0 * j[0]); s.q.rho = (R)(28.0 * j[1]); s.q.beta = (R)((8.0 / 3.0) * j[2]);
    s.x = s.y = s.z = R(0);
  }
  __device__ static void observe(const S& s, R* obs) {
    obs[0] = s.x; obs[1] = s.y; obs[2] = s.z;
    lorenz_rhs_u(s.q, s.x, s.y, s.z, R(0), R(0), R(0), obs[3], obs[4], obs[5]);
  }
  __device__ static void reset(S& s, const KParams&, const Stream& rng, R* obs) {
    double u[3];
    draw_uniform<3>(rng, TAG_RESET, -30.0, 30.0, u);  // dynamic.py:37
    s.x = (R)u[0]; s.y = (R)u[1]; s.z = (R)u[2];
    observe(s, obs);
  }
  // SPEC 1: every env of the warp carries the nominal parameters AND S == 16 (the benchmark
  // configuration): constant-bank parameters, 16 unrolled substeps, no per-interval dispatch.
  __device__ static int spec(const S& s, const KParams& p) {
    return (s.uni && p.substeps == 16) ? 1 : 0;
  }
  template <int SPEC>
  __device__ static void step_spec(S& s, const KParams& p, const float* a, const double*, R* obs, R& rew,
                                   bool& term) {
    static_assert(SPEC == 1, "unknown specialisation");
    const float lim = p.act_limit_f;
    const R g = sizeof(R) == 8 ? (R)p.act_gain : (R)p.act_gain_f;
    const R u1 = mul_keep((R)clipf(a[0], -lim, lim), g);
    const R u2 = mul_keep((R)clipf(a[1], -lim, lim), g);
    const R u3 = mul_keep((R)clipf(a[2], -lim, lim), g);
    if (sizeof(R) == 8) {
      lorenz_rk4_fixed<R, 16>((R)p.nom[0], (R)p.nom[1], (R)p.nom[2], (R)p.nom_br, s.x, s.y, s.z, u1, u2, u3,
                              (R)p.h, (R)p.hh, (R)p.h3, (R)p.h6);
    } else {
      lorenz_rk4_fixed<R, 16>((R)p.nomf[0], (R)p.nomf[1], (R)p.nomf[2], (R)p.nom_brf, s.x, s.y, s.z, u1, u2, u3,
                              (R)p.hf, (R)p.hhf, (R)p.h3f, (R)p.h6f);
    }
    // with the env's own (register) parameters: f(s) + 0 with CONSTANT parameters would need two
    // non-register operands in one DFMA, which forces the constants into registers -- and ptxas
    // then keeps using those registers inside the integrator (3-register DFMAs again)
    observe(s, obs);
    const R e = fabs(s.x) + fabs(s.y) + fabs(s.z);
    rew = -e;
    term = !(e <= R(1e6));
  }
  __device__ static void step(S& s, const KParams& p, const float* a, const double*, R* obs, R& rew,
                              bool& term) {
    const float lim = p.act_limit_f;
    const R g = sizeof(R) == 8 ? (R)p.act_gain : (R)p.act_gain_f;
    const R u1 = mul_keep((R)clipf(a[0], -lim, lim), g);
    const R u2 = mul_keep((R)clipf(a[1], -lim, lim), g);
    const R u3 = mul_keep((R)clipf(a[2], -lim, lim), g);
    if (s.uni) {
      if (sizeof(R) == 8)
        lorenz_rk4_any<R>((R)p.nom[0], (R)p.nom[1], (R)p.nom[2], (R)p.nom_br, s.x, s.y, s.z, u1, u2, u3,
                          (R)p.h, (R)p.hh, (R)p.h3, (R)p.h6, p.substeps);
      else
        lorenz_rk4_any<R>((R)p.nomf[0], (R)p.nomf[1], (R)p.nomf[2], (R)p.nom_brf, s.x, s.y, s.z, u1, u2, u3,
                          (R)p.hf, (R)p.hhf, (R)p.h3f, (R)p.h6f, p.substeps);
    } else {
      const R br = mul_rn(s.q.beta, s.q.rho);
      if (sizeof(R) == 8)
        lorenz_rk4_any<R>(s.q.sigma, s.q.rho, s.q.beta, br, s.x, s.y, s.z, u1, u2, u3,
                          (R)p.h, (R)p.hh, (R)p.h3, (R)p.h6, p.substeps);
      else
        lorenz_rk4_any<R>(s.q.sigma, s.q.rho, s.q.beta, br, s.x, s.y, s.z, u1, u2, u3,
                          (R)p.hf, (R)p.hhf, (R)p.h3f, (R)p.h6f, p.substeps);
    }
    observe(s, obs);
    const R e = fabs(s.x) + fabs(s.y) + fabs(s.z);
    rew = -e;                      // dynamic.py:84
    term = !(e <= R(1e6));         // blow-up / NaN guard (lorenz_env_transient.py:369 `reward < -1e6`)
  }
};

// ------------------------------------------------------------------------------------
// CL_ENV_PMSM_RK4 (f64).  Master/slave chaotic PMSM pair (lorenz_env_try_pmsm.py:51-58
// dynamics), RK4 x S, control clip(a, +-1) * gain on slave dx1, dx2, per-env sigma / gamma.
// planes: a0 a1 a2  b0 b1 b2  sigma gamma
struct PMSMPar { double sigma, gamma; };

__device__ __forceinline__ void pmsm_rhs_u(const PMSMPar& q, const double* x, double u1, double u2,
                                           double* d) {
  d[0] = fma(x[1], x[2], u1 - x[0]);
  d[1] = fma(q.gamma - x[0], x[2], u2 - x[1]);
  d[2] = q.sigma * (x[1] - x[2]);
}

// The integrator works on the shifted coordinate x0s = x0 - gamma for the whole control interval
// (same idea as lorenz_rhs_s): d0 = x1 x2 + (c1 - x0s), c1 = u1 - gamma ; d1 = -x0s x2 + (u2 - x1) ;
// d2 = sigma (x1 - x2): 6 FP64 instructions per evaluation instead of 7, 45 instead of 49 per substep.
__device__ __forceinline__ void pmsm_rhs_s(const double sigma, const double* x, double c1, double u2, double* d) {
  d[0] = fma(x[1], x[2], c1 - x[0]);
  d[1] = fma(-x[0], x[2], u2 - x[1]);
  d[2] = sigma * (x[1] - x[2]);
}

#define CL_PMSM_RK4_SUBSTEP(x, c1, u2)                                                                    \
  {                                                                                                        \
    double k1[3], k2[3], w[3], acc[3];                                                                     \
    pmsm_rhs_s(sigma, x, c1, u2, k1);                                                                      \
    _Pragma("unroll") for (int c = 0; c < 3; ++c) { acc[c] = fma(h6, k1[c], x[c]); w[c] = fma(hh, k1[c], x[c]); }   \
    pmsm_rhs_s(sigma, w, c1, u2, k2);                                                                      \
    _Pragma("unroll") for (int c = 0; c < 3; ++c) { acc[c] = fma(h3, k2[c], acc[c]); w[c] = fma(hh, k2[c], x[c]); } \
    pmsm_rhs_s(sigma, w, c1, u2, k1);                                                                      \
    _Pragma("unroll") for (int c = 0; c < 3; ++c) { acc[c] = fma(h3, k1[c], acc[c]); w[c] = fma(h, k1[c], x[c]); }  \
    pmsm_rhs_s(sigma, w, c1, u2, k2);                                                                      \
    _Pragma("unroll") for (int c = 0; c < 3; ++c) x[c] = fma(h6, k2[c], acc[c]);                           \
  }

// Master (free-running, a) and slave (controlled, b) advance in the SAME substep loop: two independent
// dependency chains per thread, so ~3 worker warps per scheduler see twice the instruction-level
// parallelism.  S > 0: compile-time substep count, fully unrolled (the control interval becomes one
// straight line of code); S == 0: run-time count.  Either way the operations per system, and hence the
// bits, are the same.
template <int S>
__device__ __forceinline__ void pmsm_rk4_pair(const double sigma, const double gamma, double* a, double* b,
                                              double u1, double u2, const double h, const double hh, const double h3,
                                              const double h6, int substeps) {
  double xa[3] = {a[0] - gamma, a[1], a[2]};
  double xb[3] = {b[0] - gamma, b[1], b[2]};
  const double c1a = 0.0 - gamma, c1b = u1 - gamma;
  if (S > 0) {
#pragma unroll
    for (int k = 0; k < S; ++k) {
      CL_PMSM_RK4_SUBSTEP(xa, c1a, 0.0)
      CL_PMSM_RK4_SUBSTEP(xb, c1b, u2)
    }
  } else {
#pragma unroll 2
    for (int k = 0; k < substeps; ++k) {
      CL_PMSM_RK4_SUBSTEP(xa, c1a, 0.0)
      CL_PMSM_RK4_SUBSTEP(xb, c1b, u2)
    }
  }
  a[0] = xa[0] + gamma; a[1] = xa[1]; a[2] = xa[2];
  b[0] = xb[0] + gamma; b[1] = xb[1]; b[2] = xb[2];
}

__device__ __forceinline__ void pmsm_rk4_pair_any(const double sigma, const double gamma, double* a, double* b,
                                                  double u1, double u2, const double h, const double hh,
                                                  const double h3, const double h6, int S) {
  switch (S) {  // warp-uniform
    case 4: pmsm_rk4_pair<4>(sigma, gamma, a, b, u1, u2, h, hh, h3, h6, S); break;   // the env's default
    case 8: pmsm_rk4_pair<8>(sigma, gamma, a, b, u1, u2, h, hh, h3, h6, S); break;
    case 2: pmsm_rk4_pair<2>(sigma, gamma, a, b, u1, u2, h, hh, h3, h6, S); break;
    case 1: pmsm_rk4_pair<1>(sigma, gamma, a, b, u1, u2, h, hh, h3, h6, S); break;
    default: pmsm_rk4_pair<0>(sigma, gamma, a, b, u1, u2, h, hh, h3, h6, S); break;
  }
}

struct EnvPMSMRK4 {
  typedef double real;
  enum { NSTATE = 8, NINT = 0, OBS = 6, ACT = 2, NOISE = 0 };
  struct S { double a[3], b[3]; PMSMPar q; bool uni; };
  __device__ static void load(S& s, const KParams& p, int64_t i) {
#pragma unroll
    for (int c = 0; c < 3; ++c) { s.a[c] = ldp<double>(p, c, i); s.b[c] = ldp<double>(p, 3 + c, i); }
    s.q.sigma = ldp<double>(p, 6, i); s.q.gamma = ldp<double>(p, 7, i);
  }
  __device__ static void prepare(S& s, const KParams& p, bool live) {
    const bool same = !live || (s.q.sigma == p.nom[0] && s.q.gamma == p.nom[1]);
    s.uni = __all_sync(0xffffffffu, same);
  }
  __device__ static void store(const S& s, const KParams& p, int64_t i) {
#pragma unroll
    for (int c = 0; c < 3; ++c) { stp<double>(p, c, i, s.a[c]); stp<double>(p, 3 + c, i, s.b[c]); }
    stp<double>(p, 6, i, s.q.sigma); stp<double>(p, 7, i, s.q.gamma);
  }
  __device__ static void load_sm(S& s, const double* b, unsigned lane) {
#pragma unroll
    for (int c = 0; c < 3; ++c) { s.a[c] = b[c * 32 + lane]; s.b[c] = b[(3 + c) * 32 + lane]; }
    s.q.sigma = b[6 * 32 + lane]; s.q.gamma = b[7 * 32 + lane];
  }
  __device__ static void store_sm(const S& s, double* b, unsigned lane) {
#pragma unroll
    for (int c = 0; c < 3; ++c) { b[c * 32 + lane] = s.a[c]; b[(3 + c) * 32 + lane] = s.b[c]; }
  }
  __device__ static bool uses_noise(const KParams&) { return false; }
  __device__ static bool finite(const S& s) {
    return isfinite(s.a[0] + s.a[1] + s.a[2] + s.b[0] + s.b[1] + s.b[2]);
  }
  __device__ static bool time_limit(const KParams& p, int32_t n) {
    return p.max_steps > 0 && n >= p.max_steps;
  }
  __device__ static void init_persistent(S& s, const KParams& p, const Stream& rng) {
    double j[2] = {1.0, 1.0};
    if (p.param_jitter > 0.0) draw_uniform<2>(rng, TAG_PARAM, 1.0 - p.param_jitter, 1.0 + p.param_jitter, j);
    s.q.sigma = 5.46 * j[0];  // lorenz_env_try_pmsm.py:12-13
    s.q.gamma = 20.0 * j[1];
#pragma unroll
    for (int c = 0; c < 3; ++c) { s.a[c] = 0.0; s.b[c] = 0.0; }
  }
  __device__ static void observe(const S& s, double* obs) {
    double da[3], db[3];
    pmsm_rhs_u(s.q, s.a, 0.0, 0.0, da);
    pmsm_rhs_u(s.q, s.b, 0.0, 0.0, db);
#pragma unroll
    for (int c = 0; c < 3; ++c) { obs[c] = s.a[c] - s.b[c]; obs[3 + c] = da[c] - db[c]; }
  }
  __device__ static void reset(S& s, const KParams&, const Stream& rng, double* obs) {
    double u[6];
    draw_uniform<6>(rng, TAG_RESET, -30.0, 30.0, u);  // lorenz_env_try_pmsm.py:64-65
#pragma unroll
    for (int c = 0; c < 3; ++c) { s.a[c] = u[c]; s.b[c] = u[3 + c]; }
    observe(s, obs);
  }
  // SPEC 1: nominal parameters in the whole warp -> constant-bank operands
  __device__ static int spec(const S& s, const KParams&) { return s.uni ? 1 : 0; }
  template <int SPEC>
  __device__ static void step_spec(S& s, const KParams& p, const float* a, const double*, double* obs,
                                   double& rew, bool& term) {
    static_assert(SPEC == 1, "unknown specialisation");
    const float lim = (float)p.act_limit;
    const double u1 = mul_keep((double)clipf(a[0], -lim, lim), p.act_gain);
    const double u2 = mul_keep((double)clipf(a[1], -lim, lim), p.act_gain);
    pmsm_rk4_pair_any(p.nom[0], p.nom[1], s.a, s.b, u1, u2, p.h, p.hh, p.h3, p.h6, p.substeps);
    finish(s, p, obs, rew, term);
  }
  __device__ static void finish(const S& s, const KParams& p, double* obs, double& rew, bool& term) {
    observe(s, obs);  // with the env's own (register) parameters, see EnvLorenzRK4::step_spec
    const double e0 = fabs(obs[0]), e1 = fabs(obs[1]), e2 = fabs(obs[2]);
    const double E = e0 + e1 + e2;
    rew = -E - (pow_pos(e0 + 1e-6, p.alpha) + pow_pos(e1 + 1e-6, p.alpha) + pow_pos(e2 + 1e-6, p.alpha));
    term = false;
    if (!(E <= 1000.0)) { rew = -1000.0; term = true; }  // lorenz_env_try_pmsm.py:174-176
  }
  __device__ static void step(S& s, const KParams& p, const float* a, const double*, double* obs,
                              double& rew, bool& term) {
    const float lim = (float)p.act_limit;
    const double u1 = mul_keep((double)clipf(a[0], -lim, lim), p.act_gain);
    const double u2 = mul_keep((double)clipf(a[1], -lim, lim), p.act_gain);
    if (s.uni) pmsm_rk4_pair_any(p.nom[0], p.nom[1], s.a, s.b, u1, u2, p.h, p.hh, p.h3, p.h6, p.substeps);
    else pmsm_rk4_pair_any(s.q.sigma, s.q.gamma, s.a, s.b, u1, u2, p.h, p.hh, p.h3, p.h6, p.substeps);
    finish(s, p, obs, rew, term);
  }
};

template <> struct PlainRollout<EnvLorenzRK4<double>> { enum { value = 1 }; };
template <> struct PlainRollout<EnvLorenzRK4<float>> { enum { value = 1 }; };   // issue-slot bound: same remedy
template <> struct PlainRollout<EnvPMSMRK4> { enum { value = 1 }; };
// not terminated <=> |x|+|y|+|z| <= 1e6, hence |x+y+z| <= 1e6: finite
template <> struct FiniteUnlessTerm<EnvLorenzRK4<double>> { enum { value = 1 }; };
template <> struct FiniteUnlessTerm<EnvLorenzRK4<float>> { enum { value = 1 }; };
// not terminated <=> sum |a_c - b_c| <= 1000: every difference is finite, hence every a_c, b_c (inf - x is
// inf or NaN).  The six finite values could only overflow `finite`'s sum above ~3e307 -- a magnitude at
// which the quadratic right-hand side overflows inside the same step, so it is never reached.
template <> struct FiniteUnlessTerm<EnvPMSMRK4> { enum { value = 1 }; };

}  // namespace cl
