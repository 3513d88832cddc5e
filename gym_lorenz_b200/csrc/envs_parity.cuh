// Parity env kinds: each reproduces ONE reference env class -- same scheme, same dt, same
// expression order, same dtype rules (NumPy >= 2 / NEP 50), same quirks.  This header is
// compiled ONLY into tu_parity.cu, which is built with -fmad=false so that `a*b + c`
// stays two IEEE roundings exactly like the reference's NumPy scalar arithmetic; the
// few fused operations below are explicit fma() calls.
//
// File:line citations are relative to
//   /root/reference/code/gym-lorenz/gym_lorenz/envs/
#pragma once
#include "kernels_common.cuh"

namespace cl {

// ------------------------------------------------------------------------------------
// Shared right-hand sides (expression order as written in the reference).

// dynamic.py:70-72 / :39-41  (u=10, i=28, o=8/3)
__device__ __forceinline__ void lorenz3_rhs(double x, double y, double z, double& dx, double& dy,
                                            double& dz) {
  const double o = 8.0 / 3.0;
  dx = 10.0 * (y - x);
  dy = ((28.0 * x) - y) - (x * z);
  dz = (x * y) - (o * z);
}

// lorenz_env_transient.py:323-326  (a=10, b=8/3, c=28)
__device__ __forceinline__ void lorenz4_rhs(const double* s, double* d) {
  const double b = 8.0 / 3.0;
  d[0] = (10.0 * (s[1] - s[0])) + s[3];
  d[1] = ((28.0 * s[0]) - s[1]) - (s[0] * s[2]);
  d[2] = (s[0] * s[1]) - (b * s[2]);
  d[3] = ((-s[0]) * s[1]) - (b * s[2]);
}

// lorenz_env_transient_pmsm.py:84-86 / lorenz_env_transient1.py (a=5.46, b=20)
__device__ __forceinline__ void pmsm64_rhs(double x, double y, double z, double& dx, double& dy,
                                           double& dz) {
  dx = (-x) + (y * z);
  dy = ((-y) - (x * z)) + (20.0 * z);
  dz = 5.46 * (y - z);
}

// Correctly-rounded-in-practice x**3: NumPy scalar `x1**3` goes through libm pow (one
// rounding), x*x*x has two.  p + e == x*x exactly; the residual terms restore the bits the
// second multiply would lose (lorenz_env_try.py:9).
__device__ __forceinline__ double cube_cr(double x) {
  const double p = x * x;
  const double e = fma(x, x, -p);
  const double r = p * x;
  if (!isfinite(r)) return r;  // overflow / NaN: same result class as pow(x, 3)
  const double re = fma(p, x, -r);
  return r + fma(e, x, re);
}

// hr_derivatives, lorenz_env_try.py:7-12 with (a,b,c,d,r,s,I,x_rest) =
// (1,3,1,5,0.006,4,3.2,-1.6) from :34-35.  a1/a2 are float32 values widened exactly.
__device__ __forceinline__ void hr_rhs(const double* s, double a1, double a2, double* d) {
  const double x1 = s[0], x2 = s[1], x3 = s[2];
  const double x1sq = x1 * x1;       // x1**2: single rounding either way
  const double x1cu = cube_cr(x1);   // x1**3
  d[0] = (((x2 - (1.0 * x1cu)) + (3.0 * x1sq)) - x3) + 3.2;
  d[1] = ((1.0 - (5.0 * x1sq)) - x2) + a1;
  d[2] = (0.006 * ((4.0 * (x1 - (-1.6))) - x3)) + a2;
}

// One classical RK4 step as lorenz_env_try.py:100-105 writes it:
//   k2 = f(s + dt/2*k1); k3 = f(s + dt/2*k2); k4 = f(s + dt*k3);
//   s += (dt/6.0) * (((k1 + 2*k2) + 2*k3) + k4)
__device__ __forceinline__ void hr_rk4(double* s, double a1, double a2) {
  const double dt = 0.001, hdt = 0.001 / 2, dt6 = 0.001 / 6.0;
  double k1[3], k2[3], k3[3], k4[3], w[3];
  hr_rhs(s, a1, a2, k1);
#pragma unroll
  for (int c = 0; c < 3; ++c) w[c] = s[c] + (hdt * k1[c]);
  hr_rhs(w, a1, a2, k2);
#pragma unroll
  for (int c = 0; c < 3; ++c) w[c] = s[c] + (hdt * k2[c]);
  hr_rhs(w, a1, a2, k3);
#pragma unroll
  for (int c = 0; c < 3; ++c) w[c] = s[c] + (dt * k3[c]);
  hr_rhs(w, a1, a2, k4);
#pragma unroll
  for (int c = 0; c < 3; ++c)
    s[c] = s[c] + (dt6 * (((k1[c] + (2.0 * k2[c])) + (2.0 * k3[c])) + k4[c]));
}

// ====================================================================================
// CL_ENV_LORENZ3 -- dynamic.py:5-93 (outer class).  planes: x y z t
struct EnvLorenz3 {
  typedef double real;
  enum { NSTATE = 4, NINT = 0, OBS = 6, ACT = 3, NOISE = 0 };
  struct S { double x, y, z, t; };
  __device__ static void load(S& s, const KParams& p, int64_t i) {
    s.x = ldp<double>(p, 0, i); s.y = ldp<double>(p, 1, i);
    s.z = ldp<double>(p, 2, i); s.t = ldp<double>(p, 3, i);
  }
  __device__ static void store(const S& s, const KParams& p, int64_t i) {
    stp<double>(p, 0, i, s.x); stp<double>(p, 1, i, s.y);
    stp<double>(p, 2, i, s.z); stp<double>(p, 3, i, s.t);
  }
  __device__ static bool uses_noise(const KParams&) { return false; }
  __device__ static bool finite(const S& s) { return isfinite(s.x + s.y + s.z); }
  __device__ static bool time_limit(const KParams& p, int32_t n) {
    return p.max_steps > 0 && n >= p.max_steps;
  }
  __device__ static void prepare(S&, const KParams&, bool) {}
  __device__ static void init_persistent(S&, const KParams&, const Stream&) {}
  __device__ static void observe(const S& s, double* obs) {
    obs[0] = s.x; obs[1] = s.y; obs[2] = s.z;
    lorenz3_rhs(s.x, s.y, s.z, obs[3], obs[4], obs[5]);
  }
  // dynamic.py:35-47
  __device__ static void reset(S& s, const KParams&, const Stream& rng, double* obs) {
    double u[3];
    draw_uniform<3>(rng, TAG_RESET, -30.0, 30.0, u);
    s.x = u[0]; s.y = u[1]; s.z = u[2]; s.t = 0.0;
    observe(s, obs);
  }
  // dynamic.py:61-90
  __device__ static void step(S& s, const KParams&, const float* a, const double*, double* obs,
                              double& rew, bool& term) {
    const double u1 = (double)clipf(a[0], -500.0f, 500.0f);
    const double u2 = (double)clipf(a[1], -500.0f, 500.0f);
    const double u3 = (double)clipf(a[2], -500.0f, 500.0f);
    double dx, dy, dz;
    lorenz3_rhs(s.x, s.y, s.z, dx, dy, dz);
    s.x = (s.x + (dx * 0.01)) + u1;
    s.y = (s.y + (dy * 0.01)) + u2;
    s.z = (s.z + (dz * 0.01)) + u3;
    observe(s, obs);  // state2 == 0 -> obs = state0
    rew = -(((0.0 + fabs(obs[0])) + fabs(obs[1])) + fabs(obs[2]));
    s.t = s.t + 0.01;
    term = (s.t == 10.0);  // bug-compatible: never true from accumulated 0.01
  }
};

// ====================================================================================
// CL_ENV_LORENZ3_PAIR -- dynamic.py:109-233 (nested class): target [s12, f(s12)] is drawn at
// reset and never advanced (:151-156, advance code commented :194-219).
// planes: x y z t  tx ty tz tdx tdy tdz
struct EnvLorenz3Pair {
  typedef double real;
  enum { NSTATE = 10, NINT = 0, OBS = 6, ACT = 3, NOISE = 0 };
  struct S { double x, y, z, t, g[6]; };
  __device__ static void load(S& s, const KParams& p, int64_t i) {
    s.x = ldp<double>(p, 0, i); s.y = ldp<double>(p, 1, i);
    s.z = ldp<double>(p, 2, i); s.t = ldp<double>(p, 3, i);
#pragma unroll
    for (int c = 0; c < 6; ++c) s.g[c] = ldp<double>(p, 4 + c, i);
  }
  __device__ static void store(const S& s, const KParams& p, int64_t i) {
    stp<double>(p, 0, i, s.x); stp<double>(p, 1, i, s.y);
    stp<double>(p, 2, i, s.z); stp<double>(p, 3, i, s.t);
#pragma unroll
    for (int c = 0; c < 6; ++c) stp<double>(p, 4 + c, i, s.g[c]);
  }
  __device__ static bool uses_noise(const KParams&) { return false; }
  __device__ static bool finite(const S& s) { return isfinite(s.x + s.y + s.z); }
  __device__ static bool time_limit(const KParams& p, int32_t n) {
    return p.max_steps > 0 && n >= p.max_steps;
  }
  __device__ static void prepare(S&, const KParams&, bool) {}
  __device__ static void init_persistent(S&, const KParams&, const Stream&) {}
  __device__ static void observe(const S& s, double* obs) {
    double o[6];
    o[0] = s.x; o[1] = s.y; o[2] = s.z;
    lorenz3_rhs(s.x, s.y, s.z, o[3], o[4], o[5]);
#pragma unroll
    for (int c = 0; c < 6; ++c) obs[c] = o[c] - s.g[c];
  }
  // dynamic.py:142-158
  __device__ static void reset(S& s, const KParams&, const Stream& rng, double* obs) {
    double u[6];
    draw_uniform<6>(rng, TAG_RESET, -20.0, 20.0, u);
    s.x = u[0]; s.y = u[1]; s.z = u[2]; s.t = 0.0;
    s.g[0] = u[3]; s.g[1] = u[4]; s.g[2] = u[5];
    lorenz3_rhs(u[3], u[4], u[5], s.g[3], s.g[4], s.g[5]);
    observe(s, obs);
  }
  // dynamic.py:174-230
  __device__ static void step(S& s, const KParams&, const float* a, const double*, double* obs,
                              double& rew, bool& term) {
    const double u1 = (double)clipf(a[0], -500.0f, 500.0f);
    const double u2 = (double)clipf(a[1], -500.0f, 500.0f);
    const double u3 = (double)clipf(a[2], -500.0f, 500.0f);
    double dx, dy, dz;
    lorenz3_rhs(s.x, s.y, s.z, dx, dy, dz);
    s.x = (s.x + (dx * 0.01)) + u1;
    s.y = (s.y + (dy * 0.01)) + u2;
    s.z = (s.z + (dz * 0.01)) + u3;
    observe(s, obs);
    rew = -(((0.0 + fabs(obs[0])) + fabs(obs[1])) + fabs(obs[2]));
    s.t = s.t + 0.01;
    term = (s.t == 10.0);
  }
};

// ====================================================================================
// CL_ENV_LORENZ4_PAIR -- lorenz_env_transient.py:247-376.  Two free-running 4-state systems;
// the action is clipped (:316-318) and then never used.  planes: a0..a3 b0..b3 t
struct EnvLorenz4Pair {
  typedef double real;
  enum { NSTATE = 9, NINT = 0, OBS = 8, ACT = 3, NOISE = 0 };
  struct S { double a[4], b[4], t; };
  __device__ static void load(S& s, const KParams& p, int64_t i) {
#pragma unroll
    for (int c = 0; c < 4; ++c) { s.a[c] = ldp<double>(p, c, i); s.b[c] = ldp<double>(p, 4 + c, i); }
    s.t = ldp<double>(p, 8, i);
  }
  __device__ static void store(const S& s, const KParams& p, int64_t i) {
#pragma unroll
    for (int c = 0; c < 4; ++c) { stp<double>(p, c, i, s.a[c]); stp<double>(p, 4 + c, i, s.b[c]); }
    stp<double>(p, 8, i, s.t);
  }
  __device__ static bool uses_noise(const KParams&) { return false; }
  __device__ static bool finite(const S& s) {
    return isfinite(s.a[0] + s.a[1] + s.a[2] + s.a[3] + s.b[0] + s.b[1] + s.b[2] + s.b[3]);
  }
  __device__ static bool time_limit(const KParams& p, int32_t n) {
    return p.max_steps > 0 && n >= p.max_steps;
  }
  __device__ static void prepare(S&, const KParams&, bool) {}
  __device__ static void init_persistent(S&, const KParams&, const Stream&) {}
  __device__ static void observe(const S& s, double* obs) {
    double da[4], db[4];
    lorenz4_rhs(s.a, da);
    lorenz4_rhs(s.b, db);
#pragma unroll
    for (int c = 0; c < 4; ++c) { obs[c] = s.a[c] - s.b[c]; obs[4 + c] = da[c] - db[c]; }
  }
  // lorenz_env_transient.py:275-297
  __device__ static void reset(S& s, const KParams&, const Stream& rng, double* obs) {
    double u[8];
    draw_uniform<8>(rng, TAG_RESET, 0.0, 5.0, u);
#pragma unroll
    for (int c = 0; c < 4; ++c) { s.a[c] = u[c]; s.b[c] = u[4 + c]; }
    s.t = 0.0;
    observe(s, obs);
  }
  // lorenz_env_transient.py:314-373
  __device__ static void step(S& s, const KParams&, const float*, const double*, double* obs,
                              double& rew, bool& term) {
    double d[4];
    lorenz4_rhs(s.a, d);
#pragma unroll
    for (int c = 0; c < 4; ++c) s.a[c] = s.a[c] + (d[c] * 0.001);
    lorenz4_rhs(s.b, d);
#pragma unroll
    for (int c = 0; c < 4; ++c) s.b[c] = s.b[c] + (d[c] * 0.001);
    observe(s, obs);
    rew = -((((0.0 + fabs(obs[0])) + fabs(obs[1])) + fabs(obs[2])) + fabs(obs[3]));
    s.t = s.t + 0.001;
    term = (s.t == 5.0) || (rew < -1e6);
  }
};

// ====================================================================================
// CL_ENV_HR_SYNC -- lorenz_env_try.py:13-179.  planes: m0 m1 m2 s0 s1 s2 sigma f0 f1
// (f0,f1 = filtered_action, float32 values held exactly in f64 planes)
struct EnvHRSync {
  typedef double real;
  enum { NSTATE = 9, NINT = 0, OBS = 6, ACT = 2, NOISE = 3 };
  struct S { double m[3], s[3], sigma; float f[2]; };
  __device__ static void load(S& s, const KParams& p, int64_t i) {
#pragma unroll
    for (int c = 0; c < 3; ++c) { s.m[c] = ldp<double>(p, c, i); s.s[c] = ldp<double>(p, 3 + c, i); }
    s.sigma = ldp<double>(p, 6, i);
    s.f[0] = (float)ldp<double>(p, 7, i);
    s.f[1] = (float)ldp<double>(p, 8, i);
  }
  __device__ static void store(const S& s, const KParams& p, int64_t i) {
#pragma unroll
    for (int c = 0; c < 3; ++c) { stp<double>(p, c, i, s.m[c]); stp<double>(p, 3 + c, i, s.s[c]); }
    stp<double>(p, 6, i, s.sigma);
    stp<double>(p, 7, i, (double)s.f[0]);
    stp<double>(p, 8, i, (double)s.f[1]);
  }
  __device__ static bool uses_noise(const KParams& p) { return (p.flags & CL_F_ADD_NOISE) != 0; }
  __device__ static bool finite(const S& s) {
    return isfinite(s.m[0] + s.m[1] + s.m[2] + s.s[0] + s.s[1] + s.s[2]);
  }
  __device__ static bool time_limit(const KParams& p, int32_t n) {
    return p.max_steps > 0 && n >= p.max_steps;
  }
  __device__ static void prepare(S&, const KParams&, bool) {}
  __device__ static void init_persistent(S&, const KParams&, const Stream&) {}
  // lorenz_env_try.py:49-78
  __device__ static void reset(S& s, const KParams& p, const Stream& rng, double* obs) {
    double u[6];
    draw_uniform<6>(rng, TAG_RESET, -10.0, 20.0, u);
#pragma unroll
    for (int c = 0; c < 3; ++c) { s.m[c] = u[c]; s.s[c] = u[3 + c]; }
    s.f[0] = 0.0f; s.f[1] = 0.0f;
    if (p.flags & CL_F_ADD_NOISE) {
      if (p.flags & CL_F_EVAL_MODE) {
        s.sigma = 2.0;
      } else {
        double v[1];
        draw_uniform<1>(rng, TAG_RESET + 3u, 0.0, 2.0, v);
        s.sigma = v[0];
      }
    } else {
      s.sigma = 0.0;
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      obs[c] = clipd((s.m[c] - s.s[c]) / 50.0, -1.0, 1.0);  // error part clipped in reset only (:73)
      obs[3 + c] = clipd(s.m[c] / 20.0, -1.0, 1.0);
    }
  }
  // lorenz_env_try.py:80-179
  __device__ static void step(S& s, const KParams& p, const float* a, const double* nz, double* obs,
                              double& rew, bool& term) {
    if (p.flags & CL_F_ADD_FILTER) {
      // (1 - 0.95) and 0.95 are weak Python floats -> rounded to f32, f32 arithmetic (:86)
      const float c0 = (float)(1 - 0.95), c1 = (float)0.95;
      s.f[0] = __fadd_rn(__fmul_rn(c0, s.f[0]), __fmul_rn(c1, a[0]));
      s.f[1] = __fadd_rn(__fmul_rn(c0, s.f[1]), __fmul_rn(c1, a[1]));
    } else {
      s.f[0] = a[0]; s.f[1] = a[1];
    }
    // clip(...)*100.0 evaluated in float32, then widened (:92-93; SURVEY hard part 2)
    const double a1 = (double)__fmul_rn(clipf(s.f[0], -1.0f, 1.0f), 100.0f);
    const double a2 = (double)__fmul_rn(clipf(s.f[1], -1.0f, 1.0f), 100.0f);
    hr_rk4(s.m, 0.0, 0.0);
    hr_rk4(s.s, a1, a2);
    if (p.flags & CL_F_ADD_NOISE) {
#pragma unroll
      for (int c = 0; c < 3; ++c) s.m[c] = s.m[c] + ((0.0 + s.sigma * nz[c]) * 0.001);  // :136-137
    }
    double err[3], ne[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      err[c] = s.m[c] - s.s[c];
      ne[c] = err[c] / 50.0;
      obs[c] = ne[c];  // unclipped in step (:151)
      obs[3 + c] = clipd(s.m[c] / 20.0, -1.0, 1.0);
    }
    // :165  -np.sum(|ne|) - 0.050*np.sum(np.square(action)); the action term is float32
    const float sq = __fadd_rn(__fmul_rn(a[0], a[0]), __fmul_rn(a[1], a[1]));
    const float pen = __fmul_rn((float)0.050, sq);
    rew = (-((fabs(ne[0]) + fabs(ne[1])) + fabs(ne[2]))) - (double)pen;
    term = false;
    if (fabs(err[0]) > 70.0 || fabs(err[1]) > 70.0 || fabs(err[2]) > 70.0) {
      term = true;
      rew = -2000.0;
    }
  }
};

// ====================================================================================
// CL_ENV_PMSM_SYNC -- lorenz_env_try_pmsm.py:7-184.  float32 end to end under NumPy >= 2.
// planes (f32): a0 a1 a2 (state1)  b0 b1 b2 (state2)  lambda m_t v_t ; aux_int: adam_step
struct EnvPMSMSync {
  typedef float real;
  enum { NSTATE = 9, NINT = 1, OBS = 6, ACT = 2, NOISE = 3 };
  struct S { float a[3], b[3], lam, m, v; int32_t adam; };
  __device__ static void load(S& s, const KParams& p, int64_t i) {
#pragma unroll
    for (int c = 0; c < 3; ++c) { s.a[c] = ldp<float>(p, c, i); s.b[c] = ldp<float>(p, 3 + c, i); }
    s.lam = ldp<float>(p, 6, i); s.m = ldp<float>(p, 7, i); s.v = ldp<float>(p, 8, i);
    s.adam = __ldcg(p.aux_int + i);  // L1-bypassing like every per-env load (see k_rollout_dyn)
  }
  __device__ static void store(const S& s, const KParams& p, int64_t i) {
#pragma unroll
    for (int c = 0; c < 3; ++c) { stp<float>(p, c, i, s.a[c]); stp<float>(p, 3 + c, i, s.b[c]); }
    stp<float>(p, 6, i, s.lam); stp<float>(p, 7, i, s.m); stp<float>(p, 8, i, s.v);
    p.aux_int[i] = s.adam;
  }
  __device__ static bool uses_noise(const KParams& p) { return (p.flags & CL_F_ADD_NOISE) != 0; }
  __device__ static bool finite(const S& s) {
    return isfinite(s.a[0] + s.a[1] + s.a[2] + s.b[0] + s.b[1] + s.b[2]);
  }
  // env-internal truncation at 2000 (:179-180) plus the registered TimeLimit
  __device__ static bool time_limit(const KParams& p, int32_t n) {
    return n >= 2000 || (p.max_steps > 0 && n >= p.max_steps);
  }
  __device__ static void prepare(S&, const KParams&, bool) {}
  __device__ static void init_persistent(S& s, const KParams&, const Stream&) {
    s.lam = 0.0f; s.m = 0.0f; s.v = 0.0f; s.adam = 0;  // :16,25-27
  }
  // _get_derivatives, :51-58.  g=20, sg=5.46 rounded to f32 (weak Python floats).  The
  // optional noise is a float64 array: `f32 + f64 -> f64`, rounded back to f32 by the
  // np.array(..., dtype=float32) at :58.
  __device__ static void rhs(const float* x, float a1, float a2, const double* nz, bool noisy, float* d) {
    const float g = 20.0f, sg = (float)5.46;
    const float d0 = __fadd_rn(__fadd_rn(-x[0], __fmul_rn(x[1], x[2])), a1);
    const float d1 = __fadd_rn(__fadd_rn(__fsub_rn(-x[1], __fmul_rn(x[0], x[2])), __fmul_rn(g, x[2])), a2);
    const float d2 = __fmul_rn(sg, __fsub_rn(x[1], x[2]));
    if (noisy) {
      d[0] = (float)((double)d0 + nz[0]);
      d[1] = (float)((double)d1 + nz[1]);
      d[2] = (float)((double)d2 + nz[2]);
    } else {
      d[0] = d0; d[1] = d1; d[2] = d2;
    }
  }
  // :59-75
  __device__ static void reset(S& s, const KParams&, const Stream& rng, float* obs) {
    double u[6];
    draw_uniform<6>(rng, TAG_RESET, -30.0, 30.0, u);
#pragma unroll
    for (int c = 0; c < 3; ++c) { s.a[c] = (float)u[c]; s.b[c] = (float)u[3 + c]; }
    float da[3], db[3];
    rhs(s.a, 0.0f, 0.0f, nullptr, false, da);
    rhs(s.b, 0.0f, 0.0f, nullptr, false, db);
#pragma unroll
    for (int c = 0; c < 3; ++c) { obs[c] = __fsub_rn(s.a[c], s.b[c]); obs[3 + c] = __fsub_rn(da[c], db[c]); }
  }
  // :76-184
  __device__ static void step(S& s, const KParams& p, const float* act, const double* nz, float* obs,
                              float& rew, bool& term) {
    const bool noisy = (p.flags & CL_F_ADD_NOISE) != 0;
    double n3[3] = {0.0, 0.0, 0.0};
    if (noisy) { n3[0] = 0.0 + 3.0 * nz[0]; n3[1] = 0.0 + 3.0 * nz[1]; n3[2] = 0.0 + 3.0 * nz[2]; }  // :80
    const float a1 = __fmul_rn(clipf(act[0], -1.0f, 1.0f), 50.0f);  // :81-82
    const float a2 = __fmul_rn(clipf(act[1], -1.0f, 1.0f), 50.0f);
    const float dt = (float)0.001;
    float da[3], db[3];
    rhs(s.a, 0.0f, 0.0f, nullptr, false, da);   // :88
    rhs(s.b, a1, a2, n3, noisy, db);             // :89-90
#pragma unroll
    for (int c = 0; c < 3; ++c) {                // :92-93
      s.a[c] = __fadd_rn(s.a[c], __fmul_rn(da[c], dt));
      s.b[c] = __fadd_rn(s.b[c], __fmul_rn(db[c], dt));
    }
    rhs(s.a, 0.0f, 0.0f, nullptr, false, da);   // :95
    rhs(s.b, a1, a2, n3, noisy, db);             // :96-97
    float e[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      obs[c] = __fsub_rn(s.a[c], s.b[c]);
      obs[3 + c] = __fsub_rn(da[c], db[c]);
      e[c] = fabsf(obs[c]);
    }
    const float esum = __fadd_rn(__fadd_rn(e[0], e[1]), e[2]);  // :108
    // Adam-based dual ascent on lambda, :113-138 (state persists across episodes)
    const float grad = __fsub_rn(5.0f, esum);
    s.adam += 1;
    s.m = __fadd_rn(__fmul_rn((float)0.9, s.m), __fmul_rn((float)(1 - 0.9), grad));
    s.v = __fadd_rn(__fmul_rn((float)0.999, s.v), __fmul_rn((float)(1 - 0.999), __fmul_rn(grad, grad)));
    const float bc1 = s.adam < p.bc1_n ? p.bc1[s.adam] : 1.0f;  // (float)(1 - 0.9**n)
    const float bc2 = s.adam < p.bc2_n ? p.bc2[s.adam] : 1.0f;  // (float)(1 - 0.999**n)
    const float mhat = __fdiv_rn(s.m, bc1);
    const float vhat = __fdiv_rn(s.v, bc2);
    const float upd = __fdiv_rn(__fmul_rn((float)0.001, mhat), __fadd_rn(__fsqrt_rn(vhat), (float)1e-8));
    s.lam = clipf(__fsub_rn(s.lam, upd), 0.0f, 0.5f);
    // fractional penalty :158-160 -- float32 pow; computed via double pow and rounded once
    const float al = (float)p.alpha;
    float fp[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) fp[c] = (float)pow_pos((double)__fadd_rn(e[c], (float)1e-6), (double)al);
    const float frac = __fadd_rn(__fadd_rn(fp[0], fp[1]), fp[2]);
    // :165 uses the RAW action
    const float apen = __fmul_rn(s.lam, __fadd_rn(__fmul_rn(act[0], act[0]), __fmul_rn(act[1], act[1])));
    rew = __fsub_rn(__fsub_rn(-esum, frac), apen);
    term = false;
    if (esum > 1000.0f) { rew = -1000.0f; term = true; }  // :174-176
  }
};

// ====================================================================================
// CL_ENV_PMSM_CLASSIC -- lorenz_env_transient_pmsm.py:17-137.  planes: a0 a1 a2 b0 b1 b2 t
// Slave gets u*20 (float32 product) and N(0,3) noise in the derivative on every step.
struct EnvPMSMClassic {
  typedef double real;
  enum { NSTATE = 7, NINT = 0, OBS = 6, ACT = 2, NOISE = 3 };
  struct S { double a[3], b[3], t; };
  __device__ static void load(S& s, const KParams& p, int64_t i) {
#pragma unroll
    for (int c = 0; c < 3; ++c) { s.a[c] = ldp<double>(p, c, i); s.b[c] = ldp<double>(p, 3 + c, i); }
    s.t = ldp<double>(p, 6, i);
  }
  __device__ static void store(const S& s, const KParams& p, int64_t i) {
#pragma unroll
    for (int c = 0; c < 3; ++c) { stp<double>(p, c, i, s.a[c]); stp<double>(p, 3 + c, i, s.b[c]); }
    stp<double>(p, 6, i, s.t);
  }
  __device__ static bool uses_noise(const KParams&) { return true; }
  __device__ static bool finite(const S& s) {
    return isfinite(s.a[0] + s.a[1] + s.a[2] + s.b[0] + s.b[1] + s.b[2]);
  }
  __device__ static bool time_limit(const KParams& p, int32_t n) {
    return p.max_steps > 0 && n >= p.max_steps;
  }
  __device__ static void prepare(S&, const KParams&, bool) {}
  __device__ static void init_persistent(S&, const KParams&, const Stream&) {}
  __device__ static void observe(const S& s, double* obs) {
    double da[3], db[3];
    pmsm64_rhs(s.a[0], s.a[1], s.a[2], da[0], da[1], da[2]);
    pmsm64_rhs(s.b[0], s.b[1], s.b[2], db[0], db[1], db[2]);
#pragma unroll
    for (int c = 0; c < 3; ++c) { obs[c] = s.a[c] - s.b[c]; obs[3 + c] = da[c] - db[c]; }
  }
  // :43-62
  __device__ static void reset(S& s, const KParams&, const Stream& rng, double* obs) {
    double u[6];
    draw_uniform<6>(rng, TAG_RESET, -10.0, 10.0, u);
#pragma unroll
    for (int c = 0; c < 3; ++c) { s.a[c] = u[c]; s.b[c] = u[3 + c]; }
    s.t = 0.0;
    observe(s, obs);
  }
  // :76-133
  __device__ static void step(S& s, const KParams&, const float* act, const double* nz, double* obs,
                              double& rew, bool& term) {
    const double u1 = (double)__fmul_rn(clipf(act[0], -2.0f, 2.0f), 20.0f);
    const double u2 = (double)__fmul_rn(clipf(act[1], -2.0f, 2.0f), 20.0f);
    const double n0 = 0.0 + 3.0 * nz[0], n1 = 0.0 + 3.0 * nz[1], n2 = 0.0 + 3.0 * nz[2];
    double da[3], db[3];
    pmsm64_rhs(s.a[0], s.a[1], s.a[2], da[0], da[1], da[2]);
    pmsm64_rhs(s.b[0], s.b[1], s.b[2], db[0], db[1], db[2]);
    db[0] = (db[0] + u1) + n0;
    db[1] = (db[1] + u2) + n1;
    db[2] = db[2] + n2;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      s.a[c] = s.a[c] + (da[c] * 0.01);
      s.b[c] = s.b[c] + (db[c] * 0.01);
    }
    observe(s, obs);
    const double E = ((0.0 + fabs(obs[0])) + fabs(obs[1])) + fabs(obs[2]);
    rew = (-E) - pow_pos(E, 1.0 / 10);  // :122
    s.t = s.t + 0.01;
    term = (s.t == 5.0) || (rew < -1e6);
  }
};

// ====================================================================================
// CL_ENV_PMSM_SINGLE -- lorenz_env_transient1.py: one PMSM driven to the origin by impulse
// control on x,y.  planes: x y z t
struct EnvPMSMSingle {
  typedef double real;
  enum { NSTATE = 4, NINT = 0, OBS = 6, ACT = 2, NOISE = 0 };
  struct S { double x, y, z, t; };
  __device__ static void load(S& s, const KParams& p, int64_t i) {
    s.x = ldp<double>(p, 0, i); s.y = ldp<double>(p, 1, i);
    s.z = ldp<double>(p, 2, i); s.t = ldp<double>(p, 3, i);
  }
  __device__ static void store(const S& s, const KParams& p, int64_t i) {
    stp<double>(p, 0, i, s.x); stp<double>(p, 1, i, s.y);
    stp<double>(p, 2, i, s.z); stp<double>(p, 3, i, s.t);
  }
  __device__ static bool uses_noise(const KParams&) { return false; }
  __device__ static bool finite(const S& s) { return isfinite(s.x + s.y + s.z); }
  __device__ static bool time_limit(const KParams& p, int32_t n) {
    return p.max_steps > 0 && n >= p.max_steps;
  }
  __device__ static void prepare(S&, const KParams&, bool) {}
  __device__ static void init_persistent(S&, const KParams&, const Stream&) {}
  __device__ static void observe(const S& s, double* obs) {
    obs[0] = s.x; obs[1] = s.y; obs[2] = s.z;
    pmsm64_rhs(s.x, s.y, s.z, obs[3], obs[4], obs[5]);
  }
  __device__ static void reset(S& s, const KParams&, const Stream& rng, double* obs) {
    double u[3];
    draw_uniform<3>(rng, TAG_RESET, -30.0, 30.0, u);
    s.x = u[0]; s.y = u[1]; s.z = u[2]; s.t = 0.0;
    observe(s, obs);
  }
  __device__ static void step(S& s, const KParams&, const float* a, const double*, double* obs,
                              double& rew, bool& term) {
    const double u1 = (double)clipf(a[0], -10.0f, 10.0f);
    const double u2 = (double)clipf(a[1], -10.0f, 10.0f);
    double dx, dy, dz;
    pmsm64_rhs(s.x, s.y, s.z, dx, dy, dz);
    s.x = (s.x + (dx * 0.01)) + u1;
    s.y = (s.y + (dy * 0.01)) + u2;
    s.z = s.z + (dz * 0.01);
    observe(s, obs);
    rew = -(((0.0 + fabs(obs[0])) + fabs(obs[1])) + fabs(obs[2]));
    s.t = s.t + 0.01;
    term = (s.t == 10.0);
  }
};

// ====================================================================================
// CL_ENV_MEMRISTIVE4_PAIR -- lorenz_env_transient2.py: two 4-state memristive systems
// (a=30, b=1, c=36, d=0.5, h=0.003), Euler dt 0.001; the slave gets u*100 (float32 product) on
// dx1, dx2, dx4.  planes: a0..a3 b0..b3 t
__device__ __forceinline__ void memristive4_rhs(const double* s, double* d) {
  const double w = (2.0 * s[3]) * s[3];
  d[0] = 30.0 * ((w * (s[1] - s[0])) + (0.5 * s[0]));
  d[1] = 1.0 * ((w * (s[0] - s[1])) - s[2]);
  d[2] = 36.0 * (s[1] - (0.003 * s[2]));
  d[3] = (s[1] - s[0]) - (0.01 * s[3]);
}

struct EnvMemristive4Pair {
  typedef double real;
  enum { NSTATE = 9, NINT = 0, OBS = 8, ACT = 3, NOISE = 0 };
  struct S { double a[4], b[4], t; };
  __device__ static void load(S& s, const KParams& p, int64_t i) {
#pragma unroll
    for (int c = 0; c < 4; ++c) { s.a[c] = ldp<double>(p, c, i); s.b[c] = ldp<double>(p, 4 + c, i); }
    s.t = ldp<double>(p, 8, i);
  }
  __device__ static void store(const S& s, const KParams& p, int64_t i) {
#pragma unroll
    for (int c = 0; c < 4; ++c) { stp<double>(p, c, i, s.a[c]); stp<double>(p, 4 + c, i, s.b[c]); }
    stp<double>(p, 8, i, s.t);
  }
  __device__ static bool uses_noise(const KParams&) { return false; }
  __device__ static bool finite(const S& s) {
    return isfinite(s.a[0] + s.a[1] + s.a[2] + s.a[3] + s.b[0] + s.b[1] + s.b[2] + s.b[3]);
  }
  __device__ static bool time_limit(const KParams& p, int32_t n) {
    return p.max_steps > 0 && n >= p.max_steps;
  }
  __device__ static void prepare(S&, const KParams&, bool) {}
  __device__ static void init_persistent(S&, const KParams&, const Stream&) {}
  __device__ static void observe(const S& s, double* obs) {
    double da[4], db[4];
    memristive4_rhs(s.a, da);
    memristive4_rhs(s.b, db);
#pragma unroll
    for (int c = 0; c < 4; ++c) { obs[c] = s.a[c] - s.b[c]; obs[4 + c] = da[c] - db[c]; }
  }
  __device__ static void reset(S& s, const KParams&, const Stream& rng, double* obs) {
    double u[8];
    draw_uniform<8>(rng, TAG_RESET, 0.0, 5.0, u);
#pragma unroll
    for (int c = 0; c < 4; ++c) { s.a[c] = u[c]; s.b[c] = u[4 + c]; }
    s.t = 0.0;
    observe(s, obs);
  }
  __device__ static void step(S& s, const KParams&, const float* act, const double*, double* obs,
                              double& rew, bool& term) {
    const double u1 = (double)__fmul_rn(clipf(act[0], -2.0f, 2.0f), 100.0f);
    const double u2 = (double)__fmul_rn(clipf(act[1], -2.0f, 2.0f), 100.0f);
    const double u3 = (double)__fmul_rn(clipf(act[2], -2.0f, 2.0f), 100.0f);
    double d[4];
    memristive4_rhs(s.a, d);
#pragma unroll
    for (int c = 0; c < 4; ++c) s.a[c] = s.a[c] + (d[c] * 0.001);
    memristive4_rhs(s.b, d);
    d[0] = d[0] + u1; d[1] = d[1] + u2; d[3] = d[3] + u3;
#pragma unroll
    for (int c = 0; c < 4; ++c) s.b[c] = s.b[c] + (d[c] * 0.001);
    observe(s, obs);
    const double E = (((0.0 + fabs(obs[0])) + fabs(obs[1])) + fabs(obs[2])) + fabs(obs[3]);
    rew = (-E) - pow_pos(E, 1.0 / 3);
    s.t = s.t + 0.001;
    term = (s.t == 5.0) || (rew < -1e6);
  }
};

// ====================================================================================
// CL_ENV_PMSM_FREE -- lorenz_singlecontrol.py: one uncontrolled PMSM with N(0,3) noise in the
// derivative; `step()` takes no action; `reset()` always starts at (25, 1, -1).  planes: x y z t
struct EnvPMSMFree {
  typedef double real;
  enum { NSTATE = 4, NINT = 0, OBS = 6, ACT = 2, NOISE = 3 };
  struct S { double x, y, z, t; };
  __device__ static void load(S& s, const KParams& p, int64_t i) {
    s.x = ldp<double>(p, 0, i); s.y = ldp<double>(p, 1, i);
    s.z = ldp<double>(p, 2, i); s.t = ldp<double>(p, 3, i);
  }
  __device__ static void store(const S& s, const KParams& p, int64_t i) {
    stp<double>(p, 0, i, s.x); stp<double>(p, 1, i, s.y);
    stp<double>(p, 2, i, s.z); stp<double>(p, 3, i, s.t);
  }
  __device__ static bool uses_noise(const KParams&) { return true; }
  __device__ static bool finite(const S& s) { return isfinite(s.x + s.y + s.z); }
  __device__ static bool time_limit(const KParams& p, int32_t n) {
    return p.max_steps > 0 && n >= p.max_steps;
  }
  __device__ static void prepare(S&, const KParams&, bool) {}
  __device__ static void init_persistent(S&, const KParams&, const Stream&) {}
  __device__ static void observe(const S& s, double* obs) {
    obs[0] = s.x; obs[1] = s.y; obs[2] = s.z;
    pmsm64_rhs(s.x, s.y, s.z, obs[3], obs[4], obs[5]);
  }
  __device__ static void reset(S& s, const KParams&, const Stream&, double* obs) {
    s.x = 25.0; s.y = 1.0; s.z = -1.0; s.t = 0.0;
    observe(s, obs);
  }
  __device__ static void step(S& s, const KParams&, const float*, const double* nz, double* obs,
                              double& rew, bool& term) {
    double dx, dy, dz;
    pmsm64_rhs(s.x, s.y, s.z, dx, dy, dz);
    dx = dx + (0.0 + 3.0 * nz[0]);
    dy = dy + (0.0 + 3.0 * nz[1]);
    dz = dz + (0.0 + 3.0 * nz[2]);
    s.x = s.x + (dx * 0.01);
    s.y = s.y + (dy * 0.01);
    s.z = s.z + (dz * 0.01);
    observe(s, obs);
    rew = -(((0.0 + fabs(obs[0])) + fabs(obs[1])) + fabs(obs[2]));
    s.t = s.t + 0.01;
    term = (s.t == 1000.0);
  }
};

template <> struct StepMinBlocks<EnvLorenz3> { enum { value = 5 }; };
template <> struct StepMinBlocks<EnvLorenz3Pair> { enum { value = 4 }; };
template <> struct StepMinBlocks<EnvLorenz4Pair> { enum { value = 4 }; };
template <> struct StepMinBlocks<EnvHRSync> { enum { value = 4 }; };
template <> struct StepMinBlocks<EnvPMSMSync> { enum { value = 5 }; };
template <> struct StepMinBlocks<EnvPMSMClassic> { enum { value = 4 }; };
template <> struct StepMinBlocks<EnvPMSMSingle> { enum { value = 4 }; };
template <> struct StepMinBlocks<EnvMemristive4Pair> { enum { value = 4 }; };
template <> struct StepMinBlocks<EnvPMSMFree> { enum { value = 5 }; };
template <> struct StepBlockStats<EnvMemristive4Pair> { enum { value = 1 }; };

}  // namespace cl
