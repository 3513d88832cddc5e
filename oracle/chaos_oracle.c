/* chaos_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement (plain C, scalar, one env at a time) of the reference's env `step` /
 * `reset` arithmetic, used only as the checker in tests/, __graft_entry__.smoke() and as
 * the cpu_baseline / --impl reference leg of bench.py.  The product path
 * (gym_lorenz_b200/) never links, imports or calls anything in this directory.
 *
 * Pinning: every parity kind below is checked bit-for-bit (states) / to 1 ulp (libm pow
 * results) against the UNMODIFIED reference classes executed through oracle/ref_loader.py
 * (tests/test_oracle_vs_reference.py, runs where /root/reference exists) and against the
 * committed golden vectors generated from them (tests/golden/, incl. the float32 KATs
 * recovered from the reference's PMSM_Origin_Data.xlsx).  The north-star RK4 x S kinds
 * have no reference class; they are pinned against scipy DOP853 (tests/test_rk4_vs_scipy.py).
 *
 * Build: gcc -O2 -ffp-contract=off -fno-fast-math -fopenmp -shared -fPIC  (oracle/Makefile).
 * -ffp-contract=off keeps every a*b+c as two IEEE roundings, like NumPy scalar math.
 *
 * Citations are relative to /root/reference/code/gym-lorenz/gym_lorenz/envs/.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

enum {
  K_LORENZ3 = 0, K_LORENZ3_PAIR = 1, K_LORENZ4_PAIR = 2, K_HR_SYNC = 3, K_PMSM_SYNC = 4,
  K_PMSM_CLASSIC = 5, K_PMSM_SINGLE = 6, K_LORENZ_RK4 = 7, K_LORENZ_RK4_F32 = 8, K_PMSM_RK4 = 9,
  K_MEMRISTIVE4_PAIR = 10, K_PMSM_FREE = 11
};
enum { F_ADD_NOISE = 1, F_EVAL_MODE = 2, F_ADD_FILTER = 4, F_AUTORESET = 8 };
enum { DONE_TERM = 1, DONE_TRUNC = 2 };

typedef struct orc_cfg {
  int32_t kind, flags, max_episode_steps, substeps;
  int64_t n, n_pad, env_id_base;
  uint64_t seed, step_index;
  double dt, alpha, act_limit, act_gain, param_jitter;
} orc_cfg;

/* ---- Philox4x32-10 (Salmon et al., SC'11) and the uniform/normal mappings ----------- */
static void philox(const uint32_t c_in[4], uint32_t k0, uint32_t k1, uint32_t out[4]) {
  uint32_t c0 = c_in[0], c1 = c_in[1], c2 = c_in[2], c3 = c_in[3];
  for (int r = 0; r < 10; ++r) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
  philox(ctr, key[0], key[1], out);
}

typedef struct { uint32_t id_lo, id_hi, step, k0, k1; } rng_t;
enum { TAG_RESET = 0x00000000u, TAG_NOISE = 0x10000000u, TAG_ACTION = 0x20000000u, TAG_PARAM = 0x30000000u };

static rng_t make_rng(const orc_cfg* c, int64_t i, uint64_t step) {
  uint64_t gid = (uint64_t)(c->env_id_base + i);
  rng_t r;
  r.id_lo = (uint32_t)gid; r.id_hi = (uint32_t)(gid >> 32); r.step = (uint32_t)step;
  r.k0 = (uint32_t)c->seed;
  r.k1 = (uint32_t)(c->seed >> 32) ^ (uint32_t)((step >> 32) & 0x00FFFFFFu);
  return r;
}
static void rng_draw(const rng_t* r, uint32_t tag, uint32_t out[4]) {
  uint32_t c[4] = {r->id_lo, r->id_hi, r->step, tag};
  philox(c, r->k0, r->k1, out);
}
static double u01_53(uint32_t a, uint32_t b) {
  return (double)(((uint64_t)(a >> 5) << 26) | (uint64_t)(b >> 6)) * (1.0 / 9007199254740992.0);
}
static void draw_uniform(const rng_t* r, uint32_t tag, int n, double lo, double hi, double* out) {
  for (int b = 0; b < (n + 1) / 2; ++b) {
    uint32_t w[4];
    rng_draw(r, tag + (uint32_t)b, w);
    out[2 * b] = lo + (hi - lo) * u01_53(w[0], w[1]);
    if (2 * b + 1 < n) out[2 * b + 1] = lo + (hi - lo) * u01_53(w[2], w[3]);
  }
}
static void draw_normal(const rng_t* r, uint32_t tag, int n, double* out) {
  for (int b = 0; b < (n + 1) / 2; ++b) {
    uint32_t w[4];
    rng_draw(r, tag + (uint32_t)b, w);
    double u1 = 1.0 - u01_53(w[0], w[1]), u2 = u01_53(w[2], w[3]);
    double rad = sqrt(-2.0 * log(u1)), ang = 2.0 * M_PI * u2;
    out[2 * b] = rad * cos(ang);
    if (2 * b + 1 < n) out[2 * b + 1] = rad * sin(ang);
  }
}

static float clipf(float a, float lo, float hi) { return a < lo ? lo : (a > hi ? hi : a); }
static double clipd(double a, double lo, double hi) { return a < lo ? lo : (a > hi ? hi : a); }

/* ---- right-hand sides ------------------------------------------------------------------ */
/* dynamic.py:70-72 */
static void lorenz3_rhs(const double* s, double* d) {
  const double u = 10.0, i = 28.0, o = 8.0 / 3.0;
  d[0] = u * (s[1] - s[0]);
  d[1] = i * s[0] - s[1] - s[0] * s[2];
  d[2] = s[0] * s[1] - o * s[2];
}
/* lorenz_env_transient.py:323-326 */
static void lorenz4_rhs(const double* s, double* d) {
  const double a = 10.0, b = 8.0 / 3.0, c = 28.0;
  d[0] = a * (s[1] - s[0]) + s[3];
  d[1] = c * s[0] - s[1] - s[0] * s[2];
  d[2] = s[0] * s[1] - b * s[2];
  d[3] = -s[0] * s[1] - b * s[2];
}
/* lorenz_env_transient_pmsm.py:84-86 */
static void pmsm64_rhs(const double* s, double* d) {
  const double a = 5.46, b = 20.0;
  d[0] = -s[0] + s[1] * s[2];
  d[1] = -s[1] - s[0] * s[2] + b * s[2];
  d[2] = a * (s[1] - s[2]);
}
/* lorenz_env_transient2.py step/reset right-hand side (a=30, b=1, c=36, d=0.5, h=0.003) */
static void memristive4_rhs(const double* s, double* d) {
  const double a = 30.0, b = 1.0, c = 36.0, dd = 0.5, h = 0.003;
  d[0] = a * (2 * s[3] * s[3] * (s[1] - s[0]) + dd * s[0]);
  d[1] = b * (2 * s[3] * s[3] * (s[0] - s[1]) - s[2]);
  d[2] = c * (s[1] - h * s[2]);
  d[3] = s[1] - s[0] - 0.01 * s[3];
}
/* hr_derivatives, lorenz_env_try.py:7-12 (x1**3 and x1**2 are libm pow in NumPy scalars) */
static void hr_rhs(const double* s, double a1, double a2, double* d) {
  const double a = 1.0, b = 3.0, c = 1.0, dd = 5.0, r = 0.006, sp = 4.0, I = 3.2, xr = -1.6;
  const double x1 = s[0], x2 = s[1], x3 = s[2];
  d[0] = x2 - a * pow(x1, 3.0) + b * pow(x1, 2.0) - x3 + I;
  d[1] = c - dd * pow(x1, 2.0) - x2 + a1;
  d[2] = r * (sp * (x1 - xr) - x3) + a2;
}
/* lorenz_env_try.py:100-105 */
static void hr_rk4(double* s, double a1, double a2) {
  const double dt = 0.001;
  double k1[3], k2[3], k3[3], k4[3], w[3];
  hr_rhs(s, a1, a2, k1);
  for (int c = 0; c < 3; ++c) w[c] = s[c] + dt / 2 * k1[c];
  hr_rhs(w, a1, a2, k2);
  for (int c = 0; c < 3; ++c) w[c] = s[c] + dt / 2 * k2[c];
  hr_rhs(w, a1, a2, k3);
  for (int c = 0; c < 3; ++c) w[c] = s[c] + dt * k3[c];
  hr_rhs(w, a1, a2, k4);
  for (int c = 0; c < 3; ++c) s[c] += (dt / 6.0) * (k1[c] + 2 * k2[c] + 2 * k3[c] + k4[c]);
}
/* PMSM_Sync_Env._get_derivatives, lorenz_env_try_pmsm.py:51-58 (float32; noise is float64) */
static void pmsm32_rhs(const float* x, float a1, float a2, const double* nz, int noisy, float* d) {
  const float gamma = 20.0f, sigma = (float)5.46;
  float d0 = -x[0] + x[1] * x[2] + a1;
  float d1 = -x[1] - x[0] * x[2] + gamma * x[2] + a2;
  float d2 = sigma * (x[1] - x[2]);
  if (noisy) {
    d[0] = (float)((double)d0 + nz[0]); d[1] = (float)((double)d1 + nz[1]); d[2] = (float)((double)d2 + nz[2]);
  } else {
    d[0] = d0; d[1] = d1; d[2] = d2;
  }
}

/* ---- north-star RK4 x S (no reference class; textbook RK4, ZOH control) ----------------- */
static void lorenz_par_rhs(const double* q, const double* s, const double* u, double* d) {
  d[0] = q[0] * (s[1] - s[0]) + u[0];
  d[1] = s[0] * (q[1] - s[2]) - s[1] + u[1];
  d[2] = s[0] * s[1] - q[2] * s[2] + u[2];
}
static void pmsm_par_rhs(const double* q, const double* s, const double* u, double* d) {
  d[0] = -s[0] + s[1] * s[2] + u[0];
  d[1] = -s[1] - s[0] * s[2] + q[1] * s[2] + u[1];
  d[2] = q[0] * (s[1] - s[2]);
}
typedef void (*rhs_fn)(const double*, const double*, const double*, double*);
static void rk4_sub(rhs_fn f, const double* q, double* s, const double* u, double h, int S) {
  for (int k = 0; k < S; ++k) {
    double k1[3], k2[3], k3[3], k4[3], w[3];
    f(q, s, u, k1);
    for (int c = 0; c < 3; ++c) w[c] = s[c] + 0.5 * h * k1[c];
    f(q, w, u, k2);
    for (int c = 0; c < 3; ++c) w[c] = s[c] + 0.5 * h * k2[c];
    f(q, w, u, k3);
    for (int c = 0; c < 3; ++c) w[c] = s[c] + h * k3[c];
    f(q, w, u, k4);
    for (int c = 0; c < 3; ++c) s[c] += (h / 6.0) * (k1[c] + 2.0 * k2[c] + 2.0 * k3[c] + k4[c]);
  }
}
static void lorenz_par_rhs_f(const float* q, const float* s, const float* u, float* d) {
  d[0] = q[0] * (s[1] - s[0]) + u[0];
  d[1] = s[0] * (q[1] - s[2]) - s[1] + u[1];
  d[2] = s[0] * s[1] - q[2] * s[2] + u[2];
}
static void rk4_sub_f(const float* q, float* s, const float* u, float h, int S) {
  for (int k = 0; k < S; ++k) {
    float k1[3], k2[3], k3[3], k4[3], w[3];
    lorenz_par_rhs_f(q, s, u, k1);
    for (int c = 0; c < 3; ++c) w[c] = s[c] + 0.5f * h * k1[c];
    lorenz_par_rhs_f(q, w, u, k2);
    for (int c = 0; c < 3; ++c) w[c] = s[c] + 0.5f * h * k2[c];
    lorenz_par_rhs_f(q, w, u, k3);
    for (int c = 0; c < 3; ++c) w[c] = s[c] + h * k3[c];
    lorenz_par_rhs_f(q, w, u, k4);
    for (int c = 0; c < 3; ++c) s[c] += (h / 6.0f) * (k1[c] + 2.0f * k2[c] + 2.0f * k3[c] + k4[c]);
  }
}

/* ---- per-kind geometry -------------------------------------------------------------------- */
static const int kNState[12] = {4, 10, 9, 9, 9, 7, 4, 6, 6, 8, 9, 4};
static const int kObs[12] = {6, 6, 8, 6, 6, 6, 6, 6, 6, 6, 8, 6};
static const int kAct[12] = {3, 3, 3, 2, 2, 2, 2, 3, 3, 2, 3, 2};
static int is_f32(int kind) { return kind == K_PMSM_SYNC || kind == K_LORENZ_RK4_F32; }

int orc_n_state(int kind) { return kNState[kind]; }
int orc_obs_dim(int kind) { return kObs[kind]; }
int orc_act_dim(int kind) { return kAct[kind]; }

/* One env, all state gathered into a double scratch `x[]` (f32 kinds hold exact floats). */
typedef struct { double x[12]; int32_t adam; } env_t;

static int env_finite(int kind, const env_t* e) {
  double s = 0;
  int n = (kind == K_LORENZ3 || kind == K_LORENZ3_PAIR || kind == K_PMSM_SINGLE || kind == K_LORENZ_RK4 ||
           kind == K_LORENZ_RK4_F32 || kind == K_PMSM_FREE) ? 3
          : ((kind == K_LORENZ4_PAIR || kind == K_MEMRISTIVE4_PAIR) ? 8 : 6);
  for (int c = 0; c < n; ++c) s += e->x[c];
  return isfinite(s);
}

/* observation after a state change, per kind (the part shared by reset and step) */
static void observe(int kind, const env_t* e, double* obs) {
  double d[4], d2[4];
  switch (kind) {
    case K_LORENZ3:
      lorenz3_rhs(e->x, d);
      for (int c = 0; c < 3; ++c) { obs[c] = e->x[c]; obs[3 + c] = d[c]; }
      break;
    case K_LORENZ3_PAIR:
      lorenz3_rhs(e->x, d);
      for (int c = 0; c < 3; ++c) { obs[c] = e->x[c] - e->x[4 + c]; obs[3 + c] = d[c] - e->x[7 + c]; }
      break;
    case K_LORENZ4_PAIR:
      lorenz4_rhs(e->x, d); lorenz4_rhs(e->x + 4, d2);
      for (int c = 0; c < 4; ++c) { obs[c] = e->x[c] - e->x[4 + c]; obs[4 + c] = d[c] - d2[c]; }
      break;
    case K_PMSM_CLASSIC:
      pmsm64_rhs(e->x, d); pmsm64_rhs(e->x + 3, d2);
      for (int c = 0; c < 3; ++c) { obs[c] = e->x[c] - e->x[3 + c]; obs[3 + c] = d[c] - d2[c]; }
      break;
    case K_PMSM_SINGLE: case K_PMSM_FREE:
      pmsm64_rhs(e->x, d);
      for (int c = 0; c < 3; ++c) { obs[c] = e->x[c]; obs[3 + c] = d[c]; }
      break;
    case K_MEMRISTIVE4_PAIR:
      memristive4_rhs(e->x, d); memristive4_rhs(e->x + 4, d2);
      for (int c = 0; c < 4; ++c) { obs[c] = e->x[c] - e->x[4 + c]; obs[4 + c] = d[c] - d2[c]; }
      break;
    case K_LORENZ_RK4: {
      const double z[3] = {0, 0, 0};
      lorenz_par_rhs(e->x + 3, e->x, z, d);
      for (int c = 0; c < 3; ++c) { obs[c] = e->x[c]; obs[3 + c] = d[c]; }
      break;
    }
    case K_LORENZ_RK4_F32: {
      float q[3] = {(float)e->x[3], (float)e->x[4], (float)e->x[5]};
      float s[3] = {(float)e->x[0], (float)e->x[1], (float)e->x[2]}, z[3] = {0, 0, 0}, df[3];
      lorenz_par_rhs_f(q, s, z, df);
      for (int c = 0; c < 3; ++c) { obs[c] = s[c]; obs[3 + c] = df[c]; }
      break;
    }
    case K_PMSM_RK4: {
      const double z[3] = {0, 0, 0};
      pmsm_par_rhs(e->x + 6, e->x, z, d); pmsm_par_rhs(e->x + 6, e->x + 3, z, d2);
      for (int c = 0; c < 3; ++c) { obs[c] = e->x[c] - e->x[3 + c]; obs[3 + c] = d[c] - d2[c]; }
      break;
    }
    default: break;
  }
}

static void env_reset(const orc_cfg* cfg, env_t* e, const rng_t* rng, double* obs) {
  double u[8];
  switch (cfg->kind) {
    case K_LORENZ3: case K_PMSM_SINGLE: case K_LORENZ_RK4: case K_LORENZ_RK4_F32:  /* dynamic.py:35-47 */
      draw_uniform(rng, TAG_RESET, 3, -30.0, 30.0, u);
      for (int c = 0; c < 3; ++c) e->x[c] = cfg->kind == K_LORENZ_RK4_F32 ? (double)(float)u[c] : u[c];
      if (cfg->kind == K_LORENZ3 || cfg->kind == K_PMSM_SINGLE) e->x[3] = 0.0;
      observe(cfg->kind, e, obs);
      break;
    case K_LORENZ3_PAIR: {  /* dynamic.py:142-158 */
      draw_uniform(rng, TAG_RESET, 6, -20.0, 20.0, u);
      for (int c = 0; c < 3; ++c) { e->x[c] = u[c]; e->x[4 + c] = u[3 + c]; }
      e->x[3] = 0.0;
      lorenz3_rhs(u + 3, e->x + 7);
      observe(cfg->kind, e, obs);
      break;
    }
    case K_PMSM_FREE:  /* lorenz_singlecontrol.py reset: fixed initial condition, no draw */
      e->x[0] = 25.0; e->x[1] = 1.0; e->x[2] = -1.0; e->x[3] = 0.0;
      observe(cfg->kind, e, obs);
      break;
    case K_MEMRISTIVE4_PAIR:  /* lorenz_env_transient2.py reset */
    case K_LORENZ4_PAIR:  /* lorenz_env_transient.py:275-297 */
      draw_uniform(rng, TAG_RESET, 8, 0.0, 5.0, u);
      for (int c = 0; c < 8; ++c) e->x[c] = u[c];
      e->x[8] = 0.0;
      observe(cfg->kind, e, obs);
      break;
    case K_HR_SYNC:  /* lorenz_env_try.py:49-78 */
      draw_uniform(rng, TAG_RESET, 6, -10.0, 20.0, u);
      for (int c = 0; c < 6; ++c) e->x[c] = u[c];
      e->x[7] = 0.0; e->x[8] = 0.0;
      if (cfg->flags & F_ADD_NOISE) {
        if (cfg->flags & F_EVAL_MODE) e->x[6] = 2.0;
        else { double v[1]; draw_uniform(rng, TAG_RESET + 3u, 1, 0.0, 2.0, v); e->x[6] = v[0]; }
      } else e->x[6] = 0.0;
      for (int c = 0; c < 3; ++c) {
        obs[c] = clipd((e->x[c] - e->x[3 + c]) / 50.0, -1.0, 1.0);
        obs[3 + c] = clipd(e->x[c] / 20.0, -1.0, 1.0);
      }
      break;
    case K_PMSM_SYNC: {  /* lorenz_env_try_pmsm.py:59-75 */
      draw_uniform(rng, TAG_RESET, 6, -30.0, 30.0, u);
      float a[3], b[3], da[3], db[3];
      for (int c = 0; c < 3; ++c) { a[c] = (float)u[c]; b[c] = (float)u[3 + c]; e->x[c] = a[c]; e->x[3 + c] = b[c]; }
      pmsm32_rhs(a, 0.0f, 0.0f, NULL, 0, da);
      pmsm32_rhs(b, 0.0f, 0.0f, NULL, 0, db);
      for (int c = 0; c < 3; ++c) { obs[c] = (float)(a[c] - b[c]); obs[3 + c] = (float)(da[c] - db[c]); }
      break;
    }
    case K_PMSM_CLASSIC:  /* lorenz_env_transient_pmsm.py:43-62 */
      draw_uniform(rng, TAG_RESET, 6, -10.0, 10.0, u);
      for (int c = 0; c < 6; ++c) e->x[c] = u[c];
      e->x[6] = 0.0;
      observe(cfg->kind, e, obs);
      break;
    case K_PMSM_RK4:
      draw_uniform(rng, TAG_RESET, 6, -30.0, 30.0, u);
      for (int c = 0; c < 6; ++c) e->x[c] = u[c];
      observe(cfg->kind, e, obs);
      break;
  }
}

static void env_step(const orc_cfg* cfg, env_t* e, const float* a, const double* nz, double* obs,
                     double* rew_out, int* term_out) {
  double d[4], rew = 0.0;
  int term = 0;
  switch (cfg->kind) {
    case K_LORENZ3: case K_LORENZ3_PAIR: {  /* dynamic.py:61-90 / :174-230 */
      double u1 = (double)clipf(a[0], -500.0f, 500.0f), u2 = (double)clipf(a[1], -500.0f, 500.0f),
             u3 = (double)clipf(a[2], -500.0f, 500.0f);
      lorenz3_rhs(e->x, d);
      e->x[0] = e->x[0] + d[0] * 0.01 + u1;
      e->x[1] = e->x[1] + d[1] * 0.01 + u2;
      e->x[2] = e->x[2] + d[2] * 0.01 + u3;
      observe(cfg->kind, e, obs);
      rew = 0.0; for (int c = 0; c < 3; ++c) rew = rew + fabs(obs[c]);
      rew = -rew;
      e->x[3] = e->x[3] + 0.01;
      term = (e->x[3] == 10.0);
      break;
    }
    case K_LORENZ4_PAIR: {  /* lorenz_env_transient.py:314-373 (action ignored) */
      lorenz4_rhs(e->x, d);
      for (int c = 0; c < 4; ++c) e->x[c] = e->x[c] + d[c] * 0.001;
      lorenz4_rhs(e->x + 4, d);
      for (int c = 0; c < 4; ++c) e->x[4 + c] = e->x[4 + c] + d[c] * 0.001;
      observe(cfg->kind, e, obs);
      rew = 0.0; for (int c = 0; c < 4; ++c) rew = rew + fabs(obs[c]);
      rew = -rew;
      e->x[8] = e->x[8] + 0.001;
      term = (e->x[8] == 5.0) || (rew < -1e6);
      break;
    }
    case K_HR_SYNC: {  /* lorenz_env_try.py:80-179 */
      float f0 = (float)e->x[7], f1 = (float)e->x[8];
      if (cfg->flags & F_ADD_FILTER) {
        const float c0 = (float)(1 - 0.95), c1 = (float)0.95;
        f0 = c0 * f0 + c1 * a[0]; f1 = c0 * f1 + c1 * a[1];
      } else { f0 = a[0]; f1 = a[1]; }
      e->x[7] = f0; e->x[8] = f1;
      float a1f = clipf(f0, -1.0f, 1.0f) * 100.0f, a2f = clipf(f1, -1.0f, 1.0f) * 100.0f;
      hr_rk4(e->x, 0.0, 0.0);
      hr_rk4(e->x + 3, (double)a1f, (double)a2f);
      if (cfg->flags & F_ADD_NOISE)
        for (int c = 0; c < 3; ++c) e->x[c] += (0.0 + e->x[6] * nz[c]) * 0.001;
      double err[3], ne[3];
      for (int c = 0; c < 3; ++c) {
        err[c] = e->x[c] - e->x[3 + c]; ne[c] = err[c] / 50.0;
        obs[c] = (double)(float)ne[c]; obs[3 + c] = (double)(float)clipd(e->x[c] / 20.0, -1.0, 1.0);
      }
      float sq = a[0] * a[0] + a[1] * a[1];
      float pen = (float)0.050 * sq;
      rew = -(fabs(ne[0]) + fabs(ne[1]) + fabs(ne[2])) - (double)pen;
      if (fabs(err[0]) > 70.0 || fabs(err[1]) > 70.0 || fabs(err[2]) > 70.0) { term = 1; rew = -2000.0; }
      break;
    }
    case K_PMSM_SYNC: {  /* lorenz_env_try_pmsm.py:76-184 */
      const int noisy = (cfg->flags & F_ADD_NOISE) != 0;
      double n3[3] = {0, 0, 0};
      if (noisy) for (int c = 0; c < 3; ++c) n3[c] = 0.0 + 3.0 * nz[c];
      float s1[3], s2[3], da[3], db[3];
      for (int c = 0; c < 3; ++c) { s1[c] = (float)e->x[c]; s2[c] = (float)e->x[3 + c]; }
      float lam = (float)e->x[6], m = (float)e->x[7], v = (float)e->x[8];
      float a1 = clipf(a[0], -1.0f, 1.0f) * 50.0f, a2 = clipf(a[1], -1.0f, 1.0f) * 50.0f;
      const float dt = (float)0.001;
      pmsm32_rhs(s1, 0.0f, 0.0f, NULL, 0, da);
      pmsm32_rhs(s2, a1, a2, n3, noisy, db);
      for (int c = 0; c < 3; ++c) { s1[c] = s1[c] + da[c] * dt; s2[c] = s2[c] + db[c] * dt; }
      pmsm32_rhs(s1, 0.0f, 0.0f, NULL, 0, da);
      pmsm32_rhs(s2, a1, a2, n3, noisy, db);
      float ev[3], ef[6];
      for (int c = 0; c < 3; ++c) { ef[c] = s1[c] - s2[c]; ef[3 + c] = da[c] - db[c]; ev[c] = fabsf(ef[c]); }
      float esum = ev[0] + ev[1]; esum = esum + ev[2];
      float grad = 5.0f - esum;
      e->adam += 1;
      m = (float)0.9 * m + (float)(1 - 0.9) * grad;
      v = (float)0.999 * v + (float)(1 - 0.999) * powf(grad, 2.0f);
      float mhat = m / (float)(1 - pow(0.9, (double)e->adam));
      float vhat = v / (float)(1 - pow(0.999, (double)e->adam));
      lam = lam - ((float)0.001 * mhat) / (sqrtf(vhat) + (float)1e-8);
      lam = clipf(lam, 0.0f, 0.5f);
      const float al = (float)cfg->alpha;
      float frac = powf(ev[0] + (float)1e-6, al) + powf(ev[1] + (float)1e-6, al);
      frac = frac + powf(ev[2] + (float)1e-6, al);
      float apen = lam * (powf(a[0], 2.0f) + powf(a[1], 2.0f));
      float r = -esum - frac - apen;
      if (esum > 1000.0f) { r = -1000.0f; term = 1; }
      rew = r;
      for (int c = 0; c < 3; ++c) { e->x[c] = s1[c]; e->x[3 + c] = s2[c]; }
      e->x[6] = lam; e->x[7] = m; e->x[8] = v;
      for (int c = 0; c < 6; ++c) obs[c] = ef[c];
      break;
    }
    case K_PMSM_CLASSIC: {  /* lorenz_env_transient_pmsm.py:76-133 */
      float u1f = clipf(a[0], -2.0f, 2.0f) * 20.0f, u2f = clipf(a[1], -2.0f, 2.0f) * 20.0f;
      double n0 = 0.0 + 3.0 * nz[0], n1 = 0.0 + 3.0 * nz[1], n2 = 0.0 + 3.0 * nz[2];
      double da[3], db[3];
      pmsm64_rhs(e->x, da); pmsm64_rhs(e->x + 3, db);
      db[0] = db[0] + (double)u1f + n0; db[1] = db[1] + (double)u2f + n1; db[2] = db[2] + n2;
      for (int c = 0; c < 3; ++c) { e->x[c] = e->x[c] + da[c] * 0.01; e->x[3 + c] = e->x[3 + c] + db[c] * 0.01; }
      observe(cfg->kind, e, obs);
      double E = 0.0; for (int c = 0; c < 3; ++c) E = E + fabs(obs[c]);
      rew = -E - pow(E, 1.0 / 10);
      e->x[6] = e->x[6] + 0.01;
      term = (e->x[6] == 5.0) || (rew < -1e6);
      break;
    }
    case K_PMSM_SINGLE: {  /* lorenz_env_transient1.py step */
      double u1 = (double)clipf(a[0], -10.0f, 10.0f), u2 = (double)clipf(a[1], -10.0f, 10.0f);
      pmsm64_rhs(e->x, d);
      e->x[0] = e->x[0] + d[0] * 0.01 + u1;
      e->x[1] = e->x[1] + d[1] * 0.01 + u2;
      e->x[2] = e->x[2] + d[2] * 0.01;
      observe(cfg->kind, e, obs);
      rew = 0.0; for (int c = 0; c < 3; ++c) rew = rew + fabs(obs[c]);
      rew = -rew;
      e->x[3] = e->x[3] + 0.01;
      term = (e->x[3] == 10.0);
      break;
    }
    case K_MEMRISTIVE4_PAIR: {  /* lorenz_env_transient2.py step */
      float u1 = clipf(a[0], -2.0f, 2.0f) * 100.0f, u2 = clipf(a[1], -2.0f, 2.0f) * 100.0f,
            u3 = clipf(a[2], -2.0f, 2.0f) * 100.0f;
      memristive4_rhs(e->x, d);
      for (int c = 0; c < 4; ++c) e->x[c] = e->x[c] + d[c] * 0.001;
      memristive4_rhs(e->x + 4, d);
      d[0] = d[0] + (double)u1; d[1] = d[1] + (double)u2; d[3] = d[3] + (double)u3;
      for (int c = 0; c < 4; ++c) e->x[4 + c] = e->x[4 + c] + d[c] * 0.001;
      observe(cfg->kind, e, obs);
      double E = 0.0; for (int c = 0; c < 4; ++c) E = E + fabs(obs[c]);
      rew = -E - pow(E, 1.0 / 3);
      e->x[8] = e->x[8] + 0.001;
      term = (e->x[8] == 5.0) || (rew < -1e6);
      break;
    }
    case K_PMSM_FREE: {  /* lorenz_singlecontrol.py step (takes no action) */
      double n0 = 0.0 + 3.0 * nz[0], n1 = 0.0 + 3.0 * nz[1], n2 = 0.0 + 3.0 * nz[2];
      pmsm64_rhs(e->x, d);
      d[0] = d[0] + n0; d[1] = d[1] + n1; d[2] = d[2] + n2;
      for (int c = 0; c < 3; ++c) e->x[c] = e->x[c] + d[c] * 0.01;
      observe(cfg->kind, e, obs);
      rew = 0.0; for (int c = 0; c < 3; ++c) rew = rew + fabs(obs[c]);
      rew = -rew;
      e->x[3] = e->x[3] + 0.01;
      term = (e->x[3] == 1000.0);
      break;
    }
    case K_LORENZ_RK4: {
      const float lim = (float)cfg->act_limit;
      double u[3];
      for (int c = 0; c < 3; ++c) u[c] = (double)clipf(a[c], -lim, lim) * cfg->act_gain;
      rk4_sub(lorenz_par_rhs, e->x + 3, e->x, u, cfg->dt / (double)cfg->substeps, cfg->substeps);
      observe(cfg->kind, e, obs);
      double E = fabs(e->x[0]) + fabs(e->x[1]) + fabs(e->x[2]);
      rew = -E; term = !(E <= 1e6);
      break;
    }
    case K_LORENZ_RK4_F32: {
      const float lim = (float)cfg->act_limit, g = (float)cfg->act_gain;
      float u[3], s[3], q[3];
      for (int c = 0; c < 3; ++c) { u[c] = clipf(a[c], -lim, lim) * g; s[c] = (float)e->x[c]; q[c] = (float)e->x[3 + c]; }
      rk4_sub_f(q, s, u, (float)(cfg->dt / (double)cfg->substeps), cfg->substeps);
      for (int c = 0; c < 3; ++c) e->x[c] = s[c];
      observe(cfg->kind, e, obs);
      float E = fabsf(s[0]) + fabsf(s[1]) + fabsf(s[2]);
      rew = -E; term = !(E <= 1e6f);
      break;
    }
    case K_PMSM_RK4: {
      const float lim = (float)cfg->act_limit;
      double u[3] = {(double)clipf(a[0], -lim, lim) * cfg->act_gain, (double)clipf(a[1], -lim, lim) * cfg->act_gain, 0.0};
      const double z[3] = {0, 0, 0};
      double h = cfg->dt / (double)cfg->substeps;
      rk4_sub(pmsm_par_rhs, e->x + 6, e->x, z, h, cfg->substeps);
      rk4_sub(pmsm_par_rhs, e->x + 6, e->x + 3, u, h, cfg->substeps);
      observe(cfg->kind, e, obs);
      double e0 = fabs(obs[0]), e1 = fabs(obs[1]), e2 = fabs(obs[2]), E = e0 + e1 + e2;
      rew = -E - (pow(e0 + 1e-6, cfg->alpha) + pow(e1 + 1e-6, cfg->alpha) + pow(e2 + 1e-6, cfg->alpha));
      if (!(E <= 1000.0)) { rew = -1000.0; term = 1; }
      break;
    }
  }
  *rew_out = rew;
  *term_out = term;
}

static void load_env(const orc_cfg* cfg, const void* state, const int32_t* aux, int64_t i, env_t* e) {
  const int ns = kNState[cfg->kind];
  if (is_f32(cfg->kind)) for (int c = 0; c < ns; ++c) e->x[c] = ((const float*)state)[(int64_t)c * cfg->n_pad + i];
  else for (int c = 0; c < ns; ++c) e->x[c] = ((const double*)state)[(int64_t)c * cfg->n_pad + i];
  e->adam = (cfg->kind == K_PMSM_SYNC && aux) ? aux[i] : 0;
}
static void store_env(const orc_cfg* cfg, void* state, int32_t* aux, int64_t i, const env_t* e) {
  const int ns = kNState[cfg->kind];
  if (is_f32(cfg->kind)) for (int c = 0; c < ns; ++c) ((float*)state)[(int64_t)c * cfg->n_pad + i] = (float)e->x[c];
  else for (int c = 0; c < ns; ++c) ((double*)state)[(int64_t)c * cfg->n_pad + i] = e->x[c];
  if (cfg->kind == K_PMSM_SYNC && aux) aux[i] = e->adam;
}

static int time_limit(const orc_cfg* cfg, int32_t n) {
  if (cfg->kind == K_PMSM_SYNC && n >= 2000) return 1;  /* lorenz_env_try_pmsm.py:179-180 */
  return cfg->max_episode_steps > 0 && n >= cfg->max_episode_steps;
}

/* Persistent state (constructor bodies): PMSM_SYNC Adam-dual := 0; north-star params. */
void orc_init_persistent(const orc_cfg* cfg, void* state, int32_t* aux, int32_t* ep_len, double* ep_ret) {
  for (int64_t i = 0; i < cfg->n; ++i) {
    env_t e;
    memset(&e, 0, sizeof(e));
    rng_t rng = make_rng(cfg, i, 0);
    double j[3] = {1, 1, 1};
    if (cfg->kind == K_LORENZ_RK4 || cfg->kind == K_LORENZ_RK4_F32) {
      if (cfg->param_jitter > 0) draw_uniform(&rng, TAG_PARAM, 3, 1.0 - cfg->param_jitter, 1.0 + cfg->param_jitter, j);
      e.x[3] = 10.0 * j[0]; e.x[4] = 28.0 * j[1]; e.x[5] = (8.0 / 3.0) * j[2];
    } else if (cfg->kind == K_PMSM_RK4) {
      if (cfg->param_jitter > 0) draw_uniform(&rng, TAG_PARAM, 2, 1.0 - cfg->param_jitter, 1.0 + cfg->param_jitter, j);
      e.x[6] = 5.46 * j[0]; e.x[7] = 20.0 * j[1];
    }
    store_env(cfg, state, aux, i, &e);
    ep_len[i] = 0; ep_ret[i] = 0.0;
  }
}

/* obs / term_obs are written as double SoA [obs_dim][n_pad] regardless of kind. */
void orc_reset(const orc_cfg* cfg, void* state, int32_t* aux, int32_t* ep_len, double* ep_ret,
               const uint8_t* mask, double* obs) {
  const int no = kObs[cfg->kind];
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < cfg->n; ++i) {
    if (mask && !mask[i]) continue;
    env_t e;
    double o[8];
    load_env(cfg, state, aux, i, &e);
    rng_t rng = make_rng(cfg, i, cfg->step_index);
    env_reset(cfg, &e, &rng, o);
    store_env(cfg, state, aux, i, &e);
    ep_len[i] = 0; ep_ret[i] = 0.0;
    if (obs) for (int c = 0; c < no; ++c) obs[(int64_t)c * cfg->n_pad + i] = o[c];
  }
}

/* T control intervals; action f32 SoA [T][act_dim][n_pad] (or NULL -> Philox synthetic
 * actions of amplitude synth_amp); noise f64 [3][n_pad] standard normals or NULL (Philox);
 * outputs (any may be NULL) time-major: obs f64 [T][obs_dim][n_pad], reward f64 [T][n_pad],
 * done u8 [T][n_pad]; term_obs / last_ep_* as in the C-ABI.  stats: 8 doubles accumulated. */
void orc_rollout(const orc_cfg* cfg, int T, double synth_amp, void* state, int32_t* aux, int32_t* ep_len,
                 double* ep_ret, double* stats, const float* action, const double* noise, double* obs,
                 double* reward, uint8_t* done, double* term_obs, double* last_ep_ret, int32_t* last_ep_len) {
  const int kind = cfg->kind, no = kObs[kind], na = kAct[kind];
  const int64_t np_ = cfg->n_pad;
  double st[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma omp parallel
  {
    double ls[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma omp for schedule(static)
    for (int64_t i = 0; i < cfg->n; ++i) {
      env_t e;
      load_env(cfg, state, aux, i, &e);
      int32_t len = ep_len[i];
      double ret = ep_ret[i];
      for (int t = 0; t < T; ++t) {
        const uint64_t step = cfg->step_index + (uint64_t)t;
        rng_t rng = make_rng(cfg, i, step);
        float a[3] = {0, 0, 0};
        if (action) for (int c = 0; c < na; ++c) a[c] = action[((int64_t)t * na + c) * np_ + i];
        else {
          uint32_t w[4];
          rng_draw(&rng, TAG_ACTION, w);
          for (int c = 0; c < na; ++c) a[c] = (float)synth_amp * (2.0f * ((float)(w[c] >> 8) * (1.0f / 16777216.0f)) - 1.0f);
        }
        double nz[4] = {0, 0, 0, 0};
        int want = (kind == K_PMSM_CLASSIC) || (kind == K_PMSM_FREE) || ((kind == K_HR_SYNC || kind == K_PMSM_SYNC) && (cfg->flags & F_ADD_NOISE));
        if (want) {
          if (noise) for (int c = 0; c < 3; ++c) nz[c] = noise[(int64_t)c * np_ + i];
          else draw_normal(&rng, TAG_NOISE, 3, nz);
        }
        double o[8], rew;
        int term;
        const int was_finite = env_finite(kind, &e);
        env_step(cfg, &e, a, nz, o, &rew, &term);
        len += 1; ret += rew;
        int trunc = time_limit(cfg, len);
        int dn = term || trunc;
        if (was_finite && !env_finite(kind, &e)) ls[4] += 1;  /* divergence event */
        if (dn) {
          if (term_obs) for (int c = 0; c < no; ++c) term_obs[((int64_t)t * no + c) * np_ + i] = o[c];
          if (last_ep_ret) last_ep_ret[i] = ret;
          if (last_ep_len) last_ep_len[i] = len;
          if (cfg->flags & F_AUTORESET) {
            ls[0] += 1; ls[1] += ret; ls[2] += ret * ret; ls[3] += len;
            if (term) ls[5] += 1; else ls[6] += 1;
            env_reset(cfg, &e, &rng, o);
            len = 0; ret = 0.0;
          }
        }
        if (obs) for (int c = 0; c < no; ++c) obs[((int64_t)t * no + c) * np_ + i] = o[c];
        if (reward) reward[(int64_t)t * np_ + i] = rew;
        if (done) done[(int64_t)t * np_ + i] = (uint8_t)((term ? DONE_TERM : 0) | (trunc ? DONE_TRUNC : 0));
      }
      store_env(cfg, state, aux, i, &e);
      ep_len[i] = len; ep_ret[i] = ret;
    }
#pragma omp critical
    for (int k = 0; k < 8; ++k) st[k] += ls[k];
  }
  if (stats) for (int k = 0; k < 8; ++k) stats[k] += st[k];
}

int orc_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
void orc_set_threads(int n) {
#ifdef _OPENMP
  omp_set_num_threads(n);
#else
  (void)n;
#endif
}
