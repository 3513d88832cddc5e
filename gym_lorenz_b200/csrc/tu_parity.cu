// Translation unit for the PARITY env kinds.  Built with -fmad=false (see build.py): no
// implicit mul+add contraction, so every expression rounds exactly as the reference's
// NumPy scalar arithmetic does.
#include "envs_parity.cuh"

using namespace cl;

#define CL_PARITY_KINDS(X)                    \
  X(CL_ENV_LORENZ3, EnvLorenz3)               \
  X(CL_ENV_LORENZ3_PAIR, EnvLorenz3Pair)      \
  X(CL_ENV_LORENZ4_PAIR, EnvLorenz4Pair)      \
  X(CL_ENV_HR_SYNC, EnvHRSync)                \
  X(CL_ENV_PMSM_SYNC, EnvPMSMSync)            \
  X(CL_ENV_PMSM_CLASSIC, EnvPMSMClassic)      \
  X(CL_ENV_PMSM_SINGLE, EnvPMSMSingle)        \
  X(CL_ENV_MEMRISTIVE4_PAIR, EnvMemristive4Pair) \
  X(CL_ENV_PMSM_FREE, EnvPMSMFree)

cudaError_t cl_launch_parity(int kind, const KParams& p, int mode, cudaStream_t st, int block) {
  switch (kind) {
#define X(K, E) case K: return launch_env<E>(p, mode, st, block);
    CL_PARITY_KINDS(X)
#undef X
    default: return cudaErrorInvalidValue;
  }
}

cudaError_t cl_occupancy_parity(int kind, int block, int* out) {
  switch (kind) {
#define X(K, E) case K: return cudaOccupancyMaxActiveBlocksPerMultiprocessor(out, k_step<E, false>, block, 0);
    CL_PARITY_KINDS(X)
#undef X
    default: return cudaErrorInvalidValue;
  }
}

cudaError_t cl_dyn_occupancy_parity(int kind, int chunk, int* out) {
  switch (kind) {
#define X(K, E) case K: return dyn_occupancy<E>(chunk, out);
    CL_PARITY_KINDS(X)
#undef X
    default: return cudaErrorInvalidValue;
  }
}

// ---- derivative helpers (cl_derivatives) ---------------------------------------------
// PMSM_Sync_Env._get_derivatives (lorenz_env_try_pmsm.py:51-58), called directly by
// code/lorenz_pmsm/test_evaluate.py:105-108; hr_derivatives (lorenz_env_try.py:7-12).
__global__ void k_deriv_pmsm32(const float* s, const float* a, float* out, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float x[3] = {s[i], s[n + i], s[2 * n + i]};
  float d[3];
  EnvPMSMSync::rhs(x, a ? a[i] : 0.0f, a ? a[n + i] : 0.0f, nullptr, false, d);
  out[i] = d[0]; out[n + i] = d[1]; out[2 * n + i] = d[2];
}
__global__ void k_deriv_hr(const double* s, const float* a, double* out, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double x[3] = {s[i], s[n + i], s[2 * n + i]};
  double d[3];
  hr_rhs(x, a ? (double)a[i] : 0.0, a ? (double)a[n + i] : 0.0, d);
  out[i] = d[0]; out[n + i] = d[1]; out[2 * n + i] = d[2];
}
__global__ void k_deriv_lorenz3(const double* s, double* out, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double dx, dy, dz;
  lorenz3_rhs(s[i], s[n + i], s[2 * n + i], dx, dy, dz);
  out[i] = dx; out[n + i] = dy; out[2 * n + i] = dz;
}
__global__ void k_deriv_pmsm64(const double* s, double* out, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double dx, dy, dz;
  pmsm64_rhs(s[i], s[n + i], s[2 * n + i], dx, dy, dz);
  out[i] = dx; out[n + i] = dy; out[2 * n + i] = dz;
}
__global__ void k_deriv_lorenz4(const double* s, double* out, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double x[4] = {s[i], s[n + i], s[2 * n + i], s[3 * n + i]};
  double d[4];
  lorenz4_rhs(x, d);
  for (int c = 0; c < 4; ++c) out[c * n + i] = d[c];
}

cudaError_t cl_launch_derivatives(int kind, const void* state, const float* action, void* out,
                                  int64_t n, cudaStream_t st) {
  const int block = 128;
  const unsigned grid = (unsigned)((n + block - 1) / block);
  switch (kind) {
    case CL_ENV_PMSM_SYNC: k_deriv_pmsm32<<<grid, block, 0, st>>>((const float*)state, action, (float*)out, n); break;
    case CL_ENV_HR_SYNC: k_deriv_hr<<<grid, block, 0, st>>>((const double*)state, action, (double*)out, n); break;
    case CL_ENV_LORENZ3:
    case CL_ENV_LORENZ3_PAIR: k_deriv_lorenz3<<<grid, block, 0, st>>>((const double*)state, (double*)out, n); break;
    case CL_ENV_PMSM_CLASSIC:
    case CL_ENV_PMSM_FREE:
    case CL_ENV_PMSM_SINGLE: k_deriv_pmsm64<<<grid, block, 0, st>>>((const double*)state, (double*)out, n); break;
    case CL_ENV_LORENZ4_PAIR: k_deriv_lorenz4<<<grid, block, 0, st>>>((const double*)state, (double*)out, n); break;
    default: return cudaErrorInvalidValue;
  }
  return cudaGetLastError();
}
