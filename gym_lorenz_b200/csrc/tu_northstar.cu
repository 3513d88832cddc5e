// Translation unit for the NORTH-STAR env kinds (RK4 x S, FMA contraction on) and the
// FMA-chain micro-kernel that measures the FP64 / FP32 roofline denominators.
#include "envs_northstar.cuh"

using namespace cl;

#define CL_NS_KINDS(X)                              \
  X(CL_ENV_LORENZ_RK4, EnvLorenzRK4<double>)        \
  X(CL_ENV_LORENZ_RK4_F32, EnvLorenzRK4<float>)     \
  X(CL_ENV_PMSM_RK4, EnvPMSMRK4)

cudaError_t cl_launch_northstar(int kind, const KParams& p, int mode, cudaStream_t st, int block) {
  switch (kind) {
#define X(K, E) case K: return launch_env<E>(p, mode, st, block);
    CL_NS_KINDS(X)
#undef X
    default: return cudaErrorInvalidValue;
  }
}

cudaError_t cl_occupancy_northstar(int kind, int block, int* out) {
  switch (kind) {
#define X(K, E) case K: return cudaOccupancyMaxActiveBlocksPerMultiprocessor(out, k_step<E, true>, block, 0);
    CL_NS_KINDS(X)
#undef X
    default: return cudaErrorInvalidValue;
  }
}

cudaError_t cl_dyn_occupancy_northstar(int kind, int chunk, int* out) {
  switch (kind) {
#define X(K, E) case K: return dyn_occupancy<E>(chunk, out);
    CL_NS_KINDS(X)
#undef X
    default: return cudaErrorInvalidValue;
  }
}

// ---- FMA peak: 8 independent register-resident chains per thread ---------------------
template <typename R>
__global__ void __launch_bounds__(256) k_fma_peak(int iters, R* sink) {
  R a[8];
  const R m = R(1.0) - R(1e-7), c = R(1e-9);
#pragma unroll
  for (int j = 0; j < 8; ++j) a[j] = R(threadIdx.x + j) * R(1e-3);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 16; ++r) {
#pragma unroll
      for (int j = 0; j < 8; ++j) a[j] = fma(a[j], m, c);
    }
  }
  R s = R(0);
#pragma unroll
  for (int j = 0; j < 8; ++j) s += a[j];
  if (s == R(-1.2345)) sink[0] = s;  // never true; keeps the chains alive
}

cudaError_t cl_fma_peak_launch(int dtype_bytes, int grid, int block, int iters, void* sink, cudaStream_t st) {
  if (dtype_bytes == 8) k_fma_peak<double><<<grid, block, 0, st>>>(iters, (double*)sink);
  else k_fma_peak<float><<<grid, block, 0, st>>>(iters, (float*)sink);
  return cudaGetLastError();
}
