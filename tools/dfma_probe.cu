// dfma_probe.cu -- micro-experiments behind DESIGN.md's FP64 cost model (run on the GPU box):
//   1. DFMA/DADD issue cost vs number of distinct register operands
//   2. Lorenz RK4 substep loop variants (parameter / step-size operands in registers vs
//      constant bank, block size, envs per thread) at N = 65,536 envs
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o build/dfma_probe tools/dfma_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

template <int MODE>
__global__ void __launch_bounds__(1024) k_op(int iters, const double* in, double* out) {
  double a[8], b[8], c[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { a[j] = in[threadIdx.x + j]; b[j] = in[64 + threadIdx.x + j]; c[j] = in[128 + threadIdx.x + j]; }
  const double bs = b[0], cs = c[0];
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (MODE == 0) a[j] = fma(a[j], b[j], c[j]);          // 3 distinct regs
        if (MODE == 1) a[j] = fma(a[j], b[j], 1e-9);          // 2 regs + imm/const
        if (MODE == 2) a[j] = fma(a[j], 0.9999999, 1e-9);     // 1 reg + 2 const
        if (MODE == 3) a[j] = a[j] + b[j];                    // DADD 2 regs
        if (MODE == 4) a[j] = fma(a[j], bs, c[j]);            // shared multiplier
        if (MODE == 5) a[j] = fma(a[j], bs, cs);              // shared multiplier and addend
        if (MODE == 6) a[j] = fma(b[j], c[j], a[j]);          // accumulate form
        if (MODE == 7) a[j] = a[j] * b[j];                    // DMUL 2 regs
      }
    }
  }
  double s = 0;
#pragma unroll
  for (int j = 0; j < 8; ++j) s += a[j];
  if (s == -1.2345) out[0] = s;
}

// half-warp test: only `active` lanes of each warp execute the chain
__global__ void __launch_bounds__(1024) k_half(int iters, int active, const double* in, double* out) {
  if ((threadIdx.x & 31) >= active) return;
  double a[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) a[j] = in[threadIdx.x + j];
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
#pragma unroll
      for (int j = 0; j < 8; ++j) a[j] = fma(a[j], 0.9999999, 1e-9);
    }
  }
  double s = 0;
#pragma unroll
  for (int j = 0; j < 8; ++j) s += a[j];
  if (s == -1.2345) out[0] = s;
}

// conversion throughput (11 F2F per control interval in the env kernel)
template <int MODE>
__global__ void __launch_bounds__(1024) k_cvt(int iters, const float* in, double* out) {
  float f[8]; double d[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { f[j] = in[threadIdx.x + j]; d[j] = (double)f[j] * 1.000001; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (MODE == 0) { d[j] = (double)f[j]; f[j] = __int_as_float(__float_as_int(f[j]) ^ (int)__double2loint(d[j])); }   // F2F.F64.F32 (+ cheap int ops)
        if (MODE == 1) { f[j] = (float)d[j]; d[j] = __hiloint2double(__double2hiint(d[j]), __float_as_int(f[j])); }     // F2F.F32.F64
      }
    }
  }
  double s = 0;
#pragma unroll
  for (int j = 0; j < 8; ++j) s += d[j] + (double)f[j];
  if (s == -1.2345) out[0] = s;
}

// ---------------- RK4 variants ------------------------------------------------------------
struct Par { double sigma, rho, beta; };
__device__ __forceinline__ void rhs(const Par& q, double x, double y, double z, double u1, double u2, double u3,
                                    double& dx, double& dy, double& dz) {
  dx = fma(q.sigma, y - x, u1);
  dy = fma(x, q.rho - z, u2 - y);
  dz = fma(x, y, fma(-q.beta, z, u3));
}
__device__ __forceinline__ void rk4(const Par& q, double& x, double& y, double& z, double u1, double u2, double u3,
                                    double h, double hh, double h3, double h6, int S) {
#pragma unroll 2
  for (int k = 0; k < S; ++k) {
    double k1x, k1y, k1z, kx, ky, kz, ax, ay, az;
    rhs(q, x, y, z, u1, u2, u3, k1x, k1y, k1z);
    ax = fma(h6, k1x, x); ay = fma(h6, k1y, y); az = fma(h6, k1z, z);
    rhs(q, fma(hh, k1x, x), fma(hh, k1y, y), fma(hh, k1z, z), u1, u2, u3, kx, ky, kz);
    ax = fma(h3, kx, ax); ay = fma(h3, ky, ay); az = fma(h3, kz, az);
    rhs(q, fma(hh, kx, x), fma(hh, ky, y), fma(hh, kz, z), u1, u2, u3, k1x, k1y, k1z);
    ax = fma(h3, k1x, ax); ay = fma(h3, k1y, ay); az = fma(h3, k1z, az);
    rhs(q, fma(h, k1x, x), fma(h, k1y, y), fma(h, k1z, z), u1, u2, u3, kx, ky, kz);
    x = fma(h6, kx, ax); y = fma(h6, ky, ay); z = fma(h6, kz, az);
  }
}

struct Args { double dt; int S; double h, hh, h3, h6; double sigma, rho, beta; };

// VAR 0: params in registers (loaded per env), h computed in-kernel (current product kernel)
// VAR 1: params in registers, h* from the constant bank
// VAR 2: params AND h* from the constant bank (uniform parameters)
template <int VAR, int EPT>
__global__ void __launch_bounds__(256) k_rk4(const Args a, int T, int n, const double* st, double* out) {
  const int i0 = (blockIdx.x * blockDim.x + threadIdx.x) * EPT;
  if (i0 >= n) return;
  double x[EPT], y[EPT], z[EPT];
  Par q[EPT];
#pragma unroll
  for (int e = 0; e < EPT; ++e) {
    x[e] = st[i0 + e]; y[e] = st[n + i0 + e]; z[e] = st[2 * n + i0 + e];
    if (VAR == 2) { q[e].sigma = a.sigma; q[e].rho = a.rho; q[e].beta = a.beta; }
    else { q[e].sigma = st[3 * n + i0 + e]; q[e].rho = st[4 * n + i0 + e]; q[e].beta = st[5 * n + i0 + e]; }
  }
  double h, hh, h3, h6;
  if (VAR == 0) { h = a.dt / (double)a.S; hh = 0.5 * h; h3 = h / 3.0; h6 = h / 6.0; }
  else { h = a.h; hh = a.hh; h3 = a.h3; h6 = a.h6; }
  for (int t = 0; t < T; ++t) {
#pragma unroll
    for (int e = 0; e < EPT; ++e) {
      const double u1 = 1e-3 * (double)(t & 7), u2 = -u1, u3 = 0.5 * u1;
      rk4(q[e], x[e], y[e], z[e], u1, u2, u3, h, hh, h3, h6, a.S);
    }
  }
#pragma unroll
  for (int e = 0; e < EPT; ++e) { out[i0 + e] = x[e]; out[n + i0 + e] = y[e]; out[2 * n + i0 + e] = z[e]; }
}

// VAR 2 with the two envs of a thread interleaved substep by substep (explicit ILP)
template <int VAR>
__global__ void __launch_bounds__(256) k_rk4_il2(const Args a, int T, int n, const double* st, double* out) {
  const int i0 = (blockIdx.x * blockDim.x + threadIdx.x);
  const int half = n / 2;
  if (i0 >= half) return;
  double x0 = st[i0], y0 = st[n + i0], z0 = st[2 * n + i0];
  double x1 = st[half + i0], y1 = st[n + half + i0], z1 = st[2 * n + half + i0];
  Par q0, q1;
  if (VAR == 2) { q0.sigma = a.sigma; q0.rho = a.rho; q0.beta = a.beta; q1 = q0; }
  else { q0.sigma = st[3 * n + i0]; q0.rho = st[4 * n + i0]; q0.beta = st[5 * n + i0];
         q1.sigma = st[3 * n + half + i0]; q1.rho = st[4 * n + half + i0]; q1.beta = st[5 * n + half + i0]; }
  const double h = a.h, hh = a.hh, h3 = a.h3, h6 = a.h6;
  for (int t = 0; t < T; ++t) {
    const double u1 = 1e-3 * (double)(t & 7), u2 = -u1, u3 = 0.5 * u1;
    for (int k = 0; k < a.S; ++k) {
      rk4(q0, x0, y0, z0, u1, u2, u3, h, hh, h3, h6, 1);
      rk4(q1, x1, y1, z1, u1, u2, u3, h, hh, h3, h6, 1);
    }
  }
  out[i0] = x0 + x1; out[n + i0] = y0 + y1; out[2 * n + i0] = z0 + z1;
}

// persistent-style balanced launch: exactly `wps` warps per scheduler, each thread integrates
// `reps` envs one after another (compute only) -- what the RK4 loop can reach at 3 vs 4 warps
template <int VAR>
__global__ void __launch_bounds__(128) k_rk4_bal(const Args a, int T, int reps, int n, const double* st, double* out) {
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  Par q; q.sigma = a.sigma; q.rho = a.rho; q.beta = a.beta;
  double acc = 0;
  for (int r = 0; r < reps; ++r) {
    const int i = (tid + r * 7919) % n;
    double x = st[i], y = st[n + i], z = st[2 * n + i];
    for (int t = 0; t < T; ++t) {
      const double u1 = 1e-3 * (double)(t & 7), u2 = -u1, u3 = 0.5 * u1;
      rk4(q, x, y, z, u1, u2, u3, a.h, a.hh, a.h3, a.h6, a.S);
    }
    acc += x + y + z;
  }
  out[tid % n] = acc;
}

template <typename F>
float time_ms(F launch, int reps) {
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  launch(); CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < reps; ++r) {
    CK(cudaEventRecord(e0)); launch(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
  }
  return best;
}

int main() {
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  const int sms = prop.multiProcessorCount;
  printf("device %s, %d SMs, clock %d kHz\n", prop.name, sms, prop.clockRate);
  double* in; double* out;
  std::vector<double> h(6 * 65536);
  for (size_t i = 0; i < h.size(); ++i) h[i] = 0.5 + 1e-3 * (double)(i % 97);
  for (int i = 0; i < 65536; ++i) { h[3 * 65536 + i] = 10.0; h[4 * 65536 + i] = 28.0; h[5 * 65536 + i] = 8.0 / 3.0; }
  CK(cudaMalloc(&in, h.size() * 8)); CK(cudaMalloc(&out, h.size() * 8));
  CK(cudaMemcpy(in, h.data(), h.size() * 8, cudaMemcpyHostToDevice));

  printf("\n== 1. FP64 op cost vs operand kinds (per-SMSP cycles per warp-instruction at 1965 MHz nominal) ==\n");
  const char* names[8] = {"DFMA r,r,r", "DFMA r,r,imm", "DFMA r,c,c", "DADD r,r", "DFMA r,rs,r (shared mul)",
                          "DFMA r,rs,rs", "DFMA b,c,a (acc)", "DMUL r,r"};
  const int iters = 4000;
  for (int w = 1; w <= 8; w *= 2) {
    for (int m = 0; m < 8; ++m) {
      const int block = 128 * w;   // w warps per SMSP (one block per SM)
      auto L = [&]() {
        switch (m) {
          case 0: k_op<0><<<sms, block>>>(iters, in, out); break;
          case 1: k_op<1><<<sms, block>>>(iters, in, out); break;
          case 2: k_op<2><<<sms, block>>>(iters, in, out); break;
          case 3: k_op<3><<<sms, block>>>(iters, in, out); break;
          case 4: k_op<4><<<sms, block>>>(iters, in, out); break;
          case 5: k_op<5><<<sms, block>>>(iters, in, out); break;
          case 6: k_op<6><<<sms, block>>>(iters, in, out); break;
          case 7: k_op<7><<<sms, block>>>(iters, in, out); break;
        }
      };
      float ms = time_ms(L, 5);
      double winstr_per_smsp = (double)iters * 64.0 * w;      // warp-instructions issued per SMSP
      double cyc = ms * 1e-3 * 1.965e9 / winstr_per_smsp;
      double tf = 2.0 * iters * 64.0 * 32.0 * 4 * w * sms / (ms * 1e-3) * 1e-12;
      printf("warps/SMSP=%d  %-26s %8.3f ms  %.2f cyc/warp-instr  %.2f TFLOP/s-equiv\n", w, names[m], ms, cyc, tf);
    }
  }

  printf("\n== 1b. DFMA r,c,c with only the first `active` lanes of each warp executing (4 warps/SMSP) ==\n");
  for (int active : {32, 24, 16, 8, 1}) {
    auto L = [&]() { k_half<<<sms, 512>>>(iters, active, in, out); };
    float ms = time_ms(L, 5);
    printf("active lanes=%2d  %8.3f ms  %.2f cyc/warp-instr\n", active, ms, ms * 1e-3 * 1.965e9 / ((double)iters * 64.0 * 4));
  }
  printf("\n== 1c. F2F conversion cost (4 warps/SMSP; each iteration = 32 conversions + 32 int ops per thread) ==\n");
  {
    float* fin; CK(cudaMalloc(&fin, 4096 * 4)); CK(cudaMemset(fin, 0x3c, 4096 * 4));
    for (int m = 0; m < 2; ++m) {
      auto L = [&]() { if (m == 0) k_cvt<0><<<sms, 512>>>(2000, fin, out); else k_cvt<1><<<sms, 512>>>(2000, fin, out); };
      float ms = time_ms(L, 5);
      printf("%s  %8.3f ms  %.2f cyc per warp-conversion per SMSP (upper bound, includes the int op)\n",
             m == 0 ? "F2F.F64.F32" : "F2F.F32.F64", ms, ms * 1e-3 * 1.965e9 / (2000.0 * 32.0 * 4));
    }
  }
  printf("\n== 1d. pure RK4 loop (uniform params, const operands), balanced persistent launch: warps/SMSP vs TFLOP/s(87) ==\n");
  {
    Args a; a.dt = 0.01; a.S = 16; a.h = a.dt / 16; a.hh = 0.5 * a.h; a.h3 = a.h / 3; a.h6 = a.h / 6;
    a.sigma = 10; a.rho = 28; a.beta = 8.0 / 3.0;
    for (int wps = 1; wps <= 6; ++wps) {
      const int grid = sms * wps, T = 16, reps = 4;
      auto L = [&]() { k_rk4_bal<2><<<grid, 128>>>(a, T, reps, 65536, in, out); };
      float ms = time_ms(L, 5);
      const double sub = (double)grid * 128 * reps * T * a.S;
      printf("warps/SMSP=%d  %8.3f ms  %.2f TFLOP/s(87)  = %.1f%% of 36.6\n", wps, ms, sub * 87 / (ms * 1e-3) * 1e-12,
             sub * 87 / (ms * 1e-3) * 1e-12 / 36.6 * 100);
    }
  }
  if (getenv("PROBE_SHORT")) return 0;
  printf("\n== 2. Lorenz RK4 loop variants, N=65536, S=16, T=64 (4096 substeps/env... x) ==\n");
  Args a; a.dt = 0.01; a.S = 16; a.h = a.dt / 16; a.hh = 0.5 * a.h; a.h3 = a.h / 3; a.h6 = a.h / 6;
  a.sigma = 10; a.rho = 28; a.beta = 8.0 / 3.0;
  const int n = 65536, T = 64;
  const double substeps = (double)n * T * a.S;
  for (int block : {32, 64, 128, 256}) {
    for (int var = 0; var < 3; ++var) {
      for (int ept = 1; ept <= 2; ++ept) {
        const int grid = (n / ept + block - 1) / block;
        auto L = [&]() {
          if (ept == 1) { if (var == 0) k_rk4<0, 1><<<grid, block>>>(a, T, n, in, out); if (var == 1) k_rk4<1, 1><<<grid, block>>>(a, T, n, in, out); if (var == 2) k_rk4<2, 1><<<grid, block>>>(a, T, n, in, out); }
          else { if (var == 0) k_rk4<0, 2><<<grid, block>>>(a, T, n, in, out); if (var == 1) k_rk4<1, 2><<<grid, block>>>(a, T, n, in, out); if (var == 2) k_rk4<2, 2><<<grid, block>>>(a, T, n, in, out); }
        };
        float ms = time_ms(L, 5);
        printf("block=%3d var=%d ept=%d grid=%5d  %8.3f ms  %.3e substeps/s  %.2f TFLOP/s(87)\n", block, var, ept, grid, ms,
               substeps / (ms * 1e-3), substeps * 87 / (ms * 1e-3) * 1e-12);
      }
    }
    for (int var = 1; var < 3; ++var) {
      const int grid = (n / 2 + block - 1) / block;
      auto L = [&]() { if (var == 1) k_rk4_il2<1><<<grid, block>>>(a, T, n, in, out); else k_rk4_il2<2><<<grid, block>>>(a, T, n, in, out); };
      float ms = time_ms(L, 5);
      printf("block=%3d var=%d interleave2 grid=%5d  %8.3f ms  %.3e substeps/s  %.2f TFLOP/s(87)\n", block, var, grid, ms,
             substeps / (ms * 1e-3), substeps * 87 / (ms * 1e-3) * 1e-12);
    }
  }
  printf("\n== 3. same, N=1048576 (cfg 4 size), T=8 ==\n");
  {
    const int n2 = 1048576, T2 = 8;
    double* in2; double* out2;
    std::vector<double> h2(6 * (size_t)n2);
    for (size_t i = 0; i < h2.size(); ++i) h2[i] = 0.5 + 1e-3 * (double)(i % 97);
    for (int i = 0; i < n2; ++i) { h2[3 * (size_t)n2 + i] = 10.0; h2[4 * (size_t)n2 + i] = 28.0; h2[5 * (size_t)n2 + i] = 8.0 / 3.0; }
    CK(cudaMalloc(&in2, h2.size() * 8)); CK(cudaMalloc(&out2, h2.size() * 8));
    CK(cudaMemcpy(in2, h2.data(), h2.size() * 8, cudaMemcpyHostToDevice));
    const double sub2 = (double)n2 * T2 * a.S;
    for (int block : {64, 128, 256}) {
      for (int var = 0; var < 3; ++var) {
        const int grid = (n2 + block - 1) / block;
        auto L = [&]() { if (var == 0) k_rk4<0, 1><<<grid, block>>>(a, T2, n2, in2, out2); if (var == 1) k_rk4<1, 1><<<grid, block>>>(a, T2, n2, in2, out2); if (var == 2) k_rk4<2, 1><<<grid, block>>>(a, T2, n2, in2, out2); };
        float ms = time_ms(L, 3);
        printf("block=%3d var=%d ept=1 grid=%6d  %8.3f ms  %.3e substeps/s  %.2f TFLOP/s(87)\n", block, var, grid, ms,
               sub2 / (ms * 1e-3), sub2 * 87 / (ms * 1e-3) * 1e-12);
      }
    }
  }
  return 0;
}
