// microbench.cu -- B200 micro-benchmarks that size the assembly design (DESIGN.md "Measured ceilings"):
// FP64 DFMA peak, f64 global reduction (RED) throughput under three address patterns, shared-memory f64
// atomics, plain HBM copy.  Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o microbench microbench.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <stdint.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

__global__ void k_dfma(double* out, int iters, double a, double b) {
  double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
  for (int i = 0; i < iters; ++i) {
    x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
    x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

// DFMA throughput as a function of instruction-level parallelism (independent chains per thread) and warps per SM
template <int ILP>
__global__ void k_dfma_ilp(double* out, int iters, double a, double b) {
  double x[ILP];
#pragma unroll
  for (int k = 0; k < ILP; ++k) x[k] = threadIdx.x + k;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < ILP; ++k) x[k] = fma(x[k], a, b);
  }
  double s = 0;
#pragma unroll
  for (int k = 0; k < ILP; ++k) s += x[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int ILP>
static void run_ilp(double* out, int warps_per_sm, cudaEvent_t e0, cudaEvent_t e1, int sms, double clk_hz) {
  const int iters = 20000;
  k_dfma_ilp<ILP><<<sms, 32 * warps_per_sm>>>(out, 10, 1.0000001, 1e-9);
  cudaEventRecord(e0); k_dfma_ilp<ILP><<<sms, 32 * warps_per_sm>>>(out, iters, 1.0000001, 1e-9); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double fma = (double)ILP * iters * sms * 32.0 * warps_per_sm;
  printf("  warps/SM %2d ILP %d: %.1f DFMA/clk/SM\n", warps_per_sm, ILP, fma / (ms * 1e-3) / sms / clk_hz);
}

__device__ __forceinline__ uint64_t mix(uint64_t z) {
  z ^= z >> 33; z *= 0xff51afd7ed558ccdULL; z ^= z >> 33; z *= 0xc4ceb9fe1a85ec53ULL; z ^= z >> 33; return z;
}

// mode 0: every lane a random address; 1: each warp 32 consecutive doubles at a random base;
// 2: groups of 4 lanes share a random 32 B sector; 3: mode 0 but plain (non-atomic) RMW
__global__ void k_red(double* buf, uint64_t n, int reps, int mode) {
  const uint64_t t = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  for (int r = 0; r < reps; ++r) {
    uint64_t idx;
    if (mode == 0 || mode == 3) idx = mix(t * 1315423911ULL + r) % n;
    else if (mode == 1) idx = (mix((t >> 5) * 2654435761ULL + r) % (n / 32)) * 32 + (t & 31);
    else idx = (mix((t >> 2) * 2654435761ULL + r) % (n / 4)) * 4 + (t & 3);
    if (mode == 3) buf[idx] += 1.0; else atomicAdd(buf + idx, 1.0);
  }
}

__global__ void k_smem_atomic(double* out, int reps) {
  __shared__ double acc[4096];
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) acc[i] = 0;
  __syncthreads();
  for (int r = 0; r < reps; ++r) {
    const unsigned idx = (unsigned)mix(threadIdx.x * 7919ULL + r * 104729ULL + blockIdx.x) & 4095u;
    atomicAdd(acc + idx, 1.0);
  }
  __syncthreads();
  if (threadIdx.x == 0) out[blockIdx.x] = acc[1];
}

__global__ void k_smem_rmw(double* out, int reps) {
  __shared__ double acc[4096];
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) acc[i] = 0;
  __syncthreads();
  for (int r = 0; r < reps; ++r) {
    const unsigned idx = (threadIdx.x * 17u + r * 1031u) & 4095u;   // conflict-free per warp
    acc[idx] += 1.0;
  }
  __syncthreads();
  if (threadIdx.x == 0) out[blockIdx.x] = acc[1];
}

__global__ void k_copy(const double2* __restrict__ a, double2* __restrict__ b, uint64_t n) {
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) b[i] = a[i];
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  printf("device %s, %d SMs, clock %.0f MHz\n", p.name, p.multiProcessorCount, p.clockRate / 1e3);
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  float ms;
  double* out; CK(cudaMalloc(&out, 8ull * 148 * 8 * 1024));
  {
    const int iters = 20000, blocks = 148 * 4, threads = 512;
    k_dfma<<<blocks, threads>>>(out, 100, 1.0000001, 1e-9);
    CK(cudaEventRecord(e0)); k_dfma<<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    CK(cudaEventElapsedTime(&ms, e0, e1));
    const double fl = 2.0 * 8 * iters * (double)blocks * threads;
    printf("DFMA: %.2f TFLOP/s  (%.1f DFMA/clk/SM at %.0f MHz nominal)\n", fl / ms / 1e9, fl / 2 / (ms * 1e-3) / p.multiProcessorCount / (p.clockRate * 1e3), p.clockRate / 1e3);
  }
  {
    printf("DFMA issue vs occupancy / ILP (peak = 64 per clk per SM):\n");
    for (int w : {4, 8, 16, 32}) {
      run_ilp<1>(out, w, e0, e1, p.multiProcessorCount, p.clockRate * 1e3);
      run_ilp<2>(out, w, e0, e1, p.multiProcessorCount, p.clockRate * 1e3);
      run_ilp<4>(out, w, e0, e1, p.multiProcessorCount, p.clockRate * 1e3);
    }
  }
  {
    const uint64_t n = 1ull << 28;  // 2 GiB of doubles
    double* buf; CK(cudaMalloc(&buf, n * 8)); CK(cudaMemset(buf, 0, n * 8));
    const char* names[] = {"scattered lanes", "warp-contiguous 256B", "4-lane sector groups", "plain RMW scattered (racy)"};
    for (int mode = 0; mode < 4; ++mode) {
      const int reps = 64, blocks = 148 * 16, threads = 256;
      k_red<<<blocks, threads>>>(buf, n, 2, mode);
      CK(cudaEventRecord(e0)); k_red<<<blocks, threads>>>(buf, n, reps, mode); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
      CK(cudaEventElapsedTime(&ms, e0, e1));
      const double ops = (double)reps * blocks * threads;
      printf("RED.f64 %-28s: %.1f Gop/s (%.0f GB/s of payload)\n", names[mode], ops / ms / 1e6, ops * 8 / ms / 1e6);
    }
    // L2-resident target (64 MB)
    for (int mode = 0; mode < 3; ++mode) {
      const int reps = 64, blocks = 148 * 16, threads = 256;
      CK(cudaEventRecord(e0)); k_red<<<blocks, threads>>>(buf, 1ull << 23, reps, mode); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
      CK(cudaEventElapsedTime(&ms, e0, e1));
      const double ops = (double)reps * blocks * threads;
      printf("RED.f64 L2-resident %-16s: %.1f Gop/s\n", names[mode], ops / ms / 1e6);
    }
    CK(cudaFree(buf));
  }
  {
    const int reps = 4096, blocks = 148 * 4, threads = 512;
    CK(cudaEventRecord(e0)); k_smem_atomic<<<blocks, threads>>>(out, reps); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    CK(cudaEventElapsedTime(&ms, e0, e1));
    printf("smem atomicAdd(double) random: %.1f Gop/s chip (%.2f op/clk/SM)\n", (double)reps * blocks * threads / ms / 1e6,
           (double)reps * blocks * threads / (ms * 1e-3) / p.multiProcessorCount / (p.clockRate * 1e3));
    CK(cudaEventRecord(e0)); k_smem_rmw<<<blocks, threads>>>(out, reps); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    CK(cudaEventElapsedTime(&ms, e0, e1));
    printf("smem plain RMW double conflict-free: %.1f Gop/s chip (%.2f op/clk/SM)\n", (double)reps * blocks * threads / ms / 1e6,
           (double)reps * blocks * threads / (ms * 1e-3) / p.multiProcessorCount / (p.clockRate * 1e3));
  }
  {
    const uint64_t n = 1ull << 27;  // 2 GiB each
    double2 *a, *b; CK(cudaMalloc(&a, n * 16)); CK(cudaMalloc(&b, n * 16)); CK(cudaMemset(a, 1, n * 16));
    k_copy<<<148 * 8, 512>>>(a, b, n);
    CK(cudaEventRecord(e0)); for (int i = 0; i < 5; ++i) k_copy<<<148 * 8, 512>>>(a, b, n); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    CK(cudaEventElapsedTime(&ms, e0, e1));
    printf("HBM copy: %.0f GB/s (read+write)\n", 5.0 * 2 * n * 16 / ms / 1e6);
  }
  return 0;
}
