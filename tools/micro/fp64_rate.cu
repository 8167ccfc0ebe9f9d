// Micro-benchmark: FP64 vector (DFMA) vs FP64 tensor (DMMA m8n8k4) issue rate and dependent
// latency on one SM.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_rate.bin fp64_rate.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void dfma_tput(double* out, long long* clk, int iters) {
  double a[8];
  for (int i = 0; i < 8; ++i) a[i] = threadIdx.x + i;
  const double x = 1.0000001, y = 1e-9;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = fma(a[i], x, y);
  }
  __syncthreads();
  long long t1 = clock64();
  double s = 0;
  for (int i = 0; i < 8; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}

__global__ void dfma_lat(double* out, long long* clk, int iters) {
  double a = threadIdx.x;
  const double x = 1.0000001, y = 1e-9;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) a = fma(a, x, y);
  }
  long long t1 = clock64();
  out[threadIdx.x] = a;
  if (threadIdx.x == 0) clk[0] = t1 - t0;
}

__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

__global__ void dmma_tput(double* out, long long* clk, int iters) {
  double c[8][2];
  for (int i = 0; i < 8; ++i) c[i][0] = c[i][1] = 0.0;
  const double a = 1.0 + threadIdx.x * 1e-3, b = 1.0 - threadIdx.x * 1e-3;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) dmma(c[i][0], c[i][1], a, b);
  }
  __syncthreads();
  long long t1 = clock64();
  double s = 0;
  for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}

__global__ void dmma_lat(double* out, long long* clk, int iters) {
  double c0 = 0, c1 = 0;
  const double a = 1.0 + threadIdx.x * 1e-3, b = 1.0 - threadIdx.x * 1e-3;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) dmma(c0, c1, a, b);
  }
  long long t1 = clock64();
  out[threadIdx.x] = c0 + c1;
  if (threadIdx.x == 0) clk[0] = t1 - t0;
}

int main() {
  double* out;
  long long* clk;
  cudaMalloc(&out, 1 << 20);
  cudaMalloc(&clk, 1024);
  long long h;
  const int iters = 2000;
  for (int threads : {32, 128, 256, 512, 1024}) {
    dfma_tput<<<1, threads>>>(out, clk, iters);
    cudaMemcpy(&h, clk, 8, cudaMemcpyDeviceToHost);
    printf("DFMA  %4d threads: %.2f FMA/clk/SM  (%.2f clk per warp-instr per SMSP)\n", threads,
           (double)threads * 8 * iters / h, (double)h / (8.0 * iters * ((threads + 127) / 128)));
    dmma_tput<<<1, threads>>>(out, clk, iters);
    cudaMemcpy(&h, clk, 8, cudaMemcpyDeviceToHost);
    printf("DMMA  %4d threads: %.2f FMA/clk/SM  (%.2f clk per warp-instr per SMSP)\n", threads,
           (double)(threads / 32) * 256 * 8 * iters / h, (double)h / (8.0 * iters * ((threads + 127) / 128)));
  }
  dfma_lat<<<1, 32>>>(out, clk, iters);
  cudaMemcpy(&h, clk, 8, cudaMemcpyDeviceToHost);
  printf("DFMA dependent latency: %.1f clk\n", (double)h / (8.0 * iters));
  dmma_lat<<<1, 32>>>(out, clk, iters);
  cudaMemcpy(&h, clk, 8, cudaMemcpyDeviceToHost);
  printf("DMMA dependent latency: %.1f clk\n", (double)h / (8.0 * iters));
  cudaError_t e = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(e));
  return 0;
}
