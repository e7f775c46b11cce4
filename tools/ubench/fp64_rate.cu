// Micro-benchmark (GPU needed): fp64 FMA rate of the vector pipe (DFMA) vs the tensor-core DMMA
// (mma.sync.m8n8k4.f64) on this part. Decides how the outer-product kernel of the pose-graph solver
// should issue its multiplications. Build + run:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/fp64_rate tools/ubench/fp64_rate.cu && /tmp/fp64_rate
#include <cstdio>
#include <cuda_runtime.h>

__global__ void dfma_loop(double* out, int iters) {
  double a[8], b = 1.0000001, c = 0.9999999;
  for (int i = 0; i < 8; ++i) a[i] = threadIdx.x + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = fma(a[i], b, c);
  }
  double s = 0;
  for (int i = 0; i < 8; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void dmma_loop(double* out, int iters) {
  double c[4][2];
  for (int i = 0; i < 4; ++i) c[i][0] = c[i][1] = threadIdx.x + i;
  double a = 1.0000001, b = 0.9999999;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
  }
  double s = 0;
  for (int i = 0; i < 4; ++i) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
  double* out;
  cudaMalloc(&out, sizeof(double) * 148 * 8 * 1024);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int iters = 20000;
  for (int warps = 4; warps <= 32; warps *= 2) {
    const int blocks = 148 * 2, threads = warps * 32 / 2;
    float ms;
    dfma_loop<<<blocks, threads>>>(out, 100);
    cudaEventRecord(e0);
    dfma_loop<<<blocks, threads>>>(out, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    const double f1 = 2.0 * 8 * iters * (double)blocks * threads / (ms * 1e-3) / 1e12;
    dmma_loop<<<blocks, threads>>>(out, 100);
    cudaEventRecord(e0);
    dmma_loop<<<blocks, threads>>>(out, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    // one m8n8k4 = 8*8*4 FMA = 512 flop per warp
    const double f2 = 512.0 * 4 * iters * (double)blocks * (threads / 32) / (ms * 1e-3) / 1e12;
    printf("warps/SM %2d: DFMA %.2f TFLOP/s   DMMA m8n8k4 %.2f TFLOP/s   (%s)\n", warps, f1, f2,
           cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
