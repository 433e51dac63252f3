// Issue-rate microbenchmark: three-source FFMA, FFMA with a uniform-register operand, packed FFMA2 (fma.rn.f32x2).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma_rate ffma_rate.cu && ./ffma_rate
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ITERS = 4096, CHAINS = 16;

__global__ void k_ffma(float *out, float a, float b) {
  float acc[CHAINS];
  float x = threadIdx.x * 1e-3f + a, y = b + threadIdx.x * 1e-4f;   // per-thread register operands
#pragma unroll
  for (int i = 0; i < CHAINS; ++i) acc[i] = i;
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) acc[i] = fmaf(acc[i], x, y);
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < CHAINS; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_ffma_ur(float *out, float a, float b) {
  float acc[CHAINS];
  float x = threadIdx.x * 1e-3f + 1.0f;
#pragma unroll
  for (int i = 0; i < CHAINS; ++i) acc[i] = i + x;
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) acc[i] = fmaf(x, a, acc[i]);   // a: kernel parameter -> uniform register / constant operand
  }
  float s = b;
#pragma unroll
  for (int i = 0; i < CHAINS; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_ffma2(float *out, float a, float b) {
  float2 acc[CHAINS / 2];
  float2 x = make_float2(threadIdx.x * 1e-3f + a, threadIdx.x * 2e-3f + a), y = make_float2(b + threadIdx.x * 1e-4f, b);
#pragma unroll
  for (int i = 0; i < CHAINS / 2; ++i) acc[i] = make_float2(i, -i);
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < CHAINS / 2; ++i) acc[i] = __ffma2_rn(acc[i], x, y);
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < CHAINS / 2; ++i) s += acc[i].x + acc[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename K>
static void run(const char *name, K kern, double fma_per_thread) {
  int dev, sms;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int clk;
  cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, dev);
  float *out;
  const int blocks = sms * 8, threads = 256;
  cudaMalloc(&out, sizeof(float) * blocks * threads);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  kern<<<blocks, threads>>>(out, 1.0001f, 0.5f);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  kern<<<blocks, threads>>>(out, 1.0001f, 0.5f);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  const double fma = fma_per_thread * blocks * threads;
  printf("%-28s %8.3f ms  %7.1f FMA/clk/SM at the nominal %d MHz (%.1f TFLOP/s)\n", name, ms, fma / (ms * 1e-3) / sms / (clk * 1e3), clk / 1000,
         2 * fma / (ms * 1e-3) / 1e12);
  cudaFree(out);
}

int main() {
  run("FFMA (3 register sources)", k_ffma, (double)ITERS * CHAINS);
  run("FFMA (uniform operand)", k_ffma_ur, (double)ITERS * CHAINS);
  run("FFMA2 (packed f32x2)", k_ffma2, (double)ITERS * CHAINS);
  return 0;
}
