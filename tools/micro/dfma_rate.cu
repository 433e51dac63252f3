// Issue-rate microbenchmark of the FP64 pipe on B200: DFMA, DADD, and the F32 -> F64 conversion, beside FFMA.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dfma_rate dfma_rate.cu && ./dfma_rate
// Why: kernels of this library that touch fp64 per element (warp coordinates, squared-error sums) sit below the HBM
// roofline; this measures what an fp64 instruction costs.  B200, r04: FFMA 115, DFMA 55, DADD 61 per clock per SM (fp64
// arithmetic at half the fp32 rate -- not the bound of those kernels), F32 -> F64 conversion + DADD pairs 15.7 (the
// conversion issues at 16 per clock per SM).
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ITERS = 1024, CHAINS = 16;

__global__ void k_dfma(double *out, double a, double b) {
  double acc[CHAINS];
  const double x = threadIdx.x * 1e-3 + a, y = b + threadIdx.x * 1e-4;
#pragma unroll
  for (int i = 0; i < CHAINS; ++i) acc[i] = i;
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) acc[i] = fma(acc[i], x, y);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < CHAINS; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_dadd(double *out, double a, double b) {
  double acc[CHAINS];
  const double x = threadIdx.x * 1e-3 + a;
#pragma unroll
  for (int i = 0; i < CHAINS; ++i) acc[i] = i * b;
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) acc[i] += x;
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < CHAINS; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// one F2F.F64.F32 + one DADD per step (the per-element cost of "accumulate fp32 values in fp64")
__global__ void k_cvt_dadd(double *out, double a, double b) {
  double acc[CHAINS];
  float f[CHAINS];
#pragma unroll
  for (int i = 0; i < CHAINS; ++i) { acc[i] = i * b; f[i] = (float)(threadIdx.x * 1e-3 + a + i); }
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) { acc[i] += (double)f[i]; f[i] = f[i] * 1.0001f; }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < CHAINS; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_ffma(double *out, double a, double b) {
  float acc[CHAINS];
  const float x = threadIdx.x * 1e-3f + (float)a, y = (float)b + threadIdx.x * 1e-4f;
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

template <typename K>
static void run(const char *name, K kern, double ops_per_thread) {
  int dev, sms, clk;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, dev);
  double *out;
  const int blocks = sms * 8, threads = 256;
  cudaMalloc(&out, sizeof(double) * blocks * threads);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  kern<<<blocks, threads>>>(out, 1.0001, 0.5);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  kern<<<blocks, threads>>>(out, 1.0001, 0.5);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  const double ops = ops_per_thread * blocks * threads;
  printf("%-34s %8.3f ms  %7.2f per clk per SM at the nominal %d MHz\n", name, ms, ops / (ms * 1e-3) / sms / (clk * 1e3), clk / 1000);
  cudaFree(out);
}

int main() {
  run("FFMA", k_ffma, (double)ITERS * CHAINS);
  run("DFMA", k_dfma, (double)ITERS * CHAINS);
  run("DADD", k_dadd, (double)ITERS * CHAINS);
  run("F2F.F64.F32 + DADD (pairs)", k_cvt_dadd, (double)ITERS * CHAINS);
  return 0;
}
