// Issue rates of the conversion / special-function instructions that kornia-style coordinate math in fp64 needs, on B200.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o xu_rate xu_rate.cu && ./xu_rate
// Why: ncu showed warp_rgb_kernel with the XU pipe at 100 % of its sustained peak while DRAM sat at 15-25 %: the fp64
// arithmetic itself is cheap (tools/micro/dfma_rate.cu: DFMA 55 per clock per SM), the int <-> fp64 <-> fp32 conversions,
// floor() on a double and the reciprocal seed are XU instructions.  Each kernel below runs 16 independent chains of ONE
// such instruction plus cheap glue on another pipe.
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ITERS = 512, CHAINS = 16;

#define KERNEL(NAME, TYPE, INIT, STEP)                                        \
  __global__ void NAME(double *out, double a, int n) {                         \
    TYPE acc[CHAINS];                                                          \
    _Pragma("unroll") for (int i = 0; i < CHAINS; ++i) acc[i] = INIT;          \
    for (int it = 0; it < n; ++it) {                                           \
      _Pragma("unroll") for (int i = 0; i < CHAINS; ++i) { STEP; }             \
    }                                                                          \
    double s = 0;                                                              \
    _Pragma("unroll") for (int i = 0; i < CHAINS; ++i) s += (double)acc[i];    \
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;                            \
  }

// int -> double (I2F.F64.S32) + IADD glue (result folded back through a cheap integer op so the chain is serial per i)
KERNEL(k_i2d, int, (int)threadIdx.x + i, { double d = (double)acc[i]; acc[i] += __double2hiint(d) & 3; })
// double -> int (F2I.S32.F64) + DADD glue
KERNEL(k_d2i, double, a + threadIdx.x + i, { int v = __double2int_rd(acc[i]); acc[i] += 0.25 + __hiloint2double(0, v & 1); })
// floor(double) (FRND.F64.FLOOR)
KERNEL(k_floor, double, a * 3.7 + threadIdx.x + i, { acc[i] = floor(acc[i]) + 0.3; })
// double -> float (F2F.F32.F64) + FADD glue
KERNEL(k_d2f, double, a + threadIdx.x + i, { float f = (float)acc[i]; acc[i] += 0.125 + __hiloint2double(0, __float_as_int(f) & 1); })
// float -> double (F2F.F64.F32)
KERNEL(k_f2d, float, (float)a + threadIdx.x + i, { double d = (double)acc[i]; acc[i] += __int_as_float(__double2hiint(d) & 0x3f800000); })
// MUFU.RCP (fp32 reciprocal, approximate)
__device__ __forceinline__ float rcp32(float z) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(z)); return r; }
KERNEL(k_rcp32, float, (float)a + threadIdx.x + i, { acc[i] = rcp32(acc[i]) + 1.5f; })
// MUFU.RCP64H (rcp.approx.ftz.f64: the seed of fp64 division)
__device__ __forceinline__ double rcp64h(double z) { double r; asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(z)); return r; }
KERNEL(k_rcp64h, double, a + threadIdx.x + i, { acc[i] = rcp64h(acc[i]) + 1.5; })
// baseline: DFMA chain
KERNEL(k_dfma, double, a + threadIdx.x + i, { acc[i] = fma(acc[i], 0.999, 0.25); })

template <typename K>
static void run(const char *name, K kern) {
  int dev, sms, clk;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, dev);
  double *out;
  const int blocks = sms * 8, threads = 256;
  cudaMalloc(&out, sizeof(double) * blocks * threads);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  kern<<<blocks, threads>>>(out, 1.0001, ITERS);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  kern<<<blocks, threads>>>(out, 1.0001, ITERS);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  const double ops = (double)ITERS * CHAINS * blocks * threads;
  printf("%-44s %8.3f ms  %7.2f per clk per SM at the nominal %d MHz\n", name, ms, ops / (ms * 1e-3) / sms / (clk * 1e3), clk / 1000);
  cudaFree(out);
}

int main() {
  run("DFMA (baseline)", k_dfma);
  run("I2F.F64.S32  (double)int", k_i2d);
  run("F2I.S32.F64  (int)double", k_d2i);
  run("FRND.F64.FLOOR  floor(double)", k_floor);
  run("F2F.F32.F64  (float)double", k_d2f);
  run("F2F.F64.F32  (double)float", k_f2d);
  run("MUFU.RCP  fp32 reciprocal", k_rcp32);
  run("MUFU.RCP64H  rcp.approx.ftz.f64", k_rcp64h);
  return 0;
}
