// Which tiled-TMA load configurations does this device accept?  One configuration per process (an illegal instruction kills the
// context):  ./tma_probe rank innerElems swizzle(0 none / 3 128B) l2promo(0 none / 2 256B) elem(4 fp32 / 2 bf16)
#include <cstdio>
#include <cstdlib>
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>

__global__ void probe(const __grid_constant__ CUtensorMap map, float *out, int rank, int bytes, int c0) {
  extern __shared__ __align__(128) unsigned char sm[];
  __shared__ __align__(8) unsigned long long bar;
  const uint32_t b = (uint32_t)__cvta_generic_to_shared(&bar), d = (uint32_t)__cvta_generic_to_shared(sm);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
    if (rank == 4)
      asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                   ::"r"(d), "l"((uint64_t)&map), "r"(b), "r"(c0), "r"(c0), "r"(0), "r"(0) : "memory");
    else
      asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                   ::"r"(d), "l"((uint64_t)&map), "r"(b), "r"(c0), "r"(c0), "r"(0) : "memory");
  }
  __syncthreads();
  uint32_t done = 0;
  while (!done)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(b) : "memory");
  for (int i = threadIdx.x; i < bytes / 4; i += blockDim.x) out[i] = ((float *)sm)[i];
}

int main(int argc, char **argv) {
  const int rank = atoi(argv[1]), inner = atoi(argv[2]), sw = atoi(argv[3]), l2 = atoi(argv[4]), es = atoi(argv[5]);
  const int W = 64, H = 64, B = 2;
  const int rows = argc > 6 ? atoi(argv[6]) : 20, Cc = argc > 7 ? atoi(argv[7]) : 3, c0 = argc > 8 ? atoi(argv[8]) : -2;
  float *g, *out;
  cudaMalloc(&g, (size_t)W * H * Cc * B * 4);
  cudaMemset(g, 0, (size_t)W * H * Cc * B * 4);
  cuuint64_t dims[4] = {(cuuint64_t)(W * 4 / es), (cuuint64_t)H, (cuuint64_t)Cc, (cuuint64_t)B};
  cuuint64_t strides[3] = {(cuuint64_t)W * 4, (cuuint64_t)W * H * 4, (cuuint64_t)W * H * Cc * 4};
  cuuint32_t box[4] = {(cuuint32_t)inner, (cuuint32_t)rows, (cuuint32_t)Cc, 1}, estr[4] = {1, 1, 1, 1};
  CUtensorMap m;
  void *fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  typedef CUresult (*Enc)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                          const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  CUresult r = ((Enc)fn)(&m, es == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, g, dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, (CUtensorMapSwizzle)sw, (CUtensorMapL2promotion)l2, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  const int bytes = inner * es * rows * Cc * (rank == 4 ? 1 : 1);
  cudaMalloc(&out, bytes);
  printf("rank %d inner %d x %dB swizzle %d l2 %d rows %d C %d c0 %d: encode %d, box %d bytes: ", rank, inner, es, sw, l2, rows, Cc, c0, (int)r, bytes);
  if (r != CUDA_SUCCESS) { printf("\n"); return 0; }
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 << 10);
  probe<<<1, 128, bytes + 128>>>(m, out, rank, bytes, c0);
  cudaError_t e = cudaDeviceSynchronize();
  printf("%s\n", cudaGetErrorString(e));
  return 0;
}
