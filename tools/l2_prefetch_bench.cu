// Microbenchmark: does cp.async.bulk.prefetch.tensor.2d.L2 make a later TMA stream of the same tiles faster?
// A [N, K] bf16 matrix of `mb` megabytes is streamed by `ctas` CTAs (4-stage ring of 128x64 tiles), either cold
// (L2 flushed) or after a prefetch kernel touched every tile.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/l2_prefetch_bench tools/l2_prefetch_bench.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cstring>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_2d(void* dst, const CUtensorMap* tm, int c0, int c1, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(dst)), "l"((uint64_t)tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_pf(const CUtensorMap* tm, int c0, int c1) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global [%0, {%1, %2}];" ::"l"((uint64_t)tm), "r"(c0), "r"(c1) : "memory");
}
extern __shared__ uint8_t smem_raw[];
__global__ void prefetch_kernel(const __grid_constant__ CUtensorMap tm, int boxes, int kbt) {
  if (threadIdx.x == 0)
    for (int b = blockIdx.x; b < boxes; b += gridDim.x) tma_pf(&tm, (b % kbt) * 64, (b / kbt) * 128);
}
__device__ __forceinline__ void bulk_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ const uint8_t* g_base = nullptr;
__global__ void stream_kernel(const __grid_constant__ CUtensorMap tm, int boxes, int kbt, int stages, unsigned long long* sink) {
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = (uint64_t*)(smem + stages * 16384);
  if (threadIdx.x == 0) { for (int s = 0; s < stages; ++s) mbar_init(&bars[s], 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  __syncthreads();
  if (threadIdx.x == 0) {
    const int per = (boxes + gridDim.x - 1) / gridDim.x, lo = blockIdx.x * per, hi = min(boxes, lo + per);
    int issued = 0, done = 0; unsigned long long acc = 0;
    auto issue = [&](int b, int s) { mbar_expect_tx(&bars[s], 16384); if (g_base) bulk_1d(smem + s * 16384, g_base + size_t(b) * 16384, 16384, &bars[s]); else tma_2d(smem + s * 16384, &tm, (b % kbt) * 64, (b / kbt) * 128, &bars[s]); };
    for (; issued < stages && lo + issued < hi; ++issued) issue(lo + issued, issued);
    while (done < issued) {
      const int s = done % stages;
      mbar_wait(&bars[s], (done / stages) & 1);
      acc += *(volatile unsigned*)(smem + s * 16384);
      ++done;
      if (lo + issued < hi) { issue(lo + issued, s); ++issued; }
    }
    if (acc == 0x12345) *sink = acc;
  }
}
typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main() {
  const int K = 1280, kbt = K / 64;
  uint8_t* flush; const size_t flush_bytes = size_t(512) << 20; CK(cudaMalloc(&flush, flush_bytes));
  unsigned long long* sink; CK(cudaMalloc(&sink, 8));
  void* fp = nullptr; cudaDriverEntryPointQueryResult q; CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q));
  EncodeFn enc = (EncodeFn)fp;
  CK(cudaFuncSetAttribute(stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  for (int mb : {80, 400}) {
    const size_t rows = size_t(mb) * (1 << 20) / (K * 2) / 128 * 128;
    uint8_t* buf; CK(cudaMalloc(&buf, rows * K * 2)); CK(cudaMemset(buf, 1, rows * K * 2));
    CUtensorMap tm; const cuuint64_t dims[2] = {cuuint64_t(K), rows}; const cuuint64_t strides[1] = {cuuint64_t(K) * 2};
    const cuuint32_t box[2] = {64, 128}; const cuuint32_t es[2] = {1, 1};
    enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, buf, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    const int boxes = int(rows / 128) * kbt;
    for (int mode1d = 0; mode1d < 2; ++mode1d) for (int ctas : {40, 80, 148}) for (int pf = 0; pf < 1; ++pf) {
      { const uint8_t* bp = mode1d ? buf : nullptr; CK(cudaMemcpyToSymbol(g_base, &bp, sizeof(bp))); }
      float best = 1e9f, pf_ms = 0;
      for (int it = 0; it < 3; ++it) {
        CK(cudaMemset(flush, it, flush_bytes));  // evict
        if (pf) {
          CK(cudaEventRecord(e0)); prefetch_kernel<<<148, 32>>>(tm, boxes, kbt); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
          CK(cudaEventElapsedTime(&pf_ms, e0, e1));
          CK(cudaDeviceSynchronize());
        }
        CK(cudaEventRecord(e0));
        stream_kernel<<<ctas, 32, 1024 + 4 * 16384 + 64>>>(tm, boxes, kbt, 4, sink);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
      }
      printf("%s %3d MB, %3d CTAs, prefetch %d: stream %.1f us (%.0f GB/s, %.0f GB/s per CTA)%s\n", mode1d ? "1-D contiguous" : "2-D 128x128B ", mb, ctas, pf, best * 1e3, double(boxes) * 16384 / best / 1e6, double(boxes) * 16384 / best / 1e6 / ctas, pf ? "" : "");
      if (pf) printf("        prefetch kernel itself: %.1f us\n", pf_ms * 1e3);
    }
    CK(cudaFree(buf));
  }
  return 0;
}
