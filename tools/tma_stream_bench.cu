// Microbenchmark: HBM read bandwidth of a persistent TMA streaming kernel (one CTA per SM) as a function of
// the bytes in flight per SM and of the access pattern.
//   mode 0: 1-D bulk copies of contiguous `chunk` bytes (cp.async.bulk.shared::cluster.global)
//   mode 1: 2-D tensor tiles [128 rows x 64 bf16] of a row-major [N, K] matrix (row pitch K*2 bytes), SWIZZLE_128B
//   mode 2: 2-D tensor tiles [64 rows x 64 bf16] of a contiguous [rows, 64] matrix (KV-cache like, 8 KB contiguous)
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/tma_stream_bench tools/tma_stream_bench.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(b)), "r"(parity) : "memory");
  }
}
__device__ __forceinline__ void bulk_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_2d_hint(void* dst, const CUtensorMap* tm, int c0, int c1, uint64_t* bar, uint64_t hint) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(smem_u32(dst)), "l"((uint64_t)tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "l"(hint) : "memory");
}
__device__ __forceinline__ void tma_2d(void* dst, const CUtensorMap* tm, int c0, int c1, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(dst)), "l"((uint64_t)tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}

struct Args { const uint8_t* base; long long n_chunks; int chunk; int stages; int mode; int kblocks; int issuers; int hint; int xload; };

extern __shared__ uint8_t smem_raw[];
__global__ void __launch_bounds__(128) stream_kernel(const __grid_constant__ CUtensorMap tm, const __grid_constant__ CUtensorMap tmx, Args a, unsigned long long* sink) {
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = (uint64_t*)(smem + (size_t)a.stages * (a.chunk + (a.xload ? 8192 : 0)));
  if (threadIdx.x == 0) {
    for (int s = 0; s < a.stages; ++s) mbar_init(&bars[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  // each issuing warp (lane 0) owns stages s = w, w + issuers, ...; a stage is re-armed as soon as it completes
  const int w = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0 && w < a.issuers) {
    const long long per = (a.n_chunks + gridDim.x - 1) / gridDim.x;
    const long long lo = blockIdx.x * per, hi = lo + per < a.n_chunks ? lo + per : a.n_chunks;
    const int my_stages = (a.stages - w + a.issuers - 1) / a.issuers;
    long long next = lo + w;  // chunk index; this warp takes chunks lo + w, lo + w + issuers, ...
    long long issued = 0, done = 0;
    unsigned long long acc = 0;
    auto issue = [&](long long c, int s) {
      mbar_expect_tx(&bars[s], a.chunk + (a.xload ? 8192 : 0));
      if (a.mode == 0) bulk_1d(smem + (size_t)s * a.chunk, a.base + c * a.chunk, a.chunk, &bars[s]);
      else if (a.mode == 1) {
        const long long tile = c / a.kblocks; const int kb = int(c % a.kblocks);
        uint8_t* dst = smem + (size_t)s * (a.chunk + (a.xload ? 8192 : 0));
        if (a.xload) {  // an L2-resident activation tile rides along with every weight tile (rows 0..63 of the matrix)
          tma_2d_hint(dst + a.chunk, &tmx, kb * 64, 0, &bars[s], 0x14F0000000000000ull);
        }
        if (a.hint == 0) tma_2d(dst, &tm, kb * 64, int(tile * 128), &bars[s]);
        else tma_2d_hint(dst, &tm, kb * 64, int(tile * 128), &bars[s], a.hint == 1 ? 0x12F0000000000000ull : 0x14F0000000000000ull);
      }
      else tma_2d(smem + (size_t)s * a.chunk, &tm, 0, int(c * 64), &bars[s]);
    };
    for (int i = 0; i < my_stages && next < hi; ++i, next += a.issuers) { issue(next, w + i * a.issuers); ++issued; }
    while (done < issued) {
      const int i = int(done % my_stages);
      const int s = w + i * a.issuers;
      mbar_wait(&bars[s], (done / my_stages) & 1);
      acc += *(volatile unsigned int*)(smem + (size_t)s * (a.chunk + (a.xload ? 8192 : 0)));
      ++done;
      if (next < hi) { issue(next, s); ++issued; next += a.issuers; }
    }
    if (acc == 0x123456789ull) *sink = acc;
  }
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  const size_t bytes = size_t(2) << 30;  // 2 GiB: 16x the L2
  uint8_t* buf; CK(cudaMalloc(&buf, bytes)); CK(cudaMemset(buf, 1, bytes));
  unsigned long long* sink; CK(cudaMalloc(&sink, 8));
  void* fp = nullptr; cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q));
  EncodeFn enc = (EncodeFn)fp;
  int sms = 0; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  if (getenv("MTX_SMS")) sms = atoi(getenv("MTX_SMS"));  // stream from a subset of the SMs (per-SM rate when the others are idle)
  printf("CTAs: %d\n", sms);
  CK(cudaFuncSetAttribute(stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  auto run = [&](const char* name, int mode, int chunk, int stages, int issuers, int K, int hint = 0, int xload = 0) {
    CUtensorMap tm, tmx; memset(&tm, 0, sizeof(tm)); memset(&tmx, 0, sizeof(tmx));
    Args a; a.base = buf; a.chunk = chunk; a.stages = stages; a.mode = mode; a.issuers = issuers; a.kblocks = K / 64; a.hint = hint; a.xload = xload;
    a.n_chunks = (long long)(bytes / chunk);
    if (mode == 1) {
      const cuuint64_t rows = bytes / (size_t(K) * 2);
      const cuuint64_t dims[2] = {cuuint64_t(K), rows}; const cuuint64_t strides[1] = {cuuint64_t(K) * 2};
      const cuuint32_t box[2] = {64, 128}; const cuuint32_t es[2] = {1, 1};
      if (enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, buf, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) { printf("encode failed\n"); return; }
      a.n_chunks = (long long)(rows / 128) * a.kblocks;
      const cuuint32_t boxx[2] = {64, 64};
      enc(&tmx, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, buf, dims, strides, boxx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    } else if (mode == 2) {
      const cuuint64_t rows = bytes / 128;
      const cuuint64_t dims[2] = {64, rows}; const cuuint64_t strides[1] = {128};
      const cuuint32_t box[2] = {64, cuuint32_t(chunk / 128)}; const cuuint32_t es[2] = {1, 1};
      if (enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, buf, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) { printf("encode failed\n"); return; }
    }
    const size_t smem = 1024 + size_t(stages) * (chunk + (xload ? 8192 : 0)) + 8 * stages + 64;
    if (smem > 227 * 1024) return;
    float best = 1e9f;
    for (int it = 0; it < 4; ++it) {
      CK(cudaEventRecord(e0));
      stream_kernel<<<sms, 128, smem>>>(tm, tmx, a, sink);
      CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
      float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
    }
    CK(cudaGetLastError());
    const double moved = double(a.n_chunks) * chunk;
    printf("%-34s chunk %6d stages %2d issuers %d in-flight/SM %4d KB : %7.1f GB/s\n", name, chunk, stages, issuers, stages * chunk / 1024, moved / best / 1e6);
  };
  for (int st : {4, 5, 8, 12}) run("2-D weight + X tile, no hint", 1, 16384, st, 1, 1280, 0, 1);
  for (int st : {4, 5, 8}) run("2-D weight + X tile, evict_first", 1, 16384, st, 1, 1280, 1, 1);
  return 0;
}
