// PTX wrappers for sm_100a: mbarrier, TMA (cp.async.bulk.tensor), tcgen05 / TMEM,
// programmatic dependent launch, and small bf16 helpers.
#pragma once

#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace mtx {

typedef __nv_bfloat16 bf16;

// ---------------------------------------------------------------------------------
// generic helpers
// ---------------------------------------------------------------------------------

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// Round an fp32 value to bf16 and back (round-to-nearest-even), the rounding the
// reference applies wherever an activation is a bf16 array.
__device__ __forceinline__ float bf16r(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }
// KVQuant.quantize (inference/kvcache.py:76-90) of two values with inv = MAX / scale: int8 -> the bytes u = clip(rint(x inv), -128, 127)
// + 128 (MAX 127.5), fp8 -> float8_e4m3fn(x inv) (MAX 448; round to nearest even, |x inv| <= 448 by construction).
__device__ __forceinline__ uint32_t kv_quant_pair(float x0, float x1, float inv, bool fp8) {
  if (fp8) {
    unsigned short r;
    asm("cvt.rn.satfinite.e4m3x2.f32 %0, %1, %2;" : "=h"(r) : "f"(x1 * inv), "f"(x0 * inv));
    return r;
  }
  const int q0 = int(fminf(fmaxf(rintf(x0 * inv), -128.0f), 127.0f)) + 128, q1 = int(fminf(fmaxf(rintf(x1 * inv), -128.0f), 127.0f)) + 128;
  return uint32_t(q0 | (q1 << 8));
}
__device__ __forceinline__ float kv_quant_max(bool fp8) { return fp8 ? 448.0f : 127.5f; }

__device__ __forceinline__ float bf16_lo(uint32_t packed) { return __uint_as_float(packed << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t packed) { return __uint_as_float(packed & 0xffff0000u); }

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}

__device__ __forceinline__ float ldcg_f32(const float* p) { return __ldcg(p); }

// ---------------------------------------------------------------------------------
// debug timeline: (kind, start ns, end ns) of the first CTA of every kernel of a step
// ---------------------------------------------------------------------------------

__device__ unsigned long long* d_timeline = nullptr;  // [0] = entry counter, entries from [4]

__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ int timeline_begin(int kind) {
  if (threadIdx.x == 0 && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && d_timeline != nullptr) {
    const int idx = int(atomicAdd(d_timeline, 1ull));
    if (idx < 1000) {
      d_timeline[4 + 3 * idx] = kind;
      d_timeline[4 + 3 * idx + 1] = globaltimer_ns();
      return idx;
    }
  }
  return -1;
}
__device__ __forceinline__ void timeline_end(int idx) {
  if (idx >= 0) d_timeline[4 + 3 * idx + 2] = globaltimer_ns();
}

// ---------------------------------------------------------------------------------
// programmatic dependent launch
// ---------------------------------------------------------------------------------

// Blocks until every kernel this one depends on has completed and its writes are visible.
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
// Lets the next kernel in the stream start its prologue (weight prefetch) early.
__device__ __forceinline__ void griddep_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// Named barrier 1 over the 128 "worker" threads of a CTA (GEMM epilogue warps / attention warps).
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

// ---------------------------------------------------------------------------------
// mbarrier
// ---------------------------------------------------------------------------------

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}

__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}

// A wait that lasts ~2 s is a protocol bug (wrong tx byte count, missing arrive): report and trap instead of
// hanging the GPU.  Out of line: the persistent kernel has dozens of wait sites and lives off its instruction cache.
__device__ __noinline__ void mtx_wait_timeout(int what, int a, int b) {
  printf("mtx: wait %d timed out (block %d,%d,%d thread %d: %d %d)\n", what, blockIdx.x, blockIdx.y, blockIdx.z, threadIdx.x, a, b);
  __trap();
}

// Waits for the phase with the given parity.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) mtx_wait_timeout(0, int(parity), 0);
  }
}

// ---------------------------------------------------------------------------------
// TMA
// ---------------------------------------------------------------------------------

constexpr uint64_t kEvictNormal = 0x1000000000000000ull;
constexpr uint64_t kEvictFirst = 0x12F0000000000000ull;
constexpr uint64_t kEvictLast = 0x14F0000000000000ull;

__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* tm) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(tm)) : "memory");
}

// 2-D tiled load: box at element coordinates (c0 = innermost, c1) -> smem, completion on `bar`.
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* tm, int c0, int c1, uint64_t* bar, uint64_t hint) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
      " [%0], [%1, {%3, %4}], [%2], %5;"
      :
      : "r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "l"(hint)
      : "memory");
}

// ---------------------------------------------------------------------------------
// tcgen05 / TMEM
// ---------------------------------------------------------------------------------

__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem] * B[smem], bf16 operands, fp32 accumulate, single CTA.
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      :
      : "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}

// mbarrier arrive once every previously issued tcgen05.mma of this thread has completed.
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// 16 consecutive fp32 columns of this thread's TMEM lane.
__device__ __forceinline__ void tmem_ld_x16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// Shared-memory matrix descriptor for a K-major bf16 tile whose rows are 128 bytes
// (64 elements) laid out by TMA with SWIZZLE_128B: 8-row groups are 1024 B apart.
// Field layout: cute/arch/mma_sm100_desc.hpp `SmemDescriptor` (start>>4 [0,14),
// LBO>>4 [16,30), SBO>>4 [32,46), version=1 [46,48), layout_type=2 (SW128) [61,64)).
__device__ __forceinline__ uint64_t umma_desc_sw128(const void* smem_tile) {
  const uint64_t addr = smem_u32(smem_tile);
  uint64_t d = 0;
  d |= (addr & 0x3FFFFull) >> 4;
  d |= uint64_t(1) << 16;          // leading byte offset: unused for swizzled K-major
  d |= uint64_t(1024 >> 4) << 32;  // stride byte offset between 8-row groups
  d |= uint64_t(1) << 46;          // descriptor version (Blackwell)
  d |= uint64_t(2) << 61;          // SWIZZLE_128B
  return d;
}

// Instruction descriptor, kind::f16, A = B = bf16, D = fp32, both operands K-major.
// Field layout: cute/arch/mma_sm100_desc.hpp `InstrDescriptor`.
__host__ __device__ inline uint32_t umma_idesc_bf16(int M, int N) {
  uint32_t d = 0;
  d |= 1u << 4;                // c_format  = F32
  d |= 1u << 7;                // a_format  = BF16
  d |= 1u << 10;               // b_format  = BF16
  d |= uint32_t(N >> 3) << 17; // n_dim
  d |= uint32_t(M >> 4) << 24; // m_dim
  return d;
}

// ---------------------------------------------------------------------------------
// legacy warp-level MMA (decode attention: 16 x 8 x 16, bf16 in, fp32 out) and ldmatrix
// ---------------------------------------------------------------------------------

__device__ __forceinline__ void mma_m16n8k16_bf16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__device__ __forceinline__ void ldmatrix_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr) : "memory");
}

__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr) : "memory");
}

// ---------------------------------------------------------------------------------
// Philox-4x32-10 (sampler noise; oracle/decode_ref.py:philox4x32_10 is the same stream)
// ---------------------------------------------------------------------------------

__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += 0x9E3779B9u;
    k.y += 0xBB67AE85u;
  }
  return c;
}

// Gumbel(0,1) noise for (row, vocab id) at a given step: word (v & 3) of
// philox(counter = (v >> 2, row, step, 0), key = seed).
__device__ __noinline__ float gumbel_noise(uint64_t seed, uint32_t step, uint32_t row, uint32_t v) {
  const uint4 w = philox4x32_10(make_uint4(v >> 2, row, step, 0u), make_uint2(uint32_t(seed), uint32_t(seed >> 32)));
  const uint32_t sel = v & 3u;
  const uint32_t x = sel == 0 ? w.x : sel == 1 ? w.y : sel == 2 ? w.z : w.w;
  const float u = (float(x >> 9) + 0.5f) * 1.1920928955078125e-07f;  // 2^-23
  return -logf(-logf(u));
}

// The noise of four consecutive vocabulary ids v .. v+3 with v a multiple of 4: ONE Philox block (the same values as four
// gumbel_noise calls, a quarter of the work).
__device__ __forceinline__ float4 gumbel_noise4(uint64_t seed, uint32_t step, uint32_t row, uint32_t v) {
  const uint4 w = philox4x32_10(make_uint4(v >> 2, row, step, 0u), make_uint2(uint32_t(seed), uint32_t(seed >> 32)));
  const float s = 1.1920928955078125e-07f;  // 2^-23
  return make_float4(-logf(-logf((float(w.x >> 9) + 0.5f) * s)), -logf(-logf((float(w.y >> 9) + 0.5f) * s)),
                     -logf(-logf((float(w.z >> 9) + 0.5f) * s)), -logf(-logf((float(w.w >> 9) + 0.5f) * s)));
}

}  // namespace mtx
