// XLA FFI handlers for the C ABI of include/mtx_b200.h, so that the reference's Python/JAX host code can call the
// sm_100a kernels as XLA custom calls:
//
//   jax.ffi.register_ffi_target("mtx_ragged_attention", jax.ffi.pycapsule(lib.MtxRaggedAttention), platform="CUDA")
//   out, m, l = jax.ffi.ffi_call("mtx_ragged_attention", (out_t, m_t, l_t))(q, k, v, lengths, scratch, seq_major=1, softcap=0.0)
//
// (maxtext_indextts2_b200/jax_ffi.py does the registration; INTEGRATION.md shows where MaxText calls it.)
//
// The reference has no FFI of its own; the handler signatures follow the Python-level seams it does have:
//   MtxRaggedAttention  -- AttentionOp.gpu_ragged_attention (MaxText/layers/attentions.py:761-815): (q, k, v, lengths) ->
//                          (unnormalised out, max, sum), merged by the caller (attentions.py:1454-1464)
//   MtxDecodeAttention  -- AttentionOp.__call__ in autoregressive mode over this library's two-segment cache
//   MtxQkvRopeAppend    -- Attention.query/key/value + RotaryEmbedding + KVCache append (in-place cache: input_output_aliases)
//   MtxDecodeStep       -- MaxEngine._generate_jit (MaxText/maxengine.py:868-936): the whole step on a bound engine
//   MtxOutprojResidual  -- Attention.out projection + the residual add (attentions.py:2017-2030, llama2.py:139-140)
//   MtxMlp              -- MlpBlock with its pre-norm + the residual add (linears.py:425-476, llama2.py:150-163)
//   MtxPagedAppend      -- PagedAttentionOp.update_decode_step_pages (MaxText/inference/paged_attention.py:446-471), pools donated
//   MtxPagedAttention   -- PagedAttentionOp.paged_attention_v1_decode (:302-346) on the reference's pools and PageState arrays
//
// COMPILE GUARD.  jaxlib's headers (xla/ffi/api/ffi.h) are not in this image and jax cannot be installed (no network), so
// this translation unit compiles to the single symbol mtx_jax_ffi_available() == 0 here; with the headers on the include
// path (`_lib.build()` adds jaxlib/include when jaxlib is importable) the handlers below are built.  They could not be
// exercised in this environment: the C ABI they forward to is what tests/ covers.
#include <stdint.h>

#include "../../include/mtx_b200.h"

#if defined(__has_include)
#if __has_include("xla/ffi/api/ffi.h")
#define MTX_HAVE_XLA_FFI 1
#endif
#endif

#ifdef MTX_HAVE_XLA_FFI

#include <cuda_runtime_api.h>

#include "xla/ffi/api/ffi.h"

namespace ffi = xla::ffi;

namespace {

ffi::Error Status(int rc) {
  if (rc == MTX_OK) return ffi::Error::Success();
  return ffi::Error(rc == MTX_ERR_CUDA ? ffi::ErrorCode::kInternal : ffi::ErrorCode::kInvalidArgument, mtx_last_error());
}

// (q [B,1,Hq,D] or [B,Hq,D], k / v [B,S,Hkv,D] (seq_major = 1) or [B,Hkv,S,D], lengths [B], scratch u8[...]) ->
// (out like q, max [B,Hq], sum [B,Hq])
ffi::Error RaggedAttentionImpl(cudaStream_t stream, ffi::Buffer<ffi::BF16> q, ffi::Buffer<ffi::BF16> k, ffi::Buffer<ffi::BF16> v,
                               ffi::Buffer<ffi::S32> lengths, ffi::Buffer<ffi::U8> scratch, ffi::ResultBuffer<ffi::BF16> out,
                               ffi::ResultBuffer<ffi::F32> out_max, ffi::ResultBuffer<ffi::F32> out_sum, int32_t seq_major, float softcap) {
  const auto kd = k.dimensions();
  const auto qd = q.dimensions();
  if (kd.size() != 4 || qd.size() < 3) return ffi::Error(ffi::ErrorCode::kInvalidArgument, "q must be [B,(1,)Hq,D], k/v rank 4");
  const int rows = int(kd[0]);
  const int seq_len = int(seq_major ? kd[1] : kd[2]);
  const int hkv = int(seq_major ? kd[2] : kd[1]);
  const int d = int(kd[3]);
  const int hq = int(qd[qd.size() - 2]);
  if (scratch.element_count() < mtx_ragged_attention_scratch_bytes(rows, hkv, hq, d, seq_len))
    return ffi::Error(ffi::ErrorCode::kInvalidArgument, "scratch smaller than mtx_ragged_attention_scratch_bytes()");
  return Status(mtx_ragged_attention(q.untyped_data(), k.untyped_data(), v.untyped_data(), lengths.typed_data(), out->untyped_data(),
                                     out_max->typed_data(), out_sum->typed_data(), rows, seq_len, hq, hkv, d, seq_major, softcap,
                                     scratch.untyped_data(), stream));
}

// (q [rows,Hq*D], k_cache / v_cache [slots,Hkv,T,D], plane, len0, ring_first, ring_len [rows], scratch) -> out [rows,Hq*D]
ffi::Error DecodeAttentionImpl(cudaStream_t stream, ffi::Buffer<ffi::BF16> q, ffi::Buffer<ffi::BF16> k_cache, ffi::Buffer<ffi::BF16> v_cache,
                               ffi::Buffer<ffi::S32> plane, ffi::Buffer<ffi::S32> len0, ffi::Buffer<ffi::S32> ring_first,
                               ffi::Buffer<ffi::S32> ring_len, ffi::Buffer<ffi::U8> scratch, ffi::ResultBuffer<ffi::BF16> out,
                               int32_t num_q_heads, int32_t max_prefill_len, float softcap) {
  const auto kd = k_cache.dimensions();
  if (kd.size() != 4) return ffi::Error(ffi::ErrorCode::kInvalidArgument, "k_cache must be [slots,Hkv,T,D]");
  const int rows = int(q.dimensions()[0]);
  return Status(mtx_decode_attention(q.untyped_data(), k_cache.untyped_data(), v_cache.untyped_data(), plane.typed_data(), len0.typed_data(),
                                     ring_first.typed_data(), ring_len.typed_data(), out->untyped_data(), rows, int(kd[0]), num_q_heads,
                                     int(kd[1]), int(kd[3]), max_prefill_len, int(kd[2]), softcap, scratch.untyped_data(), stream));
}

// (n [rows_padded,E], wqkv [(Hq+2Hkv)D,E], pos, plane, write_row [rows], k_cache, v_cache [planes,Hkv,T,D] (aliased to the
// results of the same index), scratch) -> (q [rows,Hq*D], k_cache, v_cache)
ffi::Error QkvRopeAppendImpl(cudaStream_t stream, ffi::Buffer<ffi::BF16> n, ffi::Buffer<ffi::BF16> wqkv, ffi::Buffer<ffi::S32> pos,
                             ffi::Buffer<ffi::S32> plane, ffi::Buffer<ffi::S32> write_row, ffi::Buffer<ffi::BF16> k_cache,
                             ffi::Buffer<ffi::BF16> v_cache, ffi::Buffer<ffi::U8> scratch, ffi::ResultBuffer<ffi::BF16> q_out,
                             ffi::ResultBuffer<ffi::BF16> k_out, ffi::ResultBuffer<ffi::BF16> v_out, int32_t num_q_heads,
                             float rope_min_timescale, float rope_max_timescale) {
  if (k_out->untyped_data() != k_cache.untyped_data() || v_out->untyped_data() != v_cache.untyped_data())
    return ffi::Error(ffi::ErrorCode::kInvalidArgument, "the caches must be donated: input_output_aliases={5: 1, 6: 2}");
  const auto kd = k_cache.dimensions();
  const int rows = int(pos.dimensions()[0]);
  return Status(mtx_qkv_rope_append(n.untyped_data(), wqkv.untyped_data(), pos.typed_data(), plane.typed_data(), write_row.typed_data(),
                                    q_out->untyped_data(), k_out->untyped_data(), v_out->untyped_data(), rows, int(n.dimensions()[1]),
                                    num_q_heads, int(kd[1]), int(kd[3]), int(kd[2]), rope_min_timescale, rope_max_timescale,
                                    scratch.untyped_data(), stream));
}

// The whole step on an engine created and bound through the C ABI (mtx_engine_create / mtx_engine_bind); `engine` is its handle
// as an integer attribute.  The decode-state operands are donated (maxengine.py:868 donate_argnums) and re-attached if XLA moved
// them: (k_cache, v_cache, tokens, next_pos, generated, prefill_len, ar_lengths, ar_index, rng_state) -> (the same nine, result [B,3]).
ffi::Error DecodeStepImpl(cudaStream_t stream, ffi::Buffer<ffi::BF16> k_cache, ffi::Buffer<ffi::BF16> v_cache, ffi::Buffer<ffi::S32> tokens,
                          ffi::Buffer<ffi::S32> next_pos, ffi::Buffer<ffi::S32> generated, ffi::Buffer<ffi::S32> prefill_len,
                          ffi::Buffer<ffi::S32> ar_lengths, ffi::Buffer<ffi::S32> ar_index, ffi::Buffer<ffi::U32> rng_state,
                          ffi::ResultBuffer<ffi::BF16> k_out, ffi::ResultBuffer<ffi::BF16> v_out, ffi::ResultBuffer<ffi::S32> tokens_out,
                          ffi::ResultBuffer<ffi::S32> next_pos_out, ffi::ResultBuffer<ffi::S32> generated_out,
                          ffi::ResultBuffer<ffi::S32> prefill_len_out, ffi::ResultBuffer<ffi::S32> ar_lengths_out,
                          ffi::ResultBuffer<ffi::S32> ar_index_out, ffi::ResultBuffer<ffi::U32> rng_out, ffi::ResultBuffer<ffi::S32> result,
                          int64_t engine, int32_t rows) {
  if (k_out->untyped_data() != k_cache.untyped_data() || tokens_out->untyped_data() != tokens.untyped_data())
    return ffi::Error(ffi::ErrorCode::kInvalidArgument, "decode state must be donated (input_output_aliases for operands 0..8)");
  mtx_decode_state s = {};  // (quantised-cache and page-pool fields stay null: the dense bf16 cache)
  s.k_cache = k_out->untyped_data();
  s.v_cache = v_out->untyped_data();
  s.tokens = tokens_out->typed_data();
  s.next_pos = next_pos_out->typed_data();
  s.generated = generated_out->typed_data();
  s.prefill_len = prefill_len_out->typed_data();
  s.ar_lengths = ar_lengths_out->typed_data();
  s.ar_index = ar_index_out->typed_data();
  s.result = result->typed_data();
  s.log_prob = nullptr;
  s.logits = nullptr;
  s.rng_state = rng_out->typed_data();
  mtx_engine* e = reinterpret_cast<mtx_engine*>(static_cast<intptr_t>(engine));
  const int rc = mtx_engine_rebind_state(e, &s);
  if (rc != MTX_OK) return Status(rc);
  return Status(mtx_decode_step(e, rows, stream));
}

// (key_pages, value_pages [Hkv, num_pages, tokens_per_page, D] (donated), key, value [B, Hkv, D] or [B, 1, Hkv, D], active_page [B],
// active_page_position [B]) -> (key_pages, value_pages)
ffi::Error PagedAppendImpl(cudaStream_t stream, ffi::Buffer<ffi::BF16> k_pages, ffi::Buffer<ffi::BF16> v_pages, ffi::Buffer<ffi::BF16> key,
                           ffi::Buffer<ffi::BF16> value, ffi::Buffer<ffi::S32> active_page, ffi::Buffer<ffi::S32> active_pos,
                           ffi::ResultBuffer<ffi::BF16> k_out, ffi::ResultBuffer<ffi::BF16> v_out) {
  const auto pd = k_pages.dimensions();
  if (pd.size() != 4) return ffi::Error(ffi::ErrorCode::kInvalidArgument, "pools must be [Hkv, num_pages, tokens_per_page, D]");
  if (k_out->untyped_data() != k_pages.untyped_data() || v_out->untyped_data() != v_pages.untyped_data())
    return ffi::Error(ffi::ErrorCode::kInvalidArgument, "the page pools must be donated (input_output_aliases {0: 0, 1: 1})");
  const int rows = int(active_page.element_count());
  return Status(mtx_paged_append(k_out->untyped_data(), v_out->untyped_data(), key.untyped_data(), value.untyped_data(), active_page.typed_data(),
                                 active_pos.typed_data(), rows, int(pd[0]), int(pd[3]), int(pd[1]), int(pd[2]), stream));
}

// (q [B, Hq*D] or [B, 1, Hq, D], key_pages, value_pages, sequence_lengths [B], page_map [B, max_pages_per_group], scratch u8[...]) -> out like q
ffi::Error PagedAttentionImpl(cudaStream_t stream, ffi::Buffer<ffi::BF16> q, ffi::Buffer<ffi::BF16> k_pages, ffi::Buffer<ffi::BF16> v_pages,
                              ffi::Buffer<ffi::S32> lengths, ffi::Buffer<ffi::S32> page_map, ffi::Buffer<ffi::U8> scratch,
                              ffi::ResultBuffer<ffi::BF16> out, int32_t num_q_heads, float softcap) {
  const auto pd = k_pages.dimensions();
  const auto md = page_map.dimensions();
  if (pd.size() != 4 || md.size() != 2) return ffi::Error(ffi::ErrorCode::kInvalidArgument, "pools must be rank 4, page_map [groups, max_pages]");
  const int rows = int(lengths.element_count());
  const int hkv = int(pd[0]), num_pages = int(pd[1]), tpp = int(pd[2]), d = int(pd[3]), max_pages = int(md[1]);
  if (scratch.element_count() < mtx_paged_attention_scratch_bytes(rows, hkv, num_q_heads, d, max_pages * tpp))
    return ffi::Error(ffi::ErrorCode::kInvalidArgument, "scratch smaller than mtx_paged_attention_scratch_bytes()");
  return Status(mtx_paged_attention(q.untyped_data(), k_pages.untyped_data(), v_pages.untyped_data(), lengths.typed_data(), page_map.typed_data(),
                                    out->untyped_data(), rows, num_q_heads, hkv, d, num_pages, tpp, max_pages, softcap, scratch.untyped_data(),
                                    stream));
}

// (attn [rows padded, Hq*D], wo [E, Hq*D], x [rows, E]) -> out [rows, E];  rows = the leading dimension of x
ffi::Error OutprojResidualImpl(cudaStream_t stream, ffi::Buffer<ffi::BF16> attn, ffi::Buffer<ffi::BF16> wo, ffi::Buffer<ffi::BF16> x,
                               ffi::ResultBuffer<ffi::BF16> out) {
  const auto wd = wo.dimensions();
  const auto xd = x.dimensions();
  if (wd.size() != 2 || xd.size() != 2) return ffi::Error(ffi::ErrorCode::kInvalidArgument, "wo must be [E, Hq*D], x [rows, E]");
  return Status(mtx_outproj_residual(attn.untyped_data(), wo.untyped_data(), x.untyped_data(), out->untyped_data(), int(xd[0]), int(wd[0]),
                                     int(wd[1]), stream));
}

// (h [rows padded, E], norm_scale [E], w01 [2M, E], wout [E, M], scratch u8[...]) -> out [rows, E];  rows = the leading dimension of out
ffi::Error MlpImpl(cudaStream_t stream, ffi::Buffer<ffi::BF16> h, ffi::Buffer<ffi::BF16> norm_scale, ffi::Buffer<ffi::BF16> w01,
                   ffi::Buffer<ffi::BF16> wout, ffi::Buffer<ffi::U8> scratch, ffi::ResultBuffer<ffi::BF16> out, float eps) {
  const auto wd = wout.dimensions();
  const auto od = out->dimensions();
  if (wd.size() != 2 || od.size() != 2) return ffi::Error(ffi::ErrorCode::kInvalidArgument, "wout must be [E, M], out [rows, E]");
  const int rows = int(od[0]), emb = int(wd[0]), mlp = int(wd[1]);
  if (scratch.element_count() < mtx_mlp_scratch_bytes(rows, emb, mlp))
    return ffi::Error(ffi::ErrorCode::kInvalidArgument, "scratch smaller than mtx_mlp_scratch_bytes()");
  return Status(mtx_mlp(h.untyped_data(), norm_scale.untyped_data(), w01.untyped_data(), wout.untyped_data(), out->untyped_data(), rows, emb, mlp,
                        eps, scratch.untyped_data(), stream));
}

}  // namespace

XLA_FFI_DEFINE_HANDLER_SYMBOL(MtxOutprojResidual, OutprojResidualImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::BF16>>()
                                  .Arg<ffi::Buffer<ffi::BF16>>()
                                  .Arg<ffi::Buffer<ffi::BF16>>()
                                  .Ret<ffi::Buffer<ffi::BF16>>());

XLA_FFI_DEFINE_HANDLER_SYMBOL(MtxMlp, MlpImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::BF16>>()
                                  .Arg<ffi::Buffer<ffi::BF16>>()
                                  .Arg<ffi::Buffer<ffi::BF16>>()
                                  .Arg<ffi::Buffer<ffi::BF16>>()
                                  .Arg<ffi::Buffer<ffi::U8>>()
                                  .Ret<ffi::Buffer<ffi::BF16>>()
                                  .Attr<float>("eps"));

XLA_FFI_DEFINE_HANDLER_SYMBOL(MtxPagedAppend, PagedAppendImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::BF16>>()
                                  .Arg<ffi::Buffer<ffi::BF16>>()
                                  .Arg<ffi::Buffer<ffi::BF16>>()
                                  .Arg<ffi::Buffer<ffi::BF16>>()
                                  .Arg<ffi::Buffer<ffi::S32>>()
                                  .Arg<ffi::Buffer<ffi::S32>>()
                                  .Ret<ffi::Buffer<ffi::BF16>>()
                                  .Ret<ffi::Buffer<ffi::BF16>>());

XLA_FFI_DEFINE_HANDLER_SYMBOL(MtxPagedAttention, PagedAttentionImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::BF16>>()
                                  .Arg<ffi::Buffer<ffi::BF16>>()
                                  .Arg<ffi::Buffer<ffi::BF16>>()
                                  .Arg<ffi::Buffer<ffi::S32>>()
                                  .Arg<ffi::Buffer<ffi::S32>>()
                                  .Arg<ffi::Buffer<ffi::U8>>()
                                  .Ret<ffi::Buffer<ffi::BF16>>()
                                  .Attr<int32_t>("num_q_heads")
                                  .Attr<float>("softcap"));

XLA_FFI_DEFINE_HANDLER_SYMBOL(MtxRaggedAttention, RaggedAttentionImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::BF16>>()
                                  .Arg<ffi::Buffer<ffi::BF16>>()
                                  .Arg<ffi::Buffer<ffi::BF16>>()
                                  .Arg<ffi::Buffer<ffi::S32>>()
                                  .Arg<ffi::Buffer<ffi::U8>>()
                                  .Ret<ffi::Buffer<ffi::BF16>>()
                                  .Ret<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>()
                                  .Attr<int32_t>("seq_major")
                                  .Attr<float>("softcap"));

XLA_FFI_DEFINE_HANDLER_SYMBOL(MtxDecodeAttention, DecodeAttentionImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::BF16>>()
                                  .Arg<ffi::Buffer<ffi::BF16>>()
                                  .Arg<ffi::Buffer<ffi::BF16>>()
                                  .Arg<ffi::Buffer<ffi::S32>>()
                                  .Arg<ffi::Buffer<ffi::S32>>()
                                  .Arg<ffi::Buffer<ffi::S32>>()
                                  .Arg<ffi::Buffer<ffi::S32>>()
                                  .Arg<ffi::Buffer<ffi::U8>>()
                                  .Ret<ffi::Buffer<ffi::BF16>>()
                                  .Attr<int32_t>("num_q_heads")
                                  .Attr<int32_t>("max_prefill_len")
                                  .Attr<float>("softcap"));

XLA_FFI_DEFINE_HANDLER_SYMBOL(MtxQkvRopeAppend, QkvRopeAppendImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::BF16>>()
                                  .Arg<ffi::Buffer<ffi::BF16>>()
                                  .Arg<ffi::Buffer<ffi::S32>>()
                                  .Arg<ffi::Buffer<ffi::S32>>()
                                  .Arg<ffi::Buffer<ffi::S32>>()
                                  .Arg<ffi::Buffer<ffi::BF16>>()
                                  .Arg<ffi::Buffer<ffi::BF16>>()
                                  .Arg<ffi::Buffer<ffi::U8>>()
                                  .Ret<ffi::Buffer<ffi::BF16>>()
                                  .Ret<ffi::Buffer<ffi::BF16>>()
                                  .Ret<ffi::Buffer<ffi::BF16>>()
                                  .Attr<int32_t>("num_q_heads")
                                  .Attr<float>("rope_min_timescale")
                                  .Attr<float>("rope_max_timescale"));

XLA_FFI_DEFINE_HANDLER_SYMBOL(MtxDecodeStep, DecodeStepImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::BF16>>()
                                  .Arg<ffi::Buffer<ffi::BF16>>()
                                  .Arg<ffi::Buffer<ffi::S32>>()
                                  .Arg<ffi::Buffer<ffi::S32>>()
                                  .Arg<ffi::Buffer<ffi::S32>>()
                                  .Arg<ffi::Buffer<ffi::S32>>()
                                  .Arg<ffi::Buffer<ffi::S32>>()
                                  .Arg<ffi::Buffer<ffi::S32>>()
                                  .Arg<ffi::Buffer<ffi::U32>>()
                                  .Ret<ffi::Buffer<ffi::BF16>>()
                                  .Ret<ffi::Buffer<ffi::BF16>>()
                                  .Ret<ffi::Buffer<ffi::S32>>()
                                  .Ret<ffi::Buffer<ffi::S32>>()
                                  .Ret<ffi::Buffer<ffi::S32>>()
                                  .Ret<ffi::Buffer<ffi::S32>>()
                                  .Ret<ffi::Buffer<ffi::S32>>()
                                  .Ret<ffi::Buffer<ffi::S32>>()
                                  .Ret<ffi::Buffer<ffi::U32>>()
                                  .Ret<ffi::Buffer<ffi::S32>>()
                                  .Attr<int64_t>("engine")
                                  .Attr<int32_t>("rows"));

extern "C" int mtx_jax_ffi_available(void) { return 1; }

#else  // no jaxlib headers on the include path

extern "C" int mtx_jax_ffi_available(void) { return 0; }

#endif
