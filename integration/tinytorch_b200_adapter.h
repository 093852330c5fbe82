/*
 * tinytorch_b200_adapter.h — the ONLY code that sees tinytorch::Tensor.  Drop this header (and libb200decode.so) next
 * to a TinyGPT checkout built with -DTINYTORCH_USE_CUDA=ON; nothing in TinyGPT/TinyTorch is modified.
 *
 *   Boundary B (per-op):   b200::adapter::registerOps()           — call from main() AFTER static initialisation
 *                          (STATIC_CALL registrars live in the op HEADERS, third_party/TinyTorch/src/Operation/
 *                          OpNNLayer.h:95-101, so a static registrar here could be overwritten by any later TU).
 *   Boundary A (per-token): b200::adapter::ModelB200 wraps a loaded GPTModel; GPTEngine keeps calling
 *                          GPTModel::forward → model()(ids) (src/model/GPTModel.h:86,99) and gets logits [1,S,V].
 *
 * Error convention of the reference: no exceptions; LOGE + ASSERT (third_party/TinyTorch/src/Utils/Macros.h:34-40).
 */
#pragma once

#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "Functions.h"
#include "Modules.h"
#include "Operations.h"
#include "Utils/CUDAUtils.h"
#include "Utils/Logger.h"
#include "b200_decode.h"
#include "model/GPTModel.h"

namespace b200::adapter {

namespace tt = tinytorch;

inline void* currentStream(const tt::Tensor& t) {
  return reinterpret_cast<void*>(tt::cuda::getCurrentCUDAStream(t.device().index).stream());
}

inline void check(int rc, const char* what) {
  if (rc != B200_OK) {
    LOGE("%s failed (%d): %s", what, rc, b200_last_error());
    ASSERT(false);
  }
}

// ---------------------------------------------------------------------------------------------- boundary B: ops
// Each function has exactly the registry's signature (OpLinalg.h:77, OpNNLayer.h:60-66, OpFused.h:14, OpElemWise.h:21)
// and falls back to the implementation that was registered before it for shapes outside the decode path.
struct Previous {
  tt::op::matmulFn matmul = nullptr;
  tt::op::rmsNormFn rmsNorm = nullptr;
  tt::op::ropeApplyFn ropeApply = nullptr;
  tt::op::flashAttentionFn flashAttention = nullptr;
  tt::op::siluMulFn siluMul = nullptr;
  tt::op::addFn add = nullptr;
};
inline Previous& previous() {
  static Previous p;
  return p;
}

inline tt::Tensor matmul(const tt::Tensor& a, const tt::Tensor& b, bool transA, bool transB, const tt::Tensor& bias) {
  // Linear::forward: a = x [B,S,k] (or [m,k]), b = W [n,k], transB = true.  Decode: B·S small.
  const bool linear_like = !transA && transB && b.dim() == 2 && a.dim() >= 2 && a.shape().back() == b.shape(1);
  const int64_t m = linear_like ? a.numel() / a.shape().back() : 0;
  if (!linear_like || m > 16 || (b.shape(1) % 8) != 0 || a.dim() < 3) return previous().matmul(a, b, transA, transB, bias);
  tt::SizeVector shape = a.shape();
  shape.back() = b.shape(0);
  tt::Tensor y = tt::Tensor::empty(shape, a.options().noGrad());
  check(b200_gemv_bf16(y.dataPtr<tt::BFloat16>(), a.dataPtr<tt::BFloat16>(), b.dataPtr<tt::BFloat16>(),
                       bias.defined() ? bias.dataPtr<tt::BFloat16>() : nullptr, m, b.shape(0), b.shape(1),
                       currentStream(a)),
        "b200_gemv_bf16");
  return y;
}

inline tt::Tensor rmsNorm(const tt::Tensor& self, tt::IntArrayView normalizedShape, const tt::Tensor& weight, float eps) {
  ASSERT(normalizedShape.size() == 1 && normalizedShape.front() == self.shape().back());
  tt::Tensor y = tt::Tensor::empty(self.shape(), self.options().noGrad());
  const int64_t dim = self.shape().back();
  check(b200_rmsnorm_bf16(y.dataPtr<tt::BFloat16>(), self.dataPtr<tt::BFloat16>(),
                          weight.defined() ? weight.dataPtr<tt::BFloat16>() : nullptr, self.numel() / dim, dim, eps,
                          currentStream(self)),
        "b200_rmsnorm_bf16");
  return y;
}

inline tt::Tensor ropeApply(const tt::Tensor& input, const tt::Tensor& rope, int64_t offset, tt::QKVLayout layout) {
  ASSERT(input.dim() == 4);
  const bool bshd = layout == tt::QKVLayout::BSHD;
  const int64_t B = input.shape(0), S = input.shape(bshd ? 1 : 2), N = input.shape(bshd ? 2 : 1), D = input.shape(3);
  tt::Tensor y = tt::Tensor::empty(input.shape(), input.options().noGrad());
  check(b200_rope_bf16(y.dataPtr<tt::BFloat16>(), input.dataPtr<tt::BFloat16>(), rope.dataPtr<float>(), B, S, N, D,
                       offset, bshd ? B200_LAYOUT_BSHD : B200_LAYOUT_BHSD, currentStream(input)),
        "b200_rope_bf16");
  return y;
}

inline tt::Tensor flashAttention(const tt::Tensor& q, const tt::Tensor& k, const tt::Tensor& v, bool isCausal) {
  const int64_t hd = q.shape(3);
  if (hd != 64 && hd != 128) return previous().flashAttention(q, k, v, isCausal);
  tt::Tensor o = tt::Tensor::empty(q.shape(), q.options().noGrad());
  check(b200_attn_bf16(o.dataPtr<tt::BFloat16>(), q.dataPtr<tt::BFloat16>(), k.dataPtr<tt::BFloat16>(),
                       v.dataPtr<tt::BFloat16>(), q.shape(0), q.shape(1), k.shape(1), q.shape(2), k.shape(2), hd,
                       isCausal ? 1 : 0, currentStream(q)),
        "b200_attn_bf16");
  return o;
}

inline tt::Tensor siluMul(const tt::Tensor& self) {
  const int64_t I = self.shape().back() / 2;
  tt::SizeVector shape = self.shape();
  shape.back() = I;
  tt::Tensor y = tt::Tensor::empty(shape, self.options().noGrad());
  check(b200_silu_mul_bf16(y.dataPtr<tt::BFloat16>(), self.dataPtr<tt::BFloat16>(), self.numel() / (2 * I), I,
                           currentStream(self)),
        "b200_silu_mul_bf16");
  return y;
}

inline tt::Tensor add(const tt::Tensor& a, const tt::Tensor& b, const tt::Scalar& alpha) {
  if (a.shape() != b.shape() || alpha.to<float>() != 1.f) return previous().add(a, b, alpha);
  tt::Tensor y = tt::Tensor::empty(a.shape(), a.options().noGrad());
  check(b200_add_bf16(y.dataPtr<tt::BFloat16>(), a.dataPtr<tt::BFloat16>(), b.dataPtr<tt::BFloat16>(), a.numel(),
                      currentStream(a)),
        "b200_add_bf16");
  return y;
}

// Call from main() (after static init, before the first forward).
inline void registerOps() {
  const tt::DispatchKey key{tt::DeviceType::CUDA, tt::DType::BFloat16};
  Previous& p = previous();
  p.matmul = tt::op::matmulRegistry::lookup(key);
  p.rmsNorm = tt::op::rmsNormRegistry::lookup(key);
  p.ropeApply = tt::op::ropeApplyRegistry::lookup(key);
  p.flashAttention = tt::op::flashAttentionRegistry::lookup(key);
  p.siluMul = tt::op::siluMulRegistry::lookup(key);
  p.add = tt::op::addRegistry::lookup(key);
  tt::op::matmulRegistry::registerImpl(key, &matmul);
  tt::op::rmsNormRegistry::registerImpl(key, &rmsNorm);
  tt::op::ropeApplyRegistry::registerImpl(key, &ropeApply);
  tt::op::flashAttentionRegistry::registerImpl(key, &flashAttention);
  tt::op::siluMulRegistry::registerImpl(key, &siluMul);
  tt::op::addRegistry::registerImpl(key, &add);
}

// ------------------------------------------------------------------------------------- boundary A: whole token
// A Module that owns ONE b200_engine built from the weights the reference's loader already placed on the device.  A
// batch of B ≤ 8 left-padded prompts (GPTEngine::generateSync, src/engine/GPTEngine.cpp:154-174) goes through the
// engine's batched step: B KV caches, ONE weight pass per decode step (b200_engine_forward with B > 1).
class B200CausalLM : public tt::nn::Module {
 public:
  // `loaded` is the reference model after ModelLoader::load (src/huggingface/ModelLoader.cpp:25-87).
  // `rope` is a RoPE module built the way the model's createModel() builds each layer's (src/model/ModelLlama.h:40-43,
  // ModelQwen2.h:34, ModelQwen3.h, ModelMistral.h): Attention keeps its own rope_ as an UNREGISTERED member
  // (src/layer/Attention.h:61-68 registers q/k/v/o_proj only), so the table is not among namedStates(); building one
  // more RoPE runs the reference's own op::ropeInit kernels and gives the identical fp32 table, owned here.
  B200CausalLM(tt::nn::Module& loaded, const b200_model_desc& desc, tinygpt::KVCacheManager* refCache, tt::nn::RoPE&& rope)
      : desc_(desc), refCache_(refCache), layers_(desc.layers), rope_(std::move(rope)) {
    ASSERT(rope_.cache().defined() && rope_.cache().dim() == 3 && rope_.cache().shape(1) == desc.head_dim);
    ASSERT(rope_.cache().device().type == tt::DeviceType::CUDA && rope_.cache().dtype() == tt::DType::Float32);
    if (desc_.max_ctx > rope_.cache().shape(0)) desc_.max_ctx = (int32_t)rope_.cache().shape(0);
    table_.rope_table = rope_.cache().dataPtr<float>();
    for (auto& [name, t] : loaded.namedStates()) {  // third_party/TinyTorch/src/Module/Module.h:43-53
      void* p = t->dataPtr<tt::BFloat16>();
      auto ends = [&](const char* s) {
        return name.size() >= strlen(s) && name.compare(name.size() - strlen(s), strlen(s), s) == 0;
      };
      if (name == "model.embed_tokens.weight") table_.embed = p;
      else if (name == "model.norm.weight") table_.final_norm = p;
      else if (name == "lm_head.weight") table_.lm_head = p;
      else if (name.rfind("model.layers.", 0) == 0) {
        const int l = std::stoi(name.substr(13));
        b200_layer_weights& w = layers_[l];
        // q/k/v and gate/up are dim-0 views of ONE merged allocation (src/layer/Linear.h:64-79): the q / gate view's
        // pointer is the merged matrix.
        if (ends("input_layernorm.weight")) w.input_norm = p;
        else if (ends("self_attn.q_proj.weight")) w.qkv_w = p;
        else if (ends("self_attn.q_proj.bias")) w.qkv_b = p;
        else if (ends("self_attn.q_norm.weight")) w.q_norm = p;
        else if (ends("self_attn.k_norm.weight")) w.k_norm = p;
        else if (ends("self_attn.o_proj.weight")) w.o_w = p;
        else if (ends("post_attention_layernorm.weight")) w.post_norm = p;
        else if (ends("mlp.gate_proj.weight")) w.gate_up_w = p;
        else if (ends("mlp.down_proj.weight")) w.down_w = p;
      }
    }
    if (table_.lm_head == nullptr) table_.lm_head = table_.embed;  // tie_word_embeddings (GPTModel.h:39-41)
    table_.layers_host = layers_.data();
  }
  ~B200CausalLM() override {
    for (b200_engine* e : engines_) b200_engine_destroy(e);
  }

  // ids [B,S] Int64 on the device → logits [B,S,V] bf16 (only the last row of each sequence is computed;
  // GPTEngine::genNextToken narrows to it, src/engine/GPTEngine.cpp:96).  A call with S > 1, or the first call after
  // GPTModel::resetCache() (non-virtual; it only clears the reference's now-unused KVCacheManager, which this module
  // re-marks with a 1-row placeholder so the next reset is visible), starts new sequences.
  tt::Tensor forward(const tt::Tensor& ids) override {
    ASSERT(ids.dim() == 2 && ids.dtype() == tt::DType::Int64);
    void* stream = currentStream(ids);
    const int64_t B = ids.shape(0), S = ids.shape(1), V = desc_.vocab;
    if (engines_.empty()) {
      b200_engine* e = nullptr;
      check(b200_engine_create(&desc_, &table_, &e), "b200_engine_create");
      engines_.push_back(e);
    }
    const bool fresh = S > 1 || (refCache_ != nullptr && refCache_->pastLength(0, 1) == 0);
    tt::Tensor logits = tt::Tensor::empty({B, S, V}, tt::Options(ids.device(), tt::DType::BFloat16).noGrad());
    if (fresh) check(b200_engine_reset(engines_[0], stream), "b200_engine_reset");
    if (S == 1) {   // decode step: [B, 1, V] is exactly the engine's last-position layout
      check(b200_engine_forward(engines_[0], ids.dataPtr<int64_t>(), B, 1, logits.dataPtr<tt::BFloat16>(), /*logits_mode=*/0,
                                stream),
            "b200_engine_forward");
    } else {        // prompt: only the last row of every sequence is computed (genNextToken narrows to it)
      tt::Tensor last = tt::Tensor::empty({B, V}, tt::Options(ids.device(), tt::DType::BFloat16).noGrad());
      check(b200_engine_forward(engines_[0], ids.dataPtr<int64_t>(), B, S, last.dataPtr<tt::BFloat16>(), /*logits_mode=*/0,
                                stream),
            "b200_engine_forward");
      for (int64_t b = 0; b < B; ++b)
        cudaMemcpyAsync(logits.dataPtr<tt::BFloat16>() + (b * S + (S - 1)) * V, last.dataPtr<tt::BFloat16>() + b * V,
                        sizeof(tt::BFloat16) * V, cudaMemcpyDeviceToDevice, reinterpret_cast<cudaStream_t>(stream));
    }
    if (fresh && refCache_ != nullptr) {  // leave a mark so that the next resetCache() can be told from a decode step
      tt::Tensor mark = tt::Tensor::empty({1, 1, 1, 1}, tt::Options(ids.device(), tt::DType::BFloat16).noGrad());
      refCache_->append(0, {mark, mark}, 1);
    }
    return logits;
  }

 private:
  b200_model_desc desc_;
  tinygpt::KVCacheManager* refCache_;
  std::vector<b200_layer_weights> layers_;
  tt::nn::RoPE rope_;
  b200_weight_table table_{};
  std::vector<b200_engine*> engines_;
};

// GPTModel whose model() is the fused module; everything else is delegated to the loaded reference model.
class ModelB200 : public tinygpt::GPTModel {
 public:
  ModelB200(std::unique_ptr<tinygpt::GPTModel> loaded, const b200_model_desc& desc, tt::nn::RoPE&& rope)
      : loaded_(std::move(loaded)), fused_(loaded_->model(), desc, &kvCache_, std::move(rope)) {
    init();
  }
  tinygpt::GPTModelType type() override { return loaded_->type(); }
  int64_t numLayers() override { return loaded_->numLayers(); }
  int64_t contextSize() override { return loaded_->contextSize(); }
  tt::nn::Module& model() override { return fused_; }
  tt::Device device() const override { return loaded_->device(); }
  bool load(const std::string& path) override { return loaded_->load(path); }

 private:
  std::unique_ptr<tinygpt::GPTModel> loaded_;
  B200CausalLM fused_;
};

}  // namespace b200::adapter
