// ref_harness.cpp — TEST INFRASTRUCTURE.  Thin C entry points over the UNMODIFIED reference (keith2018/TinyGPT,
// compiled from /root/reference by oracle/Makefile) so that Python can run the reference's own CPU implementation of
// the decode path and compare it with oracle/decode_oracle.py.  Nothing in the product links or loads this.
//
// The reference has no CPU flashAttention (third_party/TinyTorch/src/Operation/OpNNLayerCpu.cpp:16-40 registers every
// NN op except it), so a Llama-family forward traps on the CPU.  ref_register_attention_shim() registers, through the
// reference's own op registry (third_party/TinyTorch/src/Tensor/Dispatch.h:38-48), a naive fp32 attention written
// here after the loop structure of TinyFA's test oracle (tests/cpp/cpu_reference.h:14-66).  Everything else that runs
// is the reference's code: modules, ops, KV cache, model wiring.
#include <cmath>
#include <cstdint>
#include <cstring>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "Functions.h"
#include "Modules.h"
#include "model/ModelGPT2.h"
#include "model/ModelLlama.h"
#include "model/ModelMistral.h"
#include "model/ModelQwen2.h"
#include "model/ModelQwen3.h"
#include "engine/Sampler.h"

namespace tt = tinytorch;

namespace {

template <typename T>
tt::Tensor attentionShim(const tt::Tensor& q, const tt::Tensor& k, const tt::Tensor& v, bool isCausal) {
  // BSHD, GQA; fp32 math whatever the storage type; output in the storage type.
  const int64_t B = q.shape(0), Sq = q.shape(1), Hq = q.shape(2), D = q.shape(3);
  const int64_t Skv = k.shape(1), Hkv = k.shape(2);
  const int64_t G = Hq / Hkv;
  const float scale = 1.0f / std::sqrt(static_cast<float>(D));
  tt::Tensor out = tt::Tensor::empty(q.shape(), q.options().noGrad());
  const T* Q = q.dataPtr<T>();
  const T* K = k.dataPtr<T>();
  const T* V = v.dataPtr<T>();
  T* O = out.dataPtr<T>();
  std::vector<float> sc(Skv);
  for (int64_t b = 0; b < B; b++)
    for (int64_t h = 0; h < Hq; h++) {
      const int64_t hk = h / G;
      for (int64_t i = 0; i < Sq; i++) {
        const int64_t n = isCausal ? std::min<int64_t>(Skv, i + 1) : Skv;
        float mx = -INFINITY;
        for (int64_t j = 0; j < n; j++) {
          float dot = 0.f;
          for (int64_t d = 0; d < D; d++)
            dot += static_cast<float>(Q[((b * Sq + i) * Hq + h) * D + d]) *
                   static_cast<float>(K[((b * Skv + j) * Hkv + hk) * D + d]);
          sc[j] = dot * scale;
          mx = std::max(mx, sc[j]);
        }
        float sum = 0.f;
        for (int64_t j = 0; j < n; j++) {
          sc[j] = std::exp(sc[j] - mx);
          sum += sc[j];
        }
        for (int64_t d = 0; d < D; d++) {
          float acc = 0.f;
          for (int64_t j = 0; j < n; j++) acc += sc[j] * static_cast<float>(V[((b * Skv + j) * Hkv + hk) * D + d]);
          O[((b * Sq + i) * Hq + h) * D + d] = static_cast<T>(acc / sum);
        }
      }
    }
  return out;
}

tt::Tensor fromF32(const float* p, std::initializer_list<int64_t> shape) {
  tt::SizeVector sv(shape);
  tt::Tensor t = tt::Tensor::empty(sv, tt::Options(tt::Device(tt::DeviceType::CPU), tt::DType::Float32));
  std::memcpy(t.dataPtr<float>(), p, sizeof(float) * t.numel());
  return t;
}

void toF32(const tt::Tensor& t, float* out) {
  tt::Tensor f = t.dtype() == tt::DType::Float32 ? t : t.to(tt::DType::Float32);
  std::memcpy(out, f.dataPtr<float>(), sizeof(float) * f.numel());
}

}  // namespace

extern "C" {

void ref_register_attention_shim() {
  tt::op::flashAttentionRegistry::registerImpl({tt::DeviceType::CPU, tt::DType::Float32}, &attentionShim<float>);
  tt::op::flashAttentionRegistry::registerImpl({tt::DeviceType::CPU, tt::DType::BFloat16},
                                               &attentionShim<tt::BFloat16>);
}

// ---- per-op entry points, fp32 on the CPU, straight through tinytorch::function::*
void ref_rmsnorm_f32(const float* x, const float* w, float eps, int64_t rows, int64_t dim, float* out) {
  tt::NoGradGuard g;
  auto y = tt::function::rmsNorm(fromF32(x, {rows, dim}), {dim}, w ? fromF32(w, {dim}) : tt::Tensor(), eps);
  toF32(y, out);
}

void ref_rope_table_f32(int64_t hd, int64_t ctx, float theta, float factor, float high, float low, int64_t orig,
                        float* out) {
  tt::NoGradGuard g;
  std::optional<tt::RopeScalingConfig> sc;
  if (factor != 0.f) sc = tt::RopeScalingConfig{factor, high, low, orig};
  tt::nn::RoPE rope(hd, ctx, theta, sc);
  toF32(rope.cache(), out);
}

void ref_rope_apply_f32(const float* x, int64_t d0, int64_t d1, int64_t d2, int64_t d3, int bshd, int64_t hd,
                        int64_t ctx, float theta, int64_t offset, float* out) {
  tt::NoGradGuard g;
  tt::nn::RoPE rope(hd, ctx, theta);
  auto y = rope(fromF32(x, {d0, d1, d2, d3}), offset, bshd ? tt::QKVLayout::BSHD : tt::QKVLayout::BHSD);
  toF32(y, out);
}

void ref_linear_f32(const float* x, const float* W, const float* bias, int64_t b, int64_t s, int64_t n, int64_t k,
                    float* out) {
  tt::NoGradGuard g;
  auto y = tt::function::linear(fromF32(x, {b, s, k}), fromF32(W, {n, k}), bias ? fromF32(bias, {n}) : tt::Tensor());
  toF32(y, out);
}

void ref_silu_mul_f32(const float* gu, int64_t rows, int64_t I, float* out) {
  tt::NoGradGuard g;
  toF32(tt::function::siluMul(fromF32(gu, {rows, 2 * I})), out);
}

void ref_add_f32(const float* a, const float* b, int64_t n, float* out) {
  tt::NoGradGuard g;
  toF32(fromF32(a, {n}) + fromF32(b, {n}), out);
}

void ref_argmax_f32(const float* x, int64_t rows, int64_t V, int64_t* out) {
  tt::NoGradGuard g;
  auto idx = tt::function::argmax(fromF32(x, {rows, V}), -1, true);
  std::memcpy(out, idx.dataPtr<int64_t>(), sizeof(int64_t) * rows);
}

// ---- sampler: the reference's own Sampler::sample (src/engine/Sampler.cpp, compiled unmodified) with ONE hook: the
// CPU `multinomial` op is replaced, through the reference's registry, by a shim that records the probability vector
// the sampler hands it and draws by inverse CDF from a caller-supplied uniform number — so the filtering pipeline
// (temperature, top-k, top-p, min-p) that runs is the reference's, and the draw is reproducible.
static std::vector<float>* g_capturedProbs = nullptr;
static float g_uniform = 0.f;

static tt::Tensor multinomialShim(const tt::Tensor& probs, int64_t nSamples, bool /*replacement*/) {
  ASSERT(nSamples == 1 && probs.dim() == 2 && probs.shape(0) == 1);
  const int64_t n = probs.shape(1);
  tt::Tensor f = probs.dtype() == tt::DType::Float32 ? probs : probs.to(tt::DType::Float32);
  const float* p = f.dataPtr<float>();
  if (g_capturedProbs) g_capturedProbs->assign(p, p + n);
  // kMultinomialWithReplacement's rule (third_party/TinyTorch/src/Operation/OpSamplingCuda.cu:37-62): inclusive fp32
  // cumulative sum in index order, r = u * total, first index whose cdf >= r
  float total = 0.f;
  for (int64_t i = 0; i < n; i++) total += p[i];
  const float r = g_uniform * total;
  float c = 0.f;
  int64_t pick = 0;
  for (int64_t i = 0; i < n; i++) {
    c += p[i];
    if (c >= r) {
      pick = i;
      break;
    }
  }
  tt::Tensor out = tt::Tensor::empty({1, 1}, tt::Options(tt::Device(tt::DeviceType::CPU), tt::DType::Int64));
  out.dataPtr<int64_t>()[0] = pick;
  return out;
}

// logits [V] fp32 → probabilities after the reference's filtering [V] fp32, and the index drawn with uniform u.
int64_t ref_sampler_f32(const float* logits, int64_t V, float temperature, int64_t topK, float topP, float minP,
                        float u, float* probs_out) {
  tt::NoGradGuard g;
  tt::op::multinomialRegistry::registerImpl({tt::DeviceType::CPU, tt::DType::Float32}, &multinomialShim);
  std::vector<float> captured;
  g_capturedProbs = &captured;
  g_uniform = u;
  tinygpt::Sampler sampler(tinygpt::SamplerConfig(temperature, topK, topP, minP));
  tt::Tensor idx = sampler.sample(fromF32(logits, {1, V}));
  g_capturedProbs = nullptr;
  if (probs_out && (int64_t)captured.size() == V) std::memcpy(probs_out, captured.data(), sizeof(float) * V);
  if (captured.empty() && probs_out) std::memset(probs_out, 0, sizeof(float) * V);  // greedy branch: no multinomial
  return idx.to(tt::DType::Int64).dataPtr<int64_t>()[0];
}

// ---- whole model ---------------------------------------------------------------------------------------------
struct RefModelDesc {
  int32_t family;  // 0 llama, 1 qwen2, 2 qwen3, 3 mistral, 4 gpt2 (hidden = n_embd, q_heads = n_head, max_ctx = n_positions)
  int32_t hidden, layers, q_heads, kv_heads, head_dim, intermediate, vocab, max_ctx;
  float rope_theta, rms_eps;
  int32_t tie;
  float rs_factor, rs_high, rs_low;
  int32_t rs_orig;
  int32_t bf16;  // 0: Float32 model, 1: BFloat16 model (the reference's CPU bf16 arithmetic)
};

struct RefModel {
  tinygpt::huggingface::model::LlamaConfig llama;
  tinygpt::huggingface::model::QwenConfig qwen;
  tinygpt::huggingface::model::MistralConfig mistral;
  tinygpt::huggingface::model::GPT2Config gpt2;
  std::unique_ptr<tinygpt::GPTModel> model;
};

static void fillCommon(tinygpt::huggingface::model::ModelConfig& c, const RefModelDesc& d) {
  c.torchDtype = d.bf16 ? tt::DType::BFloat16 : tt::DType::Float32;
  c.vocabSize = d.vocab;
  c.hiddenSize = d.hidden;
  c.intermediateSize = d.intermediate;
  c.maxPositionEmbeddings = d.max_ctx;
  c.numAttentionHeads = d.q_heads;
  c.numHiddenLayers = d.layers;
  c.numKeyValueHeads = d.kv_heads;
  c.rmsNormEps = d.rms_eps;
  c.tieWordEmbeddings = d.tie != 0;
  c.bosTokenId = 0;
  c.eosTokenId = 0;
}

void* ref_model_create(const RefModelDesc* d) {
  ref_register_attention_shim();
  auto* m = new RefModel();
  tt::Device cpu(tt::DeviceType::CPU);
  switch (d->family) {
    case 0:
      fillCommon(m->llama, *d);
      m->llama.headDim = d->head_dim;
      m->llama.attentionBias = false;
      m->llama.ropeTheta = d->rope_theta;
      m->llama.ropeScaling = {d->rs_factor, d->rs_high, d->rs_low, d->rs_orig, "llama3"};
      m->model = std::make_unique<tinygpt::ModelLlama>(m->llama, cpu);
      break;
    case 1:
    case 2:
      fillCommon(m->qwen, *d);
      m->qwen.headDim = d->head_dim;
      m->qwen.ropeTheta = d->rope_theta;
      m->qwen.slidingWindow = 0;
      m->qwen.useSlidingWindow = false;
      m->qwen.useMRope = false;
      if (d->family == 1)
        m->model = std::make_unique<tinygpt::ModelQwen2>(m->qwen, cpu);
      else
        m->model = std::make_unique<tinygpt::ModelQwen3>(m->qwen, cpu);
      break;
    case 4:   // BASELINE config 1: GPT-2, the only family whose forward exists on the reference's CPU path unmodified
      fillCommon(m->gpt2, *d);
      m->gpt2.activationFunction = "gelu_new";
      m->gpt2.layerNormEpsilon = d->rms_eps;
      m->gpt2.nCtx = d->max_ctx;
      m->gpt2.nEmbd = d->hidden;
      m->gpt2.nHead = d->q_heads;
      m->gpt2.nLayer = d->layers;
      m->gpt2.nPositions = d->max_ctx;
      m->model = std::make_unique<tinygpt::ModelGPT2>(m->gpt2, cpu);
      break;
    default:
      fillCommon(m->mistral, *d);
      m->mistral.ropeTheta = d->rope_theta;
      m->mistral.slidingWindow = 0;
      m->mistral.useSlidingWindow = false;
      m->model = std::make_unique<tinygpt::ModelMistral>(m->mistral, cpu);
      break;
  }
  m->model->model().eval();
  return m;
}

void ref_model_destroy(void* h) { delete static_cast<RefModel*>(h); }

// Number of named states and their names/sizes (so the caller can supply data by HF name).
int64_t ref_model_num_states(void* h) { return (int64_t)static_cast<RefModel*>(h)->model->model().namedStates().size(); }

int64_t ref_model_state_info(void* h, int64_t i, char* name_out, int64_t name_cap) {
  auto st = static_cast<RefModel*>(h)->model->model().namedStates();
  std::strncpy(name_out, st[i].first.c_str(), name_cap - 1);
  name_out[name_cap - 1] = 0;
  return st[i].second->numel();
}

// Copy fp32 data into state i (converted to the model dtype by the reference's own cast).
void ref_model_set_state(void* h, int64_t i, const float* data) {
  auto st = static_cast<RefModel*>(h)->model->model().namedStates();
  tt::Tensor& t = *st[i].second;
  if (t.dtype() == tt::DType::Float32) {
    std::memcpy(t.dataPtr<float>(), data, sizeof(float) * t.numel());
  } else if (t.dtype() == tt::DType::BFloat16) {
    auto* p = t.dataPtr<tt::BFloat16>();
    for (int64_t j = 0; j < t.numel(); j++) p[j] = tt::BFloat16(data[j]);
  }
}

void ref_model_reset(void* h) { static_cast<RefModel*>(h)->model->resetCache(); }

// GPTModel::forward(ids[1,S]) → logits of the LAST position as fp32 [V]  (genNextToken's narrow, GPTEngine.cpp:94-99)
void ref_model_forward(void* h, const int64_t* ids, int64_t S, float* logits_last) {
  tt::NoGradGuard g;
  auto* m = static_cast<RefModel*>(h);
  tt::Tensor t = tt::Tensor::empty({1, S}, tt::Options(tt::Device(tt::DeviceType::CPU), tt::DType::Int64));
  std::memcpy(t.dataPtr<int64_t>(), ids, sizeof(int64_t) * S);
  tt::Tensor logits = m->model->forward(t);
  tt::Tensor last = tt::function::narrow(logits, 1, S - 1, 1).squeeze(1);
  toF32(last, logits_last);
}

}  // extern "C"
