// ref_cuda_main.cpp — TEST INFRASTRUCTURE.  Runs the UNMODIFIED reference's CUDA path (keith2018/TinyGPT: its modules,
// TinyTorch CUDA ops, cuBLAS GEMMs, TinyFA attention, its KV-cache manager and its own argmax), compiled from
// /root/reference by `make -C oracle cuda`, on a checkpoint directory written by tinygpt_b200.models.save_checkpoint:
// the bf16 parity oracle north_star names ("logits matching the reference CUDA path … bit-exact argmax ids") and the
// reference's own decode speed on the same GPU.  Nothing in the product links or loads this.
//
//   ref_cuda_decode --ckpt DIR --family {llama|qwen2|qwen3|mistral} --dims H,L,Hq,Hkv,hd,I,V,ctx --theta T --eps E
//                   --tie {0|1} [--rope-scaling f,hi,lo,orig] --ids FILE(int64) --new N [--forced FILE(int64)]
//                   --out FILE [--time-steps K] [--b200 {off|ops|engine}] [--qkv-bias 0|1] [--qk-norm 0|1]
//                   [--dump-rope FILE] [--batched 1] [--batch B]
// --batch B: the ids file holds B equally long prompts (row-major [B, S]); every step runs the model on [B, n] ids like
// GPTEngine::generateSync does for its left-padded batch (src/engine/GPTEngine.cpp:154-174).  Output: int64 [N][B]
// tokens, then float32 [N][B][V] logits.  (--forced then holds [N][B] tokens.)
// --dump-rope writes the fp32 cos/sin table of a RoPE module built as the family's createModel() builds it.
// --batched 1 (needs --forced): ONE forward over prompt + forced tokens — the reference's batched path (cuBLAS GEMM with
// m = S + N - 1, causal TinyFA) — and the logits of the same N positions: the reference against ITSELF through two of
// its own code paths, i.e. the summation-order noise floor of its CUDA arithmetic.
// --b200 exercises the DROP-IN BOUNDARY with the real reference code around it (integration/tinytorch_b200_adapter.h,
// linked against tinygpt_b200/lib/libb200decode.so):  ops    = b200::adapter::registerOps() — boundary B, the
// reference's modules call our kernels through its own op registry;  engine = the loaded GPTModel wrapped in
// b200::adapter::ModelB200 — boundary A, GPTModel::forward → virtual model() → b200_engine_forward.  Everything else
// (loader, tensors, generate loop, argmax) stays the reference's.
// Output file: int64 N tokens, then float32 [N, V] logits (last position of the prefill, then every decode step).
// With --forced the decode steps consume the given tokens (teacher forcing) instead of the reference's own argmax.
// The loop is GPTEngine::generateSync's (src/engine/GPTEngine.cpp:154-174): genNextToken = forward → narrow last →
// Sampler greedy = function::argmax(logits, -1, true) (src/engine/Sampler.cpp:23-29).
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include <unistd.h>

#include <cuda_runtime.h>

#include "Functions.h"
#include "Modules.h"
#include "model/ModelLlama.h"
#include "model/ModelMistral.h"
#include "model/ModelQwen2.h"
#include "model/ModelQwen3.h"
#include "tinytorch_b200_adapter.h"

namespace tt = tinytorch;
namespace hf = tinygpt::huggingface::model;

static std::vector<int64_t> readI64(const std::string& path) {
  std::vector<int64_t> v;
  FILE* f = fopen(path.c_str(), "rb");
  if (!f) return v;
  fseek(f, 0, SEEK_END);
  const long n = ftell(f);
  fseek(f, 0, SEEK_SET);
  v.resize(n / 8);
  if (fread(v.data(), 8, v.size(), f) != v.size()) v.clear();
  fclose(f);
  return v;
}

static std::vector<double> parseList(const char* s) {
  std::vector<double> v;
  const char* p = s;
  while (*p) {
    char* e;
    v.push_back(strtod(p, &e));
    if (e == p) break;
    p = (*e == ',') ? e + 1 : e;
  }
  return v;
}

static void fillCommon(hf::ModelConfig& c, const std::vector<double>& d, float eps, bool tie) {
  c.torchDtype = tt::DType::BFloat16;
  c.hiddenSize = (int64_t)d[0];
  c.numHiddenLayers = (int64_t)d[1];
  c.numAttentionHeads = (int64_t)d[2];
  c.numKeyValueHeads = (int64_t)d[3];
  c.intermediateSize = (int64_t)d[5];
  c.vocabSize = (int64_t)d[6];
  c.maxPositionEmbeddings = (int64_t)d[7];
  c.rmsNormEps = eps;
  c.tieWordEmbeddings = tie;
  c.bosTokenId = 0;
  c.eosTokenId = 0;
}

int main(int argc, char** argv) {
  std::string ckpt, family, idsPath, forcedPath, outPath;
  std::vector<double> dims, rs;
  float theta = 10000.f, eps = 1e-5f;
  int tie = 1, nNew = 8, timeSteps = 0, qkvBias = -1, qkNorm = -1;
  std::string b200Mode = "off", ropeDump;
  int batched = 0, batch = 1;
  for (int i = 1; i + 1 < argc; i += 2) {
    const std::string a = argv[i];
    const char* v = argv[i + 1];
    if (a == "--ckpt") ckpt = v;
    else if (a == "--family") family = v;
    else if (a == "--dims") dims = parseList(v);
    else if (a == "--theta") theta = (float)atof(v);
    else if (a == "--eps") eps = (float)atof(v);
    else if (a == "--tie") tie = atoi(v);
    else if (a == "--rope-scaling") rs = parseList(v);
    else if (a == "--ids") idsPath = v;
    else if (a == "--forced") forcedPath = v;
    else if (a == "--new") nNew = atoi(v);
    else if (a == "--out") outPath = v;
    else if (a == "--time-steps") timeSteps = atoi(v);
    else if (a == "--b200") b200Mode = v;
    else if (a == "--qkv-bias") qkvBias = atoi(v);
    else if (a == "--qk-norm") qkNorm = atoi(v);
    else if (a == "--dump-rope") ropeDump = v;
    else if (a == "--batched") batched = atoi(v);
    else if (a == "--batch") batch = atoi(v);
  }
  if (ckpt.empty() || dims.size() != 8 || idsPath.empty() || outPath.empty()) {
    fprintf(stderr, "usage: see the header of oracle/ref_cuda_main.cpp\n");
    return 2;
  }
  tt::NoGradGuard guard;
  tt::Device dev(tt::DeviceType::CUDA, 0);
  hf::LlamaConfig llama{};
  hf::QwenConfig qwen{};
  hf::MistralConfig mistral{};
  std::unique_ptr<tinygpt::GPTModel> model;
  if (family == "llama") {
    fillCommon(llama, dims, eps, tie != 0);
    llama.headDim = (int64_t)dims[4];
    llama.attentionBias = false;
    llama.ropeTheta = theta;
    if (rs.size() == 4) llama.ropeScaling = {(float)rs[0], (float)rs[1], (float)rs[2], (int64_t)rs[3], "llama3"};
    model = std::make_unique<tinygpt::ModelLlama>(llama, dev);
  } else if (family == "qwen2" || family == "qwen3") {
    fillCommon(qwen, dims, eps, tie != 0);
    qwen.headDim = (int64_t)dims[4];
    qwen.ropeTheta = theta;
    qwen.slidingWindow = 0;
    qwen.useSlidingWindow = false;
    qwen.useMRope = false;
    if (family == "qwen2") model = std::make_unique<tinygpt::ModelQwen2>(qwen, dev);
    else model = std::make_unique<tinygpt::ModelQwen3>(qwen, dev);
  } else if (family == "mistral") {
    fillCommon(mistral, dims, eps, tie != 0);
    mistral.ropeTheta = theta;
    mistral.slidingWindow = 0;
    mistral.useSlidingWindow = false;
    model = std::make_unique<tinygpt::ModelMistral>(mistral, dev);
  } else {
    fprintf(stderr, "unknown family %s\n", family.c_str());
    return 2;
  }
  // the reference's loader path: SafeTensors::load(model(), path, strict = false), then to(dtype), eval()
  // (src/model/GPTModel.h:96, src/huggingface/ModelLoader.cpp:70-87)
  if (!model->load(ckpt + "/model.safetensors")) {
    fprintf(stderr, "reference loader failed on %s\n", ckpt.c_str());
    return 3;
  }
  model->model().to(tt::DType::BFloat16);
  model->model().eval();
  auto makeRope = [&]() {
    // built exactly as the family's createModel() builds each layer's (ModelLlama.h:40-43 — always with a scaling
    // config and with the ORIGINAL context when the config carries one; ModelQwen2/3.h, ModelMistral.h: no scaling,
    // max_position_embeddings rows)
    const tt::Options ropeOpts(dev, tt::DType::BFloat16);
    std::optional<tt::RopeScalingConfig> scaling;
    int64_t ropeCtx = (int64_t)dims[7];
    if (family == "llama") {
      scaling = tinygpt::llama::convertToRopeScalingConfig(llama);
      ropeCtx = tinygpt::llama::getContextSize(llama);
    }
    return tt::nn::RoPE((int64_t)dims[4], ropeCtx, theta, scaling, ropeOpts);
  };
  if (!ropeDump.empty()) {
    tt::nn::RoPE rope = makeRope();
    tt::Tensor t = rope.cache().to(tt::Device(tt::DeviceType::CPU));
    FILE* rf = fopen(ropeDump.c_str(), "wb");
    if (!rf) return 4;
    const int64_t shp[3] = {t.shape(0), t.shape(1), t.shape(2)};
    fwrite(shp, 8, 3, rf);
    fwrite(t.dataPtr<float>(), 4, (size_t)t.numel(), rf);
    fclose(rf);
  }
  if (b200Mode == "ops") {
    b200::adapter::registerOps();           // from main(), after static initialisation
  } else if (b200Mode == "engine") {
    b200_model_desc d{};
    d.hidden = (int32_t)dims[0];
    d.layers = (int32_t)dims[1];
    d.q_heads = (int32_t)dims[2];
    d.kv_heads = (int32_t)dims[3];
    d.head_dim = (int32_t)dims[4];
    d.intermediate = (int32_t)dims[5];
    d.vocab = (int32_t)dims[6];
    d.max_ctx = (int32_t)dims[7];
    d.rms_eps = eps;
    d.qkv_bias = qkvBias >= 0 ? qkvBias : (family == "qwen2");
    d.qk_norm = qkNorm >= 0 ? qkNorm : (family == "qwen3");
    d.tp_rank = 0;
    d.tp_world = 1;
    d.tp_shard_attn = 1;
    tt::nn::RoPE rope = makeRope();
    model = std::make_unique<b200::adapter::ModelB200>(std::move(model), d, std::move(rope));
  } else if (b200Mode != "off") {
    fprintf(stderr, "--b200 must be off, ops or engine\n");
    return 2;
  }

  const std::vector<int64_t> ids = readI64(idsPath);
  const std::vector<int64_t> forced = forcedPath.empty() ? std::vector<int64_t>() : readI64(forcedPath);
  if (ids.empty() || ids.size() % batch != 0 || (!forcedPath.empty() && (int)forced.size() < nNew * batch)) {
    fprintf(stderr, "bad --ids / --forced\n");
    return 2;
  }
  const int64_t V = (int64_t)dims[6];
  const int64_t S = (int64_t)ids.size() / batch;
  if (batch > 1) {
    // batch of equally long prompts: [B, S] ids per step, every sequence's last-position logits and greedy token
    auto idsB = [&](const int64_t* p, int64_t n) {
      tt::Tensor t = tt::Tensor::empty({(int64_t)batch, n}, tt::Options(tt::Device(tt::DeviceType::CPU), tt::DType::Int64));
      std::memcpy(t.dataPtr<int64_t>(), p, sizeof(int64_t) * batch * n);
      return t.to(dev);
    };
    std::vector<int64_t> toksB;
    std::vector<float> logB((size_t)nNew * batch * V);
    std::vector<int64_t> cur(batch);
    auto stepB = [&](const tt::Tensor& in, int64_t n, int k) {
      tt::Tensor logits = model->forward(in);                                         // [B, n, V]
      tt::Tensor last = tt::function::narrow(logits, 1, n - 1, 1).squeeze(1);         // [B, V]
      tt::Tensor next = tt::function::argmax(last, -1, true).to(tt::Device(tt::DeviceType::CPU));
      tt::Tensor f = last.to(tt::DType::Float32).to(tt::Device(tt::DeviceType::CPU));
      std::memcpy(logB.data() + (size_t)k * batch * V, f.dataPtr<float>(), sizeof(float) * batch * V);
      for (int b = 0; b < batch; ++b) {
        cur[b] = next.dataPtr<int64_t>()[b];
        toksB.push_back(cur[b]);
      }
    };
    model->resetCache();
    stepB(idsB(ids.data(), S), S, 0);
    for (int k = 1; k < nNew; ++k) {
      std::vector<int64_t> in(cur);
      if (!forced.empty())
        for (int b = 0; b < batch; ++b) in[b] = forced[(size_t)(k - 1) * batch + b];
      stepB(idsB(in.data(), 1), 1, k);
    }
    FILE* fb = fopen(outPath.c_str(), "wb");
    if (!fb) return 4;
    fwrite(toksB.data(), 8, toksB.size(), fb);
    fwrite(logB.data(), 4, logB.size(), fb);
    fclose(fb);
    if (timeSteps > 0) {
      model->resetCache();
      tt::Tensor curT = tt::function::argmax(
          tt::function::narrow(model->forward(idsB(ids.data(), S)), 1, S - 1, 1).squeeze(1), -1, true);   // [B, 1]
      cudaDeviceSynchronize();
      const auto t0 = std::chrono::steady_clock::now();
      for (int k = 0; k < timeSteps; ++k) {
        tt::Tensor logits = model->forward(curT);
        curT = tt::function::argmax(tt::function::narrow(logits, 1, 0, 1).squeeze(1), -1, true);
      }
      cudaDeviceSynchronize();
      const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
      printf("{\"impl\": \"reference-cuda\", \"b200\": \"%s\", \"batch\": %d, \"tokens_per_s\": %.3f, \"us_per_step\": %.2f, "
             "\"steps\": %d, \"prompt\": %lld}\n",
             b200Mode.c_str(), batch, timeSteps * batch / dt, dt / timeSteps * 1e6, timeSteps, (long long)S);
    }
    fflush(stdout);
    _exit(0);
  }
  auto idsTensor = [&](const int64_t* p, int64_t n) {
    tt::Tensor t = tt::Tensor::empty({1, n}, tt::Options(tt::Device(tt::DeviceType::CPU), tt::DType::Int64));
    std::memcpy(t.dataPtr<int64_t>(), p, sizeof(int64_t) * n);
    return t.to(dev);
  };
  std::vector<int64_t> tokens;
  std::vector<float> logitsAll((size_t)nNew * V);
  auto step = [&](const tt::Tensor& in, int64_t n, int k) {
    tt::Tensor logits = model->forward(in);                                        // [1, n, V] bf16
    tt::Tensor last = tt::function::narrow(logits, 1, n - 1, 1).squeeze(1);        // [1, V]
    tt::Tensor next = tt::function::argmax(last, -1, true);                        // greedy sampler
    tt::Tensor f = last.to(tt::DType::Float32).to(tt::Device(tt::DeviceType::CPU));
    std::memcpy(logitsAll.data() + (size_t)k * V, f.dataPtr<float>(), sizeof(float) * V);
    const int64_t tok = next.to(tt::Device(tt::DeviceType::CPU)).dataPtr<int64_t>()[0];
    tokens.push_back(tok);
    return tok;
  };
  model->resetCache();
  if (batched && !forced.empty()) {
    std::vector<int64_t> all(ids);
    all.insert(all.end(), forced.begin(), forced.begin() + (nNew - 1));
    const int64_t n = (int64_t)all.size();
    tt::Tensor logits = model->forward(idsTensor(all.data(), n));                  // [1, n, V] bf16
    tt::Tensor tail = tt::function::narrow(logits, 1, S - 1, nNew).squeeze(0);     // [nNew, V]
    tt::Tensor next = tt::function::argmax(tail, -1, true).to(tt::Device(tt::DeviceType::CPU));
    tt::Tensor f = tail.to(tt::DType::Float32).to(tt::Device(tt::DeviceType::CPU));
    std::memcpy(logitsAll.data(), f.dataPtr<float>(), sizeof(float) * (size_t)nNew * V);
    for (int k = 0; k < nNew; ++k) tokens.push_back(next.dataPtr<int64_t>()[k]);
  } else {
    int64_t tok = step(idsTensor(ids.data(), S), S, 0);
    for (int k = 1; k < nNew; ++k) {
      const int64_t in = forced.empty() ? tok : forced[k - 1];
      tok = step(idsTensor(&in, 1), 1, k);
    }
  }
  FILE* f = fopen(outPath.c_str(), "wb");
  if (!f) return 4;
  fwrite(tokens.data(), 8, tokens.size(), f);
  fwrite(logitsAll.data(), 4, logitsAll.size(), f);
  fclose(f);

  double prefillMs = -1.0;
  if (timeSteps > 0) {
    // prefill of the whole prompt as genNextToken does it (forward over S tokens, lm_head over all S, narrow, argmax)
    model->resetCache();
    cudaDeviceSynchronize();
    const auto p0 = std::chrono::steady_clock::now();
    tt::Tensor first = tt::function::argmax(
        tt::function::narrow(model->forward(idsTensor(ids.data(), S)), 1, S - 1, 1).squeeze(1), -1, true);
    cudaDeviceSynchronize();
    prefillMs = std::chrono::duration<double>(std::chrono::steady_clock::now() - p0).count() * 1e3;
  }
  if (timeSteps > 0) {
    // the reference's generateSync decode loop, timed like bench.py times ours: device-resident token feeds the next
    // step, no logits copy; one synchronisation at the end
    model->resetCache();
    tt::Tensor cur = tt::function::argmax(
        tt::function::narrow(model->forward(idsTensor(ids.data(), S)), 1, S - 1, 1).squeeze(1), -1, true);
    cudaDeviceSynchronize();
    const auto t0 = std::chrono::steady_clock::now();
    for (int k = 0; k < timeSteps; ++k) {
      tt::Tensor logits = model->forward(cur);
      cur = tt::function::argmax(tt::function::narrow(logits, 1, 0, 1).squeeze(1), -1, true);
    }
    cudaDeviceSynchronize();
    const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    printf("{\"impl\": \"reference-cuda\", \"b200\": \"%s\", \"tokens_per_s\": %.3f, \"us_per_token\": %.2f, \"steps\": %d, "
           "\"prompt\": %lld, \"prefill_ms\": %.3f}\n",
           b200Mode.c_str(), timeSteps / dt, dt / timeSteps * 1e6, timeSteps, (long long)S, prefillMs);
  }
  fflush(stdout);
  _exit(0);  // skip static destructors (the reference's allocator asserts on teardown order)
}
