// TEST INFRASTRUCTURE.  C-ABI harness around the reference's own C++ restatement of fairseq2
// (/root/reference/ggml/examples/unity/fairseq2.cpp, compiled in place by oracle/Makefile together with ggml; nothing
// is copied).  It builds a text decoder in memory from named fp32 tensors and runs the reference's
// `generate_sequence` (fairseq2.cpp:1371-1608: embedding frontend, StandardTransformerDecoder with KV cache,
// final projection, log-softmax, beam search).  tests/golden/make_golden_beam.py uses it to produce the fixtures that
// pin oracle/unity_oracle.py::UnityOracle.beam_search; nothing under seamless_communication_b200/ may link it.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "fairseq2.h"

// defined in fairseq2.cpp, not declared in its header
extern "C" ggml_tensor* StandardTransformerEncoder_forward(fairseq2_model& model, const std::string& prefix, ggml_tensor* seqs,
                                                           ggml_tensor* padding_mask);
extern "C" ggml_tensor* StandardTransformerDecoder_forward(fairseq2_model& model, const std::string& prefix, ggml_tensor* seqs,
                                                           ggml_tensor* padding_mask, ggml_tensor* encoder_output,
                                                           ggml_tensor* encoder_padding_mask);

extern "C" ggml_tensor* ConvModule_forward(fairseq2_model& model, const std::string& prefix, ggml_tensor* seqs);
extern "C" ggml_tensor* StandardConformerEncoderLayer_forward(fairseq2_model& model, const std::string& prefix, ggml_tensor* seqs,
                                                              ggml_tensor* padding_mask);

extern "C" ggml_tensor* WaveformToFbank_forward(fairseq2_model& model, const std::string& prefix, ggml_tensor* waveform);

namespace {

std::int64_t double_bits(double v) {
  std::int64_t b;
  std::memcpy(&b, &v, sizeof(b));
  return b;
}

ggml_context* make_ctx(std::size_t bytes) {
  ggml_init_params p;
  p.mem_size = bytes;
  p.mem_buffer = nullptr;
  p.no_alloc = false;
  return ggml_init(p);
}


ggml_context* build_model(fairseq2_model& model, int n_tensors, const char** names, const float** data, const std::int64_t* d0,
                          const std::int64_t* d1, int n_modules, const char** modules, int n_layernorm,
                          const char** layernorm_names, double ln_eps, int n_attn, const char** attn_names, int num_heads,
                          int n_layers_norm_order, const char** layer_names, int unk_idx) {
  std::size_t total = 0;
  for (int i = 0; i < n_tensors; ++i) total += (std::size_t)d0[i] * (d1[i] > 0 ? d1[i] : 1) * sizeof(float);
  ggml_context* wctx = make_ctx(total + (std::size_t)n_tensors * 1024 + (1u << 20));
  model.tensors_ctx = wctx;
  for (int i = 0; i < n_tensors; ++i) {
    ggml_tensor* t = d1[i] > 0 ? ggml_new_tensor_2d(wctx, GGML_TYPE_F32, d1[i], d0[i]) : ggml_new_tensor_1d(wctx, GGML_TYPE_F32, d0[i]);
    std::memcpy(t->data, data[i], ggml_nbytes(t));
    ggml_set_name(t, names[i]);
    model.tensors[names[i]] = t;
  }
  for (int i = 0; i < n_modules; ++i)
    if (model.tensors.find(modules[i]) == model.tensors.end()) model.tensors[modules[i]] = nullptr;
  for (int i = 0; i < n_layernorm; ++i) model.layer_config[std::string(layernorm_names[i]) + ".eps"] = double_bits(ln_eps);
  for (int i = 0; i < n_attn; ++i) model.layer_config[std::string(attn_names[i]) + ".num_heads"] = num_heads;
  for (int i = 0; i < n_layers_norm_order; ++i)
    model.layer_config[std::string(layer_names[i]) + ".norm_order"] = TRANSFORMER_NORM_ORDER_PRE;
  model.vocab.token_to_id["<unk>"] = unk_idx;

  return wctx;
}

}  // namespace

extern "C" {

// Tensors arrive as a flat list: names[i] (NUL-terminated), 1-D or 2-D fp32 data in torch row-major order with
// shape (d0[i], d1[i]) (d1 = 0 for vectors).  A torch (out, in) matrix becomes a ggml tensor with ne = {in, out}.
// `modules` lists the module names that must answer has_layer() (layers, layer norms, attention blocks).
// Returns the number of finished hypotheses written (<= beam), or -1.
int fs2ref_generate(int n_tensors, const char** names, const float** data, const std::int64_t* d0, const std::int64_t* d1,
                    int n_modules, const char** modules, int n_layernorm, const char** layernorm_names, double ln_eps,
                    int n_attn, const char** attn_names, int num_heads, int n_layers_norm_order, const char** layer_names,
                    const float* encoder_output, int s_enc, int model_dim, const std::int32_t* prefix, int prefix_len,
                    int beam, int min_seq_len, float soft_a, int soft_b, int hard_max, float len_penalty, float unk_penalty,
                    int pad_idx, int unk_idx, int bos_idx, int eos_idx, int max_out_len, std::int32_t* out_tokens,
                    std::int32_t* out_lens, float* out_scores, float* out_step_scores) {
  fairseq2_model model;
  ggml_context* wctx = build_model(model, n_tensors, names, data, d0, d1, n_modules, modules, n_layernorm, layernorm_names, ln_eps,
                                  n_attn, attn_names, num_heads, n_layers_norm_order, layer_names, unk_idx);
  ggml_context* io_ctx = make_ctx((std::size_t)s_enc * model_dim * sizeof(float) + (64u << 20));
  model.ctx = io_ctx;
  ggml_tensor* enc = ggml_new_tensor_2d(io_ctx, GGML_TYPE_F32, model_dim, s_enc);
  std::memcpy(enc->data, encoder_output, ggml_nbytes(enc));
  ggml_tensor* pre = ggml_new_tensor_1d(io_ctx, GGML_TYPE_I32, prefix_len);
  std::memcpy(pre->data, prefix, sizeof(std::int32_t) * prefix_len);

  SequenceGeneratorJob job;
  job.opts.beam_size = beam;
  job.opts.min_seq_len = min_seq_len;
  job.opts.soft_max_seq_len_a = soft_a;
  job.opts.soft_max_seq_len_b = soft_b;
  job.opts.hard_max_seq_len = hard_max;
  job.opts.len_penalty = len_penalty;
  job.opts.unk_penalty = unk_penalty;
  job.opts.normalize_scores = true;
  job.opts.mem_mb = 512;
  job.prefix_seq = pre;
  job.pad_idx = pad_idx;
  job.unk_idx = unk_idx;
  job.bos_idx = bos_idx;
  job.eos_idx = eos_idx;
  job.num_threads = 1;

  ggml_context* result_ctx = make_ctx(16u << 20);
  Hypothesis* hyps = generate_sequence(model, job, enc, nullptr, result_ctx, 1);
  int n = 0;
  if (getenv("FS2REF_DEBUG")) for (int b = 0; b < beam; ++b) fprintf(stderr, "hyp %d seq=%p score=%f len=%d\n", b, (void*)hyps[b].seq, hyps[b].score, hyps[b].seq ? (int)hyps[b].seq->ne[0] : -1);
  for (int b = 0; b < beam; ++b) {
    if (hyps[b].seq == nullptr) continue;
    const int len = (int)hyps[b].seq->ne[0];
    if (len > max_out_len) return -1;
    std::memcpy(out_tokens + (std::size_t)n * max_out_len, hyps[b].seq->data, sizeof(std::int32_t) * len);
    std::memcpy(out_step_scores + (std::size_t)n * max_out_len, hyps[b].step_scores->data, sizeof(float) * len);
    out_lens[n] = len;
    out_scores[n] = hyps[b].score;
    ++n;
  }
  ggml_free(result_ctx);
  ggml_free(io_ctx);
  ggml_free(wctx);
  return n;
}

// Teacher-forced decoder pass without KV cache (StandardTransformerDecoder_forward over the whole prefix with the causal
// mask, fairseq2.cpp:1062-1094) + final projection: logits [n_tokens][vocab].  Used to check the model wiring.
int fs2ref_decoder_logits(int n_tensors, const char** names, const float** data, const std::int64_t* d0, const std::int64_t* d1,
                          int n_modules, const char** modules, int n_layernorm, const char** layernorm_names, double ln_eps,
                          int n_attn, const char** attn_names, int num_heads, int n_layers_norm_order, const char** layer_names,
                          const float* encoder_output, int s_enc, int model_dim, const std::int32_t* tokens, int n_tokens,
                          float* out_logits) {
  fairseq2_model model;
  ggml_context* wctx = build_model(model, n_tensors, names, data, d0, d1, n_modules, modules, n_layernorm, layernorm_names, ln_eps,
                                  n_attn, attn_names, num_heads, n_layers_norm_order, layer_names, 1);
  ggml_context* ctx = make_ctx(256u << 20);
  model.ctx = ctx;
  ggml_tensor* enc = ggml_new_tensor_3d(ctx, GGML_TYPE_F32, model_dim, s_enc, 1);
  std::memcpy(enc->data, encoder_output, ggml_nbytes(enc));
  ggml_tensor* seqs = ggml_new_tensor_2d(ctx, GGML_TYPE_I32, n_tokens, 1);
  std::memcpy(seqs->data, tokens, sizeof(std::int32_t) * n_tokens);
  ggml_tensor* x = TransformerEmbeddingFrontend_forward(model, "text_decoder_frontend", seqs);
  ggml_tensor* y = StandardTransformerDecoder_forward(model, "text_decoder", x, nullptr, enc, nullptr);
  ggml_tensor* logits = Linear_forward(model, "final_proj", y);
  ggml_cgraph* gf = ggml_new_graph(ctx);
  ggml_build_forward_expand(gf, logits);
  ggml_graph_compute_with_ctx(ctx, gf, 1);
  std::memcpy(out_logits, logits->data, ggml_nbytes(logits));
  const int vocab = (int)logits->ne[0];
  ggml_free(ctx);
  ggml_free(wctx);
  return vocab;
}

// StandardTransformerEncoder_forward (fairseq2.cpp:955-977: pre-LN encoder layers + final LayerNorm) over one unpadded
// sequence x [S][model_dim]; the T2U encoder of the path is this module.
int fs2ref_encoder(int n_tensors, const char** names, const float** data, const std::int64_t* d0, const std::int64_t* d1,
                   int n_modules, const char** modules, int n_layernorm, const char** layernorm_names, double ln_eps, int n_attn,
                   const char** attn_names, int num_heads, int n_layers_norm_order, const char** layer_names, const char* prefix,
                   const float* x, int seq_len, int model_dim, float* out) {
  fairseq2_model model;
  ggml_context* wctx = build_model(model, n_tensors, names, data, d0, d1, n_modules, modules, n_layernorm, layernorm_names, ln_eps,
                                  n_attn, attn_names, num_heads, n_layers_norm_order, layer_names, 1);
  ggml_context* ctx = make_ctx(256u << 20);
  model.ctx = ctx;
  ggml_tensor* seqs = ggml_new_tensor_3d(ctx, GGML_TYPE_F32, model_dim, seq_len, 1);
  std::memcpy(seqs->data, x, ggml_nbytes(seqs));
  ggml_tensor* y = StandardTransformerEncoder_forward(model, prefix, seqs, nullptr);
  ggml_cgraph* gf = ggml_new_graph(ctx);
  ggml_build_forward_expand(gf, y);
  ggml_graph_compute_with_ctx(ctx, gf, 1);
  std::memcpy(out, y->data, ggml_nbytes(y));
  ggml_free(ctx);
  ggml_free(wctx);
  return 0;
}

// The Conformer pieces of the mirror (w2v-BERT v1 variant): mode 0 = ConvModule_forward(prefix + ".conv")
// (fairseq2.cpp:698-731: LayerNorm, pointwise conv, GLU, depthwise conv, BatchNorm, SiLU, pointwise conv, + residual),
// mode 1 = StandardConformerEncoderLayer_forward(prefix) (fairseq2.cpp:733-756: the whole block, 16 heads hard-coded,
// relative positions from the `speech_encoder.pos_enc` table).  x / out: [S][model_dim].
int fs2ref_conformer(int n_tensors, const char** names, const float** data, const std::int64_t* d0, const std::int64_t* d1,
                     int n_layernorm, const char** layernorm_names, double ln_eps, const char* prefix, int mode, const float* x,
                     int seq_len, int model_dim, float* out) {
  fairseq2_model model;
  ggml_context* wctx = build_model(model, n_tensors, names, data, d0, d1, 0, nullptr, n_layernorm, layernorm_names, ln_eps, 0, nullptr,
                                  16, 0, nullptr, 1);
  ggml_context* ctx = make_ctx(512u << 20);
  model.ctx = ctx;
  ggml_tensor* seqs = ggml_new_tensor_2d(ctx, GGML_TYPE_F32, model_dim, seq_len);
  std::memcpy(seqs->data, x, ggml_nbytes(seqs));
  ggml_tensor* y = mode == 0 ? ConvModule_forward(model, std::string(prefix) + ".conv", seqs)
                             : StandardConformerEncoderLayer_forward(model, prefix, seqs, nullptr);
  ggml_cgraph* gf = ggml_new_graph(ctx);
  ggml_build_forward_expand(gf, y);
  ggml_graph_compute_with_ctx(ctx, gf, 1);
  std::memcpy(out, y->data, ggml_nbytes(y));
  ggml_free(ctx);
  ggml_free(wctx);
  return 0;
}

// WaveformToFbank_forward (fairseq2.cpp:553-602): knf log-mel frames of an already scaled (x 2**15) waveform, per-bin
// standardisation over time (ggml_norm: biased variance, eps 1e-5), an odd last frame dropped, two frames stacked per
// row (the Wav2Vec2FbankFeatureExtractor stride).  out [rows][cols] with cols = 160; returns 0, or -1 if `capacity`
// floats do not hold the result.
int fs2ref_waveform_to_fbank(const float* wave_scaled, int n_samples, float* out, int capacity, int* rows, int* cols) {
  fairseq2_model model;
  ggml_context* ctx = make_ctx(256u << 20);
  model.ctx = ctx;
  ggml_tensor* w = ggml_new_tensor_2d(ctx, GGML_TYPE_F32, n_samples, 1);
  std::memcpy(w->data, wave_scaled, ggml_nbytes(w));
  ggml_tensor* y = WaveformToFbank_forward(model, "", w);
  ggml_cgraph* gf = ggml_new_graph(ctx);
  ggml_build_forward_expand(gf, y);
  ggml_graph_compute_with_ctx(ctx, gf, 1);
  *cols = (int)y->ne[0];
  *rows = (int)y->ne[1];
  int rc = -1;
  if ((long long)(*rows) * (*cols) <= capacity) {
    std::memcpy(out, y->data, ggml_nbytes(y));
    rc = 0;
  }
  ggml_free(ctx);
  return rc;
}

}  // extern "C"
