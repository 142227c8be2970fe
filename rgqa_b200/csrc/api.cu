// rgqa_b200 — C ABI (include/rgqa_b200.h): handle, weight ingestion by reference state-dict key,
// the LXMERT / UNITER forward wiring over the kernels, and the scorers.
//
// Data layout in HBM (per handle, grown on demand for the largest (B, L) seen):
//   hidA/hidB  [Mt, 768]   activation ping-pong; LXMERT keeps the language rows [0, B*L) and the
//                          visual rows [B*L, B*L + B*36) in ONE buffer so the cross-modality layer's
//                          shared Q|K|V and output projections (one weight set for both directions,
//                          src/lxrt/modeling.py:455-459) run as a single GEMM over all Mt rows
//   qkv        [Mt, 2304]  fused Q|K|V projection, read in place (strided) by the attention kernel
//   ctx        [Mt, 768]   attention output
//   inter      [Mt, 3072]  FFN intermediate (after erf-GELU)
//   resA/resB  [Mt, 768]   fp32 residual stream (PREC_BF16 only): every LayerNorm writes its row twice,
//                          bf16 into hidX (next GEMM operand) and fp32 here (next residual add)
//   pre        [Mt, 768]   fp32 pre-LayerNorm sums (GEMM + bias + residual)
//   featb      [Mv, 2048]  bf16 copy of the RoI features when the caller hands fp32 ones
// Weights: nn.Linear weights stay [out, in] row-major (= N x K, K contiguous: the K-major B operand
// of the tcgen05 GEMM), cast to bf16 in PREC_BF16; Q, K, V are concatenated along N.
#include <cstdarg>
#include <cstdio>
#include <cmath>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "../../include/rgqa_b200.h"
#include "common.cuh"
#include "kernels.h"

namespace rgqa {

static thread_local char g_err[1024] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

namespace {

struct DevBuf {
  void* p = nullptr;
  size_t bytes = 0;
  int ensure(size_t need) {
    if (need <= bytes) return 0;
    if (p != nullptr) {
      RGQA_CUDA_OK(cudaFree(p));
      p = nullptr;
      bytes = 0;
    }
    RGQA_CUDA_OK(cudaMalloc(&p, need));
    bytes = need;
    return 0;
  }
  void release() {
    if (p != nullptr) cudaFree(p);
    p = nullptr;
    bytes = 0;
  }
};

struct RawTensor {
  float* p = nullptr;  // device fp32
  std::vector<int64_t> shape;
  int64_t numel = 0;
};

struct Linear {
  void* W = nullptr;      // [N, K] bf16 or f32 (precision of the handle)
  float* b = nullptr;     // [N] f32
  int N = 0, K = 0;
};
struct LNp {
  float* g = nullptr;
  float* b = nullptr;
};
// A consumer weight with the LayerNorm of its input folded in (deferred LayerNorm): W' = bf16(W∘gamma -
// rowmean), b' = b + W·beta.  W == nullptr when the input is not behind a LayerNorm (first layer) or in fp32 mode.
struct Folded {  // state_f16 handles: W is fp16, and a consumer WITHOUT a predecessor LayerNorm gets a plain fp16 copy here too
  void* W = nullptr;
  float* b = nullptr;
  float* c = nullptr;  // row sums of W' (the centring's rounding residue); non-null only when too large to drop (fold_stack)
};
struct AttW {
  Linear qkv;  // [3H, H]
  Linear out;  // [H, H]
  LNp ln;
  Folded qkv_f;   // folded with the LayerNorm that produced this block's input (language rows for the cross block)
  Folded qkv_f2;  // cross block only: folded with the visual branch's LayerNorm
  void* qkv_hm = nullptr;    // self-attention blocks: the Q|K|V weight this block's forward uses (folded when it has a
  float* qkv_hm_b = nullptr;  // predecessor LayerNorm) re-packed head-major [head][q|k|v][64] for qkv_attention_fused
  float* out_bb = nullptr;   // out.b + beta of the LayerNorm owed to the block's input (the producer epilogue's bias)
  float* out_bb2 = nullptr;  // cross block only: with the visual branch's beta
};
struct FfnW {
  Linear inter;  // [I, H]
  Linear out;    // [H, I]
  LNp ln;
  Folded inter_f;  // folded with the attention block's output LayerNorm
  float* out_bb = nullptr;  // out.b + beta of the attention block's output LayerNorm
};
struct LayerW {
  AttW att;
  FfnW ffn;
};
struct XLayerW {
  AttW cross, lang_self, visn_self;
  FfnW lang_ffn, visn_ffn;
};

}  // namespace
}  // namespace rgqa

using namespace rgqa;

struct rgqa_handle;

struct rgqa_handle {
  rgqa_config cfg;
  std::vector<std::string> names;           // expected reference keys
  std::map<std::string, RawTensor> raw;     // uploaded fp32 tensors
  std::vector<void*> owned;                 // packed allocations
  bool finalized = false;
  bool state_f16 = false;  // set by rgqa_finalize_weights (option "state_f16"): the deferred forward keeps its state as one fp16 word per
                           // element and every consumer weight (folded, or plain in front of the first layer) is an fp16 copy

  // embeddings (fp32)
  float *word = nullptr, *posemb = nullptr, *typeemb = nullptr;
  LNp emb_ln;
  // visual embedding
  Linear visn_fc;                 // [H, 2048]
  float *box_W = nullptr, *box_b = nullptr;  // [H, pos_dim], [H] fp32
  float *visn_W32 = nullptr;                 // [H, 2048] fp32 copy of visn_fc (tf32 operand of the one-kernel visual encoder)
  float *box_Wp = nullptr, *box_stat = nullptr;  // [H][8] zero-padded box_fc weight; closed-form LayerNorm statistics (make_box_stats)
  LNp visn_ln, box_ln, img_ln;    // img_ln: UNITER's third LN
  std::vector<LayerW> l_layers, r_layers;
  std::vector<XLayerW> x_layers;
  Linear pooler, fc0, fc3;
  Folded pooler_f;
  float *ones = nullptr, *zeros = nullptr;  // [H] identity LayerNorm
  LNp fc_ln;
  // Mahalanobis
  float *maha_mu = nullptr, *maha_P = nullptr, *maha_m = nullptr;
  void* maha_Ws = nullptr;  // bf16 [H + A, 3H]: split-bf16 rows of the stack [P; mu·P] (tensor-core Mahalanobis, bf16 handles)
  bool has_maha = false;
  float max_fold_residue = 0.f;  // largest |c[n]| over every folded consumer weight (fold_stack)
  int exact_folds = 0;           // consumers whose residue was too large to drop: they run the exact epilogue (ln_c)
  long long graph_captures = 0;  // CUDA graphs instantiated so far (a caller that never repeats its buffers keeps capturing)

  // second stream for the visual branch (language / visual layers are independent between joins)
  cudaStream_t side = nullptr;
  cudaStream_t cap = nullptr;  // capture stream for the small-batch CUDA graphs
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;

  // CUDA graphs of the whole forward, keyed by shape (small batches: inputs / outputs staged in handle-owned buffers)
  // or by shape + the caller's feature / box pointers (large batches: the 300 MB inputs are used in place)
  struct GraphEntryT {
    int seen = 0;
    int kernels = 0;
    cudaGraphExec_t exec = nullptr;
    DevBuf feat, pos, ids, seg, mask, pooled, logits, scores, labels;
  };
  struct GraphKey {
    long long shape;
    const void* feat;
    const void* pos;
    bool operator<(const GraphKey& o) const {
      if (shape != o.shape) return shape < o.shape;
      if (feat != o.feat) return feat < o.feat;
      return pos < o.pos;
    }
  };
  std::map<GraphKey, GraphEntryT> graphs;

  // workspace
  DevBuf hidA, hidB, resA, resB, qkv, ctx, inter, pre, featb, headbuf, mahabuf;
  DevBuf maha_xs, red_fc, red_maha, red_cnt;  // fused row reductions: split pooled rows, partial records, arrival counters
};

// What one forward call produces.  score_mask: bit (1 << RGQA_SCORE_*) per scorer wanted; scores [5][B] (row = scorer code),
// labels [2][B] (row 0: arg-max of sigmoid(logit), the MSP / energy label; row 1: Mahalanobis arg-max class).
struct FwdOut {
  float* pooled = nullptr;
  float* logits = nullptr;
  unsigned score_mask = 0;
  float* scores = nullptr;
  int32_t* labels = nullptr;
  bool any() const { return pooled != nullptr || logits != nullptr || score_mask != 0; }
};

namespace {

int act_dtype(const rgqa_handle* h) { return h->cfg.precision == RGQA_PREC_BF16 ? DT_BF16 : DT_F32; }
size_t act_size(const rgqa_handle* h) { return h->cfg.precision == RGQA_PREC_BF16 ? 2 : 4; }

void add_linear_names(std::vector<std::string>& v, const std::string& p) {
  v.push_back(p + ".weight");
  v.push_back(p + ".bias");
}
void add_att_names(std::vector<std::string>& v, const std::string& p, const char* self_name) {
  add_linear_names(v, p + "." + self_name + ".query");
  add_linear_names(v, p + "." + self_name + ".key");
  add_linear_names(v, p + "." + self_name + ".value");
  add_linear_names(v, p + ".output.dense");
  add_linear_names(v, p + ".output.LayerNorm");
}
void add_bert_layer_names(std::vector<std::string>& v, const std::string& p) {
  add_att_names(v, p + ".attention", "self");
  add_linear_names(v, p + ".intermediate.dense");
  add_linear_names(v, p + ".output.dense");
  add_linear_names(v, p + ".output.LayerNorm");
}

void build_names(rgqa_handle* h) {
  auto& v = h->names;
  const rgqa_config& c = h->cfg;
  if (c.arch == RGQA_ARCH_LXMERT) {
    const std::string b = "bert.";
    v.push_back(b + "embeddings.word_embeddings.weight");
    v.push_back(b + "embeddings.position_embeddings.weight");
    v.push_back(b + "embeddings.token_type_embeddings.weight");
    add_linear_names(v, b + "embeddings.LayerNorm");
    add_linear_names(v, b + "encoder.visn_fc.visn_fc");
    add_linear_names(v, b + "encoder.visn_fc.visn_layer_norm");
    add_linear_names(v, b + "encoder.visn_fc.box_fc");
    add_linear_names(v, b + "encoder.visn_fc.box_layer_norm");
    for (int i = 0; i < c.l_layers; ++i) add_bert_layer_names(v, b + "encoder.layer." + std::to_string(i));
    for (int i = 0; i < c.r_layers; ++i) add_bert_layer_names(v, b + "encoder.r_layers." + std::to_string(i));
    for (int i = 0; i < c.x_layers; ++i) {
      const std::string x = b + "encoder.x_layers." + std::to_string(i);
      add_att_names(v, x + ".visual_attention", "att");
      add_att_names(v, x + ".lang_self_att", "self");
      add_att_names(v, x + ".visn_self_att", "self");
      add_linear_names(v, x + ".lang_inter.dense");
      add_linear_names(v, x + ".lang_output.dense");
      add_linear_names(v, x + ".lang_output.LayerNorm");
      add_linear_names(v, x + ".visn_inter.dense");
      add_linear_names(v, x + ".visn_output.dense");
      add_linear_names(v, x + ".visn_output.LayerNorm");
    }
    add_linear_names(v, b + "pooler.dense");
  } else {
    const std::string u = "uniter.";
    v.push_back(u + "embeddings.word_embeddings.weight");
    v.push_back(u + "embeddings.position_embeddings.weight");
    v.push_back(u + "embeddings.token_type_embeddings.weight");
    add_linear_names(v, u + "embeddings.LayerNorm");
    add_linear_names(v, u + "img_embeddings.img_linear");
    add_linear_names(v, u + "img_embeddings.img_layer_norm");
    add_linear_names(v, u + "img_embeddings.pos_linear");
    add_linear_names(v, u + "img_embeddings.pos_layer_norm");
    add_linear_names(v, u + "img_embeddings.LayerNorm");
    for (int i = 0; i < c.l_layers; ++i) add_bert_layer_names(v, u + "encoder.layer." + std::to_string(i));
    add_linear_names(v, u + "pooler.dense");
  }
  if (c.num_answers > 0) {
    add_linear_names(v, "logit_fc.0");
    add_linear_names(v, "logit_fc.2");
    add_linear_names(v, "logit_fc.3");
  }
}

std::string strip_prefix(const char* key) {
  std::string k(key);
  const char* prefixes[] = {"module.", "lxrt_encoder.model.", "encoder.model.", "module."};
  for (const char* p : prefixes) {
    const size_t n = strlen(p);
    if (k.compare(0, n, p) == 0) k = k.substr(n);
  }
  return k;
}

int get_raw(rgqa_handle* h, const std::string& name, int64_t d0, int64_t d1, float** out) {
  auto it = h->raw.find(name);
  RGQA_CHECK(it != h->raw.end(), "weight '%s' was never loaded", name.c_str());
  const RawTensor& t = it->second;
  const int64_t want = d1 > 0 ? d0 * d1 : d0;
  RGQA_CHECK(t.numel == want, "weight '%s': expected %lld elements, got %lld", name.c_str(),
             static_cast<long long>(want), static_cast<long long>(t.numel));
  *out = t.p;
  return 0;
}

// Pack rows of up to three fp32 [n_i, K] matrices into one [sum n_i, K] matrix of the handle's precision.
int pack_matrix(rgqa_handle* h, const float* const* srcs, const int* rows, int nsrc, int K, void** out) {
  int N = 0;
  for (int i = 0; i < nsrc; ++i) N += rows[i];
  const size_t esz = act_size(h);
  void* dst = nullptr;
  RGQA_CUDA_OK(cudaMalloc(&dst, static_cast<size_t>(N) * K * esz));
  h->owned.push_back(dst);
  size_t off = 0;
  for (int i = 0; i < nsrc; ++i) {
    const long long n = static_cast<long long>(rows[i]) * K;
    if (h->cfg.precision == RGQA_PREC_BF16) {
      if (cast_f32_to_bf16(nullptr, srcs[i], static_cast<char*>(dst) + off * 2, n)) return 1;
    } else {
      RGQA_CUDA_OK(cudaMemcpy(static_cast<char*>(dst) + off * 4, srcs[i], n * 4, cudaMemcpyDeviceToDevice));
    }
    off += n;
  }
  *out = dst;
  return 0;
}

int pack_vec(rgqa_handle* h, const float* const* srcs, const int* lens, int nsrc, float** out) {
  int N = 0;
  for (int i = 0; i < nsrc; ++i) N += lens[i];
  float* dst = nullptr;
  RGQA_CUDA_OK(cudaMalloc(&dst, static_cast<size_t>(N) * 4));
  h->owned.push_back(dst);
  size_t off = 0;
  for (int i = 0; i < nsrc; ++i) {
    RGQA_CUDA_OK(cudaMemcpy(dst + off, srcs[i], static_cast<size_t>(lens[i]) * 4, cudaMemcpyDeviceToDevice));
    off += lens[i];
  }
  *out = dst;
  return 0;
}

int make_linear(rgqa_handle* h, const std::string& p, int N, int K, Linear* out) {
  float *w, *b;
  if (get_raw(h, p + ".weight", N, K, &w)) return 1;
  if (get_raw(h, p + ".bias", N, 0, &b)) return 1;
  const float* ws[1] = {w};
  const int rows[1] = {N};
  if (pack_matrix(h, ws, rows, 1, K, &out->W)) return 1;
  const float* bs[1] = {b};
  if (pack_vec(h, bs, rows, 1, &out->b)) return 1;
  out->N = N;
  out->K = K;
  return 0;
}
int make_ln(rgqa_handle* h, const std::string& p, int H, LNp* out) {
  float *g, *b;
  if (get_raw(h, p + ".weight", H, 0, &g)) return 1;
  if (get_raw(h, p + ".bias", H, 0, &b)) return 1;
  const float* gs[1] = {g};
  const float* bs[1] = {b};
  const int n[1] = {H};
  if (pack_vec(h, gs, n, 1, &out->g)) return 1;
  if (pack_vec(h, bs, n, 1, &out->b)) return 1;
  return 0;
}
int make_att(rgqa_handle* h, const std::string& p, const char* self_name, AttW* out) {
  const int H = h->cfg.hidden;
  float *wq, *wk, *wv, *bq, *bk, *bv;
  const std::string s = p + "." + self_name;
  if (get_raw(h, s + ".query.weight", H, H, &wq) || get_raw(h, s + ".key.weight", H, H, &wk) ||
      get_raw(h, s + ".value.weight", H, H, &wv) || get_raw(h, s + ".query.bias", H, 0, &bq) ||
      get_raw(h, s + ".key.bias", H, 0, &bk) || get_raw(h, s + ".value.bias", H, 0, &bv))
    return 1;
  const float* ws[3] = {wq, wk, wv};
  const float* bs[3] = {bq, bk, bv};
  const int rows[3] = {H, H, H};
  if (pack_matrix(h, ws, rows, 3, H, &out->qkv.W)) return 1;
  if (pack_vec(h, bs, rows, 3, &out->qkv.b)) return 1;
  out->qkv.N = 3 * H;
  out->qkv.K = H;
  if (make_linear(h, p + ".output.dense", H, H, &out->out)) return 1;
  return make_ln(h, p + ".output.LayerNorm", H, &out->ln);
}
int make_ffn(rgqa_handle* h, const std::string& inter, const std::string& outp, FfnW* out) {
  const int H = h->cfg.hidden, I = h->cfg.intermediate;
  if (make_linear(h, inter + ".dense", I, H, &out->inter)) return 1;
  if (make_linear(h, outp + ".dense", H, I, &out->out)) return 1;
  return make_ln(h, outp + ".LayerNorm", H, &out->ln);
}
int make_bert_layer(rgqa_handle* h, const std::string& p, LayerW* out) {
  if (make_att(h, p + ".attention", "self", &out->att)) return 1;
  return make_ffn(h, p + ".intermediate", p + ".output", &out->ffn);
}
int copy_raw(rgqa_handle* h, const std::string& name, int64_t d0, int64_t d1, float** out) {
  float* src;
  if (get_raw(h, name, d0, d1, &src)) return 1;
  const float* s[1] = {src};
  const int n[1] = {static_cast<int>(d1 > 0 ? d0 * d1 : d0)};
  return pack_vec(h, s, n, 1, out);
}

// box_fc has K <= 8: its LayerNorm statistics over the 768 outputs are closed forms of the box vector.  With v = (box, 1):
// mean_n y_n = v · (colmean(W2), mean(b2)) and var_n y_n = vᵀ C v, C = second moments of the CENTRED rows (fp64 here).
// Layout of bstat: [0..7] colmean, [8] mean(b2), [16..96] C (9 x 9, index 8 = the constant).  W2p = W2 padded to [H][8].
int make_box_stats(const float* W2_dev, const float* b2_dev, int H, int K, std::vector<float>* W2p, std::vector<float>* bstat) {
  std::vector<float> w(static_cast<size_t>(H) * K), b(H);
  RGQA_CUDA_OK(cudaMemcpy(w.data(), W2_dev, w.size() * 4, cudaMemcpyDeviceToHost));
  RGQA_CUDA_OK(cudaMemcpy(b.data(), b2_dev, b.size() * 4, cudaMemcpyDeviceToHost));
  W2p->assign(static_cast<size_t>(H) * 8, 0.f);
  bstat->assign(16 + 81, 0.f);
  double cm[9] = {0};
  for (int n = 0; n < H; ++n) {
    for (int j = 0; j < K; ++j) {
      (*W2p)[static_cast<size_t>(n) * 8 + j] = w[static_cast<size_t>(n) * K + j];
      cm[j] += w[static_cast<size_t>(n) * K + j];
    }
    cm[8] += b[n];
  }
  for (int j = 0; j < 9; ++j) cm[j] /= H;
  double C[9][9] = {{0}};
  for (int n = 0; n < H; ++n) {
    double r[9] = {0};
    for (int j = 0; j < K; ++j) r[j] = w[static_cast<size_t>(n) * K + j] - cm[j];
    r[8] = b[n] - cm[8];
    for (int a = 0; a < 9; ++a)
      for (int c = 0; c < 9; ++c) C[a][c] += r[a] * r[c];
  }
  for (int j = 0; j < 8; ++j) (*bstat)[j] = static_cast<float>(cm[j]);
  (*bstat)[8] = static_cast<float>(cm[8]);
  for (int a = 0; a < 9; ++a)
    for (int c = 0; c < 9; ++c) (*bstat)[16 + a * 9 + c] = static_cast<float>(C[a][c] / H);
  return 0;
}
int upload_vec(rgqa_handle* h, const std::vector<float>& v, float** out) {
  float* d = nullptr;
  RGQA_CUDA_OK(cudaMalloc(reinterpret_cast<void**>(&d), v.size() * sizeof(float)));
  h->owned.push_back(d);
  RGQA_CUDA_OK(cudaMemcpy(d, v.data(), v.size() * sizeof(float), cudaMemcpyHostToDevice));
  *out = d;
  return 0;
}

// ---- deferred-LayerNorm weight folding (bf16 path) ------------------------------------------------
// names[i] (+ ".weight" / ".bias") are stacked row-wise, each [rows, K]; folded with the LayerNorm `ln` of the input.
int fold_stack(rgqa_handle* h, const std::vector<std::string>& names, int rows, int K, const LNp& ln, Folded* out) {
  const size_t n = names.size();
  void* W = nullptr;
  float *b = nullptr, *c = nullptr;
  RGQA_CUDA_OK(cudaMalloc(&W, n * rows * static_cast<size_t>(K) * 2));
  h->owned.push_back(W);
  RGQA_CUDA_OK(cudaMalloc(reinterpret_cast<void**>(&b), n * rows * sizeof(float)));
  h->owned.push_back(b);
  RGQA_CUDA_OK(cudaMalloc(reinterpret_cast<void**>(&c), n * rows * sizeof(float)));  // rounding residue, unused
  h->owned.push_back(c);
  for (size_t i = 0; i < n; ++i) {
    float *w, *bias;
    if (get_raw(h, names[i] + ".weight", rows, K, &w) || get_raw(h, names[i] + ".bias", rows, 0, &bias)) return 1;
    if (fold_ln_weight(nullptr, w, bias, ln.g, ln.b, rows, K, static_cast<char*>(W) + i * rows * static_cast<size_t>(K) * 2,
                       c + i * rows, b + i * rows, h->state_f16 ? 1 : 0))
      return 1;
  }
  out->W = W;
  out->b = b;
  // Guard for the dropped - rstd * mean * c[n] term of the consumer epilogue: error-feedback rounding leaves |c| <= half an ulp
  // of a weight (~6e-5 at |w| ~ 0.03); with hidden rows at |mean| / std up to kLnMeanBound the dropped term stays below 1e-3
  // while max|c| * kLnMeanBound <= 1e-3.  Otherwise (very large weights) this consumer keeps c and runs the exact epilogue.
  {
    std::vector<float> hc(n * rows);
    RGQA_CUDA_OK(cudaMemcpy(hc.data(), c, hc.size() * sizeof(float), cudaMemcpyDeviceToHost));
    float mx = 0.f;
    for (float v : hc) mx = (std::fabs(v) > mx || !std::isfinite(v)) ? std::fabs(v) : mx;
    RGQA_CHECK(std::isfinite(mx), "deferred LayerNorm: a folded weight of '%s' leaves the fp16 range; load with rgqa_set_option(\"state_f16\", 0)",
               names[0].c_str());
    if (mx > h->max_fold_residue) h->max_fold_residue = mx;
    constexpr float kLnMeanBound = 8.0f;
    out->c = (mx * kLnMeanBound > 1e-3f) ? c : nullptr;
    if (out->c != nullptr) ++h->exact_folds;
  }
  return 0;
}
// bias + beta (both [n] f32 on the device) as one vector: the residual term LN(x) = (x-mu)*rstd*gamma + beta of a producer
// GEMM then needs no separate beta load
int add_vecs(rgqa_handle* h, const float* a, const float* b, int n, float** out) {
  std::vector<float> ha(n), hb(n);
  RGQA_CUDA_OK(cudaMemcpy(ha.data(), a, n * sizeof(float), cudaMemcpyDeviceToHost));
  RGQA_CUDA_OK(cudaMemcpy(hb.data(), b, n * sizeof(float), cudaMemcpyDeviceToHost));
  for (int i = 0; i < n; ++i) ha[i] += hb[i];
  float* d = nullptr;
  RGQA_CUDA_OK(cudaMalloc(reinterpret_cast<void**>(&d), n * sizeof(float)));
  h->owned.push_back(d);
  RGQA_CUDA_OK(cudaMemcpy(d, ha.data(), n * sizeof(float), cudaMemcpyHostToDevice));
  *out = d;
  return 0;
}
// state_f16 handles: a consumer that sits directly behind the embeddings (identity state, no LayerNorm to fold) still needs its
// weight in the state's format — a plain fp16 copy of the stacked rows, bias unchanged
int plain_f16_stack(rgqa_handle* h, const std::vector<std::string>& names, int rows, int K, Folded* out) {
  const size_t n = names.size();
  void* W = nullptr;
  float* b = nullptr;
  RGQA_CUDA_OK(cudaMalloc(&W, n * rows * static_cast<size_t>(K) * 2));
  h->owned.push_back(W);
  RGQA_CUDA_OK(cudaMalloc(reinterpret_cast<void**>(&b), n * rows * sizeof(float)));
  h->owned.push_back(b);
  for (size_t i = 0; i < n; ++i) {
    float *w, *bias;
    if (get_raw(h, names[i] + ".weight", rows, K, &w) || get_raw(h, names[i] + ".bias", rows, 0, &bias)) return 1;
    if (cast_f32_to_f16(nullptr, w, static_cast<char*>(W) + i * rows * static_cast<size_t>(K) * 2, static_cast<long long>(rows) * K)) return 1;
    RGQA_CUDA_OK(cudaMemcpyAsync(b + i * rows, bias, rows * sizeof(float), cudaMemcpyDeviceToDevice, nullptr));
  }
  out->W = W;
  out->b = b;
  out->c = nullptr;
  return 0;
}
int fold_qkv(rgqa_handle* h, const std::string& p, const char* self_name, const LNp* pred, Folded* out) {
  const std::string s = p + "." + self_name;
  if (pred == nullptr) {
    if (!h->state_f16) return 0;
    return plain_f16_stack(h, {s + ".query", s + ".key", s + ".value"}, h->cfg.hidden, h->cfg.hidden, out);
  }
  return fold_stack(h, {s + ".query", s + ".key", s + ".value"}, h->cfg.hidden, h->cfg.hidden, *pred, out);
}
// [q rows | k rows | v rows] -> [head][q | k | v][64 rows]: one 192-row tile of the fused projection + attention kernel is
// one head.  Source = the folded weight when the block sits behind a LayerNorm, else the plain one.
int pack_head_major(rgqa_handle* h, AttW* w) {
  const int H = h->cfg.hidden, heads = h->cfg.heads, hd = H / heads;
  if (w->qkv_f.c != nullptr) return 0;  // this block needs the exact consumer epilogue: it stays on the separate kernels
  const char* W = static_cast<const char*>(w->qkv_f.W != nullptr ? w->qkv_f.W : w->qkv.W);
  const float* b = w->qkv_f.W != nullptr ? w->qkv_f.b : w->qkv.b;
  char* Wo = nullptr;
  float* bo = nullptr;
  RGQA_CUDA_OK(cudaMalloc(reinterpret_cast<void**>(&Wo), static_cast<size_t>(3) * H * H * 2));
  h->owned.push_back(Wo);
  RGQA_CUDA_OK(cudaMalloc(reinterpret_cast<void**>(&bo), static_cast<size_t>(3) * H * sizeof(float)));
  h->owned.push_back(bo);
  const size_t blk = static_cast<size_t>(hd) * H * 2;  // 64 weight rows
  for (int head = 0; head < heads; ++head)
    for (int part = 0; part < 3; ++part) {
      const size_t src_row = static_cast<size_t>(part) * H + static_cast<size_t>(head) * hd;
      const size_t dst_row = (static_cast<size_t>(head) * 3 + part) * hd;
      RGQA_CUDA_OK(cudaMemcpyAsync(Wo + dst_row * H * 2, W + src_row * H * 2, blk, cudaMemcpyDeviceToDevice, nullptr));
      RGQA_CUDA_OK(cudaMemcpyAsync(bo + dst_row, b + src_row, hd * sizeof(float), cudaMemcpyDeviceToDevice, nullptr));
    }
  w->qkv_hm = Wo;
  w->qkv_hm_b = bo;
  return 0;
}
int make_identity_ln(rgqa_handle* h) {  // gamma = 1, beta = 0: the "LayerNorm" owed to an already-normalised state
  const int H = h->cfg.hidden;
  std::vector<float> one(H, 1.0f), zero(H, 0.0f);
  RGQA_CUDA_OK(cudaMalloc(reinterpret_cast<void**>(&h->ones), H * sizeof(float)));
  h->owned.push_back(h->ones);
  RGQA_CUDA_OK(cudaMalloc(reinterpret_cast<void**>(&h->zeros), H * sizeof(float)));
  h->owned.push_back(h->zeros);
  RGQA_CUDA_OK(cudaMemcpy(h->ones, one.data(), H * sizeof(float), cudaMemcpyHostToDevice));
  RGQA_CUDA_OK(cudaMemcpy(h->zeros, zero.data(), H * sizeof(float), cudaMemcpyHostToDevice));
  return 0;
}
// folds of one BertLayer whose input comes out of LayerNorm `pred` (nullptr: not behind a LayerNorm)
int fold_bert_layer(rgqa_handle* h, const std::string& p, LayerW* w, const LNp* pred) {
  if (fold_qkv(h, p + ".attention", "self", pred, &w->att.qkv_f)) return 1;
  if (pred != nullptr && add_vecs(h, w->att.out.b, pred->b, h->cfg.hidden, &w->att.out_bb)) return 1;
  if (add_vecs(h, w->ffn.out.b, w->att.ln.b, h->cfg.hidden, &w->ffn.out_bb)) return 1;
  return fold_stack(h, {p + ".intermediate.dense"}, h->cfg.intermediate, h->cfg.hidden, w->att.ln, &w->ffn.inter_f);
}

// ---- dtype-dispatched primitives over the handle's precision ------------------------------------
int gemm(rgqa_handle* h, cudaStream_t st, const void* A, long long lda, const Linear& lin, const void* resid,
         long long ldr, void* out, long long ldo, int out_f32, int M, int epi) {
  if (h->cfg.precision == RGQA_PREC_BF16)
    return gemm_bf16(st, A, lda, lin.W, lin.K, lin.b, resid, ldr, out, ldo, out_f32, M, lin.N, lin.K, epi);
  RGQA_CHECK(out_f32, "fp32 path writes fp32 only");
  return gemm_f32(st, static_cast<const float*>(A), lda, static_cast<const float*>(lin.W), lin.K, lin.b,
                  static_cast<const float*>(resid), ldr, static_cast<float*>(out), ldo, M, lin.N, lin.K, epi);
}

inline char* rows_at(void* base, long long row, long long ld, size_t esz) {
  return static_cast<char*>(base) + static_cast<size_t>(row) * ld * esz;
}

// One hidden-state buffer: `act` is the GEMM A operand in the path's activation type; `res` is the fp32
// residual stream.  In PREC_FP32 both are the same fp32 buffer; in PREC_BF16 every LayerNorm writes the
// row twice (bf16 operand + fp32 residual) so the residual path never sees a bf16 rounding.
struct Hid {
  void* act;
  float* res;
};
inline const float* resid_rows(const rgqa_handle* h, const Hid& x, long long r0) {
  return reinterpret_cast<const float*>(rows_at(x.res, r0, h->cfg.hidden, 4));
}
inline float* res_copy_rows(const rgqa_handle* h, const Hid& x, long long r0) {
  // extra fp32 copy target for row kernels; none needed when act already IS the fp32 stream
  if (static_cast<void*>(x.res) == x.act) return nullptr;
  return reinterpret_cast<float*>(rows_at(x.res, r0, h->cfg.hidden, 4));
}

// out = LayerNorm(A · Wᵀ + b + residual(in)) on rows [r0, r0 + M)  (BertAttOutput / BertOutput, modeling.py:357-361,
// :411-415).  bf16 path: ONE kernel (full 768-wide rows per CTA in TMEM, gemm_ln_tcgen05.cu); otherwise the GEMM
// writes the fp32 pre-LayerNorm sum and a row kernel normalises it.
int dense_resid_ln(rgqa_handle* h, cudaStream_t st, const void* A, long long lda, const Linear& lin, const LNp& ln,
                   const Hid& in, const Hid& out, long long r0, int M) {
  const int H = h->cfg.hidden;
  const size_t es = act_size(h);
  if (h->cfg.precision == RGQA_PREC_BF16 && option_get(OPT_GEMM_LN) != 0 && M >= 256)
    return gemm_bf16_ln(st, A, lda, lin.W, lin.K, lin.b, resid_rows(h, in, r0), H, ln.g, ln.b, h->cfg.ln_eps,
                        rows_at(out.act, r0, H, es), res_copy_rows(h, out, r0), M, lin.N, lin.K);
  float* pre = reinterpret_cast<float*>(rows_at(h->pre.p, r0, H, 4));
  if (gemm(h, st, A, lda, lin, resid_rows(h, in, r0), H, pre, H, 1, M, EPI_RESID)) return 1;
  return layernorm(st, act_dtype(h), pre, H, ln.g, ln.b, rows_at(out.act, r0, H, es), res_copy_rows(h, out, r0), M, H,
                   h->cfg.ln_eps);
}

// self-attention block + output projection + LN on rows [r0, r0 + B*S) : in -> out
int att_block(rgqa_handle* h, cudaStream_t st, const AttW& w, const Hid& in, const Hid& out, long long r0, int B, int S,
              const int32_t* key_mask) {
  const int H = h->cfg.hidden;
  const size_t es = act_size(h);
  const int dt = act_dtype(h);
  const int M = B * S;
  const int of32 = dt == DT_F32;
  char* x = rows_at(in.act, r0, H, es);
  char* qkv = rows_at(h->qkv.p, r0, 3 * H, es);
  char* ctx = rows_at(h->ctx.p, r0, H, es);
  if (gemm(h, st, x, H, w.qkv, nullptr, 0, qkv, 3 * H, of32, M, EPI_BIAS)) return 1;
  if (attention(st, dt, qkv, 3 * H, qkv + static_cast<size_t>(H) * es, 3 * H, qkv + static_cast<size_t>(2 * H) * es,
                3 * H, key_mask, ctx, H, B, S, S, h->cfg.heads))
    return 1;
  return dense_resid_ln(h, st, ctx, H, w.out, w.ln, in, out, r0, M);
}

// FFN block on rows [r0, r0 + M): in -> out
int ffn_block(rgqa_handle* h, cudaStream_t st, const FfnW& w, const Hid& in, const Hid& out, long long r0, int M) {
  const int H = h->cfg.hidden, I = h->cfg.intermediate;
  const size_t es = act_size(h);
  const int dt = act_dtype(h);
  const int of32 = dt == DT_F32;
  char* x = rows_at(in.act, r0, H, es);
  char* inter = rows_at(h->inter.p, r0, I, es);
  if (gemm(h, st, x, H, w.inter, nullptr, 0, inter, I, of32, M, EPI_GELU)) return 1;
  return dense_resid_ln(h, st, inter, I, w.out, w.ln, in, out, r0, M);
}

void drop_graphs(rgqa_handle* h) {  // captured graphs bake workspace / weight pointers: any reallocation voids them
  for (auto& kv : h->graphs) {
    if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
    kv.second.exec = nullptr;
  }
}

int ensure_workspace(rgqa_handle* h, long long Mt, long long Mv, int B) {
  const int H = h->cfg.hidden, I = h->cfg.intermediate;
  const size_t es = act_size(h);
  const void* before[4] = {h->hidA.p, h->qkv.p, h->inter.p, h->headbuf.p};
  struct Guard {
    rgqa_handle* h;
    const void* const* b;
    ~Guard() {
      if (b[0] != h->hidA.p || b[1] != h->qkv.p || b[2] != h->inter.p || b[3] != h->headbuf.p) drop_graphs(h);
    }
  } guard{h, before};
  if (h->hidA.ensure(Mt * H * es) || h->hidB.ensure(Mt * H * es) || h->qkv.ensure(Mt * 3 * H * es) ||
      h->ctx.ensure(Mt * H * es) || h->inter.ensure(Mt * I * es) || h->pre.ensure(Mt * H * 4))
    return 1;
  if (h->cfg.precision == RGQA_PREC_BF16 && (h->resA.ensure(Mt * H * 4) || h->resB.ensure(Mt * H * 4))) return 1;
  if (h->cfg.precision == RGQA_PREC_BF16 && h->featb.ensure(Mv * h->cfg.visual_feat_dim * 2)) return 1;
  // head scratch: pooled f32 [B,H] | pooled act [B,H] | h1 f32 [B,2H] | h1n act [B,2H] | logits f32 [B,A]
  const size_t head = static_cast<size_t>(B) * (H * 4 + H * es + 2 * H * 4 + 2 * H * es + (h->cfg.num_answers + 8) * 4);
  return h->headbuf.ensure(head + 1024);
}

int ensure_reduce_ws(rgqa_handle* h, DevBuf& part, int M, int N) {
  if (part.ensure(row_reduce_part_bytes(M, N))) return 1;
  const size_t need = static_cast<size_t>(row_reduce_counters(M) > 64 ? row_reduce_counters(M) : 64) * sizeof(int);
  if (need > h->red_cnt.bytes) {
    if (h->red_cnt.ensure(need * 2)) return 1;
    RGQA_CUDA_OK(cudaMemset(h->red_cnt.p, 0, h->red_cnt.bytes));  // the kernels leave every counter at zero again
  }
  return 0;
}

// Mahalanobis on the tensor cores: [t | G] = xs · [P; mu·P]^T as ONE split-bf16 GEMM (K' = 3H) whose epilogue forms
// q = <t, x> and max / arg-max_c (G_c - m_c/2); score = max - q/2 (gqa_maha.py:228-236,270).  The [B, A] matrix stays on chip.
int maha_tc(rgqa_handle* h, cudaStream_t st, const float* x, int B, float* score, int32_t* label) {
  const int H = h->cfg.hidden, A = h->cfg.num_answers;
  if (h->maha_xs.ensure(static_cast<size_t>(B) * 3 * H * 2) || ensure_reduce_ws(h, h->red_maha, B, H + A)) return 1;
  if (split_bf16x3(st, x, H, h->maha_xs.p, B, H, 0)) return 1;
  RowReduce r;
  r.kind = RED_MAHA;
  r.split = H;
  r.x = x;
  r.m = h->maha_m;
  r.part = static_cast<float*>(h->red_maha.p);
  r.counter = static_cast<int*>(h->red_cnt.p);
  r.scores = score;
  r.labels = label;
  return gemm_bf16_reduce(st, h->maha_xs.p, 3 * H, h->maha_Ws, 3 * H, nullptr, nullptr, 0, B, H + A, 3 * H, r);
}

int maha_core(cudaStream_t st, const float* x, const float* mu, const float* Psym, const float* m, int B, int A, int H,
              float* work, float* score, int32_t* label, float* full);

// pooler + answer head (+ fused scorers) on the hidden rows `hid` (first token of each sample, row stride S*H)
// `stats` != nullptr: `hid` is a deferred-LayerNorm state (un-normalised bf16 rows + statistics) and the pooler uses its
// folded weight.
int head_block(rgqa_handle* h, cudaStream_t st, const void* hid, int B, int S, const FwdOut& o, const float* stats = nullptr,
               int hid_f16 = 0 /* the rows are the fp16 single-word state */) {
  const int H = h->cfg.hidden, A = h->cfg.num_answers;
  const size_t es = act_size(h);
  const int dt = act_dtype(h);
  char* base = static_cast<char*>(h->headbuf.p);
  float* pooled = reinterpret_cast<float*>(base);
  char* pooled_act = base + static_cast<size_t>(B) * H * 4;
  float* h1 = reinterpret_cast<float*>(pooled_act + static_cast<size_t>(B) * H * es);
  char* h1n = reinterpret_cast<char*>(h1) + static_cast<size_t>(B) * 2 * H * 4;
  float* logits_ws = reinterpret_cast<float*>(h1n + static_cast<size_t>(B) * 2 * H * es);
  // 16-byte align the logits scratch
  logits_ws = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(logits_ws) + 15) & ~static_cast<uintptr_t>(15));
  const unsigned logit_scorers = o.score_mask & 0xFu;
  const bool want_maha = (o.score_mask >> RGQA_SCORE_MAHA) & 1u;
  RGQA_CHECK(o.score_mask == 0 || (o.scores != nullptr && o.labels != nullptr), "forward: scorers requested without score / label outputs");
  RGQA_CHECK(!want_maha || h->has_maha, "forward: Mahalanobis statistics were not set (rgqa_set_maha)");

  float* pdst = o.pooled != nullptr ? o.pooled : pooled;
  if (stats != nullptr) {
    if (gemm_bf16_folded(st, hid, static_cast<long long>(S) * H, h->pooler_f.W, H, h->pooler_f.b, nullptr, 0, pdst, H, 1, B, H,
                         H, EPI_TANH, stats, h->pooler_f.c, h->cfg.ln_eps, static_cast<long long>(S) * 12, hid_f16))
      return 1;
  } else if (hid_f16) {  // identity state straight into the pooler (a model without encoder layers)
    RGQA_CHECK(h->pooler_f.W != nullptr, "forward: fp16 copy of the pooler weight missing");
    if (gemm_bf16_folded(st, hid, static_cast<long long>(S) * H, h->pooler_f.W, H, h->pooler_f.b, nullptr, 0, pdst, H, 1, B, H, H, EPI_TANH,
                         nullptr, nullptr, 0.f, 12, 1))
      return 1;
  } else if (gemm(h, st, hid, static_cast<long long>(S) * H, h->pooler, nullptr, 0, pdst, H, 1, B, EPI_TANH)) {
    return 1;
  }
  if (want_maha) {  // scores row RGQA_SCORE_MAHA, labels row 1
    float* sc = o.scores + static_cast<size_t>(RGQA_SCORE_MAHA) * B;
    int32_t* lb = o.labels + B;
    if (h->maha_Ws != nullptr && option_get(OPT_MAHA_TC) != 0) {
      if (maha_tc(h, st, pdst, B, sc, lb)) return 1;
    } else {
      if (h->mahabuf.ensure((static_cast<size_t>(B) * (H + A + 1) + 16) * 4)) return 1;
      if (maha_core(st, pdst, h->maha_mu, h->maha_P, h->maha_m, B, A, H, static_cast<float*>(h->mahabuf.p), sc, lb, nullptr)) return 1;
    }
  }
  if (A <= 0 || (o.logits == nullptr && logit_scorers == 0)) return 0;
  const void* a0 = pdst;
  if (dt == DT_BF16) {
    if (cast_f32_to_bf16(st, pdst, pooled_act, static_cast<long long>(B) * H)) return 1;
    a0 = pooled_act;
  }
  if (gemm(h, st, a0, H, h->fc0, nullptr, 0, h1, 2 * H, 1, B, EPI_GELU)) return 1;
  if (layernorm(st, dt, h1, 2 * H, h->fc_ln.g, h->fc_ln.b, h1n, nullptr, B, 2 * H, h->cfg.ln_eps)) return 1;
  if (logit_scorers == 0) return gemm(h, st, h1n, 2 * H, h->fc3, nullptr, 0, o.logits, A, 1, B, EPI_BIAS);
  if (dt == DT_BF16) {
    // the answer-head GEMM reduces its own rows: MSP / energy / logsumexp leave the epilogue, the logits only if asked for
    if (ensure_reduce_ws(h, h->red_fc, B, A)) return 1;
    RowReduce r;
    r.kind = RED_LOGITS;
    r.flags = ((o.score_mask >> RGQA_SCORE_ENERGY_SUM) & 1u) | (((o.score_mask >> RGQA_SCORE_LSE) & 1u) << 1);
    r.part = static_cast<float*>(h->red_fc.p);
    r.counter = static_cast<int*>(h->red_cnt.p);
    r.scores = o.scores;
    r.labels = o.labels;
    return gemm_bf16_reduce(st, h1n, 2 * H, h->fc3.W, h->fc3.K, h->fc3.b, o.logits, A, B, A, h->fc3.K, r);
  }
  // fp32 handles: logits through the fp32 GEMM, then the row kernels
  float* lg = o.logits != nullptr ? o.logits : logits_ws;
  if (gemm(h, st, h1n, 2 * H, h->fc3, nullptr, 0, lg, A, 1, B, EPI_BIAS)) return 1;
  for (int mode = RGQA_SCORE_MSP; mode <= RGQA_SCORE_LSE; ++mode)
    if ((o.score_mask >> mode) & 1u)
      if (score_logits(st, mode, lg, A, B, A, o.scores + static_cast<size_t>(mode) * B, o.labels)) return 1;
  return 0;
}

}  // namespace

// =================================================================================================
// C ABI
// =================================================================================================
#pragma GCC visibility push(default)
extern "C" {

int rgqa_abi_version(void) { return RGQA_ABI_VERSION; }
const char* rgqa_last_error(void) { return g_err; }

int rgqa_create(const rgqa_config* cfg, rgqa_handle** out) {
  RGQA_CHECK(cfg != nullptr && out != nullptr, "rgqa_create: null argument");
  RGQA_CHECK(cfg->arch == RGQA_ARCH_LXMERT || cfg->arch == RGQA_ARCH_UNITER, "rgqa_create: unknown arch %d", cfg->arch);
  RGQA_CHECK(cfg->precision == RGQA_PREC_FP32 || cfg->precision == RGQA_PREC_BF16, "rgqa_create: unknown precision %d",
             cfg->precision);
  RGQA_CHECK(cfg->hidden == 768 && cfg->heads == 12, "rgqa_create: kernels are built for hidden=768, heads=12 (got %d, %d)",
             cfg->hidden, cfg->heads);
  RGQA_CHECK(cfg->intermediate % 64 == 0 && cfg->visual_feat_dim % 64 == 0, "rgqa_create: K dims must be multiples of 64");
  RGQA_CHECK(cfg->visual_pos_dim >= 1 && cfg->visual_pos_dim <= 8, "rgqa_create: visual_pos_dim out of range");
  RGQA_CHECK(cfg->num_rois >= 1 && cfg->num_rois <= 64, "rgqa_create: num_rois must be in [1,64]");
  RGQA_CHECK(cfg->l_layers >= 0 && cfg->x_layers >= 0 && cfg->r_layers >= 0, "rgqa_create: negative layer count");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  RGQA_CHECK(e == cudaSuccess && ndev > 0, "rgqa_create: no CUDA device (%s); this library has no CPU path",
             cudaGetErrorString(e));
  rgqa_handle* h = new rgqa_handle();
  h->cfg = *cfg;
  build_names(h);
  *out = h;
  return 0;
}

void rgqa_destroy(rgqa_handle* h) {
  if (h == nullptr) return;
  for (auto& kv : h->raw)
    if (kv.second.p != nullptr) cudaFree(kv.second.p);
  for (void* p : h->owned) cudaFree(p);
  for (auto& kv : h->graphs) {
    if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
    kv.second.feat.release(); kv.second.pos.release(); kv.second.ids.release(); kv.second.seg.release();
    kv.second.mask.release(); kv.second.pooled.release(); kv.second.logits.release();
    kv.second.scores.release(); kv.second.labels.release();
  }
  if (h->cap) cudaStreamDestroy(h->cap);
  if (h->side) cudaStreamDestroy(h->side);
  if (h->ev_fork) cudaEventDestroy(h->ev_fork);
  if (h->ev_join) cudaEventDestroy(h->ev_join);
  h->hidA.release();
  h->hidB.release();
  h->resA.release();
  h->resB.release();
  h->qkv.release();
  h->ctx.release();
  h->inter.release();
  h->pre.release();
  h->featb.release();
  h->headbuf.release();
  h->mahabuf.release();
  h->maha_xs.release();
  h->red_fc.release();
  h->red_maha.release();
  h->red_cnt.release();
  if (h->maha_Ws) cudaFree(h->maha_Ws);
  if (h->maha_mu) cudaFree(h->maha_mu);
  if (h->maha_P) cudaFree(h->maha_P);
  if (h->maha_m) cudaFree(h->maha_m);
  delete h;
}

int rgqa_num_weights(const rgqa_handle* h) { return h == nullptr ? 0 : static_cast<int>(h->names.size()); }
const char* rgqa_weight_name(const rgqa_handle* h, int index) {
  if (h == nullptr || index < 0 || index >= static_cast<int>(h->names.size())) return nullptr;
  return h->names[index].c_str();
}

int rgqa_load_weight(rgqa_handle* h, const char* ref_key, const float* data, int on_device, const int64_t* shape,
                     int ndim) {
  RGQA_CHECK(h != nullptr && ref_key != nullptr && data != nullptr, "rgqa_load_weight: null argument");
  RGQA_CHECK(ndim >= 1 && ndim <= 4, "rgqa_load_weight: bad ndim %d for '%s'", ndim, ref_key);
  const std::string key = strip_prefix(ref_key);
  bool expected = false;
  for (const auto& n : h->names)
    if (n == key) {
      expected = true;
      break;
    }
  RGQA_CHECK(expected, "rgqa_load_weight: '%s' is not a tensor of this model", ref_key);
  RawTensor t;
  t.numel = 1;
  for (int i = 0; i < ndim; ++i) {
    RGQA_CHECK(shape[i] > 0, "rgqa_load_weight: non-positive dim for '%s'", ref_key);
    t.shape.push_back(shape[i]);
    t.numel *= shape[i];
  }
  auto it = h->raw.find(key);
  if (it != h->raw.end()) {
    cudaFree(it->second.p);
    h->raw.erase(it);
  }
  RGQA_CUDA_OK(cudaMalloc(&t.p, static_cast<size_t>(t.numel) * 4));
  RGQA_CUDA_OK(cudaMemcpy(t.p, data, static_cast<size_t>(t.numel) * 4,
                          on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice));
  h->raw[key] = t;
  h->finalized = false;
  return 0;
}

int rgqa_finalize_weights(rgqa_handle* h) {
  RGQA_CHECK(h != nullptr, "rgqa_finalize_weights: null handle");
  h->max_fold_residue = 0.f;
  h->exact_folds = 0;
  for (const auto& n : h->names)
    RGQA_CHECK(h->raw.count(n) == 1, "rgqa_finalize_weights: missing tensor '%s'", n.c_str());
  drop_graphs(h);
  for (void* p : h->owned) cudaFree(p);
  h->owned.clear();
  h->l_layers.clear();
  h->r_layers.clear();
  h->x_layers.clear();
  const rgqa_config& c = h->cfg;
  const int H = c.hidden;
  if (c.arch == RGQA_ARCH_LXMERT) {
    const std::string b = "bert.";
    if (copy_raw(h, b + "embeddings.word_embeddings.weight", c.vocab_size, H, &h->word) ||
        copy_raw(h, b + "embeddings.position_embeddings.weight", c.max_position, H, &h->posemb) ||
        copy_raw(h, b + "embeddings.token_type_embeddings.weight", c.type_vocab, H, &h->typeemb) ||
        make_ln(h, b + "embeddings.LayerNorm", H, &h->emb_ln) ||
        make_linear(h, b + "encoder.visn_fc.visn_fc", H, c.visual_feat_dim, &h->visn_fc) ||
        make_ln(h, b + "encoder.visn_fc.visn_layer_norm", H, &h->visn_ln) ||
        copy_raw(h, b + "encoder.visn_fc.box_fc.weight", H, c.visual_pos_dim, &h->box_W) ||
        copy_raw(h, b + "encoder.visn_fc.box_fc.bias", H, 0, &h->box_b) ||
        make_ln(h, b + "encoder.visn_fc.box_layer_norm", H, &h->box_ln))
      return 1;
    h->l_layers.resize(c.l_layers);
    h->r_layers.resize(c.r_layers);
    h->x_layers.resize(c.x_layers);
    for (int i = 0; i < c.l_layers; ++i)
      if (make_bert_layer(h, b + "encoder.layer." + std::to_string(i), &h->l_layers[i])) return 1;
    for (int i = 0; i < c.r_layers; ++i)
      if (make_bert_layer(h, b + "encoder.r_layers." + std::to_string(i), &h->r_layers[i])) return 1;
    for (int i = 0; i < c.x_layers; ++i) {
      const std::string x = b + "encoder.x_layers." + std::to_string(i);
      XLayerW& w = h->x_layers[i];
      if (make_att(h, x + ".visual_attention", "att", &w.cross) || make_att(h, x + ".lang_self_att", "self", &w.lang_self) ||
          make_att(h, x + ".visn_self_att", "self", &w.visn_self) ||
          make_ffn(h, x + ".lang_inter", x + ".lang_output", &w.lang_ffn) ||
          make_ffn(h, x + ".visn_inter", x + ".visn_output", &w.visn_ffn))
        return 1;
    }
    if (make_linear(h, b + "pooler.dense", H, H, &h->pooler)) return 1;
    h->pooler_f = Folded{};
    if (c.precision == RGQA_PREC_BF16) {
      // one-kernel visual feature encoder: fp32 copy of visn_fc for the tf32 path, padded box_fc weight, closed-form statistics
      {
        std::vector<float> w2p, bst;
        if (copy_raw(h, b + "encoder.visn_fc.visn_fc.weight", H, c.visual_feat_dim, &h->visn_W32) ||
            make_box_stats(h->box_W, h->box_b, H, c.visual_pos_dim, &w2p, &bst) || upload_vec(h, w2p, &h->box_Wp) ||
            upload_vec(h, bst, &h->box_stat))
          return 1;
      }
      // deferred LayerNorm: every consumer of a hidden state gets a copy of its weight folded with the LayerNorm that
      // (notionally) produced that state; the embeddings' outputs are not behind one (nullptr)
      if (make_identity_ln(h)) return 1;
      h->state_f16 = option_get(OPT_STATE_F16) != 0 && H == 768;
      const LNp *pl = nullptr, *pv = nullptr;
      for (int i = 0; i < c.l_layers; ++i) {
        if (fold_bert_layer(h, b + "encoder.layer." + std::to_string(i), &h->l_layers[i], pl)) return 1;
        pl = &h->l_layers[i].ffn.ln;
      }
      for (int i = 0; i < c.r_layers; ++i) {
        if (fold_bert_layer(h, b + "encoder.r_layers." + std::to_string(i), &h->r_layers[i], pv)) return 1;
        pv = &h->r_layers[i].ffn.ln;
      }
      for (int i = 0; i < c.x_layers; ++i) {
        const std::string x = b + "encoder.x_layers." + std::to_string(i);
        XLayerW& w = h->x_layers[i];
        if (fold_qkv(h, x + ".visual_attention", "att", pl, &w.cross.qkv_f) ||
            fold_qkv(h, x + ".visual_attention", "att", pv, &w.cross.qkv_f2) ||
            fold_qkv(h, x + ".lang_self_att", "self", &w.cross.ln, &w.lang_self.qkv_f) ||
            fold_qkv(h, x + ".visn_self_att", "self", &w.cross.ln, &w.visn_self.qkv_f) ||
            fold_stack(h, {x + ".lang_inter.dense"}, c.intermediate, H, w.lang_self.ln, &w.lang_ffn.inter_f) ||
            fold_stack(h, {x + ".visn_inter.dense"}, c.intermediate, H, w.visn_self.ln, &w.visn_ffn.inter_f) ||
            (pl != nullptr && add_vecs(h, w.cross.out.b, pl->b, H, &w.cross.out_bb)) ||
            (pv != nullptr && add_vecs(h, w.cross.out.b, pv->b, H, &w.cross.out_bb2)) ||
            add_vecs(h, w.lang_self.out.b, w.cross.ln.b, H, &w.lang_self.out_bb) ||
            add_vecs(h, w.visn_self.out.b, w.cross.ln.b, H, &w.visn_self.out_bb) ||
            add_vecs(h, w.lang_ffn.out.b, w.lang_self.ln.b, H, &w.lang_ffn.out_bb) ||
            add_vecs(h, w.visn_ffn.out.b, w.visn_self.ln.b, H, &w.visn_ffn.out_bb))
          return 1;
        pl = &w.lang_ffn.ln;
        pv = &w.visn_ffn.ln;
      }
      if (pl != nullptr && fold_stack(h, {b + "pooler.dense"}, H, H, *pl, &h->pooler_f)) return 1;
      if (pl == nullptr && h->state_f16 && plain_f16_stack(h, {b + "pooler.dense"}, H, H, &h->pooler_f)) return 1;
      if (H == 768 && c.heads == 12) {
        for (auto& w : h->l_layers) if (pack_head_major(h, &w.att)) return 1;
        for (auto& w : h->r_layers) if (pack_head_major(h, &w.att)) return 1;
        for (auto& w : h->x_layers) if (pack_head_major(h, &w.lang_self) || pack_head_major(h, &w.visn_self)) return 1;
      }
    }
  } else {
    const std::string u = "uniter.";
    if (copy_raw(h, u + "embeddings.word_embeddings.weight", c.vocab_size, H, &h->word) ||
        copy_raw(h, u + "embeddings.position_embeddings.weight", c.max_position, H, &h->posemb) ||
        copy_raw(h, u + "embeddings.token_type_embeddings.weight", c.type_vocab, H, &h->typeemb) ||
        make_ln(h, u + "embeddings.LayerNorm", H, &h->emb_ln) ||
        make_linear(h, u + "img_embeddings.img_linear", H, c.visual_feat_dim, &h->visn_fc) ||
        make_ln(h, u + "img_embeddings.img_layer_norm", H, &h->visn_ln) ||
        copy_raw(h, u + "img_embeddings.pos_linear.weight", H, c.visual_pos_dim, &h->box_W) ||
        copy_raw(h, u + "img_embeddings.pos_linear.bias", H, 0, &h->box_b) ||
        make_ln(h, u + "img_embeddings.pos_layer_norm", H, &h->box_ln) ||
        make_ln(h, u + "img_embeddings.LayerNorm", H, &h->img_ln))
      return 1;
    h->l_layers.resize(c.l_layers);
    for (int i = 0; i < c.l_layers; ++i)
      if (make_bert_layer(h, u + "encoder.layer." + std::to_string(i), &h->l_layers[i])) return 1;
    if (make_linear(h, u + "pooler.dense", H, H, &h->pooler)) return 1;
    h->pooler_f = Folded{};
    if (c.precision == RGQA_PREC_BF16) {  // deferred LayerNorm: same folding as the LXMERT language stack
      if (make_identity_ln(h)) return 1;
      h->state_f16 = option_get(OPT_STATE_F16) != 0 && H == 768;
      const LNp* pl = nullptr;
      for (int i = 0; i < c.l_layers; ++i) {
        if (fold_bert_layer(h, u + "encoder.layer." + std::to_string(i), &h->l_layers[i], pl)) return 1;
        pl = &h->l_layers[i].ffn.ln;
      }
      if (pl != nullptr && fold_stack(h, {u + "pooler.dense"}, H, H, *pl, &h->pooler_f)) return 1;
      if (pl == nullptr && h->state_f16 && plain_f16_stack(h, {u + "pooler.dense"}, H, H, &h->pooler_f)) return 1;
      if (H == 768 && c.heads == 12)
        for (auto& w : h->l_layers) if (pack_head_major(h, &w.att)) return 1;
    }
  }
  if (c.num_answers > 0) {
    if (make_linear(h, "logit_fc.0", 2 * H, H, &h->fc0) || make_ln(h, "logit_fc.2", 2 * H, &h->fc_ln) ||
        make_linear(h, "logit_fc.3", c.num_answers, 2 * H, &h->fc3))
      return 1;
  }
  RGQA_CUDA_OK(cudaDeviceSynchronize());
  // the unpacked fp32 copies are no longer needed
  for (auto& kv : h->raw) {
    cudaFree(kv.second.p);
    kv.second.p = nullptr;
  }
  h->raw.clear();
  h->finalized = true;
  return 0;
}

// ---- deferred-LayerNorm forward (bf16, large batches) ----------------------------------------------------------------
// No LayerNorm kernel runs between the encoder GEMMs.  A hidden state is kept PRE-LayerNorm as a bf16 pair (hi + lo =
// the fp32 value to 2^-17) plus per-row partial (sum, sumsq); the LayerNorm itself is applied
//  - by the next GEMM that consumes the state as its A operand: gamma and the centring are folded into the weight, beta
//    into the bias, and 1/std is the scale of the epilogue's bias FMA (gemm_bf16_folded);
//  - by the next GEMM that adds the state as its residual: the epilogue normalises the hi/lo rows it loads (TMA) on the
//    fly, adds them to acc + bias and writes the new pre-LayerNorm state and its statistics (gemm_bf16_resid_hl).
// Same math as modeling.py:357-361 / :411-415, 53 fewer kernels and ~10 B per hidden element less HBM traffic per block.
namespace {
struct DState {
  char* hi;      // bf16 [Mt, H]; fp16 in the single-word form (DefCtx::f16), where it is the whole state
  char* lo;      // bf16 [Mt, H]; unused in the single-word form
  float* stats;  // [Mt, 6, 2]
};
struct DefCtx {
  rgqa_handle* h;
  int H;
  // single-word state (option "state_f16", default): the pre-LayerNorm rows are ONE fp16 word per element — A operand of the
  // consumers (fp16 x bf16 weights, kind::f16 takes the two formats independently) and residual carrier of the producers
  // (11 significand bits; a producer moves 4 608 instead of 7 680 B per row at K = 768)
  bool f16 = false;
  bool planes = false;  // Q|K|V projections are written head-major (gemm_bf16_qkv_planes) and read by attention_planes
  long long Mt = 0;     // rows of the plane tensor (all tokens of the batch)
  // Q|K|V of rows [r0, r0 + M) of `in` (LayerNorm ln still owed) into the shared qkv buffer, in the layout `planes` selects
  int qkv_proj(cudaStream_t st, const DState& in, const LNp* ln, const Linear& lin, const Folded& f, long long r0, int M) const {
    if (!planes) return consume(st, in, ln, lin, f, r0, M, rows_at(h->qkv.p, r0, 3 * H, 2), 3 * H, EPI_BIAS);
    RGQA_CHECK((ln == nullptr && !f16) || f.W != nullptr, "deferred LayerNorm: folded weight missing");
    return gemm_bf16_qkv_planes(st, hi_rows(in, r0), H, (ln || f16) ? f.W : lin.W, lin.K, (ln || f16) ? f.b : lin.b,
                                static_cast<char*>(h->qkv.p) + static_cast<size_t>(r0) * 64 * 2, Mt, h->cfg.heads, M, lin.K,
                                ln ? st_rows(in, r0) : nullptr, h->cfg.ln_eps, f16 ? 1 : 0);
  }
  // softmax(QK^T/8 + mask) V for B samples: queries rows [rq, rq + B*Lq), keys / values rows [rk, rk + B*Lk) of the qkv buffer
  int attn(cudaStream_t st, long long rq, long long rk, const int32_t* key_mask, void* ctx_rows, int B, int Lq, int Lk) const {
    if (planes) return attention_planes(st, h->qkv.p, Mt, rq, rk, key_mask, ctx_rows, H, B, Lq, Lk, h->cfg.heads);
    char* q = rows_at(h->qkv.p, rq, 3 * H, 2);
    char* kv = rows_at(h->qkv.p, rk, 3 * H, 2);
    return attention(st, DT_BF16, q, 3 * H, kv + static_cast<size_t>(H) * 2, 3 * H, kv + static_cast<size_t>(2 * H) * 2, 3 * H, key_mask,
                     ctx_rows, H, B, Lq, Lk, h->cfg.heads);
  }
  char* hi_rows(const DState& s, long long r0) const { return s.hi + static_cast<size_t>(r0) * H * 2; }
  char* lo_rows(const DState& s, long long r0) const { return s.lo + static_cast<size_t>(r0) * H * 2; }
  float* st_rows(const DState& s, long long r0) const { return s.stats + r0 * 12; }
  // out = epi(LN_ln(in) · Wᵀ + b); ln == nullptr: `in` holds already-normalised rows
  int consume(cudaStream_t st, const DState& in, const LNp* ln, const Linear& lin, const Folded& f, long long r0, int M,
              void* out, long long ldo, int epi) const {
    if (ln == nullptr) {  // identity state: plain weight (its fp16 copy when the state is fp16)
      RGQA_CHECK(!f16 || f.W != nullptr, "deferred LayerNorm: fp16 copy of the weight missing");
      return gemm_bf16_folded(st, hi_rows(in, r0), H, f16 ? f.W : lin.W, lin.K, f16 ? f.b : lin.b, nullptr, 0, out, ldo, 0, M, lin.N,
                              lin.K, epi, nullptr, nullptr, 0.f, 12, f16 ? 1 : 0);
    }
    RGQA_CHECK(f.W != nullptr, "deferred LayerNorm: folded weight missing");
    return gemm_bf16_folded(st, hi_rows(in, r0), H, f.W, lin.K, f.b, nullptr, 0, out, ldo, 0, M, lin.N, lin.K, epi,
                            st_rows(in, r0), f.c, h->cfg.ln_eps, 12, f16 ? 1 : 0);
  }
  // out (pre-LayerNorm) = A · Wᵀ + b + LN_ln(in); bias_beta = b + ln->b precomputed (nullptr: ln == nullptr, beta = 0)
  int produce(cudaStream_t st, const void* A, long long lda, const Linear& lin, const float* bias_beta, const DState& in,
              const LNp* ln, const DState& out, long long r0, int M) const {
    const float* bias = (ln != nullptr && bias_beta != nullptr) ? bias_beta : lin.b;
    const float* beta = (ln != nullptr && bias_beta == nullptr) ? ln->b : nullptr;
    if (f16)
      return gemm_bf16_resid_h16(st, A, lda, lin.W, lin.K, bias, hi_rows(in, r0), st_rows(in, r0), ln ? ln->g : h->ones, beta,
                                 h->cfg.ln_eps, hi_rows(out, r0), st_rows(out, r0), M, lin.N, lin.K);
    return gemm_bf16_resid_hl(st, A, lda, lin.W, lin.K, bias, hi_rows(in, r0), lo_rows(in, r0), st_rows(in, r0),
                              ln ? ln->g : h->ones, beta, h->cfg.ln_eps, hi_rows(out, r0), lo_rows(out, r0),
                              st_rows(out, r0), M, lin.N, lin.K);
  }
  // BertAttention (self) on rows [r0, r0 + B*S): in (LayerNorm ln_in) -> out (LayerNorm w.ln)
  int att(cudaStream_t st, const AttW& w, const DState& in, const LNp* ln_in, const DState& out, long long r0, int B, int S,
          const int32_t* key_mask) const {
    char* ctx = rows_at(h->ctx.p, r0, H, 2);
    if (!planes && w.qkv_hm != nullptr && qkv_attention_fused_ok(B, S, h->cfg.heads, H)) {
      // the projection tile is consumed where it is produced: no [rows, 2304] round trip, no attention launch
      if (qkv_attention_fused(st, hi_rows(in, r0), H, w.qkv_hm, w.qkv_hm_b, ln_in != nullptr ? st_rows(in, r0) : nullptr, h->cfg.ln_eps,
                              key_mask, ctx, H, B, S, h->cfg.heads, f16 ? 1 : 0))
        return 1;
    } else {
      if (qkv_proj(st, in, ln_in, w.qkv, w.qkv_f, r0, B * S)) return 1;
      if (attn(st, r0, r0, key_mask, ctx, B, S, S)) return 1;
    }
    return produce(st, ctx, H, w.out, w.out_bb, in, ln_in, out, r0, B * S);
  }
  // BertIntermediate + BertOutput: in (LayerNorm ln_in) -> out (LayerNorm w.ln)
  int ffn(cudaStream_t st, const FfnW& w, const DState& in, const LNp* ln_in, const DState& out, long long r0, int M) const {
    const int I = h->cfg.intermediate;
    char* inter = rows_at(h->inter.p, r0, I, 2);
    if (consume(st, in, ln_in, w.inter, w.inter_f, r0, M, inter, I, EPI_GELU)) return 1;
    return produce(st, inter, I, w.out, w.out_bb, in, ln_in, out, r0, M);
  }
};
}  // namespace

static int lxmert_forward_deferred(rgqa_handle* h, cudaStream_t st, const void* feat, int feat_dtype, const float* pos,
                                   const int32_t* ids, const int32_t* seg, const int32_t* mask, int B, int L,
                                   const FwdOut& out) {
  const rgqa_config& c = h->cfg;
  const int H = c.hidden, V = c.num_rois;
  const long long Ml = static_cast<long long>(B) * L, Mv = static_cast<long long>(B) * V, Mt = Ml + Mv;
  // resA / resB (fp32-sized) hold the lo rows in their first half and the statistics behind them
  DState cur{static_cast<char*>(h->hidA.p), static_cast<char*>(h->resA.p),
             reinterpret_cast<float*>(static_cast<char*>(h->resA.p) + static_cast<size_t>(Mt) * H * 2)};
  DState alt{static_cast<char*>(h->hidB.p), static_cast<char*>(h->resB.p),
             reinterpret_cast<float*>(static_cast<char*>(h->resB.p) + static_cast<size_t>(Mt) * H * 2)};
  DefCtx d{h, H};
  d.Mt = Mt;
  d.f16 = h->state_f16;
  const int sdt = d.f16 ? DT_F16 : DT_BF16;  // what the embedding kernels write into `hi`
  d.planes = option_get(OPT_QKV_PLANES) != 0 && attention_planes_ok(B, L, L, c.heads) && attention_planes_ok(B, V, V, c.heads) &&
             attention_planes_ok(B, L, V, c.heads) && attention_planes_ok(B, V, L, c.heads);

  const bool two = option_get(OPT_TWO_STREAMS) != 0;
  if (two && h->side == nullptr) {
    RGQA_CUDA_OK(cudaStreamCreateWithFlags(&h->side, cudaStreamNonBlocking));
    RGQA_CUDA_OK(cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
    RGQA_CUDA_OK(cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming));
  }
  cudaStream_t sv = two ? h->side : st;
  // option "sm_split": while both chains are in flight their persistent kernels get disjoint shares of the SMs (pairs), in
  // proportion to the rows they process, so that they run side by side instead of taking turns on the whole chip
  const bool split = two && option_get(OPT_SM_SPLIT) != 0;
  const int all = num_sms();
  auto share = [&](double frac) { int s = static_cast<int>(all * frac / 2.0 + 0.5) * 2; return s < 2 ? 2 : (s > all - 2 ? all - 2 : s); };
  // phase 1: c.l_layers language layers against c.r_layers visual layers; phase 2 (cross-modality layers): Ml against Mv rows
  const double w_l1 = static_cast<double>(c.l_layers) * Ml, w_v1 = static_cast<double>(c.r_layers) * Mv + 0.35 * Mv;
  const int sms_l1 = split && w_l1 > 0 && w_v1 > 0 ? share(w_l1 / (w_l1 + w_v1)) : 0, sms_v1 = sms_l1 ? all - sms_l1 : 0;
  const int sms_l2 = split ? share(static_cast<double>(Ml) / Mt) : 0, sms_v2 = sms_l2 ? all - sms_l2 : 0;
  struct BudgetReset { ~BudgetReset() { set_launch_sms(0); } } budget_reset;
  auto fork = [&]() -> int {
    if (!two) return 0;
    RGQA_CUDA_OK(cudaEventRecord(h->ev_fork, st));
    RGQA_CUDA_OK(cudaStreamWaitEvent(sv, h->ev_fork, 0));
    return 0;
  };
  auto join = [&]() -> int {
    if (!two) return 0;
    RGQA_CUDA_OK(cudaEventRecord(h->ev_join, sv));
    RGQA_CUDA_OK(cudaStreamWaitEvent(st, h->ev_join, 0));
    return 0;
  };
  if (fork()) return 1;

  // ---- embeddings (modeling.py:876) and visual feature encoder (:507-517): normalised rows -> identity states ----
  // the row kernels emit the identity state themselves (lo = bf16 remainder, statistics (0, 768)): no fp32 round trip
  set_launch_sms(sms_v1);   // the visual encoder and the relational layers (side stream)
  if (embed_ln(st, sdt, ids, seg, h->word, h->posemb, h->typeemb, h->emb_ln.g, h->emb_ln.b, cur.hi, nullptr, B, L, H,
               c.vocab_size, c.max_position, c.type_vocab, c.ln_eps, d.f16 ? nullptr : cur.lo, cur.stats))
    return 1;
  if (option_get(OPT_VISN_FUSED) != 0 && h->box_stat != nullptr && c.visual_feat_dim % 64 == 0) {
    // ONE kernel: projection (tf32 straight from fp32 features, or bf16), both LayerNorms, box branch, average, hi/lo state
    const bool f32 = feat_dtype == RGQA_DT_F32;
    if (visn_embed(sv, feat, f32 ? 1 : 0, c.visual_feat_dim, f32 ? static_cast<const void*>(h->visn_W32) : h->visn_fc.W, h->visn_fc.b,
                   h->visn_ln.g, h->visn_ln.b, pos, c.visual_pos_dim, h->box_Wp, h->box_b, h->box_ln.g, h->box_ln.b, h->box_stat,
                   c.ln_eps, d.hi_rows(cur, Ml), d.f16 ? nullptr : d.lo_rows(cur, Ml), d.st_rows(cur, Ml), static_cast<int>(Mv),
                   c.visual_feat_dim))
      return 1;
  } else {
    const void* a = feat;
    if (feat_dtype == RGQA_DT_F32) {
      if (cast_f32_to_bf16(sv, static_cast<const float*>(feat), h->featb.p, Mv * c.visual_feat_dim)) return 1;
      a = h->featb.p;
    }
    float* pre = reinterpret_cast<float*>(rows_at(h->pre.p, Ml, H, 4));
    if (gemm_bf16(sv, a, c.visual_feat_dim, h->visn_fc.W, h->visn_fc.K, h->visn_fc.b, nullptr, 0, pre, H, 1, static_cast<int>(Mv),
                  H, h->visn_fc.K, EPI_BIAS) ||
        visn_combine(sv, sdt, pre, pos, c.visual_pos_dim, h->box_W, h->box_b, h->visn_ln.g, h->visn_ln.b, h->box_ln.g,
                     h->box_ln.b, d.hi_rows(cur, Ml), nullptr, static_cast<int>(Mv), H, c.ln_eps, d.f16 ? nullptr : d.lo_rows(cur, Ml),
                     d.st_rows(cur, Ml)))
      return 1;
  }
  const LNp *ln_l = nullptr, *ln_v = nullptr;  // LayerNorm still owed to the language / visual rows of `cur`
  set_launch_sms(sms_l1);
  for (int i = 0; i < c.l_layers; ++i) {  // modeling.py:553-554
    const LayerW& w = h->l_layers[i];
    if (d.att(st, w.att, cur, ln_l, alt, 0, B, L, mask) || d.ffn(st, w.ffn, alt, &w.att.ln, cur, 0, static_cast<int>(Ml)))
      return 1;
    ln_l = &w.ffn.ln;
  }
  set_launch_sms(sms_v1);
  for (int i = 0; i < c.r_layers; ++i) {  // modeling.py:557-558
    const LayerW& w = h->r_layers[i];
    if (d.att(sv, w.att, cur, ln_v, alt, Ml, B, V, nullptr) || d.ffn(sv, w.ffn, alt, &w.att.ln, cur, Ml, static_cast<int>(Mv)))
      return 1;
    ln_v = &w.ffn.ln;
  }
  set_launch_sms(0);
  if (join()) return 1;
  for (int i = 0; i < c.x_layers; ++i) {  // modeling.py:477-488
    const XLayerW& w = h->x_layers[i];
    char* ctx = static_cast<char*>(h->ctx.p);
    // ONE BertCrossattLayer for both directions (:455-459); its Q|K|V of each modality behind that modality's LayerNorm
    if (fork()) return 1;
    set_launch_sms(sms_l2);
    if (d.qkv_proj(st, cur, ln_l, w.cross.qkv, w.cross.qkv_f, 0, static_cast<int>(Ml))) return 1;
    set_launch_sms(sms_v2);
    if (d.qkv_proj(sv, cur, ln_v, w.cross.qkv, w.cross.qkv_f2, Ml, static_cast<int>(Mv))) return 1;
    set_launch_sms(0);
    if (join()) return 1;
    // language queries over visual keys (no mask), visual queries over language keys (language mask)
    if (d.attn(st, 0, Ml, nullptr, ctx, B, L, V) || d.attn(st, Ml, 0, mask, rows_at(ctx, Ml, H, 2), B, V, L)) return 1;
    if (fork()) return 1;
    // cross output + self attention + FFN per modality (:455-475), the two chains side by side
    set_launch_sms(sms_l2);
    if (d.produce(st, ctx, H, w.cross.out, w.cross.out_bb, cur, ln_l, alt, 0, static_cast<int>(Ml)) ||
        d.att(st, w.lang_self, alt, &w.cross.ln, cur, 0, B, L, mask) ||
        d.ffn(st, w.lang_ffn, cur, &w.lang_self.ln, alt, 0, static_cast<int>(Ml)))
      return 1;
    set_launch_sms(sms_v2);
    if (d.produce(sv, rows_at(ctx, Ml, H, 2), H, w.cross.out, w.cross.out_bb2, cur, ln_v, alt, Ml, static_cast<int>(Mv)) ||
        d.att(sv, w.visn_self, alt, &w.cross.ln, cur, Ml, B, V, nullptr) ||
        d.ffn(sv, w.visn_ffn, cur, &w.visn_self.ln, alt, Ml, static_cast<int>(Mv)))
      return 1;
    set_launch_sms(0);
    if (join()) return 1;
    ln_l = &w.lang_ffn.ln;
    ln_v = &w.visn_ffn.ln;
    const DState t = cur;
    cur = alt;
    alt = t;
  }
  if (!out.any()) return 0;
  if (ln_l == nullptr) return head_block(h, st, cur.hi, B, L, out, nullptr, d.f16 ? 1 : 0);
  return head_block(h, st, cur.hi, B, L, out, cur.stats, d.f16 ? 1 : 0);
}

static int lxmert_forward_impl(rgqa_handle* h, cudaStream_t st, const void* feat, int feat_dtype, const float* pos,
                               const int32_t* ids, const int32_t* seg, const int32_t* mask, int B, int L,
                               const FwdOut& out) {
  const rgqa_config& c = h->cfg;
  const int H = c.hidden, V = c.num_rois;
  const long long Ml = static_cast<long long>(B) * L, Mv = static_cast<long long>(B) * V, Mt = Ml + Mv;
  RGQA_CHECK(Mt < (1LL << 30), "rgqa_lxmert_forward: batch too large");
  if (ensure_workspace(h, Mt, Mv, B)) return 1;
  const size_t es = act_size(h);
  const int dt = act_dtype(h);
  const bool bf = c.precision == RGQA_PREC_BF16;
  // large bf16 batches: no LayerNorm kernels at all (every encoder GEMM then runs on the CTA-pair kernel)
  if (bf && option_get(OPT_DEFER_LN) != 0 && option_get(OPT_GEMM_PAIR) != 0 && option_get(OPT_GEMM_STAGED) != 0 &&
      H == 768 && Ml >= 1024 && h->ones != nullptr)
    return lxmert_forward_deferred(h, st, feat, feat_dtype, pos, ids, seg, mask, B, L, out);
  Hid cur{h->hidA.p, bf ? static_cast<float*>(h->resA.p) : static_cast<float*>(h->hidA.p)};
  Hid alt{h->hidB.p, bf ? static_cast<float*>(h->resB.p) : static_cast<float*>(h->hidB.p)};

  // The language rows [0, Ml) and the visual rows [Ml, Mt) are independent up to the first cross-modality layer
  // and again inside each of them after the shared cross attention; the visual branch runs on a second stream
  // (fork / join with events) so one branch's kernels fill the other's tail waves.  Every workspace buffer is
  // indexed by row, so the two branches never touch the same bytes.
  const bool two = option_get(OPT_TWO_STREAMS) != 0 && B >= 64;
  if (two && h->side == nullptr) {
    RGQA_CUDA_OK(cudaStreamCreateWithFlags(&h->side, cudaStreamNonBlocking));
    RGQA_CUDA_OK(cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
    RGQA_CUDA_OK(cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming));
  }
  cudaStream_t sv = two ? h->side : st;  // stream of the visual branch
  auto fork = [&]() -> int {
    if (!two) return 0;
    RGQA_CUDA_OK(cudaEventRecord(h->ev_fork, st));
    RGQA_CUDA_OK(cudaStreamWaitEvent(sv, h->ev_fork, 0));
    return 0;
  };
  auto join = [&]() -> int {
    if (!two) return 0;
    RGQA_CUDA_OK(cudaEventRecord(h->ev_join, sv));
    RGQA_CUDA_OK(cudaStreamWaitEvent(st, h->ev_join, 0));
    return 0;
  };
  if (fork()) return 1;

  // ---- embeddings (modeling.py:876) ----
  if (embed_ln(st, dt, ids, seg, h->word, h->posemb, h->typeemb, h->emb_ln.g, h->emb_ln.b, cur.act, res_copy_rows(h, cur, 0), B, L, H, c.vocab_size,
               c.max_position, c.type_vocab, c.ln_eps))
    return 1;
  // ---- visual feature encoder (modeling.py:507-517) ----
  {
    const void* a = feat;
    if (c.precision == RGQA_PREC_BF16) {
      if (feat_dtype == RGQA_DT_F32) {
        if (cast_f32_to_bf16(sv, static_cast<const float*>(feat), h->featb.p, Mv * c.visual_feat_dim)) return 1;
        a = h->featb.p;
      }
    } else {
      RGQA_CHECK(feat_dtype == RGQA_DT_F32, "fp32 path needs fp32 features");
    }
    float* pre = reinterpret_cast<float*>(rows_at(h->pre.p, Ml, H, 4));
    if (gemm(h, sv, a, c.visual_feat_dim, h->visn_fc, nullptr, 0, pre, H, 1, static_cast<int>(Mv), EPI_BIAS)) return 1;
    if (visn_combine(sv, dt, pre, pos, c.visual_pos_dim, h->box_W, h->box_b, h->visn_ln.g, h->visn_ln.b, h->box_ln.g,
                     h->box_ln.b, rows_at(cur.act, Ml, H, es), res_copy_rows(h, cur, Ml), static_cast<int>(Mv), H, c.ln_eps))
      return 1;
  }
  // ---- language layers (modeling.py:553-554): rows [0, Ml) ----
  for (int i = 0; i < c.l_layers; ++i) {
    if (att_block(h, st, h->l_layers[i].att, cur, alt, 0, B, L, mask)) return 1;
    if (ffn_block(h, st, h->l_layers[i].ffn, alt, cur, 0, static_cast<int>(Ml))) return 1;
  }
  // ---- object-relational layers (modeling.py:557-558): rows [Ml, Mt), no key mask (entry.py:109) ----
  for (int i = 0; i < c.r_layers; ++i) {
    if (att_block(h, sv, h->r_layers[i].att, cur, alt, Ml, B, V, nullptr)) return 1;
    if (ffn_block(h, sv, h->r_layers[i].ffn, alt, cur, Ml, static_cast<int>(Mv))) return 1;
  }
  if (join()) return 1;
  // ---- cross-modality layers (modeling.py:477-488) ----
  for (int i = 0; i < c.x_layers; ++i) {
    const XLayerW& w = h->x_layers[i];
    const int of32 = dt == DT_F32;
    char* qkv = static_cast<char*>(h->qkv.p);
    char* qkv_v = rows_at(h->qkv.p, Ml, 3 * H, es);
    // cross attention: ONE BertCrossattLayer serves both directions on the pre-update inputs (:455-459)
    if (gemm(h, st, cur.act, H, w.cross.qkv, nullptr, 0, qkv, 3 * H, of32, static_cast<int>(Mt), EPI_BIAS)) return 1;
    // lang queries attend visual keys/values (no mask)
    if (attention(st, dt, qkv, 3 * H, qkv_v + static_cast<size_t>(H) * es, 3 * H, qkv_v + static_cast<size_t>(2 * H) * es,
                  3 * H, nullptr, h->ctx.p, H, B, L, V, c.heads))
      return 1;
    // visual queries attend language keys/values (language mask)
    if (attention(st, dt, qkv_v, 3 * H, qkv + static_cast<size_t>(H) * es, 3 * H, qkv + static_cast<size_t>(2 * H) * es,
                  3 * H, mask, rows_at(h->ctx.p, Ml, H, es), H, B, V, L, c.heads))
      return 1;
    if (dense_resid_ln(h, st, h->ctx.p, H, w.cross.out, w.cross.ln, cur, alt, 0, static_cast<int>(Mt))) return 1;
    // self attention (:461-465, alt -> cur) and FFN (:467-475, cur -> alt) per modality, the two chains side by side
    if (fork()) return 1;
    if (att_block(h, st, w.lang_self, alt, cur, 0, B, L, mask)) return 1;
    if (ffn_block(h, st, w.lang_ffn, cur, alt, 0, static_cast<int>(Ml))) return 1;
    if (att_block(h, sv, w.visn_self, alt, cur, Ml, B, V, nullptr)) return 1;
    if (ffn_block(h, sv, w.visn_ffn, cur, alt, Ml, static_cast<int>(Mv))) return 1;
    if (join()) return 1;
    const Hid t = cur;
    cur = alt;
    alt = t;
  }
  // ---- pooler + answer head (modeling.py:884, gqa_model.py:40-41) ----
  if (!out.any()) return 0;
  return head_block(h, st, cur.act, B, L, out);
}

// Small batches are launch-bound (170-210 kernels + ~330 tensor-map encodes per forward: ~3 ms of host time at B = 32,
// the reference scripts' default batch size).  For B <= 128 the whole forward is captured ONCE per (B, L, dtype, outputs)
// into a CUDA graph over handle-owned staging buffers; later calls are: copy inputs in, one graph launch, copy outputs
// out.  Large batches are GPU-bound on an idle host, but with one process per GPU the 8 ranks of a box share its cores
// and the ~2 ms of launch work per step starts to show (measured: 14.3 vs 12.6 ms/step at N = 8): they are captured too,
// keyed additionally by the caller's feature / box pointers, which are then used in place (a sweep feeds them from two
// alternating staging buffers); only the small id / mask inputs and the outputs go through staging copies.
// The first call of a key runs eagerly (lazy per-kernel attribute setup must not happen under capture).
static int lxmert_forward_graphed(rgqa_handle* h, void* stream, const void* feat, int feat_dtype, const float* pos,
                                 const int32_t* ids, const int32_t* seg, const int32_t* mask, int B, int L, const FwdOut& out) {
  RGQA_CHECK(h != nullptr, "rgqa_lxmert_forward: null handle");
  RGQA_CHECK(h->cfg.arch == RGQA_ARCH_LXMERT, "rgqa_lxmert_forward: handle was created for another architecture");
  RGQA_CHECK(h->finalized, "rgqa_lxmert_forward: weights are not finalized");
  RGQA_CHECK(feat != nullptr && pos != nullptr && ids != nullptr && mask != nullptr, "rgqa_lxmert_forward: null input");
  RGQA_CHECK(B > 0, "rgqa_lxmert_forward: empty batch");
  RGQA_CHECK(L >= 2 && L <= 64, "rgqa_lxmert_forward: L=%d out of range [2,64]", L);
  RGQA_CHECK((out.logits == nullptr && (out.score_mask & 0xFu) == 0) || h->cfg.num_answers > 0, "rgqa_lxmert_forward: handle has no answer head");
  RGQA_CHECK(out.score_mask == 0 || (out.scores != nullptr && out.labels != nullptr), "rgqa_forward_scored: null score / label output");
  RGQA_CHECK((out.score_mask >> (RGQA_SCORE_MAHA + 1)) == 0, "rgqa_forward_scored: unknown scorer bits 0x%x", out.score_mask);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const bool use_graph = option_get(OPT_GRAPHS) != 0 && !prof_on();
  const bool big = B > 128;  // large batch: feat / pos are used in place
  if (!use_graph) return lxmert_forward_impl(h, st, feat, feat_dtype, pos, ids, seg, mask, B, L, out);

  const rgqa_config& c = h->cfg;
  const long long key = (static_cast<long long>(B) << 20) | (static_cast<long long>(L) << 8) | (feat_dtype << 4) |
                        ((option_get(OPT_DEFER_LN) != 0) << 3) | (static_cast<long long>(option_get(OPT_QKV_PLANES) != 0) << 40) |
                        (static_cast<long long>(option_get(OPT_ATTN_TMA) != 0) << 41) |
                        (static_cast<long long>(option_get(OPT_MAHA_TC) != 0) << 42) |
                        (static_cast<long long>(option_get(OPT_VISN_FUSED) != 0) << 43) |
                        (static_cast<long long>(option_get(OPT_ATTN_FUSED) != 0) << 44) |

                        (static_cast<long long>(out.score_mask & 0x1Fu) << 48) |
                        ((seg != nullptr) << 2) | ((out.pooled != nullptr) << 1) | (out.logits != nullptr);
  if (h->graphs.size() > 16) {  // callers that never repeat a pointer: do not hoard graphs
    drop_graphs(h);
    for (auto& kv : h->graphs) {
      kv.second.feat.release(); kv.second.pos.release(); kv.second.ids.release(); kv.second.seg.release();
      kv.second.mask.release(); kv.second.pooled.release(); kv.second.logits.release();
      kv.second.scores.release(); kv.second.labels.release();
    }
    h->graphs.clear();
  }
  rgqa_handle::GraphEntryT& g = h->graphs[rgqa_handle::GraphKey{key, big ? feat : nullptr, big ? static_cast<const void*>(pos) : nullptr}];
  const size_t fe = feat_dtype == RGQA_DT_F32 ? 4 : 2;
  const size_t n_feat = static_cast<size_t>(B) * c.num_rois * c.visual_feat_dim * fe;
  const size_t n_pos = static_cast<size_t>(B) * c.num_rois * c.visual_pos_dim * 4;
  const size_t n_ids = static_cast<size_t>(B) * L * 4;
  if (g.seen == 0) {  // first sight of this shape: eager run (also grows the workspace), capture on the next call
    g.seen = 1;
    return lxmert_forward_impl(h, st, feat, feat_dtype, pos, ids, seg, mask, B, L, out);
  }
  if (g.exec == nullptr) {
    if ((!big && (g.feat.ensure(n_feat) || g.pos.ensure(n_pos))) || g.ids.ensure(n_ids) || g.seg.ensure(n_ids) || g.mask.ensure(n_ids) ||
        g.pooled.ensure(static_cast<size_t>(B) * c.hidden * 4) ||
        g.logits.ensure(static_cast<size_t>(B) * (c.num_answers > 0 ? c.num_answers : 1) * 4) ||
        g.scores.ensure(static_cast<size_t>(B) * 5 * 4) || g.labels.ensure(static_cast<size_t>(B) * 2 * 4))
      return 1;
    cudaGraph_t graph = nullptr;
    const long long n0 = launch_count();
    // capture on a handle-owned stream: the caller's stream is usually the legacy default stream, which cannot capture
    if (h->cap == nullptr) RGQA_CUDA_OK(cudaStreamCreateWithFlags(&h->cap, cudaStreamNonBlocking));
    FwdOut go;
    go.pooled = out.pooled != nullptr ? static_cast<float*>(g.pooled.p) : nullptr;
    go.logits = out.logits != nullptr ? static_cast<float*>(g.logits.p) : nullptr;
    go.score_mask = out.score_mask;
    go.scores = static_cast<float*>(g.scores.p);
    go.labels = static_cast<int32_t*>(g.labels.p);
    RGQA_CUDA_OK(cudaStreamBeginCapture(h->cap, cudaStreamCaptureModeThreadLocal));
    const int rc = lxmert_forward_impl(h, h->cap, big ? feat : g.feat.p, feat_dtype, big ? pos : static_cast<const float*>(g.pos.p),
                                       static_cast<const int32_t*>(g.ids.p),
                                       seg != nullptr ? static_cast<const int32_t*>(g.seg.p) : nullptr,
                                       static_cast<const int32_t*>(g.mask.p), B, L, go);
    const cudaError_t ce = cudaStreamEndCapture(h->cap, &graph);
    g.kernels = static_cast<int>(launch_count() - n0);
    note_launch(-g.kernels);  // captured, not launched: each replay is counted below
    if (rc != 0) {
      if (graph) cudaGraphDestroy(graph);
      return 1;
    }
    RGQA_CUDA_OK(ce);
    RGQA_CUDA_OK(cudaGraphInstantiate(&g.exec, graph, 0));
    RGQA_CUDA_OK(cudaGraphDestroy(graph));
    ++h->graph_captures;
  }
  if (!big) {
    RGQA_CUDA_OK(cudaMemcpyAsync(g.feat.p, feat, n_feat, cudaMemcpyDeviceToDevice, st));
    RGQA_CUDA_OK(cudaMemcpyAsync(g.pos.p, pos, n_pos, cudaMemcpyDeviceToDevice, st));
  }
  RGQA_CUDA_OK(cudaMemcpyAsync(g.ids.p, ids, n_ids, cudaMemcpyDeviceToDevice, st));
  if (seg != nullptr) RGQA_CUDA_OK(cudaMemcpyAsync(g.seg.p, seg, n_ids, cudaMemcpyDeviceToDevice, st));
  RGQA_CUDA_OK(cudaMemcpyAsync(g.mask.p, mask, n_ids, cudaMemcpyDeviceToDevice, st));
  RGQA_CUDA_OK(cudaGraphLaunch(g.exec, st));
  note_launch(g.kernels);
  if (out.pooled != nullptr)
    RGQA_CUDA_OK(cudaMemcpyAsync(out.pooled, g.pooled.p, static_cast<size_t>(B) * c.hidden * 4, cudaMemcpyDeviceToDevice, st));
  if (out.logits != nullptr)
    RGQA_CUDA_OK(cudaMemcpyAsync(out.logits, g.logits.p, static_cast<size_t>(B) * c.num_answers * 4, cudaMemcpyDeviceToDevice, st));
  if (out.score_mask != 0) {  // only the rows that were computed
    for (int m = RGQA_SCORE_MSP; m <= RGQA_SCORE_MAHA; ++m)
      if ((out.score_mask >> m) & 1u)
        RGQA_CUDA_OK(cudaMemcpyAsync(out.scores + static_cast<size_t>(m) * B, static_cast<float*>(g.scores.p) + static_cast<size_t>(m) * B,
                                     static_cast<size_t>(B) * 4, cudaMemcpyDeviceToDevice, st));
    if (out.score_mask & 0xFu)
      RGQA_CUDA_OK(cudaMemcpyAsync(out.labels, g.labels.p, static_cast<size_t>(B) * 4, cudaMemcpyDeviceToDevice, st));
    if ((out.score_mask >> RGQA_SCORE_MAHA) & 1u)
      RGQA_CUDA_OK(cudaMemcpyAsync(out.labels + B, static_cast<int32_t*>(g.labels.p) + B, static_cast<size_t>(B) * 4, cudaMemcpyDeviceToDevice, st));
  }
  return 0;
}

int rgqa_lxmert_forward(rgqa_handle* h, void* stream, const void* feat, int feat_dtype, const float* pos,
                        const int32_t* ids, const int32_t* seg, const int32_t* mask, int B, int L,
                        float* pooled_out, float* logits_out) {
  FwdOut out;
  out.pooled = pooled_out;
  out.logits = logits_out;
  return lxmert_forward_graphed(h, stream, feat, feat_dtype, pos, ids, seg, mask, B, L, out);
}

static int uniter_forward_impl(rgqa_handle* h, void* stream, const void* feat, int feat_dtype, const float* pos7,
                               const int32_t* ids, const int32_t* seg, const int32_t* mask, int B, int L, const FwdOut& out);

// Forward + rejection scorers in one call (north_star item 3): the answer-head GEMM's epilogue reduces its own logit rows and
// the Mahalanobis scorer runs on the tensor cores, so neither [B, A] matrix has to reach HBM (logits_out may be NULL).
int rgqa_forward_scored(rgqa_handle* h, void* stream, const void* feat, int feat_dtype, const float* pos,
                        const int32_t* ids, const int32_t* seg, const int32_t* mask, int B, int L, unsigned score_mask,
                        float* pooled_out, float* logits_out, float* scores_out, int32_t* labels_out) {
  RGQA_CHECK(h != nullptr, "rgqa_forward_scored: null handle");
  FwdOut out;
  out.pooled = pooled_out;
  out.logits = logits_out;
  out.score_mask = score_mask;
  out.scores = scores_out;
  out.labels = labels_out;
  if (h->cfg.arch == RGQA_ARCH_UNITER) return uniter_forward_impl(h, stream, feat, feat_dtype, pos, ids, seg, mask, B, L, out);
  return lxmert_forward_graphed(h, stream, feat, feat_dtype, pos, ids, seg, mask, B, L, out);
}

long long rgqa_graph_captures(const rgqa_handle* h) { return h == nullptr ? 0 : h->graph_captures; }
float rgqa_fold_residue(const rgqa_handle* h, int* exact_consumers) {
  if (h == nullptr) return 0.f;
  if (exact_consumers != nullptr) *exact_consumers = h->exact_folds;
  return h->max_fold_residue;
}

int rgqa_uniter_forward(rgqa_handle* h, void* stream, const void* feat, int feat_dtype, const float* pos7,
                        const int32_t* ids, const int32_t* seg, const int32_t* mask, int B, int L,
                        float* pooled_out, float* logits_out) {
  FwdOut out;
  out.pooled = pooled_out;
  out.logits = logits_out;
  return uniter_forward_impl(h, stream, feat, feat_dtype, pos7, ids, seg, mask, B, L, out);
}

static int uniter_forward_impl(rgqa_handle* h, void* stream, const void* feat, int feat_dtype, const float* pos7,
                               const int32_t* ids, const int32_t* seg, const int32_t* mask, int B, int L, const FwdOut& out) {
  RGQA_CHECK(h != nullptr, "rgqa_uniter_forward: null handle");
  RGQA_CHECK(h->cfg.arch == RGQA_ARCH_UNITER, "rgqa_uniter_forward: handle was created for another architecture");
  RGQA_CHECK(h->finalized, "rgqa_uniter_forward: weights are not finalized");
  RGQA_CHECK(feat != nullptr && pos7 != nullptr && ids != nullptr && mask != nullptr, "rgqa_uniter_forward: null input");
  RGQA_CHECK(B > 0, "rgqa_uniter_forward: empty batch");
  const rgqa_config& c = h->cfg;
  const int H = c.hidden, V = c.num_rois;
  const int S = L + V;
  RGQA_CHECK(L >= 2 && S <= 64, "rgqa_uniter_forward: L + num_rois = %d exceeds 64", S);
  RGQA_CHECK(c.type_vocab >= 2, "rgqa_uniter_forward: needs token type 1 for image tokens");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long long Ml = static_cast<long long>(B) * L, Mv = static_cast<long long>(B) * V, Mt = Ml + Mv;
  if (ensure_workspace(h, Mt, Mv, B)) return 1;
  // joint key mask [B, S] = [text mask | ones] lives at the tail of the head scratch
  const size_t es = act_size(h);
  const int dt = act_dtype(h);
  const bool bf = c.precision == RGQA_PREC_BF16;
  if (h->mahabuf.ensure(static_cast<size_t>(B) * S * 4 + 256)) return 1;
  int32_t* jmask = static_cast<int32_t*>(h->mahabuf.p);

  // text embeddings -> ctx buffer rows [0, Ml) (scratch), image embeddings -> ctx rows [Ml, Mt),
  // then interleave per sample into hidA as [B, S, H]
  if (embed_ln(st, dt, ids, seg, h->word, h->posemb, h->typeemb, h->emb_ln.g, h->emb_ln.b, h->ctx.p, bf ? static_cast<float*>(h->pre.p) : nullptr, B, L, H, c.vocab_size,
               c.max_position, c.type_vocab, c.ln_eps))
    return 1;
  {
    const void* a = feat;
    if (c.precision == RGQA_PREC_BF16) {
      if (feat_dtype == RGQA_DT_F32) {
        if (cast_f32_to_bf16(st, static_cast<const float*>(feat), h->featb.p, Mv * c.visual_feat_dim)) return 1;
        a = h->featb.p;
      }
    } else {
      RGQA_CHECK(feat_dtype == RGQA_DT_F32, "fp32 path needs fp32 features");
    }
    float* pre = reinterpret_cast<float*>(rows_at(h->pre.p, Ml, H, 4));
    if (gemm(h, st, a, c.visual_feat_dim, h->visn_fc, nullptr, 0, pre, H, 1, static_cast<int>(Mv), EPI_BIAS)) return 1;
    // image tokens carry token type 1 (uniter/entry.py:92, modeling.py:629)
    if (uniter_img_combine(st, dt, pre, pos7, c.visual_pos_dim, h->box_W, h->box_b, h->visn_ln.g, h->visn_ln.b,
                           h->box_ln.g, h->box_ln.b, h->typeemb + H, h->img_ln.g, h->img_ln.b,
                           rows_at(h->ctx.p, Ml, H, es), bf ? pre : nullptr, static_cast<int>(Mv), H, c.ln_eps))
      return 1;
  }
  Hid cur{h->hidA.p, bf ? static_cast<float*>(h->resA.p) : static_cast<float*>(h->hidA.p)};
  Hid alt{h->hidB.p, bf ? static_cast<float*>(h->resB.p) : static_cast<float*>(h->hidB.p)};
  // per-sample interleave: 2-D strided copies on the stream (rows of one sample are contiguous)
  RGQA_CUDA_OK(cudaMemcpy2DAsync(cur.act, static_cast<size_t>(S) * H * es, h->ctx.p, static_cast<size_t>(L) * H * es,
                                 static_cast<size_t>(L) * H * es, B, cudaMemcpyDeviceToDevice, st));
  RGQA_CUDA_OK(cudaMemcpy2DAsync(static_cast<char*>(cur.act) + static_cast<size_t>(L) * H * es,
                                 static_cast<size_t>(S) * H * es, rows_at(h->ctx.p, Ml, H, es),
                                 static_cast<size_t>(V) * H * es, static_cast<size_t>(V) * H * es, B,
                                 cudaMemcpyDeviceToDevice, st));
  // large bf16 batches: deferred LayerNorm (see lxmert_forward_deferred); the fp32 copies of the embeddings are then
  // interleaved into the idle FFN buffer and only feed identity_state
  const bool defer = bf && option_get(OPT_DEFER_LN) != 0 && option_get(OPT_GEMM_PAIR) != 0 && option_get(OPT_GEMM_STAGED) != 0 &&
                     H == 768 && Mt >= 1024 && h->ones != nullptr;
  float* res_dst = defer ? static_cast<float*>(h->inter.p) : cur.res;
  if (bf) {  // the fp32 residual copies (written into `pre` by the two embedding kernels) take the same route
    RGQA_CUDA_OK(cudaMemcpy2DAsync(res_dst, static_cast<size_t>(S) * H * 4, h->pre.p, static_cast<size_t>(L) * H * 4,
                                   static_cast<size_t>(L) * H * 4, B, cudaMemcpyDeviceToDevice, st));
    RGQA_CUDA_OK(cudaMemcpy2DAsync(res_dst + static_cast<size_t>(L) * H, static_cast<size_t>(S) * H * 4,
                                   rows_at(h->pre.p, Ml, H, 4), static_cast<size_t>(V) * H * 4,
                                   static_cast<size_t>(V) * H * 4, B, cudaMemcpyDeviceToDevice, st));
  }
  RGQA_CUDA_OK(cudaMemsetAsync(jmask, 0, static_cast<size_t>(B) * S * 4, st));
  RGQA_CUDA_OK(cudaMemcpy2DAsync(jmask, static_cast<size_t>(S) * 4, mask, static_cast<size_t>(L) * 4,
                                 static_cast<size_t>(L) * 4, B, cudaMemcpyDeviceToDevice, st));
  {
    // ones for the image tokens: reuse the cast kernel's neighbour — a tiny fill via memset of 0x01 would
    // give 0x01010101 (non-zero = keep), which is all the attention kernel tests for
    RGQA_CUDA_OK(cudaMemset2DAsync(reinterpret_cast<char*>(jmask) + static_cast<size_t>(L) * 4, static_cast<size_t>(S) * 4,
                                   1, static_cast<size_t>(V) * 4, B, st));
  }
  if (defer) {
    DState dc{static_cast<char*>(h->hidA.p), static_cast<char*>(h->resA.p),
              reinterpret_cast<float*>(static_cast<char*>(h->resA.p) + static_cast<size_t>(Mt) * H * 2)};
    DState da{static_cast<char*>(h->hidB.p), static_cast<char*>(h->resB.p),
              reinterpret_cast<float*>(static_cast<char*>(h->resB.p) + static_cast<size_t>(Mt) * H * 2)};
    DefCtx d{h, H};
    d.f16 = h->state_f16;
    if (identity_state(st, res_dst, dc.hi, d.f16 ? nullptr : dc.lo, dc.stats, Mt, H)) return 1;
    const LNp* ln = nullptr;
    for (int i = 0; i < c.l_layers; ++i) {
      const LayerW& w = h->l_layers[i];
      if (d.att(st, w.att, dc, ln, da, 0, B, S, jmask) || d.ffn(st, w.ffn, da, &w.att.ln, dc, 0, static_cast<int>(Mt))) return 1;
      ln = &w.ffn.ln;
    }
    if (!out.any()) return 0;
    return head_block(h, st, dc.hi, B, S, out, ln != nullptr ? dc.stats : nullptr, d.f16 ? 1 : 0);
  }
  for (int i = 0; i < c.l_layers; ++i) {
    if (att_block(h, st, h->l_layers[i].att, cur, alt, 0, B, S, jmask)) return 1;
    if (ffn_block(h, st, h->l_layers[i].ffn, alt, cur, 0, static_cast<int>(Mt))) return 1;
  }
  if (!out.any()) return 0;
  return head_block(h, st, cur.act, B, S, out);
}

int rgqa_set_maha(rgqa_handle* h, const float* mu, const float* precision, int on_device) {
  RGQA_CHECK(h != nullptr && mu != nullptr && precision != nullptr, "rgqa_set_maha: null argument");
  const int H = h->cfg.hidden, A = h->cfg.num_answers;
  RGQA_CHECK(A > 0, "rgqa_set_maha: handle has no answer classes");
  const cudaMemcpyKind kind = on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
  if (h->maha_mu == nullptr) RGQA_CUDA_OK(cudaMalloc(&h->maha_mu, static_cast<size_t>(A) * H * 4));
  if (h->maha_P == nullptr) RGQA_CUDA_OK(cudaMalloc(&h->maha_P, static_cast<size_t>(H) * H * 4));
  if (h->maha_m == nullptr) RGQA_CUDA_OK(cudaMalloc(&h->maha_m, static_cast<size_t>(A) * 4));
  RGQA_CUDA_OK(cudaMemcpy(h->maha_mu, mu, static_cast<size_t>(A) * H * 4, kind));
  RGQA_CUDA_OK(cudaMemcpy(h->maha_P, precision, static_cast<size_t>(H) * H * 4, kind));
  // m_c = mu_c^T P mu_c :  T = mu · P  (gemm_f32 computes A·Wᵀ, so hand it Pᵀ via a transposed copy)
  float *PT = nullptr, *T = nullptr;
  RGQA_CUDA_OK(cudaMalloc(&PT, static_cast<size_t>(H) * H * 4));
  RGQA_CUDA_OK(cudaMalloc(&T, static_cast<size_t>(A) * H * 4));
  int rc = 0;
  {
    // P is symmetric up to rounding; use (P + Pᵀ)/2 so x^T P x is evaluated on its symmetric part exactly
    std::vector<float> hp(static_cast<size_t>(H) * H), hs(static_cast<size_t>(H) * H);
    cudaMemcpy(hp.data(), h->maha_P, hp.size() * 4, cudaMemcpyDeviceToHost);
    for (int i = 0; i < H; ++i)
      for (int j = 0; j < H; ++j) hs[static_cast<size_t>(i) * H + j] = 0.5f * (hp[static_cast<size_t>(i) * H + j] + hp[static_cast<size_t>(j) * H + i]);
    cudaMemcpy(h->maha_P, hs.data(), hs.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(PT, hs.data(), hs.size() * 4, cudaMemcpyHostToDevice);
  }
  rc = gemm_f32(nullptr, h->maha_mu, H, PT, H, nullptr, nullptr, 0, T, H, A, H, H, EPI_BIAS);
  if (!rc) rc = rowdot(nullptr, T, h->maha_mu, h->maha_m, A, H);
  if (!rc && h->cfg.precision == RGQA_PREC_BF16) {
    // tensor-core form: rows of [P; mu·P] as split-bf16 operands (hi | lo | hi), see maha_tc
    if (h->maha_Ws == nullptr && cudaMalloc(&h->maha_Ws, static_cast<size_t>(H + A) * 3 * H * 2) != cudaSuccess) {
      set_error("rgqa_set_maha: out of device memory");
      rc = 1;
    }
    if (!rc) rc = split_bf16x3(nullptr, h->maha_P, H, h->maha_Ws, H, H, 1);
    if (!rc) rc = split_bf16x3(nullptr, T, H, static_cast<char*>(h->maha_Ws) + static_cast<size_t>(H) * 3 * H * 2, A, H, 1);
  }
  cudaDeviceSynchronize();
  cudaFree(PT);
  cudaFree(T);
  if (rc) return 1;
  h->has_maha = true;
  return 0;
}

extern "C++" {
namespace {
int maha_core(cudaStream_t st, const float* x, const float* mu, const float* Psym, const float* m, int B, int A,
              int H, float* work, float* score, int32_t* label, float* full) {
  float* t = work;                                   // [B, H]
  float* q = t + static_cast<size_t>(B) * H;         // [B]
  float* G = q + ((static_cast<size_t>(B) + 3) & ~static_cast<size_t>(3));  // [B, A]
  if (gemm_f32(st, x, H, Psym, H, nullptr, nullptr, 0, t, H, B, H, H, EPI_BIAS)) return 1;  // t = x·Psymᵀ = x·Psym
  if (rowdot(st, t, x, q, B, H)) return 1;
  if (gemm_f32(st, t, H, mu, H, nullptr, nullptr, 0, G, A, B, A, H, EPI_BIAS)) return 1;    // G = t·muᵀ
  return maha_finish(st, q, G, A, m, B, A, score, label, full);
}
}  // namespace
}  // extern "C++"

int rgqa_score(rgqa_handle* h, void* stream, int mode, const float* logits, const float* pooled, int B,
               float* score_out, int32_t* label_out, float* maha_full_out) {
  RGQA_CHECK(h != nullptr && score_out != nullptr && label_out != nullptr, "rgqa_score: null argument");
  RGQA_CHECK(B > 0, "rgqa_score: empty batch");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int A = h->cfg.num_answers, H = h->cfg.hidden;
  RGQA_CHECK(A > 0, "rgqa_score: handle has no answer classes");
  if (mode == RGQA_SCORE_MAHA) {
    RGQA_CHECK(h->has_maha, "rgqa_score: Mahalanobis statistics were not set (rgqa_set_maha)");
    RGQA_CHECK(pooled != nullptr, "rgqa_score: MAHA needs the pooled feature");
    if (h->maha_Ws != nullptr && option_get(OPT_MAHA_TC) != 0 && maha_full_out == nullptr &&
        reinterpret_cast<uintptr_t>(pooled) % 16 == 0)
      return maha_tc(h, st, pooled, B, score_out, label_out);
    if (h->mahabuf.ensure((static_cast<size_t>(B) * (H + A + 1) + 16) * 4)) return 1;
    return maha_core(st, pooled, h->maha_mu, h->maha_P, h->maha_m, B, A, H, static_cast<float*>(h->mahabuf.p), score_out,
                     label_out, maha_full_out);
  }
  RGQA_CHECK(logits != nullptr, "rgqa_score: this scorer needs logits");
  return score_logits(st, mode, logits, A, B, A, score_out, label_out);
}

// ---- measurement hooks ----
int rgqa_set_option(const char* name, int value) {
  RGQA_CHECK(name != nullptr, "rgqa_set_option: null name");
  if (strcmp(name, "gemm_pair") == 0) {
    option_set(OPT_GEMM_PAIR, value);
    return 0;
  }
  if (strcmp(name, "gemm_staged") == 0) {
    option_set(OPT_GEMM_STAGED, value);
    return 0;
  }
  if (strcmp(name, "graphs") == 0) {
    option_set(OPT_GRAPHS, value);
    return 0;
  }
  if (strcmp(name, "defer_ln") == 0) {
    option_set(OPT_DEFER_LN, value);
    return 0;
  }
  if (strcmp(name, "attn_tma") == 0) {
    option_set(OPT_ATTN_TMA, value);
    return 0;
  }
  if (strcmp(name, "qkv_planes") == 0) {
    option_set(OPT_QKV_PLANES, value);
    return 0;
  }
  if (strcmp(name, "two_streams") == 0) {
    option_set(OPT_TWO_STREAMS, value);
    return 0;
  }
  if (strcmp(name, "gemm_ln") == 0) {
    option_set(OPT_GEMM_LN, value);
    return 0;
  }
  if (strcmp(name, "pdl") == 0) {
    option_set(OPT_PDL, value);
    return 0;
  }
  if (strcmp(name, "maha_tc") == 0) {
    option_set(OPT_MAHA_TC, value);
    return 0;
  }
  if (strcmp(name, "visn_fused") == 0) {
    option_set(OPT_VISN_FUSED, value);
    return 0;
  }
  if (strcmp(name, "attn_fused") == 0) {
    option_set(OPT_ATTN_FUSED, value);
    return 0;
  }
  if (strcmp(name, "sm_split") == 0) {
    option_set(OPT_SM_SPLIT, value);
    return 0;
  }
  if (strcmp(name, "pair_deep") == 0) {
    option_set(OPT_PAIR_DEEP, value);
    return 0;
  }
  if (strcmp(name, "attn_straddle") == 0) {
    option_set(OPT_ATTN_STRADDLE, value);
    return 0;
  }
  if (strcmp(name, "gelu_ew16") == 0) {
    option_set(OPT_GELU_EW16, value);
    return 0;
  }
  if (strcmp(name, "state_f16") == 0) {
    option_set(OPT_STATE_F16, value);
    return 0;
  }
  if (strcmp(name, "gemm_debug") == 0) {  // measurement only: results are wrong while non-zero
    option_set(OPT_GEMM_DEBUG, value);
    return 0;
  }
  set_error("rgqa_set_option: unknown option '%s'", name);
  return 1;
}
long long rgqa_launch_count(void) { return launch_count(); }
void rgqa_prof_enable(int on) { prof_enable(on != 0); }
int rgqa_prof_collect(double* ms_sum, double* flops_sum, long long* count) {
  RGQA_CHECK(ms_sum && flops_sum && count, "rgqa_prof_collect: null argument");
  return prof_collect(ms_sum, flops_sum, count, PROF_NKINDS);
}

// ---- per-kernel entry points ----
int rgqa_gemm_bf16(void* stream, const void* A, int64_t lda, const void* W, int64_t ldw, const float* bias,
                   const void* resid, int64_t ldr, void* out, int64_t ldo, int out_dtype, int M, int N, int K,
                   int epilogue) {
  return gemm_bf16(static_cast<cudaStream_t>(stream), A, lda, W, ldw, bias, resid, ldr, out, ldo,
                   out_dtype == RGQA_DT_F32 ? 1 : 0, M, N, K, epilogue);
}
int rgqa_gemm_bf16_reduce_logits(void* stream, const void* A, int64_t lda, const void* W, int64_t ldw, const float* bias,
                                 float* logits_out, int M, int N, int K, int flags, void* work, int64_t work_bytes,
                                 float* scores, int32_t* labels) {
  RGQA_CHECK(work != nullptr && scores != nullptr && labels != nullptr, "rgqa_gemm_bf16_reduce_logits: null argument");
  const size_t part = row_reduce_part_bytes(M, N), cnt = static_cast<size_t>(row_reduce_counters(M)) * sizeof(int);
  RGQA_CHECK(static_cast<size_t>(work_bytes) >= part + cnt, "rgqa_gemm_bf16_reduce_logits: work needs %zu bytes", part + cnt);
  RGQA_CUDA_OK(cudaMemsetAsync(static_cast<char*>(work) + part, 0, cnt, static_cast<cudaStream_t>(stream)));
  RowReduce r;
  r.kind = RED_LOGITS;
  r.flags = flags;
  r.part = static_cast<float*>(work);
  r.counter = reinterpret_cast<int*>(static_cast<char*>(work) + part);
  r.scores = scores;
  r.labels = labels;
  return gemm_bf16_reduce(static_cast<cudaStream_t>(stream), A, lda, W, ldw, bias, logits_out, N, M, N, K, r);
}
int64_t rgqa_reduce_work_bytes(int M, int N) {
  return static_cast<int64_t>(row_reduce_part_bytes(M, N) + static_cast<size_t>(row_reduce_counters(M)) * sizeof(int));
}
int rgqa_qkv_attention_fused(void* stream, const void* A, int64_t lda, const void* W_hm, const float* bias_hm, const float* ln_stats,
                             float ln_eps, const int32_t* key_mask, void* ctx, int64_t ldo, int B, int S, int heads) {
  return qkv_attention_fused(static_cast<cudaStream_t>(stream), A, lda, W_hm, bias_hm, ln_stats, ln_eps, key_mask, ctx, ldo, B, S, heads);
}
int rgqa_visn_embed(void* stream, const void* feat, int feat_dtype, const void* W1, const float* b1, const float* g1,
                    const float* be1, const float* boxes, int box_dim, const float* W2, const float* b2, const float* g2,
                    const float* be2, float eps, void* out_hi, void* out_lo, float* out_stats, int M, int K) {
  RGQA_CHECK(W2 != nullptr && b2 != nullptr && box_dim >= 1 && box_dim <= 8, "rgqa_visn_embed: bad box projection");
  std::vector<float> w2p, bst;
  if (make_box_stats(W2, b2, 768, box_dim, &w2p, &bst)) return 1;
  float *dw = nullptr, *ds = nullptr;
  RGQA_CUDA_OK(cudaMalloc(reinterpret_cast<void**>(&dw), w2p.size() * 4));
  RGQA_CUDA_OK(cudaMalloc(reinterpret_cast<void**>(&ds), bst.size() * 4));
  RGQA_CUDA_OK(cudaMemcpy(dw, w2p.data(), w2p.size() * 4, cudaMemcpyHostToDevice));
  RGQA_CUDA_OK(cudaMemcpy(ds, bst.data(), bst.size() * 4, cudaMemcpyHostToDevice));
  const int rc = visn_embed(static_cast<cudaStream_t>(stream), feat, feat_dtype == RGQA_DT_F32 ? 1 : 0, K, W1, b1, g1, be1, boxes, box_dim,
                            dw, b2, g2, be2, ds, eps, out_hi, out_lo, out_stats, M, K);
  cudaStreamSynchronize(static_cast<cudaStream_t>(stream));  // test entry point: the scratch tables are freed right away
  cudaFree(dw);
  cudaFree(ds);
  return rc;
}
int rgqa_gemm_bf16_ln(void* stream, const void* A, int64_t lda, const void* W, int64_t ldw, const float* bias,
                      const float* resid, int64_t ldr, const float* gamma, const float* beta, float eps, void* out_act,
                      float* out_res, int M, int N, int K) {
  return gemm_bf16_ln(static_cast<cudaStream_t>(stream), A, lda, W, ldw, bias, resid, ldr, gamma, beta, eps, out_act,
                      out_res, M, N, K);
}
int rgqa_fold_ln_weight(void* stream, const float* W, const float* bias, const float* gamma, const float* beta, int N,
                        int K, void* W_out_bf16, float* c_out, float* bias_out) {
  return fold_ln_weight(static_cast<cudaStream_t>(stream), W, bias, gamma, beta, N, K, W_out_bf16, c_out, bias_out);
}
int rgqa_gemm_bf16_folded(void* stream, const void* A, int64_t lda, const void* W, int64_t ldw, const float* bias,
                          const void* resid, int64_t ldr, void* out, int64_t ldo, int out_dtype, int M, int N, int K,
                          int epilogue, const float* ln_stats, const float* ln_c, float eps) {
  return gemm_bf16_folded(static_cast<cudaStream_t>(stream), A, lda, W, ldw, bias, resid, ldr, out, ldo,
                          out_dtype == RGQA_DT_F32 ? 1 : 0, M, N, K, epilogue, ln_stats, ln_c, eps, 12);
}
int rgqa_gemm_bf16_resid_hl(void* stream, const void* A, int64_t lda, const void* W, int64_t ldw, const float* bias,
                            const void* in_hi, const void* in_lo, const float* in_stats, const float* in_gamma,
                            const float* in_beta, float eps, void* out_hi, void* out_lo, float* out_stats, int M, int N,
                            int K) {
  return gemm_bf16_resid_hl(static_cast<cudaStream_t>(stream), A, lda, W, ldw, bias, in_hi, in_lo, in_stats, in_gamma,
                            in_beta, eps, out_hi, out_lo, out_stats, M, N, K);
}
int rgqa_fold_ln_weight_f16(void* stream, const float* W, const float* bias, const float* gamma, const float* beta, int N,
                            int K, void* W_out_f16, float* c_out, float* bias_out) {
  return fold_ln_weight(static_cast<cudaStream_t>(stream), W, bias, gamma, beta, N, K, W_out_f16, c_out, bias_out, 1);
}
int rgqa_gemm_bf16_folded_a16(void* stream, const void* A, int64_t lda, const void* W, int64_t ldw, const float* bias, void* out,
                              int64_t ldo, int out_dtype, int M, int N, int K, int epilogue, const float* ln_stats,
                              const float* ln_c, float eps) {
  return gemm_bf16_folded(static_cast<cudaStream_t>(stream), A, lda, W, ldw, bias, nullptr, 0, out, ldo,
                          out_dtype == RGQA_DT_F32 ? 1 : 0, M, N, K, epilogue, ln_stats, ln_c, eps, 12, 1);
}
int rgqa_qkv_attention_fused_a16(void* stream, const void* A, int64_t lda, const void* W_hm, const float* bias_hm,
                                 const float* ln_stats, float ln_eps, const int32_t* key_mask, void* ctx, int64_t ldo, int B, int S,
                                 int heads) {
  return qkv_attention_fused(static_cast<cudaStream_t>(stream), A, lda, W_hm, bias_hm, ln_stats, ln_eps, key_mask, ctx, ldo, B, S,
                             heads, 1);
}
int rgqa_gemm_bf16_resid_h16(void* stream, const void* A, int64_t lda, const void* W, int64_t ldw, const float* bias,
                             const void* in_h, const float* in_stats, const float* in_gamma, const float* in_beta, float eps,
                             void* out_h, float* out_stats, int M, int N, int K) {
  return gemm_bf16_resid_h16(static_cast<cudaStream_t>(stream), A, lda, W, ldw, bias, in_h, in_stats, in_gamma, in_beta, eps,
                             out_h, out_stats, M, N, K);
}
int rgqa_gemm_f32(void* stream, const float* A, int64_t lda, const float* W, int64_t ldw, const float* bias,
                  const float* resid, int64_t ldr, float* out, int64_t ldo, int M, int N, int K, int epilogue) {
  return gemm_f32(static_cast<cudaStream_t>(stream), A, lda, W, ldw, bias, resid, ldr, out, ldo, M, N, K, epilogue);
}
int rgqa_attention(void* stream, int dtype, const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v,
                   int64_t ldv, const int32_t* key_mask, void* out, int64_t ldo, int B, int Lq, int Lk, int heads) {
  return attention(static_cast<cudaStream_t>(stream), dtype, q, ldq, k, ldk, v, ldv, key_mask, out, ldo, B, Lq, Lk, heads);
}
int rgqa_embed_ln(void* stream, int out_dtype, const int32_t* ids, const int32_t* seg, const float* word,
                  const float* pos, const float* type, const float* gamma, const float* beta, void* out, int B,
                  int L, int H, int vocab, int max_pos, int n_types, float eps) {
  return embed_ln(static_cast<cudaStream_t>(stream), out_dtype, ids, seg, word, pos, type, gamma, beta, out, nullptr, B, L,
                  H, vocab, max_pos, n_types, eps);
}
int rgqa_layernorm(void* stream, int out_dtype, const float* x, int64_t ldx, const float* gamma, const float* beta,
                   void* y, int M, int H, float eps) {
  return layernorm(static_cast<cudaStream_t>(stream), out_dtype, x, ldx, gamma, beta, y, nullptr, M, H, eps);
}
int rgqa_visn_combine(void* stream, int out_dtype, const float* x_pre, const float* boxes, int box_dim,
                      const float* Wb, const float* bb, const float* g1, const float* b1, const float* g2,
                      const float* b2, void* out, int M, int H, float eps) {
  return visn_combine(static_cast<cudaStream_t>(stream), out_dtype, x_pre, boxes, box_dim, Wb, bb, g1, b1, g2, b2, out,
                      nullptr, M, H, eps);
}
int rgqa_cast_f32_to_bf16(void* stream, const float* in, void* out, int64_t n) {
  return cast_f32_to_bf16(static_cast<cudaStream_t>(stream), in, out, n);
}
int rgqa_layernorm_bwd(void* stream, const float* dy, const float* x, const float* gamma, const float* add, float* dx, int M,
                       int H, float eps, float scale) {
  return layernorm_bwd(static_cast<cudaStream_t>(stream), dy, x, gamma, add, dx, M, H, eps, scale);
}
int rgqa_layernorm_bwd_params(void* stream, const float* dy, const float* x, const float* gamma, const float* add, float* dx,
                              float* dgamma, float* dbeta, int M, int H, float eps, float scale) {
  return layernorm_bwd_params(static_cast<cudaStream_t>(stream), dy, x, gamma, add, dx, dgamma, dbeta, M, H, eps, scale);
}
int rgqa_gelu_fwd(void* stream, const float* x, float* y, int64_t n) { return gelu_fwd(static_cast<cudaStream_t>(stream), x, y, n); }
int rgqa_gelu_bwd(void* stream, const float* dy, const float* x, float* dx, int64_t n) {
  return gelu_bwd(static_cast<cudaStream_t>(stream), dy, x, dx, n);
}
int rgqa_gelu_fwd_bf16(void* stream, const float* x, void* y_bf16, int64_t n) {
  return gelu_fwd_bf16(static_cast<cudaStream_t>(stream), x, y_bf16, n);
}
int rgqa_layernorm_dual(void* stream, const float* x, int64_t ldx, const float* gamma, const float* beta, void* y_bf16, float* y_f32,
                        int M, int H, float eps) {
  RGQA_CHECK(y_bf16 != nullptr && y_f32 != nullptr, "rgqa_layernorm_dual: null output");
  return layernorm(static_cast<cudaStream_t>(stream), DT_BF16, x, ldx, gamma, beta, y_bf16, y_f32, M, H, eps);
}
int rgqa_tanh_bwd(void* stream, const float* dy, const float* y, float* dx, int64_t n) {
  return tanh_bwd(static_cast<cudaStream_t>(stream), dy, y, dx, n);
}
int rgqa_attention_bwd(void* stream, const float* q, int64_t ldq, const float* k, int64_t ldk, const float* v, int64_t ldv,
                       const int32_t* key_mask, const float* dctx, int64_t ldo, float* dq, int64_t lddq, float* dk, int64_t lddk,
                       float* dv, int64_t lddv, int B, int Lq, int Lk, int heads, float dropout_p, uint64_t seed) {
  return attention_bwd(static_cast<cudaStream_t>(stream), q, ldq, k, ldk, v, ldv, key_mask, dctx, ldo, dq, lddq, dk, lddk, dv, lddv,
                       B, Lq, Lk, heads, dropout_p, seed);
}
int rgqa_attention_bwd_bf16(void* stream, const float* q, int64_t ldq, const float* k, int64_t ldk, const float* v, int64_t ldv,
                            const int32_t* key_mask, const float* dctx, int64_t ldo, float* dq, int64_t lddq, float* dk, int64_t lddk,
                            float* dv, int64_t lddv, int B, int Lq, int Lk, int heads) {
  return attention_bwd_bf16(static_cast<cudaStream_t>(stream), q, ldq, k, ldk, v, ldv, key_mask, dctx, ldo, dq, lddq, dk, lddk, dv, lddv,
                            B, Lq, Lk, heads);
}
int rgqa_attention_dropout(void* stream, const float* q, int64_t ldq, const float* k, int64_t ldk, const float* v, int64_t ldv,
                           const int32_t* key_mask, float* out, int64_t ldo, int B, int Lq, int Lk, int heads, float dropout_p,
                           uint64_t seed) {
  return attention(static_cast<cudaStream_t>(stream), DT_F32, q, ldq, k, ldk, v, ldv, key_mask, out, ldo, B, Lq, Lk, heads, dropout_p,
                   seed);
}
int rgqa_transpose_cast_bf16(void* stream, const float* in, int64_t ld, void* out, int R, int C) {
  RGQA_CHECK(in && out, "rgqa_transpose_cast_bf16: null argument");
  return transpose_cast_bf16(static_cast<cudaStream_t>(stream), in, ld, out, R, C);
}
int rgqa_transpose_cast_bf16_colsum(void* stream, const float* in, int64_t ld, void* out, int R, int C, float* colsum) {
  RGQA_CHECK(in && out && colsum, "rgqa_transpose_cast_bf16_colsum: null argument");
  return transpose_cast_bf16(static_cast<cudaStream_t>(stream), in, ld, out, R, C, colsum);
}
int rgqa_sign_step(void* stream, const float* x, const float* grad, float* y, int64_t n, float noise) {
  RGQA_CHECK(x && grad && y, "rgqa_sign_step: null argument");
  return sign_step(static_cast<cudaStream_t>(stream), x, grad, y, n, noise);
}
int rgqa_dropout_add(void* stream, const float* x, const float* resid, float* y, int64_t n, float p, uint64_t seed) {
  return dropout_add(static_cast<cudaStream_t>(stream), x, resid, y, n, p, seed);
}
int rgqa_gemm_f32_tn(void* stream, const float* A, int64_t lda, const float* B, int64_t ldb, float* out, int64_t ldo, int M, int N,
                     int K) {
  return gemm_f32_tn(static_cast<cudaStream_t>(stream), A, lda, B, ldb, out, ldo, M, N, K);
}
int rgqa_gemm_bf16_accum(void* stream, const void* A, int64_t lda, const void* W, int64_t ldw, float* out, int64_t ldo, int M, int N,
                         int K) {
  return gemm_bf16_accum(static_cast<cudaStream_t>(stream), A, lda, W, ldw, out, ldo, M, N, K);
}
int rgqa_gemm_bf16_accum_tn(void* stream, const void* At, int64_t ldat, const void* Bt, int64_t ldbt, float* out, int64_t ldo, int M,
                            int N, int K) {
  return gemm_bf16_accum_tn(static_cast<cudaStream_t>(stream), At, ldat, Bt, ldbt, out, ldo, M, N, K);
}
int rgqa_cast_bf16_colsum(void* stream, const float* in, int64_t ld, void* out, int64_t ldo, int R, int C, float* colsum) {
  return cast_bf16_colsum(static_cast<cudaStream_t>(stream), in, ld, out, ldo, R, C, colsum);
}
int rgqa_colsum(void* stream, const float* X, int64_t ldx, float* out, int M, int N) {
  return colsum(static_cast<cudaStream_t>(stream), X, ldx, out, M, N);
}
int rgqa_ln_param_grad(void* stream, const float* dy, const float* x, float* dgamma, float* dbeta, int M, int H, float eps,
                       float scale) {
  return ln_param_grad(static_cast<cudaStream_t>(stream), dy, x, dgamma, dbeta, M, H, eps, scale);
}
int rgqa_embedding_scatter(void* stream, const float* g, const int32_t* idx, float* table, int rows, int H, int table_rows,
                           int padding_idx) {
  return embedding_scatter(static_cast<cudaStream_t>(stream), g, idx, table, rows, H, table_rows, padding_idx);
}
int rgqa_bce_loss_grad(void* stream, const float* logit, const float* target, int B, int A, float* dlogit, double* loss_acc) {
  return bce_loss_grad(static_cast<cudaStream_t>(stream), logit, target, B, A, dlogit, loss_acc);
}
int rgqa_grad_sumsq(void* stream, const float* g, int64_t n, double* acc) {
  return grad_sumsq(static_cast<cudaStream_t>(stream), g, n, acc);
}
int rgqa_bertadam_step(void* stream, float* p, const float* g, float* m, float* v, int64_t n, float lr_scheduled, double b1,
                       double b2, float eps, float weight_decay, const double* grad_sumsq_acc, float max_norm) {
  return bertadam_step(static_cast<cudaStream_t>(stream), p, g, m, v, n, lr_scheduled, b1, b2, eps, weight_decay,
                       grad_sumsq_acc, max_norm);
}
int rgqa_estimator_accumulate(void* stream, const float* x, const float* target, int64_t ld_target, int B, int C, int D,
                              double* G, double* class_sum, double* class_count, float* w_scratch) {
  return estimator_accumulate(static_cast<cudaStream_t>(stream), x, target, ld_target, B, C, D, G, class_sum, class_count,
                              w_scratch);
}
int rgqa_score_ensemble(void* stream, const float* const* logits, int K, int method, int64_t ld, int B, int A, float* score,
                        int32_t* label) {
  RGQA_CHECK(method == 0 || method == 1, "rgqa_score_ensemble: method must be 0 (mean) or 1 (multiply)");
  return score_ensemble(static_cast<cudaStream_t>(stream), logits, K, method, ld, B, A, score, label);
}
int rgqa_score_logits(void* stream, int mode, const float* logits, int64_t ld, int B, int A, float* score,
                      int32_t* label) {
  return score_logits(static_cast<cudaStream_t>(stream), mode, logits, ld, B, A, score, label);
}
int rgqa_mahalanobis(void* stream, const float* x, const float* mu, const float* P, int B, int A, int H, float* work,
                     float* score, int32_t* label, float* full_out) {
  // stand-alone form: m_c computed on the fly into the tail of `work`; P is used as given (assumed symmetric)
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  RGQA_CHECK(x && mu && P && work && score && label, "rgqa_mahalanobis: null argument");
  RGQA_CHECK(B > 0 && A > 0 && H > 0, "rgqa_mahalanobis: empty input");
  // layout of work: [B*H + pad(B) + B*A] core | [A*H] T | [A] m   -> caller must size for it:
  // B*(H+A+1)+4 + A*(H+1) floats
  float* core = work;
  float* T = core + static_cast<size_t>(B) * H + ((static_cast<size_t>(B) + 3) & ~static_cast<size_t>(3)) + static_cast<size_t>(B) * A;
  float* m = T + static_cast<size_t>(A) * H;
  if (gemm_f32(st, mu, H, P, H, nullptr, nullptr, 0, T, H, A, H, H, EPI_BIAS)) return 1;  // T = mu·Pᵀ
  if (rowdot(st, T, mu, m, A, H)) return 1;
  return maha_core(st, x, mu, P, m, B, A, H, core, score, label, full_out);
}

}  // extern "C"
#pragma GCC visibility pop
