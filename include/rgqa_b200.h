/*
 * rgqa_b200 — C ABI of the B200-native RGQA evaluation hot path.
 *
 * This is the drop-in boundary: plain pointers and sizes, no torch types.  Every function returns 0
 * on success and non-zero on failure; the message of the last failure on the calling thread is
 * rgqa_last_error().  Nothing throws across this boundary.  All device pointers are borrowed for the
 * duration of the call; the handle owns only its packed weights and its workspace.  All work is
 * enqueued on the caller's CUDA stream (`stream` is a cudaStream_t / CUstream passed as void*).
 *
 * Reference interface each entry point replaces (paths relative to chihhuiho/RGQA):
 *   rgqa_create / rgqa_load_weight / rgqa_finalize_weights
 *        GQAModel.__init__ + load_state_dict        src/tasks/gqa_model.py:14-28, src/tasks/gqa_conf.py:375-381,
 *        LXRTEncoder.load                            src/lxrt/entry.py:126-152
 *   rgqa_lxmert_forward
 *        LXRTFeatureExtraction.forward (mode 'x')    src/lxrt/modeling.py:1020-1030
 *        + GQAModel.logit_fc                         src/tasks/gqa_model.py:22-43
 *   rgqa_uniter_forward
 *        UniterFeatureExtraction.forward             src/uniter/modeling.py:651-656
 *        + GQAUNITER.logit_fc                        src/uniter/uniter.py:23-44
 *   rgqa_score
 *        MSP scorer                                  src/tasks/gqa_conf.py:326
 *        energy scorers                              src/tasks/gqa_energy.py:131,185,204-206
 *        Mahalanobis scorer (noise = 0)              src/tasks/gqa_maha.py:228-236,270
 *   rgqa_set_maha
 *        sample_estimator() output {mean, precision} src/tasks/gqa_maha.py:120-201
 *   rgqa_gemm_* / rgqa_attention / rgqa_embed_ln / rgqa_layernorm / rgqa_visn_combine
 *        the individual nn.Module forwards of src/lxrt/modeling.py (BertAttention :320-347,
 *        BertAttOutput :357-361, BertIntermediate :398-401, BertOutput :411-415, BertEmbeddings :278-292,
 *        VisualFeatEncoder :507-517) — exported so each kernel is testable on its own.
 */
#ifndef RGQA_B200_H_
#define RGQA_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RGQA_ABI_VERSION 3

/* architectures */
#define RGQA_ARCH_LXMERT 0
#define RGQA_ARCH_UNITER 1
/* arithmetic of the path */
#define RGQA_PREC_FP32 0 /* fp32 weights + activations, CUDA-core GEMMs (tolerance 1e-4 on logits)       */
#define RGQA_PREC_BF16 1 /* bf16 weights + activations, tcgen05 GEMMs with fp32 accumulate / LN / softmax */
/* element types of borrowed buffers */
#define RGQA_DT_F32 0
#define RGQA_DT_BF16 1
#define RGQA_DT_F16 2 /* the single-word deferred-LayerNorm state: operands of the *_a16 GEMMs, output of rgqa_gemm_bf16_resid_h16 */
/* GEMM epilogues */
#define RGQA_EPI_BIAS 0
#define RGQA_EPI_GELU 1
#define RGQA_EPI_TANH 2
#define RGQA_EPI_RESID 3
/* scorers */
#define RGQA_SCORE_MSP 0
#define RGQA_SCORE_ENERGY_TOP2 1
#define RGQA_SCORE_ENERGY_SUM 2
#define RGQA_SCORE_LSE 3
#define RGQA_SCORE_MAHA 4

typedef struct rgqa_handle rgqa_handle;

typedef struct rgqa_config {
  int32_t arch;            /* RGQA_ARCH_*                                                        */
  int32_t l_layers;        /* language layers (LXMERT: 9; UNITER: 12 single-stream layers)       */
  int32_t x_layers;        /* cross-modality layers (LXMERT: 5; UNITER: 0)                       */
  int32_t r_layers;        /* object-relational layers (LXMERT: 5; UNITER: 0)                    */
  int32_t hidden;          /* 768                                                                */
  int32_t heads;           /* 12                                                                 */
  int32_t intermediate;    /* 3072                                                               */
  int32_t vocab_size;      /* 30522 (bert-base-uncased) / 28996 (bert-base-cased, UNITER)        */
  int32_t max_position;    /* 512                                                                */
  int32_t type_vocab;      /* 2                                                                  */
  int32_t visual_feat_dim; /* 2048                                                               */
  int32_t visual_pos_dim;  /* 4 (LXMERT boxes) / 7 (UNITER)                                      */
  int32_t num_rois;        /* 36                                                                 */
  int32_t num_answers;     /* 1842; 0 = encoder only (no logit_fc)                               */
  int32_t precision;       /* RGQA_PREC_*                                                        */
  float ln_eps;            /* 1e-12                                                              */
} rgqa_config;

/* ---- lifecycle ------------------------------------------------------------------------------ */
int rgqa_abi_version(void);
const char* rgqa_last_error(void);
int rgqa_create(const rgqa_config* cfg, rgqa_handle** out);
void rgqa_destroy(rgqa_handle* h);

/* Number / names of the reference state-dict tensors this handle expects (for the host shim). */
int rgqa_num_weights(const rgqa_handle* h);
const char* rgqa_weight_name(const rgqa_handle* h, int index);
/* Copy one fp32 tensor, named by its reference state-dict key (a leading "module.",
 * "lxrt_encoder.model." or "encoder.model." prefix is accepted and ignored), into the handle.
 * `data` is a host pointer (on_device = 0) or a device pointer (on_device = 1). */
int rgqa_load_weight(rgqa_handle* h, const char* ref_key, const float* data, int on_device, const int64_t* shape,
                     int ndim);
/* Verify every expected tensor arrived, then pack: fused Q|K|V, bf16 casts (PREC_BF16). */
int rgqa_finalize_weights(rgqa_handle* h);

/* ---- the hot path --------------------------------------------------------------------------- */
/* feat: [B, num_rois, visual_feat_dim] (feat_dtype F32 or BF16), pos: [B, num_rois, 4] f32,
 * ids / seg / mask: [B, L] int32 (mask: 1 = real token).  pooled_out: [B, hidden] f32 or NULL;
 * logits_out: [B, num_answers] f32 or NULL (must be NULL when num_answers == 0). */
int rgqa_lxmert_forward(rgqa_handle* h, void* stream, const void* feat, int feat_dtype, const float* pos,
                        const int32_t* ids, const int32_t* seg, const int32_t* mask, int B, int L,
                        float* pooled_out, float* logits_out);
/* UNITER: pos7 is [B, num_rois, 7] f32, the vector GQATorchDataset._uniterBoxes builds
 * (src/tasks/gqa_data.py:240-250: [y1,x1,y2,x2,w,h,w*h]); passed through untouched. */
int rgqa_uniter_forward(rgqa_handle* h, void* stream, const void* feat, int feat_dtype, const float* pos7,
                        const int32_t* ids, const int32_t* seg, const int32_t* mask, int B, int L,
                        float* pooled_out, float* logits_out);

/* Forward + rejection scorers in ONE call (either architecture; BASELINE north_star item 3).  score_mask has bit
 * (1 << RGQA_SCORE_*) set per scorer wanted.  In bf16 handles the answer-head GEMM's epilogue reduces its own logit rows
 * (arg-max of sigmoid, top-2, softplus sum, logsumexp: src/tasks/gqa_conf.py:326, gqa_energy.py:131,185,204-206) and the
 * Mahalanobis scorer (gqa_maha.py:228-236,270) is one split-bf16 tensor-core GEMM with the arg-max in its epilogue, so
 * neither [B, num_answers] matrix reaches HBM unless logits_out is given.
 * scores_out: [5][B] f32, row = RGQA_SCORE_* code (rows of scorers not requested are untouched);
 * labels_out: [2][B] int32, row 0 = arg-max of sigmoid(logit) (label of MSP / ENERGY_* / LSE), row 1 = Mahalanobis class.
 * pooled_out / logits_out as in rgqa_lxmert_forward (may be NULL). */
int rgqa_forward_scored(rgqa_handle* h, void* stream, const void* feat, int feat_dtype, const float* pos,
                        const int32_t* ids, const int32_t* seg, const int32_t* mask, int B, int L, unsigned score_mask,
                        float* pooled_out, float* logits_out, float* scores_out, int32_t* labels_out);

/* Mahalanobis statistics: class means mu [num_answers, hidden] f32 and precision P [hidden, hidden] f32. */
int rgqa_set_maha(rgqa_handle* h, const float* mu, const float* precision, int on_device);
/* score_out: [B] f32, label_out: [B] int32.  MSP / ENERGY_* / LSE read `logits` [B, num_answers];
 * MAHA reads `pooled` [B, hidden] (and optionally writes the full [B, num_answers] score matrix). */
int rgqa_score(rgqa_handle* h, void* stream, int mode, const float* logits, const float* pooled, int B,
               float* score_out, int32_t* label_out, float* maha_full_out);

/* ---- measurement hooks (bench.py) ----------------------------------------------------------- */
/* Kernel-selection switches for A/B measurements: "gemm_pair" (1 = CTA-pair tcgen05 kernel when
 * M >= 1024), "gemm_staged" (1 = swizzled-smem + TMA-store epilogue in the pair kernel), "gemm_ln" (1 = fused
 * GEMM + residual + LayerNorm kernel), "two_streams" (1 = visual branch on a second stream), "graphs" (1 = CUDA-graph replay of the forward for
 * B <= 128), "defer_ln" (1 = LXMERT bf16 batches of >= 1024 token rows run without LayerNorm kernels: hidden states stay
 * pre-LayerNorm as bf16 hi/lo pairs + row statistics, see rgqa_gemm_bf16_resid_hl), "attn_tma" (1 = TMA-fed persistent attention kernel for large batches), "qkv_planes" (1 = Q|K|V projections written head-major for that kernel), "maha_tc" (1 = Mahalanobis on the tensor
 * cores in bf16 handles), "visn_fused" (1 = one-kernel visual feature encoder), "attn_fused" (1 = self-attention inside the
 * Q|K|V projection's epilogue), "pdl", "gemm_debug". */
int rgqa_set_option(const char* name, int value);
/* Kernels enqueued by this library since it was loaded. */
long long rgqa_launch_count(void);
/* Deferred LayerNorm bookkeeping of a finalized bf16 handle: returns max |c[n]| over every folded consumer weight (c = row sums
 * of the centred bf16 weight, the term the consumer epilogue drops) and, through exact_consumers, how many consumers kept the
 * exact epilogue because their residue was too large to drop (max|c| * 8 > 1e-3). */
float rgqa_fold_residue(const rgqa_handle* h, int* exact_consumers);
/* CUDA graphs this handle has instantiated: large batches are captured per (shape, feature / box pointers), so a caller
 * that hands a fresh buffer every call keeps capturing (and pays for it); a sweep over two staging slots captures twice. */
long long rgqa_graph_captures(const rgqa_handle* h);
/* When on, every GEMM / attention launch is bracketed by CUDA events on its own stream. */
void rgqa_prof_enable(int on);
/* Sum of event-timed durations (ms), algorithmic FLOPs and launch counts per kernel kind since the last
 * collect; arrays of RGQA_PROF_NKINDS entries: 0 tcgen05 GEMM, 1 fp32 GEMM, 2 attention, 3 other. */
#define RGQA_PROF_NKINDS 4
int rgqa_prof_collect(double* ms_sum, double* flops_sum, long long* count);

/* ---- per-kernel entry points (unit tests, integration of single ops) ------------------------- */
int rgqa_gemm_bf16(void* stream, const void* A, int64_t lda, const void* W, int64_t ldw, const float* bias,
                   const void* resid, int64_t ldr, void* out, int64_t ldo, int out_dtype, int M, int N, int K,
                   int epilogue);
/* BertAttention.forward for self-attention (src/lxrt/modeling.py:320-347, :376-386) as ONE kernel: the Q|K|V projection on
 * tcgen05 with the attention of each (sample, head) in its epilogue.  A: bf16 [B*S, 768] rows (pre-LayerNorm when ln_stats
 * [B*S][12] is given, with W_hm / bias_hm folded accordingly); W_hm bf16 [2304, 768] and bias_hm f32 [2304] packed head-major
 * ([head][q|k|v][64] rows); key_mask int32 [B, S] or NULL; ctx bf16 [B*S, 768] (row stride ldo).  12 heads of 64, S <= 64. */
int rgqa_qkv_attention_fused(void* stream, const void* A, int64_t lda, const void* W_hm, const float* bias_hm, const float* ln_stats,
                             float ln_eps, const int32_t* key_mask, void* ctx, int64_t ldo, int B, int S, int heads);
/* VisualFeatEncoder.forward (src/lxrt/modeling.py:507-517) as ONE tcgen05 kernel: (LN(feat·W1ᵀ + b1) + LN(box·W2ᵀ + b2)) / 2
 * as a bf16 hi/lo pair [M, 768] (hi + lo = the fp32 value to 2^-17; out_lo == NULL: out_hi is written as ONE fp16 word per
 * element, the single-word state) and identity row statistics [M, 12].  feat [M, K] and
 * W1 [768, K] are both fp32 (feat_dtype F32: kind::tf32 MMAs read them as they are) or both bf16; box [M, box_dim] f32. */
int rgqa_visn_embed(void* stream, const void* feat, int feat_dtype, const void* W1, const float* b1, const float* g1,
                    const float* be1, const float* boxes, int box_dim, const float* W2, const float* b2, const float* g2,
                    const float* be2, float eps, void* out_hi, void* out_lo, float* out_stats, int M, int K);
/* logits = A·Wᵀ + bias with the logit scorers reduced in the epilogue (the answer head's last GEMM, gqa_model.py:27 +
 * gqa_conf.py:326 / gqa_energy.py:204-206): scores [4][M] f32 in RGQA_SCORE_* order, labels [M] int32; logits_out
 * ([M, N] f32 dense) may be NULL.  flags: bit0 also ENERGY_SUM, bit1 also LSE.  work: rgqa_reduce_work_bytes(M, N) bytes. */
int rgqa_gemm_bf16_reduce_logits(void* stream, const void* A, int64_t lda, const void* W, int64_t ldw, const float* bias,
                                 float* logits_out, int M, int N, int K, int flags, void* work, int64_t work_bytes,
                                 float* scores, int32_t* labels);
int64_t rgqa_reduce_work_bytes(int M, int N);
/* out_act (bf16 [M,768]) and out_res (f32 [M,768], may be NULL) = LayerNorm(A·Wᵀ + bias + resid) * gamma + beta,
 * one kernel; N must be 768 (BertAttOutput / BertOutput, src/lxrt/modeling.py:357-361, :411-415). */
int rgqa_gemm_bf16_ln(void* stream, const void* A, int64_t lda, const void* W, int64_t ldw, const float* bias,
                      const float* resid, int64_t ldr, const float* gamma, const float* beta, float eps, void* out_act,
                      float* out_res, int M, int N, int K);
/* Deferred LayerNorm (BertAttOutput / BertOutput, src/lxrt/modeling.py:357-361, :411-415, without ever materialising
 * the normalised rows).  A hidden state is kept pre-LayerNorm as a bf16 pair (hi, lo; value = hi + lo) plus partial row
 * statistics stats[M][6][2] = (sum, sum of squares) over each 128-column slice.
 *  - rgqa_fold_ln_weight: W'[n,k] = bf16(W[n,k]·gamma[k] − mean_k(W[n,:]·gamma)) (rows centred, so x·W'ᵀ = (x − mean x)·(W∘gamma)ᵀ),
 *    c[n] = sum_k W'[n,k] (bf16 rounding residue of the centring), bias'[n] = bias[n] + sum_k W[n,k]·beta[k]
 *    (W, bias, gamma, beta f32 device pointers; W [N,K]).
 *  - rgqa_gemm_bf16_folded: out = epilogue(rstd·(A_hi·W'ᵀ − mean·c) + bias') with mean/rstd from ln_stats; ln_c may be
 *    NULL (term dropped: it is below the bf16 rounding of W'); K = 768.
 *  - rgqa_gemm_bf16_resid_hl: v = A·Wᵀ + bias + LayerNorm_prev(in_hi + in_lo); writes v as (out_hi, out_lo, out_stats). N = 768.
 *    in_beta may be NULL when `bias` already includes the previous LayerNorm's beta. */
int rgqa_fold_ln_weight(void* stream, const float* W, const float* bias, const float* gamma, const float* beta, int N,
                        int K, void* W_out_bf16, float* c_out, float* bias_out);
int rgqa_gemm_bf16_folded(void* stream, const void* A, int64_t lda, const void* W, int64_t ldw, const float* bias,
                          const void* resid, int64_t ldr, void* out, int64_t ldo, int out_dtype, int M, int N, int K,
                          int epilogue, const float* ln_stats, const float* ln_c, float eps);
int rgqa_gemm_bf16_resid_hl(void* stream, const void* A, int64_t lda, const void* W, int64_t ldw, const float* bias,
                            const void* in_hi, const void* in_lo, const float* in_stats, const float* in_gamma,
                            const float* in_beta, float eps, void* out_hi, void* out_lo, float* out_stats, int M, int N,
                            int K);
/* Single-word form of the deferred-LayerNorm state (option "state_f16", read by rgqa_finalize_weights; the default of the bf16
 * forward): the pre-LayerNorm hidden rows are ONE fp16 word per element (11 significand bits, saturating at +-65504) + the same
 * row statistics.  The consumer GEMMs read them as an fp16 A operand against an fp16 copy of their folded weight (tcgen05
 * kind::f16 wants both operands in the same format: fp16 x bf16 is an illegal instruction on B200); every other GEMM operand of
 * the forward (context rows, GELU intermediates, their weights) stays bf16.  A producer moves 4 608 instead of 7 680 B of HBM
 * traffic per row at K = 768.
 *  - rgqa_fold_ln_weight_f16: rgqa_fold_ln_weight with W' rounded to fp16 (same centring and error feedback).
 *  - rgqa_gemm_bf16_folded_a16: rgqa_gemm_bf16_folded with A and W fp16 (ln_stats may be NULL: plain epilogue(A·Wᵀ + bias)).
 *  - rgqa_qkv_attention_fused_a16: rgqa_qkv_attention_fused with A and W_hm fp16 (Q, K, V and the context stay bf16).
 *  - rgqa_gemm_bf16_resid_h16: v = A·Wᵀ + bias + LayerNorm_prev(in_h) (A, W bf16; in_h fp16); writes out_h = fp16(v) and
 *    out_stats = the partial (sum, sumsq) of the ROUNDED row.  N = 768.  src/lxrt/modeling.py:357-361, :411-415. */
int rgqa_fold_ln_weight_f16(void* stream, const float* W, const float* bias, const float* gamma, const float* beta, int N,
                            int K, void* W_out_f16, float* c_out, float* bias_out);
int rgqa_gemm_bf16_folded_a16(void* stream, const void* A, int64_t lda, const void* W, int64_t ldw, const float* bias, void* out,
                              int64_t ldo, int out_dtype, int M, int N, int K, int epilogue, const float* ln_stats,
                              const float* ln_c, float eps);
int rgqa_qkv_attention_fused_a16(void* stream, const void* A, int64_t lda, const void* W_hm, const float* bias_hm,
                                 const float* ln_stats, float ln_eps, const int32_t* key_mask, void* ctx, int64_t ldo, int B, int S,
                                 int heads);
int rgqa_gemm_bf16_resid_h16(void* stream, const void* A, int64_t lda, const void* W, int64_t ldw, const float* bias,
                             const void* in_h, const float* in_stats, const float* in_gamma, const float* in_beta, float eps,
                             void* out_h, float* out_stats, int M, int N, int K);
int rgqa_gemm_f32(void* stream, const float* A, int64_t lda, const float* W, int64_t ldw, const float* bias,
                  const float* resid, int64_t ldr, float* out, int64_t ldo, int M, int N, int K, int epilogue);
int rgqa_attention(void* stream, int dtype, const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v,
                   int64_t ldv, const int32_t* key_mask, void* out, int64_t ldo, int B, int Lq, int Lk, int heads);
int rgqa_embed_ln(void* stream, int out_dtype, const int32_t* ids, const int32_t* seg, const float* word,
                  const float* pos, const float* type, const float* gamma, const float* beta, void* out, int B,
                  int L, int H, int vocab, int max_pos, int n_types, float eps);
int rgqa_layernorm(void* stream, int out_dtype, const float* x, int64_t ldx, const float* gamma, const float* beta,
                   void* y, int M, int H, float eps);
int rgqa_visn_combine(void* stream, int out_dtype, const float* x_pre, const float* boxes, int box_dim,
                      const float* Wb, const float* bb, const float* g1, const float* b1, const float* g2,
                      const float* b2, void* out, int M, int H, float eps);
int rgqa_cast_f32_to_bf16(void* stream, const float* in, void* out, int64_t n);
/* ---- input-gradient (dgrad) kernels, fp32 path: d loss / d (feats, boxes) for ODIN (src/tasks/gqa_odin.py:97-131) and the
 * Mahalanobis input pre-processing (src/tasks/gqa_maha.py:238-268).  Linear layers: rgqa_gemm_f32 against transposed weights
 * (EPI_RESID adds the residual branch's gradient).  Orchestration: rgqa_b200/lxrt/input_grad.py. ------------------------- */
/* dx = rstd·(g − mean(g) − xh·mean(g·xh)) (+ add), g = scale·dy·gamma, xh = (x − mean x)·rstd: BertLayerNorm backward w.r.t. its
 * input x [M, H] (src/lxrt/modeling.py:258-270); add may be NULL. */
int rgqa_layernorm_bwd(void* stream, const float* dy, const float* x, const float* gamma, const float* add, float* dx, int M,
                       int H, float eps, float scale);
/* rgqa_layernorm_bwd and rgqa_ln_param_grad (below) in ONE pass over (dy, x): dx written, dgamma / dbeta accumulated.  H = 768 or 1536. */
int rgqa_layernorm_bwd_params(void* stream, const float* dy, const float* x, const float* gamma, const float* add, float* dx,
                              float* dgamma, float* dbeta, int M, int H, float eps, float scale);
/* erf-GELU (modeling.py:112-118) forward and backward on n elements; tanh backward from the OUTPUT y (BertPooler :575-581). */
int rgqa_gelu_fwd(void* stream, const float* x, float* y, int64_t n);
int rgqa_gelu_bwd(void* stream, const float* dy, const float* x, float* dx, int64_t n);
/* Mixed-precision training step: GELU with a bf16 result (n even), and BertLayerNorm writing its row twice — bf16 (the next GEMM's
 * operand) and fp32 (the residual stream / the tape) — so that neither needs a separate cast kernel. */
int rgqa_gelu_fwd_bf16(void* stream, const float* x, void* y_bf16, int64_t n);
int rgqa_layernorm_dual(void* stream, const float* x, int64_t ldx, const float* gamma, const float* beta, void* y_bf16, float* y_f32,
                        int M, int H, float eps);
int rgqa_tanh_bwd(void* stream, const float* dy, const float* y, float* dx, int64_t n);
/* BertAttention core backward (modeling.py:330-346): probabilities recomputed from q, k, v (f32, row strides ld*), gradients
 * of the three operands written to dq / dk / dv (same [row, head*64] layout, own strides).  Lq, Lk <= 64. */
int rgqa_attention_bwd(void* stream, const float* q, int64_t ldq, const float* k, int64_t ldk, const float* v, int64_t ldv,
                       const int32_t* key_mask, const float* dctx, int64_t ldo, float* dq, int64_t lddq, float* dk, int64_t lddk,
                       float* dv, int64_t lddv, int B, int Lq, int Lk, int heads, float dropout_p, uint64_t seed);
/* The same backward with bf16 operands on the tensor cores (mma.sync, fp32 accumulation and softmax; no dropout): the
 * mixed-precision training step's form, ~2e-2 relative to rgqa_attention_bwd. */
int rgqa_attention_bwd_bf16(void* stream, const float* q, int64_t ldq, const float* k, int64_t ldk, const float* v, int64_t ldv,
                            const int32_t* key_mask, const float* dctx, int64_t ldo, float* dq, int64_t lddq, float* dk, int64_t lddk,
                            float* dv, int64_t lddv, int B, int Lq, int Lk, int heads);
/* Training-mode dropout (the reference trains with p = 0.1: hidden_dropout_prob / attention_probs_dropout_prob of BertConfig).
 * Masks come from a counter-based generator keyed by (seed, element index): the backward passes regenerate them, nothing is stored;
 * they are NOT torch's masks, so parity with the reference under dropout is statistical only.
 *  - rgqa_attention_dropout: rgqa_attention (fp32) with dropout on the attention probabilities (modeling.py:341); pass the same
 *    (dropout_p, seed) to rgqa_attention_bwd.
 *  - rgqa_dropout_add: y = dropout(x) [+ resid], n elements; the backward of the dropout is the same call on the gradient. */
int rgqa_attention_dropout(void* stream, const float* q, int64_t ldq, const float* k, int64_t ldk, const float* v, int64_t ldv,
                           const int32_t* key_mask, float* out, int64_t ldo, int B, int Lq, int Lk, int heads, float dropout_p,
                           uint64_t seed);
/* out bf16 [C, R] = transpose(in fp32 [R, C], row stride ld): operand re-layout + cast of the weight-gradient products
 * (dy^T x of src/tasks/gqa_conf.py:196's backward runs over the rows), one pass. */
int rgqa_transpose_cast_bf16(void* stream, const float* in, int64_t ld, void* out, int R, int C);
/* the same pass also ADDS the column sums of `in` to colsum [C] (the nn.Linear bias gradient of the dy being re-laid out). */
int rgqa_transpose_cast_bf16_colsum(void* stream, const float* in, int64_t ld, void* out, int R, int C, float* colsum);
/* y = x - noise * sign(grad), sign = +1 where grad >= 0 else -1: the input pre-processing step of ODIN and of the Mahalanobis
 * scorer with --noise (src/tasks/gqa_odin.py:118-120, src/tasks/gqa_maha.py:246-252); n fp32 elements. */
int rgqa_sign_step(void* stream, const float* x, const float* grad, float* y, int64_t n, float noise);
int rgqa_dropout_add(void* stream, const float* x, const float* resid, float* y, int64_t n, float p, uint64_t seed);
/* ---- weight-gradient kernels, fp32 path (the parameter half of loss.backward(), src/tasks/gqa_conf.py:196).  Every one of
 * them ADDS to its output; the caller zeroes the gradient buffers once per step.  Orchestration: rgqa_b200/lxrt/trainer.py. */
/* dW[N, K] += dY[M, N]^T · X[M, K]  (nn.Linear weight; row strides lda / ldb / ldo). */
int rgqa_gemm_f32_tn(void* stream, const float* A, int64_t lda, const float* B, int64_t ldb, float* out, int64_t ldo, int M, int N,
                     int K);
/* out[M, N] (f32, row stride ldo) += A[M, K] · W[N, K]^T with bf16 operands on the tensor cores, accumulated atomically (TMA
 * reduce-add / red.global) with the K range split over several CTAs: the weight-gradient product dW += dY^T · X of the
 * mixed-precision training step, whose output is a handful of tiles and whose K is the batch's token rows. */
int rgqa_gemm_bf16_accum(void* stream, const void* A, int64_t lda, const void* W, int64_t ldw, float* out, int64_t ldo, int M, int N,
                         int K);
/* The same accumulation from TRANSPOSED operands, out[M, N] += At[K, M]^T · Bt[K, N] (bf16, row-major, row strides ldat / ldbt): the
 * tensor cores read both tiles MN-major, so dW += dY^T · X needs no transposing copies of dY and X.  K % 64 == 0, N % 32 == 0. */
int rgqa_gemm_bf16_accum_tn(void* stream, const void* At, int64_t ldat, const void* Bt, int64_t ldbt, float* out, int64_t ldo, int M,
                            int N, int K);
/* out bf16 [R, C] (row stride ldo) = bf16(in f32 [R, C], row stride ld); colsum (may be NULL) [C] += column sums of `in`: one pass
 * over a linear layer's dy gives the bf16 operand of its dgrad / wgrad GEMMs and its bias gradient. */
int rgqa_cast_bf16_colsum(void* stream, const float* in, int64_t ld, void* out, int64_t ldo, int R, int C, float* colsum);
/* out[N] += sum over the M rows of X[M, N]  (nn.Linear bias). */
int rgqa_colsum(void* stream, const float* X, int64_t ldx, float* out, int M, int N);
/* BertLayerNorm parameters: dgamma[c] += sum_m scale·dy[m,c]·xh[m,c], dbeta[c] += sum_m scale·dy[m,c]; x is the LayerNorm INPUT
 * (xh recomputed); H = 768 or 1536. */
int rgqa_ln_param_grad(void* stream, const float* dy, const float* x, float* dgamma, float* dbeta, int M, int H, float eps,
                       float scale);
/* nn.Embedding: table[idx[r], :] += g[r, :] for r < rows; indices outside [0, table_rows) and idx == padding_idx are skipped
 * (pass -1 for none; the reference's three BertEmbeddings tables all use padding_idx = 0, src/lxrt/modeling.py:272-274). */
int rgqa_embedding_scatter(void* stream, const float* g, const int32_t* idx, float* table, int rows, int H, int table_rows,
                           int padding_idx);
/* ---- RP-finetune groundwork (the backward graph itself is not built yet) ---------------- */
/* loss = BCEWithLogitsLoss(mean)(logit, target) * A (src/tasks/gqa_conf.py:193-194): adds the loss to *loss_acc (device
 * f64, zeroed by the caller) and writes dlogit = (sigmoid(logit) - target) / B when dlogit != NULL.  [B, A] f32, dense. */
int rgqa_bce_loss_grad(void* stream, const float* logit, const float* target, int B, int A, float* dlogit, double* loss_acc);
/* *acc += sum g[i]^2 (device f64): called once per gradient tensor, gives the global norm of clip_grad_norm_ (:197). */
int rgqa_grad_sumsq(void* stream, const float* g, int64_t n, double* acc);
/* BertAdam.step for one tensor (src/lxrt/optimization.py:127-176; no bias correction): gradients scaled by
 * min(1, max_norm / (sqrt(*grad_sumsq_acc) + 1e-6)) when grad_sumsq_acc != NULL and max_norm > 0, then
 * m = b1 m + (1-b1) g, v = b2 v + (1-b2) g^2, p -= lr_scheduled * (m / (sqrt(v) + eps) + weight_decay * p). */
int rgqa_bertadam_step(void* stream, float* p, const float* g, float* m, float* v, int64_t n, float lr_scheduled, double b1,
                       double b2, float eps, float weight_decay, const double* grad_sumsq_acc, float max_norm);
/* Data pass of GQA.sample_estimator (src/tasks/gqa_maha.py:138-168) for one batch: x [B, D] f32 pooled features,
 * target [B, >= C] f32 multi-hot rows (row stride ld_target; the reference drops the last class, :141).  Adds to the
 * caller's zero-initialised fp64 device accumulators: class_count[c] += 1 and class_sum[c, :] += x_i for every class c
 * sample i is labelled with; G += w_i * x_i x_i^T with w_i = that sample's label count.  w_scratch: [B] f32 device
 * scratch.  The finish (class means, centred covariance, pseudo-inverse; :170-183) is host work over these sums. */
int rgqa_estimator_accumulate(void* stream, const float* x, const float* target, int64_t ld_target, int B, int C, int D,
                              double* G, double* class_sum, double* class_count, float* w_scratch);
/* Ensemble of K <= 8 checkpoints (src/tasks/gqa_ensemble.py:95-108, :165-177): `logits` is a HOST array of K device
 * pointers to [B, A] f32 logit matrices (row stride ld); per class sigmoid values are summed and divided by K
 * (method 0, "mean") or multiplied (method 1, "multiply") in checkpoint order; label = first arg-max, score = its value. */
int rgqa_score_ensemble(void* stream, const float* const* logits, int K, int method, int64_t ld, int B, int A, float* score,
                        int32_t* label);
int rgqa_score_logits(void* stream, int mode, const float* logits, int64_t ld, int B, int A, float* score,
                      int32_t* label);
/* Stand-alone Mahalanobis: x [B,H], mu [A,H], P [H,H] (symmetric) all f32 device pointers;
 * work: >= B*(H+A+1) + 4 + A*(H+1) floats of device scratch. */
int rgqa_mahalanobis(void* stream, const float* x, const float* mu, const float* P, int B, int A, int H,
                     float* work, float* score, int32_t* label, float* full_out);

#ifdef __cplusplus
}
#endif
#endif /* RGQA_B200_H_ */
