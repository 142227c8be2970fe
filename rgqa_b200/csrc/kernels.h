// rgqa_b200 — internal launcher declarations (C++ side, below the C ABI in include/rgqa_b200.h).
// Every launcher returns 0 on success; on failure it records a message (rgqa_last_error) and returns 1.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

struct CUtensorMap_st;  // <cuda.h>, only needed by the tcgen05 translation units

namespace rgqa {

enum Epilogue : int {
  EPI_BIAS = 0,   // out = acc + bias
  EPI_GELU = 1,   // out = gelu_erf(acc + bias)            (modeling.py:398-401 + :112-118)
  EPI_TANH = 2,   // out = tanh(acc + bias)                (modeling.py:575-581 pooler)
  EPI_RESID = 3,  // out = acc + bias + residual           (pre-LayerNorm sum of modeling.py:357-361 / :411-415)
  EPI_RESID_HL = 4,  // same sum, residual recomputed from a deferred-LayerNorm hidden state, written as a bf16 hi/lo pair
  EPI_RESID_HL_BETA = 5,  // EPI_RESID_HL with LN_prev's beta passed separately (the engine folds it into the bias at load time)
  EPI_REDUCE = 6,  // acc + bias reduced along each row inside the epilogue (scorers of the answer head / Mahalanobis), see RowReduce
  EPI_RESID_H16 = 7,       // EPI_RESID_HL with the state as ONE fp16 word per element (no lo): half the residual bytes
  EPI_RESID_H16_BETA = 8,  // ... with LN_prev's beta passed separately
  EPI_ACCUM = 9,           // out (fp32) += acc, atomically (TMA reduce-add / red.global): weight gradients, split along K
};

// Row reductions fused into a GEMM epilogue (gemm_bf16_reduce).  Every (row tile, column tile) CTA leaves one partial
// record per (row, column half); the CTA that finishes a row tile LAST (atomic arrival counter, no waiting) merges that
// tile's records and writes the final scores, so the [M, N] matrix itself never has to reach HBM.
//   RED_LOGITS  v = acc + bias[n]: arg-max of sigmoid(v) (first index on ties, torch.max), two largest v, sum of
//               log(1 + exp(v)), logsumexp  ->  scores [4][M] in ScoreMode order (MSP, ENERGY_TOP2, ENERGY_SUM, LSE)
//               src/tasks/gqa_conf.py:326, gqa_energy.py:131,185,204-206
//   RED_MAHA    columns n < split are t = x·P: q += acc * x[row, n]; columns n >= split are classes c = n - split with
//               v = acc - m[c]/2 (acc = x·P·mu_c): arg-max_c v  ->  scores[row] = max_c v - q/2 = max_c -1/2 (x-mu_c)^T P (x-mu_c)
//               src/tasks/gqa_maha.py:228-236,270
enum RowReduceKind : int { RED_NONE = 0, RED_LOGITS = 1, RED_MAHA = 2 };
struct RowReduce {
  int kind = RED_NONE;
  int flags = 0;            // RED_LOGITS: bit0 also the all-class softplus sum, bit1 also logsumexp (both cost exp/log per element)
  int split = 0;            // RED_MAHA: first class column
  const float* x = nullptr;  // RED_MAHA: fp32 [M, split] rows the t columns are dotted with
  const float* m = nullptr;  // RED_MAHA: m_c = mu_c^T P mu_c, [N - split]
  float* part = nullptr;     // workspace: [ceil(M/128)*128][2*ceil(N/128)][8] floats
  int* counter = nullptr;    // workspace: [ceil(M/128)] arrival counters, zero on entry, left zero on exit
  float* scores = nullptr;   // RED_LOGITS: [4][M]; RED_MAHA: [M]
  int32_t* labels = nullptr;  // [M]
};
// bytes of `part` / ints of `counter` for an M x N problem
inline size_t row_reduce_part_bytes(int M, int N) { return static_cast<size_t>((M + 127) / 128) * 128 * 2 * ((N + 127) / 128) * 8 * sizeof(float); }
inline int row_reduce_counters(int M) { return (M + 127) / 128; }

enum DType : int { DT_F32 = 0, DT_BF16 = 1, DT_F16 = 2 /* the fp16 deferred-LayerNorm state (embed_ln / visn_combine outputs, GEMM A operands) */ };

// ---- GEMMs -------------------------------------------------------------------------------------
// bf16 tensor-core path: A [M,K] bf16 (row stride lda), W [N,K] bf16 (row stride ldw), bias fp32,
// resid fp32 [M,*] (EPI_RESID: the fp32 residual stream), out bf16 or fp32.
int gemm_bf16(cudaStream_t stream, const void* A, long long lda, const void* W, long long ldw, const float* bias,
              const void* resid, long long ldr, void* out, long long ldo, int out_f32, int M, int N, int K, int epi);
// Same GEMM with a DEFERRED LayerNorm on its A operand: A = bf16 of the pre-LayerNorm rows, W already scaled by gamma
// and centred (fold_ln_weight), bias = b + W·beta, ln_stats = [M][6][2] partial (sum, sumsq) per row; the epilogue
// computes rstd*acc + bias[n] before the activation.  ln_c (optional) = sum_k W'[n,k], the bf16 rounding residue of the
// centring: when given, the epilogue also subtracts mu*c[n] from acc (exact form, one more FMA per element).  K = 768.
int gemm_bf16_folded(cudaStream_t stream, const void* A, long long lda, const void* W, long long ldw, const float* bias,
                     const void* resid, long long ldr, void* out, long long ldo, int out_f32, int M, int N, int K, int epi,
                     const float* ln_stats, const float* ln_c, float ln_eps, long long ln_stats_ld = 12,
                     int a_f16 = 0 /* A AND W are fp16 (the fp16 deferred-LayerNorm state against fp16 weight copies) */);
// out = A·Wᵀ + bias reduced per row in the epilogue (see RowReduce); `out` (fp32 [M, N], row stride ldo) may be nullptr: the
// matrix is then never written.  Single-CTA 128 x 128 tiles (the head GEMMs have M = batch rows).
// out[M, N] (fp32, row stride ldo) += A[M, K] · W[N, K]ᵀ, bf16 operands, accumulated ATOMICALLY so that the K range can be split
// over several CTAs (weight gradients of the training step: M, N = a weight's shape, K = the batch's token rows — a handful of
// output tiles, thousands of k-blocks).  The summation order varies from run to run, like every other wgrad kernel here.
int gemm_bf16_accum(cudaStream_t stream, const void* A, long long lda, const void* W, long long ldw, float* out, long long ldo,
                    int M, int N, int K);
// The same product from TRANSPOSED operands: out[M, N] += At[K, M]ᵀ · Bt[K, N] (both row-major, read MN-major by the tensor
// cores): dW += dYᵀ · X straight from the bf16 rows of dY and X, no transposing copies.  K % 64 == 0, N % 32 == 0.
int gemm_bf16_accum_tn(cudaStream_t stream, const void* At, long long ldat, const void* Bt, long long ldbt, float* out, long long ldo,
                       int M, int N, int K);
// bf16(in) for an fp32 [R, C] view (row stride ld) into a bf16 [R, C] view (row stride ldo), optionally adding the column sums of `in`
// to colsum [C]
int cast_bf16_colsum(cudaStream_t stream, const float* in, long long ld, void* out, long long ldo, int R, int C, float* colsum);
int gemm_bf16_reduce(cudaStream_t stream, const void* A, long long lda, const void* W, long long ldw, const float* bias,
                     float* out, long long ldo, int M, int N, int K, const RowReduce& red);
// Q|K|V projection with a head-major output: out[plane = head*3 + {q,k,v}][row][64] bf16, planes plane_rows rows apart; `out`
// points at this call's first row of plane 0.  ln_stats != nullptr: deferred-LayerNorm consumer (folded weight / bias).
int gemm_bf16_qkv_planes(cudaStream_t stream, const void* A, long long lda, const void* W, long long ldw, const float* bias,
                         void* out, long long plane_rows, int heads, int M, int K, const float* ln_stats, float ln_eps,
                         int a_f16 = 0);
// attention over the head-major planes of gemm_bf16_qkv_planes (TMA kernel only): queries are rows [q_row0, q_row0 + B*Lq) and
// keys / values rows [kv_row0, kv_row0 + B*Lk) of the same plane tensor; out [B*Lq, heads*64] bf16 (row stride ldo).
int attention_planes(cudaStream_t stream, const void* planes, long long plane_rows, long long q_row0, long long kv_row0,
                     const int32_t* key_mask, void* out, long long ldo, int B, int Lq, int Lk, int heads);
// 1 when attention_planes can serve (B, Lq, Lk): enough units for the persistent kernel and room for two slots per warp
int attention_planes_ok(int B, int Lq, int Lk, int heads);
// Producer of a deferred-LayerNorm hidden state: v = A·Wᵀ + bias + LN_prev(x_prev) with x_prev = in_hi + in_lo (bf16
// pair), LN_prev from (in_stats, in_gamma, in_beta); writes v as out_hi = bf16(v), out_lo = bf16(v - out_hi) and the
// partial (sum, sumsq) statistics [M][6][2].  N must be 768.
int gemm_bf16_resid_hl(cudaStream_t stream, const void* A, long long lda, const void* W, long long ldw, const float* bias,
                       const void* in_hi, const void* in_lo, const float* in_stats, const float* in_gamma,
                       const float* in_beta, float eps, void* out_hi, void* out_lo, float* out_stats, int M, int N, int K);
// The same producer over the single-word state: x_prev = in_h (fp16 [M, 768]), v written as out_h = fp16(v) (saturating at
// +-65504) and the partial statistics of the rounded row.  4 608 instead of 7 680 B of HBM traffic per row at K = 768.
int gemm_bf16_resid_h16(cudaStream_t stream, const void* A, long long lda, const void* W, long long ldw, const float* bias,
                        const void* in_h, const float* in_stats, const float* in_gamma, const float* in_beta, float eps,
                        void* out_h, float* out_stats, int M, int N, int K);
// W'[n,k] = bf16(W[n,k]*gamma[k] - mean_k(W[n,:]*gamma)), c[n] = sum_k float(W'[n,k]) (the bf16 rounding residue of the
// centring), bias_out[n] = bias[n] + sum_k W[n,k]*beta[k]
int fold_ln_weight(cudaStream_t stream, const float* W, const float* bias, const float* gamma, const float* beta, int N,
                   int K, void* W_out_bf16, float* c_out, float* bias_out, int f16 = 0 /* W' as fp16 instead of bf16 */);
// A normalised hidden state y (bf16 rows `hi` + their fp32 copy) as a deferred-LayerNorm state whose LayerNorm is the
// identity: lo = bf16(y - hi), stats = (0, 768, 0, ...) so that mean = 0 and rstd = 1.  lo == nullptr: the fp16 single-word
// state — `hi` is rewritten as fp16(y).
int identity_state(cudaStream_t stream, const float* y_f32, void* hi, void* lo, float* stats, long long M, int H);
// fp32 SIMT path (precise mode): everything fp32.
int gemm_f32(cudaStream_t stream, const float* A, long long lda, const float* W, long long ldw, const float* bias,
             const float* resid, long long ldr, float* out, long long ldo, int M, int N, int K, int epi);

// Self-attention inside the Q|K|V projection (qkv_attn_tcgen05.cu; modeling.py:320-347 for blocks whose queries and keys are
// the same rows): ctx[M, heads*64] = attention(LN(A)·W_hmᵀ + b_hm) with W_hm / b_hm packed head-major ([head][q|k|v][64] rows),
// A = bf16 [B*S, 768] pre-LayerNorm rows + ln_stats (or already-normalised rows and nullptr), key_mask [B, S] or nullptr.
int qkv_attention_fused(cudaStream_t stream, const void* A, long long lda, const void* W_hm, const float* bias_hm,
                        const float* ln_stats, float ln_eps, const int32_t* key_mask, void* ctx, long long ldo, int B, int S,
                        int heads, int a_f16 = 0 /* A rows are the fp16 state and W_hm is fp16 */);
int qkv_attention_fused_ok(int B, int S, int heads, int K);
// VisualFeatEncoder in ONE kernel (visn_embed_tcgen05.cu; modeling.py:507-517): (LN1(feat·W1ᵀ + b1) + LN2(box·W2ᵀ + b2)) / 2
// written as the identity deferred-LayerNorm state (hi, lo bf16 [M,768], stats [M,12]).  feat / W1 are fp32 (kind::tf32
// MMAs, no cast kernel) when feat_f32, else bf16.  W2p = box_fc weight padded to [768][8]; bstat from make_box_stats.
int visn_embed(cudaStream_t stream, const void* feat, int feat_f32, long long ldf, const void* W1, const float* b1, const float* g1,
               const float* be1, const float* boxes, int box_dim, const float* W2p, const float* b2, const float* g2,
               const float* be2, const float* bstat, float eps, void* hi, void* lo, float* stats, int M, int K);
// Full-row GEMM + LayerNorm (bf16 path): out_act/out_res = LN(A·Wᵀ + bias + resid) for N == 768, one kernel
// (BertAttOutput / BertOutput, modeling.py:357-361, :411-415).  out_res may be null.
int gemm_bf16_ln(cudaStream_t stream, const void* A, long long lda, const void* W, long long ldw, const float* bias,
                 const float* resid, long long ldr, const float* gamma, const float* beta, float eps, void* out_act,
                 float* out_res, int M, int N, int K);
// host helpers shared by the tcgen05 GEMM translation units (gemm_tcgen05.cu)
int make_tmap(::CUtensorMap_st* map, const void* ptr, long long rows, long long cols, long long ld, int box_rows,
              int elem_bytes);
// general form: explicit box width (elements) and swizzle span in bytes (0 = none, 64, 128)
int make_tmap_ex(::CUtensorMap_st* map, const void* ptr, long long rows, long long cols, long long ld, int box_rows,
                 int box_cols, int elem_bytes, int swizzle_bytes);
int make_tmap_3d(::CUtensorMap_st* map, const void* ptr, long long d0, long long d1, long long d2, long long stride1_bytes,
                 long long stride2_bytes, int b0, int b1, int b2);
int num_sms();
// SMs a persistent kernel launched NOW may occupy (its grid): num_sms() unless the engine has set a budget for the stream it is
// enqueueing on (option "sm_split": the language and visual chains of a forward run side by side on disjoint SM sets)
int launch_sms();
void set_launch_sms(int sms);   // 0 = no budget

// ---- attention (small sequences, Lq, Lk <= 64, 12 heads x 64) -----------------------------------
// q/k/v point at the first row of sample 0 (head 0); rows are ldq/ldk/ldv elements apart; sample b
// owns rows [b*Lq, (b+1)*Lq) of q/out and [b*Lk, (b+1)*Lk) of k/v.  key_mask: int32 [B, Lk] (1 = keep)
// or nullptr; masked keys get the reference's additive -10000 (modeling.py:865), not -inf.
int attention(cudaStream_t stream, int dtype, const void* q, long long ldq, const void* k, long long ldk,
              const void* v, long long ldv, const int32_t* key_mask, void* out, long long ldo, int B, int Lq,
              int Lk, int heads, float drop_p = 0.f,
              unsigned long long drop_seed = 0)  /* drop_p > 0: attention-probability dropout (fp32 training step) */;

// ---- row-wise ops ------------------------------------------------------------------------------
// LN(word[ids] + pos[0..L) + type[seg]) (modeling.py:278-292); tables and gamma/beta fp32.
int embed_ln(cudaStream_t stream, int out_dtype, const int32_t* ids, const int32_t* seg, const float* word,
             const float* pos, const float* type, const float* gamma, const float* beta, void* out, float* out32, int B,
             int L, int H, int vocab, int max_pos, int n_types, float eps, void* lo_out = nullptr, float* stats_out = nullptr);
// y = LN(x) * gamma + beta; x fp32 [M,H] (row stride ldx), y fp32 or bf16 [M,H] dense.
// Every row kernel takes an optional fp32 copy pointer (y32 / out32): in the bf16 path the hidden state is
// kept twice, bf16 as the next GEMM's A operand and fp32 as the residual stream.
int layernorm(cudaStream_t stream, int out_dtype, const float* x, long long ldx, const float* gamma,
              const float* beta, void* y, float* y32, int M, int H, float eps);
// (LN(x_pre) + LN(box·Wbᵀ + bb)) / 2 (modeling.py:507-517); x_pre = feat·Wfᵀ + bf, fp32 [M,H].
int visn_combine(cudaStream_t stream, int out_dtype, const float* x_pre, const float* boxes, int box_dim,
                 const float* Wb, const float* bb, const float* g1, const float* b1, const float* g2,
                 const float* b2, void* out, float* out32, int M, int H, float eps, void* lo_out = nullptr,
                 float* stats_out = nullptr);  // lo_out / stats_out (bf16 outputs): also emit the identity deferred-LayerNorm state
// UNITER image embedding tail: LN3( LN1(x_pre) + LN2(pos7·Wpᵀ + bp) + type_emb ) (uniter/modeling.py:606-612).
int uniter_img_combine(cudaStream_t stream, int out_dtype, const float* x_pre, const float* pos, int pos_dim,
                       const float* Wp, const float* bp, const float* g1, const float* b1, const float* g2,
                       const float* b2, const float* type_row, const float* g3, const float* b3, void* out,
                       float* out32, int M, int H, float eps);
int cast_f32_to_bf16(cudaStream_t stream, const float* in, void* out, long long n);
int cast_f32_to_f16(cudaStream_t stream, const float* in, void* out, long long n);  // load time: fp16 copies of consumer weights
// fp32 [M, K] -> bf16 [M, 3K] split operand (hi | hi | lo, or hi | lo | hi on the weight side): one bf16 GEMM over 3K then
// evaluates the product to ~16 mantissa bits per operand (the tensor-core Mahalanobis)
int split_bf16x3(cudaStream_t stream, const float* in, long long ld, void* out, long long M, int K, int weight_side);
// fp32 [R, C] (row stride ld) -> bf16 [C, R] in one pass (operands of the weight-gradient GEMMs)
// colsum != nullptr: colsum[c] += sum_r in[r, c] as well (the bias gradient rides on the same read)
int transpose_cast_bf16(cudaStream_t stream, const float* in, long long ld, void* out, int R, int C, float* colsum = nullptr);
int cast_bf16_to_f32(cudaStream_t stream, const void* in, float* out, long long n);
// copy rows: dst[r, :] = src[r*src_stride_rows .. ] for r < rows (strided gather of whole rows)
int gather_rows(cudaStream_t stream, int dtype, const void* src, long long src_ld, void* dst, long long dst_ld,
                int rows, int H);

// ---- scorers -----------------------------------------------------------------------------------
enum ScoreMode : int {
  SCORE_MSP = 0,          // max sigmoid prob, label = argmax sigmoid            (gqa_conf.py:326)
  SCORE_ENERGY_TOP2 = 1,  // sum_{top-2 logits} log(1+exp(x)), label as MSP      (gqa_energy.py:204-206)
  SCORE_ENERGY_SUM = 2,   // sum_all log(1+exp(x)), label as MSP                 (gqa_energy.py:185)
  SCORE_LSE = 3,          // logsumexp(logits), label as MSP                     (gqa_energy.py:131)
  SCORE_MAHA = 4,         // max_c -0.5 (x-mu_c)^T P (x-mu_c), label = argmax_c  (gqa_maha.py:228-236,270)
};
// input-gradient (dgrad) kernels of the fp32 path (backward.cu)
int layernorm_bwd(cudaStream_t stream, const float* dy, const float* x, const float* gamma, const float* add, float* dx, int M,
                  int H, float eps, float scale);
// layernorm_bwd and ln_param_grad (below) in one pass over (dy, x): dx as above, dgamma / dbeta ACCUMULATED
int layernorm_bwd_params(cudaStream_t stream, const float* dy, const float* x, const float* gamma, const float* add, float* dx,
                         float* dgamma, float* dbeta, int M, int H, float eps, float scale);
int gelu_fwd(cudaStream_t stream, const float* x, float* y, long long n);
int gelu_fwd_bf16(cudaStream_t stream, const float* x, void* y_bf16, long long n);   // n even
int gelu_bwd(cudaStream_t stream, const float* dy, const float* x, float* dx, long long n);
int tanh_bwd(cudaStream_t stream, const float* dy, const float* y, float* dx, long long n);
int attention_bwd(cudaStream_t stream, const float* q, long long ldq, const float* k, long long ldk, const float* v,
                  long long ldv, const int32_t* key_mask, const float* dctx, long long ldo, float* dq, long long lddq, float* dk,
                  long long lddk, float* dv, long long lddv, int B, int Lq, int Lk, int heads, float drop_p = 0.f,
                  unsigned long long drop_seed = 0);
// The same backward on the tensor cores (attention_bwd_mma.cu): q / k / v / dctx cast to bf16 on the way into shared memory, the five
// products as mma.sync m16n8k16 with fp32 accumulation, softmax in fp32; no dropout.  The mixed-precision training step's kernel.
int attention_bwd_bf16(cudaStream_t stream, const float* q, long long ldq, const float* k, long long ldk, const float* v,
                       long long ldv, const int32_t* key_mask, const float* dctx, long long ldo, float* dq, long long lddq, float* dk,
                       long long lddk, float* dv, long long lddv, int B, int Lq, int Lk, int heads);
// y = x - noise * sign(grad) (sign = +1 where grad >= 0): ODIN / Mahalanobis-with-noise input step (gqa_odin.py:118-120)
int sign_step(cudaStream_t stream, const float* x, const float* grad, float* y, long long n, float noise);
int dropout_add(cudaStream_t stream, const float* x, const float* resid, float* y, long long n, float p, unsigned long long seed);
// weight-gradient kernels (backward.cu); all ACCUMULATE into their outputs
int gemm_f32_tn(cudaStream_t stream, const float* A, long long lda, const float* B, long long ldb, float* out, long long ldo, int M,
                int N, int K);
int colsum(cudaStream_t stream, const float* X, long long ldx, float* out, int M, int N);
int ln_param_grad(cudaStream_t stream, const float* dy, const float* x, float* dgamma, float* dbeta, int M, int H, float eps,
                  float scale);
int embedding_scatter(cudaStream_t stream, const float* g, const int32_t* idx, float* table, int rows, int H, int table_rows,
                      int padding_idx);
// RP-finetune groundwork (optim.cu): BCE-with-logits loss * A and its logit gradient, global gradient norm, BertAdam update
int bce_loss_grad(cudaStream_t stream, const float* logit, const float* target, int B, int A, float* dlogit, double* loss_acc);
int grad_sumsq(cudaStream_t stream, const float* g, long long n, double* acc);
int bertadam_step(cudaStream_t stream, float* p, const float* g, float* m, float* v, long long n, float lr_scheduled, double b1,
                  double b2, float eps, float weight_decay, const double* grad_sumsq_acc, float max_norm);
// Mahalanobis estimator, data pass (gqa_maha.py:138-168): adds one batch to the fp64 accumulators G [D,D] (sum of w_i x x^T),
// class_sum [C,D], class_count [C]; w_scratch [B] f32 receives the per-sample label counts.
int estimator_accumulate(cudaStream_t stream, const float* x, const float* target, long long ld_t, int B, int C, int D,
                         double* G, double* class_sum, double* class_count, float* w_scratch);
int score_ensemble(cudaStream_t stream, const float* const* logits, int K, int multiply, long long ld, int B, int A,
                   float* score, int32_t* label);
int score_logits(cudaStream_t stream, int mode, const float* logits, long long ld, int B, int A, float* score,
                 int32_t* label);
// Mahalanobis finish: S[b,c] = -0.5 * (q[b] - 2 G[b,c] + m[c]); row max / argmax.
int maha_finish(cudaStream_t stream, const float* q, const float* G, long long ldg, const float* m, int B, int A,
                float* score, int32_t* label, float* full /* optional [B,A] */);
// q[b] = sum_j t[b,j] * x[b,j]
int rowdot(cudaStream_t stream, const float* a, const float* b, float* out, int M, int H);


// ---- runtime options (prof.cu): kernel-selection switches for A/B measurements ---------------------
enum Option : int { OPT_GEMM_PAIR = 0, OPT_GEMM_STAGED = 1, OPT_GEMM_DEBUG = 2, OPT_PDL = 3, OPT_GEMM_LN = 4, OPT_TWO_STREAMS = 5, OPT_GRAPHS = 6,
  OPT_DEFER_LN = 7,
  OPT_ATTN_TMA = 8,
  OPT_QKV_PLANES = 9,
  OPT_MAHA_TC = 10,    // Mahalanobis on the tensor cores (split-bf16, fused arg-max) in bf16 handles
  OPT_VISN_FUSED = 11, // visual feature encoder as one tcgen05 kernel (visn_embed_tcgen05.cu)
  OPT_ATTN_FUSED = 12, // self-attention inside the Q|K|V projection's epilogue (qkv_attn_tcgen05.cu)
  OPT_STATE_F16 = 13,  // read by rgqa_finalize_weights: deferred-LayerNorm state as ONE fp16 word per element (consumers run fp16 x fp16
                       // against fp16 copies of their folded weights) instead of a bf16 hi/lo pair
  OPT_GELU_EW16 = 14,  // FFN1 (erf-GELU consumer) with 16 epilogue warps instead of 8 (A/B measurement)
  OPT_ATTN_STRADDLE = 15,  // fused projection + attention, S = 36: seven whole samples per CTA pair, sample 3's K / V through DSMEM (off:
                           // 14 % fewer tiles, same time — the kernel is bound by its attention tasks, not by its main loop)
  OPT_PAIR_DEEP = 16,  // bf16-output consumers (Q|K|V, FFN1): six main-loop stages and one staging slab per group (A/B measurement)
  OPT_SM_SPLIT = 17,   // two-stream forward: give the language / visual kernels fixed shares of the SMs (A/B measurement)
  OPT_COUNT = 20 };
int option_get(int key);
void option_set(int key, int value);

// ---- measurement hooks (prof.cu) ----------------------------------------------------------------
void note_launch(int n = 1);
long long launch_count();
void prof_enable(bool on);
bool prof_on();
void* prof_begin(cudaStream_t st);
void prof_end(void* tok, cudaStream_t st, double flops, int kind);
int prof_collect(double* ms_sum, double* flops_sum, long long* count, int nkinds);
enum ProfKind : int { PROF_GEMM_TC = 0, PROF_GEMM_F32 = 1, PROF_ATTENTION = 2, PROF_OTHER = 3, PROF_NKINDS = 4 };

}  // namespace rgqa
