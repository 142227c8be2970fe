// rgqa_b200 — the realistic-VQA rejection scorers over the 1842-way answer logits and the pooled
// feature.  One warp per question; every reduction is a lane-strided pass plus warp shuffles.
//
//   MSP          score,label = sigmoid(logit).max(1)                         src/tasks/gqa_conf.py:326
//   ENERGY_TOP2  label as MSP; score = sum_{top-2} log(1 + exp(logit))       src/tasks/gqa_energy.py:204-206
//   ENERGY_SUM   label as MSP; score = sum_all log(1 + exp(logit))           src/tasks/gqa_energy.py:185
//   LSE          label as MSP; score = logsumexp(logit)                      src/tasks/gqa_energy.py:131
//   MAHA         S[b,c] = -1/2 (x_b-mu_c)^T P (x_b-mu_c); score,label = S.max(1)   src/tasks/gqa_maha.py:228-236,270
//                evaluated as -1/2 (q_b - 2 G_bc + m_c) with t = x P, q = <t,x>, G = t mu^T, m_c = mu_c^T P mu_c
//   ENSEMBLE     per-class sigmoid vectors of K checkpoints summed (then / K) or multiplied, in checkpoint order,
//                label = argmax, score = that entry                           src/tasks/gqa_ensemble.py:95-108,165-177
#include "common.cuh"
#include "kernels.h"

namespace rgqa {
namespace {

constexpr int kWarps = 4;

struct ArgMax {
  float v;
  int i;
};
// larger value wins; on equal values the lower index wins (first-occurrence semantics of torch.max)
__device__ __forceinline__ ArgMax better(ArgMax a, ArgMax b) {
  if (b.v > a.v || (b.v == a.v && b.i < a.i)) return b;
  return a;
}
__device__ __forceinline__ ArgMax warp_argmax(ArgMax a) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    ArgMax b;
    b.v = __shfl_xor_sync(0xffffffffu, a.v, o);
    b.i = __shfl_xor_sync(0xffffffffu, a.i, o);
    a = better(a, b);
  }
  return a;
}

__device__ __forceinline__ float sigmoid_ref(float x) { return 1.0f / (1.0f + expf(-x)); }
__device__ __forceinline__ float softplus_naive(float x) { return logf(1.0f + expf(x)); }  // as written in the reference

__global__ void __launch_bounds__(kWarps * 32)
score_logits_kernel(int mode, const float* __restrict__ logits, long long ld, int B, int A, float* __restrict__ score,
                    int32_t* __restrict__ label) {
  const int row = blockIdx.x * kWarps + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= B) return;
  const float* x = logits + static_cast<long long>(row) * ld;

  ArgMax best{-INFINITY, 0x7fffffff};
  float m1 = -INFINITY, m2 = -INFINITY;  // two largest logits seen by this lane
  float sp_sum = 0.f;
  for (int c = lane; c < A; c += 32) {
    const float v = x[c];
    const float s = sigmoid_ref(v);
    best = better(best, ArgMax{s, c});
    if (v > m1) {
      m2 = m1;
      m1 = v;
    } else if (v > m2) {
      m2 = v;
    }
    if (mode == SCORE_ENERGY_SUM) sp_sum += softplus_naive(v);
  }
  best = warp_argmax(best);
  float out = best.v;
  if (mode == SCORE_ENERGY_TOP2) {
    // merge per-lane (m1, m2) pairs into the global top-2 (duplicates count twice, like topk)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float o1 = __shfl_xor_sync(0xffffffffu, m1, o);
      const float o2 = __shfl_xor_sync(0xffffffffu, m2, o);
      const float n1 = fmaxf(m1, o1);
      const float n2 = fmaxf(fminf(m1, o1), fmaxf(m2, o2));
      m1 = n1;
      m2 = n2;
    }
    out = softplus_naive(m1) + (A > 1 ? softplus_naive(m2) : 0.f);
  } else if (mode == SCORE_ENERGY_SUM) {
    out = warp_sum(sp_sum);
  } else if (mode == SCORE_LSE) {
    const float gmax = warp_max(m1);
    float e = 0.f;
    for (int c = lane; c < A; c += 32) e += expf(x[c] - gmax);
    out = gmax + logf(warp_sum(e));
  }
  if (lane == 0) {
    score[row] = out;
    label[row] = best.i;
  }
}

__global__ void __launch_bounds__(kWarps * 32)
maha_finish_kernel(const float* __restrict__ q, const float* __restrict__ G, long long ldg,
                   const float* __restrict__ m, int B, int A, float* __restrict__ score,
                   int32_t* __restrict__ label, float* __restrict__ full) {
  const int row = blockIdx.x * kWarps + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= B) return;
  const float qb = q[row];
  const float* g = G + static_cast<long long>(row) * ldg;
  ArgMax best{-INFINITY, 0x7fffffff};
  for (int c = lane; c < A; c += 32) {
    const float s = -0.5f * ((qb - 2.0f * g[c]) + __ldg(m + c));
    if (full != nullptr) full[static_cast<long long>(row) * A + c] = s;
    best = better(best, ArgMax{s, c});
  }
  best = warp_argmax(best);
  if (lane == 0) {
    score[row] = best.v;
    label[row] = best.i;
  }
}

__global__ void __launch_bounds__(kWarps * 32)
rowdot_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out, int M, int H) {
  const int row = blockIdx.x * kWarps + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= M) return;
  const float* ar = a + static_cast<long long>(row) * H;
  const float* br = b + static_cast<long long>(row) * H;
  float s = 0.f;
  for (int c = lane; c < H; c += 32) s = fmaf(ar[c], br[c], s);
  s = warp_sum(s);
  if (lane == 0) out[row] = s;
}

// K <= 8 logit matrices [B, A] (one per checkpoint); float32 accumulation in checkpoint order like the reference's
// in-place numpy `+=` / `*=` on float32 arrays, then `/ K` for the mean; first-index arg-max (np.argmax)
struct EnsemblePtrs {
  const float* p[8];
};
__global__ void __launch_bounds__(kWarps * 32)
ensemble_kernel(EnsemblePtrs logits, int K, int multiply, long long ld, int B, int A, float* __restrict__ score,
                int32_t* __restrict__ label) {
  const int row = blockIdx.x * kWarps + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= B) return;
  ArgMax best{-INFINITY, 0x7fffffff};
  const float kf = static_cast<float>(K);
  for (int c = lane; c < A; c += 32) {
    float v = sigmoid_ref(logits.p[0][static_cast<long long>(row) * ld + c]);
    for (int k = 1; k < K; ++k) {
      const float s = sigmoid_ref(logits.p[k][static_cast<long long>(row) * ld + c]);
      v = multiply ? v * s : v + s;
    }
    if (!multiply) v = v / kf;
    best = better(best, ArgMax{v, c});
  }
  best = warp_argmax(best);
  if (lane == 0) {
    score[row] = best.v;
    label[row] = best.i;
  }
}

}  // namespace

int score_ensemble(cudaStream_t stream, const float* const* logits, int K, int multiply, long long ld, int B, int A,
                   float* score, int32_t* label) {
  RGQA_CHECK(logits != nullptr && K >= 1 && K <= 8, "score_ensemble: between 1 and 8 logit matrices (got %d)", K);
  RGQA_CHECK(B > 0 && A > 0 && score != nullptr && label != nullptr, "score_ensemble: bad arguments");
  EnsemblePtrs ptrs{};
  for (int k = 0; k < K; ++k) {
    RGQA_CHECK(logits[k] != nullptr, "score_ensemble: null logit matrix %d", k);
    ptrs.p[k] = logits[k];
  }
  ensemble_kernel<<<(B + kWarps - 1) / kWarps, kWarps * 32, 0, stream>>>(ptrs, K, multiply, ld, B, A, score, label);
  note_launch();
  RGQA_CUDA_OK(cudaGetLastError());
  return 0;
}

int score_logits(cudaStream_t stream, int mode, const float* logits, long long ld, int B, int A, float* score,
                 int32_t* label) {
  RGQA_CHECK(B > 0 && A > 0, "score_logits: empty input B=%d A=%d", B, A);
  RGQA_CHECK(mode >= SCORE_MSP && mode <= SCORE_LSE, "score_logits: mode %d is not a logit scorer", mode);
  score_logits_kernel<<<(B + kWarps - 1) / kWarps, kWarps * 32, 0, stream>>>(mode, logits, ld, B, A, score, label);
  note_launch();
  RGQA_CUDA_OK(cudaGetLastError());
  return 0;
}

int maha_finish(cudaStream_t stream, const float* q, const float* G, long long ldg, const float* m, int B, int A,
                float* score, int32_t* label, float* full) {
  RGQA_CHECK(B > 0 && A > 0, "maha_finish: empty input");
  maha_finish_kernel<<<(B + kWarps - 1) / kWarps, kWarps * 32, 0, stream>>>(q, G, ldg, m, B, A, score, label, full);
  note_launch();
  RGQA_CUDA_OK(cudaGetLastError());
  return 0;
}

int rowdot(cudaStream_t stream, const float* a, const float* b, float* out, int M, int H) {
  RGQA_CHECK(M > 0 && H > 0, "rowdot: empty input");
  rowdot_kernel<<<(M + kWarps - 1) / kWarps, kWarps * 32, 0, stream>>>(a, b, out, M, H);
  note_launch();
  RGQA_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace rgqa
