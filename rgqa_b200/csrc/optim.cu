// rgqa_b200 — groundwork for the RP-finetune step: the pieces of GQA.train
// (src/tasks/gqa_conf.py:148-202) that do not need the backward graph.
//
//   bce_loss_grad   loss = BCEWithLogitsLoss(mean)(logit, target) * A            gqa_conf.py:193-194
//                   dlogit = (sigmoid(logit) - target) / B                      (mean over B*A elements, times A)
//   grad_sumsq      sum of squares of a gradient tensor into an fp64 accumulator: the global norm of
//                   nn.utils.clip_grad_norm_(model.parameters(), 5.)            gqa_conf.py:197
//   bertadam_step   BertAdam.step (src/lxrt/optimization.py:127-176), one tensor:
//                   g' = g * min(1, max_norm / (norm + 1e-6));  m = b1 m + (1-b1) g';  v = b2 v + (1-b2) g'^2;
//                   u = m / (sqrt(v) + e) + wd * p;  p -= lr_scheduled * u       (no bias correction, :173-176)
#include "common.cuh"
#include "kernels.h"

namespace rgqa {
namespace {

__global__ void __launch_bounds__(256)
bce_loss_grad_kernel(const float* __restrict__ logit, const float* __restrict__ target, long long n, float inv_b,
                     float* __restrict__ dlogit, double* __restrict__ loss_acc) {
  double local = 0.0;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float z = logit[i], t = target[i];
    // max(z, 0) - z t + log(1 + exp(-|z|)): the numerically stable form torch uses
    const float l = fmaxf(z, 0.f) - z * t + log1pf(expf(-fabsf(z)));
    local += static_cast<double>(l);
    if (dlogit != nullptr) dlogit[i] = (1.0f / (1.0f + expf(-z)) - t) * inv_b;
  }
  for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
  __shared__ double part[8];
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = local;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int w = 0; w < 8; ++w) s += part[w];
    atomicAdd(loss_acc, s * static_cast<double>(inv_b));
  }
}

__global__ void __launch_bounds__(256) grad_sumsq_kernel(const float* __restrict__ g, long long n, double* __restrict__ acc) {
  double local = 0.0;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    const double v = static_cast<double>(g[i]);
    local = fma(v, v, local);
  }
  for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
  __shared__ double part[8];
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = local;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int w = 0; w < 8; ++w) s += part[w];
    atomicAdd(acc, s);
  }
}

__global__ void __launch_bounds__(256)
bertadam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, long long n,
                float lr, float b1, float omb1, float b2, float omb2, float eps, float wd, const double* __restrict__ sumsq,
                float max_norm) {
  float coef = 1.0f;
  if (sumsq != nullptr && max_norm > 0.f) {
    const float norm = static_cast<float>(sqrt(*sumsq));
    coef = fminf(max_norm / (norm + 1e-6f), 1.0f);  // torch.nn.utils.clip_grad_norm_
  }
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float gi = g[i] * coef;
    const float mi = __fadd_rn(__fmul_rn(m[i], b1), __fmul_rn(omb1, gi));                  // next_m.mul_(b1).add_(1 - b1, grad)
    const float vi = __fadd_rn(__fmul_rn(v[i], b2), __fmul_rn(__fmul_rn(omb2, gi), gi));     // .addcmul_(1 - b2, grad, grad)
    float u = __fdiv_rn(mi, __fadd_rn(__fsqrt_rn(vi), eps));
    if (wd > 0.f) u = __fadd_rn(u, __fmul_rn(wd, p[i]));
    m[i] = mi;
    v[i] = vi;
    p[i] = __fadd_rn(p[i], -__fmul_rn(lr, u));
  }
}

inline int grid_for(long long n) {
  long long g = (n + 255) / 256;
  return static_cast<int>(g > 148 * 8 ? 148 * 8 : (g < 1 ? 1 : g));
}

}  // namespace

int bce_loss_grad(cudaStream_t stream, const float* logit, const float* target, int B, int A, float* dlogit, double* loss_acc) {
  RGQA_CHECK(logit && target && loss_acc && B > 0 && A > 0, "bce_loss_grad: bad arguments");
  const long long n = static_cast<long long>(B) * A;
  bce_loss_grad_kernel<<<grid_for(n), 256, 0, stream>>>(logit, target, n, 1.0f / static_cast<float>(B), dlogit, loss_acc);
  note_launch();
  RGQA_CUDA_OK(cudaGetLastError());
  return 0;
}
int grad_sumsq(cudaStream_t stream, const float* g, long long n, double* acc) {
  RGQA_CHECK(g && acc && n > 0, "grad_sumsq: bad arguments");
  grad_sumsq_kernel<<<grid_for(n), 256, 0, stream>>>(g, n, acc);
  note_launch();
  RGQA_CUDA_OK(cudaGetLastError());
  return 0;
}
int bertadam_step(cudaStream_t stream, float* p, const float* g, float* m, float* v, long long n, float lr_scheduled, double b1,
                  double b2, float eps, float weight_decay, const double* grad_sumsq_acc, float max_norm) {
  RGQA_CHECK(p && g && m && v && n > 0, "bertadam_step: bad arguments");
  // 1 - beta is formed in double like the reference's Python scalars (1 - 0.999 in float would be off by 1.3e-5)
  bertadam_kernel<<<grid_for(n), 256, 0, stream>>>(p, g, m, v, n, lr_scheduled, static_cast<float>(b1), static_cast<float>(1.0 - b1),
                                                   static_cast<float>(b2), static_cast<float>(1.0 - b2), eps, weight_decay,
                                                   grad_sumsq_acc, max_norm);
  note_launch();
  RGQA_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace rgqa
