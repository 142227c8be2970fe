// rgqa_b200 — class-conditional Gaussian statistics for the Mahalanobis scorer, accumulated on the device.
//
// Restates the data pass of GQA.sample_estimator (src/tasks/gqa_maha.py:138-168): for every training sample and every
// class it is labelled with, the pooled feature joins that class's list.  The reference builds the lists with a Python
// double loop (O(N * 1842) `target[i][j] == 0` tests and a torch.cat per hit) on the host; what the later steps
// (:170-183) need from them are only
//     n_c = |list_c|,   s_c = sum of the features of list_c,   G = sum_c sum_{x in list_c} x x^T  (= sum_i w_i x_i x_i^T,
//     w_i = number of classes sample i is labelled with)
// which stream: each batch adds to fp64 accumulators that live in HBM (and all-reduce across ranks); the finish (class
// means, centred scatter, pseudo-inverse) is O(C*D + D^3) host work once per pass (rgqa_b200/tasks/gqa_maha.py).
#include "common.cuh"
#include "kernels.h"

namespace rgqa {
namespace {

// one CTA per sample: w_i, class sums and counts (a sample carries one or a few labels out of 1842)
__global__ void __launch_bounds__(256)
estimator_labels_kernel(const float* __restrict__ x, const float* __restrict__ target, long long ld_t, int B, int C, int D,
                        double* __restrict__ class_sum, double* __restrict__ class_count, float* __restrict__ w_out) {
  const int i = blockIdx.x;
  if (i >= B) return;
  __shared__ int hits[64];
  __shared__ int nhit;
  if (threadIdx.x == 0) nhit = 0;
  __syncthreads();
  int mine = 0;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    if (target[static_cast<long long>(i) * ld_t + c] != 0.f) {
      ++mine;
      const int slot = atomicAdd(&nhit, 1);
      if (slot < 64) hits[slot] = c;
    }
  }
  __syncthreads();
  const int n = nhit;
  if (threadIdx.x == 0) w_out[i] = static_cast<float>(n);
  (void)mine;
  if (n <= 64) {
    for (int h = 0; h < n; ++h) {
      const int c = hits[h];
      for (int d = threadIdx.x; d < D; d += blockDim.x)
        atomicAdd(class_sum + static_cast<long long>(c) * D + d, static_cast<double>(x[static_cast<long long>(i) * D + d]));
      if (threadIdx.x == 0) atomicAdd(class_count + c, 1.0);
    }
  } else {  // densely labelled sample: walk the row again
    for (int c = 0; c < C; ++c) {
      if (target[static_cast<long long>(i) * ld_t + c] == 0.f) continue;
      for (int d = threadIdx.x; d < D; d += blockDim.x)
        atomicAdd(class_sum + static_cast<long long>(c) * D + d, static_cast<double>(x[static_cast<long long>(i) * D + d]));
      if (threadIdx.x == 0) atomicAdd(class_count + c, 1.0);
    }
  }
}

// G[a, b] += sum_i w_i x[i, a] x[i, b] in fp64: one CTA per 64 x 64 tile of G (exactly one writer per element), 16 x 16
// threads with a 4 x 4 register block, samples staged 32 at a time
__global__ void __launch_bounds__(256)
estimator_scatter_kernel(const float* __restrict__ x, const float* __restrict__ w, int B, int D, double* __restrict__ G) {
  __shared__ float xa[32][65];
  __shared__ float xb[32][65];
  const int a0 = blockIdx.y * 64, b0 = blockIdx.x * 64;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  double acc[4][4];
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[r][c] = 0.0;
  for (int i0 = 0; i0 < B; i0 += 32) {
    for (int e = threadIdx.x; e < 32 * 64; e += 256) {
      const int s = e >> 6, d = e & 63;
      const int i = i0 + s;
      const bool ok = i < B;
      const float wi = ok ? w[i] : 0.f;
      xa[s][d] = (ok && a0 + d < D) ? wi * x[static_cast<long long>(i) * D + a0 + d] : 0.f;
      xb[s][d] = (ok && b0 + d < D) ? x[static_cast<long long>(i) * D + b0 + d] : 0.f;
    }
    __syncthreads();
#pragma unroll 4
    for (int s = 0; s < 32; ++s) {
      double va[4], vb[4];
#pragma unroll
      for (int r = 0; r < 4; ++r) va[r] = static_cast<double>(xa[s][ty * 4 + r]);
#pragma unroll
      for (int c = 0; c < 4; ++c) vb[c] = static_cast<double>(xb[s][tx * 4 + c]);
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[r][c] = fma(va[r], vb[c], acc[r][c]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int a = a0 + ty * 4 + r, b = b0 + tx * 4 + c;
      if (a < D && b < D) G[static_cast<long long>(a) * D + b] += acc[r][c];
    }
}

}  // namespace

int estimator_accumulate(cudaStream_t stream, const float* x, const float* target, long long ld_t, int B, int C, int D,
                         double* G, double* class_sum, double* class_count, float* w_scratch) {
  RGQA_CHECK(x && target && G && class_sum && class_count && w_scratch, "estimator_accumulate: null argument");
  RGQA_CHECK(B > 0 && C > 0 && D > 0 && ld_t >= C, "estimator_accumulate: bad shape B=%d C=%d D=%d", B, C, D);
  estimator_labels_kernel<<<B, 256, 0, stream>>>(x, target, ld_t, B, C, D, class_sum, class_count, w_scratch);
  note_launch();
  const int t = (D + 63) / 64;
  estimator_scatter_kernel<<<dim3(t, t), 256, 0, stream>>>(x, w_scratch, B, D, G);
  note_launch();
  RGQA_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace rgqa
