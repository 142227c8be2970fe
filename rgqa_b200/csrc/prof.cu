// rgqa_b200 — measurement hooks: a launch counter (every kernel this library enqueues) and optional
// CUDA-event timing around each GEMM launch on its own stream (bench.py's live roofline numbers).
// Off by default; when off the cost is one branch per launch.
#include <vector>

#include "common.cuh"
#include "kernels.h"

namespace rgqa {

namespace {
struct Span {
  cudaEvent_t a, b;
  double flops;
  int kind;
};
bool g_on = false;
long long g_launches = 0;
std::vector<Span> g_spans;
std::vector<cudaEvent_t> g_pool;

cudaEvent_t get_event() {
  if (!g_pool.empty()) {
    cudaEvent_t e = g_pool.back();
    g_pool.pop_back();
    return e;
  }
  cudaEvent_t e = nullptr;
  cudaEventCreate(&e);
  return e;
}
}  // namespace

namespace {
int g_options[OPT_COUNT] = {/*OPT_GEMM_PAIR*/ 1, /*OPT_GEMM_STAGED*/ 1, /*OPT_GEMM_DEBUG*/ 0, /*OPT_PDL (no measurable gain on B200: off)*/ 0, /*OPT_GEMM_LN (slower than GEMM + LN kernel)*/ 0, /*OPT_TWO_STREAMS*/ 1, /*OPT_GRAPHS*/ 1, /*OPT_DEFER_LN*/ 1, /*OPT_ATTN_TMA*/ 1, /*OPT_QKV_PLANES*/ 0, /*OPT_MAHA_TC*/ 1, /*OPT_VISN_FUSED*/ 1, /*OPT_ATTN_FUSED*/ 1, /*OPT_STATE_F16*/ 1, /*OPT_GELU_EW16*/ 0, /*OPT_ATTN_STRADDLE (measured: no gain, the kernel is bound by its attention tasks) */ 0};
}
bool pdl_enabled() { return g_options[OPT_PDL] != 0; }
int option_get(int key) { return (key >= 0 && key < OPT_COUNT) ? g_options[key] : 0; }
void option_set(int key, int value) {
  if (key >= 0 && key < OPT_COUNT) g_options[key] = value;
}

void note_launch(int n) { g_launches += n; }
long long launch_count() { return g_launches; }
void prof_enable(bool on) { g_on = on; }
bool prof_on() { return g_on; }

void* prof_begin(cudaStream_t st) {
  if (!g_on) return nullptr;
  Span s;
  s.a = get_event();
  s.b = get_event();
  s.flops = 0;
  s.kind = 0;
  cudaEventRecord(s.a, st);
  g_spans.push_back(s);
  return reinterpret_cast<void*>(g_spans.size());  // 1-based index
}
void prof_end(void* tok, cudaStream_t st, double flops, int kind) {
  if (tok == nullptr) return;
  Span& s = g_spans[reinterpret_cast<size_t>(tok) - 1];
  s.flops = flops;
  s.kind = kind;
  cudaEventRecord(s.b, st);
}

// Sums per kind (0 = tcgen05 GEMM, 1 = fp32 GEMM, 2 = attention, 3 = row/other); events must have completed.
int prof_collect(double* ms_sum, double* flops_sum, long long* count, int nkinds) {
  for (int k = 0; k < nkinds; ++k) {
    ms_sum[k] = 0;
    flops_sum[k] = 0;
    count[k] = 0;
  }
  for (Span& s : g_spans) {
    float ms = 0.f;
    RGQA_CUDA_OK(cudaEventSynchronize(s.b));
    RGQA_CUDA_OK(cudaEventElapsedTime(&ms, s.a, s.b));
    if (s.kind >= 0 && s.kind < nkinds) {
      ms_sum[s.kind] += ms;
      flops_sum[s.kind] += s.flops;
      count[s.kind] += 1;
    }
    g_pool.push_back(s.a);
    g_pool.push_back(s.b);
  }
  g_spans.clear();
  return 0;
}

}  // namespace rgqa
