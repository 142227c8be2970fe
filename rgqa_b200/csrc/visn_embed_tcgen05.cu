// rgqa_b200 — VisualFeatEncoder as ONE tcgen05 kernel (BASELINE north_star item 1).
//
//   x   = feat[M,2048] · W1[768,2048]ᵀ + b1          visn_fc            src/lxrt/modeling.py:507-517
//   y   = box[M,4]     · W2[768,4]ᵀ    + b2          box_fc
//   out = ( LN_1(x) + LN_2(y) ) / 2                  visn_layer_norm, box_layer_norm, (x + y) / 2
//   written as the identity deferred-LayerNorm state of the visual stream: hi = bf16(out), lo = bf16(out - hi),
//   statistics (0, 768) so that the next consumer sees mean 0 / rstd 1.
//
// Before: cast fp32 -> bf16 (302 MB read + 151 MB written at B = 1024), a GEMM writing the fp32 pre-LayerNorm rows
// (113 MB) and a row kernel reading them back: three launches and the only place the 295 KB / question input was touched
// twice.  Here the 768-wide row never leaves the SM between the projection and the two LayerNorms:
//
// * LayerNorm needs whole rows; 128 rows x 768 fp32 accumulators would be 768 TMEM columns (an SM has 512).  A CTA PAIR
//   issuing cta_group::2 MMAs with M = 128 gives each CTA 64 rows, and the hardware packs a 64 x N accumulator into N/2
//   columns x 128 lanes (column half 0 in lanes 0-63, half 1 in lanes 64-127): three N = 256 MMAs per K step hold a complete
//   64 x 768 row block in 384 TMEM columns.
// * fp32 features (the reference loader's format, src/tasks/gqa_data.py:173-238) are consumed AS THEY ARE by kind::tf32
//   MMAs (10 mantissa bits, more than the bf16 cast kept): no cast kernel, no bf16 copy.  bf16 features (the ingest store's
//   format) take kind::f16 at twice the rate.
// * box_fc has K = 4: four FMAs per output element in the epilogue.  Its LayerNorm statistics are closed forms of the box:
//   mean_n y = box · colmean(W2) + mean(b2) and var_n y = vᵀ C v with v = (box, 1) and C the (K+1) x (K+1) second-moment
//   matrix of the CENTRED rows (w2_n - colmean, b2_n - mean b2), computed once on the host in fp64 — no second reduction.
//
//   warp 0      TMA producer: per k-block its own 64 rows of A (8 KB) + its half (128 rows) of each of the three W tiles
//               (48 KB), 3-stage ring; completion bytes credited to the leader's barrier
//   warp 1      MMA issuer (leader CTA): 3 x 4 tcgen05.mma.cta_group::2 (128 x 256 x 16 bf16 | x 8 tf32) per k-block
//   warp 2      TMEM allocator (512 columns)
//   warps 4-11  epilogue, one TMEM lane (= one row, one column half) per thread, two warps per lane quarter (even / odd
//               32-column chunks); per-column parameters sit in shared memory (36 B per column, staged once per CTA):
//               pass 1  row sum of acc + b1          pass 1b  centred sum of squares (two-pass variance, as torch)
//               pass 2  both LayerNorms, average, hi/lo split; each warp transposes its 32 rows x 32 columns through a
//                       swizzled 2 KB buffer so that global stores are row-contiguous
// The accumulator is single-buffered (384 of 512 columns): the next tile's MMAs wait for pass 2, its TMA loads do not.
#include <cuda.h>

#include "common.cuh"
#include "kernels.h"

namespace rgqa {
namespace {

constexpr int N_LN = 768;
constexpr int kNT = 3;            // 256-wide W tiles
constexpr int kRowBytes = 128;    // one swizzle-atom row: 64 bf16 or 32 fp32 (tf32) along K
constexpr int kRowsPerCta = 64;   // cta_group::2, UMMA M = 128
constexpr int kABytes = kRowsPerCta * kRowBytes;          // 8 KB
constexpr int kBTileBytes = 128 * kRowBytes;              // this CTA's half (128 rows) of one 256-row W tile: 16 KB
constexpr int kStageBytes = kABytes + kNT * kBTileBytes;  // 56 KB
constexpr int kStages = 3;
constexpr int kEpiWarps = 8;         // two per TMEM lane quarter: even / odd 32-column chunks
constexpr int kThreads = (4 + kEpiWarps) * 32;
constexpr int kBufBytes = 32 * 64;   // staging: 32 rows x 64 B (32 bf16), 16-byte pieces XOR-swizzled by row (64-byte swizzle)
// per-output-column parameters staged once per CTA: {b1, g1, be1 + be2, g2}, W2[n][0..3] * g2[n], b2[n] * g2[n]
constexpr int kParamBytes = N_LN * (16 + 16 + 4);
constexpr int kSmemBytes = kStages * kStageBytes + kEpiWarps * kBufBytes + kParamBytes + 256 * 4 + 1024 + 512;

struct VisnParams {
  int M, K;           // K in elements of the feature type
  const float* b1;    // [768] visn_fc bias
  const float* g1;    // visn_layer_norm
  const float* be1;
  const float* boxes;  // [M, box_dim]
  int box_dim;
  const float* W2p;   // [768][8] box_fc weight, zero padded (columns >= 4 unused: box_dim <= 4)
  const float* b2;    // [768]
  const float* g2;    // box_layer_norm
  const float* be2;
  const float* bstat;  // [8] colmean(W2) | [1] mean(b2) | pad to 16 | [81] C row-major (9 x 9)
  float eps;
  __nv_bfloat16* hi;
  __nv_bfloat16* lo;
  float* stats;       // [M][12]
};

__device__ __forceinline__ void named_bar(int id, int threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}
// global column of chunk c (0..11) for a thread in column half `half`: tile nt = c / 4, 32-col piece j = c % 4
__device__ __forceinline__ int chunk_col(int c, int half) { return (c >> 2) * 256 + half * 128 + (c & 3) * 32; }
// TMEM column of the same chunk (each 256-wide tile occupies 128 columns)
__device__ __forceinline__ uint32_t chunk_tcol(int c) { return static_cast<uint32_t>((c >> 2) * 128 + (c & 3) * 32); }

// tf32 x tf32 -> fp32 (fp32 words in shared memory read as tf32), both operands K-major
__host__ __device__ constexpr uint32_t umma_idesc_tf32(uint32_t M, uint32_t N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((N >> 3) << 17) | ((M >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

template <bool F32IN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
visn_embed_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const VisnParams p) {
  constexpr int BK = F32IN ? 32 : 64;  // elements per 128-byte row
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* stage_bufs = smem + kStages * kStageBytes;  // [kEpiWarps][2 KB], 1024-aligned (stage size is 56 KB)
  float4* pA = reinterpret_cast<float4*>(stage_bufs + kEpiWarps * kBufBytes);  // [768] {b1, g1, be1 + be2, g2}
  float4* pW = pA + N_LN;                                                       // [768] W2[n][0..3] * g2[n]
  float* pB = reinterpret_cast<float*>(pW + N_LN);                              // [768] b2[n] * g2[n]
  float* xch = pB + N_LN;                                                       // [2][128] per-lane partial sums
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(xch + 256);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tmem_full_bar = empty_bar + kStages;
  uint64_t* tmem_empty_bar = tmem_full_bar + 1;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_empty_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1;
  const int num_pairs = gridDim.x >> 1;
  const int num_tiles = (p.M + 2 * kRowsPerCta - 1) / (2 * kRowsPerCta);
  const int num_kb = p.K / BK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tmem_full_bar, 1);
    mbar_init(tmem_empty_bar, 2 * kEpiWarps);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc_pair<512>(tmem_ptr_smem);
  // parameters are weights (never written by a preceding kernel): staged before the dependency wait
  for (int n = threadIdx.x; n < N_LN; n += kThreads) {
    const float g2 = __ldg(p.g2 + n);
    const float4 w = __ldg(reinterpret_cast<const float4*>(p.W2p + n * 8));
    pA[n] = make_float4(__ldg(p.b1 + n), __ldg(p.g1 + n), __ldg(p.be1 + n) + __ldg(p.be2 + n), g2);
    pW[n] = make_float4(w.x * g2, w.y * g2, w.z * g2, w.w * g2);
    pB[n] = __ldg(p.b2 + n) * g2;
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  pdl_launch_dependents();
  pdl_wait();

  if (warp == 0) {
    // ------------------------------ TMA producer (both CTAs) ------------------------------
    int stage = 0;
    uint32_t phase = 0;
    for (int t = pair; t < num_tiles; t += num_pairs) {
      const int m0 = t * (2 * kRowsPerCta) + static_cast<int>(rank) * kRowsPerCta;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1u);
        if (lane == 0) {
          uint8_t* sa = smem + stage * kStageBytes;
          const uint32_t leader_full = mapa_u32(smem_u32(&full_bar[stage]), 0);
          if (rank == 0) mbar_expect_tx(&full_bar[stage], 2 * kStageBytes);
          tma_load_2d_pair(sa, &tmA, leader_full, kb * BK, m0);
#pragma unroll
          for (int nt = 0; nt < kNT; ++nt)
            tma_load_2d_pair(sa + kABytes + nt * kBTileBytes, &tmB, leader_full, kb * BK, nt * 256 + static_cast<int>(rank) * 128);
        }
        __syncwarp();
        if (++stage == kStages) {
          stage = 0;
          phase ^= 1u;
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------ MMA issuer (leader CTA only) ------------------------------
    if (rank == 0) {
      constexpr uint32_t idesc = F32IN ? umma_idesc_tf32(2 * kRowsPerCta, 256) : umma_idesc_bf16(2 * kRowsPerCta, 256);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int t = pair; t < num_tiles; t += num_pairs, ++it) {
        mbar_wait(tmem_empty_bar, (it & 1u) ^ 1u);  // the previous tile's epilogue has drained the accumulators
        tc_fence_after();
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          if (lane == 0) {
            const uint32_t sa = smem_u32(smem + stage * kStageBytes);
            const uint64_t adesc = umma_smem_desc_sw128(sa);
#pragma unroll
            for (int nt = 0; nt < kNT; ++nt) {
              const uint64_t bdesc = umma_smem_desc_sw128(sa + kABytes + nt * kBTileBytes);
#pragma unroll
              for (int k = 0; k < 4; ++k) {  // 32 bytes of K per MMA: 16 bf16 or 8 tf32
                if constexpr (F32IN)
                  umma_tf32_pair(tmem_base + static_cast<uint32_t>(nt * 128), adesc + 2u * k, bdesc + 2u * k, idesc, (kb | k) != 0 ? 1u : 0u);
                else
                  umma_bf16_pair(tmem_base + static_cast<uint32_t>(nt * 128), adesc + 2u * k, bdesc + 2u * k, idesc, (kb | k) != 0 ? 1u : 0u);
              }
            }
            umma_commit_pair(&empty_bar[stage]);
            if (kb == num_kb - 1) umma_commit_pair(tmem_full_bar);
          }
          __syncwarp();
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else if (warp >= 4) {
    // ------------------------------ epilogue (both CTAs): one TMEM lane per thread, two warps per lane quarter ------------------------------
    const int e = warp - 4;
    const int q = e & 3;                    // TMEM lane quarter (== warp % 4)
    const int grp = e >> 2;                 // chunk parity handled by this warp: chunks c with (c & 1) == grp
    const int tl = q * 32 + lane;           // TMEM lane of this thread
    const int half = tl >> 6;               // column half held by this lane
    uint8_t* buf = stage_bufs + e * kBufBytes;
    const uint32_t leader_tmem_empty = mapa_u32(smem_u32(tmem_empty_bar), 0);
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    // coalesced side of the staging transpose: lane l handles row (l >> 2) + 8*i, 16-byte piece (l & 3)
    const int crow = lane >> 2, cpiece = lane & 3;
    const int sw = (lane >> 1) & 3;  // 64-byte swizzle key of this thread's own row (row index within the 32-row buffer = lane)
    const int KB = p.box_dim;
    int it = 0;
    for (int t = pair; t < num_tiles; t += num_pairs, ++it) {
      const int row_base = t * (2 * kRowsPerCta) + static_cast<int>(rank) * kRowsPerCta + (q & 1) * 32;  // first row of this warp
      const int row = row_base + lane;
      // the box branch of this row while the main loop still runs: y_n = box · W2[n] + b2[n]; its LayerNorm statistics are
      // closed forms of the box (file header)
      float bx[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) bx[j] = (j < KB && row < p.M) ? __ldg(p.boxes + static_cast<long long>(row) * KB + j) : 0.f;
      float ymean = __ldg(p.bstat + 8);
#pragma unroll
      for (int j = 0; j < 4; ++j) ymean = fmaf(bx[j], __ldg(p.bstat + j), ymean);
      float yvar = 0.f;
      {
        const float* Cm = p.bstat + 16;
        const float vv[5] = {bx[0], bx[1], bx[2], bx[3], 1.0f};
        const int idx[5] = {0, 1, 2, 3, 8};
#pragma unroll
        for (int a = 0; a < 5; ++a) {
          float acc = 0.f;
#pragma unroll
          for (int b = 0; b < 5; ++b) acc = fmaf(__ldg(Cm + idx[a] * 9 + idx[b]), vv[b], acc);
          yvar = fmaf(vv[a], acc, yvar);
        }
      }
      const float yrstd = 1.0f / sqrtf(fmaxf(yvar, 0.f) + p.eps);

      mbar_wait(tmem_full_bar, it & 1u);
      tc_fence_after();

      // ---------------- pass 1: row sum of x = acc + b1 (this warp's six chunks) ----------------
      float sum = 0.f;
#pragma unroll 1
      for (int c = grp; c < 12; c += 4) {
        uint32_t v0[32], v1[32];
        tmem_ld32(t_lane + chunk_tcol(c), v0);
        tmem_ld32(t_lane + chunk_tcol(c + 2), v1);
        const float4* a0 = pA + chunk_col(c, half);
        const float4* a1 = pA + chunk_col(c + 2, half);
        tmem_ld_wait_dep(v0);
        tmem_ld_wait_dep(v1);
        float s0 = 0.f, s1 = 0.f;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          s0 += __uint_as_float(v0[i]) + a0[i].x;
          s1 += __uint_as_float(v1[i]) + a1[i].x;
        }
        sum += s0 + s1;
      }
      // row mean: the row's other chunk parity lives in the partner warp, its other column half in lane tl ^ 64
      xch[grp * 128 + tl] = sum;
      named_bar(1, kEpiWarps * 32);
      const float mean = ((xch[tl] + xch[128 + tl]) + (xch[tl ^ 64] + xch[128 + (tl ^ 64)])) * (1.0f / N_LN);
      named_bar(1, kEpiWarps * 32);

      // ---------------- pass 1b: centred sum of squares ----------------
      float sq = 0.f;
#pragma unroll 1
      for (int c = grp; c < 12; c += 4) {
        uint32_t v0[32], v1[32];
        tmem_ld32(t_lane + chunk_tcol(c), v0);
        tmem_ld32(t_lane + chunk_tcol(c + 2), v1);
        const float4* a0 = pA + chunk_col(c, half);
        const float4* a1 = pA + chunk_col(c + 2, half);
        tmem_ld_wait_dep(v0);
        tmem_ld_wait_dep(v1);
        float s0 = 0.f, s1 = 0.f;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const float d0 = (__uint_as_float(v0[i]) + a0[i].x) - mean, d1 = (__uint_as_float(v1[i]) + a1[i].x) - mean;
          s0 = fmaf(d0, d0, s0);
          s1 = fmaf(d1, d1, s1);
        }
        sq += s0 + s1;
      }
      xch[grp * 128 + tl] = sq;
      named_bar(1, kEpiWarps * 32);
      const float rstd = 1.0f / sqrtf(((xch[tl] + xch[128 + tl]) + (xch[tl ^ 64] + xch[128 + (tl ^ 64)])) * (1.0f / N_LN) + p.eps);
      named_bar(1, kEpiWarps * 32);

      // ---------------- pass 2: both LayerNorms, average, hi/lo split, row-contiguous stores ----------------
      const bool f16_state = p.lo == nullptr;
      const float xa = rstd, xb = -mean * rstd;        // LN_1: (x - mean) * rstd = x * xa + xb
      const float ya = yrstd, yb = -ymean * yrstd;
#pragma unroll 1
      for (int c = grp; c < 12; c += 2) {
        uint32_t v[32];
        tmem_ld32(t_lane + chunk_tcol(c), v);
        const int col0 = chunk_col(c, half);
        tmem_ld_wait_dep(v);
        uint8_t* rowp = buf + lane * 64;
        uint32_t hw[16], lw[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {  // columns col0 + 2i, col0 + 2i + 1 (all lanes of the warp read the same parameters)
          float o[2];
#pragma unroll
          for (int ee = 0; ee < 2; ++ee) {
            const int n = col0 + 2 * i + ee;
            const float4 a = pA[n], w = pW[n];
            const float x = __uint_as_float(v[2 * i + ee]) + a.x;
            const float xn = fmaf(x, xa, xb);
            float yg = fmaf(bx[0], w.x, pB[n]);   // (box · W2[n] + b2[n]) * g2[n]
            yg = fmaf(bx[1], w.y, yg);
            yg = fmaf(bx[2], w.z, yg);
            yg = fmaf(bx[3], w.w, yg);
            const float yn = fmaf(yg, ya, yb * a.w);
            o[ee] = (fmaf(xn, a.y, yn) + a.z) * 0.5f;   // (LN_1(x) g1 + be1 + LN_2(y) g2 + be2) / 2
          }
          if (f16_state) {  // single-word fp16 state (p.lo == nullptr): no remainder
            hw[i] = pack_f16x2_sat(o[0], o[1]);
            lw[i] = 0u;
          } else {
            hw[i] = pack_bf16x2(o[0], o[1]);
            const float2 hf = unpack_bf16x2(hw[i]);
            lw[i] = pack_bf16x2(o[0] - hf.x, o[1] - hf.y);
          }
        }
#pragma unroll
        for (int pass = 0; pass < 2; ++pass) {  // hi rows, then lo rows, through the same 2 KB buffer
          if (pass == 1 && f16_state) break;
          const uint32_t* src = pass == 0 ? hw : lw;
#pragma unroll
          for (int i = 0; i < 4; ++i)
            *reinterpret_cast<uint4*>(rowp + ((i ^ sw) << 4)) = make_uint4(src[4 * i], src[4 * i + 1], src[4 * i + 2], src[4 * i + 3]);
          __syncwarp();
          __nv_bfloat16* dst = (pass == 0 ? p.hi : p.lo) + col0 + cpiece * 8;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int rl = crow + 8 * i;
            const int r = row_base + rl;
            const uint4 y = *reinterpret_cast<const uint4*>(buf + rl * 64 + ((cpiece ^ ((rl >> 1) & 3)) << 4));
            if (r < p.M) *reinterpret_cast<uint4*>(dst + static_cast<long long>(r) * N_LN) = y;
          }
          __syncwarp();  // the buffer is rewritten by the next pass / chunk
        }
      }
      // identity statistics: (sum 0, sumsq 768) -> mean 0, rstd 1 for the next consumer / producer
      if (half == 0 && grp == 0 && row < p.M) {
        float4* so = reinterpret_cast<float4*>(p.stats + static_cast<long long>(row) * 12);
        so[0] = make_float4(0.f, static_cast<float>(N_LN), 0.f, 0.f);
        so[1] = make_float4(0.f, 0.f, 0.f, 0.f);
        so[2] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
      // accumulators fully consumed: let the leader start the next tile's MMAs
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(leader_tmem_empty);
    }
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == 2) tmem_dealloc_pair<512>(tmem_base);
}

}  // namespace

int visn_embed(cudaStream_t stream, const void* feat, int feat_f32, long long ldf, const void* W1, const float* b1, const float* g1,
               const float* be1, const float* boxes, int box_dim, const float* W2p, const float* b2, const float* g2,
               const float* be2, const float* bstat, float eps, void* hi, void* lo, float* stats, int M, int K) {
  RGQA_CHECK(M > 0 && K > 0, "visn_embed: empty problem");
  const int bk = feat_f32 ? 32 : 64;
  RGQA_CHECK(K % bk == 0, "visn_embed: K=%d must be a multiple of %d", K, bk);
  RGQA_CHECK(box_dim >= 1 && box_dim <= 4, "visn_embed: box_dim %d out of range [1,4] (the LXMERT box projection)", box_dim);
  RGQA_CHECK(feat && W1 && b1 && g1 && be1 && boxes && W2p && b2 && g2 && be2 && bstat && hi && stats, "visn_embed: null argument");
  RGQA_CHECK(reinterpret_cast<uintptr_t>(hi) % 16 == 0 && reinterpret_cast<uintptr_t>(lo) % 16 == 0 &&
                 reinterpret_cast<uintptr_t>(stats) % 16 == 0 && reinterpret_cast<uintptr_t>(b1) % 16 == 0 &&
                 reinterpret_cast<uintptr_t>(W2p) % 16 == 0,
             "visn_embed: operands must be 16-byte aligned");
  const int esz = feat_f32 ? 4 : 2;
  CUtensorMap tmA, tmB;
  if (make_tmap(&tmA, feat, M, K, ldf, kRowsPerCta, esz)) return 1;
  if (make_tmap(&tmB, W1, N_LN, K, K, 128, esz)) return 1;
  VisnParams p;
  p.M = M;
  p.K = K;
  p.b1 = b1;
  p.g1 = g1;
  p.be1 = be1;
  p.boxes = boxes;
  p.box_dim = box_dim;
  p.W2p = W2p;
  p.b2 = b2;
  p.g2 = g2;
  p.be2 = be2;
  p.bstat = bstat;
  p.eps = eps;
  p.hi = static_cast<__nv_bfloat16*>(hi);
  p.lo = static_cast<__nv_bfloat16*>(lo);
  p.stats = stats;
  const int tiles = (M + 2 * kRowsPerCta - 1) / (2 * kRowsPerCta);
  const int max_pairs = launch_sms() / 2;
  const int pairs = tiles < max_pairs ? tiles : max_pairs;
  void* tok = prof_begin(stream);
  if (feat_f32) {
    static bool configured = false;
    if (!configured) {
      RGQA_CUDA_OK(cudaFuncSetAttribute(visn_embed_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
      configured = true;
    }
    RGQA_CUDA_OK(launch_pdl(visn_embed_kernel<true>, dim3(2 * pairs), dim3(kThreads), kSmemBytes, stream, tmA, tmB, p));
  } else {
    static bool configured = false;
    if (!configured) {
      RGQA_CUDA_OK(cudaFuncSetAttribute(visn_embed_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
      configured = true;
    }
    RGQA_CUDA_OK(launch_pdl(visn_embed_kernel<false>, dim3(2 * pairs), dim3(kThreads), kSmemBytes, stream, tmA, tmB, p));
  }
  prof_end(tok, stream, 2.0 * M * static_cast<double>(N_LN) * K, PROF_GEMM_TC);
  note_launch();
  RGQA_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace rgqa
