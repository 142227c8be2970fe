// rgqa_b200 — self-attention INSIDE the Q|K|V projection (BertAttention.forward, src/lxrt/modeling.py:320-347, for the blocks
// whose queries and keys are the same rows: BertSelfattLayer :376-386 of the language / object-relational / cross-modality
// layers, and every UNITER layer).
//
//   ctx[b, i, h*64:(h+1)*64] = softmax_j( q_i·k_j / 8 + mask[b,j] ) · v_j ,   [q|k|v] = LN(x)·W_{q|k|v}ᵀ + b
//
// Separate kernels wrote the [rows, 2304] projection to HBM (170 MB per visual layer at B = 1024) and read it straight back:
// the attention kernel ran at 0.59 of the copy bandwidth on nothing but that round trip (profiles/README.md).  Here the
// projection tile is consumed where it is produced:
//
// * The weight is packed HEAD-MAJOR ([head][q|k|v][64] rows), so one 192-column tile is exactly one head's Q, K and V.
// * Row tiles are cut at SAMPLE boundaries: a CTA's 128 accumulator rows hold ns = floor(128 / S) whole samples (6 x 20
//   language rows, 3 x 36 visual rows, 2 x 56 UNITER rows); the TMA box still fetches 128 rows, the few rows behind the last
//   whole sample are computed and ignored.  A (sample, head) attention unit therefore never straddles CTAs.
// * CTA pair (cta_group::2): one 256 x 192 x 16 MMA per step, accumulators double-buffered in TMEM (2 x 192 of 512 columns),
//   so the whole attention of tile i runs under the main loop of tile i+1.
// * Epilogue = 12 epilogue warps: (1) TMEM -> rstd * acc + bias' (deferred LayerNorm) -> bf16 -> three
//   128B-swizzled [128 x 64] matrices Q, K, V in shared memory; (2) the (sample, 16-query tile) tasks of the tile are dealt
//   over the warps: mma.sync m16n8k16 QKᵀ, quad-shuffle softmax in fp32, P·V, exactly the arithmetic of attention.cu;
//   (3) the context tile overwrites the consumed Q lines and leaves through ONE TMA store.
// HBM traffic of a self-attention block's front half: read x (1 536 B / row) + write ctx (1 536 B / row), instead of
// + 4 608 written + 4 608 read.
#include <cuda.h>

#include "common.cuh"
#include "kernels.h"

namespace rgqa {
namespace {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int HD = 64;
constexpr int kTileN = 192;                                  // one head: q | k | v
constexpr int kStageBytes = BM * BK * 2 + (kTileN / 2) * BK * 2;  // 16 KB of A + this CTA's 96 rows of W (12 KB)
constexpr int kStages = 4;
constexpr int kEpiWarps = 12;  // 3 per scheduler: the attention tasks of a tile (9 visual, 12 language) run in ONE round
constexpr int kThreads = (4 + kEpiWarps) * 32;
constexpr int kMatBytes = BM * 128;                          // one [128 x 64] bf16 matrix
constexpr int kPadBytes = 2048;                              // zeroed: V fragment rows up to 15 lines past the tile land here
constexpr int kBufBytes = 3 * kMatBytes + kPadBytes;         // Q | K | V | pad
constexpr int kAccStride = 256;                              // TMEM columns between the two accumulator buffers
constexpr int kSmemBytes = kStages * kStageBytes + 2 * kBufBytes + 1024 + 256;
// STRADDLE form (S = 36, the visual stream): a CTA PAIR owns SEVEN whole samples (252 of its 256 accumulator rows) instead of
// 2 x 3 (216): CTA 0 holds pair rows 0..123 (samples 0-2 and the first 16-query tile of sample 3), CTA 1 pair rows 124..251 (the
// rest of sample 3 and samples 4-6).  The K and V lines of the straddling sample cross the pair through distributed shared
// memory: each CTA's K / V matrices are a 160-line WINDOW of the pair tile (CTA 0: pair rows 0..143, CTA 1: pair rows 108..251;
// lines 144..159 stay zero for the fragment overruns), the 36 foreign lines arrive as st.async stores that complete on the
// receiver's mbarrier.  14 % fewer tiles for the same samples.
constexpr int kStrS = 36;
constexpr int kStrSamples = 7;                               // samples per pair tile
constexpr int kStrRows0 = 3 * kStrS + 16;                    // 124 rows in CTA 0
constexpr int kStrRowsPair = kStrSamples * kStrS;            // 252
constexpr int kStrBase1 = 3 * kStrS;                         // first pair row of CTA 1's K / V window (108)
constexpr int kStrKvLines = 160;
constexpr int kStrKvBytes = kStrKvLines * 128;
constexpr int kStrBufBytes = kMatBytes + 2 * kStrKvBytes;    // Q | K | V
constexpr int kStrHalo0 = (kStrS - 16) * 2 * 128;            // bytes CTA 0 receives per tile: K and V lines of sample rows 16..35
constexpr int kStrHalo1 = 16 * 2 * 128;                      // bytes CTA 1 receives: sample rows 0..15
constexpr int kSmemBytesStr = kStages * kStageBytes + 2 * kStrBufBytes + 1024 + 256;
static_assert(kSmemBytesStr <= 232448, "straddle form exceeds the shared memory of an SM");

constexpr float kLog2e = 1.44269504088896340736f;
constexpr float kScaleLog2e = 0.125f * kLog2e;

struct QkvAttnParams {
  int B, S, ns, heads;     // samples, rows per sample, whole samples per CTA tile, heads
  int M;                   // B * S
  const float* bias;       // [heads * 192] head-major (folded with the LayerNorm's beta when ln_stats != nullptr)
  const float* ln_stats;   // [M][6][2] or nullptr (A rows are already normalised)
  float ln_eps;
  const int32_t* key_mask;  // [B, S] (1 = keep) or nullptr
  int a_f16;                // A rows are fp16 (the single-word deferred-LayerNorm state) and so is W
  int debug;                // measurement only (rgqa_set_option "gemm_debug"): bit0 no attention tasks, bit1 no projection stores, bit2 no context store
};

__device__ __forceinline__ void named_bar_sync(int id, int threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}
__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// shared-memory address of 16-byte chunk `chunk` (0..7) of the 128-byte line at `line_addr` under the 128-byte swizzle
__device__ __forceinline__ uint32_t sw128_addr(uint32_t line_addr, int chunk) {
  return line_addr + ((static_cast<uint32_t>(chunk) ^ ((line_addr >> 7) & 7u)) << 4);
}
__device__ __forceinline__ float ex2_fast(float x) {  // same expression as attention.cu (bit-identical paths)
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_trans(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
// 16-byte store into the PEER CTA's shared memory that completes (release, cluster scope) on the peer's mbarrier
__device__ __forceinline__ void st_async_v4(uint32_t cluster_addr, const uint4& v, uint32_t cluster_bar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(cluster_addr),
               "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "r"(cluster_bar)
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t"
      "}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {  // bounded like mbar_wait: a protocol bug traps
  uint32_t spins = 0;
  unsigned long long t0 = 0;
  while (!mbar_try_wait_cluster(bar, parity)) {
    if ((++spins & 1023u) == 0) {
      unsigned long long now;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
      if (t0 == 0) t0 = now;
      if (now - t0 > 2000000000ull) {
        printf("rgqa: cluster mbarrier wait timed out (block %d thread %d)\n", blockIdx.x, threadIdx.x);
        __trap();
      }
    }
  }
}
__device__ __forceinline__ float row_rstd(const float* __restrict__ stats, int row, float eps) {
  const float4* s4 = reinterpret_cast<const float4*>(stats + static_cast<long long>(row) * 12);
  const float4 a = __ldg(s4), b = __ldg(s4 + 1), c = __ldg(s4 + 2);
  const float sum = ((a.x + a.z) + (b.x + b.z)) + (c.x + c.z);
  const float sq = ((a.y + a.w) + (b.y + b.w)) + (c.y + c.w);
  const float mu = sum * (1.0f / 768.0f);
  const float var = fmaxf(sq * (1.0f / 768.0f) - mu * mu, 0.f);
  return 1.0f / sqrtf(var + eps);
}

// KS = ceil(S / 16): 16-key steps of P·V; NT = ceil(S / 8) <= 2 KS: 8-key score tiles of QKᵀ (a tile wholly beyond the last
// key has probability exactly 0 everywhere — it is not computed at all: S = 36 runs 5 score tiles, S = 20 three)
template <int KS, int NT, bool STR>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
qkv_attn_pair_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                     const __grid_constant__ CUtensorMap tmO, const __grid_constant__ CUtensorMap tmO2, const QkvAttnParams p) {
  static_assert(NT == 2 * KS || NT == 2 * KS - 1, "NT = ceil(S / 8), KS = ceil(S / 16)");
  static_assert(!STR || (KS == 3 && NT == 5), "the straddle form is built for S = 36");
  constexpr int kBufB = STR ? kStrBufBytes : kBufBytes;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* bufs = smem + kStages * kStageBytes;  // 1024-aligned: the stage size is 28 KB
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(bufs + 2 * kBufB);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tmem_full_bar = empty_bar + kStages;
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;
  uint64_t* halo_full = tmem_empty_bar + 2;   // [2] STR: the peer's K / V lines of buffer b have landed (tx bytes)
  uint64_t* halo_free = halo_full + 2;        // [2] STR: the PEER is done reading the lines this CTA sent into its buffer b
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(halo_free + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1;
  const int num_pairs = gridDim.x >> 1;
  const int rows_cta = p.ns * p.S;                         // rows of whole samples in a CTA tile
  const int row_tiles = STR ? (p.B + kStrSamples - 1) / kStrSamples : (p.B + 2 * p.ns - 1) / (2 * p.ns);  // pair tiles along the rows
  const int num_tiles = row_tiles * p.heads;               // head fastest: the 12 tiles of a row tile reuse its A rows from L2
  constexpr int num_kb = 768 / BK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmO);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full_bar[a], 1);
      mbar_init(&tmem_empty_bar[a], 2 * kEpiWarps);
      mbar_init(&halo_full[a], 1);   // the local arrive.expect_tx of each tile; the bytes come from the peer
      mbar_init(&halo_free[a], 1);   // one remote arrive per tile (the peer's issuer thread)
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc_pair<512>(tmem_ptr_smem);
  if (warp >= 4) {  // the overrun pads must hold zeros (0 x NaN would poison P·V); nothing ever writes them
    if constexpr (STR) {  // lines 144..159 of K and V in both buffers: 2 x 2 x 2 KB
      for (int i = threadIdx.x - 128; i < 4 * 128; i += kEpiWarps * 32) {
        const int m = i >> 7, o = i & 127;
        reinterpret_cast<uint4*>(bufs + (m >> 1) * kStrBufBytes + kMatBytes + (m & 1) * kStrKvBytes + 144 * 128)[o] = make_uint4(0u, 0u, 0u, 0u);
      }
    } else {
      for (int i = threadIdx.x - 128; i < 2 * (kPadBytes / 16); i += kEpiWarps * 32) {  // 256 pieces, 384 threads
        const int b = i / (kPadBytes / 16), o = i % (kPadBytes / 16);
        reinterpret_cast<uint4*>(bufs + b * kBufBytes + 3 * kMatBytes)[o] = make_uint4(0u, 0u, 0u, 0u);
      }
    }
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
      const int rt = t / p.heads, head = t - rt * p.heads;
      const int m0 = STR ? rt * kStrRowsPair + static_cast<int>(rank) * kStrRows0
                         : (2 * rt + static_cast<int>(rank)) * rows_cta;     // first row of this CTA's samples
      const int n0 = head * kTileN + static_cast<int>(rank) * (kTileN / 2);
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1u);
        if (lane == 0) {
          uint8_t* sa = smem + stage * kStageBytes;
          uint8_t* sb = sa + BM * BK * 2;
          const uint32_t leader_full = mapa_u32(smem_u32(&full_bar[stage]), 0);
          if (rank == 0) mbar_expect_tx(&full_bar[stage], 2 * kStageBytes);
          tma_load_2d_pair(sa, &tmA, leader_full, kb * BK, m0);  // rows >= M arrive as zeros
          tma_load_2d_pair(sb, &tmB, leader_full, kb * BK, n0);
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
      const uint32_t idesc = umma_idesc_bf16(2 * BM, kTileN) & ~(p.a_f16 ? kIdescABf16 : 0u);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int t = pair; t < num_tiles; t += num_pairs, ++it) {
        const int acc = it & 1;
        mbar_wait(&tmem_empty_bar[acc], ((it >> 1) & 1u) ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * kAccStride);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          if (lane == 0) {
            const uint32_t sa = smem_u32(smem + stage * kStageBytes);
            const uint64_t adesc = umma_smem_desc_sw128(sa);
            const uint64_t bdesc = umma_smem_desc_sw128(sa + BM * BK * 2);
#pragma unroll
            for (int k = 0; k < BK / 16; ++k)
              umma_bf16_pair(d_tmem, adesc + 2u * k, bdesc + 2u * k, idesc, (kb | k) != 0 ? 1u : 0u);
            umma_commit_pair(&empty_bar[stage]);
            if (kb == num_kb - 1) umma_commit_pair(&tmem_full_bar[acc]);
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
    // ------------------------------ epilogue: projection tile -> shared memory -> attention -> context tile ------------------------------
    const int e = warp - 4;
    const int quarter = e & 3;   // TMEM lane quarter == rows 32 quarter .. + 31 of the CTA tile
    const int third = e >> 2;    // columns [64 third, 64 third + 64) of the 192-column head tile: exactly Q, K or V
    const int r_local = quarter * 32 + lane;
    const uint32_t leader_tmem_empty0 = mapa_u32(smem_u32(&tmem_empty_bar[0]), 0);
    const uint32_t leader_tmem_empty1 = mapa_u32(smem_u32(&tmem_empty_bar[1]), 0);
    const int g = lane >> 2, tq = lane & 3;
    const int MT = (p.S + 15) >> 4;          // 16-query tiles per sample
    const int tasks = STR ? (rank == 0 ? 10 : 11) : p.ns * MT;   // STR: 3 whole samples + 1 (CTA 0) or 2 (CTA 1) query tiles of sample 3
    const bool issuer = (threadIdx.x == 4 * 32);
    [[maybe_unused]] const int koff = (STR && rank != 0) ? 16 : 0;   // K / V line of local row r = r + koff (CTA 1's window starts 16 rows early)
    [[maybe_unused]] const uint32_t peer = rank ^ 1u;
    int it = 0;
    for (int t = pair; t < num_tiles; t += num_pairs, ++it) {
      const int acc = it & 1;
      const int rt = t / p.heads, head = t - rt * p.heads;
      const int sample0 = STR ? rt * kStrSamples : (2 * rt + static_cast<int>(rank)) * p.ns;   // first sample of this CTA (STR: pair) tile
      const int m0 = STR ? rt * kStrRowsPair + static_cast<int>(rank) * kStrRows0 : sample0 * p.S;
      const int row = m0 + r_local;
      uint8_t* buf = bufs + (it & 1) * kBufB;
      if constexpr (STR) {  // this tile's foreign K / V lines: the receiver arms its barrier, the bytes may already be on their way
        if (issuer) mbar_expect_tx(&halo_full[it & 1], rank == 0 ? kStrHalo0 : kStrHalo1);
      }
      const float rstd = (p.ln_stats != nullptr && row < p.M) ? row_rstd(p.ln_stats, row, p.ln_eps) : 1.0f;
      mbar_wait(&tmem_full_bar[acc], (it >> 1) & 1u);
      tc_fence_after();
      // ---- (1) this warp's 32 rows x 64 columns (one of Q / K / V): rstd * acc + bias' -> bf16 -> swizzled lines ----
      {
        const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + static_cast<uint32_t>(acc * kAccStride + third * 64);
        uint8_t* mat = STR ? (third == 0 ? buf : buf + kMatBytes + (third - 1) * kStrKvBytes) : buf + third * kMatBytes;
        const int lline = r_local + (third == 0 ? 0 : koff);
        uint8_t* line = mat + lline * 128;
        const int sw = lline & 7;
        // STR: CTA 0's rows 124..127 duplicate CTA 1's first four (their K / V lines belong to the peer's stores); the K / V lines
        // of sample 3 also go to the peer: CTA 0 rows 108..123 -> peer lines 0..15, CTA 1 rows 0..19 -> peer lines 124..143
        [[maybe_unused]] bool keep_local = true;
        [[maybe_unused]] int pline = -1;
        if constexpr (STR) {
          if (third != 0) {
            if (rank == 0) {
              keep_local = r_local < kStrRows0;
              if (r_local >= kStrBase1 && r_local < kStrRows0) pline = r_local - kStrBase1;
            } else if (r_local < kStrS - 16) {
              pline = kStrRows0 + r_local;
            }
            // the peer must be done with what this CTA sent into the same buffer two tiles ago (warp-uniform wait: whole warp)
            if (it >= 2 && ((rank == 0 && quarter == 3) || (rank != 0 && quarter == 0))) mbar_wait_cluster(&halo_free[it & 1], ((it >> 1) & 1u) ^ 1u);
          }
        }
        [[maybe_unused]] const uint32_t pmat = STR ? mapa_u32(smem_u32(mat), peer) : 0u;
        [[maybe_unused]] const uint32_t pbar = STR ? mapa_u32(smem_u32(&halo_full[it & 1]), peer) : 0u;
#pragma unroll 1
        for (int c = 0; c < 2; ++c) {
          uint32_t v[32];
          tmem_ld32(t_row + static_cast<uint32_t>(c * 32), v);
          const float4* b4 = reinterpret_cast<const float4*>(p.bias + head * kTileN + third * 64 + c * 32);
          tmem_ld_wait_dep(v);
          if ((p.debug & 2) && !STR) continue;   // (the straddle form's tile protocol needs the stores)
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float4 ba = __ldg(b4 + 2 * i), bb = __ldg(b4 + 2 * i + 1);
            uint4 o;
            o.x = pack_bf16x2(fmaf(rstd, __uint_as_float(v[8 * i + 0]), ba.x), fmaf(rstd, __uint_as_float(v[8 * i + 1]), ba.y));
            o.y = pack_bf16x2(fmaf(rstd, __uint_as_float(v[8 * i + 2]), ba.z), fmaf(rstd, __uint_as_float(v[8 * i + 3]), ba.w));
            o.z = pack_bf16x2(fmaf(rstd, __uint_as_float(v[8 * i + 4]), bb.x), fmaf(rstd, __uint_as_float(v[8 * i + 5]), bb.y));
            o.w = pack_bf16x2(fmaf(rstd, __uint_as_float(v[8 * i + 6]), bb.z), fmaf(rstd, __uint_as_float(v[8 * i + 7]), bb.w));
            if (keep_local) *reinterpret_cast<uint4*>(line + (((4 * c + i) ^ sw) << 4)) = o;
            if constexpr (STR) {
              if (pline >= 0) st_async_v4(pmat + static_cast<uint32_t>(pline * 128 + (((4 * c + i) ^ (pline & 7)) << 4)), o, pbar);
            }
          }
        }
        // all TMEM reads of this accumulator buffer are complete: hand it back to the MMA issuer
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(acc == 0 ? leader_tmem_empty0 : leader_tmem_empty1);
      }
      named_bar_sync(1, kEpiWarps * 32);  // the tile's Q, K, V are in shared memory
      if constexpr (STR) mbar_wait_cluster(&halo_full[it & 1], (it >> 1) & 1u);   // ... and the peer's lines of sample 3
      // ---- (2) attention: task = (sample u of this tile, 16-query tile mt) ----
      const uint32_t Qb = smem_u32(buf), Kb = Qb + kMatBytes, Vb = Kb + (STR ? kStrKvBytes : kMatBytes);
      for (int task = e; task < ((p.debug & 1) ? 0 : tasks); task += kEpiWarps) {
        int u, mt;
        if constexpr (STR) {
          if (rank == 0) {                       // samples 0-2 whole, query tile 0 of sample 3
            u = task / 3;
            mt = task - u * 3;
          } else if (task < 2) {                 // query tiles 1, 2 of sample 3
            u = 3;
            mt = task + 1;
          } else {                               // samples 4-6 whole
            u = 4 + (task - 2) / 3;
            mt = (task - 2) % 3;
          }
        } else {
          u = task / MT;
          mt = task - u * MT;
        }
        const int b = sample0 + u;
        if (b >= p.B) continue;  // tail tile: no such sample
        if (((p.debug & 8) && u != 0) || ((p.debug & 16) && mt != 0)) continue;  // measurement: a third of the tasks
        // first line of this sample in the Q matrix (STR, CTA 1: sample 3 starts 16 lines BEFORE its first local row) and in K / V
        const int rbq = STR ? u * kStrS - static_cast<int>(rank) * kStrRows0 : u * p.S;
        const uint32_t rb = static_cast<uint32_t>(rbq + koff);
        const int m0q = mt * 16;
        // every fragment row of this task has the same (row & 7) up to the lane's own offset: the 128B-swizzle XOR is computed
        // once per task, every ldmatrix address below is base + immediate
        const uint32_t r7 = (rb + static_cast<uint32_t>(lane & 7)) & 7u;
        const uint32_t c16 = static_cast<uint32_t>(lane >> 4), c8 = static_cast<uint32_t>(lane >> 3);
        uint32_t offq[4], offk[2];   // byte offsets of chunk 2x + (lane >> 4) (Q, V) and 4x + (lane >> 3) (K) inside a line
#pragma unroll
        for (int x = 0; x < 4; ++x) offq[x] = ((static_cast<uint32_t>(2 * x) + c16) ^ r7) << 4;
#pragma unroll
        for (int x = 0; x < 2; ++x) offk[x] = ((static_cast<uint32_t>(4 * x) + c8) ^ r7) << 4;
        const uint32_t qline = Qb + static_cast<uint32_t>(rbq + m0q + (lane & 15)) * 128u;
        const uint32_t kline = Kb + (rb + static_cast<uint32_t>(lane & 7)) * 128u;
        const uint32_t vline = Vb + (rb + static_cast<uint32_t>((lane & 7) + ((lane >> 3) & 1) * 8)) * 128u;
        // key mask of this sample as bits (one ballot per 32 keys) instead of one global load per score column
        uint32_t keep_lo = 0xffffffffu, keep_hi = 0xffffffffu;
        if (p.key_mask != nullptr) {
          const int32_t* mrow = p.key_mask + static_cast<long long>(b) * p.S;
          keep_lo = __ballot_sync(0xffffffffu, lane < p.S ? __ldg(mrow + lane) != 0 : false);
          if (NT > 4) keep_hi = __ballot_sync(0xffffffffu, lane + 32 < p.S ? __ldg(mrow + lane + 32) != 0 : false);
        }
        uint32_t qa[4][4];
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) ldsm_x4(qa[ks], qline + offq[ks]);
        float s[NT][4];
#pragma unroll
        for (int j = 0; j < NT; ++j) {
          s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f;
#pragma unroll
          for (int kp = 0; kp < 2; ++kp) {
            uint32_t kb4[4];
            ldsm_x4(kb4, kline + static_cast<uint32_t>(j * 1024) + offk[kp]);
            mma_bf16_16816(s[j], qa[2 * kp], kb4[0], kb4[1]);
            mma_bf16_16816(s[j], qa[2 * kp + 1], kb4[2], kb4[3]);
          }
        }
        // (q·k / 8 + mask) * log2(e): the softmax below is then a bare ex2 (modeling.py:333-339)
        float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
        for (int j = 0; j < NT; ++j) {
          float madd[2];
#pragma unroll
          for (int ee = 0; ee < 2; ++ee) {
            const int key = 8 * j + 2 * tq + ee;
            const uint32_t bits = (8 * j < 32) ? keep_lo : keep_hi;   // compile-time per j
            madd[ee] = key >= p.S ? -INFINITY : (((bits >> (key & 31)) & 1u) ? 0.f : -10000.f * kLog2e);
          }
          s[j][0] = fmaf(s[j][0], kScaleLog2e, madd[0]);
          s[j][1] = fmaf(s[j][1], kScaleLog2e, madd[1]);
          s[j][2] = fmaf(s[j][2], kScaleLog2e, madd[0]);
          s[j][3] = fmaf(s[j][3], kScaleLog2e, madd[1]);
          mx0 = fmaxf(mx0, fmaxf(s[j][0], s[j][1]));
          mx1 = fmaxf(mx1, fmaxf(s[j][2], s[j][3]));
        }
        mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
        mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
        mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
        mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
        float sum0 = 0.f, sum1 = 0.f;
        uint32_t pa[KS][4];
        if (NT < 2 * KS) pa[KS - 1][2] = pa[KS - 1][3] = 0u;   // the score tile that was not computed: probability 0
#pragma unroll
        for (int j = 0; j < NT; ++j) {
          const float p0 = ex2_fast(s[j][0] - mx0), p1 = ex2_fast(s[j][1] - mx0);
          const float p2 = ex2_fast(s[j][2] - mx1), p3 = ex2_fast(s[j][3] - mx1);
          sum0 += p0 + p1;
          sum1 += p2 + p3;
          pa[j >> 1][(j & 1) * 2 + 0] = pack_bf16x2(p0, p1);
          pa[j >> 1][(j & 1) * 2 + 1] = pack_bf16x2(p2, p3);
        }
        sum0 += __shfl_xor_sync(0xffffffffu, sum0, 1);
        sum0 += __shfl_xor_sync(0xffffffffu, sum0, 2);
        sum1 += __shfl_xor_sync(0xffffffffu, sum1, 1);
        sum1 += __shfl_xor_sync(0xffffffffu, sum1, 2);
        float o[8][4];
#pragma unroll
        for (int d = 0; d < 8; ++d) o[d][0] = o[d][1] = o[d][2] = o[d][3] = 0.f;
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
#pragma unroll
          for (int dp = 0; dp < 4; ++dp) {
            uint32_t vb4[4];
            ldsm_x4_trans(vb4, vline + static_cast<uint32_t>(ks * 2048) + offq[dp]);
            mma_bf16_16816(o[2 * dp], pa[ks], vb4[0], vb4[1]);
            mma_bf16_16816(o[2 * dp + 1], pa[ks], vb4[2], vb4[3]);
          }
        }
        // the 16 query lines of this task are in registers since the top: they now take the context rows (lines of queries
        // >= S belong to the next sample's task and are left alone)
        const float inv0 = 1.0f / sum0, inv1 = 1.0f / sum1;
        __syncwarp();
        const int r0 = m0q + g, r1 = r0 + 8;
        const uint32_t r7o = (rb + static_cast<uint32_t>(g)) & 7u;   // (row & 7) of both output rows (r1 = r0 + 8)
        const uint32_t l0 = Qb + static_cast<uint32_t>(rbq + r0) * 128u + 4u * tq, l1 = l0 + 1024u;
#pragma unroll
        for (int d = 0; d < 8; ++d) {
          const uint32_t off = (static_cast<uint32_t>(d) ^ r7o) << 4;
          if (r0 < p.S)
            asm volatile("st.shared.b32 [%0], %1;" ::"r"(l0 + off), "r"(pack_bf16x2(o[d][0] * inv0, o[d][1] * inv0)) : "memory");
          if (r1 < p.S)
            asm volatile("st.shared.b32 [%0], %1;" ::"r"(l1 + off), "r"(pack_bf16x2(o[d][2] * inv1, o[d][3] * inv1)) : "memory");
        }
      }
      // ---- (3) context tile out: ONE TMA store of the whole-sample rows (rows >= M are clipped by the tensor map) ----
      fence_proxy_async_smem();
      if (issuer) tma_store_wait_read<0>();  // the store of the PREVIOUS tile (other buffer) has read its lines: that buffer is
                                             // free again before any warp passes the barrier below into the next tile
      named_bar_sync(1, kEpiWarps * 32);
      if (issuer) {
        if constexpr (STR) {
          // every local task has read the peer's lines of this buffer: the peer may overwrite them two tiles on
          mbar_arrive_cluster(mapa_u32(smem_u32(&halo_free[it & 1]), peer));
          if (m0 < p.M && !(p.debug & 4)) tma_store_2d(rank == 0 ? &tmO2 : &tmO, buf, head * HD, m0);   // 124 / 128 rows
        } else {
          if (m0 < p.M && !(p.debug & 4)) tma_store_2d(&tmO, buf, head * HD, m0);
        }
        tma_store_commit();
      }
    }
    if (issuer) tma_store_wait<0>();  // shared memory must outlive the last store's read
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == 2) tmem_dealloc_pair<512>(tmem_base);
}

template <int KS, int NT, bool STR = false>
int launch_qkv_attn(cudaStream_t stream, const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmO, const QkvAttnParams& p,
                    const CUtensorMap* tmO2 = nullptr) {
  auto kern = qkv_attn_pair_kernel<KS, NT, STR>;
  constexpr int smem = STR ? kSmemBytesStr : kSmemBytes;
  static bool configured = false;
  if (!configured) {
    RGQA_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = true;
  }
  const int row_tiles = STR ? (p.B + kStrSamples - 1) / kStrSamples : (p.B + 2 * p.ns - 1) / (2 * p.ns);
  const int tiles = row_tiles * p.heads;
  const int max_pairs = launch_sms() / 2;
  const int pairs = tiles < max_pairs ? tiles : max_pairs;
  RGQA_CUDA_OK(launch_pdl(kern, dim3(2 * pairs), dim3(kThreads), smem, stream, tmA, tmB, tmO, tmO2 ? *tmO2 : tmO, p));
  return 0;
}

}  // namespace

int qkv_attention_fused_ok(int B, int S, int heads, int K) {
  return option_get(OPT_ATTN_FUSED) != 0 && heads == 12 && K == 768 && S >= 1 && S <= 64 && static_cast<long long>(B) * S >= 1024;
}

int qkv_attention_fused(cudaStream_t stream, const void* A, long long lda, const void* W_hm, const float* bias_hm,
                        const float* ln_stats, float ln_eps, const int32_t* key_mask, void* ctx, long long ldo, int B, int S,
                        int heads, int a_f16) {
  RGQA_CHECK(A && W_hm && bias_hm && ctx && B > 0, "qkv_attention_fused: null argument");
  RGQA_CHECK(heads == 12 && S >= 1 && S <= 64, "qkv_attention_fused: built for 12 heads of 64 and S <= 64 (got heads=%d S=%d)", heads, S);
  RGQA_CHECK(reinterpret_cast<uintptr_t>(bias_hm) % 16 == 0, "qkv_attention_fused: bias must be 16-byte aligned");
  const int M = B * S;
  QkvAttnParams p;
  p.B = B;
  p.S = S;
  p.ns = BM / S;
  p.heads = heads;
  p.M = M;
  p.bias = bias_hm;
  p.ln_stats = ln_stats;
  p.ln_eps = ln_eps;
  p.key_mask = key_mask;
  p.a_f16 = a_f16;
  p.debug = option_get(OPT_GEMM_DEBUG);
  CUtensorMap tmA, tmB, tmO;
  if (make_tmap(&tmA, A, M, 768, lda, BM, 2)) return 1;
  if (make_tmap(&tmB, W_hm, static_cast<long long>(heads) * kTileN, 768, 768, kTileN / 2, 2)) return 1;
  const bool straddle = S == kStrS && option_get(OPT_ATTN_STRADDLE) != 0;
  CUtensorMap tmO2;
  if (straddle) {  // CTA 1 stores 128 context rows per tile, CTA 0 124
    if (make_tmap_ex(&tmO, ctx, M, static_cast<long long>(heads) * HD, ldo, BM, HD, 2, 128) ||
        make_tmap_ex(&tmO2, ctx, M, static_cast<long long>(heads) * HD, ldo, kStrRows0, HD, 2, 128))
      return 1;
  } else if (make_tmap_ex(&tmO, ctx, M, static_cast<long long>(heads) * HD, ldo, p.ns * S, HD, 2, 128)) {
    return 1;
  }
  void* tok = prof_begin(stream);
  int rc;
  if (straddle) {
    rc = launch_qkv_attn<3, 5, true>(stream, tmA, tmB, tmO, p, &tmO2);
  } else
  switch ((S + 7) / 8) {
    case 1: rc = launch_qkv_attn<1, 1>(stream, tmA, tmB, tmO, p); break;
    case 2: rc = launch_qkv_attn<1, 2>(stream, tmA, tmB, tmO, p); break;
    case 3: rc = launch_qkv_attn<2, 3>(stream, tmA, tmB, tmO, p); break;
    case 4: rc = launch_qkv_attn<2, 4>(stream, tmA, tmB, tmO, p); break;
    case 5: rc = launch_qkv_attn<3, 5>(stream, tmA, tmB, tmO, p); break;
    case 6: rc = launch_qkv_attn<3, 6>(stream, tmA, tmB, tmO, p); break;
    case 7: rc = launch_qkv_attn<4, 7>(stream, tmA, tmB, tmO, p); break;
    default: rc = launch_qkv_attn<4, 8>(stream, tmA, tmB, tmO, p); break;
  }
  if (rc) return 1;
  // the projection's FLOPs (2 M N K over the whole-sample rows) + the attention's (4 S S 64 per head)
  prof_end(tok, stream, 2.0 * M * 2304.0 * 768.0 + 4.0 * B * static_cast<double>(S) * S * heads * HD, PROF_GEMM_TC);
  note_launch();
  RGQA_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace rgqa
