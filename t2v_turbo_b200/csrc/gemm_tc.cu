// Implicit GEMM on tcgen05 (sm_100a): the one tensor-core kernel behind every Linear, Conv2d 3x3/1x1,
// strided conv, Conv3d (3,1,1) and Conv1d(k=1) of the VideoCrafter2 UNet and the KL-VAE decoder.
//
//   out[m, n] = epi( alpha * sum_{tap, c} A[point(m) + off(tap), c] * W[n, tap * c_in + c] )
//
// Design (one CTA per SM, persistent over output tiles, 384 threads):
//   warp 0 / lane 0 : TMA producer.  A tile = 5-D box {64 ch, b1, b2, b3, b4} of the channels-last
//                     activation (im2col-free: the tap shift is a coordinate offset; out-of-range
//                     coordinates are zero-filled by TMA = conv padding); B tile = {64, BLOCK_N}
//                     of the K-major weight matrix.  Both land in 128B-swizzled smem.
//   warp 1 / lane 0 : MMA issuer. 4 x tcgen05.mma (M=128, N=BLOCK_N, K=16) per 64-wide k-block,
//                     fp32 accumulators in TMEM, double buffered (2 x BLOCK_N columns).
//   warp 2          : TMEM allocator.
//   warps 4..11     : epilogue, two warpgroups splitting the accumulator columns in 32-column chunks.
//                     tcgen05.ld (thread = output row), bias / residual / GEGLU, 16-byte stores.
//                     Overlaps the next tile's main loop.
// Pipelines: smem full/empty ring (TMA <-> MMA) and TMEM full/empty (MMA <-> epilogue).
// Split-K (small-M layers): the K range is cut into split_k slices handled by different CTAs which
// accumulate fp32 partials with vector reductions into a workspace; a finalize kernel applies the epilogue.
#include <cuda.h>
#include <string.h>

#include "../../include/t2v_b200.h"
#include "host_common.h"
#include "ptx.cuh"

namespace t2v {

struct GemmParams {
  int32_t box[4];
  int32_t ntile[4];
  int32_t n_tiles_n;
  int32_t n_tiles_mn;  // m tiles * n tiles
  int32_t num_tiles;   // n_tiles_mn * split_k
  int32_t rows_in_box;
  int32_t n_taps;
  int32_t tap_off[T2V_MAX_TAPS][4];
  int32_t tap_ch_off[T2V_MAX_TAPS];
  int32_t kb_per_tap;
  int32_t kb_src0;
  int32_t total_kb;
  int32_t split_k;
  int32_t kb_per_split;
  int32_t b_batch_dim;
  int32_t n_rows_b;  // N
  int32_t n_out;
  int64_t o_size[4];
  int64_t o_stride[4];
  int64_t r_stride[4];
  void* out;
  const void* residual;
  const float* bias;
  int64_t bias_row_stride;
  int32_t bias_dim;
  int32_t bias_div;
  float alpha;
  uint32_t flags;
  int32_t vec_ok;
  int32_t bias_vec_ok;
  uint32_t a_tile_bytes;
  float* ws;  // split-K workspace [points][N] fp32
  int32_t epi_mode;  // 0 = direct row stores, 1 = smem-staged TMA store (+ TMA residual load)
  int32_t n_stages;  // pipeline stages in use
  const float2* row_stats;  // folded LayerNorm: per input row (rstd, -rstd*mean); nullptr = off
  const float* col_sum;     // folded LayerNorm: per output column sum_k W'[n,k]
  int32_t ln_raw;           // row_stats holds raw (sum x, sum x^2) per row (accumulated by a producer GEMM)
  float ln_inv_c, ln_eps;   // 1 / channels and eps for ln_raw
  float* row_accum;         // LN == 2: fp32 [M][2], += (sum, sum of squares) of each output row
  float* col_accum;         // LN == 3: fp32 [samples][N][2], += per-channel (sum, sum of squares) of the output rows
  int32_t cs_mult[4];       // LN == 3: sample index of an output point = sum_j coord[j] * cs_mult[j]
  int32_t cs_direct;        // LN == 3: tiles span several samples -> per-half-block global reductions instead of the CTA table
  uint32_t smem_epi_off;  // offset of the epilogue staging buffers
  int32_t dbg;              // timing experiments only (results are wrong when non-zero)
};

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;
constexpr int kATileBytes = kBlockM * kBlockK * 2;  // 16 KB
constexpr int kEpiWGs = 2;                            // epilogue warpgroups (128 threads = 128 accumulator rows each); 3 measured: no gain, one pipeline stage less
constexpr int kThreads = 128 + 128 * kEpiWGs;
constexpr int kEpiWarps = 4 * kEpiWGs;
constexpr int kEpiBufBytes = 128 * 64;               // 128 rows x 32 bf16 columns, 64-byte swizzle
constexpr int kEpiBufs = 3;                          // per epilogue warpgroup
constexpr int kEpiBytes = kEpiWGs * kEpiBufs * kEpiBufBytes;
constexpr int kMaxStages = 12;
constexpr int kSmemLimit = 227 * 1024;
constexpr int kSmemBudget = kSmemLimit - kEpiBytes - 6144;

// PAIR: two CTAs of a cluster work on one 256 x BN tile with tcgen05.mma.cta_group::2 — each CTA stages its own
// 128 rows of A and HALF of the weight tile, so the L2->SM operand traffic per FLOP drops by a third to a half.
template <int BN, bool PAIR = false>
struct GemmCfg {
  static constexpr int kBTileBytes = (PAIR ? BN / 2 : BN) * kBlockK * 2;
  static constexpr int kStageBytes = kATileBytes + kBTileBytes;
  static constexpr int kStagesRaw = kSmemBudget / kStageBytes;
  static constexpr int kStages = kStagesRaw > 8 ? 8 : kStagesRaw;   // streaming mode
  static constexpr int kAccStride = BN <= 32 ? 32 : BN <= 64 ? 64 : BN <= 128 ? 128 : 256;
  static constexpr int kTmemCols = 2 * kAccStride;
  static constexpr int kBarBytes = (2 * kMaxStages + 5 + kEpiWGs * kEpiBufs) * 8 + 16 + 16 + 4 * 256 * 4;
};

// erf-GELU: g * Phi(g) with Phi(g) = sigmoid(g (c0 + c1 g^2 + c2 g^4)) — a minimax fit of logit(Phi) by an odd
// polynomial, max |error| 2.6e-5 over the reals (0.3 % of a bf16 ulp at unit scale; scripts in DESIGN.md §3).
// g^2 is clamped at 81 so the quartic term cannot turn the polynomial over for huge |g|.
// Cost: 5 FMA-pipe + 1 min + 2 MUFU (ex2, rcp); the GEGLU epilogue is FMA-pipe bound, so this matters.
__device__ __forceinline__ float gelu_erf(float g) {
  constexpr float kL = -1.4426950408889634f;  // -log2(e): sigmoid(p) = 1 / (1 + 2^(-p log2 e))
  const float g2 = fminf(g * g, 81.0f);
  float p = fmaf(g2, kL * -0.0007030335803817405f, kL * 0.07401129203239541f);
  p = fmaf(g2, p, kL * 1.5950157686368724f);
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(g * p));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
  return g * r;
}

__device__ __forceinline__ void red_add_v2(float* addr, float a, float b) {
  asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(addr), "f"(a), "f"(b) : "memory");
}
__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d)
               : "memory");
}

// split-K epilogue of 4 consecutive output columns [n0, n0 + 4) of one output point from the fp32 partial sums `wp`
// (shared by the in-kernel fix-up phase and the legacy finalize kernel).  `clear`: zero the partial sums after reading.
__device__ __forceinline__ void splitk_finalize_quad(const GemmParams& p, float* wp, int64_t point, int64_t out_off,
                                                     int64_t res_off, int64_t bias_row, int64_t smp, int n0, bool clear) {
  const bool out_f32 = (p.flags & T2V_EPI_OUT_F32) != 0;
  const bool gelu = (p.flags & T2V_EPI_GELU) != 0;
  float fsum = 0.f, fsq = 0.f;
  float cv[4] = {0.f, 0.f, 0.f, 0.f};
  float rs = 1.f, rm = 0.f;
  if (p.row_stats) {  // folded LayerNorm (A is a plain [M, K] matrix: point == row)
    const float2 st = __ldcg(p.row_stats + point);
    rs = st.x, rm = st.y;
    if (p.ln_raw) {
      const float mean = st.x * p.ln_inv_c;
      rs = rsqrtf(fmaxf(fmaf(st.y, p.ln_inv_c, -mean * mean), 0.f) + p.ln_eps);
      rm = -rs * mean;
    }
  }
  float w4[4];
  if (n0 + 3 < p.n_rows_b && (p.n_rows_b & 3) == 0) {
    const float4 t = __ldcg(reinterpret_cast<const float4*>(wp));
    w4[0] = t.x, w4[1] = t.y, w4[2] = t.z, w4[3] = t.w;
    if (clear) *reinterpret_cast<float4*>(wp) = make_float4(0.f, 0.f, 0.f, 0.f);
  } else {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      w4[j] = n0 + j < p.n_rows_b ? __ldcg(wp + j) : 0.f;
      if (clear && n0 + j < p.n_rows_b) wp[j] = 0.f;
    }
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    if (n0 + j < p.n_rows_b) {
      float v = w4[j];
      if (p.row_stats) v = fmaf(rs, v, rm * __ldg(p.col_sum + n0 + j));
      if (p.bias) v += p.bias[bias_row * p.bias_row_stride + n0 + j];
      if (gelu) v = gelu_erf(v);
      if (p.residual) v += __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p.residual)[res_off + n0 + j]);
      if (out_f32)
        reinterpret_cast<float*>(p.out)[out_off + n0 + j] = v;
      else
        reinterpret_cast<__nv_bfloat16*>(p.out)[out_off + n0 + j] = __float2bfloat16_rn(v);
      fsum += v;
      fsq = fmaf(v, v, fsq);
      cv[j] = out_f32 ? v : __bfloat162float(__float2bfloat16_rn(v));
    }
  }
  if (p.row_accum) red_add_v2(p.row_accum + 2 * point, fsum, fsq);
  if (p.col_accum && n0 + 3 < p.n_rows_b) {
    float* ca = p.col_accum + (smp * p.n_rows_b + n0) * 2;
    red_add_v4(ca, cv[0], cv[0] * cv[0], cv[1], cv[1] * cv[1]);
    red_add_v4(ca + 4, cv[2], cv[2] * cv[2], cv[3], cv[3] * cv[3]);
  }
}

// LN: 0 = plain epilogue; 1 = this GEMM CONSUMES a LayerNorm (folded: per-row statistics applied in the epilogue);
// 2 = this GEMM PRODUCES a LayerNorm input (its epilogue accumulates per-row sum / sum of squares of the output).
// 3 = this GEMM PRODUCES a GroupNorm input: per-(sample, channel) sum / sum of squares of the bf16 output, reduced
//     over the tile's rows out of the staging buffer (the GroupNorm statistics pass disappears).
// Separate instantiations keep the extra registers / instructions out of the plain kernels (measured: +10 % on the
// epilogue-bound GEMMs when the paths shared one kernel).
template <int BN, bool PAIR, int LN>
__global__ void __launch_bounds__(kThreads, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
               const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmOut,
               const __grid_constant__ CUtensorMap tmRes, const GemmParams p) {
  using Cfg = GemmCfg<BN, PAIR>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);

  uint8_t* smem_epi = smem + p.smem_epi_off;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem_epi + kEpiBytes);
  uint64_t* empty_bar = full_bar + kMaxStages;
  uint64_t* tfull_bar = empty_bar + kMaxStages;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint64_t* res_bar = tempty_bar + 2;  // [kEpiWGs][kEpiBufs]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(res_bar + kEpiWGs * kEpiBufs);
  // [2 accumulators][256] floats, 16-byte aligned for float4 reads
  float* bias_smem = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(tmem_slot + 4) + 15) & ~uintptr_t(15));

  const int warp = __shfl_sync(0xffffffffu, int(threadIdx.x >> 5), 0);  // provably warp-uniform (see the role branches)
  const int lane = threadIdx.x & 31;
  const uint32_t cta_rank = PAIR ? cluster_ctarank() : 0u;
  const int tile0 = PAIR ? int(blockIdx.x >> 1) : int(blockIdx.x);
  const int tile_step = PAIR ? int(gridDim.x >> 1) : int(gridDim.x);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA0);
    tma_prefetch_desc(&tmA1);
    tma_prefetch_desc(&tmB);
    if (p.epi_mode == 1) {
      tma_prefetch_desc(&tmOut);
      if (p.residual != nullptr) tma_prefetch_desc(&tmRes);
    }
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < p.n_stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull_bar[a], 1);
      mbar_init(&tempty_bar[a], PAIR ? 2 * kEpiWarps : kEpiWarps);  // PAIR: both CTAs' epilogues release the leader
    }
    for (int i = 0; i < kEpiWGs * kEpiBufs; ++i) mbar_init(&res_bar[i], 1);
    fence_mbar_init();
  }
  if (warp == 2) {
    if (PAIR) {
      tmem_alloc_2sm(tmem_slot, Cfg::kTmemCols);
      tmem_relinquish_2sm();
    } else {
      tmem_alloc(tmem_slot, Cfg::kTmemCols);
      tmem_relinquish();
    }
  }
  pdl_launch_dependents();
  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();  // the peer's barriers must be initialised before any remote arrive / multicast commit
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();  // everything above (barrier init, TMEM allocation, descriptor prefetch) overlaps the previous kernel

  // Role branches: `warp` is broadcast through a shuffle so that the compiler can prove it warp-uniform, and the
  // single issuing lane is chosen by elect.sync.  With a plain `threadIdx.x == k` test ptxas treats every value in
  // the branch as divergent and wraps each TMA / MMA instruction in a waterfall loop of R2UR.BROADCASTs (measured:
  // 12-24 % of the conv GEMMs' time, because the single issuing thread could no longer keep up with the tensor pipe).
  if (warp == 0 && elect_one()) {
    // ------------------------------------------------------------ TMA producer
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = tile0; tile < p.num_tiles; tile += tile_step) {
      const int ks = tile / p.n_tiles_mn;
      const int mn = tile - ks * p.n_tiles_mn;
      const int n_tile = mn % p.n_tiles_n;
      int m_tile = (mn / p.n_tiles_n) * (PAIR ? 2 : 1) + int(cta_rank);
      int o[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        o[j] = (j < 3 ? m_tile % p.ntile[j] : m_tile) * p.box[j];  // last dim unbounded: a tile past the end is all out-of-range
        m_tile /= p.ntile[j];
      }
      int bbatch = 0;
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (j == p.b_batch_dim) bbatch = o[j];
      const int b_row = PAIR ? n_tile * BN + int(cta_rank) * (BN / 2) : n_tile * BN;
      const int kb_begin = ks * p.kb_per_split;
      const int kb_end = min(kb_begin + p.kb_per_split, p.total_kb);
      int tap = kb_begin / p.kb_per_tap;
      int kb = kb_begin - tap * p.kb_per_tap;
      int idx = kb_begin;
      while (idx < kb_end) {
        // everything that depends on the tap is hoisted out of the per-k-block loop: the issuing thread has
        // ~320 cycles per k-block at BN = 160 and every dependent instruction costs it 4+ of them
        const int c1 = o[0] + p.tap_off[tap][0];
        const int c2 = o[1] + p.tap_off[tap][1];
        const int c3 = o[2] + p.tap_off[tap][2];
        const int c4 = o[3] + p.tap_off[tap][3];
        const int ch0 = p.tap_ch_off[tap];
        const int kb_hi = min(p.kb_per_tap, kb + (kb_end - idx));
        for (; kb < kb_hi; ++kb, ++idx) {
          mbar_wait_relaxed(&empty_bar[stage], phase ^ 1u);
          uint8_t* sA = smem + stage * Cfg::kStageBytes;
          uint8_t* sB = sA + kATileBytes;
          const bool src0 = kb < p.kb_src0;
          const CUtensorMap* tmA = src0 ? &tmA0 : &tmA1;
          const int ch = ch0 + (src0 ? kb : kb - p.kb_src0) * kBlockK;
          if (PAIR) {
            // The leader's barrier expects the bytes of BOTH CTAs; the peer only issues its loads (their
            // complete_tx lands on the leader's barrier).  The peer can never run a phase ahead: its stage is
            // released by the same multicast commit that follows the leader's barrier completing.  (A remote
            // arrive per k-block would cost a cluster-scope release fence on the critical path.)
            if (cta_rank == 0) mbar_expect_tx(&full_bar[stage], 2u * (p.a_tile_bytes + uint32_t(Cfg::kBTileBytes)));
            tma_load_5d_2sm(sA, tmA, &full_bar[stage], ch, c1, c2, c3, c4);
            tma_load_3d_2sm(sB, &tmB, &full_bar[stage], idx * kBlockK, b_row, 0);
          } else {
            mbar_expect_tx(&full_bar[stage], p.a_tile_bytes + uint32_t(Cfg::kBTileBytes));
            tma_load_5d(sA, tmA, &full_bar[stage], ch, c1, c2, c3, c4);
            tma_load_3d(sB, &tmB, &full_bar[stage], idx * kBlockK, b_row, bbatch);
          }
          if (++stage == p.n_stages) {
            stage = 0;
            phase ^= 1u;
          }
        }
        kb = 0;
        ++tap;
      }
    }
  } else if (warp == 1 && (!PAIR || cta_rank == 0)) {
    // ------------------------------------------------------------ MMA issuer (PAIR: the leader CTA issues for both)
    // The whole warp walks the loop — uniform control flow keeps the descriptor arithmetic on the uniform datapath — and only
    // tcgen05.mma / commit are predicated on the elected lane.  (Inside a single-thread branch every k-block cost a chain of
    // ~10 dependent R2UR / shift / add instructions before its first MMA, which bounds the narrow tiles: BN = 64 is 128
    // tensor cycles per k-block.)
    const bool leader = elect_one();
    constexpr uint32_t idesc = umma_idesc_bf16(PAIR ? 2 * kBlockM : kBlockM, BN, 0, 0);
    const uint32_t tm = __shfl_sync(0xffffffffu, tmem_base, 0);
    const uint64_t adesc0 = umma_desc_sw128(smem_u32(smem));
    const uint64_t bdesc0 = umma_desc_sw128(smem_u32(smem) + kATileBytes);
    constexpr uint64_t kStageDesc = Cfg::kStageBytes >> 4;   // one ring stage further, in descriptor (16-byte) units
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    for (int tile = tile0; tile < p.num_tiles; tile += tile_step, ++it) {
      const int ks = tile / p.n_tiles_mn;
      const int kb_begin = ks * p.kb_per_split;
      const int n_kb = min(kb_begin + p.kb_per_split, p.total_kb) - kb_begin;
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      mbar_wait_relaxed(&tempty_bar[acc], acc_phase ^ 1u);
      tc_fence_after();
      const uint32_t tmem_d = tm + acc * Cfg::kAccStride;
      for (int kb = 0; kb < n_kb; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint64_t adesc = adesc0 + uint64_t(stage) * kStageDesc;
        const uint64_t bdesc = bdesc0 + uint64_t(stage) * kStageDesc;
        if (leader) {
#pragma unroll
          for (int k = 0; k < kBlockK / 16; ++k) {
            // advance 16 elements (32 bytes) along K inside the 128B swizzle atom: +2 in addr>>4 units
            if (PAIR)
              umma_ss_2sm(tmem_d, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
            else
              umma_ss(tmem_d, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          if (PAIR)
            umma_commit_2sm(&empty_bar[stage]);
          else
            umma_commit(&empty_bar[stage]);
        }
        if (++stage == p.n_stages) {
          stage = 0;
          phase ^= 1u;
        }
      }
      if (leader) {
        if (PAIR)
          umma_commit_2sm(&tfull_bar[acc]);
        else
          umma_commit(&tfull_bar[acc]);
      }
    }
  } else if (warp >= 4 && p.epi_mode == 1) {
    // ------------------------------------------------------------ epilogue, smem-staged (8 warps = 2 warpgroups)
    // Each warpgroup owns 3 staging buffers of 128 rows x 32 bf16 (64-byte swizzle).  Per 32-column
    // output chunk: [residual chunk arrives by TMA] -> tcgen05.ld -> bias / GEGLU / +residual -> bf16 row
    // into smem -> warpgroup barrier -> one thread issues the TMA store (clips rows / columns outside the
    // output, fully coalesced, no LSU traffic).  Residual loads run two chunks ahead, the TMEM load of the
    // next chunk is issued before the current chunk is written out, and the tile's bias row is staged in
    // shared memory once per tile (broadcast reads) while the accumulator is still being computed.
    const int ew = warp - 4;
    const int lg = ew & 3;
    const int g = ew >> 2;
    const int r = lg * 32 + lane;
    const int et = ew * 32 + lane;  // 0..255 within the epilogue
    const bool lead_warp = (lg == 0);  // its elected lane issues this warpgroup's TMA loads / stores (always the same lane)
    const bool geglu = (p.flags & T2V_EPI_GEGLU) != 0;
    const bool gelu = (p.flags & T2V_EPI_GELU) != 0;
    const bool has_res = p.residual != nullptr;
    const bool alpha_one = (p.alpha == 1.0f);
    constexpr bool has_ln = (LN == 1);  // LayerNorm folded into this GEMM (host guarantees alpha == 1, plain [M, K] A)
    const int acc_cw = geglu ? 64 : 32;  // accumulator columns per 32-column output chunk
    const int chunks_per_tile = BN / acc_cw;
    uint8_t* ebuf = smem_epi + g * kEpiBufs * kEpiBufBytes;
    uint64_t* rbar = res_bar + g * kEpiBufs;
    const uint32_t res_bytes = uint32_t(p.rows_in_box) * 64u;
    const int sw = (r >> 1) & 3;
    uint32_t q = 0;  // chunks processed so far by this warpgroup (buffer = q % 3)
    // LN == 3: [4 warps of a warpgroup][BN columns][sum, sum of squares], behind the barrier / bias block
    float* cs_tab = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(bias_smem + 1024) + 15) & ~uintptr_t(15));
    if (LN == 3) {
      for (int i = et; i < 8 * BN; i += 128 * kEpiWGs) cs_tab[i] = 0.f;
      named_bar_sync(kEpiWGs + 1, 128 * kEpiWGs);
    }
    int it = 0;
    for (int tile = tile0; tile < p.num_tiles; tile += tile_step, ++it) {
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      const int n_tile = tile % p.n_tiles_n;
      int m_tile = (tile / p.n_tiles_n) * (PAIR ? 2 : 1) + int(cta_rank);
      int o[4];
      int64_t bias_row = 0, bias_row_first = 0, bias_row_last = 0;
      {
        int rr = r, rl = p.rows_in_box - 1;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          o[j] = (j < 3 ? m_tile % p.ntile[j] : m_tile) * p.box[j];
          m_tile /= p.ntile[j];
          const int ij = rr % p.box[j];
          rr /= p.box[j];
          const int il = rl % p.box[j];
          rl /= p.box[j];
          if (j == p.bias_dim) {
            bias_row = (o[j] + ij) / p.bias_div;
            bias_row_first = o[j] / p.bias_div;
            int64_t xl = o[j] + il;
            if (xl > p.o_size[j] - 1) xl = p.o_size[j] - 1;
            bias_row_last = xl / p.bias_div;
          }
        }
      }
      const int n_base = n_tile * BN;
      // (with a folded LayerNorm the column vectors are always staged; its bias, if any, is a plain [N] vector)
      const bool bias_uniform = has_ln || ((p.bias != nullptr) && (bias_row_first == bias_row_last));  // CTA-uniform
      const float* bias = p.bias ? p.bias + bias_row * p.bias_row_stride : nullptr;
      float* sbias = bias_smem + acc * 256;
      float* scs = bias_smem + 512 + acc * 256;
      if (bias_uniform) {
        if (et < BN) {
          const int n = n_base + et;
          sbias[et] = (p.bias != nullptr && n < p.n_rows_b) ? __ldg(p.bias + bias_row_first * p.bias_row_stride + n) : 0.f;
          if (has_ln) scs[et] = n < p.n_rows_b ? __ldg(p.col_sum + n) : 0.f;
        }
      }
      float ln_rs = 1.f, ln_rm = 0.f;
      if (has_ln) {
        const int64_t m = int64_t(o[0]) + r;
        if (m < p.o_size[0]) {
          const float2 st = __ldcg(p.row_stats + m);
          if (p.ln_raw) {
            const float mean = st.x * p.ln_inv_c;
            const float var = fmaxf(fmaf(st.y, p.ln_inv_c, -mean * mean), 0.f);
            ln_rs = rsqrtf(var + p.ln_eps);
            ln_rm = -ln_rs * mean;
          } else {
            ln_rs = st.x;
            ln_rm = st.y;
          }
        }
      }
      float acc_sum = 0.f, acc_sq = 0.f;  // LN == 2: this thread's row, this tile
      // LN == 3: this thread reduces channel pair (r & 15) over rows [16 (r >> 4), +16) in two half-blocks of 8 rows
      // (host: box[0] % 8 == 0 and the whole tile lies in ONE sample: box[j] == 1 wherever cs_mult[j] != 0).  The
      // column sums of every chunk go to a per-CTA shared-memory table and reach global memory once per tile.
      int64_t cs_off = -1;    // offset of this tile's (first) sample in col_accum, -1: tile outside the output
      int64_t cs_hb_off[2] = {0, 0};  // cs_direct: offset of the sample of each of this thread's two half-blocks
      int cs_nv[2] = {0, 0};  // valid rows (0..8) of each half-block: rows run along dim 0, the box may overhang the tensor
      if (LN == 3) {
        bool tile_ok = true;
        int64_t smp = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          tile_ok = tile_ok && (int64_t(o[j]) < p.o_size[j]);
          smp += int64_t(o[j]) * p.cs_mult[j];
        }
        if (tile_ok) cs_off = smp * int64_t(p.n_rows_b) * 2;
#pragma unroll
        for (int hb = 0; hb < 2; ++hb) {
          int rr = (r >> 4) * 16 + hb * 8;
          bool ok = tile_ok && rr < p.rows_in_box;
          const int i0 = rr % p.box[0];
          rr /= p.box[0];
          int64_t rel = 0;
#pragma unroll
          for (int j = 1; j < 4; ++j) {
            const int ij = rr % p.box[j];
            rr /= p.box[j];
            ok = ok && (int64_t(o[j]) + ij < p.o_size[j]);
            rel += int64_t(ij) * p.cs_mult[j];
          }
          int64_t nv = p.o_size[0] - (int64_t(o[0]) + i0);
          nv = nv < 0 ? 0 : (nv > 8 ? 8 : nv);
          cs_nv[hb] = ok ? int(nv) : 0;
          cs_hb_off[hb] = (smp + rel) * int64_t(p.n_rows_b) * 2;
        }
      }
      if (has_res && lead_warp && elect_one()) {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const int c = g + kEpiWGs * i;
          if (c < chunks_per_tile && n_base + c * acc_cw < p.n_rows_b) {
            const uint32_t buf = (q + i) % kEpiBufs;
            mbar_expect_tx(&rbar[buf], res_bytes);
            tma_load_5d(ebuf + buf * kEpiBufBytes, &tmRes, &rbar[buf], (n_base + c * acc_cw) / (acc_cw / 32), o[0], o[1],
                        o[2], o[3]);
          }
        }
      }
      if (bias_uniform) named_bar_sync(kEpiWGs + 1, 128 * kEpiWGs);  // bias row visible to all warpgroups
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + acc * Cfg::kAccStride + (static_cast<uint32_t>(lg * 32) << 16);
      bool arrived = false;
      uint32_t v[32];
      if (g < chunks_per_tile && n_base + g * acc_cw < p.n_rows_b) tmem_ld_32x32(taddr + g * acc_cw, v);
#pragma unroll 1
      for (int c = g; c < chunks_per_tile && n_base + c * acc_cw < p.n_rows_b; c += kEpiWGs) {
        const uint32_t buf = q % kEpiBufs;
        const uint32_t rphase = (q / kEpiBufs) & 1u;
        const int n0 = n_base + c * acc_cw;  // first W row (accumulator column) of this chunk
        const int oc0 = geglu ? (n0 >> 1) : n0;
        const bool has_next = (c + kEpiWGs < chunks_per_tile) && (n_base + (c + kEpiWGs) * acc_cw < p.n_rows_b);
        float f[32];
        tmem_wait_ld();
        if (p.dbg & 4) {
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = 0.f;
        } else if (geglu) {
          // packed W rows: [16 value | 16 gate] per 32 accumulator columns; two such groups per output chunk
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
            if (hh == 1) {
              tmem_ld_32x32(taddr + c * acc_cw + 32, v);
              tmem_wait_ld();
            }
            if (has_ln) {
              // out = rstd * acc + (-rstd * mean) * colsum[n] + bias'[n]   (LayerNorm folded into W', see t2v_b200.h)
              const float4* b4 = reinterpret_cast<const float4*>(sbias + c * acc_cw + hh * 32);
              const float4* s4 = reinterpret_cast<const float4*>(scs + c * acc_cw + hh * 32);
#pragma unroll
              for (int k4 = 0; k4 < 4; ++k4) {
                const float4 ba = b4[k4], sa = s4[k4], bg = b4[k4 + 4], sg = s4[k4 + 4];
                const float av[4] = {fmaf(ln_rm, sa.x, ba.x), fmaf(ln_rm, sa.y, ba.y), fmaf(ln_rm, sa.z, ba.z), fmaf(ln_rm, sa.w, ba.w)};
                const float gv[4] = {fmaf(ln_rm, sg.x, bg.x), fmaf(ln_rm, sg.y, bg.y), fmaf(ln_rm, sg.z, bg.z), fmaf(ln_rm, sg.w, bg.w)};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  const float a = fmaf(ln_rs, __uint_as_float(v[4 * k4 + i]), av[i]);
                  const float gt = fmaf(ln_rs, __uint_as_float(v[16 + 4 * k4 + i]), gv[i]);
                  f[hh * 16 + 4 * k4 + i] = a * gelu_erf(gt);
                }
              }
              continue;
            }
            float bv[32];
            if (bias_uniform) {
              const float4* b4 = reinterpret_cast<const float4*>(sbias + c * acc_cw + hh * 32);
#pragma unroll
              for (int k4 = 0; k4 < 8; ++k4) {
                const float4 t4 = b4[k4];
                bv[4 * k4 + 0] = t4.x;
                bv[4 * k4 + 1] = t4.y;
                bv[4 * k4 + 2] = t4.z;
                bv[4 * k4 + 3] = t4.w;
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) bv[j] = bias != nullptr ? __ldg(bias + n0 + hh * 32 + j) : 0.f;
            }
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              float a = __uint_as_float(v[j]);
              float gt = __uint_as_float(v[j + 16]);
              if (alpha_one) {
                a += bv[j];
                gt += bv[j + 16];
              } else {
                a = fmaf(p.alpha, a, bv[j]);
                gt = fmaf(p.alpha, gt, bv[j + 16]);
              }
              f[hh * 16 + j] = a * gelu_erf(gt);
            }
          }
        } else if (has_ln) {
          const float4* b4 = reinterpret_cast<const float4*>(sbias + c * acc_cw);
          const float4* s4 = reinterpret_cast<const float4*>(scs + c * acc_cw);
#pragma unroll
          for (int k4 = 0; k4 < 8; ++k4) {
            const float4 bv = b4[k4], sv = s4[k4];
            f[4 * k4 + 0] = fmaf(ln_rs, __uint_as_float(v[4 * k4 + 0]), fmaf(ln_rm, sv.x, bv.x));
            f[4 * k4 + 1] = fmaf(ln_rs, __uint_as_float(v[4 * k4 + 1]), fmaf(ln_rm, sv.y, bv.y));
            f[4 * k4 + 2] = fmaf(ln_rs, __uint_as_float(v[4 * k4 + 2]), fmaf(ln_rm, sv.z, bv.z));
            f[4 * k4 + 3] = fmaf(ln_rs, __uint_as_float(v[4 * k4 + 3]), fmaf(ln_rm, sv.w, bv.w));
          }
        } else if (bias_uniform) {
          const float4* b4 = reinterpret_cast<const float4*>(sbias + c * acc_cw);
#pragma unroll
          for (int k4 = 0; k4 < 8; ++k4) {
            const float4 bv = b4[k4];
            f[4 * k4 + 0] = fmaf(p.alpha, __uint_as_float(v[4 * k4 + 0]), bv.x);
            f[4 * k4 + 1] = fmaf(p.alpha, __uint_as_float(v[4 * k4 + 1]), bv.y);
            f[4 * k4 + 2] = fmaf(p.alpha, __uint_as_float(v[4 * k4 + 2]), bv.z);
            f[4 * k4 + 3] = fmaf(p.alpha, __uint_as_float(v[4 * k4 + 3]), bv.w);
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            f[j] = p.alpha * __uint_as_float(v[j]);
            if (bias != nullptr && n0 + j < p.n_rows_b) f[j] += __ldg(bias + n0 + j);
          }
        }
        if (has_next) {
          tmem_ld_32x32(taddr + (c + kEpiWGs) * acc_cw, v);  // in flight while this chunk is written out
        } else {
          // last TMEM read of this tile by this warp: release the accumulator early
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if (PAIR)
              mbar_arrive_cluster(smem_u32(&tempty_bar[acc]) & kPeerBitMask);
            else
              mbar_arrive(&tempty_bar[acc]);
          }
          arrived = true;
        }
        if (gelu && !geglu) {
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = gelu_erf(f[j]);
        }
        uint8_t* row = ebuf + buf * kEpiBufBytes + r * 64;
        if (has_res) {
          mbar_wait(&rbar[buf], rphase);
#pragma unroll
          for (int k4 = 0; k4 < 4; ++k4) {
            const uint4 rv = *reinterpret_cast<const uint4*>(row + ((k4 ^ sw) << 4));
            f[k4 * 8 + 0] += bf16_lo(rv.x);
            f[k4 * 8 + 1] += bf16_hi(rv.x);
            f[k4 * 8 + 2] += bf16_lo(rv.y);
            f[k4 * 8 + 3] += bf16_hi(rv.y);
            f[k4 * 8 + 4] += bf16_lo(rv.z);
            f[k4 * 8 + 5] += bf16_hi(rv.z);
            f[k4 * 8 + 6] += bf16_lo(rv.w);
            f[k4 * 8 + 7] += bf16_hi(rv.w);
          }
        }
        if (LN == 2) {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            acc_sum += f[j];
            acc_sq = fmaf(f[j], f[j], acc_sq);
          }
        }
#pragma unroll
        for (int k4 = 0; k4 < 4; ++k4) {
          uint4 ov;
          ov.x = pack_bf16(f[k4 * 8 + 0], f[k4 * 8 + 1]);
          ov.y = pack_bf16(f[k4 * 8 + 2], f[k4 * 8 + 3]);
          ov.z = pack_bf16(f[k4 * 8 + 4], f[k4 * 8 + 5]);
          ov.w = pack_bf16(f[k4 * 8 + 6], f[k4 * 8 + 7]);
          *reinterpret_cast<uint4*>(row + ((k4 ^ sw) << 4)) = ov;
        }
        if (!(p.dbg & 2)) {
          fence_proxy_async();
          named_bar_sync(1 + g, 128);
        }
        if (lead_warp && elect_one()) {
          if (!(p.dbg & 1)) tma_store_5d(&tmOut, ebuf + buf * kEpiBufBytes, oc0, o[0], o[1], o[2], o[3]);
          bulk_commit_group();
          bulk_wait_group_read<1>();  // every store but the one just issued has finished reading smem
          if (has_res) {
            const int c2 = c + 2 * kEpiWGs;
            if (c2 < chunks_per_tile && n_base + c2 * acc_cw < p.n_rows_b) {
              const uint32_t buf2 = (q + 2) % kEpiBufs;
              mbar_expect_tx(&rbar[buf2], res_bytes);
              tma_load_5d(ebuf + buf2 * kEpiBufBytes, &tmRes, &rbar[buf2], (n_base + c2 * acc_cw) / (acc_cw / 32), o[0],
                          o[1], o[2], o[3]);
            }
          }
        }
        if (LN == 3) {
          // column sums of this 128 x 32 chunk from the staged bf16 rows (exactly the values GroupNorm will read; issued AFTER
          // the chunk's TMA store so the store / residual pipeline is not delayed — both only read the buffer): packed fp32x2
          // accumulation of the channel pair over this thread's two 8-row half-blocks.
          const int cp = r & 15;  // channel pair: 4 bytes at word (cp & 3) of 16-byte chunk (cp >> 2)
          const uint8_t* cbase = ebuf + buf * kEpiBufBytes + (cp & 3) * 4;
          float2 s2[2], q2[2];
#pragma unroll
          for (int hb = 0; hb < 2; ++hb) {
            s2[hb] = make_float2(0.f, 0.f);
            q2[hb] = make_float2(0.f, 0.f);
            const int r0 = (r >> 4) * 16 + hb * 8;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int ii = i ^ ((lane >> 4) & 1);  // the two half-warps read rows of opposite parity: no bank conflict
              const int rw = r0 + ii;
              uint32_t u = *reinterpret_cast<const uint32_t*>(cbase + rw * 64 + ((((cp >> 2) ^ (rw >> 1)) & 3) << 4));
              if (ii >= cs_nv[hb]) u = 0u;
              const float2 v2 = make_float2(bf16_lo(u), bf16_hi(u));
              s2[hb] = add_f32x2(s2[hb], v2);
              q2[hb] = fma_f32x2(v2, v2, q2[hb]);
            }
          }
          if (p.cs_direct) {
            // tiles that span several samples (small images): one vector reduction per half-block straight to the sample's sums
#pragma unroll
            for (int hb = 0; hb < 2; ++hb)
              if (cs_nv[hb] > 0 && oc0 + 2 * cp < p.n_out)
                red_add_v4(p.col_accum + cs_hb_off[hb] + (oc0 + 2 * cp) * 2, s2[hb].x, q2[hb].x, s2[hb].y, q2[hb].y);
          } else {
            // one sample per tile: lanes l / l + 16 combined by a shuffle, then lanes 0..15 add their channel pair into THIS
            // WARP's slice of the CTA's table (no atomics: each (warp, column) entry has exactly one writer)
            float4 t4 = make_float4(s2[0].x + s2[1].x, q2[0].x + q2[1].x, s2[0].y + s2[1].y, q2[0].y + q2[1].y);
            t4.x += __shfl_xor_sync(0xffffffffu, t4.x, 16);
            t4.y += __shfl_xor_sync(0xffffffffu, t4.y, 16);
            t4.z += __shfl_xor_sync(0xffffffffu, t4.z, 16);
            t4.w += __shfl_xor_sync(0xffffffffu, t4.w, 16);
            if (lane < 16) {
              float4* tab = reinterpret_cast<float4*>(cs_tab + (lg * BN + c * 32 + 2 * cp) * 2);
              float4 a4 = *tab;
              a4.x += t4.x, a4.y += t4.y, a4.z += t4.z, a4.w += t4.w;
              *tab = a4;
            }
          }
        }
        ++q;
      }
      if (!arrived) {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
            if (PAIR)
              mbar_arrive_cluster(smem_u32(&tempty_bar[acc]) & kPeerBitMask);
            else
              mbar_arrive(&tempty_bar[acc]);
          }
      }
      if (LN == 2) {
        const int64_t m = int64_t(o[0]) + r;
        if (m < p.o_size[0] && g < chunks_per_tile && n_base + g * acc_cw < p.n_rows_b) red_add_v2(p.row_accum + 2 * m, acc_sum, acc_sq);
      }
      if (LN == 3 && !p.cs_direct) {
        // flush this warpgroup's columns (chunks c = g, g + 2, ...): sum the four warps' slices, one vector reduction per channel
        // pair to the sample's global sums, clear; the clears are ordered before the next tile's updates by its staging barriers
        named_bar_sync(1 + g, 128);
        const int c = g + kEpiWGs * (r >> 4);
        if (c < chunks_per_tile && n_base + c * 32 < p.n_rows_b) {
          float4 tv = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
          for (int w = 0; w < 4; ++w) {
            float4* t4 = reinterpret_cast<float4*>(cs_tab + (w * BN + c * 32 + 2 * (r & 15)) * 2);
            const float4 x = *t4;
            *t4 = make_float4(0.f, 0.f, 0.f, 0.f);
            tv.x += x.x, tv.y += x.y, tv.z += x.z, tv.w += x.w;
          }
          if (cs_off >= 0) red_add_v4(p.col_accum + cs_off + (n_base + c * 32 + 2 * (r & 15)) * 2, tv.x, tv.y, tv.z, tv.w);
        }
      }
    }
    if (lead_warp && elect_one()) bulk_wait_group_read<0>();
  } else if (warp >= 4) {
    // ------------------------------------------------------------ epilogue (8 warps)
    const int ew = warp - 4;
    const int lg = ew & 3;     // TMEM lane group == warp % 4
    const int half = ew >> 2;  // which interleaved set of 32-column chunks (one per warpgroup)
    const int r = lg * 32 + lane;
    const bool geglu = (p.flags & T2V_EPI_GEGLU) != 0;
    const bool out_f32 = (p.flags & T2V_EPI_OUT_F32) != 0;
    const bool gelu = (p.flags & T2V_EPI_GELU) != 0;
    const bool split = p.split_k > 1;
    int it = 0;
    for (int tile = tile0; tile < p.num_tiles; tile += tile_step, ++it) {
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      const int ks = tile / p.n_tiles_mn;
      const int mn = tile - ks * p.n_tiles_mn;
      const int n_tile = mn % p.n_tiles_n;
      int m_tile = (mn / p.n_tiles_n) * (PAIR ? 2 : 1) + int(cta_rank);
      // row -> point
      bool valid = r < p.rows_in_box;
      int64_t out_off = 0, res_off = 0, point = 0, pmul = 1;
      int64_t bias_row = 0;
      {
        int rr = r;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int oj = (j < 3 ? m_tile % p.ntile[j] : m_tile) * p.box[j];
          m_tile /= p.ntile[j];
          const int ij = rr % p.box[j];
          rr /= p.box[j];
          const int64_t x = oj + ij;
          valid = valid && (x < p.o_size[j]);
          out_off += x * p.o_stride[j];
          res_off += x * p.r_stride[j];
          point += x * pmul;
          pmul *= p.o_size[j];
          if (j == p.bias_dim) bias_row = x / p.bias_div;
        }
      }
      const float* bias = p.bias ? p.bias + bias_row * p.bias_row_stride : nullptr;

      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + acc * Cfg::kAccStride + (static_cast<uint32_t>(lg * 32) << 16);

#pragma unroll 1
      for (int c0 = half * 32; c0 < BN; c0 += 32 * kEpiWGs) {
        uint32_t v[32];
        tmem_ld_32x32(taddr + c0, v);
        tmem_wait_ld();
        const int n0 = n_tile * BN + c0;  // column in W-row space
        if (n0 < p.n_rows_b) {            // warp-uniform
          const bool full_in = (n0 + 32 <= p.n_rows_b);
          float f[32];
          if (split) {
            // fp32 partial sums -> workspace (the epilogue is applied by the finalize kernel)
            if (valid) {
              float* wp = p.ws + point * p.n_rows_b + n0;
              if (full_in && (p.n_rows_b & 3) == 0) {
#pragma unroll
                for (int q = 0; q < 8; ++q)
                  red_add_v4(wp + 4 * q, p.alpha * __uint_as_float(v[4 * q]), p.alpha * __uint_as_float(v[4 * q + 1]),
                             p.alpha * __uint_as_float(v[4 * q + 2]), p.alpha * __uint_as_float(v[4 * q + 3]));
              } else {
#pragma unroll
                for (int j = 0; j < 32; ++j)
                  if (n0 + j < p.n_rows_b) atomicAdd(wp + j, p.alpha * __uint_as_float(v[j]));
              }
            }
          } else {
            if (bias != nullptr && full_in && p.bias_vec_ok) {
              const float4* b4 = reinterpret_cast<const float4*>(bias + n0);
#pragma unroll
              for (int q = 0; q < 8; ++q) {
                const float4 bv = __ldg(b4 + q);
                f[4 * q + 0] = fmaf(p.alpha, __uint_as_float(v[4 * q + 0]), bv.x);
                f[4 * q + 1] = fmaf(p.alpha, __uint_as_float(v[4 * q + 1]), bv.y);
                f[4 * q + 2] = fmaf(p.alpha, __uint_as_float(v[4 * q + 2]), bv.z);
                f[4 * q + 3] = fmaf(p.alpha, __uint_as_float(v[4 * q + 3]), bv.w);
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                f[j] = p.alpha * __uint_as_float(v[j]);
                if (bias != nullptr && n0 + j < p.n_rows_b) f[j] += __ldg(bias + n0 + j);
              }
            }
            int ncols, oc0;
            if (geglu) {
              // packed rows: 16 value columns followed by their 16 gate columns
#pragma unroll
              for (int j = 0; j < 16; ++j) f[j] = f[j] * gelu_erf(f[j + 16]);
              ncols = 16;
              oc0 = n0 >> 1;
            } else {
              if (gelu) {
#pragma unroll
                for (int j = 0; j < 32; ++j) f[j] = gelu_erf(f[j]);
              }
              ncols = 32;
              oc0 = n0;
            }
            const bool full_chunk = (oc0 + ncols <= p.n_out);
            if (!valid) {
              // row outside the output grid: nothing to store
            } else if (p.vec_ok && full_chunk) {
              if (p.residual != nullptr) {
                const uint4* rp = reinterpret_cast<const uint4*>(
                    reinterpret_cast<const __nv_bfloat16*>(p.residual) + res_off + oc0);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                  if (q * 8 < ncols) {
                    const uint4 rv = __ldg(rp + q);
                    f[q * 8 + 0] += bf16_lo(rv.x);
                    f[q * 8 + 1] += bf16_hi(rv.x);
                    f[q * 8 + 2] += bf16_lo(rv.y);
                    f[q * 8 + 3] += bf16_hi(rv.y);
                    f[q * 8 + 4] += bf16_lo(rv.z);
                    f[q * 8 + 5] += bf16_hi(rv.z);
                    f[q * 8 + 6] += bf16_lo(rv.w);
                    f[q * 8 + 7] += bf16_hi(rv.w);
                  }
                }
              }
              if (out_f32) {
                float4* op = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + out_off + oc0);
#pragma unroll
                for (int q = 0; q < 8; ++q)
                  if (q * 4 < ncols) op[q] = make_float4(f[q * 4], f[q * 4 + 1], f[q * 4 + 2], f[q * 4 + 3]);
              } else {
                uint4* op = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.out) + out_off + oc0);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                  if (q * 8 < ncols) {
                    uint4 ov;
                    ov.x = pack_bf16(f[q * 8 + 0], f[q * 8 + 1]);
                    ov.y = pack_bf16(f[q * 8 + 2], f[q * 8 + 3]);
                    ov.z = pack_bf16(f[q * 8 + 4], f[q * 8 + 5]);
                    ov.w = pack_bf16(f[q * 8 + 6], f[q * 8 + 7]);
                    op[q] = ov;
                  }
                }
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                if (j < ncols && oc0 + j < p.n_out) {
                  float val = f[j];
                  if (p.residual != nullptr)
                    val += __bfloat162float(
                        reinterpret_cast<const __nv_bfloat16*>(p.residual)[res_off + oc0 + j]);
                  if (out_f32)
                    reinterpret_cast<float*>(p.out)[out_off + oc0 + j] = val;
                  else
                    reinterpret_cast<__nv_bfloat16*>(p.out)[out_off + oc0 + j] = __float2bfloat16_rn(val);
                }
              }
            }
          }
        }
      }
      // accumulator drained: hand the TMEM buffer back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
            if (PAIR)
              mbar_arrive_cluster(smem_u32(&tempty_bar[acc]) & kPeerBitMask);
            else
              mbar_arrive(&tempty_bar[acc]);
          }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();  // the peer may still be reading operands / signalling barriers in this CTA
  if (warp == 2) {
    tc_fence_after();
    if (PAIR)
      tmem_dealloc_2sm(tmem_base, Cfg::kTmemCols);
    else
      tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

__global__ void __launch_bounds__(256) zero_f32_kernel(float4* p, int64_t n4) {
  pdl_launch_dependents();
  pdl_wait();
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += int64_t(gridDim.x) * blockDim.x)
    p[i] = make_float4(0.f, 0.f, 0.f, 0.f);
}

// split-K finalize: out[point, n] = epi(ws[point, n]) ; one thread = 4 consecutive columns
__global__ void __launch_bounds__(256) gemm_finalize_kernel(const GemmParams p, int64_t n_points) {
  pdl_launch_dependents();
  pdl_wait();
  const int nq = (p.n_rows_b + 3) >> 2;
  const int64_t total = n_points * nq;
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += int64_t(gridDim.x) * blockDim.x) {
    const int64_t point = i / nq;
    const int n0 = int(i - point * nq) * 4;
    int64_t rem = point, out_off = 0, res_off = 0, bias_row = 0, smp = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int64_t x = rem % p.o_size[j];
      rem /= p.o_size[j];
      out_off += x * p.o_stride[j];
      res_off += x * p.r_stride[j];
      smp += x * p.cs_mult[j];
      if (j == p.bias_dim) bias_row = x / p.bias_div;
    }
    splitk_finalize_quad(p, p.ws + point * p.n_rows_b + n0, point, out_off, res_off, bias_row, smp, n0, (p.flags & T2V_WS_CLEAN) != 0);
  }
}

template <int BN, bool PAIR, int LN = 0>
static int launch_gemm(const CUtensorMap& a0, const CUtensorMap& a1, const CUtensorMap& b, const CUtensorMap& to,
                       const CUtensorMap& tr, const GemmParams& p, cudaStream_t stream) {
  using Cfg = GemmCfg<BN, PAIR>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tc_kernel<BN, PAIR, LN>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit);
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(gemm_tc)");
    configured = true;
  }
  int sms = num_sms();
  if (sms <= 0) return fail(-110, "no CUDA device");
  int grid;
  if (PAIR) {
    const int pairs = p.num_tiles < sms / 2 ? p.num_tiles : sms / 2;
    grid = 2 * pairs;
  } else {
    grid = p.num_tiles < sms ? p.num_tiles : sms;
  }
  GemmParams pp = p;
  int stages = Cfg::kStages;
  if (stages > kMaxStages) stages = kMaxStages;
  if (p.n_stages > 0 && p.n_stages < stages) stages = p.n_stages;
  if (stages < 2) return fail(-111, "gemm_tc: shared memory layout leaves %d pipeline stages", stages);
  constexpr size_t kCsTabBytes = (LN == 3) ? size_t(4 * BN * 2 * 4 + 16) : 0;  // per-warp GroupNorm-statistics tables
  while (stages > 2 && size_t(stages) * Cfg::kStageBytes + kEpiBytes + Cfg::kBarBytes + 1024 + kCsTabBytes > size_t(kSmemLimit)) --stages;
  pp.n_stages = stages;
  pp.smem_epi_off = uint32_t(stages) * uint32_t(Cfg::kStageBytes);
  const size_t smem_bytes = size_t(pp.smem_epi_off) + kEpiBytes + Cfg::kBarBytes + 1024 + kCsTabBytes;
  if (smem_bytes > size_t(kSmemLimit)) return fail(-112, "gemm_tc: %zu bytes of shared memory needed", smem_bytes);
  launch_kernel_cluster(gemm_tc_kernel<BN, PAIR, LN>, dim3(grid), dim3(kThreads), smem_bytes, stream, PAIR ? 2u : 1u, a0, a1, b,
                        to, tr, pp);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "gemm_tc launch");
  return 0;
}

static int choose_block_n(int64_t n_rows, int64_t m_tiles, int sms, bool geglu, bool wide_only, int64_t total_kb) {
  const int cands[5] = {256, 160, 128, 64, 32};
  double best = -1.0;
  int best_bn = 128;
  for (int i = 0; i < 5; ++i) {
    const int bn = cands[i];
    if (geglu && bn % 64 != 0) continue;
    if (wide_only && bn < 128) continue;  // the LayerNorm-aware kernels are only instantiated for wide tiles  // value/gate pairs live in 64-column accumulator groups (staged epilogue)
    const int64_t nt = (n_rows + bn - 1) / bn;
    const double eff_n = double(n_rows) / double(nt * bn);
    const int64_t tiles = nt * m_tiles;
    const int64_t waves = (tiles + sms - 1) / sms;
    const double eff_w = tiles >= sms ? double(tiles) / double(waves * sms) : 1.0;  // small grids use split-K
    // wide tiles halve the shared-memory / L2 operand traffic per flop
    // (measured on the UNet's shapes, scripts/gemm_bench.py with BN=...: N = 960 prefers 256-wide tiles over 160)
    const double eff_t = bn >= 256 ? 1.0 : bn >= 160 ? 0.88 : bn >= 128 ? 0.86 : bn >= 64 ? 0.70 : 0.45;
    const double score = eff_n * eff_w * eff_t;
    if (score > best) {
      best = score;
      best_bn = bn;
    }
  }
  // An under-filled single wave of 256-wide tiles (level 2: 20 x 5 = 100 tiles on 148 SMs) loses to 128-wide CTA pairs
  // when K is long enough for the pair mode to pay (measured: conv 2560x1280x11520 72 -> 63.5 us, ff.net.2
  // 2560x1280x5120 37.0 -> 32.8 us, (3,1,1) conv 27.5 -> 25.5 us; K = 1280 prefers the wide tiles).
  if (best_bn == 256 && !geglu && total_kb >= 48 && m_tiles >= 2) {
    const int64_t t256 = m_tiles * ((n_rows + 255) / 256);
    const int64_t pair128 = ((m_tiles + 1) / 2) * ((n_rows + 127) / 128);
    if (t256 < sms && pair128 >= sms / 2 && n_rows % 128 == 0) best_bn = 128;
  }
  return best_bn;
}

}  // namespace t2v

extern "C" int64_t t2v_gemm_workspace_bytes(const T2VGemmDesc* d) {
  if (!d) return -1;
  int64_t n_points = 1;
  for (int j = 0; j < T2V_MAX_DIMS; ++j) {
    if (d->o_size[j] < 1) return -1;
    n_points *= d->o_size[j];
  }
  if (d->b_rows < 1) return -1;
  return ((n_points * d->b_rows * 4 + 15) / 16) * 16;   // fp32 [points][N] partial sums of a split-K launch
}

extern "C" int64_t t2v_groupnorm_workspace_bytes(int64_t n_samples, int32_t groups, int32_t backward) {
  if (n_samples < 1 || groups < 1) return -1;
  // forward: (sum, sum of squares) per (sample, group) + 1 ticket word; backward: (sum x, sum x^2, S1, S2)
  return backward ? n_samples * groups * 4 * 4 : (n_samples * groups * 2 + 1) * 4;
}

extern "C" int t2v_gemm(const T2VGemmDesc* d, t2v_stream_t stream_) {
  using namespace t2v;
  if (!d) return fail(-1, "t2v_gemm: null descriptor");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!d->a[0] || !d->b || !d->out) return fail(-2, "t2v_gemm: null a/b/out pointer");
  if (d->n_taps < 1 || d->n_taps > T2V_MAX_TAPS) return fail(-3, "t2v_gemm: n_taps=%d", d->n_taps);
  if (d->a_ch[0] <= 0 || d->a_ch[0] % 64 != 0 || d->a_ch[1] < 0 || d->a_ch[1] % 64 != 0)
    return fail(-4, "t2v_gemm: a_ch must be positive multiples of 64 (got %d, %d)", d->a_ch[0], d->a_ch[1]);
  if (d->a_ch[1] > 0 && !d->a[1]) return fail(-5, "t2v_gemm: a[1] is null but a_ch[1]=%d", d->a_ch[1]);
  int64_t rows_in_box = 1, n_points = 1;
  for (int j = 0; j < 4; ++j) {
    if (d->box[j] < 1 || d->a_size[j] < 1 || d->o_size[j] < 1)
      return fail(-6, "t2v_gemm: box/a_size/o_size[%d] must be >= 1", j);
    rows_in_box *= d->box[j];
    n_points *= d->o_size[j];
  }
  if (rows_in_box > 128 || rows_in_box % 8 != 0)
    return fail(-7, "t2v_gemm: box product %lld must be a multiple of 8 and <= 128", (long long)rows_in_box);
  if (d->b_rows < 1) return fail(-8, "t2v_gemm: b_rows=%lld", (long long)d->b_rows);
  const bool geglu = (d->flags & T2V_EPI_GEGLU) != 0;
  if (geglu && (d->b_rows % 32 != 0 || d->n_out != d->b_rows / 2))
    return fail(-9, "t2v_gemm: GEGLU needs b_rows %% 32 == 0 and n_out == b_rows/2");
  if (!geglu && d->n_out != d->b_rows) return fail(-10, "t2v_gemm: n_out must equal b_rows");
  if (d->b_batch_dim > 3 || d->bias_dim > 3) return fail(-11, "t2v_gemm: bad dim index");
  if (d->bias && d->bias_dim >= 0 && d->bias_div < 1) return fail(-12, "t2v_gemm: bias_div must be >= 1");

  const int c_in = d->a_ch[0] + d->a_ch[1];
  const int64_t K = int64_t(d->n_taps) * c_in;

  GemmParams p;
  memset(&p, 0, sizeof(p));
  int64_t m_tiles = 1;
  for (int j = 0; j < 4; ++j) {
    p.box[j] = d->box[j];
    p.ntile[j] = int((d->o_size[j] + d->box[j] - 1) / d->box[j]);
    m_tiles *= p.ntile[j];
    p.o_size[j] = d->o_size[j];
    p.o_stride[j] = d->o_stride[j];
    p.r_stride[j] = d->r_stride[j];
  }
  int sms = num_sms();
  if (sms <= 0) return fail(-110, "t2v_gemm: no CUDA device");
  int bn = d->block_n;
  const int ln_mode = d->row_stats ? 1 : (d->row_accum ? 2 : (d->col_accum ? 3 : 0));
  bool narrow_no_split = false;
  if (bn == 0) {
    bn = choose_block_n(d->b_rows, m_tiles, sms, geglu, ln_mode != 0, K / 64);
    // Small-M layers (level 3: M = 640): when wide tiles cannot fill the SMs, 64-wide tiles over the full K beat
    // split-K (one launch instead of zero + GEMM + finalize) as long as K is short: Linear 640x1280x1280 8.0 us vs
    // 19.0 us, (3,1,1) conv 15.0 vs 20.6 us; 3x3 convs (K >= 11520) stay on split-K (30 vs 37 us).
    const int64_t tiles_wide = m_tiles * ((d->b_rows + bn - 1) / bn);
    if (d->split_k == 0 && !geglu && ln_mode != 3 && d->b_batch_dim < 0 && tiles_wide * 2 <= sms && K / 64 >= 16 && K / 64 <= 64 &&
        d->b_rows >= 64) {
      bn = 64;
      narrow_no_split = true;
    }
  }
  if (ln_mode != 0 && bn < 128 && !(bn == 64 && ln_mode != 3))
    return fail(-19, "t2v_gemm: LayerNorm-aware GEMMs need block_n >= 128 (or 64 for the row-statistics modes)");
  if (bn != 32 && bn != 64 && bn != 128 && bn != 160 && bn != 256) return fail(-13, "t2v_gemm: block_n=%d unsupported", bn);
  p.n_tiles_n = int((d->b_rows + bn - 1) / bn);
  const int64_t tiles_mn = m_tiles * p.n_tiles_n;
  p.total_kb = int(K / 64);
  // split-K: explicit, or automatic when the tile grid cannot fill the SMs
  int split = d->split_k;
  if (split < 0) return fail(-15, "t2v_gemm: split_k must be >= 0");
  if (split == 0) {
    split = 1;
    if (!narrow_no_split && !geglu && d->b_batch_dim < 0 && d->workspace && tiles_mn * 2 <= sms && p.total_kb >= 16 &&
        d->workspace_bytes >= n_points * d->b_rows * 4) {
      split = int(sms / tiles_mn);
      if (split > p.total_kb / 8) split = p.total_kb / 8;
      if (split > 32) split = 32;
      if (split < 1) split = 1;
    }
  }
  if (split > 1) {
    if (geglu) return fail(-16, "t2v_gemm: split_k is not supported with GEGLU");
    if (d->b_batch_dim >= 0) return fail(-17, "t2v_gemm: split_k is not supported for batched B");
    const int64_t need = n_points * d->b_rows * 4;
    if (!d->workspace || d->workspace_bytes < need)
      return fail(-18, "t2v_gemm: split_k needs a %lld-byte fp32 workspace", (long long)need);
    p.kb_per_split = (p.total_kb + split - 1) / split;
    split = (p.total_kb + p.kb_per_split - 1) / p.kb_per_split;  // drop empty slices
  } else {
    p.kb_per_split = p.total_kb;
  }
  p.split_k = split;
  const int64_t num_tiles = tiles_mn * split;
  if (num_tiles > 0x7fffffff) return fail(-14, "t2v_gemm: too many tiles");
  p.n_tiles_mn = int(tiles_mn);
  p.num_tiles = int(num_tiles);
  p.rows_in_box = int(rows_in_box);
  p.n_taps = d->n_taps;
  for (int t = 0; t < d->n_taps; ++t) {
    for (int j = 0; j < 4; ++j) p.tap_off[t][j] = d->tap_off[t][j];
    p.tap_ch_off[t] = d->tap_ch_off[t];
  }
  p.kb_per_tap = c_in / 64;
  p.kb_src0 = d->a_ch[0] / 64;
  p.b_batch_dim = d->b_batch_dim;
  p.n_rows_b = int(d->b_rows);
  p.n_out = d->n_out;
  p.out = d->out;
  p.residual = d->residual;
  p.bias = d->bias;
  p.bias_row_stride = d->bias_row_stride;
  p.bias_dim = d->bias ? d->bias_dim : -1;
  p.bias_div = d->bias_div < 1 ? 1 : d->bias_div;
  p.alpha = d->alpha;
  p.flags = d->flags;
  p.a_tile_bytes = uint32_t(rows_in_box * 128);
  p.ws = static_cast<float*>(d->workspace);
  p.row_stats = reinterpret_cast<const float2*>(d->row_stats);
  p.col_sum = d->col_sum;
  p.ln_raw = d->ln_raw;
  p.ln_inv_c = d->ln_channels > 0 ? 1.0f / float(d->ln_channels) : 0.f;
  p.ln_eps = d->ln_eps;
  p.row_accum = d->row_accum;
  p.col_accum = d->col_accum;
  for (int j = 0; j < 4; ++j) p.cs_mult[j] = d->cs_mult[j];
  if (d->col_accum) {
    if (d->row_stats || d->row_accum) return fail(-20, "t2v_gemm: col_accum excludes row_stats / row_accum");
    if (geglu || (d->flags & T2V_EPI_OUT_F32) || d->b_rows % 32 || d->n_out != d->b_rows || d->box[0] % 8 || d->cs_mult[0] != 0 ||
        d->b_batch_dim >= 0 || (reinterpret_cast<uintptr_t>(d->col_accum) & 15))
      return fail(-20, "t2v_gemm: col_accum needs bf16 output, N %% 32 == 0, box[0] %% 8 == 0, cs_mult[0] == 0, no batched B");
    for (int j = 1; j < 4; ++j)
      if (d->cs_mult[j] != 0 && d->box[j] != 1) p.cs_direct = 1;  // a tile spans samples (small images)
  }
  if (d->row_stats && d->row_accum) return fail(-19, "t2v_gemm: row_stats and row_accum are mutually exclusive");
  if (d->row_stats && d->ln_raw && d->ln_channels <= 0) return fail(-19, "t2v_gemm: ln_raw needs ln_channels");
  if (d->row_accum) {
    if (d->a_size[1] != 1 || d->a_size[2] != 1 || d->a_size[3] != 1 || d->n_taps != 1 || geglu || (d->flags & T2V_EPI_OUT_F32) ||
        d->b_rows % 32 || (reinterpret_cast<uintptr_t>(d->row_accum) & 7))
      return fail(-19, "t2v_gemm: row_accum needs a plain [M, K] x [N, K] GEMM with N %% 32 == 0 and bf16 output");
  }
  if (d->row_stats) {
    if (!d->col_sum) return fail(-19, "t2v_gemm: row_stats needs col_sum");
    if (d->a_size[1] != 1 || d->a_size[2] != 1 || d->a_size[3] != 1 || d->n_taps != 1 || d->alpha != 1.0f ||
        d->bias_dim >= 0 || (d->flags & T2V_EPI_OUT_F32) || (reinterpret_cast<uintptr_t>(d->row_stats) & 7))
      return fail(-19, "t2v_gemm: a folded LayerNorm needs a plain [M, K] x [N, K] GEMM, alpha = 1, [N] bias, bf16 output");
  }
  // vector epilogue: 16-byte aligned rows
  const int out_el = (d->flags & T2V_EPI_OUT_F32) ? 4 : 2;
  bool vec_ok = (d->n_out % 8 == 0) && ((reinterpret_cast<uintptr_t>(d->out) & 15) == 0);
  for (int j = 0; j < 4; ++j) vec_ok = vec_ok && ((d->o_stride[j] * out_el) % 16 == 0);
  if (d->residual) {
    vec_ok = vec_ok && ((reinterpret_cast<uintptr_t>(d->residual) & 15) == 0);
    for (int j = 0; j < 4; ++j) vec_ok = vec_ok && (d->r_stride[j] % 8 == 0);
  }
  p.vec_ok = vec_ok ? 1 : 0;
  p.bias_vec_ok = (d->bias && (reinterpret_cast<uintptr_t>(d->bias) & 15) == 0 && d->bias_row_stride % 4 == 0) ? 1 : 0;

  // staged (TMA-store) epilogue whenever the output rows are 16-byte addressable bf16
  bool staged = vec_ok && split == 1 && !(d->flags & T2V_EPI_OUT_F32) && (!geglu || bn % 64 == 0);
  if (ln_mode != 0 && split == 1 && !staged)
    return fail(-19, "t2v_gemm: LayerNorm-aware GEMMs need 16-byte aligned bf16 output rows");
  p.epi_mode = staged ? 1 : 0;
  // CTA pairs (cta_group::2): 256-row tiles, each CTA stages half of the weight tile
  const int64_t pair_tiles = ((m_tiles + 1) / 2) * p.n_tiles_n;
  // measured: +10 % at K >= 1024 (1441 TFLOP/s on 512->512 3x3 convs = the cuBLAS sustained level), a loss at small K
  const bool pair = staged && !(d->tune & 0x800) && d->b_batch_dim < 0 && bn >= 128 && m_tiles >= 2 &&
                    pair_tiles >= sms / 2 && p.total_kb >= (((d->tune >> 16) & 0xff) ? ((d->tune >> 16) & 0xff) : 15);
  if (pair) {
    p.n_tiles_mn = int(pair_tiles);
    p.num_tiles = int(pair_tiles);
  }
  p.n_stages = d->tune & 0xff;  // 0 = all stages of the instantiation
  p.dbg = (d->tune >> 12) & 0xf;  // bit0: no TMA store, bit1: no smem write/fence/barrier, bit2: no TMEM loads

  // tensor maps
  CUtensorMap tmA[2], tmB, tmOut, tmRes;
  memset(&tmOut, 0, sizeof(tmOut));
  memset(&tmRes, 0, sizeof(tmRes));
  if (staged) {
    for (int which = 0; which < (d->residual ? 2 : 1); ++which) {
      uint64_t dims[5], strides[5];
      uint32_t box[5];
      dims[0] = uint64_t(d->n_out);
      strides[0] = 2;
      box[0] = 32;
      for (int j = 0; j < 4; ++j) {
        dims[j + 1] = uint64_t(d->o_size[j]);
        strides[j + 1] = uint64_t(which == 0 ? d->o_stride[j] : d->r_stride[j]) * 2;
        box[j + 1] = uint32_t(d->box[j]);
        if (d->o_size[j] == 1 && strides[j + 1] == 0) strides[j + 1] = 16;
      }
      int rc = make_tmap_bf16(which == 0 ? &tmOut : &tmRes, which == 0 ? d->out : d->residual, 5, dims, strides, box,
                              which == 0 ? "t2v_gemm out" : "t2v_gemm residual", 64);
      if (rc) return rc;
    }
  }
  for (int s = 0; s < 2; ++s) {
    const int src = (s == 1 && d->a_ch[1] == 0) ? 0 : s;
    uint64_t dims[5], strides[5];
    uint32_t box[5];
    dims[0] = uint64_t(d->a_ch_total[src] > 0 ? d->a_ch_total[src] : d->a_ch[src]);
    strides[0] = 2;
    box[0] = 64;
    for (int j = 0; j < 4; ++j) {
      dims[j + 1] = uint64_t(d->a_size[j]);
      strides[j + 1] = uint64_t(d->a_stride[src][j]) * 2;
      box[j + 1] = uint32_t(d->box[j]);
      if (d->a_size[j] == 1 && strides[j + 1] == 0) strides[j + 1] = 16;  // size-1 dim: any legal stride
    }
    int rc = make_tmap_bf16(&tmA[s], d->a[src], 5, dims, strides, box, s == 0 ? "t2v_gemm A0" : "t2v_gemm A1");
    if (rc) return rc;
  }
  {
    const uint64_t row_stride = uint64_t(d->b_row_stride > 0 ? d->b_row_stride : K);
    uint64_t dims[3] = {uint64_t(K), uint64_t(d->b_rows), uint64_t(d->b_batches < 1 ? 1 : d->b_batches)};
    uint64_t strides[3] = {2, row_stride * 2,
                           uint64_t(d->b_batch_stride > 0 ? d->b_batch_stride : row_stride * d->b_rows) * 2};
    uint32_t box[3] = {64, uint32_t(pair ? bn / 2 : bn), 1};
    int rc = make_tmap_bf16(&tmB, d->b, 3, dims, strides, box, "t2v_gemm B");
    if (rc) return rc;
  }
  // T2V_WS_CLEAN: the workspace is zero on entry and the finalize kernel clears what it reads, so the zero kernel is
  // skipped (tune bit 0x1000000 forces it, for A/B measurements)
  if (split > 1 && (!(d->flags & T2V_WS_CLEAN) || (d->tune & 0x1000000))) {
    const int64_t n4 = (n_points * d->b_rows + 3) / 4;  // workspace is a multiple of 16 bytes
    int64_t zg = (n4 + 255) / 256;
    if (zg > sms * 4) zg = sms * 4;
    launch_kernel(zero_f32_kernel, dim3(unsigned(zg)), dim3(256), 0, stream, static_cast<float4*>(d->workspace), n4);
  }
  int rc;
#define T2V_GEMM_CASE(BN_, PAIR_, LN_) rc = launch_gemm<BN_, PAIR_, LN_>(tmA[0], tmA[1], tmB, tmOut, tmRes, p, stream)
#define T2V_GEMM_WIDE(LN_)                                                       \
  do {                                                                           \
    if (pair) {                                                                  \
      if (bn == 128) T2V_GEMM_CASE(128, true, LN_);                              \
      else if (bn == 160) T2V_GEMM_CASE(160, true, LN_);                         \
      else T2V_GEMM_CASE(256, true, LN_);                                        \
    } else {                                                                     \
      if (bn == 128) T2V_GEMM_CASE(128, false, LN_);                             \
      else if (bn == 160) T2V_GEMM_CASE(160, false, LN_);                        \
      else T2V_GEMM_CASE(256, false, LN_);                                       \
    }                                                                            \
  } while (0)
  if (ln_mode == 1 && bn == 64) {
    T2V_GEMM_CASE(64, false, 1);
  } else if (ln_mode == 2 && bn == 64) {
    T2V_GEMM_CASE(64, false, 2);
  } else if (ln_mode == 1) {
    T2V_GEMM_WIDE(1);
  } else if (ln_mode == 2) {
    T2V_GEMM_WIDE(2);
  } else if (ln_mode == 3) {
    T2V_GEMM_WIDE(3);
  } else if (bn == 32) {
    T2V_GEMM_CASE(32, false, 0);
  } else if (bn == 64) {
    T2V_GEMM_CASE(64, false, 0);
  } else {
    T2V_GEMM_WIDE(0);
  }
#undef T2V_GEMM_WIDE
#undef T2V_GEMM_CASE
  if (rc) return rc;
  if (split > 1) {
    const int64_t total = n_points * ((d->b_rows + 3) / 4);
    int64_t grid = (total + 255) / 256;
    if (grid > sms * 8) grid = sms * 8;
    launch_kernel(gemm_finalize_kernel, dim3(unsigned(grid)), dim3(256), 0, stream, p, n_points);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "gemm_finalize launch");
  }
  return 0;
}
