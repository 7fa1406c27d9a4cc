// Fused attention forward, head_dim 64, no mask, on tcgen05 (sm_100a).
//
//   O[b,i,h,:] = softmax_j(scale * Q[b,i,h,:] . K[b',j,h,:]) V[b',j,h,:],   b' = b / kv_batch_div
//
// One CTA = one (batch, head, 128-query tile); two CTAs are co-resident per SM (256 TMEM columns
// and ~112 KB shared memory each) so one CTA's softmax overlaps the other's MMAs.
//   warp 0 / lane 0 : TMA producer — Q tile once, K / V tiles (128 keys) through 2-stage rings.
//                     Q/K/V are read in place from the projection outputs via 4-D tensor maps
//                     {64, head, token, batch}: no head-split copy.
//   warp 1 / lane 0 : MMA issuer — S = Q K^T (M128 N128 K64) into TMEM, then O += P V
//                     (M128 N64 K128) with P from shared memory and V as an MN-major operand.
//   warp 2          : TMEM allocator.
//   warps 4..7      : online softmax, thread = query row: tcgen05.ld S, running max / sum in
//                     fp32 (exp2 with the scale folded in), P -> bf16 -> 128B-swizzled smem,
//                     lazy rescale of the O accumulator in TMEM, final 1/l scaling and store.
#include <cuda.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/t2v_b200.h"
#include "host_common.h"
#include "ptx.cuh"

namespace t2v {

constexpr int kAtThreads = 256;
constexpr int kKvStages = 2;
constexpr int kTile = 128;
constexpr int kQBytes = kTile * 64 * 2;   // 16 KB
constexpr int kKBytes = kTile * 64 * 2;   // 16 KB
constexpr int kPBytes = kTile * kTile * 2;  // 32 KB (two K-major 64-key sub-tiles)
constexpr int kAtSmem = kQBytes + 2 * kKvStages * kKBytes + kPBytes + 256;
constexpr int kTmemColsAttn = 256;
constexpr int kOCol = 128;

struct AttnParams {
  int32_t heads, len_q, len_k, n_q_tiles, n_kv_tiles, kv_batch_div;
  int32_t causal;  // key j visible to query i iff j <= i (CLIP text tower); 0 = no mask
  float scale_log2;
  __nv_bfloat16* o;
  int64_t o_stride_b, o_stride_t, o_stride_h;
  float* lse2;  // optional [batch][heads][len_q]
};

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__global__ void __launch_bounds__(kAtThreads, 2)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                const __grid_constant__ CUtensorMap tmV, const AttnParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();  // 128B swizzle needs 1024-byte aligned tiles
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + kQBytes;
  uint8_t* sV = sK + kKvStages * kKBytes;
  uint8_t* sP = sV + kKvStages * kKBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + kPBytes);
  uint64_t* q_full = bars;                 // 1
  uint64_t* k_full = bars + 1;             // kKvStages
  uint64_t* k_empty = k_full + kKvStages;  // kKvStages
  uint64_t* v_full = k_empty + kKvStages;
  uint64_t* v_empty = v_full + kKvStages;
  uint64_t* s_full = v_empty + kKvStages;  // 1
  uint64_t* p_full = s_full + 1;           // 1 (4 arrivals: one per row warp)
  uint64_t* pv_done = p_full + 1;          // 1
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pv_done + 1);

  const int warp = __shfl_sync(0xffffffffu, int(threadIdx.x >> 5), 0);  // provably warp-uniform: lean TMA / MMA issue code
  const int lane = threadIdx.x & 31;

  const int q_tile = blockIdx.x % p.n_q_tiles;
  const int bh = blockIdx.x / p.n_q_tiles;
  const int h = bh % p.heads;
  const int b = bh / p.heads;
  const int kvb = b / p.kv_batch_div;
  const int q0 = q_tile * kTile;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
  }
  if (warp == 1 && lane == 0) {
    mbar_init(q_full, 1);
    for (int s = 0; s < kKvStages; ++s) {
      mbar_init(&k_full[s], 1);
      mbar_init(&k_empty[s], 1);
      mbar_init(&v_full[s], 1);
      mbar_init(&v_empty[s], 1);
    }
    mbar_init(s_full, 1);
    mbar_init(p_full, 4);   // one arrival per softmax warp
    mbar_init(pv_done, 1);
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, kTmemColsAttn);
    tmem_relinquish();
  }
  pdl_launch_dependents();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();
  const int n_kv = p.n_kv_tiles;

  if (warp == 0 && elect_one()) {
    // ------------------------------------------------------------ TMA producer
    mbar_expect_tx(q_full, kQBytes);
    tma_load_4d(sQ, &tmQ, q_full, 0, h, q0, b);
    for (int j = 0; j < n_kv; ++j) {
      const int s = j % kKvStages;
      const uint32_t ph = (j / kKvStages) & 1;
      mbar_wait(&k_empty[s], ph ^ 1u);
      mbar_expect_tx(&k_full[s], kKBytes);
      tma_load_4d(sK + s * kKBytes, &tmK, &k_full[s], 0, h, j * kTile, kvb);
      mbar_wait(&v_empty[s], ph ^ 1u);
      mbar_expect_tx(&v_full[s], kKBytes);
      tma_load_4d(sV + s * kKBytes, &tmV, &v_full[s], 0, h, j * kTile, kvb);
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    // The whole warp walks the loop (uniform control flow keeps descriptor arithmetic and barrier addresses on the uniform
    // datapath); only tcgen05.mma / commit are predicated on the elected lane.  Inside a single-thread branch the issuer
    // needed ~12 dependent instructions per MMA and its own latency bounded the kernel (attn_fwd2.cu has the measurement).
    const bool leader = elect_one();
    constexpr uint32_t idesc_s = umma_idesc_bf16(128, 128, 0, 0);   // Q (K-major) x K (K-major)
    constexpr uint32_t idesc_o = umma_idesc_bf16(128, 64, 0, 1);    // P (K-major) x V (MN-major)
    const uint32_t tm = __shfl_sync(0xffffffffu, tmem_base, 0);
    const uint64_t qdesc = umma_desc_sw128(smem_u32(sQ));
    const uint64_t pdesc = umma_desc_sw128(smem_u32(sP));
    const uint64_t kdesc0 = umma_desc_sw128(smem_u32(sK));
    const uint64_t vdesc0 = umma_desc_sw128(smem_u32(sV));
    constexpr uint64_t kTileDesc = kKBytes >> 4;   // one K / V stage further, in descriptor (16-byte) units
    mbar_wait(q_full, 0);
    // Issue order per key tile j:  [p_full(j)]  S(j+1)  PV(j).  S(j+1) only needs the S columns (free once softmax(j)
    // has arrived on p_full) and K(j+1), so it is issued BEFORE PV(j): softmax(j+1) starts while PV(j) still runs and
    // the PV latency leaves the per-tile critical path.  (P(j+1) may only be written after PV(j) completed: the
    // softmax warps wait for pv_done(j) right before their first P store.)
    auto issue_s = [&](int sn, uint32_t phn) {
      mbar_wait(&k_full[sn], phn);
      tc_fence_after();
      const uint64_t kdesc = kdesc0 + uint64_t(sn) * kTileDesc;
      if (leader) {
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_ss(tm, qdesc + 2 * k, kdesc + 2 * k, idesc_s, k != 0);
        umma_commit(&k_empty[sn]);
        umma_commit(s_full);
      }
    };
    issue_s(0, 0);
    int s = 0, sn = (kKvStages > 1) ? 1 : 0;            // ring stage of key tile j / j + 1
    uint32_t ph = 0, phn = (kKvStages > 1) ? 0u : 1u;     // and their parities
    for (int j = 0; j < n_kv; ++j) {
      mbar_wait(p_full, j & 1);
      tc_fence_after();
      if (j + 1 < n_kv) issue_s(sn, phn);
      // O += P_j V_j
      mbar_wait(&v_full[s], ph);
      tc_fence_after();
      const uint64_t vdesc = vdesc0 + uint64_t(s) * kTileDesc;
      if (leader) {
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
          // A: 16 keys = 32 bytes inside the (kk/4)-th 64-key sub-tile; B: 16 key rows = 2048 bytes
          const uint64_t ad = pdesc + uint64_t((kk >> 2) * (kTile * 128 >> 4)) + 2 * (kk & 3);
          const uint64_t bd = vdesc + uint64_t(kk * (2048 >> 4));
          umma_ss(tm + kOCol, ad, bd, idesc_o, (j | kk) != 0);
        }
        umma_commit(&v_empty[s]);
        umma_commit(pv_done);
      }
      s = sn;
      ph = phn;
      if (++sn == kKvStages) {
        sn = 0;
        phn ^= 1u;
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------ softmax / correction / epilogue
    const int ew = warp - 4;
    const int r = ew * 32 + lane;
    const uint32_t lane_addr = static_cast<uint32_t>(ew * 32) << 16;
    const uint32_t s_addr = tmem_base + lane_addr;
    const uint32_t o_addr = tmem_base + lane_addr + kOCol;
    uint8_t* p_row = sP + r * 128;
    const int sw = r & 7;
    // Online softmax with a LAZY reference: exponentials are taken against m_ref, which is only raised when a
    // row's scores exceed it by more than 2^8 (then P / l / O are rescaled exactly).  Softmax is shift
    // invariant, so the result is unchanged, but the steady state is ONE pass over S per tile
    // (fma + ex2 + add + max per element) instead of a max pass followed by an exp pass.
    constexpr float kLazy = 8.0f;
    float m_ref = -INFINITY;  // exponent reference (log2 domain, scale folded in)
    float l_run = 0.f;
    for (int j = 0; j < n_kv; ++j) {
      mbar_wait(s_full, j & 1);
      tc_fence_after();
      int kv_left = p.len_k - j * kTile;  // valid keys in this tile (>= 1)
      if (p.causal) kv_left = min(kv_left, q0 + r - j * kTile + 1);  // per query row: keys up to the diagonal (may be <= 0)
      if (j == 0) {
        // first tile: exact row max
        float mx = -INFINITY;
#pragma unroll 1
        for (int c0 = 0; c0 < kTile; c0 += 32) {
          uint32_t v[32];
          tmem_ld_32x32(s_addr + c0, v);
          tmem_wait_ld();
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (c0 + i < kv_left) mx = fmaxf(mx, __uint_as_float(v[i]));
        }
        m_ref = mx * p.scale_log2;
      }
      float rs = 0.f, raw_max = -INFINITY;
      // One pass over the 128 scores of this row: the TMEM load of the next 32-column chunk is in flight while the
      // current chunk is exponentiated and written to the P tile.
      bool p_free = (j == 0);  // the P buffer (and O) may be touched once PV(j-1) has completed
      auto exp_chunk = [&](const uint32_t (&v)[32], int c0, float nref) {
        uint32_t pk[16];
        if (c0 + 32 <= kv_left) {
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            const float s0 = __uint_as_float(v[i]), s1 = __uint_as_float(v[i + 1]);
            raw_max = fmaxf(raw_max, fmaxf(s0, s1));
            const float e0 = ex2(fmaf(s0, p.scale_log2, nref));
            const float e1 = ex2(fmaf(s1, p.scale_log2, nref));
            rs += e0 + e1;
            pk[i >> 1] = pack_bf16(e0, e1);
          }
        } else {
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            const float s0 = __uint_as_float(v[i]), s1 = __uint_as_float(v[i + 1]);
            const bool ok0 = c0 + i < kv_left, ok1 = c0 + i + 1 < kv_left;
            if (ok0) raw_max = fmaxf(raw_max, s0);
            if (ok1) raw_max = fmaxf(raw_max, s1);
            const float e0 = ok0 ? ex2(fmaf(s0, p.scale_log2, nref)) : 0.f;
            const float e1 = ok1 ? ex2(fmaf(s1, p.scale_log2, nref)) : 0.f;
            rs += e0 + e1;
            pk[i >> 1] = pack_bf16(e0, e1);
          }
        }
        if (!p_free) {
          mbar_wait(pv_done, (j - 1) & 1);
          tc_fence_after();
          p_free = true;
        }
        // 32 keys = 4 chunks of 16 bytes; chunk index within the 128-key row: c0/8 + q
        uint8_t* sub = p_row + (c0 >> 6) * (kTile * 128);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int cc = ((c0 & 63) >> 3) + q;
          *reinterpret_cast<uint4*>(sub + ((cc ^ sw) << 4)) = make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
        }
      };
      auto exp_pass = [&](float ref) {
        rs = 0.f;
        raw_max = -INFINITY;
        const float nref = -ref;
        uint32_t va[32], vb[32];
        tmem_ld_32x32(s_addr, va);
        tmem_wait_ld();
        tmem_ld_32x32(s_addr + 32, vb);
        exp_chunk(va, 0, nref);
        tmem_wait_ld();
        tmem_ld_32x32(s_addr + 64, va);
        exp_chunk(vb, 32, nref);
        tmem_wait_ld();
        tmem_ld_32x32(s_addr + 96, vb);
        exp_chunk(va, 64, nref);
        tmem_wait_ld();
        exp_chunk(vb, 96, nref);
      };
      exp_pass(m_ref);
      const float tile_max = raw_max * p.scale_log2;
      if (__any_sync(0xffffffffu, tile_max > m_ref + kLazy)) {
        // rare: raise the reference for the rows that need it, rescale l and O, redo this tile's P
        const float new_ref = fmaxf(m_ref, tile_max);
        const float alpha = ex2(m_ref - new_ref);  // 1 for rows that keep their reference
        l_run *= alpha;
        if (j > 0) {
#pragma unroll
          for (int c0 = 0; c0 < 64; c0 += 16) {
            uint32_t ov[16];
            tmem_ld_32x16(o_addr + c0, ov);
            tmem_wait_ld();
#pragma unroll
            for (int i = 0; i < 16; ++i) ov[i] = __float_as_uint(__uint_as_float(ov[i]) * alpha);
            tmem_st_32x16(o_addr + c0, ov);
          }
          tmem_wait_st();
        }
        m_ref = new_ref;
        exp_pass(m_ref);
      }
      l_run += rs;
      fence_proxy_async();  // P stores (generic proxy) -> visible to the tensor core (async proxy)
      tc_fence_before();
      mbar_arrive_warp(p_full);
    }
    // the last PV must have landed before O is read
    mbar_wait(pv_done, (n_kv - 1) & 1);
    tc_fence_after();
    // epilogue: O / l -> bf16 -> global
    const int qi = q0 + r;
    const float inv_l = 1.0f / l_run;
    if (p.lse2 != nullptr && qi < p.len_q) p.lse2[(int64_t(b) * p.heads + h) * p.len_q + qi] = m_ref + log2f(l_run);
    __nv_bfloat16* orow = p.o + int64_t(b) * p.o_stride_b + int64_t(qi) * p.o_stride_t + int64_t(h) * p.o_stride_h;
#pragma unroll
    for (int c0 = 0; c0 < 64; c0 += 32) {
      uint32_t ov[32];
      tmem_ld_32x32(o_addr + c0, ov);
      tmem_wait_ld();
      if (qi < p.len_q) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint4 pk;
          pk.x = pack_bf16(__uint_as_float(ov[q * 8 + 0]) * inv_l, __uint_as_float(ov[q * 8 + 1]) * inv_l);
          pk.y = pack_bf16(__uint_as_float(ov[q * 8 + 2]) * inv_l, __uint_as_float(ov[q * 8 + 3]) * inv_l);
          pk.z = pack_bf16(__uint_as_float(ov[q * 8 + 4]) * inv_l, __uint_as_float(ov[q * 8 + 5]) * inv_l);
          pk.w = pack_bf16(__uint_as_float(ov[q * 8 + 6]) * inv_l, __uint_as_float(ov[q * 8 + 7]) * inv_l);
          *reinterpret_cast<uint4*>(orow + c0 + q * 8) = pk;
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemColsAttn);
  }
}

// attn_fwd2.cu: the two-Q-tile kernel (P in TMEM, part of the exponentials on the FMA pipe)
int launch_attn_fwd2(const T2VAttnDesc* d, const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, cudaStream_t stream);

}  // namespace t2v

extern "C" int t2v_attn_fwd(const T2VAttnDesc* d, t2v_stream_t stream_) {
  using namespace t2v;
  if (!d || !d->q || !d->k || !d->v || !d->o) return fail(-1, "t2v_attn_fwd: null pointer");
  if (d->batch < 1 || d->heads < 1 || d->len_q < 1 || d->len_k < 1) return fail(-2, "t2v_attn_fwd: bad sizes");
  if (d->kv_batch_div < 1 || d->batch % d->kv_batch_div) return fail(-3, "t2v_attn_fwd: batch %% kv_batch_div != 0");
  if (d->o_stride_b % 8 || d->o_stride_t % 8 || d->o_stride_h % 8 || (reinterpret_cast<uintptr_t>(d->o) & 15))
    return fail(-4, "t2v_attn_fwd: output must be 16-byte aligned with strides multiple of 8");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const int kvb = d->batch / d->kv_batch_div;
  CUtensorMap tq, tk, tv;
  {
    uint64_t dims[4] = {64, uint64_t(d->heads), uint64_t(d->len_q), uint64_t(d->batch)};
    uint64_t str[4] = {2, uint64_t(d->q_stride_h) * 2, uint64_t(d->q_stride_t) * 2, uint64_t(d->q_stride_b) * 2};
    uint32_t box[4] = {64, 1, 128, 1};
    if (d->heads == 1 && str[1] == 0) str[1] = 128;
    if (d->batch == 1 && str[3] == 0) str[3] = str[2] * d->len_q;
    int rc = make_tmap_bf16(&tq, d->q, 4, dims, str, box, "t2v_attn_fwd Q");
    if (rc) return rc;
  }
  {
    uint64_t dims[4] = {64, uint64_t(d->heads), uint64_t(d->len_k), uint64_t(kvb)};
    uint64_t str[4] = {2, uint64_t(d->k_stride_h) * 2, uint64_t(d->k_stride_t) * 2, uint64_t(d->k_stride_b) * 2};
    uint32_t box[4] = {64, 1, 128, 1};
    if (d->heads == 1 && str[1] == 0) str[1] = 128;
    if (kvb == 1 && str[3] == 0) str[3] = str[2] * d->len_k;
    int rc = make_tmap_bf16(&tk, d->k, 4, dims, str, box, "t2v_attn_fwd K");
    if (rc) return rc;
    uint64_t strv[4] = {2, uint64_t(d->v_stride_h) * 2, uint64_t(d->v_stride_t) * 2, uint64_t(d->v_stride_b) * 2};
    if (d->heads == 1 && strv[1] == 0) strv[1] = 128;
    if (kvb == 1 && strv[3] == 0) strv[3] = strv[2] * d->len_k;
    rc = make_tmap_bf16(&tv, d->v, 4, dims, strv, box, "t2v_attn_fwd V");
    if (rc) return rc;
  }
  // Kernel choice.  The two-Q-tile kernel (attn_fwd2.cu: P in TMEM, part of the exp2 on the FMA pipe) wins on long sequences —
  // (16, 2560, 2560, 5): 192 us vs 220 us — and loses 10-30 % on the 640- / 160-token levels and on Lk = 77 (one CTA per SM),
  // so it serves len_q, len_k >= 1024 and this single-tile kernel the rest (profiles/r02_attention.md).  T2V_ATTN_V2=0 / 1
  // forces one of them wherever both apply (read per call: scripts/attn_ablate.py switches inside one process).
  const char* v2_env = getenv("T2V_ATTN_V2");
  const bool v2_ok = !d->causal && !d->lse2;
  const bool use_v2 = v2_env != nullptr && (v2_env[0] == '0' || v2_env[0] == '1') ? v2_env[0] == '1'
                                                                                  : (d->len_q >= 1024 && d->len_k >= 1024);
  if (use_v2 && v2_ok) return launch_attn_fwd2(d, tq, tk, tv, stream);
  AttnParams p;
  p.heads = d->heads;
  p.len_q = d->len_q;
  p.len_k = d->len_k;
  p.n_q_tiles = (d->len_q + kTile - 1) / kTile;
  p.n_kv_tiles = (d->len_k + kTile - 1) / kTile;
  p.kv_batch_div = d->kv_batch_div;
  p.causal = d->causal ? 1 : 0;
  p.scale_log2 = d->scale * 1.4426950408889634f;
  p.o = static_cast<__nv_bfloat16*>(d->o);
  p.o_stride_b = d->o_stride_b;
  p.o_stride_t = d->o_stride_t;
  p.o_stride_h = d->o_stride_h;
  p.lse2 = d->lse2;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kAtSmem);
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(attn_fwd)");
    configured = true;
  }
  const int64_t grid = int64_t(d->batch) * d->heads * p.n_q_tiles;
  if (grid > 0x7fffffff) return fail(-5, "t2v_attn_fwd: grid too large");
  launch_kernel(attn_fwd_kernel, dim3(unsigned(grid)), dim3(kAtThreads), kAtSmem, stream, tq, tk, tv, p);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : cuda_fail(e, "t2v_attn_fwd launch");
}
