// Fused attention forward, head_dim 64, no mask, on tcgen05 (sm_100a) — the two-Q-tile ("ping-pong") kernel.
//
//   O[b,i,h,:] = softmax_j(scale * Q[b,i,h,:] . K[b',j,h,:]) V[b',j,h,:],   b' = b / kv_batch_div
//
// At d = 64 one 128 x 128 score tile costs 512 tensor-pipe cycles (S = Q K^T and O += P V) but 16384 exponentials, i.e.
// 1024 cycles of one SM's MUFU pipes: a kernel that sends every exp2 to the MUFU cannot exceed 50 % tensor pipe, and
// the single-tile kernel (attn_fwd.cu: one Q tile per CTA, P through shared memory) measured 27 %.  This kernel:
//   * one CTA per SM works on TWO 128-query tiles of one (batch, head) against the same K / V stream: while one
//     softmax warpgroup exponentiates S_t(j) the tensor pipe runs the other tile's S / PV MMAs;
//   * P never touches shared memory: the softmax threads write bf16 P straight into TENSOR MEMORY (tcgen05.st) and the
//     PV MMA takes its A operand from TMEM (no 32 KB smem round trip, no generic->async proxy fence);
//   * ~45 % of the exponentials are evaluated on the FMA pipe (Cody-Waite split + cubic minimax polynomial, packed
//     fp32x2 arithmetic, relative error 7.5e-5 << half a bf16 ulp of P) so that MUFU and issue slots are balanced;
//   * no running row maximum: the first key tile fixes the exponent reference exactly (row max), later tiles only
//     guard against overflow (a rare slow path raises the reference and rescales l / O exactly).  Softmax is shift
//     invariant and P / l / O carry the same factor, so results are unchanged.
// TMEM columns: S0 [0,128) S1 [128,256) P0 [256,320) P1 [320,384) O0 [384,448) O1 [448,512).
//   warps 0-15 : softmax.  Warp w works on Q tile t = w / 8, key-column half hh = (w / 4) % 2 and rows 32 (w % 4) + lane:
//                TWO threads per query row (64 of the 128 keys each), four softmax warps per scheduler — one warp per
//                scheduler and tile (thread = whole row) left the kernel latency bound at 31 % tensor pipe.  The two halves
//                of a row share only the exponent reference: they exchange their row maxima once (first key tile) and
//                meet at one named barrier per key tile, where a rare overflow flag is checked.
//   warp 16    : TMA producer — Q0 / Q1 once, K / V tiles (128 keys) through 3-stage rings, read in place from the
//                projection outputs via 4-D tensor maps {64, head, token, batch} (no head-split copy)
//   warp 17    : MMA issuer (one elected thread): per key tile and Q tile t: [P_t(j) ready] S_t(j+1) then PV_t(j)
//   warp 18    : TMEM allocator
#include <cuda.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/t2v_b200.h"
#include "host_common.h"
#include "ptx.cuh"

namespace t2v {

constexpr int kA2Threads = 608;
constexpr int kA2Stages = 3;
constexpr int kA2Tile = 128;
constexpr int kA2TileBytes = kA2Tile * 64 * 2;  // 16 KB: 128 rows x 64 bf16
constexpr int kA2Smem = 2 * kA2TileBytes + 2 * kA2Stages * kA2TileBytes + 512 + 2 * 2 * 128 * 4 + 64;
constexpr int kA2SatDefault = 0;        // 1: the saturating-FMA range reduction (ex2_poly2_sat)
constexpr int kA2PolyPairsDefault = 2;  // of the 16 column pairs of every 32-column chunk: exp2 on the FMA pipe

struct Attn2Params {
  int32_t heads, len_q, len_k, n_q_pairs, n_kv_tiles, kv_batch_div;
  float scale_log2;
  __nv_bfloat16* o;
  int64_t o_stride_b, o_stride_t, o_stride_h;
};

__device__ __forceinline__ float ex2_mufu(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// 2^x for a PAIR of arguments on the FMA / ALU pipes: x = xi + fr, xi = round(x), fr in [-0.5, 0.5];
// 2^fr by a cubic minimax polynomial (max relative error 7.5e-5), 2^xi added into the exponent field.
__device__ __forceinline__ float2 ex2_poly2(float2 x) {
  constexpr float kMagic = 12582912.f;  // 1.5 * 2^23: x + kMagic rounds x to an integer held in the low mantissa bits
  x.x = fminf(fmaxf(x.x, -125.f), 126.f);  // keep 2^xi a normal number: below, the result is ~0 either way; above, 2^126
  x.y = fminf(fmaxf(x.y, -125.f), 126.f);  // still trips the overflow guard of the row sum (slow path)
  const float2 t = add_f32x2(x, make_float2(kMagic, kMagic));
  const float2 xi = add_f32x2(t, make_float2(-kMagic, -kMagic));
  const float2 fr = fma_f32x2(xi, make_float2(-1.f, -1.f), x);
  float2 p = fma_f32x2(fr, make_float2(0.0551716685f, 0.0551716685f), make_float2(0.2426111251f, 0.2426111251f));
  p = fma_f32x2(p, fr, make_float2(0.6932609677f, 0.6932609677f));
  p = fma_f32x2(p, fr, make_float2(0.9999280572f, 0.9999280572f));
  float2 r;
  r.x = __uint_as_float(__float_as_uint(p.x) + (__float_as_uint(t.x) << 23));
  r.y = __uint_as_float(__float_as_uint(p.y) + (__float_as_uint(t.y) << 23));
  return r;
}

// 2^x' for a pair, x' = 128 y - 112 with y = sat((s c + 112 - ref) / 128) already formed by ONE saturating FMA per element:
// the saturation IS the range clamp (x' in [-112, 16]; 2^16 trips the row-sum guard), the integer part comes from one
// packed FMA, the exponent insert is one integer multiply-add per element: 10 issue slots per pair instead of 15.
__device__ __forceinline__ float sat_fma(float a, float b, float c) {
  float y;
  asm("fma.rn.sat.f32 %0, %1, %2, %3;" : "=f"(y) : "f"(a), "f"(b), "f"(c));
  return y;
}
__device__ __forceinline__ float2 ex2_poly2_sat(float2 y) {
  constexpr float kMagicM = 12582912.f - 112.f;   // t = x' + 1.5 * 2^23: round(x') sits in the low mantissa bits
  const float2 c128 = make_float2(128.f, 128.f);
  const float2 t = fma_f32x2(y, c128, make_float2(kMagicM, kMagicM));
  const float2 u = fma_f32x2(t, make_float2(-1.f, -1.f), make_float2(kMagicM, kMagicM));   // -(round(x') + 112), exact
  const float2 fr = fma_f32x2(y, c128, u);                                                   // x' - round(x')
  float2 p = fma_f32x2(fr, make_float2(0.0551716685f, 0.0551716685f), make_float2(0.2426111251f, 0.2426111251f));
  p = fma_f32x2(p, fr, make_float2(0.6932609677f, 0.6932609677f));
  p = fma_f32x2(p, fr, make_float2(0.9999280572f, 0.9999280572f));
  float2 r;
  uint32_t r0, r1;
  asm("mad.lo.u32 %0, %1, 0x800000, %2;" : "=r"(r0) : "r"(__float_as_uint(t.x)), "r"(__float_as_uint(p.x)));
  asm("mad.lo.u32 %0, %1, 0x800000, %2;" : "=r"(r1) : "r"(__float_as_uint(t.y)), "r"(__float_as_uint(p.y)));
  r.x = __uint_as_float(r0);
  r.y = __uint_as_float(r1);
  return r;
}

// kA2PolyPairs: column pairs per 32-column chunk whose exp2 runs on the FMA pipe.  kAbl != 0 are TIMING ABLATIONS (wrong
// results, scripts/attn_ablate.py only): 1 = no exponentials (P = bf16 of the shifted score), 2 = 1 + no per-tile barrier,
// 3 = the softmax warps only run the mbarrier protocol (MMA + synchronisation floor), 4 = 3 without the PV MMAs, 5 = 3 without
// the S MMAs.
template <int kA2PolyPairs, int kAbl, int kSat = 0>
__global__ void __launch_bounds__(kA2Threads, 1)
attn_fwd2_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                 const __grid_constant__ CUtensorMap tmV, const Attn2Params p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();  // 128B swizzle needs 1024-byte aligned tiles
  uint8_t* sQ = smem;                               // 2 tiles
  uint8_t* sK = sQ + 2 * kA2TileBytes;              // kA2Stages tiles
  uint8_t* sV = sK + kA2Stages * kA2TileBytes;      // kA2Stages tiles
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + kA2Stages * kA2TileBytes);
  uint64_t* q_full = bars;                    // 1
  uint64_t* k_full = bars + 1;                // kA2Stages
  uint64_t* k_empty = k_full + kA2Stages;
  uint64_t* v_full = k_empty + kA2Stages;
  uint64_t* v_empty = v_full + kA2Stages;
  uint64_t* s_full = v_empty + kA2Stages;     // 2 (per Q tile)
  uint64_t* p_full = s_full + 2;              // 2, 8 arrivals each (one per softmax warp)
  uint64_t* pv_done = p_full + 2;             // 2
  uint64_t* s_free = pv_done + 2;             // 2, 8 arrivals each: S_t(j) is in registers, its TMEM columns may be overwritten
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(s_free + 2);
  uint32_t* ovf_flag = tmem_slot + 2;                       // [2] per Q tile: some row of the tile overflowed its reference
  float* s_xchg = reinterpret_cast<float*>(bars) + 128;   // [2 tiles][2 halves][128 rows]: row maxima of the halves

  const int warp = __shfl_sync(0xffffffffu, int(threadIdx.x >> 5), 0);  // provably warp-uniform: lean TMA / MMA issue code
  const int lane = threadIdx.x & 31;

  const int q_pair = blockIdx.x % p.n_q_pairs;
  const int bh = blockIdx.x / p.n_q_pairs;
  const int h = bh % p.heads;
  const int b = bh / p.heads;
  const int kvb = b / p.kv_batch_div;
  const int q0 = q_pair * 2 * kA2Tile;

  if (warp == 16 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
  }
  if (warp == 17 && lane == 0) {
    ovf_flag[0] = ovf_flag[1] = 0u;
    mbar_init(q_full, 1);
    for (int s = 0; s < kA2Stages; ++s) {
      mbar_init(&k_full[s], 1);
      mbar_init(&k_empty[s], 1);
      mbar_init(&v_full[s], 1);
      mbar_init(&v_empty[s], 1);
    }
    for (int t = 0; t < 2; ++t) {
      mbar_init(&s_full[t], 1);
      mbar_init(&p_full[t], 8);     // one arrival per softmax warp of the tile
      mbar_init(&s_free[t], 8);
      mbar_init(&pv_done[t], 1);
    }
    fence_mbar_init();
  }
  if (warp == 18) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  pdl_launch_dependents();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();
  const int n_kv = p.n_kv_tiles;

  if (warp == 16 && elect_one()) {
    // ------------------------------------------------------------ TMA producer
    mbar_expect_tx(q_full, 2 * kA2TileBytes);
    tma_load_4d(sQ, &tmQ, q_full, 0, h, q0, b);
    tma_load_4d(sQ + kA2TileBytes, &tmQ, q_full, 0, h, q0 + kA2Tile, b);
    for (int j = 0; j < n_kv; ++j) {
      const int s = j % kA2Stages;
      const uint32_t ph = (j / kA2Stages) & 1;
      mbar_wait_relaxed(&k_empty[s], ph ^ 1u);
      mbar_expect_tx(&k_full[s], kA2TileBytes);
      tma_load_4d(sK + s * kA2TileBytes, &tmK, &k_full[s], 0, h, j * kA2Tile, kvb);
      mbar_wait_relaxed(&v_empty[s], ph ^ 1u);
      mbar_expect_tx(&v_full[s], kA2TileBytes);
      tma_load_4d(sV + s * kA2TileBytes, &tmV, &v_full[s], 0, h, j * kA2Tile, kvb);
    }
  } else if (warp == 17) {
    // ------------------------------------------------------------ MMA issuer
    // The WHOLE warp walks the loop (uniform control flow: descriptor arithmetic and barrier addresses stay on the uniform
    // datapath); only the tcgen05.mma / commit instructions themselves are predicated on the elected lane.  With the loop
    // inside a single-thread branch the issuer spent ~12 dependent instructions per MMA (64-bit adds + R2UR moves) and its
    // own instruction latency — ~2000 cycles per key tile against 1024 cycles of tensor work — was the floor of the kernel
    // (scripts/attn_ablate.py, variant a3: 126 us with the softmax warps doing nothing).
    const bool leader = elect_one();
    constexpr uint32_t idesc_s = umma_idesc_bf16(128, 128, 0, 0);   // Q (K-major) x K (K-major)
    constexpr uint32_t idesc_o = umma_idesc_bf16(128, 64, 0, 1);    // P (TMEM, K-major) x V (MN-major)
    const uint32_t tm = __shfl_sync(0xffffffffu, tmem_base, 0);
    const uint64_t qdesc0 = umma_desc_sw128(smem_u32(sQ));
    const uint64_t kdesc0 = umma_desc_sw128(smem_u32(sK));
    const uint64_t vdesc0 = umma_desc_sw128(smem_u32(sV));
    constexpr uint64_t kTileDesc = kA2TileBytes >> 4;   // one 16 KB tile further, in descriptor (16-byte) units
    mbar_wait(q_full, 0);
    // S_t(jn) into TMEM columns [128 t, 128 t + 128); stage sn / parity phn of the K ring
    auto issue_s = [&](int t, int sn, uint32_t phn) {
      if (t == 0) {
        mbar_wait(&k_full[sn], phn);
        tc_fence_after();
      }
      const uint64_t kdesc = kdesc0 + uint64_t(sn) * kTileDesc;
      const uint64_t qd = qdesc0 + uint64_t(t) * kTileDesc;
      if (leader) {
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (kAbl != 5) umma_ss(tm + t * 128, qd + 2 * k, kdesc + 2 * k, idesc_s, k != 0);
        if (t == 1) umma_commit(&k_empty[sn]);
        umma_commit(&s_full[t]);
      }
    };
    issue_s(0, 0, 0);
    issue_s(1, 0, 0);
    int s = 0, sn = (kA2Stages > 1) ? 1 : 0;          // ring stage of key tile j / j + 1
    uint32_t ph = 0, phn = (kA2Stages > 1) ? 0u : 1u;   // and their parities
    for (int j = 0; j < n_kv; ++j) {
      // The softmax threads pull S_t(j) into registers first thing and release its columns (s_free): the next score tile
      // is computed WHILE they exponentiate, so softmax_t(j+1) never waits for the tensor pipe.
      if (j + 1 < n_kv) {
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          mbar_wait(&s_free[t], j & 1);
          tc_fence_after();
          issue_s(t, sn, phn);
        }
      }
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        mbar_wait(&p_full[t], j & 1);   // P_t(j) is in TMEM
        tc_fence_after();
        if (t == 0) {
          mbar_wait(&v_full[s], ph);
          tc_fence_after();
        }
        const uint64_t vdesc = vdesc0 + uint64_t(s) * kTileDesc;
        if (leader) {
#pragma unroll
          for (int kk = 0; kk < 8; ++kk) {
            // A: 16 keys = 8 TMEM columns of packed bf16 pairs; B: 16 key rows = 2048 bytes of the MN-major V tile
            if (kAbl != 4)
              umma_ts(tm + 384 + t * 64, tm + 256 + t * 64 + kk * 8, vdesc + uint64_t(kk * (2048 >> 4)), idesc_o, (j | kk) != 0);
          }
          if (t == 1) umma_commit(&v_empty[s]);
          umma_commit(&pv_done[t]);
        }
      }
      s = sn;
      ph = phn;
      if (++sn == kA2Stages) {
        sn = 0;
        phn ^= 1u;
      }
    }
  } else if (warp < 16) {
    // ------------------------------------------------------------ softmax / epilogue
    const int t = warp >> 3;         // Q tile
    const int hh = (warp >> 2) & 1;  // key-column half of every 128-key tile
    const int ew = warp & 3;         // TMEM lane group
    const int r = ew * 32 + lane;
    const uint32_t lane_addr = static_cast<uint32_t>(ew * 32) << 16;
    const uint32_t s_addr = tmem_base + lane_addr + t * 128 + hh * 64;
    const uint32_t p_addr = tmem_base + lane_addr + 256 + t * 64 + hh * 32;
    const uint32_t o_addr = tmem_base + lane_addr + 384 + t * 64 + hh * 32;   // this thread rescales / stores 32 of the 64 columns
    float* my_x = s_xchg + (t * 2 + hh) * 128 + r;
    const float* other_x = s_xchg + (t * 2 + (hh ^ 1)) * 128 + r;
    const float c_log2 = p.scale_log2;
    float m_ref = 0.f;   // exponent reference (log2 domain, scale folded in), fixed by the first key tile
    float l_run = 0.f;   // this half's share of the row sum
    for (int j = 0; j < n_kv; ++j) {
      mbar_wait(&s_full[t], j & 1);
      tc_fence_after();
      if (kAbl >= 3) {   // ablation: the barrier protocol alone (no TMEM traffic, no arithmetic)
        tc_fence_before();
        mbar_arrive_warp(&s_free[t]);
        if (j > 0) mbar_wait(&pv_done[t], (j - 1) & 1);
        tc_fence_before();
        mbar_arrive_warp(&p_full[t]);
        continue;
      }
      const int kv_left = p.len_k - j * kA2Tile - hh * 64;  // valid keys of this half (may be <= 0 in the last tile)
      // S_t(j), this thread's 64 columns, into registers; then the TMEM columns are released for S_t(j+1)
      uint32_t va[32], vb[32];
      tmem_ld_32x32(s_addr, va);
      tmem_ld_32x32(s_addr + 32, vb);
      tmem_wait_ld();
      tc_fence_before();
      mbar_arrive_warp(&s_free[t]);
      // exact row maximum over both halves (first tile; slow path): own 64 columns, exchange through shared memory
      auto row_max = [&]() {
        float mx = -INFINITY;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          if (i < kv_left) mx = fmaxf(mx, __uint_as_float(va[i]));
          if (32 + i < kv_left) mx = fmaxf(mx, __uint_as_float(vb[i]));
        }
        *my_x = mx;
        named_bar_sync(1 + t, 256);
        mx = fmaxf(mx, *other_x);
        named_bar_sync(1 + t, 256);   // the slots may be rewritten
        return mx;
      };
      if (j == 0) m_ref = row_max() * c_log2;
      bool p_free = (j == 0);  // P_t (and O_t) may be written once PV_t(j-1) has completed
      float2 rs2;
      auto exp_chunk = [&](const uint32_t (&v)[32], int c0, float nref) {
        uint32_t pk[16];
        const float2 c2 = make_float2(c_log2, c_log2), n2 = make_float2(nref, nref);
        const float cs = c_log2 * (1.f / 128.f), ns = (nref + 112.f) * (1.f / 128.f);
        if (c0 + 32 <= kv_left) {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            float2 x, e;
            if (kSat && kAbl == 0 && i < kA2PolyPairs) {
              e = ex2_poly2_sat(make_float2(sat_fma(__uint_as_float(v[2 * i]), cs, ns), sat_fma(__uint_as_float(v[2 * i + 1]), cs, ns)));
              rs2 = add_f32x2(rs2, e);
              pk[i] = pack_bf16(e.x, e.y);
              continue;
            }
            x = fma_f32x2(make_float2(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1])), c2, n2);
            if (kAbl != 0) {
              e = x;
            } else if (i < kA2PolyPairs) {
              e = ex2_poly2(x);
            } else {
              e.x = ex2_mufu(x.x);
              e.y = ex2_mufu(x.y);
            }
            rs2 = add_f32x2(rs2, e);
            pk[i] = pack_bf16(e.x, e.y);
          }
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const bool ok0 = c0 + 2 * i < kv_left, ok1 = c0 + 2 * i + 1 < kv_left;
            const float e0 = ok0 ? ex2_mufu(fmaf(__uint_as_float(v[2 * i]), c_log2, nref)) : 0.f;
            const float e1 = ok1 ? ex2_mufu(fmaf(__uint_as_float(v[2 * i + 1]), c_log2, nref)) : 0.f;
            rs2 = add_f32x2(rs2, make_float2(e0, e1));
            pk[i] = pack_bf16(e0, e1);
          }
        }
        if (!p_free) {
          mbar_wait(&pv_done[t], (j - 1) & 1);
          tc_fence_after();
          p_free = true;
        }
        tmem_st_32x16(p_addr + (c0 >> 1), pk);   // 32 keys = 16 columns of packed bf16 pairs
      };
      auto exp_pass = [&](float ref) {
        rs2 = make_float2(0.f, 0.f);
        exp_chunk(va, 0, -ref);
        exp_chunk(vb, 32, -ref);
        return rs2.x + rs2.y;
      };
      float rs = exp_pass(m_ref);
      if (kAbl != 2 && __any_sync(0xffffffffu, !(rs < (kSat ? 6e4f : 1e30f))) && lane == 0) atomicOr(ovf_flag + t, 1u);
      if (kAbl != 2) named_bar_sync(1 + t, 256);   // the two halves of every row of this tile meet once per key tile
      if (kAbl == 0 && *reinterpret_cast<volatile uint32_t*>(ovf_flag + t) != 0u) {
        // rare: a row's scores exceed the reference by ~2^100: all 256 threads of the tile raise the reference for the rows
        // that need it, rescale l and O exactly and redo this tile's P from the scores still held in registers
        const float new_ref = fmaxf(m_ref, row_max() * c_log2);
        const float alpha = ex2_mufu(m_ref - new_ref);  // 1 for rows that keep their reference
        l_run *= alpha;
        if (warp == (t << 3) && lane == 0) ovf_flag[t] = 0u;   // (every thread of the tile is past its read: row_max synchronised twice)
        if (j > 0) {
          // (p_free is already true: exp_pass waited for PV_t(j-1)); each half rescales its 32 columns of O
#pragma unroll
          for (int c0 = 0; c0 < 32; c0 += 16) {
            uint32_t ov[16];
            tmem_ld_32x16(o_addr + c0, ov);
            tmem_wait_ld();
#pragma unroll
            for (int i = 0; i < 16; ++i) ov[i] = __float_as_uint(__uint_as_float(ov[i]) * alpha);
            tmem_st_32x16(o_addr + c0, ov);
          }
        }
        m_ref = new_ref;
        rs = exp_pass(m_ref);
      }
      l_run += rs;
      tmem_wait_st();
      tc_fence_before();
      mbar_arrive_warp(&p_full[t]);
    }
    // total row sum = the two halves' shares
    *my_x = l_run;
    named_bar_sync(1 + t, 256);
    l_run += *other_x;
    // the last PV must have landed before O is read
    mbar_wait(&pv_done[t], (n_kv - 1) & 1);
    tc_fence_after();
    // epilogue: O / l -> bf16 -> global; this thread stores columns [32 hh, 32 hh + 32) of its row
    const int qi = q0 + t * kA2Tile + r;
    const float inv_l = 1.0f / l_run;
    __nv_bfloat16* orow = p.o + int64_t(b) * p.o_stride_b + int64_t(qi) * p.o_stride_t + int64_t(h) * p.o_stride_h + hh * 32;
    uint32_t ov[32];
    tmem_ld_32x32(o_addr, ov);
    tmem_wait_ld();
    if (qi < p.len_q) {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        uint4 pk;
        pk.x = pack_bf16(__uint_as_float(ov[q * 8 + 0]) * inv_l, __uint_as_float(ov[q * 8 + 1]) * inv_l);
        pk.y = pack_bf16(__uint_as_float(ov[q * 8 + 2]) * inv_l, __uint_as_float(ov[q * 8 + 3]) * inv_l);
        pk.z = pack_bf16(__uint_as_float(ov[q * 8 + 4]) * inv_l, __uint_as_float(ov[q * 8 + 5]) * inv_l);
        pk.w = pack_bf16(__uint_as_float(ov[q * 8 + 6]) * inv_l, __uint_as_float(ov[q * 8 + 7]) * inv_l);
        *reinterpret_cast<uint4*>(orow + q * 8) = pk;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 18) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// host side of the two-tile kernel; returns < 0 for "not applicable" so that t2v_attn_fwd can use the single-tile kernel
int launch_attn_fwd2(const T2VAttnDesc* d, const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv,
                     cudaStream_t stream) {
  Attn2Params p;
  p.heads = d->heads;
  p.len_q = d->len_q;
  p.len_k = d->len_k;
  p.n_q_pairs = (d->len_q + 2 * kA2Tile - 1) / (2 * kA2Tile);
  p.n_kv_tiles = (d->len_k + kA2Tile - 1) / kA2Tile;
  p.kv_batch_div = d->kv_batch_div;
  p.scale_log2 = d->scale * 1.4426950408889634f;
  p.o = static_cast<__nv_bfloat16*>(d->o);
  p.o_stride_b = d->o_stride_b;
  p.o_stride_t = d->o_stride_t;
  p.o_stride_h = d->o_stride_h;
  using KernelT = void (*)(CUtensorMap, CUtensorMap, CUtensorMap, Attn2Params);
  // T2V_ATTN_V2_VARIANT (timing experiments, scripts/attn_ablate.py): "p<N>" / "q<N>" = N polynomial pairs per chunk with the
  // clamped / saturating-FMA range reduction, "a1" / "a2" = ablations with WRONG results.  Unset = the shipped configuration.
  struct Variant { const char* name; KernelT fn; bool configured; };
  static Variant variants[] = {
      {"", attn_fwd2_kernel<kA2PolyPairsDefault, 0, kA2SatDefault>, false},
      {"p0", attn_fwd2_kernel<0, 0, 0>, false},   {"p4", attn_fwd2_kernel<4, 0, 0>, false},
      {"p2", attn_fwd2_kernel<2, 0, 0>, false},   {"p3", attn_fwd2_kernel<3, 0, 0>, false},
      {"p5", attn_fwd2_kernel<5, 0, 0>, false},   {"p6", attn_fwd2_kernel<6, 0, 0>, false},
      {"p8", attn_fwd2_kernel<8, 0, 0>, false},   {"q2", attn_fwd2_kernel<2, 0, 1>, false},
      {"q6", attn_fwd2_kernel<6, 0, 1>, false},   {"q8", attn_fwd2_kernel<8, 0, 1>, false},   {"q3", attn_fwd2_kernel<3, 0, 1>, false},
      {"q4", attn_fwd2_kernel<4, 0, 1>, false},
      {"p7", attn_fwd2_kernel<7, 0, 0>, false},   {"p16", attn_fwd2_kernel<16, 0, 0>, false},
      {"q5", attn_fwd2_kernel<5, 0, 1>, false},   {"q7", attn_fwd2_kernel<7, 0, 1>, false},
      {"q9", attn_fwd2_kernel<9, 0, 1>, false},   {"a1", attn_fwd2_kernel<7, 1, 0>, false},
      {"a2", attn_fwd2_kernel<7, 2, 0>, false},   {"a3", attn_fwd2_kernel<7, 3, 0>, false},
      {"a4", attn_fwd2_kernel<7, 4, 0>, false},   {"a5", attn_fwd2_kernel<7, 5, 0>, false}};
  Variant* var = &variants[0];
  if (const char* name = getenv("T2V_ATTN_V2_VARIANT")) {
    for (Variant& v : variants)
      if (!strcmp(v.name, name)) var = &v;
  }
  if (!var->configured) {
    cudaError_t e = cudaFuncSetAttribute(var->fn, cudaFuncAttributeMaxDynamicSharedMemorySize, kA2Smem);
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(attn_fwd2)");
    var->configured = true;
  }
  KernelT kernel = var->fn;
  const int64_t grid = int64_t(d->batch) * d->heads * p.n_q_pairs;
  if (grid > 0x7fffffff) return fail(-5, "t2v_attn_fwd: grid too large");
  launch_kernel(kernel, dim3(unsigned(grid)), dim3(kA2Threads), kA2Smem, stream, tq, tk, tv, p);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : cuda_fail(e, "t2v_attn_fwd (two-tile kernel) launch");
}

}  // namespace t2v
