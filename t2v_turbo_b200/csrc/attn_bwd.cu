// Flash-attention backward, head_dim 64, no mask, on tcgen05 (sm_100a): the adjoint of attn_fwd.cu.
//
//   P = exp2(scale*log2e * Q K^T - lse2)      dP = dO V^T      dS = P * (dP - delta) * scale
//   dQ = dS K                                 dK = dS^T Q      dV = P^T dO
//
// ONE kernel, launched twice with the roles of queries and keys exchanged.  A CTA owns a 128-row tile of the "row"
// operand pair (R1, R2) and walks 128-column tiles of the "column" pair (C1, C2):
//
//     T  = R1 C1^T         (M128 N128 K64, TMEM columns   0..127)
//     U  = R2 C2^T         (M128 N128 K64, TMEM columns 128..255)
//     Pt = exp2(T * scale_log2 - lse2[query]),   G = Pt * (U - delta[query]) * scale        (thread = row, fp32)
//     out1 += G  C1        (M128 N64 K128, G from shared memory, C1 read in place as an MN-major operand)
//     out2 += Pt C2        (dK/dV launch only)
//
//   dQ launch   (kRowStats = true ): rows = queries: R1 = Q, R2 = dO, C1 = K, C2 = V   -> out1 = dQ
//   dK/dV launch(kRowStats = false): rows = keys:    R1 = K, R2 = V,  C1 = Q, C2 = dO  -> out1 = dK, out2 = dV
//
// (T = S in the first launch and S^T in the second; the per-query statistics index rows in the first and columns in the
// second.)  The dK/dV CTA loops over every query batch that shares its K/V batch (cross-attention: kv_batch_div frames),
// so no gradient needs atomics.  S is recomputed in both launches: 7 GEMMs of 128x128x64 per tile pair instead of 5, in
// exchange for one kernel and no global dQ accumulation.
//   warp 0 / lane 0 : TMA producer (R tiles once, C tile pairs through a 2-stage ring; 4-D maps {64, head, token, batch}:
//                     the projection outputs and dO are read in place)
//   warp 1 / lane 0 : MMA issuer;  warp 2: TMEM allocator (512 columns, one CTA per SM);  warps 4..7: the 128 row threads.
#include <cuda.h>
#include <math.h>

#include "../../include/t2v_b200.h"
#include "host_common.h"
#include "ptx.cuh"

namespace t2v {

namespace {

constexpr int kAbThreads = 256;
constexpr int kStages = 2;
constexpr int kT = 128;
constexpr int kTileBytes = kT * 64 * 2;        // 16 KB: 128 tokens x 64 channels
constexpr int kGBytes = kT * kT * 2;           // 32 KB: two K-major 64-column sub-tiles
constexpr int kStatBytes = kStages * 2 * kT * 4;
constexpr int kAbSmem = 2 * kTileBytes + 2 * kStages * kTileBytes + 2 * kGBytes + kStatBytes + 256;
constexpr int kAbTmemCols = 512;
constexpr int kColU = 128, kColO1 = 256, kColO2 = 320;

struct AttnBwdParams {
  int32_t heads, len_r, len_c, n_r_tiles, n_c_tiles, n_cb, div, len_q;
  float scale_log2, scale;
  const float* lse2;
  const float* delta;
  __nv_bfloat16* out1;
  int64_t o1_sb, o1_st, o1_sh;
  __nv_bfloat16* out2;
  int64_t o2_sb, o2_st, o2_sh;
};

__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <bool kRowStats>
__global__ void __launch_bounds__(kAbThreads, 1)
attn_bwd_kernel(const __grid_constant__ CUtensorMap tmR1, const __grid_constant__ CUtensorMap tmR2,
                const __grid_constant__ CUtensorMap tmC1, const __grid_constant__ CUtensorMap tmC2, const AttnBwdParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  uint8_t* sR1 = smem;
  uint8_t* sR2 = sR1 + kTileBytes;
  uint8_t* sC1 = sR2 + kTileBytes;
  uint8_t* sC2 = sC1 + kStages * kTileBytes;
  uint8_t* sG = sC2 + kStages * kTileBytes;
  uint8_t* sP = sG + kGBytes;
  float* s_stat = reinterpret_cast<float*>(sP + kGBytes);   // [stage][lse2 | delta][128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(s_stat) + kStatBytes);
  uint64_t* r_full = bars;                  // 1
  uint64_t* c_full = bars + 1;              // kStages
  uint64_t* c_empty = c_full + kStages;     // kStages
  uint64_t* s_full = c_empty + kStages;     // 1
  uint64_t* p_full = s_full + 1;            // 1 (4 arrivals: one per row warp)
  uint64_t* o_done = p_full + 1;            // 1
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_done + 1);

  const int warp = __shfl_sync(0xffffffffu, int(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;

  const int r_tile = blockIdx.x % p.n_r_tiles;
  const int bh = blockIdx.x / p.n_r_tiles;
  const int h = bh % p.heads;
  const int rb = bh / p.heads;
  const int r0 = r_tile * kT;
  const int cb0 = kRowStats ? rb / p.div : rb * p.div;
  const int n_it = p.n_cb * p.n_c_tiles;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmR1);
    tma_prefetch_desc(&tmR2);
    tma_prefetch_desc(&tmC1);
    tma_prefetch_desc(&tmC2);
  }
  if (warp == 1 && lane == 0) {
    mbar_init(r_full, 1);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&c_full[s], 1);
      mbar_init(&c_empty[s], 1);
    }
    mbar_init(s_full, 1);
    mbar_init(p_full, 4);   // one arrival per softmax warp
    mbar_init(o_done, 1);
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, kAbTmemCols);
    tmem_relinquish();
  }
  pdl_launch_dependents();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();

  if (warp == 0 && elect_one()) {
    // ------------------------------------------------------------ TMA producer
    mbar_expect_tx(r_full, 2 * kTileBytes);
    tma_load_4d(sR1, &tmR1, r_full, 0, h, r0, rb);
    tma_load_4d(sR2, &tmR2, r_full, 0, h, r0, rb);
    for (int it = 0; it < n_it; ++it) {
      const int s = it % kStages;
      const uint32_t ph = (it / kStages) & 1;
      const int cb = cb0 + it / p.n_c_tiles;
      const int c0 = (it % p.n_c_tiles) * kT;
      mbar_wait(&c_empty[s], ph ^ 1u);
      mbar_expect_tx(&c_full[s], 2 * kTileBytes);
      tma_load_4d(sC1 + s * kTileBytes, &tmC1, &c_full[s], 0, h, c0, cb);
      tma_load_4d(sC2 + s * kTileBytes, &tmC2, &c_full[s], 0, h, c0, cb);
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    // whole-warp loop, only tcgen05.mma / commit predicated on the elected lane (uniform-datapath descriptor arithmetic:
    // see attn_fwd2.cu for the measurement that motivated it)
    const bool leader = elect_one();
    constexpr uint32_t idesc_s = umma_idesc_bf16(128, 128, 0, 0);   // K-major x K-major (reduction over the 64 channels)
    constexpr uint32_t idesc_o = umma_idesc_bf16(128, 64, 0, 1);    // G / Pt (K-major) x C tile (MN-major: reduction over its rows)
    const uint32_t tm = __shfl_sync(0xffffffffu, tmem_base, 0);
    const uint64_t r1desc = umma_desc_sw128(smem_u32(sR1));
    const uint64_t r2desc = umma_desc_sw128(smem_u32(sR2));
    const uint64_t gdesc = umma_desc_sw128(smem_u32(sG));
    const uint64_t pdesc = umma_desc_sw128(smem_u32(sP));
    const uint64_t c1desc0 = umma_desc_sw128(smem_u32(sC1));
    const uint64_t c2desc0 = umma_desc_sw128(smem_u32(sC2));
    constexpr uint64_t kTileDesc = kTileBytes >> 4;   // one ring stage further, in descriptor (16-byte) units
    mbar_wait(r_full, 0);
    auto issue_s = [&](int sn, uint32_t phn) {
      mbar_wait(&c_full[sn], phn);
      tc_fence_after();
      const uint64_t c1desc = c1desc0 + uint64_t(sn) * kTileDesc;
      const uint64_t c2desc = c2desc0 + uint64_t(sn) * kTileDesc;
      if (leader) {
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_ss(tm, r1desc + 2 * k, c1desc + 2 * k, idesc_s, k != 0);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_ss(tm + kColU, r2desc + 2 * k, c2desc + 2 * k, idesc_s, k != 0);
        umma_commit(s_full);
      }
    };
    issue_s(0, 0);
    int s = 0, sn = (kStages > 1) ? 1 : 0;            // ring stage of iteration it / it + 1
    uint32_t phn = (kStages > 1) ? 0u : 1u;            // parity of c_full[sn]
    for (int it = 0; it < n_it; ++it) {
      mbar_wait(p_full, it & 1);
      tc_fence_after();
      if (it + 1 < n_it) issue_s(sn, phn);   // T / U columns are free: every row thread has read them before arriving
      const uint64_t c1desc = c1desc0 + uint64_t(s) * kTileDesc;
      const uint64_t c2desc = c2desc0 + uint64_t(s) * kTileDesc;
      if (leader) {
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
          // A: 16 columns = 32 bytes inside the (kk/4)-th 64-column sub-tile; B: 16 token rows of the C tile = 2048 bytes
          const uint64_t aoff = uint64_t((kk >> 2) * (kT * 128 >> 4)) + 2 * (kk & 3);
          const uint64_t boff = uint64_t(kk * (2048 >> 4));
          umma_ss(tm + kColO1, gdesc + aoff, c1desc + boff, idesc_o, (it | kk) != 0);
        }
        if (!kRowStats) {
#pragma unroll
          for (int kk = 0; kk < 8; ++kk) {
            const uint64_t aoff = uint64_t((kk >> 2) * (kT * 128 >> 4)) + 2 * (kk & 3);
            const uint64_t boff = uint64_t(kk * (2048 >> 4));
            umma_ss(tm + kColO2, pdesc + aoff, c2desc + boff, idesc_o, (it | kk) != 0);
          }
        }
        umma_commit(&c_empty[s]);
        umma_commit(o_done);
      }
      s = sn;
      if (++sn == kStages) {
        sn = 0;
        phn ^= 1u;
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------ row threads
    const int ew = warp - 4;
    const int r = ew * 32 + lane;
    const uint32_t lane_addr = static_cast<uint32_t>(ew * 32) << 16;
    const uint32_t t_addr = tmem_base + lane_addr;
    uint8_t* g_row = sG + r * 128;
    uint8_t* p_row = sP + r * 128;
    const int sw = r & 7;
    float lse_r = INFINITY, delta_r = 0.f;
    if (kRowStats && r0 + r < p.len_r) {
      const int64_t si = (int64_t(rb) * p.heads + h) * p.len_q + r0 + r;
      lse_r = __ldg(p.lse2 + si);
      delta_r = __ldg(p.delta + si);
    }
    for (int it = 0; it < n_it; ++it) {
      const int cb = cb0 + it / p.n_c_tiles;
      const int c_tile0 = (it % p.n_c_tiles) * kT;
      float* st = s_stat + (it & 1) * 2 * kT;
      if (!kRowStats) {
        // the statistics of this column tile's 128 queries (thread r loads query r); +inf lse2 zeroes padded queries
        const int qi = c_tile0 + r;
        float l = INFINITY, dl = 0.f;
        if (qi < p.len_c) {
          const int64_t si = (int64_t(cb) * p.heads + h) * p.len_q + qi;
          l = __ldg(p.lse2 + si);
          dl = __ldg(p.delta + si);
        }
        st[r] = l;
        st[kT + r] = dl;
        named_bar_sync(1, 128);
      }
      mbar_wait(s_full, it & 1);
      tc_fence_after();
      const int c_left = p.len_c - c_tile0;   // valid columns in this tile
      bool bufs_free = (it == 0);
#pragma unroll 1
      for (int c0 = 0; c0 < kT; c0 += 32) {
        uint32_t tv[32], uv[32];
        tmem_ld_32x32(t_addr + c0, tv);
        tmem_ld_32x32(t_addr + kColU + c0, uv);
        tmem_wait_ld();
        uint32_t gk[16], pk[16];
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          float pv[2], gv[2];
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const float l = kRowStats ? lse_r : st[c0 + i + e];
            const float dl = kRowStats ? delta_r : st[kT + c0 + i + e];
            float pe = ex2f(fmaf(__uint_as_float(tv[i + e]), p.scale_log2, -l));
            if (kRowStats && c0 + i + e >= c_left) pe = 0.f;   // padded keys (zero-filled K rows give T = 0, not -inf)
            pv[e] = pe;
            gv[e] = pe * (__uint_as_float(uv[i + e]) - dl) * p.scale;
          }
          gk[i >> 1] = pack_bf16(gv[0], gv[1]);
          pk[i >> 1] = pack_bf16(pv[0], pv[1]);
        }
        if (!bufs_free) {   // G / Pt of the previous iteration are still operands of its out-MMAs until o_done
          mbar_wait(o_done, (it - 1) & 1);
          tc_fence_after();
          bufs_free = true;
        }
        const int sub = (c0 >> 6) * (kT * 128);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int cc = ((c0 & 63) >> 3) + q;
          *reinterpret_cast<uint4*>(g_row + sub + ((cc ^ sw) << 4)) = make_uint4(gk[4 * q], gk[4 * q + 1], gk[4 * q + 2], gk[4 * q + 3]);
          if (!kRowStats)
            *reinterpret_cast<uint4*>(p_row + sub + ((cc ^ sw) << 4)) = make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
        }
      }
      fence_proxy_async();
      tc_fence_before();
      mbar_arrive_warp(p_full);
    }
    mbar_wait(o_done, (n_it - 1) & 1);
    tc_fence_after();
    // epilogue: fp32 accumulators -> bf16 -> global
    const int ri = r0 + r;
    auto store_rows = [&](uint32_t col, __nv_bfloat16* base, int64_t sb, int64_t stt, int64_t sh) {
      __nv_bfloat16* orow = base + int64_t(rb) * sb + int64_t(ri) * stt + int64_t(h) * sh;
#pragma unroll
      for (int c0 = 0; c0 < 64; c0 += 32) {
        uint32_t ov[32];
        tmem_ld_32x32(t_addr + col + c0, ov);
        tmem_wait_ld();
        if (ri < p.len_r) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            uint4 pkv;
            pkv.x = pack_bf16(__uint_as_float(ov[q * 8 + 0]), __uint_as_float(ov[q * 8 + 1]));
            pkv.y = pack_bf16(__uint_as_float(ov[q * 8 + 2]), __uint_as_float(ov[q * 8 + 3]));
            pkv.z = pack_bf16(__uint_as_float(ov[q * 8 + 4]), __uint_as_float(ov[q * 8 + 5]));
            pkv.w = pack_bf16(__uint_as_float(ov[q * 8 + 6]), __uint_as_float(ov[q * 8 + 7]));
            *reinterpret_cast<uint4*>(orow + c0 + q * 8) = pkv;
          }
        }
      }
    };
    store_rows(kColO1, p.out1, p.o1_sb, p.o1_st, p.o1_sh);
    if (!kRowStats) store_rows(kColO2, p.out2, p.o2_sb, p.o2_st, p.o2_sh);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kAbTmemCols);
  }
}

int make_map(CUtensorMap* tm, const void* base, int heads, int len, int batch, int64_t sb, int64_t st, int64_t sh, const char* what) {
  uint64_t dims[4] = {64, uint64_t(heads), uint64_t(len), uint64_t(batch)};
  uint64_t str[4] = {2, uint64_t(sh) * 2, uint64_t(st) * 2, uint64_t(sb) * 2};
  uint32_t box[4] = {64, 1, 128, 1};
  if (heads == 1 && str[1] == 0) str[1] = 128;
  if (batch == 1 && str[3] == 0) str[3] = str[2] * len;
  return make_tmap_bf16(tm, base, 4, dims, str, box, what);
}

bool bad_out(const void* p, int64_t sb, int64_t st, int64_t sh) {
  return (sb % 8) || (st % 8) || (sh % 8) || (reinterpret_cast<uintptr_t>(p) & 15);
}

}  // namespace
}  // namespace t2v

extern "C" int t2v_attn_bwd(const T2VAttnBwdDesc* d, t2v_stream_t stream_) {
  using namespace t2v;
  if (!d || !d->q || !d->k || !d->v || !d->d_o || !d->lse2 || !d->delta) return fail(-1, "t2v_attn_bwd: null pointer");
  if (d->batch < 1 || d->heads < 1 || d->len_q < 1 || d->len_k < 1) return fail(-2, "t2v_attn_bwd: bad sizes");
  if (d->kv_batch_div < 1 || d->batch % d->kv_batch_div) return fail(-3, "t2v_attn_bwd: batch %% kv_batch_div != 0");
  if ((d->dk == nullptr) != (d->dv == nullptr)) return fail(-4, "t2v_attn_bwd: dk and dv are produced together");
  if (d->dq && bad_out(d->dq, d->dq_stride_b, d->dq_stride_t, d->dq_stride_h)) return fail(-5, "t2v_attn_bwd: dq alignment / strides");
  if (d->dk && (bad_out(d->dk, d->dk_stride_b, d->dk_stride_t, d->dk_stride_h) || bad_out(d->dv, d->dv_stride_b, d->dv_stride_t, d->dv_stride_h)))
    return fail(-5, "t2v_attn_bwd: dk / dv alignment / strides");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const int kvb = d->batch / d->kv_batch_div;
  CUtensorMap tq, tk, tv, tdo;
  int rc;
  if ((rc = make_map(&tq, d->q, d->heads, d->len_q, d->batch, d->q_stride_b, d->q_stride_t, d->q_stride_h, "t2v_attn_bwd Q"))) return rc;
  if ((rc = make_map(&tdo, d->d_o, d->heads, d->len_q, d->batch, d->do_stride_b, d->do_stride_t, d->do_stride_h, "t2v_attn_bwd dO"))) return rc;
  if ((rc = make_map(&tk, d->k, d->heads, d->len_k, kvb, d->k_stride_b, d->k_stride_t, d->k_stride_h, "t2v_attn_bwd K"))) return rc;
  if ((rc = make_map(&tv, d->v, d->heads, d->len_k, kvb, d->v_stride_b, d->v_stride_t, d->v_stride_h, "t2v_attn_bwd V"))) return rc;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(attn_bwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kAbSmem);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(attn_bwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kAbSmem);
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(attn_bwd)");
    configured = true;
  }
  AttnBwdParams p;
  p.heads = d->heads;
  p.len_q = d->len_q;
  p.div = d->kv_batch_div;
  p.scale = d->scale;
  p.scale_log2 = d->scale * 1.4426950408889634f;
  p.lse2 = d->lse2;
  p.delta = d->delta;
  const int nq = (d->len_q + kT - 1) / kT, nk = (d->len_k + kT - 1) / kT;
  if (d->dq) {
    p.len_r = d->len_q; p.len_c = d->len_k; p.n_r_tiles = nq; p.n_c_tiles = nk; p.n_cb = 1;
    p.out1 = static_cast<__nv_bfloat16*>(d->dq); p.o1_sb = d->dq_stride_b; p.o1_st = d->dq_stride_t; p.o1_sh = d->dq_stride_h;
    p.out2 = nullptr; p.o2_sb = p.o2_st = p.o2_sh = 0;
    const int64_t grid = int64_t(d->batch) * d->heads * nq;
    if (grid > 0x7fffffff) return fail(-6, "t2v_attn_bwd: grid too large");
    launch_kernel(attn_bwd_kernel<true>, dim3(unsigned(grid)), dim3(kAbThreads), kAbSmem, stream, tq, tdo, tk, tv, p);
  }
  if (d->dk) {
    p.len_r = d->len_k; p.len_c = d->len_q; p.n_r_tiles = nk; p.n_c_tiles = nq; p.n_cb = d->kv_batch_div;
    p.out1 = static_cast<__nv_bfloat16*>(d->dk); p.o1_sb = d->dk_stride_b; p.o1_st = d->dk_stride_t; p.o1_sh = d->dk_stride_h;
    p.out2 = static_cast<__nv_bfloat16*>(d->dv); p.o2_sb = d->dv_stride_b; p.o2_st = d->dv_stride_t; p.o2_sh = d->dv_stride_h;
    const int64_t grid = int64_t(kvb) * d->heads * nk;
    if (grid > 0x7fffffff) return fail(-6, "t2v_attn_bwd: grid too large");
    launch_kernel(attn_bwd_kernel<false>, dim3(unsigned(grid)), dim3(kAbThreads), kAbSmem, stream, tk, tv, tq, tdo, p);
  }
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : cuda_fail(e, "t2v_attn_bwd launch");
}
