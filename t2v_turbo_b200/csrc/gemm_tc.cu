// Implicit GEMM on tcgen05 (sm_100a): the one tensor-core kernel behind every Linear, Conv2d 3x3/1x1,
// strided conv, Conv3d (3,1,1) and Conv1d(k=1) of the VideoCrafter2 UNet and the KL-VAE decoder.
//
//   out[m, n] = epi( alpha * sum_{tap, c} A[point(m) + off(tap), c] * W[n, tap * c_in + c] )
//
// Design (one CTA per SM, persistent over output tiles):
//   warp 0 / lane 0 : TMA producer.  A tile = 5-D box {64 ch, b1, b2, b3, b4} of the channels-last
//                     activation (im2col-free: the tap shift is a coordinate offset; out-of-range
//                     coordinates are zero-filled by TMA = conv padding); B tile = {64, BLOCK_N}
//                     of the K-major weight matrix.  Both land in 128B-swizzled smem.
//   warp 1 / lane 0 : MMA issuer. 4 x tcgen05.mma (M=128, N=BLOCK_N, K=16) per 64-wide k-block,
//                     fp32 accumulators in TMEM, double buffered (2 x BLOCK_N columns).
//   warp 2          : TMEM allocator.
//   warps 4..7      : epilogue.  tcgen05.ld (thread = output row), bias / residual / GEGLU,
//                     bf16 (or fp32) vector stores.  Overlaps the next tile's main loop.
// Pipelines: smem full/empty ring (TMA <-> MMA) and TMEM full/empty (MMA <-> epilogue).
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
  int32_t num_tiles;
  int32_t rows_in_box;
  int32_t n_taps;
  int32_t tap_off[T2V_MAX_TAPS][4];
  int32_t tap_ch_off[T2V_MAX_TAPS];
  int32_t kb_per_tap;
  int32_t kb_src0;
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
  uint32_t a_tile_bytes;
};

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;
constexpr int kATileBytes = kBlockM * kBlockK * 2;  // 16 KB
constexpr int kThreads = 256;
constexpr int kSmemBudget = 200 * 1024;

template <int BN>
struct GemmCfg {
  static constexpr int kBTileBytes = BN * kBlockK * 2;
  static constexpr int kStageBytes = kATileBytes + kBTileBytes;
  static constexpr int kStagesRaw = kSmemBudget / kStageBytes;
  static constexpr int kStages = kStagesRaw > 8 ? 8 : kStagesRaw;
  static constexpr int kAccStride = BN <= 32 ? 32 : BN <= 64 ? 64 : BN <= 128 ? 128 : 256;
  static constexpr int kTmemCols = 2 * kAccStride;
  static constexpr int kBarBytes = (2 * kStages + 4) * 8 + 16;
  static constexpr int kSmemBytes = kStages * kStageBytes + kBarBytes + 1024;
};

__device__ __forceinline__ float gelu_erf(float x) {
  return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f));
}

template <int BN>
__global__ void __launch_bounds__(kThreads, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
               const __grid_constant__ CUtensorMap tmB, const GemmParams p) {
  using Cfg = GemmCfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);

  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + Cfg::kStages * Cfg::kStageBytes);
  uint64_t* empty_bar = full_bar + Cfg::kStages;
  uint64_t* tfull_bar = empty_bar + Cfg::kStages;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA0);
    tma_prefetch_desc(&tmA1);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < Cfg::kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull_bar[a], 1);
      mbar_init(&tempty_bar[a], 4);
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, Cfg::kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int total_kb = p.n_taps * p.kb_per_tap;

  if (warp == 0 && lane == 0) {
    // ------------------------------------------------------------ TMA producer
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      const int n_tile = tile % p.n_tiles_n;
      int m_tile = tile / p.n_tiles_n;
      int o[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        o[j] = (m_tile % p.ntile[j]) * p.box[j];
        m_tile /= p.ntile[j];
      }
      const int bbatch = p.b_batch_dim >= 0 ? o[p.b_batch_dim] : 0;
      for (int tap = 0; tap < p.n_taps; ++tap) {
        const int c1 = o[0] + p.tap_off[tap][0];
        const int c2 = o[1] + p.tap_off[tap][1];
        const int c3 = o[2] + p.tap_off[tap][2];
        const int c4 = o[3] + p.tap_off[tap][3];
        const int ch0 = p.tap_ch_off[tap];
        for (int kb = 0; kb < p.kb_per_tap; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1u);
          uint8_t* sA = smem + stage * Cfg::kStageBytes;
          uint8_t* sB = sA + kATileBytes;
          mbar_expect_tx(&full_bar[stage], p.a_tile_bytes + Cfg::kBTileBytes);
          if (kb < p.kb_src0)
            tma_load_5d(sA, &tmA0, &full_bar[stage], ch0 + kb * kBlockK, c1, c2, c3, c4);
          else
            tma_load_5d(sA, &tmA1, &full_bar[stage], ch0 + (kb - p.kb_src0) * kBlockK, c1, c2, c3,
                        c4);
          tma_load_3d(sB, &tmB, &full_bar[stage], (tap * p.kb_per_tap + kb) * kBlockK, n_tile * BN,
                      bbatch);
          if (++stage == Cfg::kStages) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else if (warp == 1 && lane == 0) {
    // ------------------------------------------------------------ MMA issuer
    constexpr uint32_t idesc = umma_idesc_bf16(kBlockM, BN, 0, 0);
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      mbar_wait(&tempty_bar[acc], acc_phase ^ 1u);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + acc * Cfg::kAccStride;
      for (int kb = 0; kb < total_kb; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(smem + stage * Cfg::kStageBytes);
        const uint64_t adesc = umma_desc_sw128(a_addr);
        const uint64_t bdesc = umma_desc_sw128(a_addr + kATileBytes);
#pragma unroll
        for (int k = 0; k < kBlockK / 16; ++k) {
          // advance 16 elements (32 bytes) along K inside the 128B swizzle atom: +2 in addr>>4 units
          umma_ss(tmem_d, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
        }
        umma_commit(&empty_bar[stage]);
        if (++stage == Cfg::kStages) {
          stage = 0;
          phase ^= 1u;
        }
      }
      umma_commit(&tfull_bar[acc]);
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------ epilogue
    const int ew = warp - 4;  // == warp % 4 -> TMEM lanes [32*ew, 32*ew+32)
    const int r = ew * 32 + lane;
    const bool geglu = (p.flags & T2V_EPI_GEGLU) != 0;
    const bool out_f32 = (p.flags & T2V_EPI_OUT_F32) != 0;
    const bool gelu = (p.flags & T2V_EPI_GELU) != 0;
    int it = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      const int n_tile = tile % p.n_tiles_n;
      int m_tile = tile / p.n_tiles_n;
      // row -> point
      bool valid = r < p.rows_in_box;
      int64_t out_off = 0, res_off = 0;
      int64_t bias_row = 0;
      {
        int rr = r;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int oj = (m_tile % p.ntile[j]) * p.box[j];
          m_tile /= p.ntile[j];
          const int ij = rr % p.box[j];
          rr /= p.box[j];
          const int64_t x = oj + ij;
          valid = valid && (x < p.o_size[j]);
          out_off += x * p.o_stride[j];
          res_off += x * p.r_stride[j];
          if (j == p.bias_dim) bias_row = x / p.bias_div;
        }
      }
      const float* bias = p.bias ? p.bias + bias_row * p.bias_row_stride : nullptr;

      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + acc * Cfg::kAccStride + (static_cast<uint32_t>(ew * 32) << 16);

#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 32) {
        uint32_t v[32];
        tmem_ld_32x32(taddr + c0, v);
        tmem_wait_ld();
        const int n0 = n_tile * BN + c0;  // column in W-row space
        if (n0 >= p.n_rows_b) continue;   // warp-uniform
        float f[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          f[j] = p.alpha * __uint_as_float(v[j]);
          if (bias != nullptr && n0 + j < p.n_rows_b) f[j] += __ldg(bias + n0 + j);
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
      // accumulator drained: hand the TMEM buffer back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[acc]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

template <int BN>
static int launch_gemm(const CUtensorMap& a0, const CUtensorMap& a1, const CUtensorMap& b,
                       const GemmParams& p, cudaStream_t stream) {
  using Cfg = GemmCfg<BN>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tc_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         Cfg::kSmemBytes);
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(gemm_tc)");
    configured = true;
  }
  int sms = num_sms();
  if (sms <= 0) return fail(-110, "no CUDA device");
  int grid = p.num_tiles < sms ? p.num_tiles : sms;
  gemm_tc_kernel<BN><<<grid, kThreads, Cfg::kSmemBytes, stream>>>(a0, a1, b, p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "gemm_tc launch");
  return 0;
}

static int choose_block_n(int64_t n_rows, int64_t m_tiles, int sms, bool geglu) {
  const int cands[4] = {256, 160, 128, 64};
  double best = -1.0;
  int best_bn = 128;
  for (int i = 0; i < 4; ++i) {
    const int bn = cands[i];
    const int64_t nt = (n_rows + bn - 1) / bn;
    const double eff_n = double(n_rows) / double(nt * bn);
    const int64_t tiles = nt * m_tiles;
    const int64_t waves = (tiles + sms - 1) / sms;
    const double eff_w = double(tiles) / double(waves * sms);
    // wide tiles halve the shared-memory operand traffic per flop
    const double eff_t = bn >= 256 ? 1.0 : bn >= 160 ? 0.97 : bn >= 128 ? 0.93 : 0.80;
    const double score = eff_n * eff_w * eff_t;
    if (score > best) {
      best = score;
      best_bn = bn;
    }
  }
  (void)geglu;
  return best_bn;
}

}  // namespace t2v

extern "C" int t2v_gemm(const T2VGemmDesc* d, t2v_stream_t stream_) {
  using namespace t2v;
  if (!d) return fail(-1, "t2v_gemm: null descriptor");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!d->a[0] || !d->b || !d->out) return fail(-2, "t2v_gemm: null a/b/out pointer");
  if (d->n_taps < 1 || d->n_taps > T2V_MAX_TAPS) return fail(-3, "t2v_gemm: n_taps=%d", d->n_taps);
  if (d->a_ch[0] <= 0 || d->a_ch[0] % 64 != 0 || d->a_ch[1] < 0 || d->a_ch[1] % 64 != 0)
    return fail(-4, "t2v_gemm: a_ch must be positive multiples of 64 (got %d, %d)", d->a_ch[0], d->a_ch[1]);
  if (d->a_ch[1] > 0 && !d->a[1]) return fail(-5, "t2v_gemm: a[1] is null but a_ch[1]=%d", d->a_ch[1]);
  int64_t rows_in_box = 1;
  for (int j = 0; j < 4; ++j) {
    if (d->box[j] < 1 || d->a_size[j] < 1 || d->o_size[j] < 1)
      return fail(-6, "t2v_gemm: box/a_size/o_size[%d] must be >= 1", j);
    rows_in_box *= d->box[j];
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
  if (bn == 0) bn = choose_block_n(d->b_rows, m_tiles, sms, geglu);
  if (bn != 64 && bn != 128 && bn != 160 && bn != 256) return fail(-13, "t2v_gemm: block_n=%d unsupported", bn);
  p.n_tiles_n = int((d->b_rows + bn - 1) / bn);
  const int64_t num_tiles = m_tiles * p.n_tiles_n;
  if (num_tiles > 0x7fffffff) return fail(-14, "t2v_gemm: too many tiles");
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
  // vector epilogue: 16-byte aligned rows
  const int out_el = (d->flags & T2V_EPI_OUT_F32) ? 4 : 2;
  bool vec_ok = (d->n_out % 8 == 0) && ((reinterpret_cast<uintptr_t>(d->out) & 15) == 0);
  for (int j = 0; j < 4; ++j) vec_ok = vec_ok && ((d->o_stride[j] * out_el) % 16 == 0);
  if (d->residual) {
    vec_ok = vec_ok && ((reinterpret_cast<uintptr_t>(d->residual) & 15) == 0);
    for (int j = 0; j < 4; ++j) vec_ok = vec_ok && (d->r_stride[j] % 8 == 0);
  }
  p.vec_ok = vec_ok ? 1 : 0;

  // tensor maps
  CUtensorMap tmA[2], tmB;
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
    uint64_t dims[3] = {uint64_t(K), uint64_t(d->b_rows), uint64_t(d->b_batches < 1 ? 1 : d->b_batches)};
    const uint64_t row_stride = uint64_t(d->b_row_stride > 0 ? d->b_row_stride : K);
    uint64_t strides[3] = {2, row_stride * 2,
                           uint64_t(d->b_batch_stride > 0 ? d->b_batch_stride : row_stride * d->b_rows) * 2};
    uint32_t box[3] = {64, uint32_t(bn), 1};
    int rc = make_tmap_bf16(&tmB, d->b, 3, dims, strides, box, "t2v_gemm B");
    if (rc) return rc;
  }
  switch (bn) {
    case 64: return launch_gemm<64>(tmA[0], tmA[1], tmB, p, stream);
    case 128: return launch_gemm<128>(tmA[0], tmA[1], tmB, p, stream);
    case 160: return launch_gemm<160>(tmA[0], tmA[1], tmB, p, stream);
    default: return launch_gemm<256>(tmA[0], tmA[1], tmB, p, stream);
  }
}
