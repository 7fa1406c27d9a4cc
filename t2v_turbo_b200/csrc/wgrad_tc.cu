// Weight-gradient GEMM on tcgen05 (sm_100a) for the LoRA parameters of the consistency-distillation step
// (utils/lora.py:19-230: lora_down has the base layer's kernel, lora_up is 1x1; train_t2v_turbo_v1_lora.py:1190).
//
//   out[j, c, tap] += alpha * sum_points  A[point + off(tap), c] * B[point, j]        c < a_ch,  j < b_cols <= 64
//
// Both operands are read IN PLACE from their natural channels-last layouts as MN-MAJOR UMMA operands: the reduction
// runs over points (tokens / pixels), and for a fixed point the channels of A and the columns of B are contiguous —
// exactly what a TMA box {64 channels, <=128 points} delivers.  No transposed copies of activations or gradients, no
// im2col: a tap is a coordinate offset of the A box, the conv padding is TMA out-of-bounds zero fill (the same tensor-map
// machinery as the forward implicit GEMM, gemm_tc.cu).
//   lora_down.weight grad : A = layer input x (taps of the base kernel), B = d(lora_down output)   [r, Cin, taps]
//   lora_up.weight grad   : A = scale * mask * dy,                        B = lora_down output      [Cout, r]
// One CTA = (128-channel block, tap, slice of the point tiles); fp32 accumulator [128 x 64] in TMEM; the partial sums of
// the slices are added into the caller's fp32 buffer — the LoRA-gradient ARENA — with red.global.add (the buffer is
// accumulated into, never overwritten: zero it once per optimizer step).
//   warp 0 : TMA producer (two 64-channel boxes of A + one box of B per stage, 4-stage ring)
//   warp 1 : MMA issuer (M128 N64 K16, A and B MN-major)     warps 2..5 : epilogue (thread = channel row)
#include <cuda.h>
#include <string.h>

#include "../../include/t2v_b200.h"
#include "host_common.h"
#include "ptx.cuh"

namespace t2v {

constexpr int kWgThreads = 192;
constexpr int kWgStages = 4;
constexpr int kWgABytes = 2 * 128 * 64 * 2;   // two 64-channel blocks of up to 128 points
constexpr int kWgBBytes = 128 * 64 * 2;
constexpr int kWgStageBytes = kWgABytes + kWgBBytes;
constexpr int kWgSmem = kWgStages * kWgStageBytes + 256 + 1024;

struct WgradParams {
  int32_t box[4], ntile[4];
  int32_t rows_in_box, n_point_tiles, n_cblocks, n_taps, tiles_per_slice, n_slices;
  int32_t tap_off[T2V_MAX_TAPS][4];
  int32_t a_ch, b_cols;
  float* out;
  int64_t out_j_stride, out_c_stride, out_tap_stride;
  float alpha;
};

// MN-major operand tile [k][64 bf16] with 128-byte swizzle: 8-k atoms of 1024 B (SBO); blocks of 64 MN elements LBO apart
__device__ __forceinline__ uint64_t umma_desc_sw128_mn(uint32_t smem_addr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

__global__ void __launch_bounds__(kWgThreads, 1)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const WgradParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kWgStages * kWgStageBytes);
  uint64_t* empty_bar = full_bar + kWgStages;
  uint64_t* acc_bar = empty_bar + kWgStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_bar + 1);

  const int warp = __shfl_sync(0xffffffffu, int(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  // blockIdx.x = (slice * n_taps + tap) * n_cblocks + cblock
  const int cb = blockIdx.x % p.n_cblocks;
  const int tap = (blockIdx.x / p.n_cblocks) % p.n_taps;
  const int slice = blockIdx.x / (p.n_cblocks * p.n_taps);
  const int t_begin = slice * p.tiles_per_slice;
  const int t_end = min(t_begin + p.tiles_per_slice, p.n_point_tiles);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < kWgStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(acc_bar, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 64);
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
    const uint32_t a_bytes = uint32_t(p.rows_in_box) * 128u;
    int stage = 0;
    uint32_t phase = 0;
    for (int t = t_begin; t < t_end; ++t) {
      int m = t, o[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        o[j] = (j < 3 ? m % p.ntile[j] : m) * p.box[j];
        m /= p.ntile[j];
      }
      mbar_wait_relaxed(&empty_bar[stage], phase ^ 1u);
      uint8_t* sA = smem + stage * kWgStageBytes;
      mbar_expect_tx(&full_bar[stage], 3u * a_bytes);
      const int c1 = o[0] + p.tap_off[tap][0], c2 = o[1] + p.tap_off[tap][1], c3 = o[2] + p.tap_off[tap][2],
                c4 = o[3] + p.tap_off[tap][3];
      tma_load_5d(sA, &tmA, &full_bar[stage], cb * 128, c1, c2, c3, c4);
      tma_load_5d(sA + 128 * 128, &tmA, &full_bar[stage], cb * 128 + 64, c1, c2, c3, c4);
      tma_load_5d(sA + kWgABytes, &tmB, &full_bar[stage], 0, o[0], o[1], o[2], o[3]);
      if (++stage == kWgStages) {
        stage = 0;
        phase ^= 1u;
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer: D[128 ch x 64] += A^T (MN-major) * B (MN-major)
    // whole-warp loop (uniform-datapath descriptor arithmetic), tcgen05.mma / commit on the elected lane only: an N = 64 MMA is
    // 32 tensor cycles, less than the ~12 dependent instructions per MMA a single-thread issue loop needs
    const bool leader = elect_one();
    constexpr uint32_t idesc = umma_idesc_bf16(128, 64, 1, 1);
    const uint32_t tm = __shfl_sync(0xffffffffu, tmem_base, 0);
    const int k_steps = p.rows_in_box >> 4;
    const uint64_t adesc0 = umma_desc_sw128_mn(smem_u32(smem), 128 * 128);
    const uint64_t bdesc0 = umma_desc_sw128_mn(smem_u32(smem) + kWgABytes, 128 * 128);
    constexpr uint64_t kStageDesc = kWgStageBytes >> 4;   // one ring stage further, in descriptor (16-byte) units
    int stage = 0;
    uint32_t phase = 0;
    for (int t = t_begin; t < t_end; ++t) {
      mbar_wait(&full_bar[stage], phase);
      tc_fence_after();
      const uint64_t adesc = adesc0 + uint64_t(stage) * kStageDesc;
      const uint64_t bdesc = bdesc0 + uint64_t(stage) * kStageDesc;
      if (leader) {
        for (int kk = 0; kk < k_steps; ++kk)   // 16 points = 16 rows of 128 bytes
          umma_ss(tm, adesc + uint64_t(kk * (2048 >> 4)), bdesc + uint64_t(kk * (2048 >> 4)), idesc, (t > t_begin || kk > 0) ? 1u : 0u);
        umma_commit(&empty_bar[stage]);
      }
      if (++stage == kWgStages) {
        stage = 0;
        phase ^= 1u;
      }
    }
    if (leader) umma_commit(acc_bar);
  } else if (warp >= 2 && t_end > t_begin) {
    // ------------------------------------------------------------ epilogue: thread = channel row, red.add into the arena
    const int lg = warp & 3;               // TMEM lane group of this warp
    const int c = cb * 128 + lg * 32 + lane;
    mbar_wait(acc_bar, 0);
    tc_fence_after();
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(lg * 32) << 16);
    float* orow = p.out + int64_t(c) * p.out_c_stride + int64_t(tap) * p.out_tap_stride;
#pragma unroll 1
    for (int j0 = 0; j0 < 64; j0 += 32) {
      uint32_t v[32];
      tmem_ld_32x32(taddr + j0, v);
      tmem_wait_ld();
      if (c < p.a_ch) {
        if (p.out_j_stride == 1 && ((reinterpret_cast<uintptr_t>(orow + j0) & 15) == 0) && j0 + 32 <= p.b_cols) {
#pragma unroll
          for (int q = 0; q < 8; ++q)
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(orow + j0 + 4 * q),
                         "f"(p.alpha * __uint_as_float(v[4 * q])), "f"(p.alpha * __uint_as_float(v[4 * q + 1])),
                         "f"(p.alpha * __uint_as_float(v[4 * q + 2])), "f"(p.alpha * __uint_as_float(v[4 * q + 3]))
                         : "memory");
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (j0 + j < p.b_cols) atomicAdd(orow + int64_t(j0 + j) * p.out_j_stride, p.alpha * __uint_as_float(v[j]));
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 64);
  }
}

}  // namespace t2v

extern "C" int t2v_wgrad(const T2VWgradDesc* d, t2v_stream_t stream_) {
  using namespace t2v;
  if (!d || !d->a || !d->b || !d->out) return fail(-1, "t2v_wgrad: null pointer");
  if (d->a_ch < 8 || d->a_ch % 8 || d->b_cols < 8 || d->b_cols > 64 || d->b_cols % 8)
    return fail(-2, "t2v_wgrad: a_ch must be a multiple of 8, b_cols a multiple of 8 in [8, 64] (got %d, %d)", d->a_ch, d->b_cols);
  if (d->n_taps < 1 || d->n_taps > T2V_MAX_TAPS) return fail(-3, "t2v_wgrad: n_taps=%d", d->n_taps);
  int64_t rows = 1, tiles = 1;
  WgradParams p;
  memset(&p, 0, sizeof(p));
  for (int j = 0; j < 4; ++j) {
    if (d->box[j] < 1 || d->o_size[j] < 1 || d->a_size[j] < 1) return fail(-4, "t2v_wgrad: box / sizes must be >= 1");
    rows *= d->box[j];
    p.box[j] = d->box[j];
    p.ntile[j] = int((d->o_size[j] + d->box[j] - 1) / d->box[j]);
    tiles *= p.ntile[j];
  }
  if (rows > 128 || rows % 16) return fail(-5, "t2v_wgrad: box product %lld must be a multiple of 16 and <= 128", (long long)rows);
  if (tiles > 0x7fffffff) return fail(-6, "t2v_wgrad: too many point tiles");
  int sms = num_sms();
  if (sms <= 0) return fail(-110, "t2v_wgrad: no CUDA device");
  p.rows_in_box = int(rows);
  p.n_point_tiles = int(tiles);
  p.n_cblocks = (d->a_ch + 127) / 128;
  p.n_taps = d->n_taps;
  for (int t = 0; t < d->n_taps; ++t)
    for (int j = 0; j < 4; ++j) p.tap_off[t][j] = d->tap_off[t][j];
  // slices of the point range: about two CTAs per SM in total, at least 4 tiles per slice (amortises the pipeline fill)
  int64_t want = (2 * int64_t(sms) + p.n_cblocks * p.n_taps - 1) / (int64_t(p.n_cblocks) * p.n_taps);
  if (want < 1) want = 1;
  int64_t tps = (tiles + want - 1) / want;
  if (tps < 4) tps = tiles < 4 ? tiles : 4;
  p.tiles_per_slice = int(tps);
  p.n_slices = int((tiles + tps - 1) / tps);
  p.a_ch = d->a_ch;
  p.b_cols = d->b_cols;
  p.out = d->out;
  p.out_j_stride = d->out_j_stride;
  p.out_c_stride = d->out_c_stride;
  p.out_tap_stride = d->out_tap_stride;
  p.alpha = d->alpha;
  CUtensorMap tmA, tmB;
  {
    uint64_t dims[5], strides[5];
    uint32_t box[5];
    dims[0] = uint64_t(d->a_ch);
    strides[0] = 2;
    box[0] = 64;
    for (int j = 0; j < 4; ++j) {
      dims[j + 1] = uint64_t(d->a_size[j]);
      strides[j + 1] = uint64_t(d->a_stride[j]) * 2;
      box[j + 1] = uint32_t(d->box[j]);
      if (d->a_size[j] == 1 && strides[j + 1] == 0) strides[j + 1] = 16;
    }
    int rc = make_tmap_bf16(&tmA, d->a, 5, dims, strides, box, "t2v_wgrad A");
    if (rc) return rc;
    dims[0] = uint64_t(d->b_cols);
    for (int j = 0; j < 4; ++j) {
      dims[j + 1] = uint64_t(d->o_size[j]);
      strides[j + 1] = uint64_t(d->b_stride[j]) * 2;
      if (d->o_size[j] == 1 && strides[j + 1] == 0) strides[j + 1] = 16;
    }
    rc = make_tmap_bf16(&tmB, d->b, 5, dims, strides, box, "t2v_wgrad B");
    if (rc) return rc;
  }
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kWgSmem);
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(wgrad_tc)");
    configured = true;
  }
  const int64_t grid = int64_t(p.n_slices) * p.n_taps * p.n_cblocks;
  launch_kernel(wgrad_tc_kernel, dim3(unsigned(grid)), dim3(kWgThreads), kWgSmem, static_cast<cudaStream_t>(stream_), tmA, tmB, p);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : cuda_fail(e, "t2v_wgrad launch");
}
