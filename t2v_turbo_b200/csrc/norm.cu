// GroupNorm(+SiLU) and LayerNorm over channels-last bf16 activations (HBM-bound passes).
//
// GroupNorm has two implementations.  When one sample fits the shared memory of a thread-block cluster (every
// UNet GroupNorm except the widest concatenated ones) ONE kernel does everything: the cluster's CTAs bulk-copy
// their row slabs into shared memory, exchange per-group partial sums through distributed shared memory, and
// normalise out of shared memory — the activation is read from HBM/L2 once and no workspace is touched.
// Otherwise (VAE frames: up to 160k rows x 512 channels per sample) it
// is two kernels: a statistics pass (per-(sample, group) sum / sum-of-squares, fp32
// accumulation in registers -> shared -> one global atomic per group per block) and an apply pass
// that folds mean / rstd / gamma / beta into one fma per element followed by SiLU.  Both passes
// give each thread a FIXED 8-channel column (16-byte vector loads, fully coalesced across a row)
// and walk rows, so the per-channel scale / shift live in registers.
// The input may be the channel concatenation of two tensors (UNet skip connections).
#include <cuda_bf16.h>

#include "../../include/t2v_b200.h"
#include "host_common.h"
#include "ptx.cuh"

namespace t2v {

struct GnParams {
  const __nv_bfloat16* x0;
  const __nv_bfloat16* x1;
  int32_t ncv0;  // 8-channel vectors in source 0
  int32_t ncv;   // total 8-channel vectors
  int64_t rs0, rs1;
  __nv_bfloat16* out;
  int64_t out_rs;
  const float* gamma;
  const float* beta;
  int64_t rows_per_sample;
  int32_t rows_per_block;
  int32_t cpg;     // channels per group
  int32_t groups;
  float eps;
  int32_t silu;
  float* ws;
  unsigned* ticket;  // lives right after the sums in the workspace
  // statistics supplied by the producing GEMMs (T2VGemmDesc.col_accum): per source fp32 [samples*cs_group][ch][2]
  const float* cs0;
  const float* cs1;
  int32_t cs_group;  // producer samples per GroupNorm sample (temporal GroupNorm over per-frame sums: t)
};

__device__ __forceinline__ void unpack8(const uint4& v, float (&f)[8]) {
  f[0] = bf16_lo(v.x); f[1] = bf16_hi(v.x);
  f[2] = bf16_lo(v.y); f[3] = bf16_hi(v.y);
  f[4] = bf16_lo(v.z); f[5] = bf16_hi(v.z);
  f[6] = bf16_lo(v.w); f[7] = bf16_hi(v.w);
}

constexpr int kGnThreads = 256;

// grid: (blocks_per_sample, n_samples)
__global__ void __launch_bounds__(kGnThreads) gn_stats_kernel(const GnParams p) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float s_acc[2 * 64];
  for (int i = threadIdx.x; i < 2 * p.groups; i += blockDim.x) s_acc[i] = 0.f;
  __syncthreads();
  const int sample = blockIdx.y;
  const int64_t row_begin = int64_t(blockIdx.x) * p.rows_per_block;
  int64_t row_end = row_begin + p.rows_per_block;
  if (row_end > p.rows_per_sample) row_end = p.rows_per_sample;
  const int tpr = p.ncv < kGnThreads ? p.ncv : kGnThreads;  // threads per row
  const int rpp = kGnThreads / tpr;                         // rows per pass
  const int rr = threadIdx.x / tpr;
  const int64_t base_row = int64_t(sample) * p.rows_per_sample;
  if (rr < rpp) {
    for (int cv = threadIdx.x % tpr; cv < p.ncv; cv += tpr) {
      const bool src1 = cv >= p.ncv0;
      const __nv_bfloat16* xp = src1 ? p.x1 + (cv - p.ncv0) * 8 : p.x0 + cv * 8;
      const int64_t rs = src1 ? p.rs1 : p.rs0;
      float s[8], ss[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) s[j] = ss[j] = 0.f;
      // 4 independent 16-byte loads in flight per thread (memory-level parallelism)
      for (int64_t r = row_begin + rr; r < row_end; r += 4 * rpp) {
        uint4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int64_t ru = r + int64_t(u) * rpp;
          v[u] = ru < row_end ? __ldg(reinterpret_cast<const uint4*>(xp + (base_row + ru) * rs)) : make_uint4(0, 0, 0, 0);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          float f[8];
          unpack8(v[u], f);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            s[j] += f[j];
            ss[j] += f[j] * f[j];
          }
        }
      }
      // merge runs of channels that fall into the same group
      int g_prev = (cv * 8) / p.cpg;
      float a = 0.f, b = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int g = (cv * 8 + j) / p.cpg;
        if (g != g_prev) {
          atomicAdd(&s_acc[2 * g_prev], a);
          atomicAdd(&s_acc[2 * g_prev + 1], b);
          a = b = 0.f;
          g_prev = g;
        }
        a += s[j];
        b += ss[j];
      }
      atomicAdd(&s_acc[2 * g_prev], a);
      atomicAdd(&s_acc[2 * g_prev + 1], b);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * p.groups; i += blockDim.x)
    atomicAdd(&p.ws[int64_t(sample) * 2 * p.groups + i], s_acc[i]);
}

// The statistics workspace is SELF-CLEANING: every apply block copies its sample's sums to shared memory, then
// takes a ticket; the block that draws the last ticket knows every block has read the sums and zeroes them (and
// the ticket counter), so the next GroupNorm call finds a zeroed workspace without a memset / zero kernel.
//
// SiLU(y) = y * sigmoid(y) = h + h * tanh(h) with h = y / 2: the 1/2 is folded into the per-channel scale / shift, so an
// element costs FFMA + MUFU.TANH + FFMA (was FFMA, FMUL, MUFU.EX2, FADD, MUFU.RCP, FMUL: the XU pipe sat at 64 % and
// the kernel at 4.4 TB/s of the 6.5 TB/s the statistics pass reaches; profiles/r02_groupnorm.md).  tanh.approx is good
// to 2^-11 relative, i.e. an absolute error <= |h| * 4.9e-4 -- below the bf16 rounding of the output for y > -3 and
// under 1.3e-3 absolute in the negative tail.
template <bool kSilu>
__global__ void __launch_bounds__(kGnThreads) gn_apply_kernel(const GnParams p) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float s_ws[2 * 64];
  __shared__ int s_last;
  if (p.cs0 != nullptr) {
    // group statistics from the per-channel sums the producing GEMMs accumulated: no statistics pass, no workspace
    for (int i = threadIdx.x; i < 2 * p.groups; i += blockDim.x) s_ws[i] = 0.f;
    __syncthreads();
    const int c_total = p.ncv * 8, c0n = p.ncv0 * 8;
    for (int c = threadIdx.x; c < c_total; c += blockDim.x) {
      const bool src1 = c >= c0n;
      const float* cs = src1 ? p.cs1 : p.cs0;
      const int cl = src1 ? c - c0n : c;
      const int cn = src1 ? c_total - c0n : c0n;
      float a = 0.f, b = 0.f;
      for (int k = 0; k < p.cs_group; ++k) {
        const float2 v = __ldcg(reinterpret_cast<const float2*>(cs) + (int64_t(blockIdx.y) * p.cs_group + k) * cn + cl);
        a += v.x;
        b += v.y;
      }
      atomicAdd(&s_ws[2 * (c / p.cpg)], a);
      atomicAdd(&s_ws[2 * (c / p.cpg) + 1], b);
    }
    __syncthreads();
  } else {
    const float* wsg = p.ws + int64_t(blockIdx.y) * 2 * p.groups;
    for (int i = threadIdx.x; i < 2 * p.groups; i += blockDim.x) s_ws[i] = __ldcg(wsg + i);
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence();
      const unsigned total = gridDim.x * gridDim.y;
      const unsigned ticket = atomicAdd(p.ticket, 1u);
      s_last = (ticket == total - 1) ? 1 : 0;
    }
    __syncthreads();
    if (s_last) {
      const int n = 2 * p.groups * int(gridDim.y);
      for (int i = threadIdx.x; i < n; i += blockDim.x) p.ws[i] = 0.f;
      if (threadIdx.x == 0) *p.ticket = 0u;
    }
  }
  const int sample = blockIdx.y;
  const int64_t row_begin = int64_t(blockIdx.x) * p.rows_per_block;
  int64_t row_end = row_begin + p.rows_per_block;
  if (row_end > p.rows_per_sample) row_end = p.rows_per_sample;
  const int tpr = p.ncv < kGnThreads ? p.ncv : kGnThreads;
  const int rpp = kGnThreads / tpr;
  const int rr = threadIdx.x / tpr;
  if (rr >= rpp) return;  // (after the block-wide ticket protocol above)
  const int64_t base_row = int64_t(sample) * p.rows_per_sample;
  const float inv_n = 1.0f / (float(p.rows_per_sample) * float(p.cpg));
  const float* ws = s_ws;
  for (int cv = threadIdx.x % tpr; cv < p.ncv; cv += tpr) {
    float sc[8], sh[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = cv * 8 + j;
      const int g = c / p.cpg;
      const float mean = ws[2 * g] * inv_n;
      float var = ws[2 * g + 1] * inv_n - mean * mean;
      var = var < 0.f ? 0.f : var;
      const float rstd = rsqrtf(var + p.eps);
      const float ga = __ldg(p.gamma + c) * (kSilu ? 0.5f : 1.0f);
      sc[j] = rstd * ga;
      sh[j] = __ldg(p.beta + c) * (kSilu ? 0.5f : 1.0f) - mean * rstd * ga;
    }
    const bool src1 = cv >= p.ncv0;
    const __nv_bfloat16* xp = src1 ? p.x1 + (cv - p.ncv0) * 8 : p.x0 + cv * 8;
    const int64_t rs = src1 ? p.rs1 : p.rs0;
    __nv_bfloat16* op = p.out + cv * 8;
    for (int64_t r = row_begin + rr; r < row_end; r += 4 * rpp) {
      uint4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int64_t ru = r + int64_t(u) * rpp;
        if (ru < row_end) v[u] = __ldg(reinterpret_cast<const uint4*>(xp + (base_row + ru) * rs));
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int64_t ru = r + int64_t(u) * rpp;
        if (ru < row_end) {
          float f[8];
          unpack8(v[u], f);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float y = fmaf(f[j], sc[j], sh[j]);
            if (kSilu) y = fmaf(y, tanh_approx(y), y);
            f[j] = y;
          }
          uint4 o;
          o.x = pack_bf16(f[0], f[1]);
          o.y = pack_bf16(f[2], f[3]);
          o.z = pack_bf16(f[4], f[5]);
          o.w = pack_bf16(f[6], f[7]);
          *reinterpret_cast<uint4*>(op + (base_row + ru) * p.out_rs) = o;
        }
      }
    }
  }
}

// ------------------------------------------------------------------ single-kernel GroupNorm: one cluster per sample
constexpr int kGcThreads = 512;
constexpr int kGcMaxSlab = 200 * 1024;   // per-CTA slab limit (1 CTA / SM); <= 110 KB lets two CTAs share an SM

// grid = n_samples * CL CTAs, cluster = CL CTAs = one sample; CTA `rank` owns rows [rank*rpb, (rank+1)*rpb).
// dynamic smem: [rpb rows x ncv x 16 B slab][2*groups partial sums][2*groups totals][mbarrier]
__global__ void __launch_bounds__(kGcThreads) gn_cluster_kernel(const GnParams p) {
  extern __shared__ __align__(128) uint8_t gsm[];
  const uint32_t cl = cluster_nctarank();
  const uint32_t rank = cluster_ctarank();
  const int sample = blockIdx.x / cl;
  const int64_t row_begin = int64_t(rank) * p.rows_per_block;
  int64_t row_end = row_begin + p.rows_per_block;
  if (row_end > p.rows_per_sample) row_end = p.rows_per_sample;
  const int nrows = row_end > row_begin ? int(row_end - row_begin) : 0;
  const uint32_t row_bytes = uint32_t(p.ncv) * 16u;
  uint4* slab = reinterpret_cast<uint4*>(gsm);
  float* s_acc = reinterpret_cast<float*>(gsm + size_t(p.rows_per_block) * row_bytes);
  float* s_tot = s_acc + 128;
  uint64_t* bar = reinterpret_cast<uint64_t*>(s_tot + 128);
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
  }
  for (int i = threadIdx.x; i < 2 * p.groups; i += blockDim.x) s_acc[i] = 0.f;
  pdl_launch_dependents();
  __syncthreads();
  pdl_wait();
  const int64_t base_row = int64_t(sample) * p.rows_per_sample + row_begin;
  if (threadIdx.x < 32) {
    // one bulk copy per row and source, issued by the 32 lanes of warp 0; all land on one mbarrier
    if (threadIdx.x == 0) mbar_expect_tx(bar, uint32_t(nrows) * row_bytes);
    __syncwarp();
    const uint32_t b0 = uint32_t(p.ncv0) * 16u, b1 = row_bytes - b0;
    if (b1 == 0 && p.rs0 == int64_t(p.ncv0) * 8) {
      // dense rows: the slab is one contiguous range -> a few large copies instead of one per row
      constexpr uint32_t kChunk = 8192;
      const uint32_t total = uint32_t(nrows) * row_bytes;
      const uint8_t* src = reinterpret_cast<const uint8_t*>(p.x0 + base_row * p.rs0);
      for (uint32_t off = threadIdx.x * kChunk; off < total; off += 32 * kChunk)
        bulk_load_1d(gsm + off, src + off, total - off < kChunk ? total - off : kChunk, bar);
    } else {
      for (int r = threadIdx.x; r < nrows; r += 32) {
        uint8_t* dst = gsm + size_t(r) * row_bytes;
        bulk_load_1d(dst, p.x0 + (base_row + r) * p.rs0, b0, bar);
        if (b1) bulk_load_1d(dst + b0, p.x1 + (base_row + r) * p.rs1, b1, bar);
      }
    }
  }
  mbar_wait(bar, 0);

  const int tpr = p.ncv < kGcThreads ? p.ncv : kGcThreads;  // threads per row
  const int rpp = kGcThreads / tpr;                         // rows per pass
  const int rr = threadIdx.x / tpr;
  const int cv0 = threadIdx.x % tpr;
  if (rr < rpp) {
    for (int cv = cv0; cv < p.ncv; cv += tpr) {
      float s[8], ss[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) s[j] = ss[j] = 0.f;
      for (int r = rr; r < nrows; r += rpp) {
        float f[8];
        unpack8(slab[size_t(r) * p.ncv + cv], f);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          s[j] += f[j];
          ss[j] += f[j] * f[j];
        }
      }
      int g_prev = (cv * 8) / p.cpg;
      float a = 0.f, b = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int g = (cv * 8 + j) / p.cpg;
        if (g != g_prev) {
          atomicAdd(&s_acc[2 * g_prev], a);
          atomicAdd(&s_acc[2 * g_prev + 1], b);
          a = b = 0.f;
          g_prev = g;
        }
        a += s[j];
        b += ss[j];
      }
      atomicAdd(&s_acc[2 * g_prev], a);
      atomicAdd(&s_acc[2 * g_prev + 1], b);
    }
  }
  cluster_sync_all();  // every CTA's partial sums are final (release / acquire at cluster scope)
  if (threadIdx.x < 2 * p.groups) {
    float t = 0.f;
    for (uint32_t k = 0; k < cl; ++k) t += ld_dsmem_f32(s_acc + threadIdx.x, k);
    s_tot[threadIdx.x] = t;
  }
  cluster_sync_all();  // totals visible block-wide; no CTA runs ahead (or exits) while a peer still reads its sums

  if (rr >= rpp) return;
  const float inv_n = 1.0f / (float(p.rows_per_sample) * float(p.cpg));
  for (int cv = cv0; cv < p.ncv; cv += tpr) {
    float sc[8], sh[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = cv * 8 + j;
      const int g = c / p.cpg;
      const float mean = s_tot[2 * g] * inv_n;
      float var = s_tot[2 * g + 1] * inv_n - mean * mean;
      var = var < 0.f ? 0.f : var;
      const float rstd = rsqrtf(var + p.eps);
      const float half = p.silu ? 0.5f : 1.0f;  // SiLU through tanh of y / 2, as gn_apply_kernel
      const float ga = __ldg(p.gamma + c) * half;
      sc[j] = rstd * ga;
      sh[j] = __ldg(p.beta + c) * half - mean * rstd * ga;
    }
    __nv_bfloat16* op = p.out + cv * 8 + base_row * p.out_rs;
    for (int r = rr; r < nrows; r += rpp) {
      float f[8];
      unpack8(slab[size_t(r) * p.ncv + cv], f);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float y = fmaf(f[j], sc[j], sh[j]);
        if (p.silu) y = fmaf(y, tanh_approx(y), y);
        f[j] = y;
      }
      uint4 o;
      o.x = pack_bf16(f[0], f[1]);
      o.y = pack_bf16(f[2], f[3]);
      o.z = pack_bf16(f[4], f[5]);
      o.w = pack_bf16(f[6], f[7]);
      *reinterpret_cast<uint4*>(op + int64_t(r) * p.out_rs) = o;
    }
  }
}

// ------------------------------------------------------------------ LayerNorm: one warp per row
struct LnParams {
  const __nv_bfloat16* x;
  int64_t xs;
  __nv_bfloat16* out;
  int64_t os;
  const float* gamma;
  const float* beta;
  int64_t rows;
  int32_t nvec;  // C / 8
  float eps;
  float inv_c;
};

template <int MAXV>  // vectors per lane
__global__ void __launch_bounds__(256) ln_kernel(const LnParams p) {
  pdl_launch_dependents();
  pdl_wait();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t row = int64_t(blockIdx.x) * 8 + warp;
  if (row >= p.rows) return;
  const __nv_bfloat16* xr = p.x + row * p.xs;
  float f[MAXV][8];
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int v = lane + i * 32;
    if (v < p.nvec) {
      const uint4 u = __ldg(reinterpret_cast<const uint4*>(xr + v * 8));
      unpack8(u, f[i]);
#pragma unroll
      for (int j = 0; j < 8; ++j) sum += f[i][j];
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float mean = sum * p.inv_c;
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int v = lane + i * 32;
    if (v < p.nvec) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float d = f[i][j] - mean;
        sq += d * d;
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
  const float rstd = rsqrtf(sq * p.inv_c + p.eps);
  __nv_bfloat16* orow = p.out + row * p.os;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int v = lane + i * 32;
    if (v < p.nvec) {
      const float4 g0 = __ldg(reinterpret_cast<const float4*>(p.gamma + v * 8));
      const float4 g1 = __ldg(reinterpret_cast<const float4*>(p.gamma + v * 8 + 4));
      const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.beta + v * 8));
      const float4 b1 = __ldg(reinterpret_cast<const float4*>(p.beta + v * 8 + 4));
      const float g[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
      const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
      float y[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) y[j] = (f[i][j] - mean) * rstd * g[j] + b[j];
      uint4 o;
      o.x = pack_bf16(y[0], y[1]);
      o.y = pack_bf16(y[2], y[3]);
      o.z = pack_bf16(y[4], y[5]);
      o.w = pack_bf16(y[6], y[7]);
      *reinterpret_cast<uint4*>(orow + v * 8) = o;
    }
  }
}

// statistics only: one warp per row -> (rstd, -rstd * mean); two-pass in registers like ln_kernel
template <int MAXV>
__global__ void __launch_bounds__(256) ln_stats_kernel(const LnParams p, float2* __restrict__ stats) {
  pdl_launch_dependents();
  pdl_wait();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t row = int64_t(blockIdx.x) * 8 + warp;
  if (row >= p.rows) return;
  const __nv_bfloat16* xr = p.x + row * p.xs;
  float f[MAXV][8];
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int v = lane + i * 32;
    if (v < p.nvec) {
      const uint4 u = __ldg(reinterpret_cast<const uint4*>(xr + v * 8));
      unpack8(u, f[i]);
#pragma unroll
      for (int j = 0; j < 8; ++j) sum += f[i][j];
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float mean = sum * p.inv_c;
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int v = lane + i * 32;
    if (v < p.nvec) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float d = f[i][j] - mean;
        sq += d * d;
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
  if (lane == 0) {
    const float rstd = rsqrtf(sq * p.inv_c + p.eps);
    stats[row] = make_float2(rstd, -rstd * mean);
  }
}

}  // namespace t2v

extern "C" int t2v_groupnorm(const T2VGroupNormDesc* d, t2v_stream_t stream_) {
  using namespace t2v;
  if (!d) return fail(-1, "t2v_groupnorm: null descriptor");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const int C = d->ch[0] + d->ch[1];
  if (!d->x[0] || !d->out || !d->gamma || !d->beta || (!d->workspace && !d->chan_sums[0]))
    return fail(-2, "t2v_groupnorm: null pointer");
  if (d->ch[0] <= 0 || d->ch[0] % 8 || d->ch[1] < 0 || d->ch[1] % 8)
    return fail(-3, "t2v_groupnorm: channel counts must be multiples of 8");
  if (d->ch[1] > 0 && !d->x[1]) return fail(-4, "t2v_groupnorm: x[1] null");
  if (d->groups < 1 || d->groups > 64 || C % d->groups) return fail(-5, "t2v_groupnorm: bad groups");
  if (d->rows_per_sample < 1 || d->rows % d->rows_per_sample) return fail(-6, "t2v_groupnorm: rows %% rows_per_sample != 0");
  if (d->x_row_stride[0] % 8 || d->out_row_stride % 8 || (d->ch[1] && d->x_row_stride[1] % 8))
    return fail(-7, "t2v_groupnorm: row strides must be multiples of 8 elements");
  const int64_t n_samples = d->rows / d->rows_per_sample;
  if (n_samples > 65535) return fail(-8, "t2v_groupnorm: too many samples");
  GnParams p;
  p.x0 = static_cast<const __nv_bfloat16*>(d->x[0]);
  p.x1 = static_cast<const __nv_bfloat16*>(d->x[1]);
  p.ncv0 = d->ch[0] / 8;
  p.ncv = C / 8;
  p.rs0 = d->x_row_stride[0];
  p.rs1 = d->x_row_stride[1];
  p.out = static_cast<__nv_bfloat16*>(d->out);
  p.out_rs = d->out_row_stride;
  p.gamma = d->gamma;
  p.beta = d->beta;
  p.rows_per_sample = d->rows_per_sample;
  p.cpg = C / d->groups;
  p.groups = d->groups;
  p.eps = d->eps;
  p.silu = d->silu;
  p.ws = d->workspace;
  p.cs0 = d->chan_sums[0];
  p.cs1 = d->chan_sums[1];
  p.cs_group = d->chan_group < 1 ? 1 : d->chan_group;
  const bool have_sums = d->chan_sums[0] != nullptr;
  if (have_sums && d->ch[1] > 0 && !d->chan_sums[1]) return fail(-10, "t2v_groupnorm: chan_sums[1] missing for the second source");
  if (!have_sums) p.cs0 = p.cs1 = nullptr;
  int sms = num_sms();
  if (sms <= 0) return fail(-110, "t2v_groupnorm: no CUDA device");
  // ---- single-kernel path: one thread-block cluster per sample, the sample lives in the cluster's shared memory
  // Measured on B200 (scripts/gn_bench.py): the cluster kernel wins only while the whole tensor is small (its load /
  // reduce / store phases run in lock-step across the GPU); above ~4 MB the statistics + apply pair streams better.
  const bool small = d->rows * int64_t(C) * 2 <= (4 << 20);
  if (!have_sums && (d->mode == 2 || (d->mode == 0 && small))) {
    static int max_cl = -1;  // largest usable cluster size (16 needs the non-portable opt-in)
    if (max_cl < 0) {
      cudaError_t e1 = cudaFuncSetAttribute(gn_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kGcMaxSlab + 2048);
      cudaError_t e2 = cudaFuncSetAttribute(gn_cluster_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
      max_cl = (e1 == cudaSuccess) ? (e2 == cudaSuccess ? 16 : 8) : 0;
      (void)cudaGetLastError();
    }
    const int64_t row_bytes = int64_t(C) * 2;
    const bool aligned = (reinterpret_cast<uintptr_t>(d->x[0]) % 16 == 0) &&
                         (d->ch[1] == 0 || reinterpret_cast<uintptr_t>(d->x[1]) % 16 == 0);
    while (max_cl >= 1 && aligned) {
      int cl = max_cl;
      while (cl > 1 && (d->rows_per_sample + cl - 1) / cl < 8) cl >>= 1;  // keep >= 8 rows per CTA
      const int64_t rpb = (d->rows_per_sample + cl - 1) / cl;
      const int64_t slab = rpb * row_bytes;
      if (slab > kGcMaxSlab || n_samples * cl > 0x7fffffff) break;
      p.rows_per_block = int(rpb);
      p.ticket = nullptr;
      const size_t smem = size_t(slab) + 2 * 128 * sizeof(float) + 16;
      cudaError_t e = launch_kernel_cluster(gn_cluster_kernel, dim3(unsigned(n_samples * cl)), dim3(kGcThreads), smem,
                                            stream, unsigned(cl), p);
      if (e == cudaSuccess) return 0;
      (void)cudaGetLastError();
      if (cl < max_cl) break;  // the failure is not about the cluster size
      max_cl = cl / 2;         // this device cannot co-schedule clusters of `cl` CTAs: never ask again
    }
    if (d->mode == 2) return fail(-9, "t2v_groupnorm: the single-kernel cluster path is not available for this shape");
  }
  const int tpr = p.ncv < kGnThreads ? p.ncv : kGnThreads;
  const int rpp = kGnThreads / tpr;
  // 8 blocks per SM (two waves of the 4 resident ones): measured 1-12 % faster than one wave on every shape of the
  // bs = 8 step; 48-register variants (5 resident blocks) spill and lose 15-20 % (profiles/r02_groupnorm.md)
  int64_t want_blocks = (int64_t(sms) * 8 + n_samples - 1) / n_samples;
  if (want_blocks < 1) want_blocks = 1;
  int64_t rpb = (d->rows_per_sample + want_blocks - 1) / want_blocks;
  const int64_t min_rpb = int64_t(rpp) * 4;
  if (rpb < min_rpb) rpb = min_rpb;
  rpb = (rpb + rpp - 1) / rpp * rpp;
  p.rows_per_block = int(rpb);
  const int64_t bps = (d->rows_per_sample + rpb - 1) / rpb;
  // workspace layout: [n_samples][groups][2] sums + one ticket word; zero on entry (self-cleaning, see gn_apply_kernel)
  p.ticket = reinterpret_cast<unsigned*>(d->workspace + 2 * d->groups * n_samples);
  cudaError_t e;
  dim3 grid((unsigned)bps, (unsigned)n_samples);
  if (!have_sums) launch_kernel(gn_stats_kernel, dim3(grid), dim3(kGnThreads), 0, stream, p);
  if (p.silu) launch_kernel(gn_apply_kernel<true>, dim3(grid), dim3(kGnThreads), 0, stream, p);
  else launch_kernel(gn_apply_kernel<false>, dim3(grid), dim3(kGnThreads), 0, stream, p);
  e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "t2v_groupnorm launch");
  return 0;
}

extern "C" int t2v_layernorm(const T2VLayerNormDesc* d, t2v_stream_t stream_) {
  using namespace t2v;
  if (!d) return fail(-1, "t2v_layernorm: null descriptor");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!d->x || !d->out || !d->gamma || !d->beta) return fail(-2, "t2v_layernorm: null pointer");
  if (d->channels <= 0 || d->channels % 8 || d->channels > 2048)
    return fail(-3, "t2v_layernorm: channels must be a multiple of 8, <= 2048");
  if (d->x_row_stride % 8 || d->out_row_stride % 8) return fail(-4, "t2v_layernorm: strides must be multiples of 8");
  if (d->rows <= 0) return 0;
  LnParams p;
  p.x = static_cast<const __nv_bfloat16*>(d->x);
  p.xs = d->x_row_stride;
  p.out = static_cast<__nv_bfloat16*>(d->out);
  p.os = d->out_row_stride;
  p.gamma = d->gamma;
  p.beta = d->beta;
  p.rows = d->rows;
  p.nvec = d->channels / 8;
  p.eps = d->eps;
  p.inv_c = 1.0f / float(d->channels);
  const unsigned grid = unsigned((d->rows + 7) / 8);
  const int vpl = (p.nvec + 31) / 32;
  if (vpl <= 2) launch_kernel(ln_kernel<2>, dim3(grid), dim3(256), 0, stream, p);
  else if (vpl <= 5) launch_kernel(ln_kernel<5>, dim3(grid), dim3(256), 0, stream, p);
  else launch_kernel(ln_kernel<8>, dim3(grid), dim3(256), 0, stream, p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "t2v_layernorm launch");
  return 0;
}

extern "C" int t2v_layernorm_stats(const T2VLayerNormDesc* d, float* stats, t2v_stream_t stream_) {
  using namespace t2v;
  if (!d || !d->x || !stats) return fail(-1, "t2v_layernorm_stats: null pointer");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (d->channels <= 0 || d->channels % 8 || d->channels > 2048)
    return fail(-3, "t2v_layernorm_stats: channels must be a multiple of 8, <= 2048");
  if (d->x_row_stride % 8 || (reinterpret_cast<uintptr_t>(stats) & 7)) return fail(-4, "t2v_layernorm_stats: bad stride / alignment");
  if (d->rows <= 0) return 0;
  LnParams p;
  p.x = static_cast<const __nv_bfloat16*>(d->x);
  p.xs = d->x_row_stride;
  p.out = nullptr;
  p.os = 0;
  p.gamma = p.beta = nullptr;
  p.rows = d->rows;
  p.nvec = d->channels / 8;
  p.eps = d->eps;
  p.inv_c = 1.0f / float(d->channels);
  const unsigned grid = unsigned((d->rows + 7) / 8);
  const int vpl = (p.nvec + 31) / 32;
  float2* st = reinterpret_cast<float2*>(stats);
  if (vpl <= 2) launch_kernel(ln_stats_kernel<2>, dim3(grid), dim3(256), 0, stream, p, st);
  else if (vpl <= 5) launch_kernel(ln_stats_kernel<5>, dim3(grid), dim3(256), 0, stream, p, st);
  else launch_kernel(ln_stats_kernel<8>, dim3(grid), dim3(256), 0, stream, p, st);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "t2v_layernorm_stats launch");
  return 0;
}
