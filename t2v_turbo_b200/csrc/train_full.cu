// SIMT kernels added for the training paths beyond the v1 LoRA step:
//  * the FULL fine-tune step (train_latent_t2v_turbo_v2.py:945-1276: every UNet parameter trains, an EMA copy is the target network):
//    affine gradients of GroupNorm(+SiLU) and LayerNorm, EMA update of the target parameters (utils/common_utils.py:308-319).
//    Weight gradients reuse t2v_wgrad, bias gradients t2v_colsum_samples, input gradients the v1 step's kernels (train_bwd.cu);
//  * `vae.decode` WITH grad (the reward terms, train_t2v_turbo_v1_lora.py:1043-1099): the softmax adjoint of the KL-VAE AttnBlock.
// All HBM bound: one pass over the operands, fp32 accumulation, per-CTA shared-memory reduction, one global atomic per channel per
// CTA into the fp32 gradient arena.  No TMA / tcgen05 here, so this file also compiles for the host emulation (tests/cuda_emu).
#include <math.h>

#include "../../include/t2v_b200.h"
#ifdef T2V_HOST_EMU   // tests/cuda_emu: the SIMT kernels of this file compiled by g++ and run on CPU threads (test infrastructure only)
#include "cuda_emu.h"
#else
#include <cuda_bf16.h>

#include "host_common.h"
#include "ptx.cuh"
#endif

namespace t2v {

namespace {

constexpr int kThreads = 256;
constexpr int kMaxGnChannels = 2560;

__device__ __forceinline__ void unpack8(const uint4& v, float (&f)[8]) {
  f[0] = bf16_lo(v.x); f[1] = bf16_hi(v.x);
  f[2] = bf16_lo(v.y); f[3] = bf16_hi(v.y);
  f[4] = bf16_lo(v.z); f[5] = bf16_hi(v.z);
  f[6] = bf16_lo(v.w); f[7] = bf16_hi(v.w);
}

// ------------------------------------------------------------------------------------------------ GroupNorm affine gradients
// y = act(xh * gamma + beta), xh = (x - mean_g) * rstd_g.   dpre = dy * act'(pre);   dgamma[c] += sum dpre * xh,  dbeta[c] += sum dpre
// over every row of every sample.  The statistics are the (sum x, sum x^2) the input-gradient kernel (t2v_groupnorm_bwd) left in
// its workspace: fp32 [n_samples][groups][4], slots 0 and 1.  Thread mapping = gn_bwd_kernel's (8-channel column per thread).
struct GnAffParams {
  const __nv_bfloat16* x; int64_t x_rs;
  const __nv_bfloat16* dy; int64_t dy_rs;
  const float* gamma; const float* beta;
  float* dgamma; float* dbeta;
  int64_t rows_per_sample;
  int32_t rows_per_block, ncv, cpg, groups, silu;
  float eps;
  const float* ws;
};

__global__ void __launch_bounds__(kThreads) gn_affine_grad_kernel(const GnAffParams p) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float s_g[kMaxGnChannels];
  __shared__ float s_b[kMaxGnChannels];
  __shared__ float s_st[2 * 64];
  const int sample = blockIdx.y;
  const int nch = p.ncv * 8;
  for (int i = threadIdx.x; i < nch; i += blockDim.x) s_g[i] = s_b[i] = 0.f;
  for (int i = threadIdx.x; i < 2 * p.groups; i += blockDim.x)
    s_st[i] = __ldcg(p.ws + (int64_t(sample) * p.groups + (i >> 1)) * 4 + (i & 1));
  __syncthreads();
  const int64_t row_begin = int64_t(blockIdx.x) * p.rows_per_block;
  int64_t row_end = row_begin + p.rows_per_block;
  if (row_end > p.rows_per_sample) row_end = p.rows_per_sample;
  const int tpr = p.ncv < kThreads ? p.ncv : kThreads;
  const int rpp = kThreads / tpr;
  const int rr = threadIdx.x / tpr;
  const int64_t base_row = int64_t(sample) * p.rows_per_sample;
  const float inv_n = 1.0f / (float(p.rows_per_sample) * float(p.cpg));
  if (rr < rpp) {
    for (int cv = threadIdx.x % tpr; cv < p.ncv; cv += tpr) {
      float mean[8], rstd[8], ga[8], be[8], ag[8], ab[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int c = cv * 8 + j;
        const int g = c / p.cpg;
        mean[j] = s_st[2 * g] * inv_n;
        float var = s_st[2 * g + 1] * inv_n - mean[j] * mean[j];
        var = var < 0.f ? 0.f : var;
        rstd[j] = rsqrtf(var + p.eps);
        ga[j] = __ldg(p.gamma + c);
        be[j] = __ldg(p.beta + c);
        ag[j] = ab[j] = 0.f;
      }
      for (int64_t r = row_begin + rr; r < row_end; r += rpp) {
        float xf[8], dyf[8];
        unpack8(__ldg(reinterpret_cast<const uint4*>(p.x + (base_row + r) * p.x_rs + cv * 8)), xf);
        unpack8(__ldg(reinterpret_cast<const uint4*>(p.dy + (base_row + r) * p.dy_rs + cv * 8)), dyf);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float xh = (xf[j] - mean[j]) * rstd[j];
          float d = dyf[j];
          if (p.silu) {
            const float pre = fmaf(xh, ga[j], be[j]);
            const float sg = 1.0f / (1.0f + __expf(-pre));
            d *= sg * fmaf(pre, 1.0f - sg, 1.0f);
          }
          ab[j] += d;
          ag[j] = fmaf(d, xh, ag[j]);
        }
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        atomicAdd(&s_g[cv * 8 + j], ag[j]);
        atomicAdd(&s_b[cv * 8 + j], ab[j]);
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < nch; i += blockDim.x) {
    atomicAdd(p.dgamma + i, s_g[i]);
    atomicAdd(p.dbeta + i, s_b[i]);
  }
}

// ------------------------------------------------------------------------------------------------ LayerNorm affine gradients
// y = xh * gamma + beta per row (C = 64 * kNI):  dgamma[c] += sum_rows dy * xh,  dbeta[c] += sum_rows dy.
// One warp per row, the row in registers (ln_bwd_kernel's mapping: lane holds channels i * 64 + 2 * lane + {0, 1}).
template <int kNI>
__global__ void __launch_bounds__(kThreads) ln_affine_grad_kernel(const __nv_bfloat16* __restrict__ x, int64_t x_rs,
                                                                  const __nv_bfloat16* __restrict__ dy, int64_t dy_rs,
                                                                  float* __restrict__ dgamma, float* __restrict__ dbeta, int64_t rows,
                                                                  float eps) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float s_g[64 * kNI];
  __shared__ float s_b[64 * kNI];
  for (int i = threadIdx.x; i < 64 * kNI; i += blockDim.x) s_g[i] = s_b[i] = 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int64_t warps = int64_t(gridDim.x) * (kThreads / 32);
  constexpr float inv_c = 1.0f / float(64 * kNI);
  float ag[2 * kNI], ab[2 * kNI];
#pragma unroll
  for (int i = 0; i < 2 * kNI; ++i) ag[i] = ab[i] = 0.f;
  for (int64_t r = int64_t(blockIdx.x) * (kThreads / 32) + (threadIdx.x >> 5); r < rows; r += warps) {
    float xv[2 * kNI];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < kNI; ++i) {
      const uint32_t u = __ldg(reinterpret_cast<const uint32_t*>(x + r * x_rs) + i * 32 + lane);
      xv[2 * i] = bf16_lo(u);
      xv[2 * i + 1] = bf16_hi(u);
      s += xv[2 * i] + xv[2 * i + 1];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s * inv_c;
    float v = 0.f;
#pragma unroll
    for (int i = 0; i < 2 * kNI; ++i) {
      xv[i] -= mean;
      v = fmaf(xv[i], xv[i], v);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const float rstd = rsqrtf(v * inv_c + eps);
#pragma unroll
    for (int i = 0; i < kNI; ++i) {
      const uint32_t u = __ldg(reinterpret_cast<const uint32_t*>(dy + r * dy_rs) + i * 32 + lane);
      const float d0 = bf16_lo(u), d1 = bf16_hi(u);
      ab[2 * i] += d0;
      ab[2 * i + 1] += d1;
      ag[2 * i] = fmaf(d0, xv[2 * i] * rstd, ag[2 * i]);
      ag[2 * i + 1] = fmaf(d1, xv[2 * i + 1] * rstd, ag[2 * i + 1]);
    }
  }
#pragma unroll
  for (int i = 0; i < kNI; ++i) {
    atomicAdd(&s_g[i * 64 + 2 * lane], ag[2 * i]);
    atomicAdd(&s_g[i * 64 + 2 * lane + 1], ag[2 * i + 1]);
    atomicAdd(&s_b[i * 64 + 2 * lane], ab[2 * i]);
    atomicAdd(&s_b[i * 64 + 2 * lane + 1], ab[2 * i + 1]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 64 * kNI; i += blockDim.x) {
    atomicAdd(dgamma + i, s_g[i]);
    atomicAdd(dbeta + i, s_b[i]);
  }
}

// ------------------------------------------------------------------------------------------------ EMA of the target parameters
// utils/common_utils.py:308-319: targ.mul_(rate).add_(src, alpha = 1 - rate), over the flat fp32 arenas.
__global__ void __launch_bounds__(kThreads) ema_update_kernel(float* __restrict__ target, const float* __restrict__ src, int64_t n4,
                                                              int64_t n, float rate) {
  pdl_launch_dependents();
  pdl_wait();
  const float alpha = 1.0f - rate;
  const int64_t stride = int64_t(gridDim.x) * blockDim.x;
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 t = reinterpret_cast<float4*>(target)[i];
    const float4 s = __ldg(reinterpret_cast<const float4*>(src) + i);
    t.x = fmaf(s.x, alpha, t.x * rate);
    t.y = fmaf(s.y, alpha, t.y * rate);
    t.z = fmaf(s.z, alpha, t.z * rate);
    t.w = fmaf(s.w, alpha, t.w * rate);
    reinterpret_cast<float4*>(target)[i] = t;
  }
  for (int64_t i = n4 * 4 + int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride)
    target[i] = fmaf(src[i], alpha, target[i] * rate);
}

// ------------------------------------------------------------------------------------------------ softmax backward over rows
// P = softmax(scale * S) row-wise (t2v_softmax_rows).  In place on dP:  dS[r, c] = scale * P[r, c] * (dP[r, c] - sum_j dP[r, j] P[r, j]).
// The single-head 512-channel attention of the KL-VAE decoder's mid block (ae_modules.py:48-73), whose backward the reward
// terms of the training scripts need (vae.decode WITH grad, train_t2v_turbo_v1_lora.py:1055-1098).  One CTA per row.
__global__ void __launch_bounds__(kThreads) softmax_bwd_rows_kernel(__nv_bfloat16* __restrict__ dp, int64_t dp_rs,
                                                                    const __nv_bfloat16* __restrict__ p, int64_t p_rs, int cols, float scale) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float red[kThreads / 32];
  const uint16_t* prow = reinterpret_cast<const uint16_t*>(p + int64_t(blockIdx.x) * p_rs);
  uint16_t* drow = reinterpret_cast<uint16_t*>(dp + int64_t(blockIdx.x) * dp_rs);
  float dot = 0.f;
  for (int c = threadIdx.x; c < cols; c += blockDim.x) dot = fmaf(bf16_lo(drow[c]), bf16_lo(prow[c]), dot);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = dot;
  __syncthreads();
  dot = 0.f;
#pragma unroll
  for (int i = 0; i < kThreads / 32; ++i) dot += red[i];
  for (int c = threadIdx.x; c < cols; c += blockDim.x) {
    const float v = scale * bf16_lo(prow[c]) * (bf16_lo(drow[c]) - dot);
    drow[c] = uint16_t(pack_bf16(v, 0.f) & 0xFFFFu);
  }
}

// ------------------------------------------------------------------------------------------------ temporal attention: d(probs) -> dq, dk
// The motion-prior score of the v2 preprocessing (motion_prior_sample.py:59-84) differentiates a loss on the temporal attention
// PROBABILITIES P = softmax(scale q k^T) (exported by t2v_attn_short_fwd) w.r.t. the latents: this is the adjoint of that export,
//   dS = P * (dP - rowsum(dP * P)),   dq = scale * dS k,   dk = scale * dS^T q          per (sequence, head), len <= 16, head dim 64.
// One warp per (sequence, head): q, k staged in shared memory as fp32, lane i owns query row i (softmax row in registers), then
// row i of dq and row i of dk.  Tokens are laid out [(outer, t, inner), heads * 64] like t2v_attn_short_fwd's operands.
constexpr int kPbWarps = 4;
constexpr int kPbMaxLen = 16;

struct ProbsBwdParams {
  const __nv_bfloat16* q; const __nv_bfloat16* k;
  const float* dp;
  __nv_bfloat16* dq; __nv_bfloat16* dk;
  int64_t q_rs, k_rs;                 // row strides (elements) of q / k; dq / dk are contiguous [rows, heads * 64]
  int32_t n_outer, n_inner, heads, len;
  float scale;
};

__global__ void __launch_bounds__(kPbWarps * 32) attn_short_probs_bwd_kernel(const ProbsBwdParams p) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float s_q[kPbWarps][kPbMaxLen][64];
  __shared__ float s_k[kPbWarps][kPbMaxLen][64];
  __shared__ float s_ds[kPbWarps][kPbMaxLen][kPbMaxLen + 1];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t n_items = int64_t(p.n_outer) * p.n_inner * p.heads;
  const int t = p.len;
  const int64_t o_rs = int64_t(p.heads) * 64;
  for (int64_t item = int64_t(blockIdx.x) * kPbWarps + warp; item < n_items; item += int64_t(gridDim.x) * kPbWarps) {
    const int head = int(item % p.heads);
    const int64_t seq = item / p.heads;
    const int64_t outer = seq / p.n_inner, inner = seq % p.n_inner;
    // token row of frame r: (outer * len + r) * n_inner + inner
    for (int r = 0; r < t; ++r) {
      const int64_t row = (outer * t + r) * p.n_inner + inner;
      const uint32_t uq = __ldg(reinterpret_cast<const uint32_t*>(p.q + row * p.q_rs + head * 64) + lane);
      const uint32_t uk = __ldg(reinterpret_cast<const uint32_t*>(p.k + row * p.k_rs + head * 64) + lane);
      s_q[warp][r][2 * lane] = bf16_lo(uq);
      s_q[warp][r][2 * lane + 1] = bf16_hi(uq);
      s_k[warp][r][2 * lane] = bf16_lo(uk);
      s_k[warp][r][2 * lane + 1] = bf16_hi(uk);
    }
    __syncwarp();
    if (lane < t) {
      float sc[kPbMaxLen];
      float mx = -INFINITY;
#pragma unroll
      for (int j = 0; j < kPbMaxLen; ++j) {
        float a = 0.f;
        if (j < t) {
          for (int d = 0; d < 64; ++d) a = fmaf(s_q[warp][lane][d], s_k[warp][j][d], a);
          a *= p.scale;
          mx = fmaxf(mx, a);
        }
        sc[j] = a;
      }
      float sum = 0.f;
#pragma unroll
      for (int j = 0; j < kPbMaxLen; ++j) {
        sc[j] = j < t ? __expf(sc[j] - mx) : 0.f;
        sum += sc[j];
      }
      const float inv = 1.0f / sum;
      const float* dprow = p.dp + (item * t + lane) * t;
      float dpv[kPbMaxLen];
      float dot = 0.f;
#pragma unroll
      for (int j = 0; j < kPbMaxLen; ++j) {
        sc[j] *= inv;
        dpv[j] = j < t ? __ldg(dprow + j) : 0.f;
        dot = fmaf(dpv[j], sc[j], dot);
      }
#pragma unroll
      for (int j = 0; j < kPbMaxLen; ++j)
        if (j < t) s_ds[warp][lane][j] = sc[j] * (dpv[j] - dot);
    }
    __syncwarp();
    if (lane < t) {
      const int64_t row = (outer * t + lane) * p.n_inner + inner;
      uint32_t* oq = reinterpret_cast<uint32_t*>(p.dq + row * o_rs + head * 64);
      uint32_t* ok = reinterpret_cast<uint32_t*>(p.dk + row * o_rs + head * 64);
      for (int d = 0; d < 64; d += 2) {
        float q0 = 0.f, q1 = 0.f, k0 = 0.f, k1 = 0.f;
        for (int j = 0; j < t; ++j) {
          const float a = s_ds[warp][lane][j], b = s_ds[warp][j][lane];
          q0 = fmaf(a, s_k[warp][j][d], q0);
          q1 = fmaf(a, s_k[warp][j][d + 1], q1);
          k0 = fmaf(b, s_q[warp][j][d], k0);
          k1 = fmaf(b, s_q[warp][j][d + 1], k1);
        }
        oq[d >> 1] = pack_bf16(q0 * p.scale, q1 * p.scale);
        ok[d >> 1] = pack_bf16(k0 * p.scale, k1 * p.scale);
      }
    }
    __syncwarp();
  }
}

}  // namespace

}  // namespace t2v

extern "C" int t2v_attn_short_probs_bwd(const void* q, int64_t q_row_stride, const void* k, int64_t k_row_stride, const float* d_probs,
                                        void* dq, void* dk, int32_t n_outer, int32_t n_inner, int32_t heads, int32_t len, float scale,
                                        t2v_stream_t s) {
  using namespace t2v;
  if (!q || !k || !d_probs || !dq || !dk) return fail(-1, "t2v_attn_short_probs_bwd: null pointer");
  if (n_outer < 1 || n_inner < 1 || heads < 1 || len < 1 || len > kPbMaxLen)
    return fail(-2, "t2v_attn_short_probs_bwd: need n_outer, n_inner, heads >= 1 and 1 <= len <= %d (got len %d)", kPbMaxLen, len);
  if (q_row_stride % 2 || k_row_stride % 2 || q_row_stride < int64_t(heads) * 64 || k_row_stride < int64_t(heads) * 64 ||
      ((reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(k) | reinterpret_cast<uintptr_t>(dq) | reinterpret_cast<uintptr_t>(dk)) & 3))
    return fail(-3, "t2v_attn_short_probs_bwd: rows must be 4-byte aligned with strides >= heads * 64");
  ProbsBwdParams p;
  p.q = static_cast<const __nv_bfloat16*>(q); p.k = static_cast<const __nv_bfloat16*>(k);
  p.dp = d_probs;
  p.dq = static_cast<__nv_bfloat16*>(dq); p.dk = static_cast<__nv_bfloat16*>(dk);
  p.q_rs = q_row_stride; p.k_rs = k_row_stride;
  p.n_outer = n_outer; p.n_inner = n_inner; p.heads = heads; p.len = len; p.scale = scale;
  const int64_t n_items = int64_t(n_outer) * n_inner * heads;
  const int sms = num_sms() > 0 ? num_sms() : 148;
  int64_t g = (n_items + kPbWarps - 1) / kPbWarps;
  if (g > int64_t(sms) * 8) g = int64_t(sms) * 8;
  launch_kernel(attn_short_probs_bwd_kernel, dim3(unsigned(g)), dim3(kPbWarps * 32), 0, static_cast<cudaStream_t>(s), p);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : cuda_fail(e, "t2v_attn_short_probs_bwd launch");
}

extern "C" int t2v_softmax_bwd_rows(void* dp, int64_t dp_row_stride, const void* p, int64_t p_row_stride, int64_t rows, int32_t cols,
                                    float scale, t2v_stream_t s) {
  using namespace t2v;
  if (!dp || !p || rows < 1 || cols < 1) return fail(-1, "t2v_softmax_bwd_rows: bad argument");
  if (rows > 0x7fffffff) return fail(-2, "t2v_softmax_bwd_rows: too many rows");
  if (dp_row_stride < cols || p_row_stride < cols) return fail(-3, "t2v_softmax_bwd_rows: row strides must be >= cols");
  launch_kernel(softmax_bwd_rows_kernel, dim3(unsigned(rows)), dim3(kThreads), 0, static_cast<cudaStream_t>(s),
                static_cast<__nv_bfloat16*>(dp), dp_row_stride, static_cast<const __nv_bfloat16*>(p), p_row_stride, cols, scale);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : cuda_fail(e, "t2v_softmax_bwd_rows launch");
}

extern "C" int t2v_groupnorm_affine_grad(const void* x, int64_t x_row_stride, const void* dy, int64_t dy_row_stride, const float* gamma,
                                         const float* beta, const float* stats_ws, float* dgamma, float* dbeta, int64_t rows,
                                         int64_t rows_per_sample, int32_t channels, int32_t groups, float eps, int32_t silu,
                                         t2v_stream_t s) {
  using namespace t2v;
  if (!x || !dy || !gamma || !beta || !stats_ws || !dgamma || !dbeta) return fail(-1, "t2v_groupnorm_affine_grad: null pointer");
  const int c = channels;
  if (c < 8 || c % 8 || c > kMaxGnChannels || groups < 1 || groups > 64 || c % groups)
    return fail(-2, "t2v_groupnorm_affine_grad: channels %% 8, channels <= %d, groups <= 64, channels %% groups", kMaxGnChannels);
  if (rows < 1 || rows_per_sample < 1 || rows % rows_per_sample) return fail(-3, "t2v_groupnorm_affine_grad: rows %% rows_per_sample != 0");
  if (x_row_stride % 8 || dy_row_stride % 8 || ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(dy)) & 15))
    return fail(-4, "t2v_groupnorm_affine_grad: x / dy must be 16-byte aligned with row strides multiple of 8");
  GnAffParams p;
  p.x = static_cast<const __nv_bfloat16*>(x); p.x_rs = x_row_stride;
  p.dy = static_cast<const __nv_bfloat16*>(dy); p.dy_rs = dy_row_stride;
  p.gamma = gamma; p.beta = beta; p.dgamma = dgamma; p.dbeta = dbeta;
  p.rows_per_sample = rows_per_sample;
  p.ncv = c / 8; p.cpg = c / groups; p.groups = groups; p.silu = silu; p.eps = eps;
  p.ws = stats_ws;
  const int64_t n_samples = rows / rows_per_sample;
  if (n_samples > 65535) return fail(-5, "t2v_groupnorm_affine_grad: too many samples");
  const int sms = num_sms() > 0 ? num_sms() : 148;
  const int tpr = p.ncv < kThreads ? p.ncv : kThreads;
  const int rpp = kThreads / tpr;
  // two CTAs per SM in total: every CTA ends with one global atomic per channel, so fewer, longer CTAs than the dx kernel
  int64_t want_blocks = (int64_t(sms) * 2 + n_samples - 1) / n_samples;
  int64_t rpb = (rows_per_sample + want_blocks - 1) / want_blocks;
  if (rpb < int64_t(rpp) * 4) rpb = int64_t(rpp) * 4;
  rpb = (rpb + rpp - 1) / rpp * rpp;
  p.rows_per_block = int(rpb);
  const int64_t bps = (rows_per_sample + rpb - 1) / rpb;
  launch_kernel(gn_affine_grad_kernel, dim3((unsigned)bps, (unsigned)n_samples), dim3(kThreads), 0, static_cast<cudaStream_t>(s), p);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : cuda_fail(e, "t2v_groupnorm_affine_grad launch");
}

extern "C" int t2v_layernorm_affine_grad(const void* x, int64_t x_row_stride, const void* dy, int64_t dy_row_stride, float* dgamma,
                                         float* dbeta, int64_t rows, int32_t channels, float eps, t2v_stream_t s) {
  using namespace t2v;
  if (!x || !dy || !dgamma || !dbeta || rows < 1) return fail(-1, "t2v_layernorm_affine_grad: bad argument");
  if (x_row_stride % 2 || dy_row_stride % 2 || ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(dy)) & 3))
    return fail(-2, "t2v_layernorm_affine_grad: rows must be 4-byte aligned (even strides)");
  const int sms = num_sms() > 0 ? num_sms() : 148;
  int64_t g = (rows + 7) / 8;
  if (g > int64_t(sms) * 2) g = int64_t(sms) * 2;
  cudaStream_t st = static_cast<cudaStream_t>(s);
  auto X = static_cast<const __nv_bfloat16*>(x);
  auto DY = static_cast<const __nv_bfloat16*>(dy);
#define T2V_LN_AFF(NI) launch_kernel(ln_affine_grad_kernel<NI>, dim3(unsigned(g)), dim3(kThreads), 0, st, X, x_row_stride, DY, dy_row_stride, \
                                     dgamma, dbeta, rows, eps)
  switch (channels) {
    case 64: T2V_LN_AFF(1); break;
    case 128: T2V_LN_AFF(2); break;
    case 256: T2V_LN_AFF(4); break;
    case 320: T2V_LN_AFF(5); break;
    case 512: T2V_LN_AFF(8); break;
    case 640: T2V_LN_AFF(10); break;
    case 1024: T2V_LN_AFF(16); break;
    case 1280: T2V_LN_AFF(20); break;
    default: return fail(-3, "t2v_layernorm_affine_grad: channels must be one of 64, 128, 256, 320, 512, 640, 1024, 1280 (got %d)", channels);
  }
#undef T2V_LN_AFF
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : cuda_fail(e, "t2v_layernorm_affine_grad launch");
}

extern "C" int t2v_ema_update(float* target, const float* src, int64_t n, float rate, t2v_stream_t s) {
  using namespace t2v;
  if (!target || !src || n < 1) return fail(-1, "t2v_ema_update: bad argument");
  if (rate < 0.f || rate > 1.f) return fail(-2, "t2v_ema_update: rate %f outside [0, 1]", double(rate));
  const bool vec = ((reinterpret_cast<uintptr_t>(target) | reinterpret_cast<uintptr_t>(src)) & 15) == 0;
  const int64_t n4 = vec ? n / 4 : 0;
  const int sms = num_sms() > 0 ? num_sms() : 148;
  int64_t g = ((vec ? n4 : n) + kThreads - 1) / kThreads;
  if (g > int64_t(sms) * 16) g = int64_t(sms) * 16;
  if (g < 1) g = 1;
  launch_kernel(ema_update_kernel, dim3(unsigned(g)), dim3(kThreads), 0, static_cast<cudaStream_t>(s), target, src, n4, n, rate);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : cuda_fail(e, "t2v_ema_update launch");
}
