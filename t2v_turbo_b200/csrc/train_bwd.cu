// Backward of the non-GEMM layers of the student UNet (train_t2v_turbo_v1_lora.py:1190 `accelerator.backward`): GroupNorm(+SiLU),
// LayerNorm, GEGLU, SiLU, the broadcast timestep-embedding add, nearest-2x upsampling / stride-2 subsampling adjoints and
// gradient accumulation.  All HBM bound: channels-last bf16 activations and gradients, fp32 arithmetic and statistics.
// Only the LoRA weights train (:862-906), so no kernel here produces gamma / beta / bias gradients.
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

__device__ __forceinline__ void unpack8f(const uint4& v, float (&f)[8]) {
  f[0] = bf16_lo(v.x); f[1] = bf16_hi(v.x);
  f[2] = bf16_lo(v.y); f[3] = bf16_hi(v.y);
  f[4] = bf16_lo(v.z); f[5] = bf16_hi(v.z);
  f[6] = bf16_lo(v.w); f[7] = bf16_hi(v.w);
}
__device__ __forceinline__ uint4 pack8f(const float (&f)[8]) {
  uint4 o;
  o.x = pack_bf16(f[0], f[1]);
  o.y = pack_bf16(f[2], f[3]);
  o.z = pack_bf16(f[4], f[5]);
  o.w = pack_bf16(f[6], f[7]);
  return o;
}
__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + __expf(-x)); }

unsigned grid_1d(int64_t n_items) {
  const int sms = num_sms();
  int64_t g = (n_items + kThreads - 1) / kThreads;
  const int64_t cap = int64_t(sms > 0 ? sms : 148) * 16;
  if (g > cap) g = cap;
  return unsigned(g < 1 ? 1 : g);
}

// ------------------------------------------------------------------------------------------------ GroupNorm backward
// y = act(xh * gamma + beta), xh = (x - mean_g) * rstd_g over (rows_per_sample x cpg) per (sample, group).
//   dpre = dy * act'(pre);  dyh = dpre * gamma;  S1 = sum dyh, S2 = sum dyh * xh  (per sample, group)
//   dx   = rstd * (dyh - (S1 + xh * S2) / N)  (+ dx_add)
// Three passes over a fixed 8-channel column per thread (the forward kernels' mapping, norm.cu): statistics of x,
// the two gradient sums, the apply.  workspace: fp32 [n_samples][groups][4] = (sum x, sum x^2, S1, S2), zero on entry.
struct GnBwdParams {
  const __nv_bfloat16* x; int64_t x_rs;
  const __nv_bfloat16* dy; int64_t dy_rs;
  const __nv_bfloat16* add; int64_t add_rs;
  __nv_bfloat16* dx; int64_t dx_rs;
  const float* gamma; const float* beta;
  int64_t rows_per_sample;
  int32_t rows_per_block, ncv, cpg, groups, silu;
  float eps;
  float* ws;
};

template <int kPass>
__global__ void __launch_bounds__(kThreads) gn_bwd_kernel(const GnBwdParams p) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float s_acc[2 * 64];
  __shared__ float s_st[4 * 64];
  const int sample = blockIdx.y;
  float* wsg = p.ws + int64_t(sample) * 4 * p.groups;
  if (kPass < 2) {
    for (int i = threadIdx.x; i < 2 * p.groups; i += blockDim.x) s_acc[i] = 0.f;
  }
  if (kPass > 0) {
    for (int i = threadIdx.x; i < 4 * p.groups; i += blockDim.x) s_st[i] = __ldcg(wsg + i);
  }
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
      float mean[8], rstd[8], ga[8], be[8], k1[8], k2[8];
      if (kPass > 0) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int c = cv * 8 + j;
          const int g = c / p.cpg;
          mean[j] = s_st[4 * g] * inv_n;
          float var = s_st[4 * g + 1] * inv_n - mean[j] * mean[j];
          var = var < 0.f ? 0.f : var;
          rstd[j] = rsqrtf(var + p.eps);
          ga[j] = __ldg(p.gamma + c);
          be[j] = __ldg(p.beta + c);
          k1[j] = s_st[4 * g + 2] * inv_n;
          k2[j] = s_st[4 * g + 3] * inv_n;
        }
      }
      float a0[8], a1[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) a0[j] = a1[j] = 0.f;
      for (int64_t r = row_begin + rr; r < row_end; r += rpp) {
        float xf[8];
        unpack8f(__ldg(reinterpret_cast<const uint4*>(p.x + (base_row + r) * p.x_rs + cv * 8)), xf);
        if (kPass == 0) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            a0[j] += xf[j];
            a1[j] += xf[j] * xf[j];
          }
        } else {
          float dyf[8];
          unpack8f(__ldg(reinterpret_cast<const uint4*>(p.dy + (base_row + r) * p.dy_rs + cv * 8)), dyf);
          float out[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float xh = (xf[j] - mean[j]) * rstd[j];
            float d = dyf[j];
            if (p.silu) {
              const float pre = fmaf(xh, ga[j], be[j]);
              const float sg = sigmoidf_(pre);
              d *= sg * fmaf(pre, 1.0f - sg, 1.0f);
            }
            const float dyh = d * ga[j];
            if (kPass == 1) {
              a0[j] += dyh;
              a1[j] = fmaf(dyh, xh, a1[j]);
            } else {
              out[j] = rstd[j] * (dyh - k1[j] - xh * k2[j]);
            }
          }
          if (kPass == 2) {
            if (p.add != nullptr) {
              float af[8];
              unpack8f(__ldg(reinterpret_cast<const uint4*>(p.add + (base_row + r) * p.add_rs + cv * 8)), af);
#pragma unroll
              for (int j = 0; j < 8; ++j) out[j] += af[j];
            }
            *reinterpret_cast<uint4*>(p.dx + (base_row + r) * p.dx_rs + cv * 8) = pack8f(out);
          }
        }
      }
      if (kPass < 2) {
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
          a += a0[j];
          b += a1[j];
        }
        atomicAdd(&s_acc[2 * g_prev], a);
        atomicAdd(&s_acc[2 * g_prev + 1], b);
      }
    }
  }
  if (kPass < 2) {
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * p.groups; i += blockDim.x)
      atomicAdd(&wsg[4 * (i >> 1) + 2 * kPass + (i & 1)], s_acc[i]);
  }
}

// ------------------------------------------------------------------------------------------------ LayerNorm backward
// One warp per row, the row in registers (C = 64 * kNI <= 1280): mean / rstd recomputed, then
//   g = dy * gamma;  dx = rstd * (g - mean(g) - xh * mean(g * xh))  (+ dx_add)
template <int kNI>
__global__ void __launch_bounds__(kThreads) ln_bwd_kernel(const __nv_bfloat16* __restrict__ x, int64_t x_rs,
                                                          const __nv_bfloat16* __restrict__ dy, int64_t dy_rs,
                                                          const __nv_bfloat16* __restrict__ add, int64_t add_rs,
                                                          __nv_bfloat16* __restrict__ dx, int64_t dx_rs,
                                                          const float* __restrict__ gamma, int64_t rows, float eps) {
  pdl_launch_dependents();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int64_t warps = int64_t(gridDim.x) * (kThreads / 32);
  constexpr float inv_c = 1.0f / float(64 * kNI);
  for (int64_t r = int64_t(blockIdx.x) * (kThreads / 32) + (threadIdx.x >> 5); r < rows; r += warps) {
    float xv[2 * kNI], gv[2 * kNI];
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
    float m1 = 0.f, m2 = 0.f;
#pragma unroll
    for (int i = 0; i < kNI; ++i) {
      const uint32_t u = __ldg(reinterpret_cast<const uint32_t*>(dy + r * dy_rs) + i * 32 + lane);
      const float2 gm = __ldg(reinterpret_cast<const float2*>(gamma) + i * 32 + lane);
      xv[2 * i] *= rstd;
      xv[2 * i + 1] *= rstd;
      gv[2 * i] = bf16_lo(u) * gm.x;
      gv[2 * i + 1] = bf16_hi(u) * gm.y;
      m1 += gv[2 * i] + gv[2 * i + 1];
      m2 = fmaf(gv[2 * i], xv[2 * i], fmaf(gv[2 * i + 1], xv[2 * i + 1], m2));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      m1 += __shfl_xor_sync(0xffffffffu, m1, o);
      m2 += __shfl_xor_sync(0xffffffffu, m2, o);
    }
    m1 *= inv_c;
    m2 *= inv_c;
#pragma unroll
    for (int i = 0; i < kNI; ++i) {
      float o0 = rstd * (gv[2 * i] - m1 - xv[2 * i] * m2);
      float o1 = rstd * (gv[2 * i + 1] - m1 - xv[2 * i + 1] * m2);
      if (add != nullptr) {
        const uint32_t u = __ldg(reinterpret_cast<const uint32_t*>(add + r * add_rs) + i * 32 + lane);
        o0 += bf16_lo(u);
        o1 += bf16_hi(u);
      }
      reinterpret_cast<uint32_t*>(dx + r * dx_rs)[i * 32 + lane] = pack_bf16(o0, o1);
    }
  }
}

// ------------------------------------------------------------------------------------------------ GEGLU forward / backward
// attention.py:516-523: a, gate = proj(x).chunk(2, -1); out = a * gelu(gate)   (erf GELU)
__device__ __forceinline__ float gelu_f(float g) { return 0.5f * g * (1.0f + erff(g * 0.70710678118654752f)); }
__device__ __forceinline__ float gelu_grad_f(float g) {
  return 0.5f * (1.0f + erff(g * 0.70710678118654752f)) + g * 0.3989422804014327f * __expf(-0.5f * g * g);
}

// mode 0: out[r, :I] = a * gelu(gate);  mode 1: dpre[r, :2I] from (pre, dout)
template <int kMode>
__global__ void __launch_bounds__(kThreads) geglu_kernel(const __nv_bfloat16* __restrict__ pre, int64_t pre_rs,
                                                         const __nv_bfloat16* __restrict__ dout, int64_t dout_rs,
                                                         __nv_bfloat16* __restrict__ out, int64_t out_rs, int64_t rows, int32_t inner8) {
  pdl_launch_dependents();
  pdl_wait();
  const int64_t total = rows * inner8;
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += int64_t(gridDim.x) * blockDim.x) {
    const int64_t r = i / inner8;
    const int c = int(i - r * inner8) * 8;
    float a[8], g[8];
    unpack8f(__ldg(reinterpret_cast<const uint4*>(pre + r * pre_rs + c)), a);
    unpack8f(__ldg(reinterpret_cast<const uint4*>(pre + r * pre_rs + inner8 * 8 + c)), g);
    if (kMode == 0) {
      float o[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = a[j] * gelu_f(g[j]);
      *reinterpret_cast<uint4*>(out + r * out_rs + c) = pack8f(o);
    } else {
      float d[8], da[8], dg[8];
      unpack8f(__ldg(reinterpret_cast<const uint4*>(dout + r * dout_rs + c)), d);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        da[j] = d[j] * gelu_f(g[j]);
        dg[j] = d[j] * a[j] * gelu_grad_f(g[j]);
      }
      *reinterpret_cast<uint4*>(out + r * out_rs + c) = pack8f(da);
      *reinterpret_cast<uint4*>(out + r * out_rs + inner8 * 8 + c) = pack8f(dg);
    }
  }
}

// ------------------------------------------------------------------------------------------------ elementwise 2-D (row-strided) ops
// op 0: out = a + b          op 1: out = silu(a)          op 2: out = b * silu'(a)   (a = pre-activation, b = upstream grad)
template <int kOp>
__global__ void __launch_bounds__(kThreads) ew2d_kernel(const __nv_bfloat16* __restrict__ a, int64_t a_rs,
                                                        const __nv_bfloat16* __restrict__ b, int64_t b_rs,
                                                        __nv_bfloat16* __restrict__ out, int64_t out_rs, int64_t rows, int32_t c8) {
  pdl_launch_dependents();
  pdl_wait();
  const int64_t total = rows * c8;
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += int64_t(gridDim.x) * blockDim.x) {
    const int64_t r = i / c8;
    const int c = int(i - r * c8) * 8;
    float x[8], y[8], o[8];
    unpack8f(__ldg(reinterpret_cast<const uint4*>(a + r * a_rs + c)), x);
    if (kOp != 1) unpack8f(__ldg(reinterpret_cast<const uint4*>(b + r * b_rs + c)), y);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (kOp == 0) {
        o[j] = x[j] + y[j];
      } else if (kOp == 1) {
        o[j] = x[j] * sigmoidf_(x[j]);
      } else {
        const float sg = sigmoidf_(x[j]);
        o[j] = y[j] * sg * fmaf(x[j], 1.0f - sg, 1.0f);
      }
    }
    *reinterpret_cast<uint4*>(out + r * out_rs + c) = pack8f(o);
  }
}

// ------------------------------------------------------------------------------------------------ per-sample column sums
// out[s, c] += sum over the rows of sample s of x[row, c]   (adjoint of the broadcast emb add, openaimodel3d.py:237-246)
__global__ void __launch_bounds__(kThreads) colsum_kernel(const __nv_bfloat16* __restrict__ x, int64_t x_rs, float* __restrict__ out,
                                                          int64_t rows_per_sample, int32_t rows_per_block, int32_t ncv) {
  pdl_launch_dependents();
  pdl_wait();
  const int sample = blockIdx.y;
  const int64_t row_begin = int64_t(blockIdx.x) * rows_per_block;
  int64_t row_end = row_begin + rows_per_block;
  if (row_end > rows_per_sample) row_end = rows_per_sample;
  const int tpr = ncv < kThreads ? ncv : kThreads;
  const int rpp = kThreads / tpr;
  const int rr = threadIdx.x / tpr;
  if (rr >= rpp) return;
  const int64_t base_row = int64_t(sample) * rows_per_sample;
  for (int cv = threadIdx.x % tpr; cv < ncv; cv += tpr) {
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    for (int64_t r = row_begin + rr; r < row_end; r += rpp) {
      float f[8];
      unpack8f(__ldg(reinterpret_cast<const uint4*>(x + (base_row + r) * x_rs + cv * 8)), f);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += f[j];
    }
    float* o = out + int64_t(sample) * ncv * 8 + cv * 8;
#pragma unroll
    for (int j = 0; j < 8; ++j) atomicAdd(o + j, acc[j]);
  }
}

// ------------------------------------------------------------------------------------------------ 2x resampling adjoints
// mode 0: out[n, y, x] = in[n, 2y, 2x]                         (subsample: a stride-2 conv = the stride-1 conv sampled)
// mode 1: out[n, 2y, 2x] = in[n, y, x], 0 elsewhere            (its adjoint: zero stuffing)
// mode 2: out[n, y, x] = sum_{dy,dx in {0,1}} in[n, 2y+dy, 2x+dx]  (adjoint of nearest-neighbour 2x upsampling)
template <int kMode>
__global__ void __launch_bounds__(kThreads) resample2x_kernel(const __nv_bfloat16* __restrict__ in, __nv_bfloat16* __restrict__ out,
                                                              int64_t n, int32_t ho, int32_t wo, int32_t c8) {
  pdl_launch_dependents();
  pdl_wait();
  // (ho, wo) = extents of the OUTPUT grid
  const int64_t total = n * ho * wo * c8;
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += int64_t(gridDim.x) * blockDim.x) {
    const int cv = int(i % c8);
    int64_t pix = i / c8;
    const int x = int(pix % wo);
    pix /= wo;
    const int y = int(pix % ho);
    const int64_t f = pix / ho;
    const int64_t c = int64_t(c8) * 8;
    uint4 o;
    if (kMode == 0) {
      o = __ldg(reinterpret_cast<const uint4*>(in + ((f * (2 * ho) + 2 * y) * (2 * wo) + 2 * x) * c + cv * 8));
    } else if (kMode == 1) {
      if ((x | y) & 1) o = make_uint4(0, 0, 0, 0);
      else o = __ldg(reinterpret_cast<const uint4*>(in + ((f * (ho / 2) + (y >> 1)) * (wo / 2) + (x >> 1)) * c + cv * 8));
    } else {
      float acc[8], t[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] = 0.f;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        unpack8f(__ldg(reinterpret_cast<const uint4*>(in + ((f * (2 * ho) + 2 * y + (q >> 1)) * (2 * wo) + 2 * x + (q & 1)) * c + cv * 8)), t);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] += t[j];
      }
      o = pack8f(acc);
    }
    *reinterpret_cast<uint4*>(out + i * 8) = o;
  }
}

// ------------------------------------------------------------------------------------------------ attention backward helper
// delta[b, h, i] = sum_d dO[b, i, h, d] * O[b, i, h, d]   (the softmax-Jacobian row term of the flash backward)
__global__ void __launch_bounds__(kThreads) attn_delta_kernel(const __nv_bfloat16* __restrict__ o, int64_t o_sb, int64_t o_st, int64_t o_sh,
                                                              const __nv_bfloat16* __restrict__ d_o, int64_t d_sb, int64_t d_st, int64_t d_sh,
                                                              float* __restrict__ delta, int32_t batch, int32_t heads, int32_t len) {
  pdl_launch_dependents();
  pdl_wait();
  const int64_t total = int64_t(batch) * heads * len;
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += int64_t(gridDim.x) * blockDim.x) {
    const int t = int(i % len);
    const int h = int((i / len) % heads);
    const int64_t b = i / (int64_t(len) * heads);
    const uint4* po = reinterpret_cast<const uint4*>(o + b * o_sb + int64_t(t) * o_st + int64_t(h) * o_sh);
    const uint4* pd = reinterpret_cast<const uint4*>(d_o + b * d_sb + int64_t(t) * d_st + int64_t(h) * d_sh);
    float acc = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      float a[8], g[8];
      unpack8f(__ldg(po + k), a);
      unpack8f(__ldg(pd + k), g);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc = fmaf(a[j], g[j], acc);
    }
    delta[i] = acc;
  }
}

bool misaligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) != 0; }

}  // namespace
}  // namespace t2v

extern "C" int t2v_groupnorm_bwd(const T2VGroupNormBwdDesc* d, t2v_stream_t s) {
  using namespace t2v;
  if (!d || !d->x || !d->dy || !d->dx || !d->gamma || !d->beta || !d->workspace) return fail(-1, "t2v_groupnorm_bwd: null pointer");
  const int c = d->channels;
  if (c < 8 || c % 8 || d->groups < 1 || d->groups > 64 || c % d->groups) return fail(-2, "t2v_groupnorm_bwd: channels %% 8, groups <= 64, channels %% groups");
  if (d->rows < 1 || d->rows_per_sample < 1 || d->rows % d->rows_per_sample) return fail(-3, "t2v_groupnorm_bwd: rows %% rows_per_sample != 0");
  if (d->x_row_stride % 8 || d->dy_row_stride % 8 || d->dx_row_stride % 8 || (d->dx_add && d->dx_add_row_stride % 8) || misaligned16(d->x) ||
      misaligned16(d->dy) || misaligned16(d->dx) || (d->dx_add && misaligned16(d->dx_add)))
    return fail(-4, "t2v_groupnorm_bwd: tensors must be 16-byte aligned with row strides multiple of 8");
  GnBwdParams p;
  p.x = static_cast<const __nv_bfloat16*>(d->x); p.x_rs = d->x_row_stride;
  p.dy = static_cast<const __nv_bfloat16*>(d->dy); p.dy_rs = d->dy_row_stride;
  p.add = static_cast<const __nv_bfloat16*>(d->dx_add); p.add_rs = d->dx_add_row_stride;
  p.dx = static_cast<__nv_bfloat16*>(d->dx); p.dx_rs = d->dx_row_stride;
  p.gamma = d->gamma; p.beta = d->beta;
  p.rows_per_sample = d->rows_per_sample;
  p.ncv = c / 8; p.cpg = c / d->groups; p.groups = d->groups; p.silu = d->silu; p.eps = d->eps;
  p.ws = d->workspace;
  const int64_t n_samples = d->rows / d->rows_per_sample;
  const int sms = num_sms() > 0 ? num_sms() : 148;
  const int tpr = p.ncv < kThreads ? p.ncv : kThreads;
  const int rpp = kThreads / tpr;
  int64_t want_blocks = (int64_t(sms) * 8 + n_samples - 1) / n_samples;
  int64_t rpb = (d->rows_per_sample + want_blocks - 1) / want_blocks;
  if (rpb < int64_t(rpp) * 4) rpb = int64_t(rpp) * 4;
  rpb = (rpb + rpp - 1) / rpp * rpp;
  p.rows_per_block = int(rpb);
  const int64_t bps = (d->rows_per_sample + rpb - 1) / rpb;
  if (n_samples > 65535) return fail(-5, "t2v_groupnorm_bwd: too many samples");
  cudaStream_t st = static_cast<cudaStream_t>(s);
  dim3 grid((unsigned)bps, (unsigned)n_samples);
  launch_kernel(gn_bwd_kernel<0>, grid, dim3(kThreads), 0, st, p);
  launch_kernel(gn_bwd_kernel<1>, grid, dim3(kThreads), 0, st, p);
  launch_kernel(gn_bwd_kernel<2>, grid, dim3(kThreads), 0, st, p);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : cuda_fail(e, "t2v_groupnorm_bwd launch");
}

extern "C" int t2v_layernorm_bwd(const void* x, int64_t x_row_stride, const void* dy, int64_t dy_row_stride, const void* dx_add,
                                 int64_t dx_add_row_stride, void* dx, int64_t dx_row_stride, const float* gamma, int64_t rows,
                                 int32_t channels, float eps, t2v_stream_t s) {
  using namespace t2v;
  if (!x || !dy || !dx || !gamma || rows < 1) return fail(-1, "t2v_layernorm_bwd: bad argument");
  if (x_row_stride % 2 || dy_row_stride % 2 || dx_row_stride % 2 || (dx_add && dx_add_row_stride % 2) ||
      ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(dx) | reinterpret_cast<uintptr_t>(dx_add)) & 3) ||
      (reinterpret_cast<uintptr_t>(gamma) & 7))
    return fail(-2, "t2v_layernorm_bwd: rows must be 4-byte aligned (even strides), gamma 8-byte aligned");
  const int sms = num_sms() > 0 ? num_sms() : 148;
  int64_t g = (rows + 7) / 8;
  if (g > int64_t(sms) * 16) g = int64_t(sms) * 16;
  cudaStream_t st = static_cast<cudaStream_t>(s);
  auto X = static_cast<const __nv_bfloat16*>(x);
  auto DY = static_cast<const __nv_bfloat16*>(dy);
  auto AD = static_cast<const __nv_bfloat16*>(dx_add);
  auto DX = static_cast<__nv_bfloat16*>(dx);
#define T2V_LN_BWD(NI) launch_kernel(ln_bwd_kernel<NI>, dim3(unsigned(g)), dim3(kThreads), 0, st, X, x_row_stride, DY, dy_row_stride, AD, \
                                     dx_add_row_stride, DX, dx_row_stride, gamma, rows, eps)
  switch (channels) {
    case 64: T2V_LN_BWD(1); break;
    case 128: T2V_LN_BWD(2); break;
    case 256: T2V_LN_BWD(4); break;
    case 320: T2V_LN_BWD(5); break;
    case 512: T2V_LN_BWD(8); break;
    case 640: T2V_LN_BWD(10); break;
    case 1024: T2V_LN_BWD(16); break;
    case 1280: T2V_LN_BWD(20); break;
    default: return fail(-3, "t2v_layernorm_bwd: channels must be one of 64, 128, 256, 320, 512, 640, 1024, 1280 (got %d)", channels);
  }
#undef T2V_LN_BWD
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : cuda_fail(e, "t2v_layernorm_bwd launch");
}

extern "C" int t2v_geglu(const void* pre, int64_t pre_row_stride, const void* dout, int64_t dout_row_stride, void* out,
                         int64_t out_row_stride, int64_t rows, int32_t inner, t2v_stream_t s) {
  using namespace t2v;
  if (!pre || !out || rows < 1 || inner < 8 || inner % 8) return fail(-1, "t2v_geglu: bad argument (inner %% 8 == 0)");
  if (pre_row_stride % 8 || out_row_stride % 8 || (dout && dout_row_stride % 8) || misaligned16(pre) || misaligned16(out) ||
      (dout && misaligned16(dout)))
    return fail(-2, "t2v_geglu: tensors must be 16-byte aligned with row strides multiple of 8");
  cudaStream_t st = static_cast<cudaStream_t>(s);
  const unsigned g = grid_1d(rows * (inner / 8));
  auto P = static_cast<const __nv_bfloat16*>(pre);
  auto D = static_cast<const __nv_bfloat16*>(dout);
  auto O = static_cast<__nv_bfloat16*>(out);
  if (dout == nullptr) launch_kernel(geglu_kernel<0>, dim3(g), dim3(kThreads), 0, st, P, pre_row_stride, D, dout_row_stride, O, out_row_stride, rows, inner / 8);
  else launch_kernel(geglu_kernel<1>, dim3(g), dim3(kThreads), 0, st, P, pre_row_stride, D, dout_row_stride, O, out_row_stride, rows, inner / 8);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : cuda_fail(e, "t2v_geglu launch");
}

extern "C" int t2v_ew2d(int32_t op, const void* a, int64_t a_row_stride, const void* b, int64_t b_row_stride, void* out,
                        int64_t out_row_stride, int64_t rows, int32_t channels, t2v_stream_t s) {
  using namespace t2v;
  if (!a || !out || rows < 1 || channels < 8 || channels % 8 || op < 0 || op > 2 || (op != 1 && !b)) return fail(-1, "t2v_ew2d: bad argument");
  if (a_row_stride % 8 || out_row_stride % 8 || (b && b_row_stride % 8) || misaligned16(a) || misaligned16(out) || (b && misaligned16(b)))
    return fail(-2, "t2v_ew2d: tensors must be 16-byte aligned with row strides multiple of 8");
  cudaStream_t st = static_cast<cudaStream_t>(s);
  const unsigned g = grid_1d(rows * (channels / 8));
  auto A = static_cast<const __nv_bfloat16*>(a);
  auto B = static_cast<const __nv_bfloat16*>(b);
  auto O = static_cast<__nv_bfloat16*>(out);
  if (op == 0) launch_kernel(ew2d_kernel<0>, dim3(g), dim3(kThreads), 0, st, A, a_row_stride, B, b_row_stride, O, out_row_stride, rows, channels / 8);
  else if (op == 1) launch_kernel(ew2d_kernel<1>, dim3(g), dim3(kThreads), 0, st, A, a_row_stride, B, b_row_stride, O, out_row_stride, rows, channels / 8);
  else launch_kernel(ew2d_kernel<2>, dim3(g), dim3(kThreads), 0, st, A, a_row_stride, B, b_row_stride, O, out_row_stride, rows, channels / 8);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : cuda_fail(e, "t2v_ew2d launch");
}

extern "C" int t2v_colsum_samples(const void* x, int64_t x_row_stride, float* out, int64_t rows, int64_t rows_per_sample, int32_t channels,
                                  t2v_stream_t s) {
  using namespace t2v;
  if (!x || !out || rows < 1 || rows_per_sample < 1 || rows % rows_per_sample || channels < 8 || channels % 8)
    return fail(-1, "t2v_colsum_samples: bad argument");
  if (x_row_stride % 8 || misaligned16(x)) return fail(-2, "t2v_colsum_samples: x must be 16-byte aligned with a row stride multiple of 8");
  const int64_t n_samples = rows / rows_per_sample;
  if (n_samples > 65535) return fail(-3, "t2v_colsum_samples: too many samples");
  const int sms = num_sms() > 0 ? num_sms() : 148;
  const int ncv = channels / 8;
  const int tpr = ncv < kThreads ? ncv : kThreads;
  const int rpp = kThreads / tpr;
  int64_t want_blocks = (int64_t(sms) * 4 + n_samples - 1) / n_samples;
  int64_t rpb = (rows_per_sample + want_blocks - 1) / want_blocks;
  if (rpb < int64_t(rpp) * 8) rpb = int64_t(rpp) * 8;
  rpb = (rpb + rpp - 1) / rpp * rpp;
  const int64_t bps = (rows_per_sample + rpb - 1) / rpb;
  launch_kernel(colsum_kernel, dim3((unsigned)bps, (unsigned)n_samples), dim3(kThreads), 0, static_cast<cudaStream_t>(s),
                static_cast<const __nv_bfloat16*>(x), x_row_stride, out, rows_per_sample, int32_t(rpb), ncv);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : cuda_fail(e, "t2v_colsum_samples launch");
}

extern "C" int t2v_resample2x(int32_t mode, const void* in, void* out, int64_t n, int32_t h_out, int32_t w_out, int32_t channels,
                              t2v_stream_t s) {
  using namespace t2v;
  if (!in || !out || n < 1 || h_out < 1 || w_out < 1 || channels < 8 || channels % 8 || mode < 0 || mode > 2) return fail(-1, "t2v_resample2x: bad argument");
  if (mode == 1 && ((h_out | w_out) & 1)) return fail(-2, "t2v_resample2x: zero stuffing needs even output extents");
  if (misaligned16(in) || misaligned16(out)) return fail(-3, "t2v_resample2x: tensors must be 16-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(s);
  const unsigned g = grid_1d(n * h_out * w_out * (channels / 8));
  auto I = static_cast<const __nv_bfloat16*>(in);
  auto O = static_cast<__nv_bfloat16*>(out);
  if (mode == 0) launch_kernel(resample2x_kernel<0>, dim3(g), dim3(kThreads), 0, st, I, O, n, h_out, w_out, channels / 8);
  else if (mode == 1) launch_kernel(resample2x_kernel<1>, dim3(g), dim3(kThreads), 0, st, I, O, n, h_out, w_out, channels / 8);
  else launch_kernel(resample2x_kernel<2>, dim3(g), dim3(kThreads), 0, st, I, O, n, h_out, w_out, channels / 8);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : cuda_fail(e, "t2v_resample2x launch");
}

extern "C" int t2v_attn_delta(const void* o, int64_t o_stride_b, int64_t o_stride_t, int64_t o_stride_h, const void* d_o,
                              int64_t do_stride_b, int64_t do_stride_t, int64_t do_stride_h, float* delta, int32_t batch, int32_t heads,
                              int32_t len, t2v_stream_t s) {
  using namespace t2v;
  if (!o || !d_o || !delta || batch < 1 || heads < 1 || len < 1) return fail(-1, "t2v_attn_delta: bad argument");
  if ((o_stride_b | o_stride_t | o_stride_h | do_stride_b | do_stride_t | do_stride_h) % 8 || misaligned16(o) || misaligned16(d_o))
    return fail(-2, "t2v_attn_delta: tensors must be 16-byte aligned with strides multiple of 8");
  launch_kernel(attn_delta_kernel, dim3(grid_1d(int64_t(batch) * heads * len)), dim3(kThreads), 0, static_cast<cudaStream_t>(s),
                static_cast<const __nv_bfloat16*>(o), o_stride_b, o_stride_t, o_stride_h, static_cast<const __nv_bfloat16*>(d_o), do_stride_b,
                do_stride_t, do_stride_h, delta, batch, heads, len);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : cuda_fail(e, "t2v_attn_delta launch");
}
