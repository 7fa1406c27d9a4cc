// HBM-bound kernels of the consistency-distillation step (train_t2v_turbo_v1_lora.py:943-1196) around the tensor-core
// GEMMs: the LoRA branch's dropout * scale (forward and adjoint), the MSE distillation loss with its gradient, the
// squared gradient norm for clipping and ONE fused AdamW launch over the flat fp32 LoRA arenas (575 layers, 117 M values).
#include "../../include/t2v_b200.h"
#ifdef T2V_HOST_EMU   // tests/cuda_emu: the SIMT kernels of this file compiled by g++ and run on CPU threads (test infrastructure only)
#include "cuda_emu.h"
#else
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "host_common.h"
#include "ptx.cuh"
#endif

namespace t2v {

static inline unsigned grid_1d(int64_t n, int per_thread = 1) {
  int sms = num_sms();
  int64_t g = (n + 256 * int64_t(per_thread) - 1) / (256 * int64_t(per_thread));
  const int64_t cap = int64_t(sms > 0 ? sms : 148) * 16;
  if (g > cap) g = cap;
  return unsigned(g < 1 ? 1 : g);
}

// Philox4x32-10 (Salmon et al., SC'11): counter (c0..c3), key (k0, k1) -> 128 random bits.
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += 0x9E3779B9u;
    k.y += 0xBB67AE85u;
  }
  return c;
}

// out = x * scale * keep (+ addend), keep ~ Bernoulli(keep_prob) drawn IN the kernel; the keep-mask (1 byte / element) is written for the
// backward.  8 bf16 per thread = one Philox call: element j of the group keeps iff its 16 random bits are < thresh
// (thresh = round(keep_prob * 65536): the keep rate is exact to 2^-17).  Stream = (*seed, call_id, element index / 8):
// reproducible, independent across calls (call_id) and steps (*seed is advanced on the device between steps, so a captured
// CUDA graph draws fresh masks on every replay).
__global__ void __launch_bounds__(256) dropout_scale_kernel(const uint4* __restrict__ x, const uint4* __restrict__ addend,
                                                            uint4* __restrict__ out, uint2* __restrict__ mask_out, int64_t n8,
                                                            uint32_t thresh, float scale, const uint64_t* __restrict__ seed,
                                                            uint32_t call_id) {
  pdl_launch_dependents();
  pdl_wait();
  const uint64_t sd = *seed;
  const uint2 key = make_uint2(uint32_t(sd), uint32_t(sd >> 32));
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n8; i += int64_t(gridDim.x) * blockDim.x) {
    const uint4 r = philox4x32_10(make_uint4(uint32_t(i), uint32_t(uint64_t(i) >> 32), call_id, 0x74327662u), key);
    const uint4 v = __ldg(x + i);
    const uint4 ad = addend != nullptr ? __ldg(addend + i) : make_uint4(0u, 0u, 0u, 0u);   // bf16 zeros
    const uint32_t rr[4] = {r.x, r.y, r.z, r.w};
    const uint32_t vv[4] = {v.x, v.y, v.z, v.w};
    const uint32_t aa[4] = {ad.x, ad.y, ad.z, ad.w};
    uint32_t oo[4], m[2] = {0u, 0u};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const bool k0 = (rr[j] & 0xffffu) < thresh, k1 = (rr[j] >> 16) < thresh;
      oo[j] = pack_bf16((k0 ? bf16_lo(vv[j]) * scale : 0.f) + bf16_lo(aa[j]), (k1 ? bf16_hi(vv[j]) * scale : 0.f) + bf16_hi(aa[j]));
      m[j >> 1] |= (uint32_t(k0) | (uint32_t(k1) << 8)) << (16 * (j & 1));
    }
    out[i] = make_uint4(oo[0], oo[1], oo[2], oo[3]);
    mask_out[i] = make_uint2(m[0], m[1]);
  }
}

// out = x * scale * mask, 8 bf16 per thread
__global__ void __launch_bounds__(256) scale_mask_kernel(const uint4* __restrict__ x, const uint8_t* __restrict__ mask,
                                                         uint4* __restrict__ out, int64_t n8, float scale) {
  pdl_launch_dependents();
  pdl_wait();
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n8; i += int64_t(gridDim.x) * blockDim.x) {
    const uint4 v = __ldg(x + i);
    float f[8] = {bf16_lo(v.x), bf16_hi(v.x), bf16_lo(v.y), bf16_hi(v.y), bf16_lo(v.z), bf16_hi(v.z), bf16_lo(v.w), bf16_hi(v.w)};
    if (mask != nullptr) {
      const uint2 m = __ldg(reinterpret_cast<const uint2*>(mask) + i);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const uint32_t w = j < 4 ? m.x : m.y;
        f[j] = ((w >> (8 * (j & 3))) & 0xffu) ? f[j] * scale : 0.f;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] *= scale;
    }
    uint4 o;
    o.x = pack_bf16(f[0], f[1]);
    o.y = pack_bf16(f[2], f[3]);
    o.z = pack_bf16(f[4], f[5]);
    o.w = pack_bf16(f[6], f[7]);
    out[i] = o;
  }
}

__global__ void __launch_bounds__(256) adamw_kernel(float4* __restrict__ p, const float4* __restrict__ g, float4* __restrict__ m,
                                                    float4* __restrict__ v, int64_t n4, float lr, float b1, float b2, float eps,
                                                    float wd, float bc1, float bc2_sqrt, float gscale) {
  pdl_launch_dependents();
  pdl_wait();
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += int64_t(gridDim.x) * blockDim.x) {
    float4 pp = p[i], mm = m[i], vv = v[i];
    const float4 gg = __ldg(g + i);
    float* pa = reinterpret_cast<float*>(&pp);
    float* ma = reinterpret_cast<float*>(&mm);
    float* va = reinterpret_cast<float*>(&vv);
    const float* ga = reinterpret_cast<const float*>(&gg);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float gr = ga[j] * gscale;
      float w = pa[j] * (1.f - lr * wd);                       // decoupled weight decay (torch.optim.AdamW)
      ma[j] = b1 * ma[j] + (1.f - b1) * gr;
      va[j] = b2 * va[j] + (1.f - b2) * gr * gr;
      const float denom = sqrtf(va[j]) / bc2_sqrt + eps;
      pa[j] = w - (lr / bc1) * (ma[j] / denom);
    }
    p[i] = pp;
    m[i] = mm;
    v[i] = vv;
  }
}

__global__ void __launch_bounds__(256) sum_squares_kernel(const float4* __restrict__ x, int64_t n4, const float* __restrict__ tail,
                                                          int n_tail, float* out) {
  pdl_launch_dependents();
  pdl_wait();
  float acc = 0.f;
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += int64_t(gridDim.x) * blockDim.x) {
    const float4 v = __ldg(x + i);
    acc = fmaf(v.x, v.x, fmaf(v.y, v.y, fmaf(v.z, v.z, fmaf(v.w, v.w, acc))));
  }
  if (blockIdx.x == 0 && int(threadIdx.x) < n_tail) acc = fmaf(tail[threadIdx.x], tail[threadIdx.x], acc);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  __shared__ float s[8];
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < 8; ++i) t += s[i];
    atomicAdd(out, t);
  }
}

__device__ __forceinline__ float ld_any(const void* p, int64_t i, int dt) {
  if (dt == 0) return __bfloat162float(static_cast<const __nv_bfloat16*>(p)[i]);
  if (dt == 1) return __half2float(static_cast<const __half*>(p)[i]);
  return static_cast<const float*>(p)[i];
}
__device__ __forceinline__ void st_any(void* p, int64_t i, int dt, float v) {
  if (dt == 0) static_cast<__nv_bfloat16*>(p)[i] = __float2bfloat16_rn(v);
  else if (dt == 1) static_cast<__half*>(p)[i] = __float2half_rn(v);
  else static_cast<float*>(p)[i] = v;
}

// huber_c < 0: mean squared error; else the pseudo-Huber loss of utils/common_utils.py:302-304, mean(sqrt(d^2 + c^2) - c)
__global__ void __launch_bounds__(256) mse_loss_grad_kernel(const void* a, const void* b, void* grad, float* loss, int64_t n, int dt,
                                                            float inv_n, float gscale, float huber_c) {
  pdl_launch_dependents();
  pdl_wait();
  float acc = 0.f;
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x) {
    const float d = ld_any(a, i, dt) - ld_any(b, i, dt);
    if (huber_c < 0.f) {
      acc = fmaf(d, d, acc);
      if (grad != nullptr) st_any(grad, i, dt, 2.f * d * inv_n * gscale);
    } else {
      const float r = sqrtf(fmaf(d, d, huber_c * huber_c));
      acc += r - huber_c;
      if (grad != nullptr) st_any(grad, i, dt, d / r * inv_n * gscale);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  __shared__ float s[8];
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < 8; ++i) t += s[i];
    atomicAdd(loss, t * inv_n);
  }
}

}  // namespace t2v

extern "C" int t2v_scale_mask(const void* x, const uint8_t* mask, void* out, int64_t n, float scale, t2v_stream_t s) {
  using namespace t2v;
  if (!x || !out || n < 1 || n % 8) return fail(-1, "t2v_scale_mask: n must be a positive multiple of 8");
  if ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out)) & 15 || (mask && (reinterpret_cast<uintptr_t>(mask) & 7)))
    return fail(-2, "t2v_scale_mask: pointers must be 16-byte (mask: 8-byte) aligned");
  launch_kernel(scale_mask_kernel, dim3(grid_1d(n / 8)), dim3(256), 0, static_cast<cudaStream_t>(s), static_cast<const uint4*>(x), mask,
                static_cast<uint4*>(out), n / 8, scale);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : cuda_fail(e, "t2v_scale_mask launch");
}

extern "C" int t2v_dropout_scale(const void* x, const void* addend, void* out, uint8_t* mask_out, int64_t n, float keep_prob,
                                 float scale, const uint64_t* seed, uint32_t call_id, t2v_stream_t s) {
  using namespace t2v;
  if (!x || !out || !mask_out || !seed || n < 1 || n % 8) return fail(-1, "t2v_dropout_scale: n must be a positive multiple of 8");
  if (!(keep_prob > 0.f) || keep_prob > 1.f) return fail(-1, "t2v_dropout_scale: keep_prob must be in (0, 1]");
  if ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(addend)) & 15 ||
      (reinterpret_cast<uintptr_t>(mask_out) & 7) || (reinterpret_cast<uintptr_t>(seed) & 7))
    return fail(-2, "t2v_dropout_scale: x / out must be 16-byte, mask_out / seed 8-byte aligned");
  const uint32_t thresh = uint32_t(lrintf(keep_prob * 65536.f));
  launch_kernel(dropout_scale_kernel, dim3(grid_1d(n / 8)), dim3(256), 0, static_cast<cudaStream_t>(s), static_cast<const uint4*>(x),
                static_cast<const uint4*>(addend), static_cast<uint4*>(out), reinterpret_cast<uint2*>(mask_out), n / 8, thresh, scale, seed,
                call_id);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : cuda_fail(e, "t2v_dropout_scale launch");
}

extern "C" int t2v_adamw_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, float lr, float beta1,
                              float beta2, float eps, float weight_decay, int32_t step, float grad_scale, t2v_stream_t s) {
  using namespace t2v;
  if (!param || !grad || !exp_avg || !exp_avg_sq || n < 4 || n % 4 || step < 1) return fail(-1, "t2v_adamw_step: bad argument (n %% 4 == 0, step >= 1)");
  if ((reinterpret_cast<uintptr_t>(param) | reinterpret_cast<uintptr_t>(grad) | reinterpret_cast<uintptr_t>(exp_avg) |
       reinterpret_cast<uintptr_t>(exp_avg_sq)) & 15)
    return fail(-2, "t2v_adamw_step: arenas must be 16-byte aligned");
  const float bc1 = 1.f - powf(beta1, float(step));
  const float bc2_sqrt = sqrtf(1.f - powf(beta2, float(step)));
  launch_kernel(adamw_kernel, dim3(grid_1d(n / 4)), dim3(256), 0, static_cast<cudaStream_t>(s), reinterpret_cast<float4*>(param),
                reinterpret_cast<const float4*>(grad), reinterpret_cast<float4*>(exp_avg), reinterpret_cast<float4*>(exp_avg_sq), n / 4,
                lr, beta1, beta2, eps, weight_decay, bc1, bc2_sqrt, grad_scale);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : cuda_fail(e, "t2v_adamw_step launch");
}

extern "C" int t2v_sum_squares(const float* x, int64_t n, float* out, t2v_stream_t s) {
  using namespace t2v;
  if (!x || !out || n < 1 || (reinterpret_cast<uintptr_t>(x) & 15)) return fail(-1, "t2v_sum_squares: bad argument");
  launch_kernel(sum_squares_kernel, dim3(grid_1d(n / 4 + 1)), dim3(256), 0, static_cast<cudaStream_t>(s), reinterpret_cast<const float4*>(x),
                n / 4, x + (n / 4) * 4, int(n % 4), out);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : cuda_fail(e, "t2v_sum_squares launch");
}

extern "C" int t2v_mse_loss_grad(const void* a, const void* b, void* grad, float* loss, int64_t n, int32_t dtype, float grad_scale,
                                 t2v_stream_t s) {
  using namespace t2v;
  if (!a || !b || !loss || n < 1 || dtype < 0 || dtype > 2) return fail(-1, "t2v_mse_loss_grad: bad argument");
  launch_kernel(mse_loss_grad_kernel, dim3(grid_1d(n)), dim3(256), 0, static_cast<cudaStream_t>(s), a, b, grad, loss, n, dtype,
                1.0f / float(n), grad_scale, -1.0f);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : cuda_fail(e, "t2v_mse_loss_grad launch");
}

extern "C" int t2v_huber_loss_grad(const void* a, const void* b, void* grad, float* loss, int64_t n, int32_t dtype, float huber_c,
                                   float grad_scale, t2v_stream_t s) {
  using namespace t2v;
  if (!a || !b || !loss || n < 1 || dtype < 0 || dtype > 2 || !(huber_c >= 0.f)) return fail(-1, "t2v_huber_loss_grad: bad argument");
  launch_kernel(mse_loss_grad_kernel, dim3(grid_1d(n)), dim3(256), 0, static_cast<cudaStream_t>(s), a, b, grad, loss, n, dtype,
                1.0f / float(n), grad_scale, huber_c);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : cuda_fail(e, "t2v_huber_loss_grad launch");
}
