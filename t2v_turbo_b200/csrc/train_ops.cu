// HBM-bound kernels of the consistency-distillation step (train_t2v_turbo_v1_lora.py:943-1196) around the tensor-core
// GEMMs: the LoRA branch's dropout * scale (forward and adjoint), the MSE distillation loss with its gradient, the
// squared gradient norm for clipping and ONE fused AdamW launch over the flat fp32 LoRA arenas (575 layers, 117 M values).
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "../../include/t2v_b200.h"
#include "host_common.h"
#include "ptx.cuh"

namespace t2v {

static inline unsigned grid_1d(int64_t n, int per_thread = 1) {
  int sms = num_sms();
  int64_t g = (n + 256 * int64_t(per_thread) - 1) / (256 * int64_t(per_thread));
  const int64_t cap = int64_t(sms > 0 ? sms : 148) * 16;
  if (g > cap) g = cap;
  return unsigned(g < 1 ? 1 : g);
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
