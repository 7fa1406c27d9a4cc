// Small / HBM-bound kernels around the tensor-core path: embedding MLP rows (M = batch), sinusoidal
// embeddings, the 4-channel latent convolution, layout conversion at the UNet boundary, nearest
// upsampling, channel concat, row softmax, the fused LCM scheduler step and weight packing.
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

__device__ __forceinline__ float silu_f(float x) { return x / (1.0f + __expf(-x)); }

__device__ __forceinline__ float load_as_float(const void* p, int64_t i, int dtype) {
  if (dtype == 0) return __bfloat162float(static_cast<const __nv_bfloat16*>(p)[i]);
  if (dtype == 1) return __half2float(static_cast<const __half*>(p)[i]);
  return static_cast<const float*>(p)[i];
}
__device__ __forceinline__ void store_from_float(void* p, int64_t i, int dtype, float v) {
  if (dtype == 0) static_cast<__nv_bfloat16*>(p)[i] = __float2bfloat16_rn(v);
  else if (dtype == 1) static_cast<__half*>(p)[i] = __float2half_rn(v);
  else static_cast<float*>(p)[i] = v;
}
__device__ __forceinline__ float round_to(float v, int dtype) {
  if (dtype == 0) return __bfloat162float(__float2bfloat16_rn(v));
  if (dtype == 1) return __half2float(__float2half_rn(v));
  return v;
}

// ------------------------------------------------------------------ small-M linear: warp per output column
constexpr int kSlMaxM = 8;
__global__ void __launch_bounds__(256) small_linear_kernel(const T2VSmallLinearDesc d) {
  pdl_launch_dependents();
  pdl_wait();
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= d.n) return;
  const __nv_bfloat16* w = static_cast<const __nv_bfloat16*>(d.w) + int64_t(warp) * d.k;
  for (int m0 = 0; m0 < d.m; m0 += kSlMaxM) {
    float acc[kSlMaxM];
#pragma unroll
    for (int i = 0; i < kSlMaxM; ++i) acc[i] = 0.f;
    for (int k = lane * 8; k < d.k; k += 256) {
      const uint4 wv = __ldg(reinterpret_cast<const uint4*>(w + k));
      float wf[8] = {bf16_lo(wv.x), bf16_hi(wv.x), bf16_lo(wv.y), bf16_hi(wv.y),
                     bf16_lo(wv.z), bf16_hi(wv.z), bf16_lo(wv.w), bf16_hi(wv.w)};
#pragma unroll
      for (int i = 0; i < kSlMaxM; ++i) {
        if (m0 + i < d.m) {
          const float* xr = d.x + int64_t(m0 + i) * d.x_row_stride + k;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float xv = __ldg(xr + j);
            if (d.silu_in) xv = silu_f(xv);
            acc[i] = fmaf(xv, wf[j], acc[i]);
          }
        }
      }
    }
#pragma unroll
    for (int i = 0; i < kSlMaxM; ++i) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], o);
    }
    if (lane == 0) {
#pragma unroll
      for (int i = 0; i < kSlMaxM; ++i) {
        if (m0 + i < d.m) {
          float v = acc[i];
          if (d.bias) v += d.bias[warp];
          if (d.round_bf16) v = round_to(v, 0);
          if (d.add) v += d.add[int64_t(m0 + i) * d.add_row_stride + warp];
          if (d.silu_out) v = silu_f(v);
          if (d.round_bf16) v = round_to(v, 0);
          d.out[int64_t(m0 + i) * d.out_row_stride + warp] = v;
        }
      }
    }
  }
}

// ------------------------------------------------------------------ sinusoidal embeddings
__global__ void sinusoidal_kernel(const float* t, const float* freqs, float* out, int m, int half,
                                  int sin_first, int round_bf16) {
  pdl_launch_dependents();
  pdl_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m * half) return;
  const int r = i / half, c = i % half;
  const float arg = t[r] * freqs[c];
  float s = sinf(arg), co = cosf(arg);
  if (round_bf16) {
    s = round_to(s, 0);
    co = round_to(co, 0);
  }
  float* o = out + int64_t(r) * 2 * half;
  if (sin_first) {
    o[c] = s;
    o[half + c] = co;
  } else {
    o[c] = co;
    o[half + c] = s;
  }
}

// ------------------------------------------------------------------ 3x3 conv, tiny C_in (latent -> features)
// one thread = one output pixel x 8 output channels; weights staged in smem as fp32 [tap*CIN][Cout]
// (lanes of a warp read consecutive output channels -> conflict-free 16-byte shared loads).
template <int CIN>
__global__ void __launch_bounds__(256)
conv3x3_small_kernel(const __nv_bfloat16* __restrict__ in, const __nv_bfloat16* __restrict__ w,
                     const float* __restrict__ bias, __nv_bfloat16* __restrict__ out, int n, int h,
                     int wd, int cout) {
  pdl_launch_dependents();
  pdl_wait();
  T2V_DYN_SMEM(float, s_w);  // [9*CIN][cout]
  for (int i = threadIdx.x; i < cout * 9 * CIN; i += blockDim.x) {
    const int oc = i / (9 * CIN), q = i % (9 * CIN);  // global layout [cout][9*CIN]
    s_w[q * cout + oc] = __bfloat162float(w[i]);
  }
  __syncthreads();
  const int ocv = cout / 8;
  const int64_t total = int64_t(n) * h * wd * ocv;
  for (int64_t idx = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; idx < total;
       idx += int64_t(gridDim.x) * blockDim.x) {
    const int cv = int(idx % ocv);
    const int64_t pix = idx / ocv;
    const int x = int(pix % wd);
    const int y = int((pix / wd) % h);
    const int64_t img = pix / (int64_t(wd) * h);
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = bias ? bias[cv * 8 + j] : 0.f;
#pragma unroll
    for (int ky = 0; ky < 3; ++ky)
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const int yy = y + ky - 1, xx = x + kx - 1;
        if (yy < 0 || yy >= h || xx < 0 || xx >= wd) continue;
        const __nv_bfloat16* ip = in + ((img * h + yy) * wd + xx) * CIN;
        float pv[CIN];
        if (CIN == 4) {
          const uint2 u = __ldg(reinterpret_cast<const uint2*>(ip));
          pv[0] = bf16_lo(u.x); pv[1] = bf16_hi(u.x); pv[2] = bf16_lo(u.y); pv[3] = bf16_hi(u.y);
        } else {
#pragma unroll
          for (int c = 0; c < CIN; ++c) pv[c] = __bfloat162float(ip[c]);
        }
#pragma unroll
        for (int c = 0; c < CIN; ++c) {
          const float4* wr = reinterpret_cast<const float4*>(s_w + ((ky * 3 + kx) * CIN + c) * cout + cv * 8);
          const float4 w0 = wr[0], w1 = wr[1];
          acc[0] = fmaf(pv[c], w0.x, acc[0]); acc[1] = fmaf(pv[c], w0.y, acc[1]);
          acc[2] = fmaf(pv[c], w0.z, acc[2]); acc[3] = fmaf(pv[c], w0.w, acc[3]);
          acc[4] = fmaf(pv[c], w1.x, acc[4]); acc[5] = fmaf(pv[c], w1.y, acc[5]);
          acc[6] = fmaf(pv[c], w1.z, acc[6]); acc[7] = fmaf(pv[c], w1.w, acc[7]);
        }
      }
    uint4 o;
    o.x = pack_bf16(acc[0], acc[1]);
    o.y = pack_bf16(acc[2], acc[3]);
    o.z = pack_bf16(acc[4], acc[5]);
    o.w = pack_bf16(acc[6], acc[7]);
    *reinterpret_cast<uint4*>(out + pix * cout + cv * 8) = o;
  }
}

// 4-channel fast path (the latent convs: UNet input_blocks.0.0, VAE conv_in / encoder conv_in): one thread = 8 output
// channels x 4 consecutive pixels of a row.  The 3 x 6 input patch lives in registers and every weight fetched from shared
// memory is applied to 4 pixels: 16 FMAs per shared-memory load instead of 4 (the one-pixel kernel above is bound by
// shared-memory bandwidth: 1.36 ms for a 335 MB output that streams in ~60 us).
__global__ void __launch_bounds__(256)
conv3x3_cin4_x4_kernel(const __nv_bfloat16* __restrict__ in, const __nv_bfloat16* __restrict__ w, const float* __restrict__ bias,
                       __nv_bfloat16* __restrict__ out, int n, int h, int wd, int cout) {
  pdl_launch_dependents();
  pdl_wait();
  T2V_DYN_SMEM(float, s_w);  // [36][cout]
  for (int i = threadIdx.x; i < cout * 36; i += blockDim.x) {
    const int oc = i / 36, q = i % 36;  // global layout [cout][9 taps][4]
    s_w[q * cout + oc] = __bfloat162float(w[i]);
  }
  __syncthreads();
  const int ocv = cout / 8;
  const int wq = wd / 4;
  const int64_t total = int64_t(n) * h * wq * ocv;
  for (int64_t idx = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; idx < total; idx += int64_t(gridDim.x) * blockDim.x) {
    const int cv = int(idx % ocv);
    const int64_t pg = idx / ocv;
    const int x0 = int(pg % wq) * 4;
    const int y = int((pg / wq) % h);
    const int64_t img = pg / (int64_t(wq) * h);
    float pv[3][6][4];
#pragma unroll
    for (int dy = 0; dy < 3; ++dy) {
      const int yy = y + dy - 1;
#pragma unroll
      for (int dx = 0; dx < 6; ++dx) {
        const int xx = x0 + dx - 1;
        uint2 u = make_uint2(0u, 0u);
        if (yy >= 0 && yy < h && xx >= 0 && xx < wd) u = __ldg(reinterpret_cast<const uint2*>(in + ((img * h + yy) * wd + xx) * 4));
        pv[dy][dx][0] = bf16_lo(u.x); pv[dy][dx][1] = bf16_hi(u.x); pv[dy][dx][2] = bf16_lo(u.y); pv[dy][dx][3] = bf16_hi(u.y);
      }
    }
    float acc[4][8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float b = bias ? bias[cv * 8 + j] : 0.f;
#pragma unroll
      for (int px = 0; px < 4; ++px) acc[px][j] = b;
    }
#pragma unroll
    for (int ky = 0; ky < 3; ++ky)
#pragma unroll
      for (int kx = 0; kx < 3; ++kx)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const float4* wr = reinterpret_cast<const float4*>(s_w + ((ky * 3 + kx) * 4 + c) * cout + cv * 8);
          const float4 w0 = wr[0], w1 = wr[1];
          const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
          for (int px = 0; px < 4; ++px) {
            const float v = pv[ky][px + kx][c];
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[px][j] = fmaf(v, wv[j], acc[px][j]);
          }
        }
    __nv_bfloat16* op = out + ((img * h + y) * int64_t(wd) + x0) * cout + cv * 8;
#pragma unroll
    for (int px = 0; px < 4; ++px) {
      uint4 o;
      o.x = pack_bf16(acc[px][0], acc[px][1]);
      o.y = pack_bf16(acc[px][2], acc[px][3]);
      o.z = pack_bf16(acc[px][4], acc[px][5]);
      o.w = pack_bf16(acc[px][6], acc[px][7]);
      *reinterpret_cast<uint4*>(op + int64_t(px) * cout) = o;
    }
  }
}

// ------------------------------------------------------------------ layout conversion
__global__ void bcthw_to_frames_kernel(const void* in, int in_dtype, __nv_bfloat16* out, int b, int c,
                                       int t, int h, int w, float scale) {
  pdl_launch_dependents();
  pdl_wait();
  const int64_t total = int64_t(b) * t * h * w * c;
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += int64_t(gridDim.x) * blockDim.x) {
    const int ci = int(i % c);
    int64_t r = i / c;
    const int x = int(r % w); r /= w;
    const int y = int(r % h); r /= h;
    const int ti = int(r % t);
    const int bi = int(r / t);
    const int64_t src = (((int64_t(bi) * c + ci) * t + ti) * h + y) * w + x;
    out[i] = __float2bfloat16_rn(load_as_float(in, src, in_dtype) * scale);
  }
}
__global__ void bcthw_to_frames_mix_kernel(const void* in, int in_dtype, __nv_bfloat16* out, int b, int c,
                                           int t, int h, int w, float scale, const float* mix,
                                           const float* bias) {
  pdl_launch_dependents();
  pdl_wait();
  const int64_t total = int64_t(b) * t * h * w;
  const int64_t plane = int64_t(t) * h * w;
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += int64_t(gridDim.x) * blockDim.x) {
    const int64_t bi = i / plane, rem = i % plane;  // rem = (t, y, x) flattened, same order in both layouts
    float v[8];
    for (int ci = 0; ci < c; ++ci) v[ci] = load_as_float(in, (bi * c + ci) * plane + rem, in_dtype) * scale;
    for (int o = 0; o < c; ++o) {
      float a = bias ? bias[o] : 0.f;
      for (int ci = 0; ci < c; ++ci) a = fmaf(mix[o * c + ci], v[ci], a);
      out[i * c + o] = __float2bfloat16_rn(a);
    }
  }
}
// [B,C,T,H,W] -> [B*T,H,W,c_pad] bf16 with channels c..c_pad-1 zero (RGB -> 4-channel frames for the small-Cin conv)
__global__ void bcthw_to_frames_pad_kernel(const void* in, int in_dtype, __nv_bfloat16* out, int b, int c, int c_pad,
                                           int t, int h, int w, float scale) {
  pdl_launch_dependents();
  pdl_wait();
  const int64_t total = int64_t(b) * t * h * w;
  const int64_t plane = int64_t(t) * h * w;
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += int64_t(gridDim.x) * blockDim.x) {
    const int64_t bi = i / plane, rem = i % plane;
    for (int o = 0; o < c_pad; ++o)
      out[i * c_pad + o] = __float2bfloat16_rn(o < c ? load_as_float(in, (bi * c + o) * plane + rem, in_dtype) * scale : 0.f);
  }
}
__global__ void frames_to_bcthw_kernel(const __nv_bfloat16* in, int c_pad, void* out, int out_dtype,
                                       int b, int c, int t, int h, int w) {
  pdl_launch_dependents();
  pdl_wait();
  const int64_t total = int64_t(b) * c * t * h * w;
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += int64_t(gridDim.x) * blockDim.x) {
    int64_t r = i;
    const int x = int(r % w); r /= w;
    const int y = int(r % h); r /= h;
    const int ti = int(r % t); r /= t;
    const int ci = int(r % c);
    const int bi = int(r / c);
    const int64_t src = (((int64_t(bi) * t + ti) * h + y) * w + x) * c_pad + ci;
    store_from_float(out, i, out_dtype, __bfloat162float(in[src]));
  }
}
__global__ void upsample2x_kernel(const uint4* in, uint4* out, int n, int h, int w, int cv) {
  pdl_launch_dependents();
  pdl_wait();
  const int64_t total = int64_t(n) * 2 * h * 2 * w * cv;
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += int64_t(gridDim.x) * blockDim.x) {
    const int c = int(i % cv);
    int64_t r = i / cv;
    const int x = int(r % (2 * w)); r /= (2 * w);
    const int y = int(r % (2 * h));
    const int64_t img = r / (2 * h);
    out[i] = __ldg(in + ((img * h + (y >> 1)) * w + (x >> 1)) * cv + c);
  }
}
__global__ void concat_kernel(const uint4* a, int cva, const uint4* b, int cvb, uint4* out, int64_t rows) {
  pdl_launch_dependents();
  pdl_wait();
  const int cv = cva + cvb;
  const int64_t total = rows * cv;
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += int64_t(gridDim.x) * blockDim.x) {
    const int c = int(i % cv);
    const int64_t r = i / cv;
    out[i] = c < cva ? __ldg(a + r * cva + c) : __ldg(b + r * cvb + (c - cva));
  }
}

// ------------------------------------------------------------------ row softmax (bf16 in place), block per row
__global__ void __launch_bounds__(256) softmax_rows_kernel(__nv_bfloat16* x, int cols, int64_t row_stride, float scale) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float red[32];
  __nv_bfloat16* row = x + int64_t(blockIdx.x) * row_stride;
  float mx = -INFINITY;
  for (int c = threadIdx.x; c < cols; c += blockDim.x) mx = fmaxf(mx, __bfloat162float(row[c]) * scale);
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
  __syncthreads();
  mx = red[0];
  for (int i = 1; i < (blockDim.x >> 5); ++i) mx = fmaxf(mx, red[i]);
  __syncthreads();
  float sum = 0.f;
  for (int c = threadIdx.x; c < cols; c += blockDim.x) sum += __expf(__bfloat162float(row[c]) * scale - mx);
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = sum;
  __syncthreads();
  sum = 0.f;
  for (int i = 0; i < (blockDim.x >> 5); ++i) sum += red[i];
  const float inv = 1.0f / sum;
  for (int c = threadIdx.x; c < cols; c += blockDim.x)
    row[c] = __float2bfloat16_rn(__expf(__bfloat162float(row[c]) * scale - mx) * inv);
}

// ------------------------------------------------------------------ posterior sample of the KL-VAE encoder
// moments: fp32 channels-last [B*T, H, W, 2*zc] (mean | logvar); noise: fp32 [B*T, zc, H, W] (the reference's layout,
// may be null = posterior mode); out: [B, zc, T, H, W] in out_dtype = scale * (mean + exp(0.5 clamp(logvar,-30,20)) * noise)
__global__ void gaussian_sample_kernel(const float* __restrict__ mom, const float* __restrict__ noise, void* out,
                                       int out_dtype, int b, int t, int h, int w, int zc, float scale) {
  pdl_launch_dependents();
  pdl_wait();
  const int64_t plane = int64_t(h) * w;
  const int64_t total = int64_t(b) * t * plane * zc;
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += int64_t(gridDim.x) * blockDim.x) {
    // i indexes the output [b][zc][t][h*w]
    const int64_t px = i % plane;
    int64_t r = i / plane;
    const int ti = int(r % t);
    r /= t;
    const int c = int(r % zc);
    const int bi = int(r / zc);
    const int64_t f = int64_t(bi) * t + ti;
    const float* m = mom + (f * plane + px) * (2 * zc);
    float z = m[c];
    if (noise != nullptr) {
      const float lv = fminf(fmaxf(m[zc + c], -30.f), 20.f);
      z = fmaf(expf(0.5f * lv), noise[(f * zc + c) * plane + px], z);
    }
    store_from_float(out, i, out_dtype, scale * z);
  }
}

// ------------------------------------------------------------------ text-tower token + positional embedding, video -> uint8
__global__ void embedding_gather_kernel(const void* table, const void* pos, int dt, const int64_t* __restrict__ ids,
                                        __nv_bfloat16* out, int64_t n, int width, int ctx, int vocab) {
  pdl_launch_dependents();
  pdl_wait();
  const int64_t total = n * width;
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += int64_t(gridDim.x) * blockDim.x) {
    const int64_t row = i / width;
    const int c = int(i - row * width);
    int64_t id = ids[row];
    id = id < 0 ? 0 : (id >= vocab ? vocab - 1 : id);
    float v = load_as_float(table, id * width + c, dt);
    if (pos != nullptr) v = round_to(v + load_as_float(pos, (row % ctx) * int64_t(width) + c, dt), dt);   // x + positional_embedding in the model dtype
    out[i] = __float2bfloat16_rn(v);
  }
}

__global__ void video_to_uint8_kernel(const void* video, int dt, uint8_t* out, int b, int t, int h, int w) {
  pdl_launch_dependents();
  pdl_wait();
  const int64_t plane = int64_t(h) * w;
  const int64_t total = int64_t(b) * t * plane * 3;
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += int64_t(gridDim.x) * blockDim.x) {
    // i indexes the output [b][t][h*w][3]
    const int c = int(i % 3);
    int64_t r = i / 3;
    const int64_t px = r % plane;
    r /= plane;
    const int ti = int(r % t);
    const int bi = int(r / t);
    float v = load_as_float(video, ((int64_t(bi) * 3 + c) * t + ti) * plane + px, dt);
    v = fminf(fmaxf(v, -1.0f), 1.0f);
    v = (v + 1.0f) / 2.0f;
    out[i] = static_cast<uint8_t>(v * 255.0f);   // truncation, like torch's .to(torch.uint8)
  }
}

// ------------------------------------------------------------------ per-row scaled sum (add_noise, DDIM solver lines)
__global__ void scale_add_rows_kernel(const void* x, const void* y, const float* __restrict__ a, const float* __restrict__ b,
                                      void* out, int64_t rows, int64_t row_len, int dt) {
  pdl_launch_dependents();
  pdl_wait();
  const int64_t n = rows * row_len;
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x) {
    const int64_t r = i / row_len;
    float v = round_to(__ldg(a + r) * load_as_float(x, i, dt), dt);
    if (y != nullptr) v = round_to(v + round_to(__ldg(b + r) * load_as_float(y, i, dt), dt), dt);
    store_from_float(out, i, dt, v);
  }
}

// ------------------------------------------------------------------ fused LCM step
__global__ void lcm_step_kernel(const void* x, const void* eps, const void* noise, void* prev, void* den,
                                int64_t n, int dt, float sa_inv, float sb, float c_skip, float c_out,
                                float sap, float sbp) {
  pdl_launch_dependents();
  pdl_wait();
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += int64_t(gridDim.x) * blockDim.x) {
    const float xv = load_as_float(x, i, dt);
    const float ev = load_as_float(eps, i, dt);
    // every tensor op of the reference rounds to the tensor dtype
    const float t1 = round_to(sb * ev, dt);
    const float t2 = round_to(xv - t1, dt);
    const float x0 = round_to(t2 * sa_inv, dt);
    const float t3 = round_to(c_out * x0, dt);
    const float t4 = round_to(c_skip * xv, dt);
    const float dv = round_to(t3 + t4, dt);
    store_from_float(den, i, dt, dv);
    if (noise != nullptr) {
      const float nv = load_as_float(noise, i, dt);
      const float t5 = round_to(sap * dv, dt);
      const float t6 = round_to(sbp * nv, dt);
      store_from_float(prev, i, dt, t5 + t6);
    } else {
      store_from_float(prev, i, dt, dv);
    }
  }
}

// ------------------------------------------------------------------ weight packing
__global__ void pack_conv_weight_kernel(const void* w, int wdt, __nv_bfloat16* out, int cout, int cin, int taps) {
  pdl_launch_dependents();
  pdl_wait();
  const int64_t total = int64_t(cout) * cin * taps;
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += int64_t(gridDim.x) * blockDim.x) {
    const int c = int(i % cin);
    const int tp = int((i / cin) % taps);
    const int64_t o = i / (int64_t(cin) * taps);
    out[i] = __float2bfloat16_rn(load_as_float(w, (o * cin + c) * taps + tp, wdt));
  }
}
// packed row index for original row r (r < inner: value row, else gate row), blocks of 16
__device__ __forceinline__ int64_t geglu_packed_row(int64_t r, int inner) {
  const bool gate = r >= inner;
  const int64_t j = gate ? r - inner : r;
  return (j / 16) * 32 + (gate ? 16 : 0) + (j % 16);
}
__global__ void pack_geglu_kernel(const void* w, int wdt, __nv_bfloat16* out, const void* bias, int bdt,
                                  float* bias_out, int inner, int k) {
  pdl_launch_dependents();
  pdl_wait();
  const int64_t total = int64_t(2) * inner * k;
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += int64_t(gridDim.x) * blockDim.x) {
    const int64_t r = i / k;
    const int c = int(i % k);
    const int64_t pr = geglu_packed_row(r, inner);
    out[pr * k + c] = __float2bfloat16_rn(load_as_float(w, i, wdt));
    if (c == 0 && bias != nullptr) bias_out[pr] = load_as_float(bias, r, bdt);
  }
}

static inline unsigned grid_for(int64_t total, int threads = 256) {
  int64_t g = (total + threads - 1) / threads;
  const int64_t cap = 148 * 16;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return unsigned(g);
}

// one resident wave: every block stages the whole weight tensor in shared memory exactly once
static inline unsigned small_conv_grid(int64_t total) {
  int64_t g = (total + 255) / 256;
  const int64_t cap = 2 * (num_sms() > 0 ? num_sms() : 148);
  if (g > cap) g = cap;
  return unsigned(g < 1 ? 1 : g);
}

}  // namespace t2v

using namespace t2v;

extern "C" int t2v_small_linear(const T2VSmallLinearDesc* d, t2v_stream_t s) {
  if (!d || !d->x || !d->w || !d->out) return fail(-1, "t2v_small_linear: null pointer");
  if (d->k % 8 || d->k <= 0 || d->n <= 0 || d->m <= 0) return fail(-2, "t2v_small_linear: k must be a positive multiple of 8");
  const int warps_per_block = 8;
  const unsigned grid = unsigned((d->n + warps_per_block - 1) / warps_per_block);
  launch_kernel(small_linear_kernel, dim3(grid), dim3(256), 0, static_cast<cudaStream_t>(s), *d);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : cuda_fail(e, "t2v_small_linear launch");
}

extern "C" int t2v_sinusoidal_embedding(const float* t, const float* freqs, float* out, int32_t m,
                                        int32_t half, int32_t sin_first, int32_t round_bf16,
                                        t2v_stream_t s) {
  if (!t || !freqs || !out || m <= 0 || half <= 0) return fail(-1, "t2v_sinusoidal_embedding: bad argument");
  launch_kernel(sinusoidal_kernel, dim3(grid_for(int64_t(m) * half)), dim3(256), 0, static_cast<cudaStream_t>(s), t, freqs, out, m, half, sin_first, round_bf16);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : cuda_fail(e, "t2v_sinusoidal_embedding launch");
}

extern "C" int t2v_conv3x3_small_cin(const void* in, const void* w, const float* bias, void* out,
                                     int32_t n, int32_t h, int32_t wd, int32_t cin, int32_t cout,
                                     t2v_stream_t s) {
  if (!in || !w || !out) return fail(-1, "t2v_conv3x3_small_cin: null pointer");
  if (cout % 8 || cout <= 0) return fail(-2, "t2v_conv3x3_small_cin: cout must be a multiple of 8");
  const size_t smem = size_t(cout) * 9 * cin * sizeof(float);
  if (smem > 96 * 1024) return fail(-3, "t2v_conv3x3_small_cin: weights do not fit shared memory");
  const int64_t total = int64_t(n) * h * wd * (cout / 8);
  cudaStream_t st = static_cast<cudaStream_t>(s);
  const __nv_bfloat16* ip = static_cast<const __nv_bfloat16*>(in);
  const __nv_bfloat16* wp = static_cast<const __nv_bfloat16*>(w);
  __nv_bfloat16* op = static_cast<__nv_bfloat16*>(out);
  cudaError_t e = cudaSuccess;
  if (cin == 4 && wd % 4 == 0) {
    static bool cfg = false;
    if (!cfg) { e = cudaFuncSetAttribute(conv3x3_cin4_x4_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024); cfg = true; }
    if (e == cudaSuccess) launch_kernel(conv3x3_cin4_x4_kernel, dim3(small_conv_grid(total / 4)), dim3(256), smem, st, ip, wp, bias, op, n, h, wd, cout);
  } else if (cin == 4) {
    static bool cfg = false;
    if (!cfg) { e = cudaFuncSetAttribute(conv3x3_small_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024); cfg = true; }
    if (e == cudaSuccess) launch_kernel(conv3x3_small_kernel<4>, dim3(small_conv_grid(total)), dim3(256), smem, st, ip, wp, bias, op, n, h, wd, cout);
  } else if (cin == 8) {
    static bool cfg = false;
    if (!cfg) { e = cudaFuncSetAttribute(conv3x3_small_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024); cfg = true; }
    if (e == cudaSuccess) launch_kernel(conv3x3_small_kernel<8>, dim3(small_conv_grid(total)), dim3(256), smem, st, ip, wp, bias, op, n, h, wd, cout);
  } else {
    return fail(-4, "t2v_conv3x3_small_cin: cin must be 4 or 8 (got %d)", cin);
  }
  if (e == cudaSuccess) e = cudaGetLastError();
  return e == cudaSuccess ? 0 : cuda_fail(e, "t2v_conv3x3_small_cin launch");
}

extern "C" int t2v_bcthw_to_frames(const void* in, int32_t in_dtype, void* out, int32_t b, int32_t c,
                                   int32_t t, int32_t h, int32_t w, float scale, t2v_stream_t s) {
  if (!in || !out) return fail(-1, "t2v_bcthw_to_frames: null pointer");
  const int64_t total = int64_t(b) * c * t * h * w;
  launch_kernel(bcthw_to_frames_kernel, dim3(grid_for(total)), dim3(256), 0, static_cast<cudaStream_t>(s), in, in_dtype, static_cast<__nv_bfloat16*>(out), b, c, t, h, w, scale);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : cuda_fail(e, "t2v_bcthw_to_frames launch");
}

extern "C" int t2v_bcthw_to_frames_mix(const void* in, int32_t in_dtype, void* out, int32_t b, int32_t c,
                                       int32_t t, int32_t h, int32_t w, float scale, const float* mix,
                                       const float* bias, t2v_stream_t s) {
  if (!in || !out || !mix) return fail(-1, "t2v_bcthw_to_frames_mix: null pointer");
  if (c < 1 || c > 8) return fail(-2, "t2v_bcthw_to_frames_mix: c must be in [1,8]");
  const int64_t total = int64_t(b) * t * h * w;
  launch_kernel(bcthw_to_frames_mix_kernel, dim3(grid_for(total)), dim3(256), 0, static_cast<cudaStream_t>(s), in, in_dtype, static_cast<__nv_bfloat16*>(out), b, c, t, h, w, scale, mix, bias);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : cuda_fail(e, "t2v_bcthw_to_frames_mix launch");
}

extern "C" int t2v_bcthw_to_frames_pad(const void* in, int32_t in_dtype, void* out, int32_t b, int32_t c, int32_t c_pad,
                                       int32_t t, int32_t h, int32_t w, float scale, t2v_stream_t s) {
  if (!in || !out) return fail(-1, "t2v_bcthw_to_frames_pad: null pointer");
  if (c < 1 || c_pad < c || c_pad > 64) return fail(-2, "t2v_bcthw_to_frames_pad: need 1 <= c <= c_pad <= 64");
  const int64_t total = int64_t(b) * t * h * w;
  launch_kernel(bcthw_to_frames_pad_kernel, dim3(grid_for(total)), dim3(256), 0, static_cast<cudaStream_t>(s), in, in_dtype, static_cast<__nv_bfloat16*>(out), b, c, c_pad, t, h, w, scale);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : cuda_fail(e, "t2v_bcthw_to_frames_pad launch");
}

extern "C" int t2v_frames_to_bcthw(const void* in, int32_t c_pad, void* out, int32_t out_dtype, int32_t b,
                                   int32_t c, int32_t t, int32_t h, int32_t w, t2v_stream_t s) {
  if (!in || !out) return fail(-1, "t2v_frames_to_bcthw: null pointer");
  const int64_t total = int64_t(b) * c * t * h * w;
  launch_kernel(frames_to_bcthw_kernel, dim3(grid_for(total)), dim3(256), 0, static_cast<cudaStream_t>(s), static_cast<const __nv_bfloat16*>(in), c_pad, out, out_dtype, b, c, t, h, w);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : cuda_fail(e, "t2v_frames_to_bcthw launch");
}

extern "C" int t2v_upsample_nearest2x(const void* in, void* out, int32_t n, int32_t h, int32_t w, int32_t c,
                                      t2v_stream_t s) {
  if (!in || !out || c % 8) return fail(-1, "t2v_upsample_nearest2x: bad argument");
  const int64_t total = int64_t(n) * 4 * h * w * (c / 8);
  launch_kernel(upsample2x_kernel, dim3(grid_for(total)), dim3(256), 0, static_cast<cudaStream_t>(s), static_cast<const uint4*>(in), static_cast<uint4*>(out), n, h, w, c / 8);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : cuda_fail(e, "t2v_upsample_nearest2x launch");
}

extern "C" int t2v_concat_channels(const void* a, int32_t ca, const void* b, int32_t cb, void* out,
                                   int64_t rows, t2v_stream_t s) {
  if (!a || !b || !out || ca % 8 || cb % 8) return fail(-1, "t2v_concat_channels: bad argument");
  launch_kernel(concat_kernel, dim3(grid_for(rows * ((ca + cb) / 8))), dim3(256), 0, static_cast<cudaStream_t>(s), static_cast<const uint4*>(a), ca / 8, static_cast<const uint4*>(b), cb / 8, static_cast<uint4*>(out), rows);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : cuda_fail(e, "t2v_concat_channels launch");
}

extern "C" int t2v_softmax_rows(void* x, int64_t rows, int32_t cols, int64_t row_stride, float scale,
                                t2v_stream_t s) {
  if (!x || rows <= 0 || cols <= 0) return fail(-1, "t2v_softmax_rows: bad argument");
  if (rows > 0x7fffffff) return fail(-2, "t2v_softmax_rows: too many rows");
  launch_kernel(softmax_rows_kernel, dim3(unsigned(rows)), dim3(256), 0, static_cast<cudaStream_t>(s), static_cast<__nv_bfloat16*>(x), cols, row_stride, scale);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : cuda_fail(e, "t2v_softmax_rows launch");
}

extern "C" int t2v_gaussian_sample(const float* moments, const float* noise, void* out, int32_t out_dtype, int32_t b,
                                   int32_t t, int32_t h, int32_t w, int32_t zc, float scale, t2v_stream_t s) {
  if (!moments || !out || b < 1 || t < 1 || h < 1 || w < 1 || zc < 1) return fail(-1, "t2v_gaussian_sample: bad argument");
  if (out_dtype < 0 || out_dtype > 2) return fail(-2, "t2v_gaussian_sample: dtype must be 0 (bf16), 1 (fp16) or 2 (fp32)");
  const int64_t total = int64_t(b) * t * h * w * zc;
  launch_kernel(gaussian_sample_kernel, dim3(grid_for(total)), dim3(256), 0, static_cast<cudaStream_t>(s), moments, noise, out,
                out_dtype, b, t, h, w, zc, scale);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : cuda_fail(e, "t2v_gaussian_sample launch");
}

extern "C" int t2v_embedding_gather(const void* table, const void* pos, int32_t dtype, const int64_t* ids, void* out, int64_t n,
                                    int32_t width, int32_t ctx, int32_t vocab, t2v_stream_t s) {
  if (!table || !ids || !out || n < 1 || width < 1 || ctx < 1 || vocab < 1 || dtype < 0 || dtype > 2)
    return fail(-1, "t2v_embedding_gather: bad argument");
  launch_kernel(embedding_gather_kernel, dim3(grid_for(n * width)), dim3(256), 0, static_cast<cudaStream_t>(s), table, pos, dtype, ids,
                static_cast<__nv_bfloat16*>(out), n, width, ctx, vocab);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : cuda_fail(e, "t2v_embedding_gather launch");
}

extern "C" int t2v_video_to_uint8(const void* video, int32_t dtype, uint8_t* out, int32_t b, int32_t t, int32_t h, int32_t w,
                                  t2v_stream_t s) {
  if (!video || !out || b < 1 || t < 1 || h < 1 || w < 1 || dtype < 0 || dtype > 2) return fail(-1, "t2v_video_to_uint8: bad argument");
  launch_kernel(video_to_uint8_kernel, dim3(grid_for(int64_t(b) * t * h * w * 3)), dim3(256), 0, static_cast<cudaStream_t>(s), video, dtype,
                out, b, t, h, w);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : cuda_fail(e, "t2v_video_to_uint8 launch");
}

extern "C" int t2v_scale_add_rows(const void* x, const void* y, const float* a, const float* b, void* out, int64_t rows,
                                  int64_t row_len, int32_t dtype, t2v_stream_t s) {
  if (!x || !a || !out || rows < 1 || row_len < 1 || (y && !b)) return fail(-1, "t2v_scale_add_rows: bad argument");
  if (dtype < 0 || dtype > 2) return fail(-2, "t2v_scale_add_rows: dtype must be 0 (bf16), 1 (fp16) or 2 (fp32)");
  launch_kernel(scale_add_rows_kernel, dim3(grid_for(rows * row_len)), dim3(256), 0, static_cast<cudaStream_t>(s), x, y, a, b, out,
                rows, row_len, dtype);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : cuda_fail(e, "t2v_scale_add_rows launch");
}

extern "C" int t2v_lcm_step(const void* x, const void* eps, const void* noise, void* prev, void* denoised,
                            int64_t n, int32_t dtype, float inv_sqrt_alpha_t, float sqrt_beta_t, float c_skip,
                            float c_out, float sqrt_alpha_prev, float sqrt_beta_prev, t2v_stream_t s) {
  if (!x || !eps || !prev || !denoised || n <= 0) return fail(-1, "t2v_lcm_step: bad argument");
  if (dtype < 0 || dtype > 2) return fail(-2, "t2v_lcm_step: dtype must be 0 (bf16), 1 (fp16) or 2 (fp32)");
  launch_kernel(lcm_step_kernel, dim3(grid_for(n)), dim3(256), 0, static_cast<cudaStream_t>(s), x, eps, noise, prev, denoised, n, dtype, inv_sqrt_alpha_t, sqrt_beta_t, c_skip, c_out, sqrt_alpha_prev, sqrt_beta_prev);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : cuda_fail(e, "t2v_lcm_step launch");
}

extern "C" int t2v_pack_conv_weight(const void* w, int32_t w_dtype, void* out, int32_t cout, int32_t cin,
                                    int32_t taps, t2v_stream_t s) {
  if (!w || !out) return fail(-1, "t2v_pack_conv_weight: null pointer");
  launch_kernel(pack_conv_weight_kernel, dim3(grid_for(int64_t(cout) * cin * taps)), dim3(256), 0, static_cast<cudaStream_t>(s), w, w_dtype, static_cast<__nv_bfloat16*>(out), cout, cin, taps);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : cuda_fail(e, "t2v_pack_conv_weight launch");
}

extern "C" int t2v_pack_geglu_rows(const void* w, int32_t w_dtype, void* out, const void* bias,
                                   int32_t bias_dtype, float* bias_out, int32_t inner, int32_t k,
                                   t2v_stream_t s) {
  if (!w || !out || inner % 16) return fail(-1, "t2v_pack_geglu_rows: inner must be a multiple of 16");
  launch_kernel(pack_geglu_kernel, dim3(grid_for(int64_t(2) * inner * k)), dim3(256), 0, static_cast<cudaStream_t>(s), w, w_dtype, static_cast<__nv_bfloat16*>(out), bias, bias_dtype, bias_out, inner, k);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : cuda_fail(e, "t2v_pack_geglu_rows launch");
}
