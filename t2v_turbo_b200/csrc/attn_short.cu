// Short-sequence attention (len <= 32, head_dim 64) for the temporal transformer layers:
// every pixel attends over its T frames.  Sequences are read in place from the [B*T, H*W, C]
// activation through strides (token stride = H*W*C) — no "(b hw) t c" regrouping copy.
//
// v1 mapping: one thread = one query row; a 128-thread block handles 128/len_pad (sequence, head)
// tasks whose K and V rows are staged in shared memory with 16-byte coalesced loads.
// The work is HBM/latency-bound (4 x 128 B per token); fp32 softmax and accumulation.
#include <cuda_bf16.h>

#include "../../include/t2v_b200.h"
#include "host_common.h"
#include "ptx.cuh"

namespace t2v {

constexpr int kSaThreads = 128;

template <int LEN_PAD>  // 16 or 32
__global__ void __launch_bounds__(kSaThreads) attn_short_kernel(const T2VShortAttnDesc d, int64_t n_tasks) {
  constexpr int TASKS = kSaThreads / LEN_PAD;
  __shared__ __align__(16) __nv_bfloat16 s_k[TASKS][LEN_PAD][64];
  __shared__ __align__(16) __nv_bfloat16 s_v[TASKS][LEN_PAD][64];
  const int len = d.len;
  const int64_t task0 = int64_t(blockIdx.x) * TASKS;

  // cooperative K/V staging: one 16-byte chunk per thread per iteration
  for (int idx = threadIdx.x; idx < TASKS * LEN_PAD * 8; idx += kSaThreads) {
    const int chunk = idx & 7;
    const int row = (idx >> 3) % LEN_PAD;
    const int tl = idx / (8 * LEN_PAD);
    const int64_t task = task0 + tl;
    uint4 kv = make_uint4(0, 0, 0, 0), vv = make_uint4(0, 0, 0, 0);
    if (task < n_tasks && row < len) {
      const int h = int(task % d.heads);
      const int64_t seq = task / d.heads;
      const int64_t inner = seq % d.n_seq_inner, outer = seq / d.n_seq_inner;
      const __nv_bfloat16* kp = static_cast<const __nv_bfloat16*>(d.k) + outer * d.k_stride_outer +
                                inner * d.k_stride_inner + int64_t(row) * d.k_stride_t + h * d.k_stride_h;
      const __nv_bfloat16* vp = static_cast<const __nv_bfloat16*>(d.v) + outer * d.v_stride_outer +
                                inner * d.v_stride_inner + int64_t(row) * d.v_stride_t + h * d.v_stride_h;
      kv = __ldg(reinterpret_cast<const uint4*>(kp) + chunk);
      vv = __ldg(reinterpret_cast<const uint4*>(vp) + chunk);
    }
    *reinterpret_cast<uint4*>(&s_k[tl][row][chunk * 8]) = kv;
    *reinterpret_cast<uint4*>(&s_v[tl][row][chunk * 8]) = vv;
  }
  __syncthreads();

  const int tl = threadIdx.x / LEN_PAD;
  const int qi = threadIdx.x % LEN_PAD;
  const int64_t task = task0 + tl;
  if (task >= n_tasks || qi >= len) return;
  const int h = int(task % d.heads);
  const int64_t seq = task / d.heads;
  const int64_t inner = seq % d.n_seq_inner, outer = seq / d.n_seq_inner;
  const __nv_bfloat16* qp = static_cast<const __nv_bfloat16*>(d.q) + outer * d.q_stride_outer +
                            inner * d.q_stride_inner + int64_t(qi) * d.q_stride_t + h * d.q_stride_h;
  uint4 qv[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) qv[c] = __ldg(reinterpret_cast<const uint4*>(qp) + c);

  float s[LEN_PAD];
  float mx = -INFINITY;
#pragma unroll
  for (int j = 0; j < LEN_PAD; ++j) {
    float acc = 0.f;
    const uint4* kr = reinterpret_cast<const uint4*>(&s_k[tl][j][0]);
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const uint4 kk = kr[c];
      acc = fmaf(bf16_lo(qv[c].x), bf16_lo(kk.x), acc);
      acc = fmaf(bf16_hi(qv[c].x), bf16_hi(kk.x), acc);
      acc = fmaf(bf16_lo(qv[c].y), bf16_lo(kk.y), acc);
      acc = fmaf(bf16_hi(qv[c].y), bf16_hi(kk.y), acc);
      acc = fmaf(bf16_lo(qv[c].z), bf16_lo(kk.z), acc);
      acc = fmaf(bf16_hi(qv[c].z), bf16_hi(kk.z), acc);
      acc = fmaf(bf16_lo(qv[c].w), bf16_lo(kk.w), acc);
      acc = fmaf(bf16_hi(qv[c].w), bf16_hi(kk.w), acc);
    }
    s[j] = j < len ? acc * d.scale : -INFINITY;
    mx = fmaxf(mx, s[j]);
  }
  float sum = 0.f;
#pragma unroll
  for (int j = 0; j < LEN_PAD; ++j) {
    s[j] = __expf(s[j] - mx);
    sum += s[j];
  }
  const float inv = 1.0f / sum;
  __nv_bfloat16* op = static_cast<__nv_bfloat16*>(d.o) + outer * d.o_stride_outer +
                      inner * d.o_stride_inner + int64_t(qi) * d.o_stride_t + h * d.o_stride_h;
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    float o[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) o[e] = 0.f;
#pragma unroll
    for (int j = 0; j < LEN_PAD; ++j) {
      const uint4 vv = *reinterpret_cast<const uint4*>(&s_v[tl][j][c * 8]);
      const float pj = s[j];
      o[0] = fmaf(pj, bf16_lo(vv.x), o[0]);
      o[1] = fmaf(pj, bf16_hi(vv.x), o[1]);
      o[2] = fmaf(pj, bf16_lo(vv.y), o[2]);
      o[3] = fmaf(pj, bf16_hi(vv.y), o[3]);
      o[4] = fmaf(pj, bf16_lo(vv.z), o[4]);
      o[5] = fmaf(pj, bf16_hi(vv.z), o[5]);
      o[6] = fmaf(pj, bf16_lo(vv.w), o[6]);
      o[7] = fmaf(pj, bf16_hi(vv.w), o[7]);
    }
    uint4 ov;
    ov.x = pack_bf16(o[0] * inv, o[1] * inv);
    ov.y = pack_bf16(o[2] * inv, o[3] * inv);
    ov.z = pack_bf16(o[4] * inv, o[5] * inv);
    ov.w = pack_bf16(o[6] * inv, o[7] * inv);
    reinterpret_cast<uint4*>(op)[c] = ov;
  }
}

}  // namespace t2v

extern "C" int t2v_attn_short_fwd(const T2VShortAttnDesc* d, t2v_stream_t stream_) {
  using namespace t2v;
  if (!d || !d->q || !d->k || !d->v || !d->o) return fail(-1, "t2v_attn_short_fwd: null pointer");
  if (d->len < 1 || d->len > 32) return fail(-2, "t2v_attn_short_fwd: len must be in [1,32] (got %d)", d->len);
  if (d->heads < 1 || d->n_seq_inner < 1 || d->n_seq_outer < 1) return fail(-3, "t2v_attn_short_fwd: bad sizes");
  const int64_t strides[] = {d->q_stride_outer, d->q_stride_inner, d->q_stride_t, d->q_stride_h,
                             d->k_stride_outer, d->k_stride_inner, d->k_stride_t, d->k_stride_h,
                             d->v_stride_outer, d->v_stride_inner, d->v_stride_t, d->v_stride_h,
                             d->o_stride_outer, d->o_stride_inner, d->o_stride_t, d->o_stride_h};
  for (int64_t s : strides)
    if (s % 8) return fail(-4, "t2v_attn_short_fwd: strides must be multiples of 8 elements");
  const void* ptrs[] = {d->q, d->k, d->v, d->o};
  for (const void* p : ptrs)
    if (reinterpret_cast<uintptr_t>(p) & 15) return fail(-5, "t2v_attn_short_fwd: pointers must be 16-byte aligned");
  const int64_t n_tasks = int64_t(d->n_seq_outer) * d->n_seq_inner * d->heads;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (d->len <= 16) {
    const int64_t blocks = (n_tasks + (kSaThreads / 16) - 1) / (kSaThreads / 16);
    attn_short_kernel<16><<<unsigned(blocks), kSaThreads, 0, stream>>>(*d, n_tasks);
  } else {
    const int64_t blocks = (n_tasks + (kSaThreads / 32) - 1) / (kSaThreads / 32);
    attn_short_kernel<32><<<unsigned(blocks), kSaThreads, 0, stream>>>(*d, n_tasks);
  }
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : cuda_fail(e, "t2v_attn_short_fwd launch");
}
