// Short-sequence attention backward (len <= 16, head_dim 64): the adjoint of attn_short.cu for the temporal
// self-attention layers (every pixel attends over its T = 16 frames; attention.py:471-513).
//
// One warp per (sequence, head).  The problem is five 16x16x64 products — far below one tcgen05 tile — and the kernel is
// bound by its 7 x 2 KB of global traffic per task, so the arithmetic stays on the FMA pipe in fp32:
//   S = scale Q K^T, P = softmax(S) (recomputed), dP = dO V^T, delta_i = sum_j P_ij dP_ij, dS = P (dP - delta) scale
//   dV = P^T dO, dK = dS^T Q, dQ = dS K
// Q / K / V / dO are read in place through the forward's strides (token stride = H*W*C) and dQ / dK / dV written the same way.
#include <cuda_bf16.h>

#include "../../include/t2v_b200.h"
#include "host_common.h"
#include "ptx.cuh"

namespace t2v {

namespace {

constexpr int kWarps = 4;
constexpr int kL = 16;

struct Tile {          // one warp's working set
  float q[kL][65], k[kL][65], v[kL][65], g[kL][65];   // fp32, +1 padding: column reads are conflict free
  float p[kL][kL + 1], ds[kL][kL + 1];
};

__device__ __forceinline__ void load_tile(float (&dst)[kL][65], const __nv_bfloat16* base, int64_t stride_t, int len, int lane) {
  // 16 rows x 8 chunks of 16 bytes; lane handles chunk (lane & 7) of rows (lane >> 3) + 4 i
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int row = (lane >> 3) + 4 * i;
    const int ch = lane & 7;
    uint4 u = make_uint4(0, 0, 0, 0);
    if (row < len) u = __ldg(reinterpret_cast<const uint4*>(base + int64_t(row) * stride_t) + ch);
    float* d = &dst[row][ch * 8];
    d[0] = bf16_lo(u.x); d[1] = bf16_hi(u.x); d[2] = bf16_lo(u.y); d[3] = bf16_hi(u.y);
    d[4] = bf16_lo(u.z); d[5] = bf16_hi(u.z); d[6] = bf16_lo(u.w); d[7] = bf16_hi(u.w);
  }
}

__global__ void __launch_bounds__(kWarps * 32) attn_short_bwd_kernel(const T2VShortAttnBwdDesc d, int64_t n_tasks) {
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ __align__(16) uint8_t smem_raw[];
  Tile* tiles = reinterpret_cast<Tile*>(smem_raw);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t task = int64_t(blockIdx.x) * kWarps + warp;
  if (task >= n_tasks) return;
  Tile& t = tiles[warp];
  const T2VShortAttnDesc& f = d.fwd;
  const int len = f.len;
  const int h = int(task % f.heads);
  const int64_t seq = task / f.heads;
  const int64_t inner = seq % f.n_seq_inner, outer = seq / f.n_seq_inner;
  const int64_t qo = outer * f.q_stride_outer + inner * f.q_stride_inner + int64_t(h) * f.q_stride_h;
  const int64_t ko = outer * f.k_stride_outer + inner * f.k_stride_inner + int64_t(h) * f.k_stride_h;
  const int64_t vo = outer * f.v_stride_outer + inner * f.v_stride_inner + int64_t(h) * f.v_stride_h;
  const int64_t oo = outer * f.o_stride_outer + inner * f.o_stride_inner + int64_t(h) * f.o_stride_h;
  load_tile(t.q, static_cast<const __nv_bfloat16*>(f.q) + qo, f.q_stride_t, len, lane);
  load_tile(t.k, static_cast<const __nv_bfloat16*>(f.k) + ko, f.k_stride_t, len, lane);
  load_tile(t.v, static_cast<const __nv_bfloat16*>(f.v) + vo, f.v_stride_t, len, lane);
  load_tile(t.g, static_cast<const __nv_bfloat16*>(d.d_o) + oo, f.o_stride_t, len, lane);
  __syncwarp();
  // ---- S and dP: lane owns row i = lane >> 1, columns j0..j0+7 with j0 = (lane & 1) * 8
  const int i = lane >> 1, j0 = (lane & 1) * 8;
  float s[8], dp[8];
#pragma unroll
  for (int jj = 0; jj < 8; ++jj) s[jj] = dp[jj] = 0.f;
  for (int c = 0; c < 64; ++c) {
    const float qv = t.q[i][c], gv = t.g[i][c];
#pragma unroll
    for (int jj = 0; jj < 8; ++jj) {
      s[jj] = fmaf(qv, t.k[j0 + jj][c], s[jj]);
      dp[jj] = fmaf(gv, t.v[j0 + jj][c], dp[jj]);
    }
  }
  float mx = -INFINITY;
#pragma unroll
  for (int jj = 0; jj < 8; ++jj) {
    s[jj] = (j0 + jj < len) ? s[jj] * f.scale : -INFINITY;
    mx = fmaxf(mx, s[jj]);
  }
  mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
  float sum = 0.f;
#pragma unroll
  for (int jj = 0; jj < 8; ++jj) {
    s[jj] = __expf(s[jj] - mx);
    sum += s[jj];
  }
  sum += __shfl_xor_sync(0xffffffffu, sum, 1);
  const float inv = 1.0f / sum;
  float delta = 0.f;
#pragma unroll
  for (int jj = 0; jj < 8; ++jj) {
    s[jj] *= inv;
    delta = fmaf(s[jj], dp[jj], delta);
  }
  delta += __shfl_xor_sync(0xffffffffu, delta, 1);
  const bool row_ok = i < len;
#pragma unroll
  for (int jj = 0; jj < 8; ++jj) {
    const float pv = row_ok ? s[jj] : 0.f;
    t.p[i][j0 + jj] = pv;
    t.ds[i][j0 + jj] = pv * (dp[jj] - delta) * f.scale;
  }
  __syncwarp();
  // ---- dQ / dK / dV: lane owns row r = lane >> 1, channels c0..c0+31 with c0 = (lane & 1) * 32
  const int r = lane >> 1, c0 = (lane & 1) * 32;
  auto store_row = [&](void* base, int64_t off, int64_t stride_t, const float (&acc)[32]) {
    if (r >= len) return;
    uint4* o = reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(base) + off + int64_t(r) * stride_t + c0);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      uint4 u;
      u.x = pack_bf16(acc[q * 8 + 0], acc[q * 8 + 1]);
      u.y = pack_bf16(acc[q * 8 + 2], acc[q * 8 + 3]);
      u.z = pack_bf16(acc[q * 8 + 4], acc[q * 8 + 5]);
      u.w = pack_bf16(acc[q * 8 + 6], acc[q * 8 + 7]);
      o[q] = u;
    }
  };
  float acc[32];
  // dQ[r] = sum_j dS[r][j] K[j]
#pragma unroll
  for (int c = 0; c < 32; ++c) acc[c] = 0.f;
  for (int j = 0; j < kL; ++j) {
    const float w = t.ds[r][j];
#pragma unroll
    for (int c = 0; c < 32; ++c) acc[c] = fmaf(w, t.k[j][c0 + c], acc[c]);
  }
  store_row(d.dq, qo, f.q_stride_t, acc);
  // dK[r] = sum_i dS[i][r] Q[i]
#pragma unroll
  for (int c = 0; c < 32; ++c) acc[c] = 0.f;
  for (int ii = 0; ii < kL; ++ii) {
    const float w = t.ds[ii][r];
#pragma unroll
    for (int c = 0; c < 32; ++c) acc[c] = fmaf(w, t.q[ii][c0 + c], acc[c]);
  }
  store_row(d.dk, ko, f.k_stride_t, acc);
  // dV[r] = sum_i P[i][r] dO[i]
#pragma unroll
  for (int c = 0; c < 32; ++c) acc[c] = 0.f;
  for (int ii = 0; ii < kL; ++ii) {
    const float w = t.p[ii][r];
#pragma unroll
    for (int c = 0; c < 32; ++c) acc[c] = fmaf(w, t.g[ii][c0 + c], acc[c]);
  }
  store_row(d.dv, vo, f.v_stride_t, acc);
}

}  // namespace
}  // namespace t2v

extern "C" int t2v_attn_short_bwd(const T2VShortAttnBwdDesc* d, t2v_stream_t stream_) {
  using namespace t2v;
  if (!d || !d->fwd.q || !d->fwd.k || !d->fwd.v || !d->d_o || !d->dq || !d->dk || !d->dv) return fail(-1, "t2v_attn_short_bwd: null pointer");
  const T2VShortAttnDesc& f = d->fwd;
  if (f.len < 1 || f.len > kL) return fail(-2, "t2v_attn_short_bwd: len must be in [1,16] (got %d)", f.len);
  if (f.heads < 1 || f.n_seq_inner < 1 || f.n_seq_outer < 1) return fail(-3, "t2v_attn_short_bwd: bad sizes");
  const int64_t strides[] = {f.q_stride_outer, f.q_stride_inner, f.q_stride_t, f.q_stride_h, f.k_stride_outer, f.k_stride_inner,
                             f.k_stride_t,     f.k_stride_h,     f.v_stride_outer, f.v_stride_inner, f.v_stride_t, f.v_stride_h,
                             f.o_stride_outer, f.o_stride_inner, f.o_stride_t, f.o_stride_h};
  for (int64_t s : strides)
    if (s % 8) return fail(-4, "t2v_attn_short_bwd: strides must be multiples of 8 elements");
  const void* ptrs[] = {f.q, f.k, f.v, d->d_o, d->dq, d->dk, d->dv};
  for (const void* p : ptrs)
    if (reinterpret_cast<uintptr_t>(p) & 15) return fail(-5, "t2v_attn_short_bwd: pointers must be 16-byte aligned");
  const int64_t n_tasks = int64_t(f.n_seq_outer) * f.n_seq_inner * f.heads;
  const size_t smem = sizeof(Tile) * kWarps;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(attn_short_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(attn_short_bwd)");
    configured = true;
  }
  const int64_t blocks = (n_tasks + kWarps - 1) / kWarps;
  launch_kernel(attn_short_bwd_kernel, dim3(unsigned(blocks)), dim3(kWarps * 32), smem, static_cast<cudaStream_t>(stream_), *d, n_tasks);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : cuda_fail(e, "t2v_attn_short_bwd launch");
}
