// Short-sequence attention (len <= 32, head_dim 64) for the temporal transformer layers:
// every pixel attends over its T frames.  Sequences are read in place from the [B*T, H*W, C]
// activation through strides (token stride = H*W*C) — no "(b hw) t c" regrouping copy.
//
// len <= 16 (the T2V-Turbo case, T = 16): one WARP per (sequence, head) task on mma.sync m16n8k16
// (the 16x16x64 problem is exactly one MMA row block; tcgen05's 128-row tiles would waste 8x):
//   cp.async 16-byte coalesced loads of the Q/K/V rows into XOR-swizzled smem, ldmatrix fragments,
//   S = Q K^T (8 MMAs), quad-shuffle softmax in fp32, P re-used in registers as the A operand,
//   O = P V (8 MMAs, V through ldmatrix.trans), O staged through smem for 16-byte coalesced stores.
//   The work is HBM-bound: 4 x 128 B per token per layer.
// 16 < len <= 32: scalar kernel (thread = query row), kept for generality.
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "../../include/t2v_b200.h"
#include "host_common.h"
#include "ptx.cuh"

namespace t2v {

__device__ __forceinline__ void store_prob(void* p, int dtype, int64_t i, float v) {
  if (dtype == 0) static_cast<__nv_bfloat16*>(p)[i] = __float2bfloat16_rn(v);
  else if (dtype == 1) static_cast<__half*>(p)[i] = __float2half_rn(v);
  else static_cast<float*>(p)[i] = v;
}

// ------------------------------------------------------------------------------------------------
// mma.sync path, len <= 16
// ------------------------------------------------------------------------------------------------
constexpr int kSaWarps = 4;

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src, bool valid) {
  const uint32_t d = smem_u32(smem_dst);
  const int sz = valid ? 16 : 0;  // src-size 0 => zero fill
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gmem_src), "r"(sz) : "memory");
}
__device__ __forceinline__ void ldmatrix_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// tile = 16 rows x 64 bf16 (128 B per row), 16-byte chunk c of row r stored at chunk (c ^ (r & 7))
__device__ __forceinline__ uint32_t tile_off(int row, int chunk) { return uint32_t(row * 128 + ((chunk ^ (row & 7)) << 4)); }

__global__ void __launch_bounds__(kSaWarps * 32) attn_short_mma_kernel(const T2VShortAttnDesc d, int64_t n_tasks) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ __align__(128) uint8_t smem[kSaWarps][3][16 * 128];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t task = int64_t(blockIdx.x) * kSaWarps + warp;
  if (task >= n_tasks) return;
  const int len = d.len;
  const int h = int(task % d.heads);
  const int64_t seq = task / d.heads;
  const int64_t inner = seq % d.n_seq_inner, outer = seq / d.n_seq_inner;
  const __nv_bfloat16* qp = static_cast<const __nv_bfloat16*>(d.q) + outer * d.q_stride_outer + inner * d.q_stride_inner + h * d.q_stride_h;
  const __nv_bfloat16* kp = static_cast<const __nv_bfloat16*>(d.k) + outer * d.k_stride_outer + inner * d.k_stride_inner + h * d.k_stride_h;
  const __nv_bfloat16* vp = static_cast<const __nv_bfloat16*>(d.v) + outer * d.v_stride_outer + inner * d.v_stride_inner + h * d.v_stride_h;
  uint8_t* sQ = smem[warp][0];
  uint8_t* sK = smem[warp][1];
  uint8_t* sV = smem[warp][2];
  // 16 rows x 8 chunks per matrix = 128 chunks -> 4 per lane; 8 consecutive lanes cover one 128-byte row
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int id = lane + 32 * i;
    const int row = id >> 3, c = id & 7;
    const bool ok = row < len;
    const int rs = ok ? row : 0;
    cp_async16(sQ + tile_off(row, c), qp + int64_t(rs) * d.q_stride_t + c * 8, ok);
    cp_async16(sK + tile_off(row, c), kp + int64_t(rs) * d.k_stride_t + c * 8, ok);
    cp_async16(sV + tile_off(row, c), vp + int64_t(rs) * d.v_stride_t + c * 8, ok);
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncwarp();

  const int g = lane >> 2, t = lane & 3;
  // ---- S = Q K^T : 2 n-tiles (8 keys each) x 4 k-steps
  float s[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
  const uint32_t q_base = smem_u32(sQ), k_base = smem_u32(sK), v_base = smem_u32(sV);
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) {
    uint32_t a[4], b[4];
    {  // A fragment: matrices (rows 0-7 | 8-15) x (k chunk 2ks | 2ks+1)
      const int row = (lane & 7) + ((lane >> 3) & 1) * 8;
      const int chunk = 2 * ks + (lane >> 4);
      ldmatrix_x4(q_base + tile_off(row, chunk), a[0], a[1], a[2], a[3]);
    }
    {  // B fragments for both n-tiles: matrix m -> keys (m>>1)*8.., k chunk 2ks + (m&1)
      const int m = lane >> 3;
      const int row = (lane & 7) + (m >> 1) * 8;
      const int chunk = 2 * ks + (m & 1);
      ldmatrix_x4(k_base + tile_off(row, chunk), b[0], b[1], b[2], b[3]);
    }
    mma_bf16_16816(s[0], a, b[0], b[1]);
    mma_bf16_16816(s[1], a, b[2], b[3]);
  }
  // ---- softmax over the 16 keys of rows g and g+8 (values spread over the 4 lanes of a quad)
  const float sl2 = d.scale * 1.4426950408889634f;
  float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
  for (int j = 0; j < 2; ++j) {
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const bool ok = (j * 8 + 2 * t + e) < len;
      s[j][e] = ok ? s[j][e] * sl2 : -INFINITY;
      s[j][2 + e] = ok ? s[j][2 + e] * sl2 : -INFINITY;
      mx0 = fmaxf(mx0, s[j][e]);
      mx1 = fmaxf(mx1, s[j][2 + e]);
    }
  }
  mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
  mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
  mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
  mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
  float sum0 = 0.f, sum1 = 0.f;
#pragma unroll
  for (int j = 0; j < 2; ++j) {
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      s[j][e] = exp2f(s[j][e] - mx0);
      s[j][2 + e] = exp2f(s[j][2 + e] - mx1);
      sum0 += s[j][e];
      sum1 += s[j][2 + e];
    }
  }
  sum0 += __shfl_xor_sync(0xffffffffu, sum0, 1);
  sum0 += __shfl_xor_sync(0xffffffffu, sum0, 2);
  sum1 += __shfl_xor_sync(0xffffffffu, sum1, 1);
  sum1 += __shfl_xor_sync(0xffffffffu, sum1, 2);
  const float inv0 = 1.0f / sum0, inv1 = 1.0f / sum1;
  if (d.probs != nullptr) {
    // attention-probability export (attention.py:124-126, `record_attn_probs`): probs[(seq * heads + h)][query][key]
    const int64_t base = task * int64_t(len) * len;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int col = j * 8 + 2 * t + e;
        if (col < len) {
          if (g < len) store_prob(d.probs, d.probs_dtype, base + int64_t(g) * len + col, s[j][e] * inv0);
          if (g + 8 < len) store_prob(d.probs, d.probs_dtype, base + int64_t(g + 8) * len + col, s[j][2 + e] * inv1);
        }
      }
    }
  }
  // P (bf16) as the A operand of O = P V: k = key index
  uint32_t pa[4];
  pa[0] = pack_bf16(s[0][0], s[0][1]);
  pa[1] = pack_bf16(s[0][2], s[0][3]);
  pa[2] = pack_bf16(s[1][0], s[1][1]);
  pa[3] = pack_bf16(s[1][2], s[1][3]);
  // ---- O = P V : 8 n-tiles (8 channels each), one k-step (16 keys); V fragments via ldmatrix.trans
  float o[8][4];
#pragma unroll
  for (int j = 0; j < 8; ++j) o[j][0] = o[j][1] = o[j][2] = o[j][3] = 0.f;
#pragma unroll
  for (int jp = 0; jp < 4; ++jp) {
    uint32_t b[4];
    const int m = lane >> 3;
    const int row = (lane & 7) + (m & 1) * 8;  // key
    const int chunk = 2 * jp + (m >> 1);       // channel chunk
    ldmatrix_x4_trans(v_base + tile_off(row, chunk), b[0], b[1], b[2], b[3]);
    mma_bf16_16816(o[2 * jp], pa, b[0], b[1]);
    mma_bf16_16816(o[2 * jp + 1], pa, b[2], b[3]);
  }
  // ---- stage O (bf16) in the Q tile, then coalesced 16-byte stores
  __syncwarp();
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    // row g: channels j*8 + 2t, +1 ; row g+8 likewise
    *reinterpret_cast<uint32_t*>(sQ + tile_off(g, j) + t * 4) = pack_bf16(o[j][0] * inv0, o[j][1] * inv0);
    *reinterpret_cast<uint32_t*>(sQ + tile_off(g + 8, j) + t * 4) = pack_bf16(o[j][2] * inv1, o[j][3] * inv1);
  }
  __syncwarp();
  __nv_bfloat16* op = static_cast<__nv_bfloat16*>(d.o) + outer * d.o_stride_outer + inner * d.o_stride_inner + h * d.o_stride_h;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int id = lane + 32 * i;
    const int row = id >> 3, c = id & 7;
    if (row < len) {
      const uint4 val = *reinterpret_cast<const uint4*>(sQ + tile_off(row, c));
      *reinterpret_cast<uint4*>(op + int64_t(row) * d.o_stride_t + c * 8) = val;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// scalar path, 16 < len <= 32: thread = query row
// ------------------------------------------------------------------------------------------------
constexpr int kSaThreads = 128;

template <int LEN_PAD>
__global__ void __launch_bounds__(kSaThreads) attn_short_kernel(const T2VShortAttnDesc d, int64_t n_tasks) {
  pdl_launch_dependents();
  pdl_wait();
  constexpr int TASKS = kSaThreads / LEN_PAD;
  __shared__ __align__(16) __nv_bfloat16 s_k[TASKS][LEN_PAD][64];
  __shared__ __align__(16) __nv_bfloat16 s_v[TASKS][LEN_PAD][64];
  const int len = d.len;
  const int64_t task0 = int64_t(blockIdx.x) * TASKS;
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
  if (d.probs != nullptr) {
    const int64_t base = (task * int64_t(len) + qi) * len;
#pragma unroll
    for (int j = 0; j < LEN_PAD; ++j)
      if (j < len) store_prob(d.probs, d.probs_dtype, base + j, s[j] * inv);
  }
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
  if (d->probs && (d->probs_dtype < 0 || d->probs_dtype > 2)) return fail(-6, "t2v_attn_short_fwd: probs_dtype must be 0 (bf16), 1 (fp16) or 2 (fp32)");
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
    const int64_t blocks = (n_tasks + kSaWarps - 1) / kSaWarps;
    launch_kernel(attn_short_mma_kernel, dim3(unsigned(blocks)), dim3(kSaWarps * 32), 0, stream, *d, n_tasks);
  } else {
    const int64_t blocks = (n_tasks + (kSaThreads / 32) - 1) / (kSaThreads / 32);
    launch_kernel(attn_short_kernel<32>, dim3(unsigned(blocks)), dim3(kSaThreads), 0, stream, *d, n_tasks);
  }
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : cuda_fail(e, "t2v_attn_short_fwd launch");
}
