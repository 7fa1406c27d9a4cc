// TEST INFRASTRUCTURE (never part of the product build): a minimal host emulation of the CUDA SIMT model, just large enough to
// compile the NON-tensor-core kernels of csrc/train_bwd.cu and csrc/train_full.cu with g++ (-DT2V_HOST_EMU -x c++) and run them on
// CPU threads, so that their indexing / reduction logic is executed in the `-m "not gpu"` suite (tests/test_cuda_emu_cpu.py).
//
// Model: blocks run one after another; inside a block every CUDA thread is an OS thread.  __syncthreads is a std::barrier over
// the block, warp shuffles exchange through a per-warp buffer between two warp barriers, __shared__ variables are function-level
// statics (one block is live at a time), atomicAdd is std::atomic_ref.  A thread that returns from the kernel drops out of the
// barriers, as an exited CUDA thread does.  What this does NOT model: memory ordering subtleties, bank conflicts, PDL, TMA,
// tcgen05 — kernels that use those are GPU-only.  The emulator itself is validated on kernels that ARE parity-tested on B200
// (t2v_groupnorm_bwd, t2v_layernorm_bwd, t2v_colsum_samples, t2v_geglu) before it is trusted on the GPU-unverified ones.
#pragma once
#include <math.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <barrier>
#include <memory>
#include <thread>
#include <vector>

struct dim3 {
  unsigned x, y, z;
  constexpr dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct uint4 { uint32_t x, y, z, w; };
struct float2 { float x, y; };
struct float4 { float x, y, z, w; };
struct uint2 { uint32_t x, y; };
inline uint2 make_uint2(uint32_t x, uint32_t y) { return uint2{x, y}; }
inline float2 make_float2(float x, float y) { return float2{x, y}; }
inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { return uint4{x, y, z, w}; }
inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
struct __nv_bfloat16 { uint16_t bits; };
static_assert(sizeof(__nv_bfloat16) == 2, "bf16 storage");

typedef void* cudaStream_t;
typedef int cudaError_t;
constexpr cudaError_t cudaSuccess = 0;
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
inline const char* cudaGetErrorString(cudaError_t) { return "host emulation"; }

namespace emu {
struct Warp {
  uint64_t buf[32];
  std::barrier<> bar;
  explicit Warp(int n) : bar(n) {}
};
struct Ctx {
  dim3 tid, bid, bdim, gdim;
  std::barrier<>* block_bar = nullptr;
  Warp* warp = nullptr;
  void* dyn_smem = nullptr;          // the block's dynamic shared memory (launch parameter `smem` bytes)
};
inline thread_local Ctx ctx;
}  // namespace emu

#define threadIdx (emu::ctx.tid)
#define blockIdx (emu::ctx.bid)
#define blockDim (emu::ctx.bdim)
#define gridDim (emu::ctx.gdim)
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define __shared__ static
#define T2V_DYN_SMEM(type, name) type* name = reinterpret_cast<type*>(emu::ctx.dyn_smem)
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
template <class F> inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) { return cudaSuccess; }

inline void __syncthreads() { emu::ctx.block_bar->arrive_and_wait(); }

template <class T>
inline T __shfl_xor_sync(unsigned, T v, int lane_mask) {
  static_assert(sizeof(T) <= 8, "shuffle of at most 64 bits");
  emu::Warp* w = emu::ctx.warp;
  const int lane = int(emu::ctx.tid.x & 31u);
  uint64_t b = 0;
  memcpy(&b, &v, sizeof(T));
  w->buf[lane] = b;
  w->bar.arrive_and_wait();
  const uint64_t r = w->buf[lane ^ lane_mask];
  w->bar.arrive_and_wait();
  T out;
  memcpy(&out, &r, sizeof(T));
  return out;
}

inline float atomicAdd(float* p, float v) { return std::atomic_ref<float>(*p).fetch_add(v, std::memory_order_relaxed); }
template <class T> inline T __ldg(const T* p) { return *p; }
template <class T> inline T __ldcg(const T* p) { return *p; }
inline float __expf(float x) { return expf(x); }
inline float rsqrtf(float x) { return 1.0f / sqrtf(x); }
inline uint32_t __umulhi(uint32_t a, uint32_t b) { return uint32_t((uint64_t(a) * uint64_t(b)) >> 32); }
inline void __syncwarp(unsigned = 0xffffffffu) { emu::ctx.warp->bar.arrive_and_wait(); }
inline float __uint_as_float(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
inline uint32_t __float_as_uint(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }

struct __half { _Float16 v; };
static_assert(sizeof(__half) == 2, "fp16 storage");
inline float __half2float(__half h) { return float(h.v); }
inline __half __float2half_rn(float f) { return __half{_Float16(f)}; }
inline float __bfloat162float(__nv_bfloat16 b) { return __uint_as_float(uint32_t(b.bits) << 16); }
inline __nv_bfloat16 __float2bfloat16_rn(float f) {
  uint32_t u = __float_as_uint(f);
  if ((u & 0x7FFFFFFFu) > 0x7F800000u) return __nv_bfloat16{uint16_t((u >> 16) | 0x40u)};
  u += 0x7FFFu + ((u >> 16) & 1u);
  return __nv_bfloat16{uint16_t(u >> 16)};
}

namespace t2v {

inline char* last_error_buf() { static thread_local char buf[512]; return buf; }
inline int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(last_error_buf(), 512, fmt, ap);
  va_end(ap);
  return code;
}
inline int cuda_fail(cudaError_t e, const char* what) { return fail(int(e), "%s", what); }
inline int num_sms() {                     // a small "device" keeps the emulated grids small
  const char* e = getenv("T2V_EMU_SMS");
  return e ? atoi(e) : 2;
}
inline void pdl_launch_dependents() {}
inline void pdl_wait() {}

inline uint16_t f32_to_bf16_rne(float f) {
  uint32_t u = __float_as_uint(f);
  if ((u & 0x7FFFFFFFu) > 0x7F800000u) return uint16_t((u >> 16) | 0x40u);     // NaN
  u += 0x7FFFu + ((u >> 16) & 1u);
  return uint16_t(u >> 16);
}
inline uint32_t pack_bf16(float a, float b) { return uint32_t(f32_to_bf16_rne(a)) | (uint32_t(f32_to_bf16_rne(b)) << 16); }
inline float bf16_lo(uint32_t u) { return __uint_as_float(u << 16); }
inline float bf16_hi(uint32_t u) { return __uint_as_float(u & 0xFFFF0000u); }

template <typename... KArgs, typename... Args>
inline cudaError_t launch_kernel(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t, Args... args) {
  const unsigned nt = block.x * block.y * block.z;
  std::vector<uint64_t> dyn((smem + 7) / 8 + 1);
  for (unsigned bz = 0; bz < grid.z; ++bz)
    for (unsigned by = 0; by < grid.y; ++by)
      for (unsigned bx = 0; bx < grid.x; ++bx) {
        std::barrier<> bar(nt);
        std::vector<std::unique_ptr<emu::Warp>> warps;
        for (unsigned w = 0; w * 32 < nt; ++w) warps.emplace_back(new emu::Warp(int(nt - w * 32 < 32 ? nt - w * 32 : 32)));
        std::vector<std::thread> threads;
        threads.reserve(nt);
        for (unsigned t = 0; t < nt; ++t)
          threads.emplace_back([&, t]() {
            emu::ctx.tid = dim3(t % block.x, (t / block.x) % block.y, t / (block.x * block.y));
            emu::ctx.bid = dim3(bx, by, bz);
            emu::ctx.bdim = block;
            emu::ctx.gdim = grid;
            emu::ctx.block_bar = &bar;
            emu::ctx.warp = warps[t / 32].get();
            emu::ctx.dyn_smem = dyn.data();
            kernel(KArgs(args)...);
            emu::ctx.warp->bar.arrive_and_drop();      // an exited thread no longer takes part in barriers
            bar.arrive_and_drop();
          });
        for (auto& th : threads) th.join();
      }
  return cudaSuccess;
}

}  // namespace t2v
