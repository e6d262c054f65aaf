// Common device/host helpers for the toad_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

#include "../../include/toad_b200.h"

#define TOAD_CUDA_TRY(expr)                                        \
  do {                                                             \
    cudaError_t _e = (expr);                                       \
    if (_e != cudaSuccess) return (int)_e;                         \
  } while (0)

#define TOAD_TRY(expr)                                             \
  do {                                                             \
    int _r = (expr);                                               \
    if (_r != 0) return _r;                                        \
  } while (0)

namespace toad {

constexpr int kWarp = 32;

__host__ __device__ inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }
__host__ __device__ inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// ---- split-precision helpers -------------------------------------------------
// v ~= hi + lo, hi = bf16_rn(v), lo = bf16_rn(v - hi).  (v - hi) is exact in fp32.
__device__ __forceinline__ void split2(float v0, float v1, uint32_t& hi, uint32_t& lo) {
  // packed cvt (v0 in the low half) and bit-level widening: 6 instructions per pair (the __nv_bfloat162
  // intrinsics unpack / re-pack the halves with two extra PRMTs per pair in the converter's inner loop)
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(v1), "f"(v0));
  const float h0 = __uint_as_float(hi << 16), h1 = __uint_as_float(hi & 0xFFFF0000u);
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(v1 - h1), "f"(v0 - h0));
}

__device__ __forceinline__ float bf16lo_to_f32(uint32_t packed) { return __uint_as_float(packed << 16); }
__device__ __forceinline__ float bf16hi_to_f32(uint32_t packed) { return __uint_as_float(packed & 0xFFFF0000u); }

// Accurate transcendental helpers for the gate epilogue (abs error ~1e-7).
__device__ __forceinline__ float sigmoid_acc(float z) { return 1.0f / (1.0f + expf(-z)); }
__device__ __forceinline__ float tanh_acc(float z) { return tanhf(z); }

// Counter-based dropout mask (training, nn.Dropout(0.25), reference models/model_toad.py:27-29,60-64):
// keep element `idx` of activation `layer` iff splitmix64(seed, layer, idx) >> 32 >= thresh, thresh = p * 2^32.
// Stateless, so forward and tests regenerate the same mask from (seed, layer, idx).
struct DropoutCfg {
  unsigned long long seed;
  uint32_t thresh;  // 0 = dropout off
  float scale;      // 1 / (1 - p)
};
__host__ __device__ __forceinline__ uint32_t dropout_hash(unsigned long long seed, uint32_t layer, unsigned long long idx) {
  unsigned long long z = seed + 0x9E3779B97F4A7C15ull * (idx + 1ull) + static_cast<unsigned long long>(layer) * 0xD1B54A32D192ED03ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z ^= z >> 31;
  return static_cast<uint32_t>(z >> 32);
}
__device__ __forceinline__ float dropout_apply(const DropoutCfg& d, uint32_t layer, unsigned long long idx, float v) {
  if (d.thresh == 0u) return v;
  return dropout_hash(d.seed, layer, idx) >= d.thresh ? v * d.scale : 0.f;
}
enum { DROP_H1 = 1, DROP_H = 2, DROP_A = 3, DROP_B = 4 };

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// 256-bit global store (sm_100: STG.256): one full 32 B sector per lane and instruction, for the epilogues whose lanes
// each own a contiguous run of a row (a register -> global store of v4 words fills only half a sector at a time).
__device__ __forceinline__ void st_global_v8(void* p, const uint32_t (&r)[32], int i0) {
  asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(r[i0]), "r"(r[i0 + 1]), "r"(r[i0 + 2]),
               "r"(r[i0 + 3]), "r"(r[i0 + 4]), "r"(r[i0 + 5]), "r"(r[i0 + 6]), "r"(r[i0 + 7])
               : "memory");
}

// 128-bit streaming load that does not pollute L1.
__device__ __forceinline__ float4 ld_stream_f4(const float* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ uint4 ld_stream_u4(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}

}  // namespace toad
