// Small kernels around the convolution GEMMs of the truncated ResNet-50 trunk
// (reference models/resnet_custom.py:19-109, eval mode): BatchNorm folding + weight re-layout,
// the 7x7/s2 stem's im2col, 3x3/s2 max-pool and the global average pool.  Activations live in
// HBM as NHWC (hi, lo) bf16 planes -- the operand format of the tcgen05 split-bf16 GEMM -- so a
// 1x1 convolution is a plain GEMM over the plane and a 3x3 one is 9 shifted TMA box loads.
#pragma once
#include "common.cuh"
#include <cuda_fp16.h>

namespace toad {
namespace resnet {

// w [Co, Ci, KH, KW] fp32 + BN(gamma, beta, running_mean, running_var) ->
//   planes [Co, Kpad] with K order (kh, kw, ci) (zero padded), bias[Co] = beta - mean * gamma / sqrt(var + eps)
// (resnet_custom.py:38-47: conv (bias-free) followed by BatchNorm2d in eval mode).
// HALF = false: (hi, lo) bf16 planes; HALF = true: (hi, lo) fp16 planes (hi = fp16(v), lo = fp16(v - hi): ~21 bits)
template <bool HALF>
__global__ void fold_conv_kernel(const float* __restrict__ w, const float* __restrict__ gamma,
                                 const float* __restrict__ beta, const float* __restrict__ mean,
                                 const float* __restrict__ var, float eps, __nv_bfloat16* __restrict__ hi,
                                 __nv_bfloat16* __restrict__ lo, float* __restrict__ bias, int Co, int Ci, int KH, int KW,
                                 int Kpad, int stem_layout) {
  const int co = blockIdx.x;
  const float scale = gamma[co] / sqrtf(var[co] + eps);
  if (threadIdx.x == 0) bias[co] = beta[co] - mean[co] * scale;
  const int K = KH * KW * Ci;
  for (int k = threadIdx.x; k < Kpad; k += blockDim.x) {
    float v = 0.f;
    if (stem_layout) {  // the fused stem's K order (stem.cuh): k = (ci*KH + kh)*8 + (kw + 1), slot 0 of each group zero
      const int g = k >> 3, e = k & 7;
      const int ci = g / KH, kh = g - ci * KH;
      if (g < Ci * KH && e >= 1 && e <= KW) v = w[((static_cast<int64_t>(co) * Ci + ci) * KH + kh) * KW + (e - 1)] * scale;
    } else if (k < K) {
      const int tap = k / Ci, ci = k % Ci;
      const int kh = tap / KW, kw = tap % KW;
      v = w[((static_cast<int64_t>(co) * Ci + ci) * KH + kh) * KW + kw] * scale;
    }
    if (HALF) {
      const __half h = __float2half_rn(v);
      reinterpret_cast<__half*>(hi)[static_cast<int64_t>(co) * Kpad + k] = h;
      reinterpret_cast<__half*>(lo)[static_cast<int64_t>(co) * Kpad + k] = __float2half_rn(v - __half2float(h));
    } else {
      const __nv_bfloat16 h = __float2bfloat16_rn(v);
      hi[static_cast<int64_t>(co) * Kpad + k] = h;
      lo[static_cast<int64_t>(co) * Kpad + k] = __float2bfloat16_rn(v - __bfloat162float(h));
    }
  }
}

// Stem im2col (conv1: 7x7, stride 2, pad 3, 3 input channels; resnet_custom.py:61,97).
// x NCHW fp32 [B,3,H,W] -> planes [B*Ho*Wo, 192], column k = (kh*7 + kw)*3 + c (147 real, rest 0).
// One thread produces 8 consecutive columns of one row (one 16 B store per plane).
__device__ __forceinline__ uint32_t pack_h2(float v0, float v1) {  // v0 in the low half; finite saturation
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(v1), "f"(v0));
  return r;
}
__device__ __forceinline__ float2 unpack_h2(uint32_t w) {
  const __half2 h = *reinterpret_cast<const __half2*>(&w);
  return __half22float2(h);
}
// 8 consecutive channel values of an activation held as (hi, lo) bf16 planes or as one fp16 plane
template <bool HALF>
__device__ __forceinline__ void load8(const __nv_bfloat16* hi, const __nv_bfloat16* lo, int64_t off, float (&v)[8]) {
  const uint4 vh = *reinterpret_cast<const uint4*>(hi + off);
  const uint32_t uh[4] = {vh.x, vh.y, vh.z, vh.w};
  if (HALF) {
#pragma unroll
    for (int e = 0; e < 4; ++e) { const float2 f = unpack_h2(uh[e]); v[2 * e] = f.x; v[2 * e + 1] = f.y; }
  } else {
    const uint4 vl = *reinterpret_cast<const uint4*>(lo + off);
    const uint32_t ul[4] = {vl.x, vl.y, vl.z, vl.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      v[2 * e] = bf16lo_to_f32(uh[e]) + bf16lo_to_f32(ul[e]);
      v[2 * e + 1] = bf16hi_to_f32(uh[e]) + bf16hi_to_f32(ul[e]);
    }
  }
}
template <bool HALF>
__device__ __forceinline__ void store8(__nv_bfloat16* hi, __nv_bfloat16* lo, int64_t off, const float (&v)[8]) {
  uint32_t h[4], l[4];
  if (HALF) {
#pragma unroll
    for (int e = 0; e < 4; ++e) h[e] = pack_h2(v[2 * e], v[2 * e + 1]);
    *reinterpret_cast<uint4*>(hi + off) = make_uint4(h[0], h[1], h[2], h[3]);
  } else {
#pragma unroll
    for (int e = 0; e < 4; ++e) split2(v[2 * e], v[2 * e + 1], h[e], l[e]);
    *reinterpret_cast<uint4*>(hi + off) = make_uint4(h[0], h[1], h[2], h[3]);
    *reinterpret_cast<uint4*>(lo + off) = make_uint4(l[0], l[1], l[2], l[3]);
  }
}

constexpr int STEM_K = 147, STEM_KPAD = 192;
constexpr int STEM_THREADS = 192;  // 24 column groups x 8 output pixels per pass
// One CTA per output row (b, oh): the 7 input rows x 3 channels it touches are staged in shared memory with
// coalesced 128-bit loads (zero padded: 3 columns each side, rows outside the image), then thread (kg, ow) gathers
// its 8 columns through 8 per-thread constant offsets -- no div/mod and no scattered global loads in the loop.
template <bool HALF>
__global__ void __launch_bounds__(STEM_THREADS) stem_im2col_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ hi,
                                                                   __nv_bfloat16* __restrict__ lo, int B, int H, int W, int Ho,
                                                                   int Wo) {
  extern __shared__ float s_in[];  // [3][7][W + 8]: element (c, kh, iw + 4)
  const int PW = W + 8;
  const int b = blockIdx.x / Ho, oh = blockIdx.x % Ho;
  const int tid = threadIdx.x;
  for (int i = tid; i < 21 * 2; i += STEM_THREADS) {  // the two 4-wide pad strips of every staged row
    const int r = i >> 1, side = i & 1;
    *reinterpret_cast<float4*>(s_in + r * PW + (side ? W + 4 : 0)) = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  const int W4 = W / 4;
  for (int i = tid; i < 21 * W4; i += STEM_THREADS) {
    const int r = i / W4, q = i - r * W4;  // r = c * 7 + kh
    const int c = r / 7, kh = r - c * 7;
    const int ih = oh * 2 + kh - 3;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (ih >= 0 && ih < H) v = ld_stream_f4(x + ((static_cast<int64_t>(b) * 3 + c) * H + ih) * W + q * 4);
    *reinterpret_cast<float4*>(s_in + r * PW + 4 + q * 4) = v;
  }
  const int kg = tid % 24, ow0 = tid / 24;
  int off[8];  // smem offset of column k = kg*8 + e for ow = 0 (-1: zero padding column)
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const int k = kg * 8 + e;
    const int c = k % 3, tap = k / 3;
    const int kh = tap / 7, kw = tap % 7;
    off[e] = k < STEM_K ? (c * 7 + kh) * PW + 4 + kw - 3 : -1;
  }
  __syncthreads();
  const int64_t row0 = (static_cast<int64_t>(b) * Ho + oh) * Wo;
  for (int ow = ow0; ow < Wo; ow += STEM_THREADS / 24) {
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = off[e] >= 0 ? s_in[off[e] + 2 * ow] : 0.f;
    store8<HALF>(hi, lo, (row0 + ow) * STEM_KPAD + kg * 8, v);
  }
}
template <bool HALF>
inline int launch_stem_im2col(const float* x, __nv_bfloat16* hi, __nv_bfloat16* lo, int B, int H, int W, int Ho, int Wo,
                              cudaStream_t stream) {
  if (W % 4 != 0) return TOAD_ERR_UNSUPPORTED;
  const int smem = 21 * (W + 8) * static_cast<int>(sizeof(float));
  if (smem > 48 * 1024) TOAD_CUDA_TRY(cudaFuncSetAttribute(stem_im2col_kernel<HALF>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  stem_im2col_kernel<HALF><<<static_cast<unsigned>(B) * Ho, STEM_THREADS, smem, stream>>>(x, hi, lo, B, H, W, Ho, Wo);
  TOAD_CUDA_TRY(cudaGetLastError());
  return 0;
}

// MaxPool2d(3, stride 2, pad 1) on NHWC planes (resnet_custom.py:64,100).  Thread = (output pixel, 8 channels).
template <bool HALF>
__global__ void maxpool3x3s2_kernel(const __nv_bfloat16* __restrict__ in_hi, const __nv_bfloat16* __restrict__ in_lo,
                                    __nv_bfloat16* __restrict__ out_hi, __nv_bfloat16* __restrict__ out_lo, int B, int H,
                                    int W, int C) {
  const int Ho = H / 2, Wo = W / 2, CG8 = C / 8;
  const int64_t gid = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t total = static_cast<int64_t>(B) * Ho * Wo * CG8;
  if (gid >= total) return;
  const int cg = static_cast<int>(gid % CG8);
  const int64_t pix = gid / CG8;
  const int ow = static_cast<int>(pix % Wo), oh = static_cast<int>((pix / Wo) % Ho);
  const int b = static_cast<int>(pix / (static_cast<int64_t>(Wo) * Ho));
  float m[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) m[e] = -INFINITY;
  for (int kh = 0; kh < 3; ++kh) {
    const int ih = oh * 2 + kh - 1;
    if (ih < 0 || ih >= H) continue;
    for (int kw = 0; kw < 3; ++kw) {
      const int iw = ow * 2 + kw - 1;
      if (iw < 0 || iw >= W) continue;
      const int64_t off = ((static_cast<int64_t>(b) * H + ih) * W + iw) * C + cg * 8;
      float v[8];
      load8<HALF>(in_hi, in_lo, off, v);
#pragma unroll
      for (int e = 0; e < 8; ++e) m[e] = fmaxf(m[e], v[e]);
    }
  }
  store8<HALF>(out_hi, out_lo, pix * C + cg * 8, m);  // exact: m is one of the stored inputs
}

// AdaptiveAvgPool2d(1) + view (resnet_custom.py:106-107): planes [B, HW, C] -> fp32 [B, C].
// Block = (image, 256-channel slab); 8 pixel-groups x 32 lanes x 8 channels, fixed-order smem reduce.
template <bool HALF>
__global__ void avgpool_kernel(const __nv_bfloat16* __restrict__ in_hi, const __nv_bfloat16* __restrict__ in_lo,
                               float* __restrict__ out, int HW, int C) {
  __shared__ float red[8][256];
  const int b = blockIdx.x, c0 = blockIdx.y * 256;
  const int lane = threadIdx.x & 31, grp = threadIdx.x >> 5;  // 256 threads
  float acc[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) acc[e] = 0.f;
  for (int p = grp; p < HW; p += 8) {
    const int64_t off = (static_cast<int64_t>(b) * HW + p) * C + c0 + lane * 8;
    float v[8];
    load8<HALF>(in_hi, in_lo, off, v);
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] += v[e];
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) red[grp][lane * 8 + e] = acc[e];
  __syncthreads();
  const int c = threadIdx.x;
  float s = 0.f;
#pragma unroll
  for (int g = 0; g < 8; ++g) s += red[g][c];
  out[static_cast<int64_t>(b) * C + c0 + c] = s / static_cast<float>(HW);
}

}  // namespace resnet
}  // namespace toad
