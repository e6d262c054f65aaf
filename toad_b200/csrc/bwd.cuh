// Backward of the TOAD forward: gradients of the 14 parameter tensors given the upstream
// gradients of `logits` and `site_logits` (what autograd computes for
// utils/core_utils_mtl_concat.py:231 on the graph of models/model_toad.py:90-107).
// No gradient flows to the features x (leaf data, core_utils_mtl_concat.py:201).
//
// Small HBM-bound kernels live here; the five large contractions (2 dgrad, 3 wgrad) run on
// the fp32 CUDA-core GEMM in sgemm_simt.cuh.  All cross-CTA sums are two-stage with a fixed
// reduction order (deterministic).
#pragma once
#include "common.cuh"

namespace toad {
namespace bwd {

constexpr int H = 512;
constexpr int T = 2;

// ---- heads: dW/db of classifier & site_classifier, dM[t] = dlogits_t . W_t[:, :H], sdot[t] = dM[t].M[t]
__global__ void heads_bwd_kernel(const float* __restrict__ dlogits, const float* __restrict__ dsite,
                                 const float* __restrict__ features, const float* __restrict__ wcls,
                                 const float* __restrict__ wsite, int n_classes, float* __restrict__ g_wcls,
                                 float* __restrict__ g_bcls, float* __restrict__ g_wsite, float* __restrict__ g_bsite,
                                 float* __restrict__ dM, float* __restrict__ sdot) {
  __shared__ float s_part[2][32];
  const int tid = threadIdx.x;  // 1024 threads
  for (int i = tid; i < n_classes * (H + 1); i += blockDim.x)
    g_wcls[i] = __ldg(dlogits + i / (H + 1)) * __ldg(features + i % (H + 1));
  for (int i = tid; i < 2 * (H + 1); i += blockDim.x)
    g_wsite[i] = __ldg(dsite + i / (H + 1)) * __ldg(features + (H + 1) + i % (H + 1));
  if (tid < n_classes) g_bcls[tid] = dlogits[tid];
  if (tid < 2) g_bsite[tid] = dsite[tid];
  // dM: thread j < H -> task 0, H <= tid < 2H -> task 1
  const int t = tid / H, j = tid % H;
  float d = 0.f;
  if (t == 0) {
    for (int c = 0; c < n_classes; ++c) d = fmaf(__ldg(dlogits + c), __ldg(wcls + c * (H + 1) + j), d);
  } else {
    for (int c = 0; c < 2; ++c) d = fmaf(__ldg(dsite + c), __ldg(wsite + c * (H + 1) + j), d);
  }
  dM[tid] = d;
  float prod = d * __ldg(features + t * (H + 1) + j);
  prod = warp_sum(prod);
  if ((tid & 31) == 0) s_part[t][(tid % H) >> 5] = prod;
  __syncthreads();
  if (tid < 2) {
    float s = 0.f;
    for (int w = 0; w < H / 32; ++w) s += s_part[tid][w];
    sdot[tid] = s;
  }
}

// ---- softmax-pooling backward per row: P[n,t], dA[n,t] = P (dM_t.h_n - sdot_t)   (warp per row)
// PLANES: h is given as its saved (hi, lo) bf16 planes (tensor-core path) instead of fp32.
template <bool PLANES>
__global__ void pool_bwd_kernel(const float* __restrict__ h, const __nv_bfloat16* __restrict__ h_hi,
                                const __nv_bfloat16* __restrict__ h_lo, const float* __restrict__ a_raw,
                                const float* __restrict__ stats, const float* __restrict__ dM,
                                const float* __restrict__ sdot, float* __restrict__ P, float* __restrict__ dA,
                                int64_t N) {
  __shared__ float s_dM[T * H];
  for (int i = threadIdx.x; i < T * H; i += blockDim.x) s_dM[i] = dM[i];
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  for (int64_t row = static_cast<int64_t>(blockIdx.x) * wpb + (threadIdx.x >> 5); row < N;
       row += static_cast<int64_t>(gridDim.x) * wpb) {
    float d0 = 0.f, d1 = 0.f;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int c = 4 * (lane + 32 * q);
      float4 v;
      if (PLANES) {
        const uint2 uh = *reinterpret_cast<const uint2*>(h_hi + row * H + c);
        const uint2 ul = *reinterpret_cast<const uint2*>(h_lo + row * H + c);
        v = make_float4(bf16lo_to_f32(uh.x) + bf16lo_to_f32(ul.x), bf16hi_to_f32(uh.x) + bf16hi_to_f32(ul.x),
                        bf16lo_to_f32(uh.y) + bf16lo_to_f32(ul.y), bf16hi_to_f32(uh.y) + bf16hi_to_f32(ul.y));
      } else {
        v = ld_stream_f4(h + row * H + c);
      }
      d0 += v.x * s_dM[c] + v.y * s_dM[c + 1] + v.z * s_dM[c + 2] + v.w * s_dM[c + 3];
      d1 += v.x * s_dM[H + c] + v.y * s_dM[H + c + 1] + v.z * s_dM[H + c + 2] + v.w * s_dM[H + c + 3];
    }
    d0 = warp_sum(d0);
    d1 = warp_sum(d1);
    if (lane < T) {
      const float d = lane == 0 ? d0 : d1;
      const float pr = expf(a_raw[static_cast<int64_t>(lane) * N + row] - stats[2 * lane]) / stats[2 * lane + 1];
      P[row * T + lane] = pr;
      dA[row * T + lane] = pr * (d - sdot[lane]);
    }
  }
}

// ---- gate backward: dab[n] = [da_pre | db_pre]; per-CTA partials of dWc, dba, dbb, dbc
// block = D threads (thread j owns gate column j); partial layout per CTA: [dWc0[D] dWc1[D] dba[D] dbb[D] dbc[2]]
// PLANES: dab goes out as (hi, lo) bf16 planes -- the operand format of the tensor-core dgrad / wgrad -- instead of fp32.
// Thread (rg, cq) = (tid / (D/4), tid % (D/4)) owns gate columns 4cq..4cq+3 of rows r0 + rg, r0 + rg + 4, ...: 128-bit
// loads of a and b, one 8-byte store per plane and branch; the 4 row groups are added in order through shared memory.
template <bool PLANES>
__global__ void gate_bwd_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                const float* __restrict__ dA, const float* __restrict__ wc, float* __restrict__ dab,
                                __nv_bfloat16* __restrict__ dab_hi, __nv_bfloat16* __restrict__ dab_lo,
                                float* __restrict__ part, int64_t N, int D, int rows_per_block, float keep) {
  extern __shared__ float gb_sm[];  // [4 row groups][4 * D + 2]
  // a, b are the saved POST-dropout activations (a_post = a*mask/keep); keep == 1 without dropout.
  const float inv_keep = 1.f / keep;
  const int quads = D / 4;
  const int rg = threadIdx.x / quads, cq = threadIdx.x - rg * quads, j0 = 4 * cq;
  const int64_t r0 = static_cast<int64_t>(blockIdx.x) * rows_per_block;
  int64_t r1 = r0 + rows_per_block;
  if (r1 > N) r1 = N;
  const float4 w0 = __ldg(reinterpret_cast<const float4*>(wc + j0)), w1 = __ldg(reinterpret_cast<const float4*>(wc + D + j0));
  const float w0v[4] = {w0.x, w0.y, w0.z, w0.w}, w1v[4] = {w1.x, w1.y, w1.z, w1.w};
  float gw0[4] = {0.f, 0.f, 0.f, 0.f}, gw1[4] = {0.f, 0.f, 0.f, 0.f}, gba[4] = {0.f, 0.f, 0.f, 0.f}, gbb[4] = {0.f, 0.f, 0.f, 0.f};
  float gc0 = 0.f, gc1 = 0.f;
#pragma unroll 2
  for (int64_t row = r0 + rg; row < r1; row += 4) {
    const float2 da = __ldg(reinterpret_cast<const float2*>(dA + row * T));
    const float4 a4 = ld_stream_f4(a + row * D + j0), b4 = ld_stream_f4(b + row * D + j0);
    const float av4[4] = {a4.x, a4.y, a4.z, a4.w}, bv4[4] = {b4.x, b4.y, b4.z, b4.w};
    float dap[4], dbp[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float av = av4[e], bv = bv4[e];
      const float g = av * bv;
      gw0[e] = fmaf(da.x, g, gw0[e]);
      gw1[e] = fmaf(da.y, g, gw1[e]);
      const float dg = da.x * w0v[e] + da.y * w1v[e];
      const float a_pre = av * keep;  // tanh output where kept (0 where dropped: no gradient there)
      dap[e] = (keep < 1.f && av == 0.f) ? 0.f : dg * bv * (1.f - a_pre * a_pre) * inv_keep;
      dbp[e] = dg * av * bv * (1.f - bv * keep);
      gba[e] += dap[e];
      gbb[e] += dbp[e];
    }
    gc0 += da.x;
    gc1 += da.y;
    if (PLANES) {
      uint32_t h[2], l[2];
      split2(dap[0], dap[1], h[0], l[0]);
      split2(dap[2], dap[3], h[1], l[1]);
      *reinterpret_cast<uint2*>(dab_hi + row * 2 * D + j0) = make_uint2(h[0], h[1]);
      *reinterpret_cast<uint2*>(dab_lo + row * 2 * D + j0) = make_uint2(l[0], l[1]);
      split2(dbp[0], dbp[1], h[0], l[0]);
      split2(dbp[2], dbp[3], h[1], l[1]);
      *reinterpret_cast<uint2*>(dab_hi + row * 2 * D + D + j0) = make_uint2(h[0], h[1]);
      *reinterpret_cast<uint2*>(dab_lo + row * 2 * D + D + j0) = make_uint2(l[0], l[1]);
    } else {
      *reinterpret_cast<float4*>(dab + row * 2 * D + j0) = make_float4(dap[0], dap[1], dap[2], dap[3]);
      *reinterpret_cast<float4*>(dab + row * 2 * D + D + j0) = make_float4(dbp[0], dbp[1], dbp[2], dbp[3]);
    }
  }
  float* sm = gb_sm + rg * (4 * D + 2);
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    sm[j0 + e] = gw0[e];
    sm[D + j0 + e] = gw1[e];
    sm[2 * D + j0 + e] = gba[e];
    sm[3 * D + j0 + e] = gbb[e];
  }
  if (cq == 0) { sm[4 * D] = gc0; sm[4 * D + 1] = gc1; }  // (every thread of a row group sees the same rows)
  __syncthreads();
  float* mine = part + static_cast<int64_t>(blockIdx.x) * (4 * D + 2);
  for (int i = threadIdx.x; i < 4 * D + 2; i += blockDim.x) {
    float t = 0.f;
#pragma unroll
    for (int q = 0; q < 4; ++q) t += gb_sm[q * (4 * D + 2) + i];
    mine[i] = t;
  }
}
inline size_t gate_bwd_smem(int D) { return static_cast<size_t>(4) * (4 * D + 2) * sizeof(float); }

// ---- column sums of a [N, C] matrix: per-CTA partials [gridDim.x][C] (thread per column)
__global__ void colsum_kernel(const float* __restrict__ m, float* __restrict__ part, int64_t N, int C,
                              int rows_per_block) {
  const int c = threadIdx.x;
  const int64_t r0 = static_cast<int64_t>(blockIdx.x) * rows_per_block;
  int64_t r1 = r0 + rows_per_block;
  if (r1 > N) r1 = N;
  float s = 0.f;
  for (int64_t row = r0; row < r1; ++row) s += m[row * C + c];
  part[static_cast<int64_t>(blockIdx.x) * C + c] = s;
}

// out[i] = sum_z part[z*stride + i], fixed order
__global__ void reduce_strided_kernel(const float* __restrict__ part, float* __restrict__ out, int64_t n,
                                      int64_t stride, int splits) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float s = 0.f;
  for (int z = 0; z < splits; ++z) s += part[static_cast<int64_t>(z) * stride + i];
  out[i] = s;
}

// ---- transposes feeding the tensor-core wgrad GEMMs (dW = dY^T . X needs both operands K-major in the
// patch index): out planes [C, ldT] (hi, lo) with out[c][r] = in[r][c]; 32 x 32 tiles through smem.
__global__ void transpose_split_kernel(const float* __restrict__ in, int64_t R, int C, int64_t ld_in,
                                       __nv_bfloat16* __restrict__ out_hi, __nv_bfloat16* __restrict__ out_lo,
                                       int64_t ldT) {
  __shared__ float tile[32][33];
  const int64_t r0 = static_cast<int64_t>(blockIdx.x) * 32;
  const int c0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int64_t r = r0 + i;
    const int c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (r < R && c < C) ? in[r * ld_in + c] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int c = c0 + i;
    const int64_t r = r0 + threadIdx.x;
    if (c < C && r < R) {
      const float v = tile[threadIdx.x][i];
      const __nv_bfloat16 h = __float2bfloat16_rn(v);
      out_hi[static_cast<int64_t>(c) * ldT + r] = h;
      out_lo[static_cast<int64_t>(c) * ldT + r] = __float2bfloat16_rn(v - __bfloat162float(h));
    }
  }
}

inline int launch_transpose_split(const float* in, int64_t R, int C, int64_t ld_in, __nv_bfloat16* out_hi,
                                  __nv_bfloat16* out_lo, int64_t ldT, cudaStream_t stream) {
  if (R <= 0 || C <= 0) return 0;
  dim3 grid(static_cast<unsigned>((R + 31) / 32), static_cast<unsigned>((C + 31) / 32));
  transpose_split_kernel<<<grid, dim3(32, 8), 0, stream>>>(in, R, C, ld_in, out_hi, out_lo, ldT);
  TOAD_CUDA_TRY(cudaGetLastError());
  return 0;
}

// (hi, lo) planes [R, C] -> transposed planes [C, ldT]
__global__ void transpose_planes_kernel(const __nv_bfloat16* __restrict__ in_hi, const __nv_bfloat16* __restrict__ in_lo,
                                        int64_t R, int C, __nv_bfloat16* __restrict__ out_hi,
                                        __nv_bfloat16* __restrict__ out_lo, int64_t ldT) {
  __shared__ __nv_bfloat16 th[32][34], tl[32][34];
  const int64_t r0 = static_cast<int64_t>(blockIdx.x) * 32;
  const int c0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int64_t r = r0 + i;
    const int c = c0 + threadIdx.x;
    const bool ok = r < R && c < C;
    th[i][threadIdx.x] = ok ? in_hi[r * C + c] : __float2bfloat16_rn(0.f);
    tl[i][threadIdx.x] = ok ? in_lo[r * C + c] : __float2bfloat16_rn(0.f);
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int c = c0 + i;
    const int64_t r = r0 + threadIdx.x;
    if (c < C && r < R) {
      out_hi[static_cast<int64_t>(c) * ldT + r] = th[threadIdx.x][i];
      out_lo[static_cast<int64_t>(c) * ldT + r] = tl[threadIdx.x][i];
    }
  }
}

inline int launch_transpose_planes(const __nv_bfloat16* in_hi, const __nv_bfloat16* in_lo, int64_t R, int C,
                                   __nv_bfloat16* out_hi, __nv_bfloat16* out_lo, int64_t ldT, cudaStream_t stream) {
  if (R <= 0 || C <= 0) return 0;
  dim3 grid(static_cast<unsigned>((R + 31) / 32), static_cast<unsigned>((C + 31) / 32));
  transpose_planes_kernel<<<grid, dim3(32, 8), 0, stream>>>(in_hi, in_lo, R, C, out_hi, out_lo, ldT);
  TOAD_CUDA_TRY(cudaGetLastError());
  return 0;
}

// column sums of (hi + lo) planes [N, C] (C % 8 == 0): per-CTA partials [gridDim.x][C].  Thread (x, y) owns columns
// 8x..8x+7 (one 128-bit load per plane) of rows r0 + y, r0 + y + blockDim.y, ...; the y lanes are added in order.
__global__ void colsum_planes_kernel(const __nv_bfloat16* __restrict__ hi, const __nv_bfloat16* __restrict__ lo,
                                     float* __restrict__ part, int64_t N, int C, int rows_per_block) {
  extern __shared__ float cs_sm[];  // [blockDim.y][C]
  const int c8 = threadIdx.x;
  const int64_t r0 = static_cast<int64_t>(blockIdx.x) * rows_per_block;
  int64_t r1 = r0 + rows_per_block;
  if (r1 > N) r1 = N;
  float s[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll 2
  for (int64_t row = r0 + threadIdx.y; row < r1; row += blockDim.y) {
    const uint4 vh = ld_stream_u4(hi + row * C + c8 * 8), vl = ld_stream_u4(lo + row * C + c8 * 8);
    const uint32_t uh[4] = {vh.x, vh.y, vh.z, vh.w}, ul[4] = {vl.x, vl.y, vl.z, vl.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      s[2 * e] += bf16lo_to_f32(uh[e]) + bf16lo_to_f32(ul[e]);
      s[2 * e + 1] += bf16hi_to_f32(uh[e]) + bf16hi_to_f32(ul[e]);
    }
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) cs_sm[threadIdx.y * C + c8 * 8 + e] = s[e];
  __syncthreads();
  const int tid = threadIdx.y * blockDim.x + threadIdx.x;
  float* mine = part + static_cast<int64_t>(blockIdx.x) * C;
  for (int c = tid; c < C; c += blockDim.x * blockDim.y) {
    float t = 0.f;
    for (int y = 0; y < static_cast<int>(blockDim.y); ++y) t += cs_sm[y * C + c];
    mine[c] = t;
  }
}

inline int launch_colsum_planes(const __nv_bfloat16* hi, const __nv_bfloat16* lo, float* part, int64_t N, int C,
                                int blocks, cudaStream_t stream) {
  if (C % 8 != 0 || C / 8 > 1024) return TOAD_ERR_UNSUPPORTED;
  int ry = 512 / (C / 8);
  if (ry > 16) ry = 16;
  if (ry < 1) ry = 1;
  const int rpb = static_cast<int>((N + blocks - 1) / blocks);
  colsum_planes_kernel<<<blocks, dim3(C / 8, ry), static_cast<size_t>(ry) * C * sizeof(float), stream>>>(
      hi, lo, part, N, C, rpb);
  TOAD_CUDA_TRY(cudaGetLastError());
  return 0;
}

// Same sum for many splits over few columns (per-CTA partials of bias / score-weight gradients): a 32-column x 32-lane
// CTA, lane y adds splits y, y+32, ... and the 32 lane sums are added in lane order -- fixed order, so deterministic.
// Up to four column segments [seg_begin[i], seg_begin[i+1]) go to separate destinations in one launch.
struct ReduceSegs {
  float* dst[4];
  int begin[5];
};
__global__ void reduce_many_splits_kernel(const float* __restrict__ part, ReduceSegs segs, int n, int64_t stride,
                                          int splits) {
  __shared__ float sm[32][33];
  const int col = blockIdx.x * 32 + threadIdx.x;
  float s = 0.f;
  if (col < n)
    for (int z = threadIdx.y; z < splits; z += 32) s += part[static_cast<int64_t>(z) * stride + col];
  sm[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && col < n) {
    float t = 0.f;
#pragma unroll
    for (int y = 0; y < 32; ++y) t += sm[y][threadIdx.x];
    int sgi = 0;
#pragma unroll
    for (int i = 1; i < 4; ++i)
      if (col >= segs.begin[i]) sgi = i;
    segs.dst[sgi][col - segs.begin[sgi]] = t;
  }
}

inline int launch_reduce_segs(const float* part, const ReduceSegs& segs, int n, int64_t stride, int splits,
                              cudaStream_t stream) {
  if (n <= 0) return 0;
  reduce_many_splits_kernel<<<static_cast<unsigned>((n + 31) / 32), dim3(32, 32), 0, stream>>>(part, segs, n, stride, splits);
  TOAD_CUDA_TRY(cudaGetLastError());
  return 0;
}

inline int launch_reduce_strided(const float* part, float* out, int64_t n, int64_t stride, int splits,
                                 cudaStream_t stream) {
  if (n <= 0) return 0;
  if (splits > 16 && n <= 4096) {
    ReduceSegs segs{};
    const int big = 1 << 30;
    segs.dst[0] = out; segs.begin[0] = 0; segs.begin[1] = big; segs.begin[2] = big; segs.begin[3] = big; segs.begin[4] = big;
    return launch_reduce_segs(part, segs, static_cast<int>(n), stride, splits, stream);
  }
  reduce_strided_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, stream>>>(part, out, n, stride, splits);
  TOAD_CUDA_TRY(cudaGetLastError());
  return 0;
}

}  // namespace bwd
}  // namespace toad
