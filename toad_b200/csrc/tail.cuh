// HBM-bound tail of the TOAD forward (reference models/model_toad.py:92-107):
//   A_raw = scores^T ; P = softmax(A_raw, dim=1) over all N patches, per task ;
//   M = P . h ; M = cat(M, sex) ; the two linear heads ; top-1 ; softmaxes.
// One launch: every CTA streams a contiguous chunk of rows of h exactly once (128-bit
// coalesced loads, warp-per-row, register accumulators), keeps (max, sum-exp, weighted sum)
// partials, and the last CTA to finish (atomic ticket) merges the partials in fixed order
// and evaluates the heads -- deterministic, no float atomics.
//
// Also here: the small weight-preparation kernels of the tensor-core path and the
// attention_c contraction of the fp32 path.
#pragma once
#include "common.cuh"

namespace toad {
namespace tail {

constexpr int H = 512;          // hid_dim (model_toad.py:56, both size_args)
constexpr int T = 2;            // n_tasks (model_toad.py:66)
constexpr int THREADS = 1024;   // 32 warps, one CTA per SM
constexpr int WARPS = THREADS / 32;
constexpr int MAX_CHUNK = 2048; // rows per CTA (scores cached in smem)
constexpr int MAX_BLOCKS = 1024;   // (max,sum) staging of the merge fits s_score: 2*T*MAX_BLOCKS floats
constexpr int PART_STRIDE = T * (H + 2);  // per-CTA partial: acc[T][H], then (m,l)[T]

enum { H_F32 = 0, H_SPLIT = 1 };

struct TailParams {
  const float* part;   // [n_parts][N][T] score partials (no bias)
  int32_t n_parts;
  const float* bc;     // [T]
  float* a_raw;        // [T][N]
  const float* h_f32;  // [N, H]          (H_F32)
  const __nv_bfloat16* h_hi;  // [N, H]   (H_SPLIT)
  const __nv_bfloat16* h_lo;
  int64_t N;
  int32_t rows_per_block;
  const float* sex;    // [1]
  const float* wcls; const float* bcls; int32_t n_classes;
  const float* wsite; const float* bsite;
  float* features; float* logits; float* y_prob; int64_t* y_hat;
  float* site_logits; float* site_prob; int64_t* site_hat; float* stats;
  float* blk_part;     // [gridDim.x][PART_STRIDE]
  unsigned int* ticket;
  int32_t attention_only;
};

inline int tail_blocks(int64_t n, int sms) {
  int64_t b = static_cast<int64_t>(sms);  // one 1024-thread CTA per SM: fewer partials for the final merge
  const int64_t need = (n + MAX_CHUNK - 1) / MAX_CHUNK;
  if (b < need) b = need;
  const int64_t most = (n + 31) / 32;  // at least 32 rows per CTA
  if (b > most) b = most;
  if (b < 1) b = 1;
  return static_cast<int>(b);
}

template <int H_MODE>
__global__ void __launch_bounds__(THREADS) pool_heads_kernel(const TailParams p) {
  extern __shared__ float dsm[];               // [WARPS][T*H] cross-warp reduction buffer
  __shared__ float s_score[T][MAX_CHUNK];      // this CTA's scores, then reused by the merge
  __shared__ float s_red[T][WARPS];
  __shared__ float s_m[T], s_l[T];
  __shared__ unsigned int s_ticket;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t r0 = static_cast<int64_t>(blockIdx.x) * p.rows_per_block;
  int64_t r1 = r0 + p.rows_per_block;
  if (r1 > p.N) r1 = p.N;
  const int rows = r1 > r0 ? static_cast<int>(r1 - r0) : 0;

  // ---- phase 1: finish the scores (sum split-N partials + bias), write A_raw, local max
  float lmax[T] = {-INFINITY, -INFINITY};
  for (int i = tid; i < rows; i += THREADS) {
    const int64_t row = r0 + i;
#pragma unroll
    for (int t = 0; t < T; ++t) {
      float s = 0.f;
      for (int z = 0; z < p.n_parts; ++z) s += __ldg(p.part + (static_cast<int64_t>(z) * p.N + row) * T + t);
      s += __ldg(p.bc + t);
      p.a_raw[static_cast<int64_t>(t) * p.N + row] = s;
      s_score[t][i] = s;
      lmax[t] = fmaxf(lmax[t], s);
    }
  }
  if (p.attention_only) return;
#pragma unroll
  for (int t = 0; t < T; ++t) {
    const float m = warp_max(lmax[t]);
    if (lane == 0) s_red[t][warp] = m;
  }
  __syncthreads();
  if (tid < T) {
    float m = -INFINITY;
    for (int w = 0; w < WARPS; ++w) m = fmaxf(m, s_red[tid][w]);
    s_m[tid] = m;
  }
  __syncthreads();
  const float m0 = s_m[0], m1 = s_m[1];

  // ---- phase 2: stream h once; warp per row, 16 columns per lane
  float acc0[16], acc1[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) { acc0[i] = 0.f; acc1[i] = 0.f; }
  float l0 = 0.f, l1 = 0.f;
  for (int i = warp; i < rows; i += WARPS) {
    const int64_t row = r0 + i;
    const float p0 = expf(s_score[0][i] - m0);
    const float p1 = expf(s_score[1][i] - m1);
    l0 += p0;
    l1 += p1;
    float hv[16];
    if (H_MODE == H_F32) {
      const float* hr = p.h_f32 + row * H;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float4 v = ld_stream_f4(hr + 4 * (lane + 32 * q));
        hv[4 * q] = v.x; hv[4 * q + 1] = v.y; hv[4 * q + 2] = v.z; hv[4 * q + 3] = v.w;
      }
    } else {
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const uint4 vh = ld_stream_u4(p.h_hi + row * H + q * 256 + lane * 8);
        const uint4 vl = ld_stream_u4(p.h_lo + row * H + q * 256 + lane * 8);
        const uint32_t uh[4] = {vh.x, vh.y, vh.z, vh.w}, ul[4] = {vl.x, vl.y, vl.z, vl.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          hv[8 * q + 2 * e] = bf16lo_to_f32(uh[e]) + bf16lo_to_f32(ul[e]);
          hv[8 * q + 2 * e + 1] = bf16hi_to_f32(uh[e]) + bf16hi_to_f32(ul[e]);
        }
      }
    }
#pragma unroll
    for (int e = 0; e < 16; ++e) {
      acc0[e] = fmaf(p0, hv[e], acc0[e]);
      acc1[e] = fmaf(p1, hv[e], acc1[e]);
    }
  }
  // column owned by accumulator e of this lane
  auto col_of = [&](int e) -> int {
    return H_MODE == H_F32 ? 4 * (lane + 32 * (e >> 2)) + (e & 3) : (e >> 3) * 256 + lane * 8 + (e & 7);
  };
#pragma unroll
  for (int e = 0; e < 16; ++e) {
    dsm[warp * (T * H) + col_of(e)] = acc0[e];
    dsm[warp * (T * H) + H + col_of(e)] = acc1[e];
  }
  if (lane == 0) { s_red[0][warp] = l0; s_red[1][warp] = l1; }
  __syncthreads();
  float* mine = p.blk_part + static_cast<int64_t>(blockIdx.x) * PART_STRIDE;
  for (int c = tid; c < T * H; c += THREADS) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < WARPS; ++w) s += dsm[w * (T * H) + c];
    mine[c] = s;
  }
  if (tid < T) {
    float l = 0.f;
    for (int w = 0; w < WARPS; ++w) l += s_red[tid][w];
    mine[T * H + 2 * tid] = s_m[tid];
    mine[T * H + 2 * tid + 1] = l;
  }

  // ---- phase 3: last CTA merges all partials in fixed order and evaluates the heads
  __threadfence();
  __syncthreads();
  if (tid == 0) s_ticket = atomicAdd(p.ticket, 1u);
  __syncthreads();
  if (s_ticket != gridDim.x - 1) return;
  __threadfence();
  const int nb = gridDim.x;
  // stage every CTA's (max, sum) pair in smem with parallel L2 loads, then two threads fold them
  // serially from smem (fixed order, ~3k cycles) instead of chasing ~300 dependent L2 latencies.
  float* s_scale = &s_score[0][0];  // [nb][T] max -> scale;  [nb][T] sums behind it (nb <= MAX_BLOCKS)
  float* s_lb = s_scale + nb * T;
  for (int i = tid; i < nb * T; i += THREADS) {
    const int b = i / T, t = i % T;
    s_scale[i] = __ldcg(p.blk_part + static_cast<int64_t>(b) * PART_STRIDE + T * H + 2 * t);
    s_lb[i] = __ldcg(p.blk_part + static_cast<int64_t>(b) * PART_STRIDE + T * H + 2 * t + 1);
  }
  __syncthreads();
  if (tid < T) {
    float m = -INFINITY;
    for (int b = 0; b < nb; ++b) m = fmaxf(m, s_scale[b * T + tid]);
    s_m[tid] = m;
  }
  __syncthreads();
  for (int i = tid; i < nb * T; i += THREADS) s_scale[i] = expf(s_scale[i] - s_m[i % T]);  // exp(-inf) = 0: empty CTA
  __syncthreads();
  if (tid < T) {
    float l = 0.f;
    for (int b = 0; b < nb; ++b) l = fmaf(s_scale[b * T + tid], s_lb[b * T + tid], l);
    s_l[tid] = l;
    p.stats[2 * tid] = s_m[tid];
    p.stats[2 * tid + 1] = l;
  }
  __syncthreads();
  // weighted sum of the per-CTA accumulators: 2 thread groups x 256 float4 columns, 8 independent
  // 16 B L2 loads in flight per thread, partial sums combined in fixed order through smem.
  float* s_feat = dsm;                 // [T][H+1]
  float4* s_half = reinterpret_cast<float4*>(dsm + 4096);  // [4][256] float4 scratch (dsm is 128 KB)
  const float sexv = __ldg(p.sex);
  {
    const int c4 = tid & 255, bg = tid >> 8;  // 4 partial-groups x 256 float4 columns
    const int t = c4 >> 7;  // 128 float4 per task
    const int bper = (nb + 3) / 4;
    const int b0 = bg * bper, b1 = (b0 + bper) < nb ? (b0 + bper) : nb;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    int b = b0;
    for (; b + 8 <= b1; b += 8) {
      float4 v8[8];
#pragma unroll
      for (int u = 0; u < 8; ++u)
        v8[u] = __ldcg(reinterpret_cast<const float4*>(p.blk_part + static_cast<int64_t>(b + u) * PART_STRIDE) + c4);
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const float sc = s_scale[(b + u) * T + t];
        acc.x = fmaf(sc, v8[u].x, acc.x); acc.y = fmaf(sc, v8[u].y, acc.y);
        acc.z = fmaf(sc, v8[u].z, acc.z); acc.w = fmaf(sc, v8[u].w, acc.w);
      }
    }
    for (; b < b1; ++b) {
      const float4 v = __ldcg(reinterpret_cast<const float4*>(p.blk_part + static_cast<int64_t>(b) * PART_STRIDE) + c4);
      const float sc = s_scale[b * T + t];
      acc.x = fmaf(sc, v.x, acc.x); acc.y = fmaf(sc, v.y, acc.y);
      acc.z = fmaf(sc, v.z, acc.z); acc.w = fmaf(sc, v.w, acc.w);
    }
    s_half[bg * 256 + c4] = acc;
  }
  __syncthreads();
  if (tid < 256) {
    const float4 a0 = s_half[tid], a1 = s_half[256 + tid], a2 = s_half[512 + tid], a3 = s_half[768 + tid];
    const int t = tid >> 7, j = (tid & 127) * 4;
    const float inv = 1.0f / s_l[t];
    const float v[4] = {(((a0.x + a1.x) + a2.x) + a3.x) * inv, (((a0.y + a1.y) + a2.y) + a3.y) * inv,
                        (((a0.z + a1.z) + a2.z) + a3.z) * inv, (((a0.w + a1.w) + a2.w) + a3.w) * inv};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      s_feat[t * (H + 1) + j + e] = v[e];
      p.features[t * (H + 1) + j + e] = v[e];
    }
  }
  if (tid < T) {
    s_feat[tid * (H + 1) + H] = sexv;
    p.features[tid * (H + 1) + H] = sexv;
  }
  __syncthreads();
  // heads: task 0 pooled vector -> classifier, task 1 -> site_classifier (model_toad.py:101,105)
  float* s_logit = dsm + T * (H + 1);  // [n_classes + 2]
  const int n_out = p.n_classes + 2;
  for (int o = warp; o < n_out; o += WARPS) {
    const bool is_site = o >= p.n_classes;
    const float* w = is_site ? p.wsite + static_cast<int64_t>(o - p.n_classes) * (H + 1) : p.wcls + static_cast<int64_t>(o) * (H + 1);
    const float* f = s_feat + (is_site ? (H + 1) : 0);
    float wv[17];  // (H+1)/32 rounded up: all weight loads in flight before the first FMA
#pragma unroll
    for (int q = 0; q < 17; ++q) wv[q] = (lane + 32 * q) < H + 1 ? __ldg(w + lane + 32 * q) : 0.f;
    float s = 0.f;
#pragma unroll
    for (int q = 0; q < 17; ++q) s = fmaf(wv[q], (lane + 32 * q) < H + 1 ? f[lane + 32 * q] : 0.f, s);
    s = warp_sum(s);
    if (lane == 0) s_logit[o] = s + (is_site ? __ldg(p.bsite + o - p.n_classes) : __ldg(p.bcls + o));
  }
  __syncthreads();
  if (tid < 2) {
    const int off = tid == 0 ? 0 : p.n_classes;
    const int cnt = tid == 0 ? p.n_classes : 2;
    float* lg = tid == 0 ? p.logits : p.site_logits;
    float* pr = tid == 0 ? p.y_prob : p.site_prob;
    int64_t* hat = tid == 0 ? p.y_hat : p.site_hat;
    float mx = -INFINITY;
    int arg = 0;
    for (int c = 0; c < cnt; ++c) {
      const float v = s_logit[off + c];
      lg[c] = v;
      if (v > mx) { mx = v; arg = c; }
    }
    float sum = 0.f;
    for (int c = 0; c < cnt; ++c) sum += expf(s_logit[off + c] - mx);
    for (int c = 0; c < cnt; ++c) pr[c] = expf(s_logit[off + c] - mx) / sum;
    hat[0] = arg;
  }
  if (tid == 0) *p.ticket = 0u;  // ready for the next launch on this workspace
}

template <int H_MODE>
int launch_tail(const TailParams& p_in, int sms, cudaStream_t stream) {
  TailParams p = p_in;
  if (p.N <= 0) return TOAD_ERR_ARG;
  const int nb = tail_blocks(p.N, sms);
  if (nb > MAX_BLOCKS) return TOAD_ERR_UNSUPPORTED;
  p.rows_per_block = static_cast<int32_t>((p.N + nb - 1) / nb);
  if (p.rows_per_block > MAX_CHUNK) return TOAD_ERR_UNSUPPORTED;
  const int dyn = WARPS * T * H * static_cast<int>(sizeof(float));
  auto kern = pool_heads_kernel<H_MODE>;
  TOAD_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn));
  kern<<<nb, THREADS, dyn, stream>>>(p);
  TOAD_CUDA_TRY(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------------------------
// weight preparation for the tensor-core path
// ---------------------------------------------------------------------------------------------
// (hi, lo) bf16 planes of a dense fp32 array (4 elements per thread).
__global__ void split_planes_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ hi,
                                    __nv_bfloat16* __restrict__ lo, int64_t n4) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  const float4 v = *reinterpret_cast<const float4*>(src + 4 * i);
  uint32_t h0, l0, h1, l1;
  split2(v.x, v.y, h0, l0);
  split2(v.z, v.w, h1, l1);
  *reinterpret_cast<uint2*>(hi + 4 * i) = make_uint2(h0, h1);
  *reinterpret_cast<uint2*>(lo + 4 * i) = make_uint2(l0, l1);
}

inline int launch_split_planes(const float* src, __nv_bfloat16* hi, __nv_bfloat16* lo, int64_t n, cudaStream_t stream) {
  if (n % 4 != 0) return TOAD_ERR_UNSUPPORTED;
  const int64_t n4 = n / 4;
  if (n4 == 0) return 0;
  split_planes_kernel<<<static_cast<unsigned>((n4 + 255) / 256), 256, 0, stream>>>(src, hi, lo, n4);
  TOAD_CUDA_TRY(cudaGetLastError());
  return 0;
}

// Gate weights packed for the EPI_GATE tile layout: packed row r of the [2D, K] operand is
//   tile = r / 2*half, within = r % (2*half): rows [0,half) <- Wa[tile*half + .], [half,2*half) <- Wb[...]
__global__ void split_gate_weights_kernel(const float* __restrict__ wa, const float* __restrict__ wb,
                                          __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo,
                                          int D, int K, int half) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;  // 4-element group
  const int64_t n4 = static_cast<int64_t>(2) * D * K / 4;
  if (i >= n4) return;
  const int64_t e = 4 * i;
  const int r = static_cast<int>(e / K), k = static_cast<int>(e % K);
  const int tile = r / (2 * half), within = r % (2 * half);
  const int j = tile * half + (within % half);
  const float* src = (within < half ? wa : wb) + static_cast<int64_t>(j) * K + k;
  const float4 v = *reinterpret_cast<const float4*>(src);
  uint32_t h0, l0, h1, l1;
  split2(v.x, v.y, h0, l0);
  split2(v.z, v.w, h1, l1);
  *reinterpret_cast<uint2*>(hi + e) = make_uint2(h0, h1);
  *reinterpret_cast<uint2*>(lo + e) = make_uint2(l0, l1);
}

inline int launch_split_gate_weights(const float* wa, const float* wb, __nv_bfloat16* hi, __nv_bfloat16* lo, int D,
                                     int K, int half, cudaStream_t stream) {
  if (K % 4 != 0 || D % half != 0) return TOAD_ERR_UNSUPPORTED;
  const int64_t n4 = static_cast<int64_t>(2) * D * K / 4;
  split_gate_weights_kernel<<<static_cast<unsigned>((n4 + 255) / 256), 256, 0, stream>>>(wa, wb, hi, lo, D, K, half);
  TOAD_CUDA_TRY(cudaGetLastError());
  return 0;
}

// fp32 path: part[n][t] = sum_j a[n,j]*b[n,j]*wc[t,j]   (attention_c without bias; warp per row)
__global__ void attn_c_kernel(const float* __restrict__ a, const float* __restrict__ b, const float* __restrict__ wc,
                              float* __restrict__ part, int64_t N, int D, int ntasks) {
  const int64_t row = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= N) return;
  float s[4] = {0.f, 0.f, 0.f, 0.f};
  for (int j = lane; j < D; j += 32) {
    const float g = a[row * D + j] * b[row * D + j];
#pragma unroll
    for (int t = 0; t < 4; ++t)
      if (t < ntasks) s[t] = fmaf(g, __ldg(wc + t * D + j), s[t]);
  }
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    if (t < ntasks) {
      const float v = warp_sum(s[t]);
      if (lane == 0) part[row * ntasks + t] = v;
    }
  }
}

inline int launch_attn_c(const float* a, const float* b, const float* wc, float* part, int64_t N, int D, int ntasks,
                         cudaStream_t stream) {
  if (N <= 0) return 0;
  attn_c_kernel<<<static_cast<unsigned>((N + 7) / 8), 256, 0, stream>>>(a, b, wc, part, N, D, ntasks);
  TOAD_CUDA_TRY(cudaGetLastError());
  return 0;
}

// A_out[n][t] = sum_z part[z][n][t] + bc[t]   (standalone Attn_Net_Gated output, [N, n_tasks])
__global__ void finish_scores_kernel(const float* __restrict__ part, int n_parts, const float* __restrict__ bc,
                                     float* __restrict__ out, int64_t N, int ntasks) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= N * ntasks) return;
  float s = 0.f;
  for (int z = 0; z < n_parts; ++z) s += part[static_cast<int64_t>(z) * N * ntasks + i];
  out[i] = s + __ldg(bc + (i % ntasks));
}

inline int launch_finish_scores(const float* part, int n_parts, const float* bc, float* out, int64_t N, int ntasks,
                                cudaStream_t stream) {
  const int64_t n = N * ntasks;
  if (n <= 0) return 0;
  finish_scores_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, stream>>>(part, n_parts, bc, out, N, ntasks);
  TOAD_CUDA_TRY(cudaGetLastError());
  return 0;
}

}  // namespace tail
}  // namespace toad
