// HBM-bound tail of the TOAD forward (reference models/model_toad.py:92-107):
//   A_raw = scores^T ; P = softmax(A_raw, dim=1) over all N patches, per task ;
//   M = P . h ; M = cat(M, sex) ; the two linear heads ; top-1 ; softmaxes.
// One launch: every CTA streams a contiguous chunk of rows of h exactly once (128-bit
// coalesced loads, warp-per-row with 4 rows in flight, register accumulators), keeps (max,
// sum-exp, weighted sum) partials; atomic tickets elect the last CTA of each group of >= 8 to
// fold the group's partials and the last group to fold the group partials and evaluate the
// heads, all in fixed order -- deterministic, no float atomics.
// (-DTOAD_TAIL_DEBUG + TOAD_TAIL_STOP=k builds a profiling variant that leaves after phase k:
//  how the per-phase times in DESIGN.md section 7 were taken.)
//
// Also here: the small weight-preparation kernels of the tensor-core path and the
// attention_c contraction of the fp32 path.
#pragma once
#include "common.cuh"

namespace toad {
namespace tail {

constexpr int H = 512;          // hid_dim (model_toad.py:56, both size_args)
constexpr int T = 2;            // n_tasks (model_toad.py:66)
constexpr int THREADS = 512;    // 16 warps, one CTA per SM, <= 128 registers: 4 rows (8 KB) in flight per warp
constexpr int WARPS = THREADS / 32;
constexpr int ROWS_IN_FLIGHT = 4;
constexpr int MIN_GROUP = 8;    // two-level merge: the last CTA of each group of >= 8 folds the group, the last group folds all
constexpr int MAX_GROUPS = 32;
constexpr int MAX_HEADS_SMEM = 64;  // head rows staged in shared memory (64 x 513 floats = 128 KB)
constexpr int MAX_CHUNK = 2048; // rows per CTA (scores cached in smem)
constexpr int PART_STRIDE = T * (H + 2);  // one partial: acc[T][H], then (m,l)[T]

enum { H_F32 = 0, H_SPLIT = 1 };

constexpr int MAX_BATCH = 16;   // slides per launch (grid.y) of the batched forward
struct TailBatch {              // slide s = rows [off[s], off[s+1]) of the trunk's row space
  int32_t n_slides;
  int32_t pad;
  int64_t off[MAX_BATCH + 1];
};

// One launch serves grid.y slides whose rows sit back to back in the trunk's arrays (toad_fwd_batch); the plain
// forward is the n_slides = 1 case.  Per-slide outputs / scratch are slide-major blocks behind the given pointers.
struct TailParams {
  const float* part;   // [n_parts][stride][T] score partials (no bias)
  int32_t n_parts;
  const float* bc;     // [T]
  float* a_raw;        // [T][stride]
  int64_t stride;      // total rows of the trunk arrays (= N for a single slide)
  TailBatch batch;
  int32_t parts_per_slide;  // partial blocks per slide in blk_part (gridDim.x + MAX_GROUPS)
  const float* h_f32;  // [N, H]          (H_F32)
  const __nv_bfloat16* h_hi;  // [N, H]   (H_SPLIT)
  const __nv_bfloat16* h_lo;
  int64_t N;           // (host side only: rows of the largest slide, sizes the grid)
  int32_t rows_per_block;
  const float* sex;    // [n_slides]
  const float* wcls; const float* bcls; int32_t n_classes;
  const float* wsite; const float* bsite;
  float* features; float* logits; float* y_prob; int64_t* y_hat;
  float* site_logits; float* site_prob; int64_t* site_hat; float* stats;
  float* blk_part;     // per slide [gridDim.x + MAX_GROUPS][PART_STRIDE]: per-CTA partials, then the group partials
  unsigned int* ticket;  // per slide 64 zero-initialised counters (self-resetting): global, then one per group
  int32_t group;       // CTAs per group
  int32_t heads_in_smem;  // head weights staged in shared memory (n_classes + 2 <= MAX_HEADS_SMEM)
#ifdef TOAD_TAIL_DEBUG
  int32_t dbg_stop;    // profiling builds only: leave the kernel after phase dbg_stop
#endif
  int32_t attention_only;
};

inline int tail_blocks(int64_t n, int sms) {
  int64_t b = static_cast<int64_t>(sms);  // one CTA per SM
  const int64_t need = (n + MAX_CHUNK - 1) / MAX_CHUNK;
  if (b < need) b = need;
  const int64_t most = (n + 31) / 32;  // at least 32 rows per CTA
  if (b > most) b = most;
  if (b < 1) b = 1;
  return static_cast<int>(b);
}

inline int tail_group(int nb) {
  int g = (nb + MAX_GROUPS - 1) / MAX_GROUPS;
  return g < MIN_GROUP ? MIN_GROUP : g;
}

// Fold n partials {acc[T][H], (m, l)[T]} at src (stride PART_STRIDE, read through L2) in fixed order:
//   m = max_b m_b ;  l = sum_b exp(m_b - m) l_b ;  acc = sum_b exp(m_b - m) acc_b
// Results: s_m / s_l (valid after the call), and acc: thread c4 < 256 returns float4 column c4 of the flattened
// [T][H] accumulator (task c4 >> 7).  All THREADS threads of the CTA call it.  s_stage: 3*T*MAX_FOLD floats of
// scratch, s_half: 256 float4.  Latency-lean: the first accumulator loads are in flight before anything waits;
// warp t derives task t's maximum, weights and sum with one (m, l) load per lane and warp shuffles (n <= 32: a
// single L2 round trip); one barrier publishes the weights, one joins the two halves of the sum.
constexpr int MAX_FOLD = 512;
__device__ __forceinline__ float4 fold_partials(const float* __restrict__ src, int n, float* s_stage, float4* s_half,
                                                float* s_m, float* s_l) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int c4 = tid & 255, bg = tid >> 8;  // 2 thread groups x 256 float4 columns
  const int t = c4 >> 7;                    // 128 float4 per task
  const int bper = (n + 1) / 2;
  const int b0 = bg * bper, b1 = (b0 + bper) < n ? (b0 + bper) : n;
  float4 v8[8];
#pragma unroll
  for (int u = 0; u < 8; ++u)
    if (b0 + u < b1) v8[u] = __ldcg(reinterpret_cast<const float4*>(src + static_cast<int64_t>(b0 + u) * PART_STRIDE) + c4);
  float* s_wt = s_stage;                                            // [T][n] weights exp(m_b - m)
  float2* s_ml = reinterpret_cast<float2*>(s_stage + T * MAX_FOLD);  // [T][n] (m_b, l_b)
  if (warp < T) {
    const float* ml = src + T * H + 2 * warp;
    float mloc = -INFINITY;
    for (int b = lane; b < n; b += 32) {
      const float2 v = __ldcg(reinterpret_cast<const float2*>(ml + static_cast<int64_t>(b) * PART_STRIDE));
      s_ml[warp * n + b] = v;
      mloc = fmaxf(mloc, v.x);
    }
    const float m = warp_max(mloc);
    float lloc = 0.f;
    for (int b = lane; b < n; b += 32) {
      const float2 v = s_ml[warp * n + b];  // (written by this same lane)
      const float w = v.x == -INFINITY ? 0.f : expf(v.x - m);  // empty partial (also when all are empty): weight 0
      s_wt[warp * n + b] = w;
      lloc = fmaf(w, v.y, lloc);
    }
    const float l = warp_sum(lloc);
    if (lane == 0) { s_m[warp] = m; s_l[warp] = l; }
  }
  __syncthreads();
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int b = b0; b < b1; b += 8) {
    float4 cur[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) cur[u] = v8[u];
#pragma unroll
    for (int u = 0; u < 8; ++u)  // next batch in flight while this one is consumed
      if (b + 8 + u < b1) v8[u] = __ldcg(reinterpret_cast<const float4*>(src + static_cast<int64_t>(b + 8 + u) * PART_STRIDE) + c4);
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      if (b + u < b1) {
        const float sc = s_wt[t * n + b + u];
        acc.x = fmaf(sc, cur[u].x, acc.x); acc.y = fmaf(sc, cur[u].y, acc.y);
        acc.z = fmaf(sc, cur[u].z, acc.z); acc.w = fmaf(sc, cur[u].w, acc.w);
      }
    }
  }
  if (bg == 1) s_half[c4] = acc;
  __syncthreads();
  if (bg == 0) {
    const float4 o = s_half[c4];
    acc = make_float4(acc.x + o.x, acc.y + o.y, acc.z + o.z, acc.w + o.w);
  }
  return acc;
}

template <int H_MODE>
__global__ void __launch_bounds__(THREADS, 1) pool_heads_kernel(const TailParams p) {
  extern __shared__ float dsm[];               // [WARPS][T*H] cross-warp reduction buffer
  __shared__ float s_score[T][MAX_CHUNK];      // this CTA's scores, then exp(score - max); reused by the merge
  __shared__ float s_red[T][WARPS];
  __shared__ float s_m[T], s_l[T];
  __shared__ unsigned int s_ticket;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int slide = blockIdx.y;
  const int64_t row_base = p.batch.off[slide];               // first row of this slide in the trunk arrays
  const int64_t n_rows = p.batch.off[slide + 1] - row_base;
  const int64_t rpb = (n_rows + gridDim.x - 1) / gridDim.x;  // <= MAX_CHUNK (checked by the host for the largest slide)
  const int64_t l0_ = static_cast<int64_t>(blockIdx.x) * rpb;
  int64_t l1_ = l0_ + rpb;
  if (l1_ > n_rows) l1_ = n_rows;
  const int rows = l1_ > l0_ ? static_cast<int>(l1_ - l0_) : 0;
  const int64_t r0 = row_base + l0_;                         // global row of this CTA's first row

  // programmatic dependent launch: this grid may have been started while the gate GEMM's last wave was still
  // running; nothing above touches global memory (see gemm_tc.cuh: pdl_wait_then_release)
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  // head weights -> shared memory in the background (used only by the CTA that ends up evaluating the heads)
  float* s_w = dsm + WARPS * T * H;            // [n_classes + 2][H + 1]
  if (p.heads_in_smem && !p.attention_only) {
    const int n_cls = p.n_classes * (H + 1), n_all = n_cls + 2 * (H + 1);
    for (int i = tid; i < n_all; i += THREADS) {
      const float* src = i < n_cls ? p.wcls + i : p.wsite + (i - n_cls);
      asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(static_cast<uint32_t>(__cvta_generic_to_shared(s_w + i))), "l"(src) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  }

  // ---- phase 1: finish the scores (sum split-N partials + bias), write A_raw, local max
  float lmax[T] = {-INFINITY, -INFINITY};
  const float bc0 = __ldg(p.bc), bc1 = __ldg(p.bc + 1);
  for (int i = tid; i < rows; i += THREADS) {
    const int64_t row = r0 + i;
    float s0 = 0.f, s1 = 0.f;
    for (int z0 = 0; z0 < p.n_parts; z0 += 8) {  // 8 independent loads in flight (n_parts = 6 for "big"), summed in order
      float2 v[8];
#pragma unroll
      for (int z = 0; z < 8; ++z)
        v[z] = z0 + z < p.n_parts ? __ldg(reinterpret_cast<const float2*>(p.part + (static_cast<int64_t>(z0 + z) * p.stride + row) * T))
                                  : make_float2(0.f, 0.f);
#pragma unroll
      for (int z = 0; z < 8; ++z) {
        s0 += v[z].x;
        s1 += v[z].y;
      }
    }
    s0 += bc0;
    s1 += bc1;
    p.a_raw[row] = s0;
    p.a_raw[p.stride + row] = s1;
    s_score[0][i] = s0;
    s_score[1][i] = s1;
    lmax[0] = fmaxf(lmax[0], s0);
    lmax[1] = fmaxf(lmax[1], s1);
  }
  if (p.attention_only) return;
#ifdef TOAD_TAIL_DEBUG
  if (p.dbg_stop == 1) return;
#endif
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
  // softmax numerators once per row (not once per lane of the streaming warp), and their sums
  float l0 = 0.f, l1 = 0.f;
  for (int i = tid; i < rows; i += THREADS) {
    const float e0 = expf(s_score[0][i] - m0), e1 = expf(s_score[1][i] - m1);
    s_score[0][i] = e0;
    s_score[1][i] = e1;
    l0 += e0;
    l1 += e1;
  }
  l0 = warp_sum(l0);
  l1 = warp_sum(l1);
  __syncthreads();  // s_red (maxima) consumed; numerators visible to every warp
  if (lane == 0) { s_red[0][warp] = l0; s_red[1][warp] = l1; }

  // ---- phase 2: stream h once; warp per row, 16 columns per lane, ROWS_IN_FLIGHT rows (8 KB per warp) in flight
  float acc0[16], acc1[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) { acc0[i] = 0.f; acc1[i] = 0.f; }
  for (int i = warp; i < rows; i += WARPS * ROWS_IN_FLIGHT) {
    uint4 raw[ROWS_IN_FLIGHT][4];
#pragma unroll
    for (int u = 0; u < ROWS_IN_FLIGHT; ++u) {
      const int iu = i + u * WARPS;
      const int64_t row = r0 + (iu < rows ? iu : i);  // past the chunk: re-read row i (its weight is zeroed below)
      if (H_MODE == H_F32) {
        const float* hr = p.h_f32 + row * H;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float4 v = ld_stream_f4(hr + 4 * (lane + 32 * q));
          raw[u][q] = make_uint4(__float_as_uint(v.x), __float_as_uint(v.y), __float_as_uint(v.z), __float_as_uint(v.w));
        }
      } else {
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          raw[u][2 * q] = ld_stream_u4(p.h_hi + row * H + q * 256 + lane * 8);
          raw[u][2 * q + 1] = ld_stream_u4(p.h_lo + row * H + q * 256 + lane * 8);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < ROWS_IN_FLIGHT; ++u) {
      const int iu = i + u * WARPS;
      const bool ok = iu < rows;
      const float p0 = ok ? s_score[0][iu] : 0.f, p1 = ok ? s_score[1][iu] : 0.f;
      float hv[16];
      if (H_MODE == H_F32) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          hv[4 * q] = __uint_as_float(raw[u][q].x); hv[4 * q + 1] = __uint_as_float(raw[u][q].y);
          hv[4 * q + 2] = __uint_as_float(raw[u][q].z); hv[4 * q + 3] = __uint_as_float(raw[u][q].w);
        }
      } else {
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          const uint4 vh = raw[u][2 * q], vl = raw[u][2 * q + 1];
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
  }
#ifdef TOAD_TAIL_DEBUG
  if (p.dbg_stop == 2) { if (acc0[0] + acc1[3] == 123.456f) p.stats[0] = 1.f; return; }
#endif
  // column owned by accumulator e of this lane
  auto col_of = [&](int e) -> int {
    return H_MODE == H_F32 ? 4 * (lane + 32 * (e >> 2)) + (e & 3) : (e >> 3) * 256 + lane * 8 + (e & 7);
  };
#pragma unroll
  for (int e = 0; e < 16; ++e) {
    dsm[warp * (T * H) + col_of(e)] = acc0[e];
    dsm[warp * (T * H) + H + col_of(e)] = acc1[e];
  }
  __syncthreads();
  float* const blk_part = p.blk_part + static_cast<int64_t>(slide) * p.parts_per_slide * PART_STRIDE;
  unsigned int* const ticket = p.ticket + slide * 64;
  float* mine = blk_part + static_cast<int64_t>(blockIdx.x) * PART_STRIDE;
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

  // ---- phase 3: two-level merge, fixed order, no float atomics.  The last CTA of each group (atomic ticket)
  // folds the group's partials; the last group to finish folds the group partials and evaluates the heads.
  // (One CTA folding all 148 partials -- 600 KB through one SM's L2 port -- was half of this kernel's time.)
  const int nb = gridDim.x;
  const int grp = blockIdx.x / p.group, n_groups = (nb + p.group - 1) / p.group;
  const int g0 = grp * p.group, g_cnt = (g0 + p.group <= nb ? p.group : nb - g0);
  float* grp_part = blk_part + static_cast<int64_t>(nb) * PART_STRIDE;
  float4* s_half = reinterpret_cast<float4*>(dsm + 4096);  // [2][256] float4 scratch
  __threadfence();
  __syncthreads();
  if (tid == 0) s_ticket = atomicAdd(ticket + 1 + grp, 1u);
  __syncthreads();
  if (s_ticket != static_cast<unsigned int>(g_cnt - 1)) return;
#ifdef TOAD_TAIL_DEBUG
  if (p.dbg_stop == 3) { if (tid == 0) ticket[1 + grp] = 0u; return; }
#endif
  __threadfence();
  {
    const float4 v = fold_partials(blk_part + static_cast<int64_t>(g0) * PART_STRIDE, g_cnt, &s_score[0][0], s_half, s_m, s_l);
    float* gp = grp_part + static_cast<int64_t>(grp) * PART_STRIDE;
    if (tid < 256) reinterpret_cast<float4*>(gp)[tid] = v;
    if (tid < T) {
      gp[T * H + 2 * tid] = s_m[tid];
      gp[T * H + 2 * tid + 1] = s_l[tid];
    }
    if (tid == 0) ticket[1 + grp] = 0u;  // ready for the next launch on this workspace
  }
  __threadfence();
  __syncthreads();
  if (tid == 0) s_ticket = atomicAdd(ticket, 1u);
  __syncthreads();
  if (s_ticket != static_cast<unsigned int>(n_groups - 1)) return;
#ifdef TOAD_TAIL_DEBUG
  if (p.dbg_stop == 4) { if (tid == 0) *ticket = 0u; return; }
#endif
  __threadfence();
  const float4 pooled = fold_partials(grp_part, n_groups, &s_score[0][0], s_half, s_m, s_l);
  // this slide's result blocks
  float* const o_stats = p.stats + slide * 2 * T;
  float* const o_feat = p.features + static_cast<int64_t>(slide) * T * (H + 1);
  float* const o_logits = p.logits + static_cast<int64_t>(slide) * p.n_classes;
  float* const o_prob = p.y_prob + static_cast<int64_t>(slide) * p.n_classes;
  float* const o_slogits = p.site_logits + slide * 2;
  float* const o_sprob = p.site_prob + slide * 2;
  if (tid < T) {
    o_stats[2 * tid] = s_m[tid];
    o_stats[2 * tid + 1] = s_l[tid];
  }
  float* s_feat = dsm;                 // [T][H+1]
  const float sexv = __ldg(p.sex + slide);
  if (tid < 256) {
    const int t = tid >> 7, j = (tid & 127) * 4;
    const float inv = 1.0f / s_l[t];
    const float v[4] = {pooled.x * inv, pooled.y * inv, pooled.z * inv, pooled.w * inv};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      s_feat[t * (H + 1) + j + e] = v[e];
      o_feat[t * (H + 1) + j + e] = v[e];
    }
  }
  if (tid < T) {
    s_feat[tid * (H + 1) + H] = sexv;
    o_feat[tid * (H + 1) + H] = sexv;
  }
  __syncthreads();
#ifdef TOAD_TAIL_DEBUG
  if (p.dbg_stop == 5) { if (tid == 0) *ticket = 0u; return; }
#endif
  // heads: task 0 pooled vector -> classifier, task 1 -> site_classifier (model_toad.py:101,105).  The weight
  // rows were copied to shared memory by cp.async at kernel start (every CTA: only the last one gets here, and
  // none knows in advance), so no global latency is left on this serial path.
  float* s_logit = dsm + T * (H + 1);  // [n_classes + 2]
  const int n_out = p.n_classes + 2;
  if (p.heads_in_smem) {
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncthreads();
  }
  for (int o = warp; o < n_out; o += WARPS) {
    const bool is_site = o >= p.n_classes;
    const float* wg = is_site ? p.wsite + static_cast<int64_t>(o - p.n_classes) * (H + 1) : p.wcls + static_cast<int64_t>(o) * (H + 1);
    const float* w = p.heads_in_smem ? s_w + o * (H + 1) : wg;
    const float* f = s_feat + (is_site ? (H + 1) : 0);
    float wv[17];  // (H+1)/32 rounded up: all weight loads in flight before the first FMA
#pragma unroll
    for (int q = 0; q < 17; ++q) wv[q] = (lane + 32 * q) < H + 1 ? w[lane + 32 * q] : 0.f;
    float s = 0.f;
#pragma unroll
    for (int q = 0; q < 17; ++q) s = fmaf(wv[q], (lane + 32 * q) < H + 1 ? f[lane + 32 * q] : 0.f, s);
    s = warp_sum(s);
    if (lane == 0) s_logit[o] = s + (is_site ? __ldg(p.bsite + o - p.n_classes) : __ldg(p.bcls + o));
  }
  __syncthreads();
  if (warp < 2) {  // warp 0: classifier softmax / top-1, warp 1: site (ties -> lowest index, like torch.topk)
    const int off = warp == 0 ? 0 : p.n_classes;
    const int cnt = warp == 0 ? p.n_classes : 2;
    float* lg = warp == 0 ? o_logits : o_slogits;
    float* pr = warp == 0 ? o_prob : o_sprob;
    int64_t* hat = (warp == 0 ? p.y_hat : p.site_hat) + slide;
    float mx = -INFINITY;
    int arg = 0x7fffffff;
    for (int c = lane; c < cnt; c += 32) {
      const float v = s_logit[off + c];
      lg[c] = v;
      if (v > mx) { mx = v; arg = c; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float om = __shfl_xor_sync(0xffffffffu, mx, o);
      const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
      if (om > mx || (om == mx && oa < arg)) { mx = om; arg = oa; }
    }
    float sum = 0.f;
    for (int c = lane; c < cnt; c += 32) sum += expf(s_logit[off + c] - mx);
    sum = warp_sum(sum);
    for (int c = lane; c < cnt; c += 32) pr[c] = expf(s_logit[off + c] - mx) / sum;
    if (lane == 0) hat[0] = arg == 0x7fffffff ? 0 : arg;
  }
  if (tid == 0) *ticket = 0u;  // ready for the next launch on this workspace
}

template <int H_MODE>
int launch_tail(const TailParams& p_in, int sms, cudaStream_t stream, bool pdl = false) {
  TailParams p = p_in;
  if (p.N <= 0) return TOAD_ERR_ARG;
  if (p.n_classes + 2 > 1024) return TOAD_ERR_UNSUPPORTED;
  if (p.batch.n_slides == 0) {  // plain forward: one slide covering all N rows
    p.batch.n_slides = 1;
    p.batch.off[0] = 0;
    p.batch.off[1] = p.N;
    p.stride = p.N;
  }
  if (p.batch.n_slides < 1 || p.batch.n_slides > MAX_BATCH) return TOAD_ERR_ARG;
  const int nb = tail_blocks(p.N, sms);   // (N = rows of the largest slide)
  p.parts_per_slide = nb + MAX_GROUPS;
  p.group = tail_group(nb);
#ifdef TOAD_TAIL_DEBUG
  { const char* e = getenv("TOAD_TAIL_STOP"); p.dbg_stop = e ? atoi(e) : 0; }
#endif
  if (p.group > MAX_FOLD) return TOAD_ERR_UNSUPPORTED;  // > 16k CTAs = 33M patches
  p.rows_per_block = static_cast<int32_t>((p.N + nb - 1) / nb);
  if (p.rows_per_block > MAX_CHUNK) return TOAD_ERR_UNSUPPORTED;
  p.heads_in_smem = (p.n_classes + 2) <= MAX_HEADS_SMEM ? 1 : 0;
  const int dyn = (WARPS * T * H + (p.heads_in_smem ? (p.n_classes + 2) * (H + 1) : 0)) * static_cast<int>(sizeof(float));
  auto kern = pool_heads_kernel<H_MODE>;
  TOAD_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn));
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(nb, p.batch.n_slides);
  cfg.blockDim = dim3(THREADS);
  cfg.dynamicSmemBytes = dyn;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;   // see pool_heads_kernel: griddepcontrol.wait
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  TOAD_CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, p));
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
