// top-k of one attention-score row (config 5: torch.topk over results['A'][t], SURVEY.md F6;
// the reference's own uses are k=1 at models/model_toad.py:102,106).
// Exact and deterministic: a 3-pass (11 + 11 + 10 bit) radix select of the k-th largest key over ALL SMs (one
// cooperative launch; per-CTA shared-memory histograms merged into global ones, a grid barrier per pass), an
// ordered compaction (ties at the k-th value resolved to the lowest indices, like torch.topk), and a bitonic sort
// of the k winners by (value desc, index asc) in one CTA.  The row (<= a few MB) is read 4 times and stays in L2;
// the time is the 4 grid barriers, not the bytes.
#pragma once
#include <cooperative_groups.h>

#include "common.cuh"

namespace toad {
namespace topk {

namespace cg = cooperative_groups;

constexpr int THREADS = 1024;
constexpr int KMAX = 2048;
constexpr int MAX_CTAS = 256;
constexpr int BINS = 2048;  // 11-bit digits (the last pass uses 10 bits)

struct TopkWs {                       // caller-owned, zeroed by launch_topk before the kernel
  unsigned int hist[3][BINS];
  unsigned int n_items;               // append cursor of the strictly-greater items
  unsigned int pad[15];
  unsigned int cta_eq[MAX_CTAS][BINS / 2];  // per-CTA histogram of the last pass (for the index-ordered tie rule)
  unsigned long long items[KMAX];
};

__device__ __forceinline__ uint32_t f2key(float f) {
  const uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);  // ascending uint order == ascending float order
}
__device__ __forceinline__ float key2f(uint32_t k) {
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7FFFFFFFu) : ~k);
}

__global__ void __launch_bounds__(THREADS, 1) topk_kernel(const float* __restrict__ scores, int64_t n, int k, TopkWs* ws,
                                                          float* __restrict__ out_vals, int64_t* __restrict__ out_idx) {
  __shared__ unsigned int sh[BINS];            // this CTA's histogram of the current digit
  __shared__ unsigned long long items[KMAX];   // (CTA 0) the winners, sorted in place; also scan scratch
  __shared__ unsigned int s_wsum[32], s_d, s_cum, s_warp_gt[32], s_warp_eq[32], s_base_eq, s_gt_base;
  cg::grid_group grid = cg::this_grid();
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t per = (n + gridDim.x - 1) / gridDim.x;   // contiguous chunk per CTA (ties are resolved in index order)
  const int64_t lo = static_cast<int64_t>(blockIdx.x) * per;
  const int64_t hi = (lo + per) < n ? (lo + per) : n;

  // ---- radix select: after the loop `prefix` is the key of the k-th largest element
  uint32_t prefix = 0, mask = 0;
  unsigned int need = static_cast<unsigned int>(k);  // how many still to take among keys matching prefix
  for (int pass = 0; pass < 3; ++pass) {
    const int shift = pass == 0 ? 21 : (pass == 1 ? 10 : 0);
    const int bins = pass == 2 ? BINS / 2 : BINS;
    for (int b = tid; b < BINS; b += THREADS) sh[b] = 0;
    __syncthreads();
    for (int64_t i = lo + tid; i < hi; i += THREADS) {
      const uint32_t key = f2key(__ldg(scores + i));
      if ((key & mask) == prefix) atomicAdd(&sh[(key >> shift) & (bins - 1)], 1u);
    }
    __syncthreads();
    for (int b = tid; b < bins; b += THREADS) {
      const unsigned int c = sh[b];
      if (c != 0) atomicAdd(&ws->hist[pass][b], c);
      if (pass == 2) ws->cta_eq[blockIdx.x][b] = c;
    }
    __threadfence();
    grid.sync();
    // every CTA finds the digit where the count from the top crosses `need` (same global histogram -> same answer)
    unsigned int* g = reinterpret_cast<unsigned int*>(items);  // [BINS] scratch
    for (int b = tid; b < BINS; b += THREADS) g[b] = b < bins ? __ldcg(&ws->hist[pass][b]) : 0u;
    __syncthreads();
    const unsigned int c_hi = g[2 * tid + 1], c_lo = g[2 * tid];
    // suffix sums over threads (thread t owns bins 2t, 2t+1): inclusive scan in reversed thread order
    unsigned int v = c_hi + c_lo, incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned int u = __shfl_down_sync(0xffffffffu, incl, o);
      if (lane + o < 32) incl += u;
    }
    if (lane == 0) s_wsum[warp] = incl;  // total of this warp
    __syncthreads();
    unsigned int above = incl - v;       // bins of higher threads in this warp
    for (int w2 = warp + 1; w2 < 32; ++w2) above += s_wsum[w2];
    if (above < need && need <= above + c_hi) { s_d = 2 * tid + 1; s_cum = above; }
    else if (above + c_hi < need && need <= above + c_hi + c_lo) { s_d = 2 * tid; s_cum = above + c_hi; }
    __syncthreads();
    prefix |= static_cast<uint32_t>(s_d) << shift;
    need -= s_cum;
    mask |= static_cast<uint32_t>(bins - 1) << shift;
    __syncthreads();
  }
  const uint32_t kth = prefix;  // `need` elements equal to kth are taken, lowest indices first
  const unsigned int n_gt = static_cast<unsigned int>(k) - need;
  const unsigned int d_last = kth & (BINS / 2 - 1);

  // ---- compaction: strictly greater items are appended in any order (the sort fixes it); items equal to the
  // k-th key go to the reserved slots [n_gt, k) in index order, this CTA starting after the lower CTAs' ties
  {
    unsigned int part = 0;
    for (int c = tid; c < static_cast<int>(blockIdx.x); c += THREADS) part += __ldcg(&ws->cta_eq[c][d_last]);
    part = __reduce_add_sync(0xffffffffu, part);
    if (lane == 0) s_wsum[warp] = part;
    if (tid == 0) s_base_eq = 0;
    __syncthreads();
    if (tid == 0) {
      unsigned int t = 0;
      for (int w2 = 0; w2 < 32; ++w2) t += s_wsum[w2];
      s_base_eq = t;  // ties in lower CTAs
    }
    __syncthreads();
  }
  for (int64_t base = lo; base < hi; base += THREADS) {
    const int64_t i = base + tid;
    uint32_t key = 0;
    bool gt = false, eq = false;
    if (i < hi) {
      key = f2key(__ldg(scores + i));
      gt = key > kth;
      eq = key == kth;
    }
    const unsigned int bgt = __ballot_sync(0xffffffffu, gt), beq = __ballot_sync(0xffffffffu, eq);
    if (lane == 0) { s_warp_gt[warp] = __popc(bgt); s_warp_eq[warp] = __popc(beq); }
    __syncthreads();
    if (tid == 0) {
      unsigned int tg = 0;
      for (int w2 = 0; w2 < 32; ++w2) tg += s_warp_gt[w2];
      s_gt_base = tg != 0 ? atomicAdd(&ws->n_items, tg) : 0u;
    }
    __syncthreads();
    unsigned int off_gt = s_gt_base, off_eq = s_base_eq;
    for (int w2 = 0; w2 < warp; ++w2) { off_gt += s_warp_gt[w2]; off_eq += s_warp_eq[w2]; }
    const unsigned int lt_mask = (1u << lane) - 1u;
    const unsigned long long item = (static_cast<unsigned long long>(key) << 32) |
                                    static_cast<unsigned long long>(0xFFFFFFFFu - static_cast<uint32_t>(i));
    if (gt) ws->items[off_gt + __popc(bgt & lt_mask)] = item;
    if (eq) {
      const unsigned int pos = off_eq + __popc(beq & lt_mask);
      if (pos < need) ws->items[n_gt + pos] = item;
    }
    __syncthreads();
    if (tid == 0) {
      unsigned int te = 0;
      for (int w2 = 0; w2 < 32; ++w2) te += s_warp_eq[w2];
      s_base_eq += te;
    }
    __syncthreads();
  }
  __threadfence();
  grid.sync();
  if (blockIdx.x != 0) return;

  // ---- bitonic sort, descending on the 64-bit item (key, ~index): padding zeros sink to the end
  int np2 = 1;
  while (np2 < k) np2 <<= 1;
  for (int i = tid; i < np2; i += THREADS) items[i] = i < k ? __ldcg(&ws->items[i]) : 0ull;
  __syncthreads();
  for (int size = 2; size <= np2; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int i = tid; i < np2 / 2; i += THREADS) {
        const int a = 2 * i - (i & (stride - 1));
        const int b2 = a + stride;
        const bool desc = ((a & size) == 0);
        const unsigned long long x = items[a], y = items[b2];
        if ((x < y) == desc) { items[a] = y; items[b2] = x; }
      }
      __syncthreads();
    }
  }
  for (int i = tid; i < k; i += THREADS) {
    const unsigned long long it = items[i];
    out_vals[i] = key2f(static_cast<uint32_t>(it >> 32));
    out_idx[i] = static_cast<int64_t>(0xFFFFFFFFu - static_cast<uint32_t>(it & 0xFFFFFFFFull));
  }
}

inline size_t topk_workspace_bytes() { return sizeof(TopkWs); }

inline int launch_topk(const float* scores, int64_t n, int k, float* out_vals, int64_t* out_idx, void* workspace,
                       size_t workspace_bytes, int sms, cudaStream_t stream) {
  if (k <= 0 || k > KMAX || n < k || n > 0xFFFFFFFFll) return TOAD_ERR_UNSUPPORTED;
  if (workspace == nullptr || workspace_bytes < sizeof(TopkWs) || (reinterpret_cast<uintptr_t>(workspace) & 15) != 0)
    return TOAD_ERR_WORKSPACE;
  TopkWs* ws = static_cast<TopkWs*>(workspace);
  TOAD_CUDA_TRY(cudaMemsetAsync(ws, 0, offsetof(TopkWs, cta_eq), stream));  // histograms + cursor
  int64_t ctas = (n + 2 * THREADS - 1) / (2 * THREADS);   // >= 2048 elements per CTA
  if (ctas > sms) ctas = sms;
  if (ctas > MAX_CTAS) ctas = MAX_CTAS;
  if (ctas < 1) ctas = 1;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(static_cast<unsigned>(ctas));
  cfg.blockDim = dim3(THREADS);
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;  // grid.sync(): every CTA must be resident
  attr[0].val.cooperative = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  TOAD_CUDA_TRY(cudaLaunchKernelEx(&cfg, topk_kernel, scores, n, k, ws, out_vals, out_idx));
  TOAD_CUDA_TRY(cudaGetLastError());
  return 0;
}

// out[i, :] = table[idx[i], :] for rows of `row_bytes` bytes (4-byte words): the coordinates of the top-k patches
// (the reference's h5 bags carry `coords` [N, 2] next to `features`, datasets/dataset_mtl_concat.py:377-383).
__global__ void gather_rows_kernel(const uint32_t* __restrict__ table, int64_t n_rows, int row_words,
                                   const int64_t* __restrict__ idx, int k, uint32_t* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= k * row_words) return;
  const int r = i / row_words, c = i - r * row_words;
  const int64_t src = idx[r];
  out[i] = (src >= 0 && src < n_rows) ? table[src * row_words + c] : 0u;
}

inline int launch_gather_rows(const void* table, int64_t n_rows, int row_bytes, const int64_t* idx, int k, void* out,
                              cudaStream_t stream) {
  if (row_bytes <= 0 || row_bytes % 4 != 0 || k <= 0 || n_rows <= 0) return TOAD_ERR_ARG;
  const int words = row_bytes / 4, total = k * words;
  gather_rows_kernel<<<(total + 255) / 256, 256, 0, stream>>>(static_cast<const uint32_t*>(table), n_rows, words, idx, k,
                                                              static_cast<uint32_t*>(out));
  TOAD_CUDA_TRY(cudaGetLastError());
  return 0;
}

}  // namespace topk
}  // namespace toad
