// top-k of one attention-score row (config 5: torch.topk over results['A'][t], SURVEY.md F6;
// the reference's own uses are k=1 at models/model_toad.py:102,106).
// Exact: 4-pass 8-bit radix select of the k-th largest key, ordered compaction (ties resolved
// to the lowest indices), then a bitonic sort of the k winners by (value desc, index asc).
// Single CTA: the row is at most a few hundred thousand floats and stays L2-resident.
#pragma once
#include "common.cuh"

namespace toad {
namespace topk {

constexpr int THREADS = 1024;
constexpr int KMAX = 2048;

__device__ __forceinline__ uint32_t f2key(float f) {
  const uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);  // ascending uint order == ascending float order
}
__device__ __forceinline__ float key2f(uint32_t k) {
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7FFFFFFFu) : ~k);
}

__global__ void __launch_bounds__(THREADS) topk_kernel(const float* __restrict__ scores, int64_t n, int k,
                                                       float* __restrict__ out_vals, int64_t* __restrict__ out_idx) {
  __shared__ unsigned int hist[256];
  __shared__ unsigned long long items[KMAX];
  __shared__ unsigned int s_prefix, s_need, s_warp_gt[32], s_warp_eq[32], s_base_gt, s_base_eq;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  // ---- radix select: after the loop `prefix` is the key of the k-th largest element
  uint32_t prefix = 0, mask = 0;
  unsigned int need = static_cast<unsigned int>(k);  // how many still to take among keys matching prefix
  for (int shift = 24; shift >= 0; shift -= 8) {
    if (tid < 256) hist[tid] = 0;
    __syncthreads();
    for (int64_t i = tid; i < n; i += THREADS) {
      const uint32_t key = f2key(scores[i]);
      if ((key & mask) == prefix) atomicAdd(&hist[(key >> shift) & 0xFF], 1u);
    }
    __syncthreads();
    if (tid == 0) {
      unsigned int cum = 0;
      int d = 255;
      for (; d > 0; --d) {
        if (cum + hist[d] >= need) break;
        cum += hist[d];
      }
      s_prefix = prefix | (static_cast<uint32_t>(d) << shift);
      s_need = need - cum;
    }
    __syncthreads();
    prefix = s_prefix;
    need = s_need;
    mask |= 0xFFu << shift;
    __syncthreads();
  }
  const uint32_t kth = prefix;  // `need` elements equal to kth are taken, lowest indices first
  const unsigned int n_gt = static_cast<unsigned int>(k) - need;

  // ---- ordered compaction
  if (tid == 0) { s_base_gt = 0; s_base_eq = 0; }
  for (int i = tid; i < KMAX; i += THREADS) items[i] = 0ull;
  __syncthreads();
  for (int64_t base = 0; base < n; base += THREADS) {
    const int64_t i = base + tid;
    uint32_t key = 0;
    bool gt = false, eq = false;
    if (i < n) {
      key = f2key(scores[i]);
      gt = key > kth;
      eq = key == kth;
    }
    const unsigned int bgt = __ballot_sync(0xffffffffu, gt), beq = __ballot_sync(0xffffffffu, eq);
    if (lane == 0) { s_warp_gt[warp] = __popc(bgt); s_warp_eq[warp] = __popc(beq); }
    __syncthreads();
    unsigned int off_gt = s_base_gt, off_eq = s_base_eq;
    for (int w = 0; w < warp; ++w) { off_gt += s_warp_gt[w]; off_eq += s_warp_eq[w]; }
    const unsigned int lt_mask = (1u << lane) - 1u;
    const unsigned long long item = (static_cast<unsigned long long>(key) << 32) |
                                    static_cast<unsigned long long>(0xFFFFFFFFu - static_cast<uint32_t>(i));
    if (gt) items[off_gt + __popc(bgt & lt_mask)] = item;
    if (eq) {
      const unsigned int pos = off_eq + __popc(beq & lt_mask);
      if (pos < need) items[n_gt + pos] = item;
    }
    __syncthreads();
    if (tid == 0) {
      unsigned int tg = 0, te = 0;
      for (int w = 0; w < THREADS / 32; ++w) { tg += s_warp_gt[w]; te += s_warp_eq[w]; }
      s_base_gt += tg;
      s_base_eq += te;
    }
    __syncthreads();
  }

  // ---- bitonic sort, descending on the 64-bit item (key, ~index): padding zeros sink to the end
  int np2 = 1;
  while (np2 < k) np2 <<= 1;
  for (int size = 2; size <= np2; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int i = tid; i < np2 / 2; i += THREADS) {
        const int lo = 2 * i - (i & (stride - 1));
        const int hi = lo + stride;
        const bool desc = ((lo & size) == 0);
        const unsigned long long a = items[lo], b = items[hi];
        if ((a < b) == desc) { items[lo] = b; items[hi] = a; }
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

inline int launch_topk(const float* scores, int64_t n, int k, float* out_vals, int64_t* out_idx, cudaStream_t stream) {
  if (k <= 0 || k > KMAX || n < k || n > 0xFFFFFFFFll) return TOAD_ERR_UNSUPPORTED;
  topk_kernel<<<1, THREADS, 0, stream>>>(scores, n, k, out_vals, out_idx);
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
