// Loss and optimizer step of the reference training loop, one launch each (sm_100a).
//
//   loss = 0.75 * CE(logits, label) + 0.25 * CE(site_logits, site)     utils/core_utils_mtl_concat.py:213-215
//   optim.Adam(params, lr, weight_decay)                               utils/utils.py:65, core_utils:231-234
//
// In the eager loop these are ~25 + ~110 tiny launches per step (softmax, nll, scalar ops, eight
// multi-tensor passes) -- about 0.25 ms of a 1.7 ms step at N = 50k and most of the step for small bags.
#pragma once
#include "common.cuh"

namespace toad {
namespace train {

__device__ __forceinline__ float warp_max_f(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_sum_f(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Two cross-entropy heads (batch of one slide), their weighted sum and d(loss)/d(logits).
// Warp 0: the n_classes-way head, warp 1: the 2-way site head.  loss3 = {total, cls, site}.
// A target outside [0, C) makes that head's loss NaN (the reference raises a device-side assert).
__global__ void ce_loss_grad_kernel(const float* __restrict__ logits, const float* __restrict__ site_logits, int C,
                                    const int64_t* __restrict__ label, const int64_t* __restrict__ site, float w_cls,
                                    float w_site, float* __restrict__ loss3, float* __restrict__ dlogits,
                                    float* __restrict__ dsite) {
  __shared__ float s_loss[2];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp < 2) {
    const float* z = warp == 0 ? logits : site_logits;
    float* dz = warp == 0 ? dlogits : dsite;
    const int n = warp == 0 ? C : 2;
    const int64_t y = warp == 0 ? *label : *site;
    const float w = warp == 0 ? w_cls : w_site;
    float m = -INFINITY;
    for (int i = lane; i < n; i += 32) m = fmaxf(m, z[i]);
    m = warp_max_f(m);
    float s = 0.f;
    for (int i = lane; i < n; i += 32) s += expf(z[i] - m);
    s = warp_sum_f(s);
    const float lse = m + logf(s);
    const bool ok = y >= 0 && y < n;
    for (int i = lane; i < n; i += 32) {
      const float pr = expf(z[i] - lse);
      dz[i] = w * (pr - ((ok && i == static_cast<int>(y)) ? 1.f : 0.f));
    }
    if (lane == 0) s_loss[warp] = ok ? lse - z[y] : __int_as_float(0x7fc00000);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    loss3[1] = s_loss[0];
    loss3[2] = s_loss[1];
    loss3[0] = w_cls * s_loss[0] + w_site * s_loss[1];
  }
}

// torch.optim.Adam (L2 weight decay added to the gradient, bias-corrected, amsgrad off) over all 14 parameter
// tensors in one launch.  Gradient and both moments are flat buffers in toad_param_offsets order; the
// parameters stay where the module keeps them (14 pointers).
struct AdamArgs {
  float* p[14];
  int64_t off[15];
  const float* g;
  float* m;
  float* v;
  float grad_scale;  // applied to g first (1/world_size after a summing all-reduce)
  float wd, b1, b2, eps;
  float one_m_b1, one_m_b2;  // 1 - beta computed in double on the host (as torch's python scalars are)
  float step_size;     // lr / (1 - b1^t)
  float bc2_sqrt;      // sqrt(1 - b2^t)
};

__global__ void adam_kernel(const AdamArgs a) {
  const int64_t total = a.off[14];
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    int t = 0;
#pragma unroll
    for (int k = 1; k < 14; ++k) t += (i >= a.off[k]) ? 1 : 0;
    float* pp = a.p[t] + (i - a.off[t]);
    const float p = *pp;
    const float g = fmaf(a.wd, p, a.g[i] * a.grad_scale);
    float m = a.m[i], v = a.v[i];
    m = m + (g - m) * a.one_m_b1;                  // exp_avg.lerp_(grad, 1 - beta1)
    v = fmaf(a.one_m_b2, g * g, v * a.b2);         // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
    const float denom = sqrtf(v) / a.bc2_sqrt + a.eps;
    a.m[i] = m;
    a.v[i] = v;
    *pp = p - a.step_size * (m / denom);
  }
}

}  // namespace train
}  // namespace toad
