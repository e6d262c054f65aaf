// fp32 CUDA-core GEMM with fused epilogues (sm_100a).  C[M,N] = epi( sum_k A(m,k) * B(n,k) ).
//
// Used where the tensor-core path does not apply yet: the fp32 training forward that keeps
// activations for the backward, and the backward's dgrad / wgrad contractions
// (autograd of models/model_toad.py:91 under utils/core_utils_mtl_concat.py:231).
// Operands are addressed through (row stride, k stride) so one kernel covers
// NT (nn.Linear forward), NN (dgrad) and TN (wgrad, K = patches, split-K over blockIdx.z).
#pragma once
#include "common.cuh"

namespace toad {
namespace simt {

enum Epi {
  EPI_STORE = 0,         // raw store (split-K partials go to c + z*M*N)
  EPI_BIAS_RELU = 1,     // relu(acc + bias[n])
  EPI_BIAS_TANH = 2,     // tanh(acc + bias[n])
  EPI_BIAS_SIGMOID = 3,  // sigmoid(acc + bias[n])
  EPI_BIAS = 4,          // acc + bias[n]
  EPI_RELUMASK = 5,      // acc * (mask[m,n] > 0)
  EPI_POOL_RELUMASK = 6  // (acc + c_in[m,n] + p0[m]*v0[n] + p1[m]*v1[n]) * (mask[m,n] > 0)
};

struct SgemmParams {
  const float* a; int64_t a_rs, a_ks;  // A(m,k) = a[m*a_rs + k*a_ks]
  const float* b; int64_t b_rs, b_ks;  // B(n,k) = b[n*b_rs + k*b_ks]
  float* c; int64_t ldc;
  int64_t M; int32_t N; int64_t K;
  int64_t k_chunk;      // K range per blockIdx.z (== K when gridDim.z == 1)
  const float* bias;    // [N]
  const float* mask;    // [M, ldmask]
  int64_t ldmask;
  const float* p0; const float* p1;  // [M] (stride p_stride)
  int64_t p_stride;
  const float* v0; const float* v1;  // [N]
  int32_t accumulate;   // EPI_RELUMASK/POOL: add existing c before masking
  float out_scale;      // EPI_RELUMASK/POOL: multiply the masked result (1/(1-p) when the mask tensor is post-dropout); 0 -> 1
  DropoutCfg drop;      // EPI_BIAS_RELU/TANH/SIGMOID: dropout on the activation (element index m*N + n)
  uint32_t drop_layer;
};

constexpr int BM = 128, BN = 128, BK = 16, TM = 8, TN = 8, NTHREADS = 256;

// A_KC / B_KC: operand is contiguous along k (true) or along its row index m / n (false).
template <bool A_KC, bool B_KC, int EPI>
__global__ void __launch_bounds__(NTHREADS) sgemm_kernel(const SgemmParams p) {
  __shared__ float As[2][BK][BM + 4];
  __shared__ float Bs[2][BK][BN + 4];
  const int tid = threadIdx.x;
  const int64_t m0 = static_cast<int64_t>(blockIdx.y) * BM;
  const int n0 = blockIdx.x * BN;
  const int64_t k_begin = static_cast<int64_t>(blockIdx.z) * p.k_chunk;
  const int64_t k_end = (k_begin + p.k_chunk < p.K) ? k_begin + p.k_chunk : p.K;
  const int tx = tid % 16, ty = tid / 16;  // thread computes rows ty*8.., cols tx*8..

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  float ra[8], rb[8];
  auto gload = [&](int64_t k0) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      int mm, kk;
      if (A_KC) { kk = tid % BK; mm = tid / BK + i * (NTHREADS / BK); }
      else      { mm = tid % BM; kk = tid / BM + i * (NTHREADS / BM); }
      const int64_t gm = m0 + mm, gk = k0 + kk;
      ra[i] = (gm < p.M && gk < k_end) ? __ldg(p.a + gm * p.a_rs + gk * p.a_ks) : 0.f;
      int nn, kb;
      if (B_KC) { kb = tid % BK; nn = tid / BK + i * (NTHREADS / BK); }
      else      { nn = tid % BN; kb = tid / BN + i * (NTHREADS / BN); }
      const int64_t gn = n0 + nn, gkb = k0 + kb;
      rb[i] = (gn < p.N && gkb < k_end) ? __ldg(p.b + gn * p.b_rs + gkb * p.b_ks) : 0.f;
    }
  };
  auto sstore = [&](int buf) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      int mm, kk;
      if (A_KC) { kk = tid % BK; mm = tid / BK + i * (NTHREADS / BK); }
      else      { mm = tid % BM; kk = tid / BM + i * (NTHREADS / BM); }
      As[buf][kk][mm] = ra[i];
      int nn, kb;
      if (B_KC) { kb = tid % BK; nn = tid / BK + i * (NTHREADS / BK); }
      else      { nn = tid % BN; kb = tid / BN + i * (NTHREADS / BN); }
      Bs[buf][kb][nn] = rb[i];
    }
  };

  int buf = 0;
  if (k_begin < k_end) {
    gload(k_begin);
    sstore(0);
  }
  __syncthreads();
  for (int64_t k0 = k_begin; k0 < k_end; k0 += BK) {
    const bool has_next = k0 + BK < k_end;
    if (has_next) gload(k0 + BK);
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float a[TM], b[TN];
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * TM]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * TM + 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * TN]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * TN + 4]);
      a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w; a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
      b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w; b[4] = b1.x; b[5] = b1.y; b[6] = b1.z; b[7] = b1.w;
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (has_next) {
      sstore(buf ^ 1);
      __syncthreads();
      buf ^= 1;
    }
  }

  float* cbase = p.c;
  if (EPI == EPI_STORE) cbase += static_cast<int64_t>(blockIdx.z) * p.M * p.ldc;
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int64_t gm = m0 + ty * TM + i;
    if (gm >= p.M) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int gn = n0 + tx * TN + j;
      if (gn >= p.N) continue;
      float v = acc[i][j];
      if (EPI == EPI_BIAS_RELU) v = fmaxf(v + __ldg(p.bias + gn), 0.f);
      else if (EPI == EPI_BIAS_TANH) v = tanh_acc(v + __ldg(p.bias + gn));
      else if (EPI == EPI_BIAS_SIGMOID) v = sigmoid_acc(v + __ldg(p.bias + gn));
      if (EPI == EPI_BIAS_RELU || EPI == EPI_BIAS_TANH || EPI == EPI_BIAS_SIGMOID)
        v = dropout_apply(p.drop, p.drop_layer, static_cast<unsigned long long>(gm) * p.N + gn, v);
      else if (EPI == EPI_BIAS) v = v + (p.bias ? __ldg(p.bias + gn) : 0.f);
      else if (EPI == EPI_RELUMASK || EPI == EPI_POOL_RELUMASK) {
        if (p.accumulate) v += cbase[gm * p.ldc + gn];
        if (EPI == EPI_POOL_RELUMASK)
          v += __ldg(p.p0 + gm * p.p_stride) * __ldg(p.v0 + gn) + __ldg(p.p1 + gm * p.p_stride) * __ldg(p.v1 + gn);
        v = (__ldg(p.mask + gm * p.ldmask + gn) > 0.f) ? v : 0.f;
        if (p.out_scale != 0.f) v *= p.out_scale;
      }
      cbase[gm * p.ldc + gn] = v;
    }
  }
}

template <bool A_KC, bool B_KC, int EPI>
int launch_sgemm(const SgemmParams& p, int splits, cudaStream_t stream) {
  if (p.M <= 0 || p.N <= 0) return 0;
  dim3 grid(static_cast<unsigned>((p.N + BN - 1) / BN), static_cast<unsigned>((p.M + BM - 1) / BM),
            static_cast<unsigned>(splits));
  sgemm_kernel<A_KC, B_KC, EPI><<<grid, NTHREADS, 0, stream>>>(p);
  TOAD_CUDA_TRY(cudaGetLastError());
  return 0;
}

// out[i] = sum_z part[z*n + i] in fixed z order (deterministic split-K / partial reduction).
__global__ void reduce_partials_kernel(const float* __restrict__ part, float* __restrict__ out, int64_t n, int splits) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float s = 0.f;
  for (int z = 0; z < splits; ++z) s += part[static_cast<int64_t>(z) * n + i];
  out[i] = s;
}

inline int launch_reduce_partials(const float* part, float* out, int64_t n, int splits, cudaStream_t stream) {
  if (n <= 0) return 0;
  reduce_partials_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, stream>>>(part, out, n, splits);
  TOAD_CUDA_TRY(cudaGetLastError());
  return 0;
}

}  // namespace simt
}  // namespace toad
