// Probe: can a K-major SWIZZLE_128B UMMA operand start at an arbitrary 128-byte row of a TMA-written tile?
//
// A halo-reusing 3x3 convolution would load an input patch ONCE and feed its nine taps to tcgen05.mma as nine views of
// that patch shifted by whole pixels (= 128-byte rows).  The descriptor's start address is then not 1024-byte aligned
// and the question is what the hardware does with the swizzle phase; the matrix descriptor has a 3-bit "base offset"
// field (bits 49..51) for exactly that.  This program measures it: A = a [272 x 64] fp16 matrix in smem (TMA,
// SWIZZLE_128B), B = the 64 x 64 identity, D = A_view . B^T should equal rows s .. s+127 of A.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -I toad_b200/csrc -std=c++17 -o build_probe/probe_umma_shift tools/probe_umma_shift.cu -lcuda   (build_probe/ is git-ignored but travels with gpurun)
#include "gemm_tc.cuh"
#include <cstdio>
#include <vector>

using namespace toad;
using namespace toad::tc;

constexpr int ROWS = 272, KC = 64, NB = 64;

__global__ void __launch_bounds__(128, 1)
probe_kernel(const __grid_constant__ CUtensorMap ta, const __grid_constant__ CUtensorMap tb, int shift, int mode, float* out) {
  extern __shared__ uint8_t raw[];
  __shared__ uint64_t bar_load, bar_mma;
  __shared__ uint32_t tmem_slot;
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  const uint32_t sA = base, sB = base + 40960;   // A: 272 rows x 128 B = 34816 B (room to 40 KB), B: 64 x 128 B
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bar_load), 1);
    mbar_init(smem_u32(&bar_mma), 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<1>(smem_u32(&tmem_slot), 64);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (threadIdx.x == 0) {
    mbar_expect_tx(smem_u32(&bar_load), ROWS * 128 + NB * 128);
    tma_load_2d<1>(sA, &ta, smem_u32(&bar_load), 0, 0);                  // rows 0..135
    tma_load_2d<1>(sA + 136 * 128, &ta, smem_u32(&bar_load), 0, 136);    // rows 136..271
    tma_load_2d<1>(sB, &tb, smem_u32(&bar_load), 0, 0);
    mbar_wait(smem_u32(&bar_load), 0);
    tc_fence_after();
    constexpr uint32_t idesc = make_idesc_f16(128, NB);
    for (int k = 0; k < 4; ++k) {
      const uint32_t a_addr = sA + shift * 128 + k * 32;
      uint64_t ad = make_kmajor_sw128_desc(a_addr);
      if (mode == 1) ad |= static_cast<uint64_t>((a_addr >> 7) & 7) << 49;
      if (mode == 2) ad |= static_cast<uint64_t>((8 - ((a_addr >> 7) & 7)) & 7) << 49;
      umma_bf16<1>(tmem, ad, make_kmajor_sw128_desc(sB + k * 32), idesc, k > 0);
    }
    umma_commit<1>(smem_u32(&bar_mma));
  }
  __syncthreads();
  mbar_wait(smem_u32(&bar_mma), 0);
  tc_fence_after();
  for (int c = 0; c < 2; ++c) {
    uint32_t v[32];
    tmem_ld32(tmem + (static_cast<uint32_t>(warp * 32) << 16) + c * 32, v);
    tmem_ld_wait();
    for (int j = 0; j < 32; ++j) out[(warp * 32 + lane) * NB + c * 32 + j] = __uint_as_float(v[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<1>(tmem, 64);
}

static float a_val(int r, int k) { return static_cast<float>((r * 7 + k * 3) % 13 - 6) + 0.25f * static_cast<float>(r % 3); }

int main() {
  std::vector<__half> ha(ROWS * KC), hb(NB * KC);
  for (int r = 0; r < ROWS; ++r) for (int k = 0; k < KC; ++k) ha[r * KC + k] = __float2half(a_val(r, k));
  for (int n = 0; n < NB; ++n) for (int k = 0; k < KC; ++k) hb[n * KC + k] = __float2half(n == k ? 1.f : 0.f);
  __half *da, *db; float* dout;
  cudaMalloc(&da, ha.size() * 2); cudaMalloc(&db, hb.size() * 2); cudaMalloc(&dout, 128 * NB * 4);
  cudaMemcpy(da, ha.data(), ha.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(db, hb.data(), hb.size() * 2, cudaMemcpyHostToDevice);
  CUtensorMap ta, tb;
  if (make_bf16_tmap(&ta, da, ROWS, KC, 136) || make_bf16_tmap(&tb, db, NB, KC, NB)) { printf("tensor map failed\n"); return 1; }
  const int smem = 40960 + 8192 + 1024;
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int shifts[] = {0, 8, 1, 2, 3, 4, 5, 7, 9, 33, 66, 67, 68, 130, 137};
  std::vector<float> out(128 * NB);
  for (int mode = 0; mode < 3; ++mode) {
    for (int s : shifts) {
      cudaMemset(dout, 0xff, 128 * NB * 4);
      probe_kernel<<<1, 128, smem>>>(ta, tb, s, mode, dout);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("mode %d shift %d: %s\n", mode, s, cudaGetErrorString(e)); return 2; }
      cudaMemcpy(out.data(), dout, out.size() * 4, cudaMemcpyDeviceToHost);
      int bad = 0, first = -1;
      for (int m = 0; m < 128; ++m) for (int n = 0; n < NB; ++n)
        if (out[m * NB + n] != a_val(s + m, n)) { if (first < 0) first = m * NB + n; ++bad; }
      printf("base_offset mode %d  shift %3d rows: %s (%d of %d wrong", mode, s, bad ? "MISMATCH" : "exact", bad, 128 * NB);
      if (bad) printf("; first at m=%d n=%d got %g want %g", first / NB, first % NB, out[first], a_val(s + first / NB, first % NB));
      printf(")\n");
    }
  }
  return 0;
}
