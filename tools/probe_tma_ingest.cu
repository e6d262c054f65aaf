// Probe: what bounds the L2 -> shared-memory operand feed of the trunk's GEMMs (measured ~24 B/clk/SM when all 148 SMs
// stream TMA boxes), and does cluster multicast lift it?
//
// Every CTA runs a 4-deep ring of 32 KB TMA loads (two [64 x 128 rows] SWIZZLE_128B boxes) from an L2-resident matrix and
// does nothing else.  Modes:
//   0 unique     every CTA streams its own rows (distinct L2 lines per CTA)
//   1 shared     the CTAs of a cluster load the SAME rows, each with its own unicast loads
//   2 multicast  the CTAs of a cluster load the same rows: each issues 1/csz of the tile with .multicast::cluster to all
// Reported: bytes landed in shared memory per SM clock per SM, and the aggregate TB/s.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -I toad_b200/csrc -o build_probe/probe_tma_ingest tools/probe_tma_ingest.cu -lcuda
#include "gemm_tc.cuh"
#include <cstdio>
#include <vector>

using namespace toad;
using namespace toad::tc;

constexpr int STAGES = 4;
constexpr int TILE_ROWS = 128;                 // matrix rows of 256 B (128 fp16 columns)
constexpr int TILE_BYTES = TILE_ROWS * 256;    // 32 KB = two [64 columns x 128 rows] boxes

__device__ __forceinline__ void tma_load_2d_mc(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "h"(mask)
      : "memory");
}

template <int CSZ>
__global__ void __launch_bounds__(128, 1)
ingest_kernel(const __grid_constant__ CUtensorMap tm, int mode, int iters, int rows_total, long long* cycles) {
  extern __shared__ uint8_t raw[];
  __shared__ __align__(8) uint64_t bar_full[STAGES];
  __shared__ __align__(8) uint64_t bar_empty[STAGES];   // multicast: every CTA of the cluster has consumed the stage
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  const uint32_t rank = CSZ > 1 ? cluster_ctarank() : 0u;
  const int cluster_id = blockIdx.x / CSZ;
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(smem_u32(&bar_full[s]), 1);
      mbar_init(smem_u32(&bar_empty[s]), CSZ);
    }
    fence_barrier_init();
  }
  __syncthreads();
  if (CSZ > 1) cluster_sync_all();
  const long long t0 = clock64();
  if (threadIdx.x == 0) {
    // producer and consumer in one thread: keep STAGES loads in flight
    const int n_clusters = gridDim.x / CSZ;
    for (int it = 0; it < iters + STAGES; ++it) {
      const int s = it % STAGES;
      if (it >= STAGES) {  // consume the load issued STAGES iterations ago
        mbar_wait(smem_u32(&bar_full[s]), ((it / STAGES) - 1) & 1);
        if (mode == 2) {   // tell every CTA of the cluster that this CTA's copy of the stage is free
          for (uint32_t r = 0; r < CSZ; ++r) mbar_arrive_cluster(smem_u32(&bar_empty[s]), r);
        }
      }
      if (it < iters) {
        const int unit = (mode == 0) ? (it * gridDim.x + blockIdx.x) : (it * n_clusters + cluster_id);
        const int row0 = (unit % (rows_total / TILE_ROWS)) * TILE_ROWS;
        const uint32_t dst = base + s * TILE_BYTES;
        const uint32_t fb = smem_u32(&bar_full[s]);
        mbar_expect_tx(fb, TILE_BYTES);
        if (mode == 2) {
          if (it >= STAGES) mbar_wait(smem_u32(&bar_empty[s]), ((it / STAGES) - 1) & 1);   // all peers released it
          constexpr int PART = TILE_ROWS / CSZ;   // this CTA's slice of the tile, delivered to every CTA
          const uint16_t mask = static_cast<uint16_t>((1u << CSZ) - 1u);
          for (int h = 0; h < 2; ++h)   // this CTA's [64 x PART] slice of each column half, delivered to every CTA of the cluster
            tma_load_2d_mc(dst + h * (TILE_BYTES / 2) + rank * PART * 128, &tm, fb, h * 64, row0 + static_cast<int>(rank) * PART, mask);
        } else {
          for (int h = 0; h < 2; ++h) tma_load_2d<1>(dst + h * (TILE_BYTES / 2), &tm, fb, h * 64, row0);
        }
      }
    }
  }
  __syncthreads();
  if (CSZ > 1) cluster_sync_all();
  if (threadIdx.x == 0) cycles[blockIdx.x] = clock64() - t0;
}

template <int CSZ>
static void run(const CUtensorMap& tm_full, const CUtensorMap& tm_part, int rows_total, long long* dcyc, int sms) {
  const int smem = STAGES * TILE_BYTES + 1024;
  cudaFuncSetAttribute(ingest_kernel<CSZ>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(ingest_kernel<CSZ>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  const int grid = sms / CSZ * CSZ;
  const int iters = 2000;
  for (int mode = 0; mode < 3; ++mode) {
    if (CSZ == 1 && mode > 0) continue;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CSZ; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    float best = 1e30f;
    long long cyc = 0;
    for (int rep = 0; rep < 3; ++rep) {
      cudaEvent_t e0, e1;
      cudaEventCreate(&e0); cudaEventCreate(&e1);
      cudaEventRecord(e0);
      cudaError_t le = cudaLaunchKernelEx(&cfg, ingest_kernel<CSZ>, mode == 2 ? tm_part : tm_full, mode, iters, rows_total, dcyc);
      cudaEventRecord(e1);
      cudaError_t e = cudaDeviceSynchronize();
      if (le != cudaSuccess || e != cudaSuccess) { printf("csz %d mode %d: %s / %s\n", CSZ, mode, cudaGetErrorString(le), cudaGetErrorString(e)); return; }
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      if (ms < best) {
        best = ms;
        std::vector<long long> h(grid);
        cudaMemcpy(h.data(), dcyc, grid * sizeof(long long), cudaMemcpyDeviceToHost);
        cyc = 0;
        for (long long v : h) cyc = v > cyc ? v : cyc;
      }
    }
    const double bytes_per_cta = static_cast<double>(iters) * TILE_BYTES;
    const char* names[3] = {"unique", "shared-unicast", "multicast"};
    printf("cluster %d  %-15s grid %3d: %7.3f ms  %6.2f B/clk/SM landed  aggregate %6.2f TB/s landed\n", CSZ, names[mode], grid, best,
           bytes_per_cta / static_cast<double>(cyc), bytes_per_cta * grid / (best * 1e-3) / 1e12);
  }
}

int main() {
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  // [rows_total x 128 columns] fp16 matrix = 48 MB: resident in the 126 MB L2 after the first pass
  const int rows_total = 192 * 1024;
  __half* d;
  cudaMalloc(&d, static_cast<size_t>(rows_total) * 256);
  cudaMemset(d, 0, static_cast<size_t>(rows_total) * 256);
  long long* dcyc;
  cudaMalloc(&dcyc, 1024 * sizeof(long long));
  // box = 64 columns x 128 rows (one half tile of 16 KB); multicast parts use 128 / csz rows
  for (int csz : {1, 2, 4}) {
    CUtensorMap tm_full, tm_part;
    if (make_bf16_tmap(&tm_full, d, rows_total, 128, 128) || make_bf16_tmap(&tm_part, d, rows_total, 128, 128 / csz)) { printf("tmap failed\n"); return 1; }
    if (csz == 1) run<1>(tm_full, tm_part, rows_total, dcyc, sms);
    if (csz == 2) run<2>(tm_full, tm_part, rows_total, dcyc, sms);
    if (csz == 4) run<4>(tm_full, tm_part, rows_total, dcyc, sms);
  }
  return 0;
}
