// tcgen05 (5th-gen tensor core) GEMM / implicit-GEMM convolution, sm_100a.
//
//   D[M,N] = epilogue( A[M,K] . B[N,K]^T )      fp32 accumulation in TMEM, two operand precisions (PREC):
//   PREC_BF16X3  A ~= A_hi + A_lo, B ~= B_hi + B_lo (bf16 pairs); A.B^T ~= A_hi.B_hi + A_hi.B_lo + A_lo.B_hi
//   PREC_F16X2   A = one fp16 plane, B ~= B_hi + B_lo (fp16 pairs);  A.B^T ~= A.B_hi + A.B_lo
//
// PREC_BF16X3 is the nn.Linear building block of the TOAD trunk (reference: models/model_toad.py:59,62 fc layers
// and :37-38 the gated-attention pair) and of its backward, at fp32-class accuracy (dropped terms are O(2^-16)
// relative per product); both precisions serve the convolutions of the ResNet trunk (models/resnet_custom.py:35-55).
//
// Structure (one persistent CTA per SM -- or a CTA pair per 256-row tile, cta_group::2 --, K in 64-wide blocks):
//   warp 0      TMA producer: B_hi/B_lo tiles and the A plane(s): 2-D boxes of a matrix, or 4-D NHWC boxes (A_CONV)
//   warp 1      MMA issuer: one elected thread issues tcgen05.mma (UMMA 128*CG x <=256 x 16)
//   warp 2      TMEM allocator
//   warps 4..   epilogue, 4 per set (TMEM lane quarters): 2 sets for plane-fed kernels, 4 for the wide fp16 tiles, 1-2
//               for fp32-fed ones: tcgen05.ld accumulator -> registers -> bias / residual / activation (or the dgrad
//               extras, or the gate) -> the warp's private 64B-swizzled smem staging tile -> the warp's own TMA
//               stores; no CTA-wide barrier, TMEM handed back right after the warp's last tcgen05.ld
//   last 8      (A_F32 only) A converter: fp32 rows from global -> (hi,lo) bf16 pairs written straight into the
//               128B-swizzled UMMA operand layout, loads one K block ahead
// Pipelines: smem ring (full/empty mbarriers) between producer/converter and MMA, and a TMEM accumulator ring over
// all 512 columns (tmem_full/tmem_empty) between MMA and epilogue, so the epilogue of tile i overlaps the main loop
// of the following tiles.
#pragma once
#include "common.cuh"
#include <cuda_fp16.h>
#include <cstdlib>

namespace toad {
namespace tc {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;  // 64 bf16 = 128 B = one SWIZZLE_128B row
constexpr int UMMA_K = 16;
constexpr int A_TILE_BYTES = BLOCK_M * BLOCK_K * 2;

enum { EPI_LINEAR = 0, EPI_GATE = 1, EPI_DGRAD = 2 };
// EPI_DGRAD: the linear epilogue's staging / store path with the backward's extras (pooling term, ReLU mask, scale)
// instead of the forward's (bias, residual, ReLU, dropout): two instantiations keep each one's live registers low.
__host__ __device__ constexpr bool epi_is_linear(int epi) { return epi == EPI_LINEAR || epi == EPI_DGRAD; }
enum { A_F32 = 0, A_SPLIT = 1, A_CONV = 2, A_MN = 3 };
// Operand precision of one GEMM:
//   PREC_BF16X3  A ~= A_hi + A_lo, B ~= B_hi + B_lo as bf16 pairs, 3 passes (A_hi.B_hi + A_hi.B_lo + A_lo.B_hi): fp32-class
//                accuracy (dropped terms O(2^-16)); activations cost 4 B / element in HBM.  The TOAD head and the
//                "exact" ResNet mode.
//   PREC_F16X2   A is ONE fp16 plane (2 B / element), B ~= B_hi + B_lo as fp16 pairs, 2 passes (A.B_hi + A.B_lo): the
//                weights keep ~21 bits, the activations are rounded to fp16 (11 bits, the TF32 input precision cuDNN
//                uses for the reference's convolutions by default) where a layer stores them.  The ResNet trunk's
//                default: half the activation traffic, 2/3 of the tensor work.
enum { PREC_BF16X3 = 0, PREC_F16X2 = 1 };
// A_CONV: implicit-GEMM convolution over NHWC (hi,lo) planes.  A_MN: both operands MN-major -- A(m,k) and B(n,k)
// are read from planes stored [k, m] / [k, n] (the wgrad dW = dY^T . X with k = patch index: no transposes).

// CG = 1: one CTA per tile (UMMA M=128).  CG = 2: a CTA pair (cluster of 2, cta_group::2) shares one
// 256 x BLOCK_N tile: each CTA stages its own 128 rows of A and HALF of the B tile, the leader CTA
// issues UMMA M=256 reading both CTAs' shared memory, accumulators land in each CTA's own TMEM.
// Halves the per-SM weight traffic (L2->smem and smem->tensor core) and frees room for a 3rd stage.
template <int BLOCK_N, int CG = 1, int OUT_BUFS = 1, int PREC = PREC_BF16X3>
struct Cfg {
  static_assert(BLOCK_N == 64 || BLOCK_N == 128 || BLOCK_N == 256 || BLOCK_N == 512, "BLOCK_N");
  static_assert(CG == 1 || CG == 2, "CG");
  static_assert(BLOCK_N < 512 || CG == 2, "512-wide tiles are CTA-pair only");
  // BLOCK_N = 512 ("wide"): two UMMA N=256 halves per K step sharing ONE staged A tile, so a row tile's A
  // operand is loaded / converted once for all 512 output columns.  The accumulator then fills TMEM
  // (512 columns): a single accumulator buffer, the epilogue no longer overlaps the next main loop.
  static constexpr int N_SUB = BLOCK_N == 512 ? 2 : 1;
  static constexpr int UMMA_N = BLOCK_N / N_SUB;
  // accumulator ring: all 512 TMEM columns (one CTA per SM owns them anyway).  Narrow tiles with a short K loop
  // (the 1x1 convolutions: one K block per tile) are otherwise paced by the MMA -> epilogue -> MMA round trip.
  static constexpr int ACC_STAGES = 512 / BLOCK_N > 8 ? 8 : 512 / BLOCK_N;
  static constexpr int B_ROWS = BLOCK_N / CG;        // B rows staged by one CTA (N_SUB sub-tiles of B_SUB_ROWS)
  static constexpr int B_SUB_ROWS = UMMA_N / CG;     // rows of one sub-tile staged by one CTA
  static constexpr int B_TILE_BYTES = B_ROWS * BLOCK_K * 2;
  static constexpr int A_PLANES = PREC == PREC_F16X2 ? 1 : 2;
  static constexpr int B_OFF = A_PLANES * A_TILE_BYTES;  // stage layout: A plane(s) | B_hi | B_lo
  static constexpr int STAGE_BYTES = A_PLANES * A_TILE_BYTES + 2 * B_TILE_BYTES;
  // OUT_BUFS x 32 KB of staging for the epilogue's TMA stores, carved into per-warp private (hi, lo) tiles of
  // 32 rows x 32 columns (4 KB): 4 epilogue warps get 2 * OUT_BUFS tiles each, 8 warps OUT_BUFS each.  With one
  // tile a warp's store must finish reading smem before its next chunk is staged; with two it overlaps.
  // OUT_BUFS = 2 costs one operand stage: for store-/epilogue-bound shapes (ResNet).
  static_assert(OUT_BUFS == 1 || OUT_BUFS == 2, "OUT_BUFS");
  // (fp16 mode keeps the bias vector in 4 KB of static shared memory: leave room for it under the 227 KB limit)
  static constexpr int RING_BYTES = 192 * 1024 - (OUT_BUFS - 1) * 32 * 1024 - (PREC == PREC_F16X2 ? 8 * 1024 : 0);
  static constexpr int STAGES = RING_BYTES / STAGE_BYTES > 6 ? 6 : RING_BYTES / STAGE_BYTES;
  static_assert(STAGES >= 2, "need at least a double-buffered operand ring");
  static constexpr int TMEM_COLS = ACC_STAGES * BLOCK_N;  // accumulator ring (128/256/512: powers of 2)
  static constexpr int OUT_STAGE_BYTES = 2 * BLOCK_M * 128;  // (hi, lo) 128 x 64 bf16 staging tiles for the TMA store
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024;  // + slack for 1024 B alignment
  static constexpr int SMEM_BYTES_LINEAR = SMEM_BYTES + OUT_BUFS * OUT_STAGE_BYTES;
  static_assert(SMEM_BYTES_LINEAR + (PREC == PREC_F16X2 ? 4096 : 0) + 1024 <= 227 * 1024, "dynamic + static shared memory");
};

struct GemmTcParams {
  // A operand when A_MODE == A_F32 (lda = row stride in floats); for plane-fed modes lda / ldb are the
  // row strides of the (hi, lo) planes in elements (0 = K; must be multiples of 8)
  const float* a_f32;
  int64_t lda;
  int64_t ldb;
  int64_t M;
  int32_t N;
  int32_t K;
  // EPI_LINEAR
  const float* bias;  // [N] or nullptr
  int32_t relu;
  float* out_f32;  // nullable
  int64_t ld_f32;
  __nv_bfloat16* out_hi;  // nullable (both or neither)
  __nv_bfloat16* out_lo;
  int64_t ld_split;
  // EPI_GATE: tile columns [0,BLOCK_N/2) hold the tanh branch, [BLOCK_N/2,BLOCK_N) the sigmoid
  // branch of gate columns j0 = n_tile*BLOCK_N/2 ...
  const float* gate_ba;
  const float* gate_bb;
  const float* gate_wc;  // [ntasks, D]
  int32_t gate_D;
  int32_t gate_ntasks;  // 1..4
  float* gate_part;     // [sets * n_tiles][M][ntasks] partial scores (no bias): one per (tile, epilogue warp set);
                        // sets = 2 for plane-fed A, 1 for fp32-fed A
  int64_t gate_part_rows;  // rows of one partial plane (0 = M): a row slice of a larger bag writes into the bag's planes
  float* gate_a;        // nullable [M, D]
  float* gate_b;        // nullable [M, D]
  // A_CONV: rows are output pixels (b, oh, ow) raster; K blocks walk (tap, 64-channel chunk); the A tile
  // of one K block is one 4-D TMA box {64 ch, wb, hb, bb} of the NHWC input, shifted by the tap, with
  // out-of-image pixels zero-filled by TMA (= the convolution's zero padding).
  int32_t conv_kw;       // taps per filter row (1 or 3)
  int32_t conv_cchunks;  // C_in / 64
  int32_t conv_stride;   // 1 or 2 (encoded in the tensor map's elementStrides)
  int32_t conv_pad;      // 0 or 1
  int32_t conv_Wo;
  int32_t conv_Ho;
  // M tiling of an A_CONV GEMM: every 128-row M sub-tile is one box {64 ch, wb = Wo, hb, bb} of output pixels -- hb
  // full rows of one image (bb = 1, conv_tpi tiles per image, the last one clipped at Ho) or bb whole images (hb = Ho).
  // Only the first conv_tile_rows = wb*hb*bb (<= 128) rows of a tile are real; of those rows_valid (a prefix) exist.
  // Output rows stay the dense raster [B*Ho*Wo, N]: no power-of-two restriction on the image size.
  int32_t conv_B;
  int32_t conv_hb;
  int32_t conv_bb;
  int32_t conv_tpi;
  int32_t conv_tile_rows;
  int32_t conv_mtiles;
  // EPI_LINEAR residual: out = act(acc + bias + (res_hi + res_lo)), planes [M, ld_res]
  const __nv_bfloat16* res_hi;
  const __nv_bfloat16* res_lo;
  int64_t ld_res;
  // training-mode dropout on the epilogue's activation: EPI_LINEAR uses drop_layer with element index
  // row*N + col; EPI_GATE uses DROP_A / DROP_B with index row*D + gate column.
  DropoutCfg drop;
  uint32_t drop_layer;
  int64_t drop_row0;      // global index of row 0 (a row slice of a bag keeps the bag-wide dropout mask)
  int32_t pdl;            // launch programmatically dependent on the preceding kernel of the stream (see pdl_enabled)
  // split-K (TMA-fed A only): the K range is cut into k_splits slices of kb_per_split K blocks; slice s
  // writes its fp32 partial tile to out_f32 + s * M * ld_f32 (bias only in slice 0); a reduction follows.
  int32_t k_splits;       // 0 or 1 = no split
  int32_t kb_per_split;
  // dgrad extras (plane-fed EPI_LINEAR): out = (acc + pool_p[m][0]*pool_v[n] + pool_p[m][1]*pool_v[N+n])
  //                                              * (mask_f32[m, n] > 0) * out_scale
  const float* pool_p;    // [M, 2] or nullptr
  const float* pool_v;    // [2, N]
  const float* mask_f32;  // [M, ld_mask] or nullptr
  const __nv_bfloat16* mask_bf16;  // the same mask given as the hi plane of the activation (hi > 0 <=> value > 0), or nullptr
  int64_t ld_mask;
  float out_scale;        // 0 = 1
  // fp16 plane output written into a zero-bordered ("padded") plane for conv3x3_halo.cuh: raster row r of the
  // [B*H*W, N] result lands at position pad_pos(r, H, W) of a [pad_positions(B, H, W), N] plane (W % 32 == 0).  0 = off.
  int32_t out_pad_H, out_pad_W;
};

// zero-bordered activation plane of conv3x3_halo.cuh: padded position of raster row r = (b*H + h)*W + w
__host__ __device__ __forceinline__ int64_t pad_pos(int64_t r, int H, int W) {
  const int64_t bh = r / W;
  const int w = static_cast<int>(r - bh * W);
  const int64_t b = bh / H;
  const int h = static_cast<int>(bh - b * H);
  return (b * (H + 1) + h + 1) * (W + 1) + w + 1;
}
// positions of a padded plane holding B images (one trailing border row and one more cell: every tap of every pixel
// stays inside the matrix)
inline int64_t pad_positions(int64_t B, int H, int W) { return (B * (H + 1) + 1) * (W + 1) + 1; }

// ---------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// arrive on the barrier at the same offset in CTA `rank` of this cluster.  Default semantics (.release.cta): the
// explicit .release.cluster form lowers to MEMBAR.ALL.GPU + ERRBAR, which also waits for the caller's in-flight
// global prefetch loads (7 % of the fc1 converter warps' samples in profiles/r1m).  What crosses the CTA boundary
// here is ordered by other means: smem operand writes by fence.proxy.async, TMEM reads by tcgen05.wait::ld +
// tcgen05.fence::before_thread_sync.
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar, uint32_t rank) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}"
      ::"r"(bar), "r"(rank)
      : "memory");
}
__device__ __forceinline__ uint32_t mapa_cluster(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// Bounded spin: a protocol bug traps (launch error) instead of hanging the device.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  long long t0 = 0;
  for (uint32_t spin = 0;; ++spin) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    if (ok) break;
    if (spin == 1024) t0 = clock64();
    if (spin > 1024 && (spin & 1023) == 0 && clock64() - t0 > 4000000000LL) __trap();
  }
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire;" ::: "memory");
}

template <int CG>
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  if (CG == 1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
  } else {  // data to this CTA's smem, transaction bytes to the (cluster-mapped) leader barrier `bar`
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
  }
}
// smem tile -> global through the tensor map (rows/cols beyond the tensor are clipped by the hardware)
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(map), "r"(src), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int PENDING>
__device__ __forceinline__ void tma_store_wait_read() {  // all but the PENDING most recent store groups have read smem
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(PENDING) : "memory");
}
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
template <int CG>
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2,
                                            int c3) {
  if (CG == 1) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
  } else {
    asm volatile(
        "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
  }
}

// packed fp32x2 add (Blackwell FADD2): (a0, a1) += (b0, b1)
__device__ __forceinline__ void add_f32x2(uint32_t& a0, uint32_t& a1, float b0, float b1) {
  unsigned long long x, y;
  asm("mov.b64 %0, {%1, %2};" : "=l"(x) : "r"(a0), "r"(a1));
  asm("mov.b64 %0, {%1, %2};" : "=l"(y) : "f"(b0), "f"(b1));
  asm("add.rn.f32x2 %0, %0, %1;" : "+l"(x) : "l"(y));
  asm("mov.b64 {%0, %1}, %2;" : "=r"(a0), "=r"(a1) : "l"(x));
}
// mixed-precision adds (FHADD): a0 += fp16(low half of w), a1 += fp16(high half of w), exact fp32 results
__device__ __forceinline__ void add_f16x2_to_f32(uint32_t& a0, uint32_t& a1, uint32_t w) {
  asm("{\n\t.reg .b16 l, h;\n\t.reg .f32 x, y;\n\t"
      "mov.b32 {l, h}, %2;\n\tmov.b32 x, %0;\n\tmov.b32 y, %1;\n\t"
      "add.rn.f32.f16 x, l, x;\n\tadd.rn.f32.f16 y, h, y;\n\t"
      "mov.b32 %0, x;\n\tmov.b32 %1, y;\n\t}"
      : "+r"(a0), "+r"(a1) : "r"(w));
}
// two fp32 bit patterns -> packed fp16x2 (v0 in the low half) with ReLU and finite saturation in the conversion
__device__ __forceinline__ uint32_t pack_relu_f16x2(uint32_t v0, uint32_t v1) {
  uint32_t r;
  asm("cvt.rn.relu.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(__uint_as_float(v1)), "f"(__uint_as_float(v0)));
  return r;
}
// Programmatic dependent launch (PDL).  Every kernel of a chain runs `wait` (all memory operations of the preceding
// grid are complete and visible) BEFORE `launch_dependents`, and touches no kernel-written global memory before the
// wait: its successor can then start its prologue (barrier init, TMEM allocation, descriptor prefetch) on SMs the
// last wave of this grid has left, but -- by induction -- never runs ahead of a grid it depends on transitively.
// Without the launch attribute both instructions are no-ops.
__device__ __forceinline__ void pdl_wait_then_release() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

template <int CG>
__device__ __forceinline__ void tmem_alloc(uint32_t slot_smem, uint32_t ncols) {
  if (CG == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot_smem), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  } else {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot_smem), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
}
template <int CG>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  if (CG == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
  else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem] (+)= A[smem desc] . B[smem desc]; bf16 inputs, fp32 accumulate.
template <int CG>
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  if (CG == 1) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
  }
}
// mbarrier arrive (on every CTA of the pair for CG == 2) once all previously issued tcgen05.mma of this
// thread have completed.
template <int CG>
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  if (CG == 1) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
  } else {
    const uint16_t mask = 3;
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"(mask)
                 : "memory");
  }
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// UMMA shared-memory matrix descriptor: K-major operand, SWIZZLE_128B, rows of 128 B,
// 8-row groups 1024 B apart (SBO), descriptor version 1 (Blackwell).
__device__ __forceinline__ uint64_t make_kmajor_sw128_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr >> 4) & 0x3FFF);  // start address        bits [0,14)
  d |= static_cast<uint64_t>(1) << 16;                     // leading byte offset  bits [16,30) (unused for SW128 K-major)
  d |= static_cast<uint64_t>(1024 >> 4) << 32;             // stride byte offset   bits [32,46)
  d |= static_cast<uint64_t>(1) << 46;                     // version = 1          bits [46,48)
  d |= static_cast<uint64_t>(2) << 61;                     // SWIZZLE_128B         bits [61,64)
  return d;
}
// MN-major operand, SWIZZLE_128B: the tile is a row of [64 k-rows x 64 mn-elements (128 B)] boxes (one TMA box
// each, 8 KB); inside a box 8-row groups are 1024 B apart (SBO), boxes along MN are 8192 B apart (LBO).
__device__ __forceinline__ uint64_t make_mnmajor_sw128_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr >> 4) & 0x3FFF);
  d |= static_cast<uint64_t>(8192 >> 4) << 16;  // leading byte offset: next 64-element block along MN
  d |= static_cast<uint64_t>(1024 >> 4) << 32;  // stride byte offset: next 8 k-rows
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}
// Instruction descriptor: D=f32, A=B=bf16, M = 128*CG, N = BLOCK_N; mn_major sets both operands MN-major.
__host__ __device__ constexpr uint32_t make_idesc_bf16(int m, int n, bool mn_major = false) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (mn_major ? ((1u << 15) | (1u << 16)) : 0u) |
         (static_cast<uint32_t>(n >> 3) << 17) | (static_cast<uint32_t>(m >> 4) << 24);
}
// the same with fp16 operands (a_format = b_format = 0), K-major
__host__ __device__ constexpr uint32_t make_idesc_f16(int m, int n) {
  return (1u << 4) | (static_cast<uint32_t>(n >> 3) << 17) | (static_cast<uint32_t>(m >> 4) << 24);
}

// sigmoid / tanh on the SFU (ex2.approx + rcp.approx): abs error ~2e-7, far below the score tolerance.
__device__ __forceinline__ float fast_sigmoid(float z) {
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-1.4426950408889634f * z));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
  return r;
}
__device__ __forceinline__ float fast_tanh(float z) { return fmaf(2.0f, fast_sigmoid(2.0f * z), -1.0f); }

// ---------------------------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------------------------
struct ConvTile { int b0, oh0, rows_valid; int64_t out_row0; };
__device__ __forceinline__ ConvTile conv_tile(const GemmTcParams& p, int t) {
  ConvTile c;
  if (p.conv_tpi == 1) {  // bb whole images per tile
    c.b0 = t * p.conv_bb; c.oh0 = 0;
    int nb = p.conv_B - c.b0;
    nb = nb < 0 ? 0 : (nb > p.conv_bb ? p.conv_bb : nb);
    c.rows_valid = nb * p.conv_Ho * p.conv_Wo;
  } else {                // hb rows of one image
    c.b0 = t / p.conv_tpi; c.oh0 = (t - c.b0 * p.conv_tpi) * p.conv_hb;
    int nh = p.conv_Ho - c.oh0;
    nh = nh > p.conv_hb ? p.conv_hb : nh;
    c.rows_valid = c.b0 < p.conv_B ? nh * p.conv_Wo : 0;
  }
  c.out_row0 = (static_cast<int64_t>(c.b0) * p.conv_Ho + c.oh0) * p.conv_Wo;
  return c;
}
__device__ __forceinline__ float2 h2_to_f2(uint32_t w) {
  const __half2 h = *reinterpret_cast<const __half2*>(&w);
  return __half22float2(h);
}
// two floats -> packed fp16x2 (v0 in the low half), round to nearest, finite saturation (no inf in the activations)
__device__ __forceinline__ uint32_t pack_f16x2(float v0, float v1) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(v1), "f"(v0));
  return r;
}

constexpr int GATE_SMEM_FLOATS = 4 * 1024;  // ba | bb | wc rows (<= 2 tasks staged) for D <= 1024

// Epilogue warp sets (x4 warps, one per TMEM lane quarter).  Plane-fed kernels run 2 sets (384 threads).  The
// fp32-fed kernels also carry 8 converter warps: 1 set = 512 threads / 128 registers each; 2 sets (linear epilogue
// only) = 640 threads, where setmaxnreg moves registers from the 4 control warps to the converters.
#ifndef TOAD_F32_LINEAR_EPI_SETS
#define TOAD_F32_LINEAR_EPI_SETS 2
#endif
// The fp16 (ResNet) kernels with tiles of >= 128 columns run FOUR sets (16 epilogue warps, 640 threads): their short-K
// 1x1 convolutions are paced by the epilogue's instruction stream (ncu profiles/r2o: issue slots 30 % busy with two
// epilogue warps per scheduler, no memory unit above 50 %), so the cure is more warps in flight, not fewer bytes.
template <int A_MODE, int EPI, int BLOCK_N = 256, int PREC = PREC_BF16X3>
__host__ __device__ constexpr int epi_sets() {
  return A_MODE != A_F32 ? ((PREC == PREC_F16X2 && BLOCK_N >= 128) ? 4 : 2) : (epi_is_linear(EPI) ? TOAD_F32_LINEAR_EPI_SETS : 1);
}
template <int A_MODE, int EPI, int BLOCK_N = 256, int PREC = PREC_BF16X3>
__host__ __device__ constexpr int cta_threads() {
  return A_MODE == A_F32 ? (4 + 4 * epi_sets<A_MODE, EPI>() + 8) * 32 : (4 + 4 * epi_sets<A_MODE, EPI, BLOCK_N, PREC>()) * 32;
}
template <int REGS>
__device__ __forceinline__ void setmaxnreg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(REGS)); }
template <int REGS>
__device__ __forceinline__ void setmaxnreg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(REGS)); }

template <int BLOCK_N, int A_MODE, int EPI, int CG, int OUT_BUFS = 1, int PREC = PREC_BF16X3>
__global__ void __launch_bounds__(cta_threads<A_MODE, EPI, BLOCK_N, PREC>(), 1)
gemm_bf16x3_kernel(const __grid_constant__ CUtensorMap tm_a_hi, const __grid_constant__ CUtensorMap tm_a_lo,
                   const __grid_constant__ CUtensorMap tm_b_hi, const __grid_constant__ CUtensorMap tm_b_lo,
                   const __grid_constant__ CUtensorMap tm_o_hi, const __grid_constant__ CUtensorMap tm_o_lo,
                   const GemmTcParams p) {
  using C = Cfg<BLOCK_N, CG, OUT_BUFS, PREC>;
  static_assert(EPI != EPI_GATE || BLOCK_N <= 256, "the gate epilogue pairs two 128-column halves of a 256-wide tile");
  static_assert(PREC == PREC_BF16X3 || ((A_MODE == A_SPLIT || A_MODE == A_CONV) && EPI == EPI_LINEAR),
                "the fp16 single-plane mode exists for the plane-fed linear / convolution GEMMs");
  // epilogue warp sets: plane-fed kernels run 8 epilogue warps (two per TMEM lane quarter, each taking one
  // 32-column half of every 64-column chunk); the fp32-fed kernels keep 4 so that, with their 8 converter
  // warps, the CTA stays at 512 threads / 128 registers.
  constexpr int EPI_SETS = epi_sets<A_MODE, EPI, BLOCK_N, PREC>();
  constexpr int CONV_WARP0 = 4 + 4 * EPI_SETS;
  constexpr int STAGES = C::STAGES;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_full_b[STAGES];
  __shared__ __align__(8) uint64_t bar_full_a[STAGES];
  __shared__ __align__(8) uint64_t bar_empty[STAGES];
  __shared__ __align__(8) uint64_t bar_tmem_full[C::ACC_STAGES];
  __shared__ __align__(8) uint64_t bar_tmem_empty[C::ACC_STAGES];
  // fp16 mode: the whole bias vector staged once per CTA (the epilogue reads it with broadcast 128-bit loads)
  constexpr int BIAS_SMEM = PREC == PREC_F16X2 ? 1024 : 1;
  __shared__ __align__(16) float s_bias[BIAS_SMEM];
  __shared__ uint32_t tmem_slot;
  __shared__ float s_gate[EPI == EPI_GATE ? GATE_SMEM_FLOATS : 1];

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t cta_rank = CG == 2 ? cluster_ctarank() : 0u;
  const bool is_leader = cta_rank == 0;
  const uint32_t tiles_base = (smem_u32(smem_raw) + 1023u) & ~1023u;

  // work units: (128*CG) x BLOCK_N tiles, n fastest; this CTA owns rows [m0, m0+128) of its unit
  const int m_units = A_MODE == A_CONV ? (p.conv_mtiles + CG - 1) / CG
                                       : static_cast<int>((p.M + BLOCK_M * CG - 1) / (BLOCK_M * CG));
  const int n_tiles = p.N / BLOCK_N;
  const int mn_tiles = m_units * n_tiles;
  const int k_splits = (A_MODE != A_F32 && p.k_splits > 1) ? p.k_splits : 1;
  const int num_tiles = mn_tiles * k_splits;  // tile = split * mn_tiles + (m_unit * n_tiles + n_tile)
  const int num_kb = (p.K + BLOCK_K - 1) / BLOCK_K;  // TMA zero-fills a K tail (A_F32 callers keep K % 64 == 0)
  const int kb_per = k_splits > 1 ? p.kb_per_split : num_kb;
  const int unit0 = blockIdx.x / CG;
  const int unit_stride = gridDim.x / CG;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(smem_u32(&bar_full_b[s]), CG);      // producer of each CTA of the pair
      mbar_init(smem_u32(&bar_full_a[s]), 8 * CG);  // one arrive per converter warp (8 per CTA)
      mbar_init(smem_u32(&bar_empty[s]), 1);
    }
    for (int a = 0; a < C::ACC_STAGES; ++a) {
      mbar_init(smem_u32(&bar_tmem_full[a]), 1);
      mbar_init(smem_u32(&bar_tmem_empty[a]), 4 * EPI_SETS * CG);  // one arrive per epilogue warp
    }
    fence_barrier_init();
  }
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tm_b_hi);
    prefetch_tmap(&tm_b_lo);
    if (A_MODE != A_F32) {
      prefetch_tmap(&tm_a_hi);
      if (C::A_PLANES == 2) prefetch_tmap(&tm_a_lo);
    }
  }
  pdl_wait_then_release();
  if (PREC == PREC_F16X2) {
    for (int i = threadIdx.x; i < p.N && i < BIAS_SMEM; i += blockDim.x) s_bias[i] = p.bias != nullptr ? __ldg(p.bias + i) : 0.f;
  }
  if (EPI == EPI_GATE) {  // stage the gate biases and (up to 2) score rows once per CTA
    const int D = p.gate_D;
    for (int i = threadIdx.x; i < D; i += blockDim.x) {
      s_gate[i] = __ldg(p.gate_ba + i);
      s_gate[1024 + i] = __ldg(p.gate_bb + i);
      s_gate[2048 + i] = __ldg(p.gate_wc + i);
      s_gate[3072 + i] = p.gate_ntasks > 1 ? __ldg(p.gate_wc + D + i) : 0.f;
    }
  }
  if (warp == 2) tmem_alloc<CG>(smem_u32(&tmem_slot), C::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  if (CG == 2) cluster_sync_all();  // peer barriers initialised before any remote arrive / multicast
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;

#ifdef TOAD_SETMAXNREG
  if (A_MODE == A_F32 && EPI_SETS == 2) {  // 640 threads start at 96 registers: 4 control warps -> 40, 8 converter warps -> 120
    if (warp < 4) setmaxnreg_dec<40>();
    else if (warp >= CONV_WARP0) setmaxnreg_inc<120>();
  }
#endif
  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (all lanes loop, lane 0 issues)
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = unit0; tile < num_tiles; tile += unit_stride) {
      const int split = tile / mn_tiles, mn = tile - split * mn_tiles;
      const int kb0 = split * kb_per, kb1 = (kb0 + kb_per) < num_kb ? (kb0 + kb_per) : num_kb;
      const int n0 = (mn % n_tiles) * BLOCK_N + static_cast<int>(cta_rank) * C::B_SUB_ROWS;
      const int m0 = (mn / n_tiles) * (BLOCK_M * CG) + static_cast<int>(cta_rank) * BLOCK_M;
      // per-tile geometry once (the K loop below is paced by this one warp: no divisions inside it)
      ConvTile ct{};
      int tap_kh = 0, tap_kw = 0, tap_cc = 0;
      if (A_MODE == A_CONV) {
        ct = conv_tile(p, (mn / n_tiles) * CG + static_cast<int>(cta_rank));
        const int tap = kb0 / p.conv_cchunks;
        tap_cc = kb0 - tap * p.conv_cchunks;
        tap_kh = tap / p.conv_kw;
        tap_kw = tap - tap_kh * p.conv_kw;
      }
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(smem_u32(&bar_empty[stage]), phase ^ 1);
        if (lane == 0) {
          const uint32_t sa = tiles_base + stage * C::STAGE_BYTES;
          const uint32_t fb_local = smem_u32(&bar_full_b[stage]);
          // (a convolution's A box holds conv_tile_rows <= 128 pixels; out-of-image parts are zero-filled AND counted)
          const uint32_t a_bytes = A_MODE == A_F32 ? 0u : (A_MODE == A_CONV ? static_cast<uint32_t>(p.conv_tile_rows) * 128u
                                                                            : static_cast<uint32_t>(A_TILE_BYTES));
          const uint32_t bytes = 2 * C::B_TILE_BYTES + C::A_PLANES * a_bytes;
          uint32_t fb = fb_local;
          if (CG == 1) {
            mbar_expect_tx(fb_local, bytes);
          } else {
            fb = mapa_cluster(fb_local, 0);  // transaction bytes of both CTAs go to the leader's barrier
            if (is_leader) mbar_expect_tx(fb_local, 2 * bytes);
            else mbar_arrive_cluster(fb_local, 0);
          }
          if (A_MODE == A_SPLIT) {
            tma_load_2d<CG>(sa, &tm_a_hi, fb, kb * BLOCK_K, m0);
            if (C::A_PLANES == 2) tma_load_2d<CG>(sa + A_TILE_BYTES, &tm_a_lo, fb, kb * BLOCK_K, m0);
          } else if (A_MODE == A_MN) {
#pragma unroll
            for (int j = 0; j < BLOCK_M / 64; ++j) {  // [64 patches x 64 m] boxes
              tma_load_2d<CG>(sa + j * 8192, &tm_a_hi, fb, m0 + j * 64, kb * BLOCK_K);
              tma_load_2d<CG>(sa + A_TILE_BYTES + j * 8192, &tm_a_lo, fb, m0 + j * 64, kb * BLOCK_K);
            }
          } else if (A_MODE == A_CONV) {
            const int cw = tap_kw - p.conv_pad;  // (tiles span the full output width)
            const int ch = ct.oh0 * p.conv_stride + tap_kh - p.conv_pad;
            tma_load_4d<CG>(sa, &tm_a_hi, fb, tap_cc * BLOCK_K, cw, ch, ct.b0);
            if (C::A_PLANES == 2) tma_load_4d<CG>(sa + A_TILE_BYTES, &tm_a_lo, fb, tap_cc * BLOCK_K, cw, ch, ct.b0);
          }
          if (A_MODE == A_MN) {
#pragma unroll
            for (int j = 0; j < C::B_SUB_ROWS / 64; ++j) {  // [64 patches x 64 n] boxes of this CTA's share of B
              tma_load_2d<CG>(sa + C::B_OFF + j * 8192, &tm_b_hi, fb, n0 + j * 64, kb * BLOCK_K);
              tma_load_2d<CG>(sa + C::B_OFF + C::B_TILE_BYTES + j * 8192, &tm_b_lo, fb, n0 + j * 64, kb * BLOCK_K);
            }
          } else {
#pragma unroll
            for (int hs = 0; hs < C::N_SUB; ++hs) {  // sub-tile hs = rows [n0 + hs*UMMA_N, +B_SUB_ROWS) of this CTA's share
              const uint32_t so = hs * (C::B_SUB_ROWS * 128);
              tma_load_2d<CG>(sa + C::B_OFF + so, &tm_b_hi, fb, kb * BLOCK_K, n0 + hs * C::UMMA_N);
              tma_load_2d<CG>(sa + C::B_OFF + C::B_TILE_BYTES + so, &tm_b_lo, fb, kb * BLOCK_K, n0 + hs * C::UMMA_N);
            }
          }
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
        if (A_MODE == A_CONV) {  // next K block = next (tap, channel chunk), all lanes in step
          if (++tap_cc == p.conv_cchunks) {
            tap_cc = 0;
            if (++tap_kw == p.conv_kw) { tap_kw = 0; ++tap_kh; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (leader CTA only)
    if (is_leader) {
      constexpr uint32_t idesc = PREC == PREC_F16X2 ? make_idesc_f16(BLOCK_M * CG, C::UMMA_N)
                                                    : make_idesc_bf16(BLOCK_M * CG, C::UMMA_N, A_MODE == A_MN);
      // descriptor start-address step per UMMA K (=16): K-major +32 B inside the swizzled row; MN-major +16 k-rows
      constexpr uint64_t KSTEP = A_MODE == A_MN ? ((16 * 128) >> 4) : ((UMMA_K * 2) >> 4);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = unit0; tile < num_tiles; tile += unit_stride, ++it) {
        const int acc = it % C::ACC_STAGES;
        mbar_wait(smem_u32(&bar_tmem_empty[acc]), ((it / C::ACC_STAGES) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_tmem0 = tmem_base + acc * BLOCK_N;
        const int split = tile / mn_tiles;
        const int kb0 = split * kb_per, kb1 = (kb0 + kb_per) < num_kb ? (kb0 + kb_per) : num_kb;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(smem_u32(&bar_full_b[stage]), phase);
          if (A_MODE == A_F32) mbar_wait(smem_u32(&bar_full_a[stage]), phase);
          tc_fence_after();
          if (lane == 0) {
            const uint32_t sa = tiles_base + stage * C::STAGE_BYTES;
            const uint64_t a_hi = A_MODE == A_MN ? make_mnmajor_sw128_desc(sa) : make_kmajor_sw128_desc(sa);
            const uint64_t a_lo = A_MODE == A_MN ? make_mnmajor_sw128_desc(sa + A_TILE_BYTES) : make_kmajor_sw128_desc(sa + A_TILE_BYTES);
#pragma unroll
            for (int hs = 0; hs < C::N_SUB; ++hs) {
              const uint32_t so = hs * (C::B_SUB_ROWS * 128);
              const uint64_t b_hi = A_MODE == A_MN ? make_mnmajor_sw128_desc(sa + C::B_OFF + so)
                                                   : make_kmajor_sw128_desc(sa + C::B_OFF + so);
              const uint64_t b_lo = A_MODE == A_MN ? make_mnmajor_sw128_desc(sa + C::B_OFF + C::B_TILE_BYTES + so)
                                                   : make_kmajor_sw128_desc(sa + C::B_OFF + C::B_TILE_BYTES + so);
              const uint32_t d_tmem = d_tmem0 + hs * C::UMMA_N;
#pragma unroll
              for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
                const uint64_t koff = static_cast<uint64_t>(k) * KSTEP;
                umma_bf16<CG>(d_tmem, a_hi + koff, b_hi + koff, idesc, (kb > kb0) || (k != 0));
              }
#pragma unroll
              for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
                const uint64_t koff = static_cast<uint64_t>(k) * KSTEP;
                umma_bf16<CG>(d_tmem, a_hi + koff, b_lo + koff, idesc, 1);
              }
              if (PREC == PREC_BF16X3) {
#pragma unroll
                for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
                  const uint64_t koff = static_cast<uint64_t>(k) * KSTEP;
                  umma_bf16<CG>(d_tmem, a_lo + koff, b_hi + koff, idesc, 1);
                }
              }
            }
            umma_commit<CG>(smem_u32(&bar_empty[stage]));  // smem slot free (in both CTAs) once these MMAs retire
            if (kb == kb1 - 1) umma_commit<CG>(smem_u32(&bar_tmem_full[acc]));  // accumulator ready
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp >= 4 && warp < CONV_WARP0) {
    // ------------------------------------------------------------------ epilogue (4 * EPI_SETS warps)
    // warp -> (TMEM lane quarter ew = warp % 4, column half eh): the two warps of a lane quarter split every
    // 64-column chunk into its two 32-column halves.
    const int ew = (warp - 4) & 3;  // == warp % 4: the TMEM lane quarter this warp may access
    const int eh = (warp - 4) >> 2;
    int it = 0;
    uint32_t out_chunk = 0;  // running count of staged chunks (selects the staging buffer)
    for (int tile = unit0; tile < num_tiles; tile += unit_stride, ++it) {
      const int split = tile / mn_tiles, mn = tile - split * mn_tiles;
      const int n_tile = mn % n_tiles;
      const int n0 = n_tile * BLOCK_N;
      const int m0 = (mn / n_tiles) * (BLOCK_M * CG) + static_cast<int>(cta_rank) * BLOCK_M;
      const int acc = it % C::ACC_STAGES;
      // rows of this warp: a plain GEMM tile is 128 dense rows (rows >= M are clipped by the TMA store); a convolution
      // tile holds rows_valid <= 128 real output pixels starting at raster row out_row0 (see GemmTcParams::conv_*)
      int64_t tile_row0 = m0;
      int warp_rows = 32;
      if (A_MODE == A_CONV) {
        const ConvTile ct = conv_tile(p, (mn / n_tiles) * CG + static_cast<int>(cta_rank));
        tile_row0 = ct.out_row0;
        warp_rows = ct.rows_valid - ew * 32;
        warp_rows = warp_rows < 0 ? 0 : (warp_rows > 32 ? 32 : warp_rows);
      }
      const int64_t row = tile_row0 + ew * 32 + lane;
      const bool row_ok = A_MODE == A_CONV ? lane < warp_rows : row < p.M;
      // full warps go through the staged TMA store; a convolution warp cut by the end of its tile stores its valid
      // rows straight from registers (a TMA box cannot be clipped inside the tensor)
      const bool use_tma = A_MODE != A_CONV || warp_rows == 32;
      // residual operand (ResNet shortcut): this warp's NEXT 32-column chunk is fetched one chunk ahead (and the
      // first one before waiting for the accumulator), hiding the global latency that otherwise serialises the
      // epilogue of the short-K 1x1 convolutions.
      const bool has_res = EPI == EPI_LINEAR && A_MODE != A_F32 && p.res_hi != nullptr;  // (compile-time false for EPI_DGRAD)
      uint4 res_h[4], res_l[4];
      auto load_res = [&](int c_) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          res_h[q] = make_uint4(0u, 0u, 0u, 0u);
          res_l[q] = make_uint4(0u, 0u, 0u, 0u);
          if (has_res && row_ok) {
            const int64_t o = row * p.ld_res + n0 + c_ * 32 + q * 8;
            res_h[q] = *reinterpret_cast<const uint4*>(p.res_hi + o);
            if (PREC == PREC_BF16X3) res_l[q] = *reinterpret_cast<const uint4*>(p.res_lo + o);
          }
        }
      };
      // dgrad ReLU mask (the saved activation's bf16 hi plane, 64 B per row and chunk): fetched one chunk ahead too
      const bool has_mask16 = EPI == EPI_DGRAD && A_MODE != A_F32 && p.mask_bf16 != nullptr;
      uint4 msk[4];
      auto load_mask = [&](int c_) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          msk[q] = make_uint4(0x3f803f80u, 0x3f803f80u, 0x3f803f80u, 0x3f803f80u);  // bf16 1.0 pairs: keep
          if (has_mask16 && row_ok) msk[q] = *reinterpret_cast<const uint4*>(p.mask_bf16 + row * p.ld_mask + n0 + c_ * 32 + q * 8);
        }
      };
      // dgrad pooling term p[m][0] * v[0][n] + p[m][1] * v[1][n]: the row's two weights once per tile, the chunk's two
      // 32-column slices of v held one value per lane (fetched a chunk ahead, broadcast by shuffles like the bias)
      const bool has_pool = EPI == EPI_DGRAD && A_MODE != A_F32 && p.pool_p != nullptr;
      float pool_p0 = 0.f, pool_p1 = 0.f, pv0_nxt = 0.f, pv1_nxt = 0.f;
      auto load_pool = [&](int c_) {
        if (has_pool) {
          pv0_nxt = __ldg(p.pool_v + n0 + c_ * 32 + lane);
          pv1_nxt = __ldg(p.pool_v + p.N + n0 + c_ * 32 + lane);
        }
      };
      if (has_pool && row_ok) { pool_p0 = __ldg(p.pool_p + row * 2); pool_p1 = __ldg(p.pool_p + row * 2 + 1); }
      // bias of a 32-column chunk: lane j holds bias[col0 + j] (one coalesced load, fetched one chunk ahead);
      // the value of column i is broadcast with a shuffle where it is added.
      const bool has_bias = EPI == EPI_LINEAR && PREC == PREC_BF16X3 && p.bias != nullptr && split == 0;  // (fp16 mode: s_bias)
      auto load_bias = [&](int c_) { return has_bias ? __ldg(p.bias + n0 + c_ * 32 + lane) : 0.f; };
      float bias_nxt = 0.f;
      if (EPI == EPI_LINEAR) {
        bias_nxt = load_bias(eh);
        if (A_MODE != A_F32) load_res(eh);
      }
      if (EPI == EPI_DGRAD && A_MODE != A_F32) { load_mask(eh); load_pool(eh); }
      mbar_wait(smem_u32(&bar_tmem_full[acc]), (it / C::ACC_STAGES) & 1);
      tc_fence_after();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + acc * BLOCK_N;

      if (epi_is_linear(EPI)) {
        // 32 columns at a time, every warp on its own: TMEM -> registers -> bias/residual/ReLU -> (hi,lo) bf16 ->
        // the warp's PRIVATE 64B-swizzled staging tile (32 rows x 64 B per plane) -> one TMA store per plane
        // (rows >= M clipped by TMA).  No CTA-wide barrier: the warps drift, so one warp's TMEM / store latency
        // is covered by the others' arithmetic; with WARP_BUFS = 2 a warp's own store overlaps its next chunk.
        constexpr int N_CHUNKS = BLOCK_N / 32;
        // (the fp16 mode stages ONE 2 KB plane per chunk in the same 4 KB slots: twice the buffers)
        constexpr int WARP_STAGE_BYTES = OUT_BUFS * 8192 / EPI_SETS;   // OUT_BUFS x 32 KB over 4 * EPI_SETS warps
        constexpr int WARP_BUFS = WARP_STAGE_BYTES / (PREC == PREC_F16X2 ? 2048 : 4096);
        static_assert(WARP_BUFS >= 1, "staging");
        const uint32_t my_stage = tiles_base + STAGES * C::STAGE_BYTES + (eh * 4 + ew) * WARP_STAGE_BYTES;
        // (fp16 mode: the <= 4 chunks of a warp fully unrolled, so the one-chunk-ahead residual prefetch rotates through
        // renamed registers instead of 16 moves per chunk)
        constexpr int CHUNK_UNROLL = PREC == PREC_F16X2 ? 4 : 1;
#pragma unroll CHUNK_UNROLL
        for (int c = eh; c < N_CHUNKS; c += EPI_SETS) {
          const bool last = c + EPI_SETS >= N_CHUNKS;
          uint32_t r[32];
          tmem_ld32(t_row + c * 32, r);
          const float bias_cur = bias_nxt;
          uint4 cur_h[4], cur_l[4], cur_m[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) { cur_h[q] = res_h[q]; cur_l[q] = res_l[q]; cur_m[q] = msk[q]; }
          const float pv0_cur = pv0_nxt, pv1_cur = pv1_nxt;
          // fp16 mode: the first half of the chunk's bias (4 broadcast 128-bit loads, batched: their shared-memory
          // latency -- long while the tensor pipe and the TMA own the shared-memory port -- overlaps the TMEM load);
          // the next chunk's residual is fetched AFTER this chunk's has been consumed (same registers, no copies).
          uint4 bias_a[4];
          if (PREC == PREC_F16X2) {
            const uint32_t sb = smem_u32(s_bias + n0 + c * 32);
#pragma unroll
            for (int q = 0; q < 4; ++q)
              asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(bias_a[q].x), "=r"(bias_a[q].y), "=r"(bias_a[q].z), "=r"(bias_a[q].w)
                           : "r"(sb + q * 16));
          }
          if (!last && PREC != PREC_F16X2) {  // next chunk's bias / residual (forward) or mask / pooling vectors (dgrad), in flight below
            if (EPI == EPI_LINEAR) {
              bias_nxt = load_bias(c + EPI_SETS);
              if (A_MODE != A_F32) load_res(c + EPI_SETS);
            } else if (A_MODE != A_F32) {
              load_mask(c + EPI_SETS);
              load_pool(c + EPI_SETS);
            }
          }
          tmem_ld_wait();
          if (last) {  // this warp has drained its share of the accumulator: hand TMEM back before the arithmetic
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
              if (CG == 1) mbar_arrive(smem_u32(&bar_tmem_empty[acc]));
              else mbar_arrive_cluster(smem_u32(&bar_tmem_empty[acc]), 0);
            }
          }
          const int col0 = n0 + c * 32;
          if (EPI == EPI_LINEAR && PREC == PREC_BF16X3) {
#pragma unroll
            for (int i = 0; i < 32; ++i)
              r[i] = __float_as_uint(__uint_as_float(r[i]) + __shfl_sync(0xffffffffu, bias_cur, i));
          }
          if (PREC == PREC_F16X2) {
            // lean fp16-mode arithmetic: bias from shared memory (8 broadcast 128-bit loads) with packed fp32x2 adds,
            // the residual's fp16 halves added straight into the fp32 values (FHADD); ReLU rides on the final conversion
            uint4 bias_b[4];
            {
              const uint32_t sb = smem_u32(s_bias + col0 + 16);
#pragma unroll
              for (int q = 0; q < 4; ++q)
                asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(bias_b[q].x), "=r"(bias_b[q].y), "=r"(bias_b[q].z), "=r"(bias_b[q].w)
                             : "r"(sb + q * 16));
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              add_f32x2(r[4 * q], r[4 * q + 1], __uint_as_float(bias_a[q].x), __uint_as_float(bias_a[q].y));
              add_f32x2(r[4 * q + 2], r[4 * q + 3], __uint_as_float(bias_a[q].z), __uint_as_float(bias_a[q].w));
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              add_f32x2(r[16 + 4 * q], r[16 + 4 * q + 1], __uint_as_float(bias_b[q].x), __uint_as_float(bias_b[q].y));
              add_f32x2(r[16 + 4 * q + 2], r[16 + 4 * q + 3], __uint_as_float(bias_b[q].z), __uint_as_float(bias_b[q].w));
            }
            if (has_res) {
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const uint4 vh = res_h[q];
                add_f16x2_to_f32(r[q * 8], r[q * 8 + 1], vh.x);
                add_f16x2_to_f32(r[q * 8 + 2], r[q * 8 + 3], vh.y);
                add_f16x2_to_f32(r[q * 8 + 4], r[q * 8 + 5], vh.z);
                add_f16x2_to_f32(r[q * 8 + 6], r[q * 8 + 7], vh.w);
              }
            }
            if (!last) load_res(c + EPI_SETS);   // lands during this chunk's stores and the next chunk's TMEM wait
          }
          if (has_res && PREC == PREC_BF16X3) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const uint4 vh = cur_h[q], vl = cur_l[q];
              const uint32_t uh[4] = {vh.x, vh.y, vh.z, vh.w}, ul[4] = {vl.x, vl.y, vl.z, vl.w};
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                const uint32_t wh = uh[e >> 1], wl = ul[e >> 1];
                const float rv = (e & 1) ? (bf16hi_to_f32(wh) + bf16hi_to_f32(wl)) : (bf16lo_to_f32(wh) + bf16lo_to_f32(wl));
                r[q * 8 + e] = __float_as_uint(__uint_as_float(r[q * 8 + e]) + rv);
              }
            }
          }
          if (EPI == EPI_LINEAR && PREC == PREC_BF16X3 && p.relu) {
#pragma unroll
            for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(fmaxf(__uint_as_float(r[i]), 0.0f));
          }
          if (EPI == EPI_DGRAD && A_MODE != A_F32) {
            // backward dgrad epilogue: + pooling term, ReLU mask of the saved activation, 1/keep
            if (has_pool) {
#pragma unroll
              for (int i = 0; i < 32; ++i) {
                const float t = fmaf(pool_p0, __shfl_sync(0xffffffffu, pv0_cur, i), __uint_as_float(r[i]));
                r[i] = __float_as_uint(fmaf(pool_p1, __shfl_sync(0xffffffffu, pv1_cur, i), t));
              }
            }
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              float4 mk = make_float4(1.f, 1.f, 1.f, 1.f);
              if (p.mask_f32 != nullptr && row_ok)
                mk = *reinterpret_cast<const float4*>(p.mask_f32 + row * p.ld_mask + col0 + q * 4);
              if (p.mask_bf16 != nullptr) {  // (prefetched; rows past M carry the keep pattern and are clipped by TMA)
                const uint4 m4 = cur_m[q >> 1];
                const uint32_t mx = (q & 1) ? m4.z : m4.x, my = (q & 1) ? m4.w : m4.y;
                mk = make_float4(bf16lo_to_f32(mx), bf16hi_to_f32(mx), bf16lo_to_f32(my), bf16hi_to_f32(my));
              }
              const float mv[4] = {mk.x, mk.y, mk.z, mk.w};
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const int i = q * 4 + e;
                float t = __uint_as_float(r[i]);
                t = mv[e] > 0.f ? t : 0.f;
                if (p.out_scale != 0.f) t *= p.out_scale;
                r[i] = __float_as_uint(t);
              }
            }
          }
          if (EPI == EPI_LINEAR && PREC == PREC_BF16X3 && p.drop.thresh != 0u) {  // training only: one big uniform branch, never predicated into the hot path
            const unsigned long long e0 = static_cast<unsigned long long>(row + p.drop_row0) * p.N + col0;
#pragma unroll
            for (int i = 0; i < 32; ++i)
              r[i] = __float_as_uint(dropout_apply(p.drop, p.drop_layer, e0 + i, __uint_as_float(r[i])));
          }
          if (PREC == PREC_BF16X3 && row_ok && p.out_f32 != nullptr) {
            float* dst = p.out_f32 + (static_cast<int64_t>(split) * p.M + row) * p.ld_f32 + col0;  // 32 B aligned (ld % 8 == 0)
#pragma unroll
            for (int i = 0; i < 4; ++i) st_global_v8(dst + 8 * i, r, 8 * i);
          }
          if (p.out_hi != nullptr && PREC == PREC_F16X2) {
            // one fp16 plane: 32 rows x 64 B per chunk
            const uint32_t buf = my_stage + (out_chunk % WARP_BUFS) * 2048;
            ++out_chunk;
            if (use_tma) {  // staging buffer free again?  (this warp's store that last used it must have finished READING it)
              if (lane == 0) tma_store_wait_read<WARP_BUFS - 1>();
              __syncwarp();
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {  // 8 columns -> one 16 B chunk
              uint32_t h[4];
              if (p.relu) {
#pragma unroll
                for (int e = 0; e < 4; ++e) h[e] = pack_relu_f16x2(r[8 * q + 2 * e], r[8 * q + 2 * e + 1]);
              } else {
#pragma unroll
                for (int e = 0; e < 4; ++e) h[e] = pack_f16x2(__uint_as_float(r[8 * q + 2 * e]), __uint_as_float(r[8 * q + 2 * e + 1]));
              }
              if (use_tma) {
                const uint32_t off = lane * 64 + ((q ^ ((lane >> 1) & 3)) << 4);  // SWIZZLE_64B: chunk ^= row bits [1,2]
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(buf + off), "r"(h[0]), "r"(h[1]), "r"(h[2]), "r"(h[3]) : "memory");
              } else if (row_ok) {
                *reinterpret_cast<uint4*>(p.out_hi + row * p.ld_split + col0 + q * 8) = make_uint4(h[0], h[1], h[2], h[3]);
              }
            }
            if (use_tma) {
              fence_proxy_async_smem();
              __syncwarp();
              if (lane == 0) {
                int r0 = static_cast<int>(tile_row0) + ew * 32;
                bool inside = true;
                if (p.out_pad_W != 0) {  // 32 rows of one image row -> 32 consecutive padded positions (no hardware clipping)
                  inside = r0 < p.M;
                  r0 = static_cast<int>(pad_pos(r0, p.out_pad_H, p.out_pad_W));
                }
                if (inside) tma_store_2d(&tm_o_hi, buf, col0, r0);
                tma_store_commit();
              }
            }
          }
          if (p.out_hi != nullptr && PREC == PREC_BF16X3) {
            const uint32_t buf_hi = my_stage + (out_chunk % WARP_BUFS) * 4096, buf_lo = buf_hi + 2048;
            ++out_chunk;
            // staging buffer free again?  (this warp's store that last used it must have finished READING it)
            if (use_tma) {
              if (lane == 0) tma_store_wait_read<WARP_BUFS - 1>();
              __syncwarp();
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {  // 8 columns -> one 16 B chunk per plane
              uint32_t hi[4], lo[4];
#pragma unroll
              for (int e = 0; e < 4; ++e)
                split2(__uint_as_float(r[8 * q + 2 * e]), __uint_as_float(r[8 * q + 2 * e + 1]), hi[e], lo[e]);
              if (use_tma) {
                const uint32_t off = lane * 64 + ((q ^ ((lane >> 1) & 3)) << 4);  // SWIZZLE_64B: chunk ^= row bits [1,2]
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(buf_hi + off), "r"(hi[0]), "r"(hi[1]),
                             "r"(hi[2]), "r"(hi[3]) : "memory");
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(buf_lo + off), "r"(lo[0]), "r"(lo[1]),
                             "r"(lo[2]), "r"(lo[3]) : "memory");
              } else if (row_ok) {
                *reinterpret_cast<uint4*>(p.out_hi + row * p.ld_split + col0 + q * 8) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                *reinterpret_cast<uint4*>(p.out_lo + row * p.ld_split + col0 + q * 8) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
              }
            }
            if (use_tma) {
              fence_proxy_async_smem();
              __syncwarp();
              if (lane == 0) {
                tma_store_2d(&tm_o_hi, buf_hi, col0, static_cast<int>(tile_row0) + ew * 32);
                tma_store_2d(&tm_o_lo, buf_lo, col0, static_cast<int>(tile_row0) + ew * 32);
                tma_store_commit();
              }
            }
          }
        }
      } else {  // EPI_GATE
        constexpr int HALF = BLOCK_N / 2;
        const int j0 = n_tile * HALF;
        float s[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 1
        for (int c = eh; c < HALF / 32; c += EPI_SETS) {  // the warps of a lane quarter alternate 32-column chunks
          uint32_t ra[32], rb[32];
          tmem_ld32(t_row + c * 32, ra);
          tmem_ld32(t_row + HALF + c * 32, rb);
          tmem_ld_wait();
          const int jc = j0 + c * 32;
          const bool save = row_ok && p.gate_a != nullptr;
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            ra[i] = __float_as_uint(fast_tanh(__uint_as_float(ra[i]) + s_gate[jc + i]));
            rb[i] = __float_as_uint(fast_sigmoid(__uint_as_float(rb[i]) + s_gate[1024 + jc + i]));
          }
          if (p.drop.thresh != 0u) {  // training only: one big uniform branch (see EPI_LINEAR)
            const unsigned long long e0 = static_cast<unsigned long long>(row + p.drop_row0) * p.gate_D + jc;
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              ra[i] = __float_as_uint(dropout_apply(p.drop, DROP_A, e0 + i, __uint_as_float(ra[i])));
              rb[i] = __float_as_uint(dropout_apply(p.drop, DROP_B, e0 + i, __uint_as_float(rb[i])));
            }
          }
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const float g = __uint_as_float(ra[i]) * __uint_as_float(rb[i]);
            s[0] = fmaf(g, s_gate[2048 + jc + i], s[0]);
            s[1] = fmaf(g, s_gate[3072 + jc + i], s[1]);
            if (p.gate_ntasks > 2) {  // rare: extra tasks read their score rows from global
              s[2] = fmaf(g, __ldg(p.gate_wc + 2 * p.gate_D + jc + i), s[2]);
              if (p.gate_ntasks > 3) s[3] = fmaf(g, __ldg(p.gate_wc + 3 * p.gate_D + jc + i), s[3]);
            }
          }
          if (save) {
            float* da = p.gate_a + row * p.gate_D + jc;  // 32 B aligned: gate_D % 8 == 0, jc % 32 == 0
            float* db = p.gate_b + row * p.gate_D + jc;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              st_global_v8(da + 8 * i, ra, 8 * i);
              st_global_v8(db + 8 * i, rb, 8 * i);
            }
          }
        }
        if (row_ok) {
          const int64_t plane_rows = p.gate_part_rows > 0 ? p.gate_part_rows : p.M;
          float* dst = p.gate_part + (static_cast<int64_t>(n_tile * EPI_SETS + eh) * plane_rows + row) * p.gate_ntasks;
#pragma unroll
          for (int t = 0; t < 4; ++t)
            if (t < p.gate_ntasks) dst[t] = s[t];
        }
      }
      if (!epi_is_linear(EPI)) {  // (the linear epilogue hands TMEM back right after its last tcgen05.ld)
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (CG == 1) mbar_arrive(smem_u32(&bar_tmem_empty[acc]));
          else mbar_arrive_cluster(smem_u32(&bar_tmem_empty[acc]), 0);
        }
      }
    }
    if (epi_is_linear(EPI) && lane == 0) tma_store_wait_all();  // this warp's own bulk stores
  } else if (A_MODE == A_F32 && warp >= CONV_WARP0) {
    // ------------------------------------------------------------------ A converter (8 warps)
    // Each half-warp streams one 256 B row segment (64 fp32) per load instruction; a thread turns its
    // 4 floats into 4 (hi) + 4 (lo) bf16 = one 8 B store into each swizzled tile.
    const int cw = warp - CONV_WARP0;  // 0..7: rows cw*16 .. cw*16+15 of the tile
    const int hw = lane >> 4, l16 = lane & 15;
    const int my_tiles = unit0 < num_tiles ? (num_tiles - unit0 + unit_stride - 1) / unit_stride : 0;
    const int64_t total = static_cast<int64_t>(my_tiles) * num_kb;
    float4 buf0[8], buf1[8];
    auto issue = [&](int64_t g, float4(&buf)[8]) {
      const int tile = unit0 + static_cast<int>(g / num_kb) * unit_stride;
      const int kb = static_cast<int>(g % num_kb);
      const int64_t m0 = static_cast<int64_t>(tile / n_tiles) * (BLOCK_M * CG) + cta_rank * BLOCK_M;
      const float* base = p.a_f32 + kb * BLOCK_K + l16 * 4;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int64_t row = m0 + cw * 16 + j * 2 + hw;
        buf[j] = row < p.M ? ld_stream_f4(base + row * p.lda) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    };
    int stage = 0;
    uint32_t phase = 0;
    auto convert = [&](float4(&cur)[8]) {
      mbar_wait(smem_u32(&bar_empty[stage]), phase ^ 1);
      const uint32_t sa = tiles_base + stage * C::STAGE_BYTES;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int r = cw * 16 + j * 2 + hw;
        uint32_t h0, l0, h1, l1;
        split2(cur[j].x, cur[j].y, h0, l0);
        split2(cur[j].z, cur[j].w, h1, l1);
        const uint32_t off = r * 128 + (((l16 >> 1) ^ (r & 7)) << 4) + ((l16 & 1) << 3);
        asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(sa + off), "r"(h0), "r"(h1) : "memory");
        asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(sa + A_TILE_BYTES + off), "r"(l0), "r"(l1) : "memory");
      }
      fence_proxy_async_smem();  // make generic-proxy smem writes visible to the tensor core
      __syncwarp();
      if (lane == 0) {
        if (CG == 1) mbar_arrive(smem_u32(&bar_full_a[stage]));
        else mbar_arrive_cluster(smem_u32(&bar_full_a[stage]), 0);
      }
      if (++stage == STAGES) { stage = 0; phase ^= 1; }
    };
    // loads run one K block ahead of the conversion (two rotating register buffers)
    if (total > 0) issue(0, buf0);
    for (int64_t g = 0; g < total; g += 2) {
      if (g + 1 < total) issue(g + 1, buf1);
      convert(buf0);
      if (g + 1 >= total) break;
      if (g + 2 < total) issue(g + 2, buf0);
      convert(buf1);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (CG == 2) cluster_sync_all();  // the leader's MMAs read the peer's smem; nobody leaves early
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<CG>(tmem_base, C::TMEM_COLS);
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = []() -> PFN_encodeTiled {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess) return nullptr;
    if (q != cudaDriverEntryPointSuccess) return nullptr;
    return reinterpret_cast<PFN_encodeTiled>(f);
  }();
  return fn;
}

// Tensor map over a row-major bf16 matrix [rows, cols]; box = 64 columns x box_rows rows, 128B swizzle.
inline int make_bf16_tmap(CUtensorMap* map, const void* ptr, int64_t rows, int64_t cols, int box_rows,
                          int64_t ld = 0) {
  PFN_encodeTiled enc = get_encode_fn();
  if (enc == nullptr) return TOAD_ERR_DRIVER;
  if (ld == 0) ld = cols;
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(ld) * 2};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(BLOCK_K), static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : TOAD_ERR_DRIVER;
}

// Tensor map for the epilogue's per-warp TMA stores: box = 32 columns (64 B) x 32 rows, 64B swizzle.
inline int make_bf16_out_tmap(CUtensorMap* map, const void* ptr, int64_t rows, int64_t cols, int64_t ld) {
  PFN_encodeTiled enc = get_encode_fn();
  if (enc == nullptr) return TOAD_ERR_DRIVER;
  if (ld == 0) ld = cols;
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(ld) * 2};
  cuuint32_t box[2] = {32, 32};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : TOAD_ERR_DRIVER;
}

// Programmatic dependent launch is OPT-IN (TOAD_B200_PDL=1 in the environment).  Measured on B200 (profiles/
// r2h, r2i): on ONE stream it hides the launch gap + prologue of the next kernel (+3 % slides/s at N = 10k, +-0 at
// 50k); with slides in flight on SEVERAL streams the early-launched CTAs sit on SMs (214 KB of shared memory each)
// waiting for their predecessor while another stream's runnable kernel needs those SMs: 3-stream throughput fell from
// ~3500 to 2100 slides/s.  The library cannot know how many streams its caller uses, so the default is off for the
// TOAD head; the ResNet trunk (45 short dependent launches on one stream) sets GemmTcParams::pdl itself.
inline bool pdl_enabled() {
  static bool v = []() {
    const char* e = getenv("TOAD_B200_PDL");
    return e != nullptr && e[0] == '1';
  }();
  return v;
}

inline int sm_count() {
  static int n = []() {
    int dev = 0, v = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) return 148;
    return v;
  }();
  return n;
}

// 4-D tensor map over an NHWC bf16 plane [B, H, W, C]: box = 64 channels x (wb x hb x bb) pixels sampled with
// `stride` in W and H (boxDim = pixels * stride, elementStrides = stride), 128B swizzle, zero OOB fill.
inline int make_nhwc_tmap(CUtensorMap* map, const void* ptr, int64_t B, int64_t H, int64_t W, int64_t C, int wb, int hb,
                          int bb, int stride) {
  PFN_encodeTiled enc = get_encode_fn();
  if (enc == nullptr) return TOAD_ERR_DRIVER;
  cuuint64_t dims[4] = {static_cast<cuuint64_t>(C), static_cast<cuuint64_t>(W), static_cast<cuuint64_t>(H),
                        static_cast<cuuint64_t>(B)};
  cuuint64_t strides[3] = {static_cast<cuuint64_t>(C) * 2, static_cast<cuuint64_t>(W) * C * 2,
                           static_cast<cuuint64_t>(H) * W * C * 2};
  cuuint32_t box[4] = {static_cast<cuuint32_t>(BLOCK_K), static_cast<cuuint32_t>(wb * stride),
                       static_cast<cuuint32_t>(hb * stride), static_cast<cuuint32_t>(bb)};
  cuuint32_t estr[4] = {1, static_cast<cuuint32_t>(stride), static_cast<cuuint32_t>(stride), 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : TOAD_ERR_DRIVER;
}

// Launch with ready-made A tensor maps (unused for A_F32).  B operand: planes b_hi/b_lo [N, K] bf16.
// Launch with ready-made A and B tensor maps.
template <int BLOCK_N, int A_MODE, int EPI, int CG, int OUT_BUFS = 1, int PREC = PREC_BF16X3>
int launch_gemm_maps_b(const GemmTcParams& p, const CUtensorMap& ta_hi, const CUtensorMap& ta_lo,
                       const CUtensorMap& tb_hi, const CUtensorMap& tb_lo, cudaStream_t stream) {
  using C = Cfg<BLOCK_N, CG, OUT_BUFS, PREC>;
  if (p.M <= 0) return 0;
  if ((A_MODE == A_F32 && p.K % BLOCK_K != 0) || p.N % BLOCK_N != 0 || p.K <= 0 || p.N <= 0) return TOAD_ERR_UNSUPPORTED;
  if (p.k_splits > 1 && (A_MODE == A_F32 || EPI != EPI_LINEAR || p.out_hi != nullptr || p.kb_per_split <= 0)) return TOAD_ERR_ARG;
  if (EPI != EPI_DGRAD && (p.pool_p != nullptr || p.mask_f32 != nullptr || p.mask_bf16 != nullptr || p.out_scale != 0.f)) return TOAD_ERR_ARG;
  if (EPI == EPI_DGRAD && (p.bias != nullptr || p.res_hi != nullptr || p.relu != 0 || p.drop.thresh != 0u)) return TOAD_ERR_ARG;
  if (EPI == EPI_GATE && (p.gate_D > 1024 || p.gate_ntasks < 1 || p.gate_ntasks > 4)) return TOAD_ERR_UNSUPPORTED;
  if (PREC == PREC_F16X2 && (p.N > 1024 || p.out_f32 != nullptr || p.k_splits > 1 || p.drop.thresh != 0u)) return TOAD_ERR_UNSUPPORTED;
  // the epilogues' register -> global stores are 256-bit: 32-byte aligned rows
  if (epi_is_linear(EPI) && p.out_f32 != nullptr && ((reinterpret_cast<uintptr_t>(p.out_f32) & 31) != 0 || p.ld_f32 % 8 != 0))
    return TOAD_ERR_ARG;
  if (EPI == EPI_GATE && p.gate_a != nullptr &&
      (((reinterpret_cast<uintptr_t>(p.gate_a) | reinterpret_cast<uintptr_t>(p.gate_b)) & 31) != 0 || p.gate_D % 8 != 0))
    return TOAD_ERR_ARG;
  CUtensorMap to_hi = tb_hi, to_lo = tb_lo;
  if (epi_is_linear(EPI) && p.out_hi != nullptr) {
    if ((PREC == PREC_BF16X3 && p.out_lo == nullptr) || p.ld_split % 8 != 0) return TOAD_ERR_ARG;
    int64_t out_rows = p.M;
    if (p.out_pad_W != 0) {
      const int64_t img = static_cast<int64_t>(p.out_pad_H) * p.out_pad_W;
      if (PREC != PREC_F16X2 || A_MODE == A_CONV || p.out_pad_W % 32 != 0 || p.out_pad_H <= 0 || p.M % img != 0) return TOAD_ERR_ARG;
      out_rows = pad_positions(p.M / img, p.out_pad_H, p.out_pad_W);
    }
    TOAD_TRY(make_bf16_out_tmap(&to_hi, p.out_hi, out_rows, p.N, p.ld_split));
    if (PREC == PREC_BF16X3) TOAD_TRY(make_bf16_out_tmap(&to_lo, p.out_lo, p.M, p.N, p.ld_split));
  }
  constexpr int kSmem = epi_is_linear(EPI) ? C::SMEM_BYTES_LINEAR : C::SMEM_BYTES;
  auto kern = gemm_bf16x3_kernel<BLOCK_N, A_MODE, EPI, CG, OUT_BUFS, PREC>;
  {  // once per instantiation and device (one process drives one GPU; a repeated call would only cost host time)
    static int attr_dev = -1;
    int dev = 0;
    TOAD_CUDA_TRY(cudaGetDevice(&dev));
    if (attr_dev != dev) {
      TOAD_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
      attr_dev = dev;
    }
  }
  const int64_t m_units = A_MODE == A_CONV ? (p.conv_mtiles + CG - 1) / CG : (p.M + BLOCK_M * CG - 1) / (BLOCK_M * CG);
  const int64_t units = m_units * (p.N / BLOCK_N) * (p.k_splits > 1 ? p.k_splits : 1);
  const int64_t max_units = sm_count() / CG;
  const int grid = static_cast<int>(units < max_units ? units : max_units) * CG;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(static_cast<unsigned>(grid));
  cfg.blockDim = dim3(cta_threads<A_MODE, EPI, BLOCK_N, PREC>());
  cfg.dynamicSmemBytes = kSmem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (CG > 1) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = CG;
    attr[na].val.clusterDim.y = 1;
    attr[na].val.clusterDim.z = 1;
    ++na;
  }
  if (pdl_enabled() || p.pdl != 0) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  TOAD_CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, ta_hi, ta_lo, tb_hi, tb_lo, to_hi, to_lo, p));
  TOAD_CUDA_TRY(cudaGetLastError());
  return 0;
}

// Launch with ready-made A tensor maps (unused for A_F32).  B operand: planes b_hi/b_lo [N, K] bf16.
template <int BLOCK_N, int A_MODE, int EPI, int CG, int OUT_BUFS = 1, int PREC = PREC_BF16X3>
int launch_gemm_maps(const GemmTcParams& p, const CUtensorMap& ta_hi, const CUtensorMap& ta_lo,
                     const __nv_bfloat16* b_hi, const __nv_bfloat16* b_lo, cudaStream_t stream) {
  using C = Cfg<BLOCK_N, CG, OUT_BUFS, PREC>;
  if (p.M <= 0) return 0;
  if (p.K <= 0 || p.N <= 0) return TOAD_ERR_UNSUPPORTED;
  CUtensorMap tb_hi, tb_lo;
  const int64_t ldb = p.ldb > 0 ? p.ldb : p.K;
  TOAD_TRY(make_bf16_tmap(&tb_hi, b_hi, p.N, p.K, C::B_SUB_ROWS, ldb));
  TOAD_TRY(make_bf16_tmap(&tb_lo, b_lo, p.N, p.K, C::B_SUB_ROWS, ldb));
  return launch_gemm_maps_b<BLOCK_N, A_MODE, EPI, CG, OUT_BUFS, PREC>(p, ta_hi, ta_lo, tb_hi, tb_lo, stream);
}

// A operand: fp32 (a_f32/lda in p) when A_MODE == A_F32, else the planes a_hi/a_lo [M, K] bf16.
template <int BLOCK_N, int A_MODE, int EPI, int CG = 1, int OUT_BUFS = 1, int PREC = PREC_BF16X3>
int launch_gemm(const GemmTcParams& p, const __nv_bfloat16* a_hi, const __nv_bfloat16* a_lo,
                const __nv_bfloat16* b_hi, const __nv_bfloat16* b_lo, cudaStream_t stream) {
  static_assert(A_MODE != A_CONV && A_MODE != A_MN, "use launch_conv_gemm / launch_gemm_mn");
  if (p.M <= 0) return 0;
  CUtensorMap ta_hi, ta_lo;
  if (A_MODE == A_SPLIT) {
    if (p.K <= 0) return TOAD_ERR_UNSUPPORTED;
    const int64_t lda = p.lda > 0 ? p.lda : p.K;  // plane row stride in elements (multiple of 8)
    TOAD_TRY(make_bf16_tmap(&ta_hi, a_hi, p.M, p.K, BLOCK_M, lda));
    if (PREC == PREC_BF16X3) TOAD_TRY(make_bf16_tmap(&ta_lo, a_lo, p.M, p.K, BLOCK_M, lda));
    else ta_lo = ta_hi;
  } else {
    TOAD_TRY(make_bf16_tmap(&ta_hi, b_hi, p.N, p.K, 64));  // placeholders, never dereferenced
    ta_lo = ta_hi;
  }
  return launch_gemm_maps<BLOCK_N, A_MODE, EPI, CG, OUT_BUFS, PREC>(p, ta_hi, ta_lo, b_hi, b_lo, stream);
}

// wgrad-style GEMM with both operands MN-major: C[M, N] = sum_k A(m,k) B(n,k) with A stored as planes [K, M]
// (row stride lda) and B as planes [K, N] (row stride ldb) -- i.e. dW = dY^T . X straight from the natural
// [patch, channel] layouts, K = patches.  Split-K / fp32 partial output as configured in p.
template <int BLOCK_N, int CG>
int launch_gemm_mn(GemmTcParams p, const __nv_bfloat16* a_hi, const __nv_bfloat16* a_lo, const __nv_bfloat16* b_hi,
                   const __nv_bfloat16* b_lo, cudaStream_t stream) {
  static_assert(BLOCK_N <= 256, "MN-major path uses one UMMA N tile");
  if (p.M <= 0 || p.K <= 0) return 0;
  if (p.M % 64 != 0 || p.N % BLOCK_N != 0) return TOAD_ERR_UNSUPPORTED;
  const int64_t lda = p.lda > 0 ? p.lda : p.M, ldb = p.ldb > 0 ? p.ldb : p.N;
  CUtensorMap ta_hi, ta_lo, tb_hi, tb_lo;
  // tensor maps over the [K, channels] planes: box = 64 channels (inner, 128 B) x 64 k-rows
  TOAD_TRY(make_bf16_tmap(&ta_hi, a_hi, p.K, p.M, 64, lda));
  TOAD_TRY(make_bf16_tmap(&ta_lo, a_lo, p.K, p.M, 64, lda));
  TOAD_TRY(make_bf16_tmap(&tb_hi, b_hi, p.K, p.N, 64, ldb));
  TOAD_TRY(make_bf16_tmap(&tb_lo, b_lo, p.K, p.N, 64, ldb));
  return launch_gemm_maps_b<BLOCK_N, A_MN, EPI_LINEAR, CG, 1>(p, ta_hi, ta_lo, tb_hi, tb_lo, stream);
}

// Convolution as implicit GEMM: input planes NHWC [B, H, W, Cin] ((hi, lo) bf16, or one fp16 plane for PREC_F16X2),
// weights [Cout, taps*Cin] with K order (kh, kw, cin), output planes [B*Ho*Wo, Cout].  ksize in {1, 3}; stride in
// {1, 2}; pad = ksize/2.  Any image size with Wo <= 128 (and even H, W for stride 2): see GemmTcParams::conv_*.
template <int BLOCK_N, int CG, int OUT_BUFS = 1, int PREC = PREC_BF16X3>
int launch_conv_gemm(GemmTcParams p, const __nv_bfloat16* in_hi, const __nv_bfloat16* in_lo, int B, int H, int W,
                     int Cin, int ksize, int stride, const __nv_bfloat16* w_hi, const __nv_bfloat16* w_lo,
                     cudaStream_t stream) {
  if (Cin % BLOCK_K != 0 || (ksize != 1 && ksize != 3) || (stride != 1 && stride != 2)) return TOAD_ERR_UNSUPPORTED;
  if (H % stride != 0 || W % stride != 0) return TOAD_ERR_UNSUPPORTED;
  const int Ho = H / stride, Wo = W / stride;
  if (Wo > BLOCK_M || Wo < 1 || Ho < 1) return TOAD_ERR_UNSUPPORTED;
  // <= 128 output pixels per M tile = wb x hb x bb: full rows first, then whole images
  const int wb = Wo;
  int hb, bb, tpi;
  if (Ho * Wo <= BLOCK_M) { hb = Ho; bb = BLOCK_M / (Ho * Wo); tpi = 1; }
  else { hb = BLOCK_M / Wo; bb = 1; tpi = (Ho + hb - 1) / hb; }
  if (hb * stride > 256 || wb * stride > 256) return TOAD_ERR_UNSUPPORTED;  // TMA box limits
  p.M = static_cast<int64_t>(B) * Ho * Wo;
  p.K = ksize * ksize * Cin;
  p.conv_kw = ksize;
  p.conv_cchunks = Cin / BLOCK_K;
  p.conv_stride = stride;
  p.conv_pad = ksize / 2;
  p.conv_Wo = Wo;
  p.conv_Ho = Ho;
  p.conv_B = B;
  p.conv_hb = hb;
  p.conv_bb = bb;
  p.conv_tpi = tpi;
  p.conv_tile_rows = wb * hb * bb;
  p.conv_mtiles = tpi == 1 ? (B + bb - 1) / bb : B * tpi;
  CUtensorMap ta_hi, ta_lo;
  TOAD_TRY(make_nhwc_tmap(&ta_hi, in_hi, B, H, W, Cin, wb, hb, bb, stride));
  if (PREC == PREC_BF16X3) TOAD_TRY(make_nhwc_tmap(&ta_lo, in_lo, B, H, W, Cin, wb, hb, bb, stride));
  else ta_lo = ta_hi;
  return launch_gemm_maps<BLOCK_N, A_CONV, EPI_LINEAR, CG, OUT_BUFS, PREC>(p, ta_hi, ta_lo, w_hi, w_lo, stream);
}

}  // namespace tc
}  // namespace toad
