// 3x3 / stride-1 convolution that loads every input pixel ONCE per tile ("halo reuse"), sm_100a, fp16 planes.
//
// The bottleneck blocks' conv2 (reference: models/resnet_custom.py:23-24,44-46: conv3x3 -> bn2 -> relu) of layer1
// (64 -> 64 channels at 64 x 64) and layer2 (128 -> 128 at 32 x 32) are short-N, long-K problems: as a tap-by-tap implicit
// GEMM (gemm_tc.cuh, A_CONV) they re-fetch the activation nine times and the weights once per tile from L2 and
// run two narrow UMMAs per K step (~1000 cycles per K block, tensor pipe 26 % / 50 % busy, profiles/r2y).  Here
//   * the WEIGHTS of the CTA's 64 output channels stay resident in shared memory for the whole kernel (a CTA pair
//     splits them: 32 rows x 9*Cin x 2 fp16 planes = 72 KB per CTA for Cin = 64, 144 KB for Cin = 128), and
//   * the ACTIVATION is read from a zero-bordered ("padded") plane in which a pixel's 3x3 neighbourhood sits at fixed
//     row offsets: position q = (b*(H+1) + h + 1)*(W+1) + (w + 1) of a [P, Cin] matrix (one shared zero row between
//     images, one shared zero column between rows).  A tile = 128 consecutive positions; its patch = positions
//     q0-(W+2) .. q0+127+(W+2) is loaded once by TMA (SWIZZLE_128B), and tap (kh, kw) is the SAME shared-memory
//     patch viewed from row kh*(W+1)+kw on: a UMMA descriptor whose start address is advanced by whole 128-byte rows.
//     (tcgen05 applies the 128B swizzle to absolute shared-memory address bits, so a start that is not 1024-byte
//     aligned needs nothing else: measured by tools/probe_umma_shift.cu, profiles/r2_umma_shift_probe.txt.)
// L2 -> SM traffic per 128 x 64 output tile drops from 9 x 24 KB to one 33 KB patch.
//
// The two fp16 weight planes (w ~= w_hi + w_lo) are not two passes over K but 64 more accumulator COLUMNS: a CTA's
// [32 hi rows | 32 lo rows] form one 64-row B operand, the pair issues 256 x 128 x 16 UMMAs, and the epilogue adds the
// two column halves.  A 256 x 64 UMMA re-reads its A operand from shared memory for half the math (measured: 76
// cycles per 256 x 64 x 16 instruction where the tensor pipe needs 32); folding the planes into N halves the A reads
// and the number of instructions.
//
// The 128 positions of a tile include the shared border cells (1 in W+1, plus one row in H+1): their results are
// computed and dropped; the epilogue writes only real pixels, into the ordinary unpadded NHWC plane.
//
// Roles (CTA pair, cta_group::2): warp 0 TMA producer (weights once, then one patch per tile and 64-channel block),
// warp 1 MMA issuer (leader CTA), warp 2 TMEM allocator, warps 4.. epilogue (sets of 4 alternate tiles).
#pragma once
#include "gemm_tc.cuh"

namespace toad {
namespace halo {

using namespace tc;

constexpr int ACC_STAGES = 4;
constexpr int ACC_COLS = 128;   // 64 output channels x (hi, lo) weight planes
constexpr int W_TILE_BYTES = 32 * 128;   // one CTA's 32 weight rows x 64 k of one plane

template <int KB>   // Cin = 64 * KB
struct HaloCfg {
  static constexpr int MAX_RB = KB == 1 ? 136 : 104;         // patch = 2 TMA boxes of RB rows (RB % 8 == 0)
  static constexpr int MAX_W = MAX_RB - 66;                   // 128 + 2 (W + 2) <= 2 RB
  static constexpr int EPI_SETS = KB == 1 ? 2 : 1;
  static constexpr int THREADS = 128 + 128 * EPI_SETS;
  static constexpr int W_BYTES = 9 * KB * 2 * W_TILE_BYTES;
  static constexpr int STAGE_BYTES = 2 * MAX_RB * 128;
  static constexpr int OUT_BYTES = EPI_SETS * 4 * 4096;
  static constexpr int SMEM_BYTES = W_BYTES + 2 * STAGE_BYTES + OUT_BYTES + 1024;
  static_assert(SMEM_BYTES <= 227 * 1024, "shared memory");
};

struct HaloParams {
  const float* bias;     // [Cout] folded BatchNorm bias
  __nv_bfloat16* out;    // [B*H*W, Cout] fp16 bits, NHWC
  int32_t B, H, W, Cout;
  int32_t n_groups;      // Cout / 64
  int32_t m_units;       // ceil(P / 256): pairs of 128-position tiles
  int32_t rb;            // rows per patch box
  int32_t relu;
};

template <int KB>
__global__ void __launch_bounds__(HaloCfg<KB>::THREADS, 1)
conv3x3_halo_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_w_hi,
                    const __grid_constant__ CUtensorMap tm_w_lo, const HaloParams p) {
  using C = HaloCfg<KB>;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_w;
  __shared__ __align__(8) uint64_t bar_full[2];
  __shared__ __align__(8) uint64_t bar_empty[2];
  __shared__ __align__(8) uint64_t bar_acc_full[ACC_STAGES];
  __shared__ __align__(8) uint64_t bar_acc_empty[ACC_STAGES];
  __shared__ __align__(16) float s_bias[64];
  __shared__ uint32_t tmem_slot;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t cta_rank = cluster_ctarank();
  const bool is_leader = cta_rank == 0;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sW = base, sA = base + C::W_BYTES, sOut = sA + 2 * C::STAGE_BYTES;
  const int pitch = p.W + 1;

  // a pair keeps ONE group of 64 output channels (its resident weights) and walks the position tiles
  const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
  const int group = pair % p.n_groups;
  const int unit0 = pair / p.n_groups, unit_stride = n_pairs / p.n_groups;

  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bar_w), 2);
    for (int s = 0; s < 2; ++s) {
      mbar_init(smem_u32(&bar_full[s]), 2);
      mbar_init(smem_u32(&bar_empty[s]), 1);
    }
    for (int a = 0; a < ACC_STAGES; ++a) {
      mbar_init(smem_u32(&bar_acc_full[a]), 1);
      mbar_init(smem_u32(&bar_acc_empty[a]), 4 * 2);  // the 4 warps of one epilogue set, in both CTAs
    }
    fence_barrier_init();
  }
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tm_a);
    prefetch_tmap(&tm_w_hi);
    prefetch_tmap(&tm_w_lo);
  }
  pdl_wait_then_release();
  if (threadIdx.x < 64) s_bias[threadIdx.x] = p.bias != nullptr ? __ldg(p.bias + group * 64 + threadIdx.x) : 0.f;
  if (warp == 2) tmem_alloc<2>(smem_u32(&tmem_slot), ACC_STAGES * ACC_COLS);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {  // resident weights: this CTA's 32 rows of the group, all 9*KB K blocks, both planes
      const uint32_t bw_local = smem_u32(&bar_w);
      const uint32_t bw = mapa_cluster(bw_local, 0);
      if (is_leader) mbar_expect_tx(bw_local, 2 * C::W_BYTES);
      else mbar_arrive_cluster(bw_local, 0);
      const int n0 = group * 64 + static_cast<int>(cta_rank) * 32;
      for (int t = 0; t < 9 * KB; ++t) {
        tma_load_2d<2>(sW + (2 * t) * W_TILE_BYTES, &tm_w_hi, bw, t * 64, n0);
        tma_load_2d<2>(sW + (2 * t + 1) * W_TILE_BYTES, &tm_w_lo, bw, t * 64, n0);
      }
    }
    __syncwarp();
    int stage = 0;
    uint32_t phase = 0;
    const uint32_t stage_tx = 2u * static_cast<uint32_t>(p.rb) * 128u;
    for (int unit = unit0; unit < p.m_units; unit += unit_stride) {
      const int q_first = (unit * 2 + static_cast<int>(cta_rank)) * 128 - (p.W + 2);   // position of patch row 0
      for (int kb = 0; kb < KB; ++kb) {
        mbar_wait(smem_u32(&bar_empty[stage]), phase ^ 1);
        if (lane == 0) {
          const uint32_t sa = sA + stage * C::STAGE_BYTES;
          const uint32_t fb_local = smem_u32(&bar_full[stage]);
          const uint32_t fb = mapa_cluster(fb_local, 0);
          if (is_leader) mbar_expect_tx(fb_local, 2 * stage_tx);
          else mbar_arrive_cluster(fb_local, 0);
          tma_load_2d<2>(sa, &tm_a, fb, kb * 64, q_first);                       // (rows outside [0, P) arrive as zeros)
          tma_load_2d<2>(sa + p.rb * 128, &tm_a, fb, kb * 64, q_first + p.rb);
        }
        __syncwarp();
        if (++stage == 2) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (leader CTA)
    if (is_leader) {
      constexpr uint32_t idesc = make_idesc_f16(256, ACC_COLS);
      mbar_wait(smem_u32(&bar_w), 0);
      int stage = 0, it = 0;
      uint32_t phase = 0;
      for (int unit = unit0; unit < p.m_units; unit += unit_stride, ++it) {
        const int acc = it % ACC_STAGES;
        mbar_wait(smem_u32(&bar_acc_empty[acc]), ((it / ACC_STAGES) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * ACC_COLS;
        for (int kb = 0; kb < KB; ++kb) {
          mbar_wait(smem_u32(&bar_full[stage]), phase);
          tc_fence_after();
          if (lane == 0) {
            const uint32_t sa = sA + stage * C::STAGE_BYTES;
#pragma unroll 1
            for (int kh = 0; kh < 3; ++kh) {
#pragma unroll
              for (int kw = 0; kw < 3; ++kw) {
                // tap (kh, kw) of position q is position q + (kh-1)*(W+1) + (kw-1): patch row kh*(W+1) + kw onwards
                const uint64_t a_d = make_kmajor_sw128_desc(sa + static_cast<uint32_t>(kh * pitch + kw) * 128u);
                const uint32_t wt = sW + static_cast<uint32_t>(2 * ((kh * 3 + kw) * KB + kb)) * W_TILE_BYTES;
                const uint64_t w_d = make_kmajor_sw128_desc(wt);   // 64 rows: this CTA's [hi | lo] tiles, contiguous
#pragma unroll
                for (int k = 0; k < 4; ++k)
                  umma_bf16<2>(d_tmem, a_d + static_cast<uint64_t>(2 * k), w_d + static_cast<uint64_t>(2 * k), idesc,
                               (kb | kh | kw | k) != 0);
              }
            }
            umma_commit<2>(smem_u32(&bar_empty[stage]));
            if (kb == KB - 1) umma_commit<2>(smem_u32(&bar_acc_full[acc]));
          }
          __syncwarp();
          if (++stage == 2) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ epilogue: set eh takes tiles it % EPI_SETS == eh
    const int ew = (warp - 4) & 3, eh = (warp - 4) >> 2;
    const uint32_t stg = sOut + static_cast<uint32_t>(warp - 4) * 4096u;
    int it = 0;
    for (int unit = unit0; unit < p.m_units; unit += unit_stride, ++it) {
      if (it % C::EPI_SETS != eh) continue;
      const int acc = it % ACC_STAGES;
      // this lane's position -> real pixel?  raster row of the output plane, or -1
      const int64_t q = static_cast<int64_t>(unit * 2 + static_cast<int>(cta_rank)) * 128 + ew * 32 + lane;
      int64_t dst_row = -1;
      {
        const int64_t rr = q / pitch;
        const int cc = static_cast<int>(q - rr * pitch);
        if (cc >= 1 && rr >= 1) {
          const int64_t b = (rr - 1) / (p.H + 1);
          const int h = static_cast<int>(rr - 1 - b * (p.H + 1));
          if (h < p.H && b < p.B) dst_row = (b * p.H + h) * p.W + (cc - 1);
        }
      }
      mbar_wait(smem_u32(&bar_acc_full[acc]), (it / ACC_STAGES) & 1);
      tc_fence_after();
      // accumulator columns: [CTA0: hi ch 0..31 | lo ch 0..31 | CTA1: hi ch 32..63 | lo ch 32..63] of the group
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + acc * ACC_COLS;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t v[32], vl[32];
        tmem_ld32(t_row + c * 64, v);
        tmem_ld32(t_row + c * 64 + 32, vl);
        tmem_ld_wait();
        if (c == 1) {  // accumulator drained by this warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(smem_u32(&bar_acc_empty[acc]), 0);
        }
        // hi + lo, bias, ReLU, fp16; the lane's 128-byte row parked in the warp's staging tile (16 B chunk index ^= row & 7)
        const float4* sb = reinterpret_cast<const float4*>(s_bias + c * 32);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 b4 = sb[j];
          add_f32x2(v[4 * j], v[4 * j + 1], __uint_as_float(vl[4 * j]) + b4.x, __uint_as_float(vl[4 * j + 1]) + b4.y);
          add_f32x2(v[4 * j + 2], v[4 * j + 3], __uint_as_float(vl[4 * j + 2]) + b4.z, __uint_as_float(vl[4 * j + 3]) + b4.w);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          uint32_t h[4];
#pragma unroll
          for (int e = 0; e < 4; ++e)
            h[e] = p.relu ? pack_relu_f16x2(v[8 * j + 2 * e], v[8 * j + 2 * e + 1])
                          : pack_f16x2(__uint_as_float(v[8 * j + 2 * e]), __uint_as_float(v[8 * j + 2 * e + 1]));
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(stg + lane * 128 + (((c * 4 + j) ^ (lane & 7)) << 4)),
                       "r"(h[0]), "r"(h[1]), "r"(h[2]), "r"(h[3]) : "memory");
        }
      }
      __syncwarp();
      // coalesced stores: each instruction writes 4 rows x 128 B; border positions are skipped
      const int sub = lane >> 3, ch = lane & 7;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int r = i * 4 + sub;
        const int64_t dr = __shfl_sync(0xffffffffu, dst_row, r);
        uint4 t;
        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(t.x), "=r"(t.y), "=r"(t.z), "=r"(t.w)
                     : "r"(stg + r * 128 + ((ch ^ (r & 7)) << 4)) : "memory");
        if (dr >= 0) *reinterpret_cast<uint4*>(p.out + dr * p.Cout + group * 64 + ch * 8) = t;
      }
      __syncwarp();
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 2) tmem_dealloc<2>(tmem_base, ACC_STAGES * ACC_COLS);
}

// zero the border cells of a padded plane [P, C] (the interior is written by the producing convolution): one warp per
// padded row -- a border row is cleared whole, of an image row only cell 0 (the column shared with the previous row);
// the last warp also clears the plane's trailing cell
__global__ void zero_borders_kernel(__nv_bfloat16* __restrict__ plane, int64_t n_rows, int H, int W, int C) {
  const int64_t rr = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (rr >= n_rows) return;
  const int lane = threadIdx.x & 31;
  const int cpr = C / 8;   // 16-byte chunks per cell
  const bool border_row = rr % (H + 1) == 0;
  int cells = border_row ? (W + 1) : 1;
  if (rr == n_rows - 1) ++cells;   // (the cell after the last row)
  uint4* row = reinterpret_cast<uint4*>(plane + rr * (W + 1) * C);
  for (int i = lane; i < cells * cpr; i += 32) row[i] = make_uint4(0u, 0u, 0u, 0u);
}
inline int launch_zero_borders(__nv_bfloat16* plane, int B, int H, int W, int C, cudaStream_t stream) {
  const int64_t n_rows = static_cast<int64_t>(B) * (H + 1) + 1;   // pad_positions = n_rows * (W + 1) + 1
  zero_borders_kernel<<<static_cast<unsigned>((n_rows + 7) / 8), 256, 0, stream>>>(plane, n_rows, H, W, C);
  TOAD_CUDA_TRY(cudaGetLastError());
  return 0;
}

template <int KB>
inline bool halo_supported(int H, int W, int Cin, int Cout) {
  return Cin == 64 * KB && Cout % 64 == 0 && W % 32 == 0 && W <= HaloCfg<KB>::MAX_W && H >= 1;
}

// out [B*H*W, Cout] = act(conv3x3(in) + bias): in = padded plane [pad_positions(B, H, W), Cin] with zero borders,
// weights (w_hi, w_lo) [Cout, 9*Cin] fp16 planes in K order (kh, kw, cin).
template <int KB>
int launch_conv3x3_halo(const __nv_bfloat16* in_padded, const __nv_bfloat16* w_hi, const __nv_bfloat16* w_lo, const float* bias,
                        __nv_bfloat16* out, int B, int H, int W, int Cout, bool relu, bool pdl, cudaStream_t stream) {
  using C = HaloCfg<KB>;
  constexpr int Cin = 64 * KB;
  if (B <= 0) return 0;
  if (!halo_supported<KB>(H, W, Cin, Cout)) return TOAD_ERR_UNSUPPORTED;
  const int64_t P = pad_positions(B, H, W);
  if (P > 0x7fffffff - 1024 || static_cast<int64_t>(B) * H * W > 0x7fffffff) return TOAD_ERR_UNSUPPORTED;
  HaloParams p{};
  p.bias = bias; p.out = out; p.B = B; p.H = H; p.W = W; p.Cout = Cout; p.relu = relu ? 1 : 0;
  p.n_groups = Cout / 64;
  p.m_units = static_cast<int32_t>((P + 255) / 256);
  const int patch_rows = 128 + 2 * (W + 2);
  p.rb = ((patch_rows + 1) / 2 + 7) / 8 * 8;
  CUtensorMap ta, tw_hi, tw_lo;
  TOAD_TRY(make_bf16_tmap(&ta, in_padded, P, Cin, p.rb));
  TOAD_TRY(make_bf16_tmap(&tw_hi, w_hi, Cout, 9 * Cin, 32));
  TOAD_TRY(make_bf16_tmap(&tw_lo, w_lo, Cout, 9 * Cin, 32));
  auto kern = conv3x3_halo_kernel<KB>;
  {
    static int attr_dev = -1;
    int dev = 0;
    TOAD_CUDA_TRY(cudaGetDevice(&dev));
    if (attr_dev != dev) {
      TOAD_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
      attr_dev = dev;
    }
  }
  int pairs = sm_count() / 2;
  pairs -= pairs % p.n_groups;
  const int64_t units = static_cast<int64_t>(p.m_units) * p.n_groups;
  if (units < pairs) pairs = static_cast<int>(units);   // (units is a multiple of n_groups)
  if (pairs < p.n_groups) return TOAD_ERR_UNSUPPORTED;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(static_cast<unsigned>(pairs * 2));
  cfg.blockDim = dim3(C::THREADS);
  cfg.dynamicSmemBytes = C::SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  int na = 0;
  attr[na].id = cudaLaunchAttributeClusterDimension;
  attr[na].val.clusterDim.x = 2;
  attr[na].val.clusterDim.y = 1;
  attr[na].val.clusterDim.z = 1;
  ++na;
  if (pdl) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  TOAD_CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, ta, tw_hi, tw_lo, p));
  TOAD_CUDA_TRY(cudaGetLastError());
  return 0;
}

}  // namespace halo
}  // namespace toad
