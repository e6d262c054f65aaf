// Fused stem of resnet50_baseline (reference models/resnet_custom.py:61-63, 97-99): conv1 7x7 / stride 2 / pad 3
// (3 -> 64 channels) + folded BatchNorm + ReLU as an IMPLICIT GEMM on tcgen05 -- the im2col operand is built in shared
// memory, tile by tile, straight from the NCHW fp32 image and never exists in HBM (the explicit im2col plane is
// 6.3 MB per 256 x 256 patch in fp16: written once, read once, ~12 % of the trunk's time and ~20 % of its DRAM bytes).
//
// One persistent CTA per SM (cta_group::1).  A tile = 128 consecutive output pixels of one output row (b, oh):
//   A [128 pixels x 192]   K order chosen for the gather: column k = (c*7 + kh)*8 + (kw + 1), i.e. one 8-column group
//                          per (channel, tap row), kw fastest, slot 0 of each group and the groups 21..23 being zero
//                          padding.  Output pixel r needs input columns 2r - 3 .. 2r + 3 of that image row, so a group
//                          is EIGHT CONSECUTIVE staged floats starting at 2r: four 64-bit shared loads per lane, and a
//                          warp's 32 consecutive pixels read 64 consecutive words -- no offset table, no bank conflict.
//   B [64 channels x 192]  the folded weights in the same K order as fp16 (hi, lo) planes (toad_resnet_prepare with
//                          the stem layout), loaded ONCE by TMA
//   D = A.B_hi + A.B_lo    fp32 in TMEM (two 64-column accumulators), fp16 single-plane mode of gemm_tc.cuh
// Warps: 0 = loads B, 1 = MMA issuer, 2 = TMEM allocator, 4-7 = epilogue (bias + ReLU in the fp16 conversion, 128 B
// per pixel stored by its own lane), 8-15 = A builders: they stage the 21 input rows (3 channels x 7 taps) the tile
// touches with coalesced 128-bit loads (zero padded; issued one tile ahead into registers, so their latency hides
// behind the gather of the current tile), then write the swizzled K-major UMMA tile, one work item = (column group,
// 32 pixels).  A is double buffered: building tile i+1 overlaps the MMAs of tile i, whose epilogue overlaps the MMAs
// of tile i+1.
//
// POOL = true (output rows of <= 128 pixels, i.e. W <= 256) also folds MaxPool2d(3, stride 2, pad 1)
// (resnet_custom.py:64,100) into the epilogue: a CTA walks BANDS of consecutive output rows of one image, the epilogue
// warps park each finished row (fp16, swizzled) in a 3-row shared-memory ring, and every second row they emit one
// pooled row (3 x 3 window maximum over the ring) -- the 128 x 128 x 64 stem activation never exists in HBM.  A band
// yields 8 pooled rows from 17 stem rows (one halo row recomputed per band: +6 %).
#pragma once
#include "gemm_tc.cuh"

namespace toad {
namespace stem {

constexpr int K_REAL = 147, K_PAD = 192, KBLOCKS = 3;
constexpr int COUT = 64;
constexpr int TILE_PIX = 128;
constexpr int BUILD_WARPS = 8, BUILD_THREADS = BUILD_WARPS * 32;
constexpr int THREADS = (8 + BUILD_WARPS) * 32;         // 512
constexpr int STG_W = 2 * TILE_PIX + 8;                 // 264 staged input columns per row ...
constexpr int STG_ROWS = 21;                            // (channel, kh) = the 21 real 8-column groups
constexpr int STG_LOADS = (STG_ROWS * (STG_W / 4) + BUILD_THREADS - 1) / BUILD_THREADS;  // float4 loads per thread and tile
constexpr int REAL_GROUPS = STG_ROWS;                   // column groups 0..20 are real, 21..23 zero padding
constexpr int A_KB_BYTES = TILE_PIX * 128;              // 16 KB: one K block of the A tile
constexpr int A_BUF_BYTES = KBLOCKS * A_KB_BYTES;       // 48 KB
constexpr int B_KB_BYTES = COUT * 128;                  // 8 KB per plane and K block
constexpr int ACC_COLS = 2 * COUT;                      // accumulator columns: [x . w_hi | x . w_lo] (the planes ride in N, see below)
constexpr int B_BYTES = KBLOCKS * 2 * B_KB_BYTES;       // 48 KB
constexpr int STG_BYTES = STG_ROWS * STG_W * 4;         // 22 KB
constexpr int SMEM_BYTES = B_BYTES + 2 * A_BUF_BYTES + STG_BYTES + 1024;
constexpr int POOL_BAND = 8;                            // pooled rows per band
constexpr int RING_ROW_BYTES = TILE_PIX * COUT * 2;     // 16 KB: one stem row, fp16
constexpr int SMEM_BYTES_POOL = SMEM_BYTES + 3 * RING_ROW_BYTES;
static_assert(SMEM_BYTES_POOL + 1024 <= 227 * 1024, "shared memory");

struct StemParams {
  const float* x;        // [B, 3, H, W] fp32 NCHW
  const float* bias;     // [64] folded BatchNorm bias
  __nv_bfloat16* out;    // fp16 bits, NHWC: [B*Ho*Wo, 64] (POOL = false) or the pooled [B*(Ho/2)*(Wo/2), 64] (POOL = true)
  int32_t B, H, W, Ho, Wo;
  int32_t tiles_per_row; // ceil(Wo / 128)   (POOL: 1)
  int32_t n_tiles;       // B * Ho * tiles_per_row (POOL = false)
  int32_t bands_per_img; // ceil((Ho/2) / POOL_BAND)   (POOL = true)
  int32_t n_bands;       // B * bands_per_img
};

// The sequence of tiles (= 128-pixel pieces of output rows) one CTA walks; identical in every warp role.
template <bool POOL>
struct TileWalk {
  int tile;                        // plain: linear tile index
  int band, sr, sr_last, p0;       // pooled: band index, current / last stem row of the band, first pooled row
  int b, oh, ow0;                  // the current tile
  bool valid;
  __device__ __forceinline__ void set_band(const StemParams& p) {
    valid = band < p.n_bands;
    if (!valid) return;
    b = band / p.bands_per_img;
    const int j = band - b * p.bands_per_img;
    p0 = j * POOL_BAND;
    int p1 = p0 + POOL_BAND;
    if (p1 > p.Ho / 2) p1 = p.Ho / 2;
    sr = p0 > 0 ? 2 * p0 - 1 : 0;
    sr_last = 2 * p1 - 1;
    oh = sr; ow0 = 0;
  }
  __device__ __forceinline__ void set_tile(const StemParams& p) {
    valid = tile < p.n_tiles;
    const int row_tile = tile / p.tiles_per_row;
    ow0 = (tile - row_tile * p.tiles_per_row) * TILE_PIX;
    b = row_tile / p.Ho; oh = row_tile - b * p.Ho;
  }
  __device__ __forceinline__ void init(const StemParams& p) {
    if (POOL) { band = blockIdx.x; set_band(p); } else { tile = blockIdx.x; set_tile(p); }
  }
  __device__ __forceinline__ void next(const StemParams& p) {
    if (POOL) {
      if (++sr > sr_last) { band += gridDim.x; set_band(p); } else { oh = sr; }
    } else {
      tile += gridDim.x; set_tile(p);
    }
  }
};

template <bool POOL>
__global__ void __launch_bounds__(THREADS, 1)
stem_conv_kernel(const __grid_constant__ CUtensorMap tm_b_hi, const __grid_constant__ CUtensorMap tm_b_lo, const StemParams p) {
  using namespace tc;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_b, bar_a_full[2], bar_a_empty[2], bar_acc_full[2], bar_acc_empty[2];
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(16) float s_bias[COUT];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sB = base, sA = base + B_BYTES, sStg = sA + 2 * A_BUF_BYTES, sRing = sStg + STG_BYTES;
  float* const stg = reinterpret_cast<float*>(smem_raw + (sStg - smem_u32(smem_raw)));

  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bar_b), 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(smem_u32(&bar_a_full[i]), BUILD_WARPS);
      mbar_init(smem_u32(&bar_a_empty[i]), 1);
      mbar_init(smem_u32(&bar_acc_full[i]), 1);
      mbar_init(smem_u32(&bar_acc_empty[i]), 4);
    }
    fence_barrier_init();
  }
  if (threadIdx.x < COUT) s_bias[threadIdx.x] = __ldg(p.bias + threadIdx.x);
  if (warp == 2) tmem_alloc<1>(smem_u32(&tmem_slot), 2 * ACC_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;

  if (warp == 0) {
    // ---------------------------------------------------------------- the weights, once
    if (lane == 0) {
      prefetch_tmap(&tm_b_hi);
      prefetch_tmap(&tm_b_lo);
      mbar_expect_tx(smem_u32(&bar_b), B_BYTES);
      for (int kb = 0; kb < KBLOCKS; ++kb) {
        tma_load_2d<1>(sB + kb * 2 * B_KB_BYTES, &tm_b_hi, smem_u32(&bar_b), kb * 64, 0);
        tma_load_2d<1>(sB + kb * 2 * B_KB_BYTES + B_KB_BYTES, &tm_b_lo, smem_u32(&bar_b), kb * 64, 0);
      }
    }
  } else if (warp == 1) {
    // ---------------------------------------------------------------- MMA issuer
    // one 128 x 128 x 16 UMMA per K step: B = the K block's [64 hi rows | 64 lo rows] (contiguous), so A is read from
    // shared memory once for both weight planes; the epilogue adds the two column halves
    constexpr uint32_t idesc = make_idesc_f16(TILE_PIX, ACC_COLS);
    mbar_wait(smem_u32(&bar_b), 0);
    int it = 0;
    TileWalk<POOL> tw;
    for (tw.init(p); tw.valid; tw.next(p), ++it) {
      const int buf = it & 1;
      const uint32_t par = (it >> 1) & 1;
      mbar_wait(smem_u32(&bar_acc_empty[buf]), par ^ 1);
      mbar_wait(smem_u32(&bar_a_full[buf]), par);
      tc_fence_after();
      if (lane == 0) {
        const uint32_t d_tmem = tmem_base + buf * ACC_COLS;
#pragma unroll
        for (int kb = 0; kb < KBLOCKS; ++kb) {
          const uint64_t a = make_kmajor_sw128_desc(sA + buf * A_BUF_BYTES + kb * A_KB_BYTES);
          const uint64_t b = make_kmajor_sw128_desc(sB + kb * 2 * B_KB_BYTES);
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
            const uint64_t koff = static_cast<uint64_t>(k) * ((UMMA_K * 2) >> 4);
            umma_bf16<1>(d_tmem, a + koff, b + koff, idesc, (kb > 0) || (k != 0));
          }
        }
        umma_commit<1>(smem_u32(&bar_a_empty[buf]));   // the A buffer may be rebuilt once these MMAs retire
        umma_commit<1>(smem_u32(&bar_acc_full[buf]));  // accumulator ready
      }
      __syncwarp();
    }
  } else if (warp >= 4 && warp < 8) {
    // ---------------------------------------------------------------- epilogue: lane = output pixel
    const int ew = warp - 4;
    int it = 0;
    TileWalk<POOL> tw;
    for (tw.init(p); tw.valid; tw.next(p), ++it) {
      const int buf = it & 1;
      const int r = ew * 32 + lane;
      const bool ok = tw.ow0 + r < p.Wo;
      __nv_bfloat16* dst = p.out + ((static_cast<int64_t>(tw.b) * p.Ho + tw.oh) * p.Wo + tw.ow0 + r) * COUT;   // (POOL = false)
      const uint32_t ring_row = sRing + (POOL ? (tw.sr % 3) : 0) * RING_ROW_BYTES;
      mbar_wait(smem_u32(&bar_acc_full[buf]), (it >> 1) & 1);
      tc_fence_after();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + buf * ACC_COLS;
#pragma unroll
      for (int c = 0; c < COUT / 32; ++c) {
        uint32_t v[32], vl[32];
        tmem_ld32(t_row + c * 32, v);
        tmem_ld32(t_row + COUT + c * 32, vl);
        tmem_ld_wait();
        if (c == COUT / 32 - 1) {  // accumulator drained by this warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(smem_u32(&bar_acc_empty[buf]));
        }
        const float4* sb = reinterpret_cast<const float4*>(s_bias + c * 32);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float4 b4 = sb[q];
          add_f32x2(v[4 * q], v[4 * q + 1], __uint_as_float(vl[4 * q]) + b4.x, __uint_as_float(vl[4 * q + 1]) + b4.y);
          add_f32x2(v[4 * q + 2], v[4 * q + 3], __uint_as_float(vl[4 * q + 2]) + b4.z, __uint_as_float(vl[4 * q + 3]) + b4.w);
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint32_t h[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) h[e] = pack_relu_f16x2(v[8 * q + 2 * e], v[8 * q + 2 * e + 1]);
          if (POOL) {  // park the pixel's 16-byte chunk j = 4c + q in the row ring (chunk index ^= pixel & 7: conflict-free)
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(ring_row + r * 128 + (((c * 4 + q) ^ (r & 7)) << 4)),
                         "r"(h[0]), "r"(h[1]), "r"(h[2]), "r"(h[3]) : "memory");
          } else if (ok) {
            *reinterpret_cast<uint4*>(dst + c * 32 + q * 8) = make_uint4(h[0], h[1], h[2], h[3]);
          }
        }
      }
      if (POOL && (tw.sr & 1) && tw.sr >= 2 * tw.p0 + 1) {
        // stem rows sr-2, sr-1, sr are in the ring: emit pooled row ph = (sr - 1) / 2 (MaxPool2d(3, 2, 1); the values are
        // post-ReLU, so the zero padding of the window edges is simply left out of the maximum)
        asm volatile("bar.sync 2, 128;" ::: "memory");   // every epilogue warp has parked its part of row sr
        const int ph = (tw.sr - 1) >> 1, W2 = p.Wo >> 1;
        __nv_bfloat16* prow = p.out + (static_cast<int64_t>(tw.b) * (p.Ho >> 1) + ph) * W2 * COUT;
        for (int idx = ew * 32 + lane; idx < W2 * 8; idx += 128) {
          const int pw = idx >> 3, j = idx & 7;
          __half2 m[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) m[e] = __float2half2_rn(0.f);
#pragma unroll
          for (int dr = 0; dr < 3; ++dr) {
            const int rr = tw.sr - 2 + dr;
            if (rr < 0) continue;
            const uint32_t rb = sRing + (rr % 3) * RING_ROW_BYTES;
#pragma unroll
            for (int dc = 0; dc < 3; ++dc) {
              const int ow = 2 * pw - 1 + dc;
              if (ow < 0) continue;
              uint4 t;
              asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(t.x), "=r"(t.y), "=r"(t.z), "=r"(t.w)
                           : "r"(rb + ow * 128 + ((j ^ (ow & 7)) << 4)) : "memory");
              m[0] = __hmax2(m[0], *reinterpret_cast<const __half2*>(&t.x));
              m[1] = __hmax2(m[1], *reinterpret_cast<const __half2*>(&t.y));
              m[2] = __hmax2(m[2], *reinterpret_cast<const __half2*>(&t.z));
              m[3] = __hmax2(m[3], *reinterpret_cast<const __half2*>(&t.w));
            }
          }
          *reinterpret_cast<uint4*>(prow + static_cast<int64_t>(pw) * COUT + j * 8) =
              make_uint4(*reinterpret_cast<const uint32_t*>(&m[0]), *reinterpret_cast<const uint32_t*>(&m[1]),
                         *reinterpret_cast<const uint32_t*>(&m[2]), *reinterpret_cast<const uint32_t*>(&m[3]));
        }
        asm volatile("bar.sync 2, 128;" ::: "memory");   // the ring slot of row sr-2 may be overwritten by row sr+1
      }
    }
  } else if (warp >= 8) {
    // ---------------------------------------------------------------- A builders (256 threads)
    const int u = threadIdx.x - 8 * 32, bw = warp - 8;
    const int W4 = p.W >> 2;
    // the padding columns (k >= 152) of both A buffers: zero once, never written again
    constexpr int PAD_GROUPS = 24 - REAL_GROUPS;  // 3
    for (int i = u; i < 2 * TILE_PIX * PAD_GROUPS; i += BUILD_THREADS) {
      const int buf = i / (TILE_PIX * PAD_GROUPS), rem = i - buf * (TILE_PIX * PAD_GROUPS);
      const int r = rem / PAD_GROUPS, jc = REAL_GROUPS - 16 + (rem - r * PAD_GROUPS);
      asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(sA + buf * A_BUF_BYTES + 2 * A_KB_BYTES + r * 128 + ((jc ^ (r & 7)) << 4)), "r"(0u) : "memory");
    }
    float4 pre[STG_LOADS];
    // float4 i of the tile's staging = input columns 4*(q0 + q) .. +3 of row ih = 2*oh + kh - 3 of channel c
    auto issue_loads = [&](const TileWalk<POOL>& t) {
      const int b = t.b, oh = t.oh;
      const int q0 = (2 * t.ow0 - 4) >> 2;  // (may be -1)
#pragma unroll
      for (int j = 0; j < STG_LOADS; ++j) {
        const int i = u + j * BUILD_THREADS;
        const int rr = i / (STG_W / 4), q = i - rr * (STG_W / 4);
        const int c = rr / 7, kh = rr - c * 7;
        const int ih = 2 * oh + kh - 3, q4 = q0 + q;
        pre[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (rr < STG_ROWS && t.valid && ih >= 0 && ih < p.H && q4 >= 0 && q4 < W4)
          pre[j] = ld_stream_f4(p.x + ((static_cast<int64_t>(b) * 3 + c) * p.H + ih) * p.W + q4 * 4);
      }
    };
    TileWalk<POOL> tw, nxt;
    tw.init(p);
    nxt = tw;
    issue_loads(nxt);
    int it = 0;
    for (; tw.valid; tw.next(p), ++it) {
      const int buf = it & 1;
      // everybody has finished gathering the previous tile from the staging rows
      asm volatile("bar.sync 1, %0;" ::"n"(BUILD_THREADS) : "memory");
#pragma unroll
      for (int j = 0; j < STG_LOADS; ++j) {   // staged column s of row (c, kh) = input column 2*ow0 - 4 + s
        const int i = u + j * BUILD_THREADS;
        if (i < STG_ROWS * (STG_W / 4)) reinterpret_cast<float4*>(stg)[i] = pre[j];
      }
      asm volatile("bar.sync 1, %0;" ::"n"(BUILD_THREADS) : "memory");
      nxt.next(p);
      issue_loads(nxt);                // next tile's rows: in flight during this tile's gather
      // the MMAs that read this A buffer two tiles ago have retired
      mbar_wait(smem_u32(&bar_a_empty[buf]), ((it >> 1) & 1) ^ 1);
      // work item = (column group cg = (c, kh), 32-pixel group pg): lane = pixel r; the group's 8 columns are the
      // staged floats 2r .. 2r+7 of row cg (kw = -1 .. 6; the kw = -1 slot meets a zero weight)
      for (int item = bw; item < REAL_GROUPS * 4; item += BUILD_WARPS) {
        const int cg = item >> 2, pg = item & 3;
        const int r = pg * 32 + lane;
        const float2* src = reinterpret_cast<const float2*>(stg + cg * STG_W + 2 * r);
        const float2 v0 = src[0], v1 = src[1], v2 = src[2], v3 = src[3];
        const uint32_t h0 = pack_f16x2(v0.x, v0.y), h1 = pack_f16x2(v1.x, v1.y), h2 = pack_f16x2(v2.x, v2.y), h3 = pack_f16x2(v3.x, v3.y);
        const uint32_t dst = sA + buf * A_BUF_BYTES + (cg >> 3) * A_KB_BYTES + r * 128 + (((cg & 7) ^ (r & 7)) << 4);  // SWIZZLE_128B
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(h0), "r"(h1), "r"(h2), "r"(h3) : "memory");
      }
      fence_proxy_async_smem();  // generic-proxy writes -> visible to the tensor core
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&bar_a_full[buf]));
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<1>(tmem_base, 2 * ACC_COLS);
  }
}

// x [B, 3, H, W] fp32 -> relu(conv7x7s2(x) * bn_scale + bn_bias) as fp16 NHWC:
//   pool = false: out [B*(H/2)*(W/2), 64];   pool = true (W <= 256, H % 4 == 0): out [B*(H/4)*(W/4), 64] after MaxPool2d(3, 2, 1).
// w_hi / w_lo: [64, 192] fp16 planes in the stem's K order (toad_resnet_prepare), bias [64].  H, W even, W % 4 == 0.
inline bool stem_can_pool(int H, int W) { return W <= 2 * TILE_PIX && (H % 4) == 0 && (W % 4) == 0; }

inline int launch_stem_fused(const float* x, const __nv_bfloat16* w_hi, const __nv_bfloat16* w_lo, const float* bias,
                             __nv_bfloat16* out, int B, int H, int W, bool pool, cudaStream_t stream) {
  if (B <= 0) return 0;
  if ((H & 1) || (W & 3) || H < 2 || W < 4) return TOAD_ERR_UNSUPPORTED;
  if (pool && !stem_can_pool(H, W)) return TOAD_ERR_UNSUPPORTED;
  StemParams p{};
  p.x = x; p.bias = bias; p.out = out; p.B = B; p.H = H; p.W = W; p.Ho = H / 2; p.Wo = W / 2;
  p.tiles_per_row = (p.Wo + TILE_PIX - 1) / TILE_PIX;
  const int64_t tiles = static_cast<int64_t>(B) * p.Ho * p.tiles_per_row;
  if (tiles > 0x7fffffff) return TOAD_ERR_UNSUPPORTED;
  p.n_tiles = static_cast<int32_t>(tiles);
  p.bands_per_img = (p.Ho / 2 + POOL_BAND - 1) / POOL_BAND;
  p.n_bands = B * p.bands_per_img;
  CUtensorMap tb_hi, tb_lo;
  TOAD_TRY(tc::make_bf16_tmap(&tb_hi, w_hi, COUT, K_PAD, COUT, K_PAD));
  TOAD_TRY(tc::make_bf16_tmap(&tb_lo, w_lo, COUT, K_PAD, COUT, K_PAD));
  {
    static int attr_dev = -1;
    int dev = 0;
    TOAD_CUDA_TRY(cudaGetDevice(&dev));
    if (attr_dev != dev) {
      TOAD_CUDA_TRY(cudaFuncSetAttribute(stem_conv_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
      TOAD_CUDA_TRY(cudaFuncSetAttribute(stem_conv_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES_POOL));
      attr_dev = dev;
    }
  }
  const int sms = tc::sm_count();
  const int64_t units = pool ? p.n_bands : tiles;
  const int grid = static_cast<int>(units < sms ? units : sms);
  if (pool) stem_conv_kernel<true><<<grid, THREADS, SMEM_BYTES_POOL, stream>>>(tb_hi, tb_lo, p);
  else stem_conv_kernel<false><<<grid, THREADS, SMEM_BYTES, stream>>>(tb_hi, tb_lo, p);
  TOAD_CUDA_TRY(cudaGetLastError());
  return 0;
}

}  // namespace stem
}  // namespace toad
