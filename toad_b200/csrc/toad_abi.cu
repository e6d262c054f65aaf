// extern "C" entry points of libtoad_b200.so (see include/toad_b200.h for the contract).
#include "common.cuh"
#include "gemm_tc.cuh"
#include "sgemm_simt.cuh"
#include "tail.cuh"
#include "bwd.cuh"
#include "topk.cuh"
#include "resnet.cuh"
#include "stem.cuh"
#include "conv3x3_halo.cuh"
#include "train.cuh"
#include <cmath>
#include <cstdlib>
#include <new>

namespace {

using namespace toad;
typedef __nv_bfloat16 bf16;

constexpr int kSMs = 148;       // B200; used only to size workspaces and tail grids
constexpr int kGateHalf = 128;  // gate columns per EPI_GATE tile (BLOCK_N 256 / 2)

struct Carver {
  uint8_t* base;
  size_t off;
  explicit Carver(void* b) : base(static_cast<uint8_t*>(b)), off(0) {}
  template <class T>
  T* take(size_t count) {
    off = align_up(off, 256);
    T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
    off += count * sizeof(T);
    return p;
  }
};

int check_dims(const toad_dims_t* d) {
  if (d == nullptr) return TOAD_ERR_ARG;
  if (d->hid_dim != tail::H || d->n_tasks != tail::T) return TOAD_ERR_UNSUPPORTED;
  if (d->in_dim <= 0 || d->in_dim % 64 != 0) return TOAD_ERR_UNSUPPORTED;
  if (d->attn_dim <= 0 || d->attn_dim % kGateHalf != 0 || d->attn_dim > 1024) return TOAD_ERR_UNSUPPORTED;
  if (d->n_classes < 1 || d->n_classes > 1024) return TOAD_ERR_UNSUPPORTED;
  return 0;
}

struct FwdWs {
  unsigned int* ticket;
  float* blk_part;
  float* part;  // [n_parts][N][T]
  int n_parts;
  // tensor-core path
  bf16 *w1_hi, *w1_lo, *w2_hi, *w2_lo, *wab_hi, *wab_lo, *h1_hi, *h1_lo, *h_hi, *h_lo;
  // fp32 path (when the caller gave no `saved`)
  float *h1, *h, *a, *b;
  size_t bytes;
};

FwdWs carve_fwd(const toad_dims_t* d, int64_t n, uint32_t flags, void* base, int n_slides = 1) {
  FwdWs w{};
  Carver c(base);
  const int64_t Hd = d->hid_dim, D = d->attn_dim, L = d->in_dim;
  const bool simt = (flags & TOAD_FLAG_SIMT_FP32) != 0;
  w.ticket = c.take<unsigned int>(64 * tail::MAX_BATCH);  // (fixed size: the weight planes behind it keep their offsets)
  if (!simt) {  // weight planes first: their offsets do not depend on n (TOAD_FLAG_REUSE_WEIGHT_PLANES)
    w.w1_hi = c.take<bf16>(Hd * L);  w.w1_lo = c.take<bf16>(Hd * L);
    w.w2_hi = c.take<bf16>(Hd * Hd); w.w2_lo = c.take<bf16>(Hd * Hd);
    w.wab_hi = c.take<bf16>(2 * D * Hd); w.wab_lo = c.take<bf16>(2 * D * Hd);
  }
  w.blk_part = c.take<float>(static_cast<size_t>(tail::tail_blocks(n, kSMs) + tail::MAX_GROUPS) * tail::PART_STRIDE * n_slides);
  w.n_parts = simt ? 1 : 2 * static_cast<int>(D / kGateHalf);  // (tile, epilogue warp set) partials
  w.part = c.take<float>(static_cast<size_t>(w.n_parts) * n * d->n_tasks);
  if (!simt) {
    w.h1_hi = c.take<bf16>(n * Hd); w.h1_lo = c.take<bf16>(n * Hd);
    w.h_hi = c.take<bf16>(n * Hd);  w.h_lo = c.take<bf16>(n * Hd);
  } else if (!(flags & TOAD_FLAG_SAVE_ACTS)) {
    w.h1 = c.take<float>(n * Hd); w.h = c.take<float>(n * Hd);
    w.a = c.take<float>(n * D);   w.b = c.take<float>(n * D);
  }
  w.bytes = align_up(c.off, 256);
  return w;
}

struct Prof {
  cudaEvent_t* ev;  // [max_calls][TOAD_N_STAGES + 1]
  int max_calls;
  int n;
};
inline int prof_mark(Prof* p, int idx, cudaStream_t st) {
  if (p == nullptr || p->n >= p->max_calls) return 0;
  TOAD_CUDA_TRY(cudaEventRecord(p->ev[p->n * (TOAD_N_STAGES + 1) + idx], st));
  return 0;
}

DropoutCfg make_drop(const toad_saved_t* sv, uint32_t flags) {
  DropoutCfg d{};
  if ((flags & TOAD_FLAG_DROPOUT) && sv != nullptr && sv->dropout_p > 0.f) {
    d.seed = sv->dropout_seed;
    const double t = static_cast<double>(sv->dropout_p) * 4294967296.0;
    d.thresh = t >= 4294967295.0 ? 0xFFFFFFFFu : static_cast<uint32_t>(t);
    d.scale = 1.0f / (1.0f - sv->dropout_p);
  }
  return d;
}

int check_ws(const void* ws, size_t have, size_t need) {
  if (ws == nullptr || (reinterpret_cast<uintptr_t>(ws) & 255) != 0 || have < need) return TOAD_ERR_WORKSPACE;
  return 0;
}

int run_tail(const toad_dims_t* d, const toad_params_t* P, int64_t n, const float* sex, const toad_fwd_out_t* out,
             const FwdWs& w, const float* h_f32, const bf16* h_hi, const bf16* h_lo, bool attention_only,
             cudaStream_t st, const tail::TailBatch* batch = nullptr) {
  tail::TailParams t{};
  if (batch != nullptr) {  // several slides back to back in the n rows: grid.y = slide, grid.x sized by the largest
    t.batch = *batch;
    t.stride = n;
    int64_t largest = 0;
    for (int s = 0; s < batch->n_slides; ++s)
      if (batch->off[s + 1] - batch->off[s] > largest) largest = batch->off[s + 1] - batch->off[s];
    n = largest;
  }
  t.part = w.part; t.n_parts = w.n_parts; t.bc = P->bc; t.a_raw = out->a_raw;
  t.h_f32 = h_f32; t.h_hi = h_hi; t.h_lo = h_lo; t.N = n; t.sex = sex;
  t.wcls = P->wcls; t.bcls = P->bcls; t.n_classes = d->n_classes; t.wsite = P->wsite; t.bsite = P->bsite;
  t.features = out->features; t.logits = out->logits; t.y_prob = out->y_prob; t.y_hat = out->y_hat;
  t.site_logits = out->site_logits; t.site_prob = out->site_prob; t.site_hat = out->site_hat;
  t.stats = out->softmax_stats; t.blk_part = w.blk_part; t.ticket = w.ticket;
  t.attention_only = attention_only ? 1 : 0;
  // (tensor-core path: launched programmatically dependent on the gate GEMM)
  return h_f32 ? tail::launch_tail<tail::H_F32>(t, kSMs, st) : tail::launch_tail<tail::H_SPLIT>(t, kSMs, st, tc::pdl_enabled());
}

simt::SgemmParams linear_params(const float* x, int64_t ldx, const float* w, const float* bias, float* y, int64_t m,
                                int n, int64_t k) {
  simt::SgemmParams s{};
  s.a = x; s.a_rs = ldx; s.a_ks = 1;
  s.b = w; s.b_rs = k; s.b_ks = 1;
  s.c = y; s.ldc = n; s.M = m; s.N = n; s.K = k; s.k_chunk = k; s.bias = bias;
  return s;
}

}  // namespace

extern "C" int toad_abi_version(void) { return TOAD_ABI_VERSION; }

#ifndef TOAD_BUILD_ID
#define TOAD_BUILD_ID "unknown"
#endif
extern "C" const char* toad_build_id(void) { return TOAD_BUILD_ID; }

extern "C" uint32_t toad_dropout_hash(uint64_t seed, uint32_t layer, uint64_t index) {
  return dropout_hash(seed, layer, index);
}

extern "C" const char* toad_error_string(int code) {
  switch (code) {
    case TOAD_OK: return "ok";
    case TOAD_ERR_ARG: return "invalid argument (null pointer or bad size)";
    case TOAD_ERR_WORKSPACE: return "workspace too small or not 256-byte aligned";
    case TOAD_ERR_UNSUPPORTED: return "dimensions not supported by the sm_100a kernels";
    case TOAD_ERR_DRIVER: return "could not obtain cuTensorMapEncodeTiled / encode a TMA descriptor";
    default: return code > 0 ? cudaGetErrorString(static_cast<cudaError_t>(code)) : "unknown error";
  }
}

extern "C" int toad_param_offsets(const toad_dims_t* d, int64_t offsets[15]) {
  if (d == nullptr || offsets == nullptr) return TOAD_ERR_ARG;
  const int64_t L = d->in_dim, Hd = d->hid_dim, D = d->attn_dim, T = d->n_tasks, C = d->n_classes;
  const int64_t sizes[14] = {Hd * L, Hd, Hd * Hd, Hd, D * Hd, D, D * Hd, D, T * D, T, C * (Hd + 1), C, 2 * (Hd + 1), 2};
  int64_t o = 0;
  for (int i = 0; i < 14; ++i) { offsets[i] = o; o += sizes[i]; }
  offsets[14] = o;
  return 0;
}

extern "C" int toad_fwd_workspace_bytes(const toad_dims_t* d, int64_t n, uint32_t flags, size_t* bytes) {
  TOAD_TRY(check_dims(d));
  if (bytes == nullptr || n <= 0) return TOAD_ERR_ARG;
  *bytes = carve_fwd(d, n, flags, nullptr).bytes;
  return 0;
}

static int fwd_impl(const toad_dims_t* d, const toad_params_t* P, const float* x, int64_t n, const float* sex,
                    const toad_fwd_out_t* out, const toad_saved_t* saved, void* workspace, size_t workspace_bytes,
                    uint32_t flags, toad_stream_t stream, Prof* prof, const tail::TailBatch* batch = nullptr) {
  TOAD_TRY(check_dims(d));
  if (P == nullptr || x == nullptr || out == nullptr || out->a_raw == nullptr || n <= 0) return TOAD_ERR_ARG;
  const bool attn_only = (flags & TOAD_FLAG_ATTENTION_ONLY) != 0;
  const bool save = (flags & TOAD_FLAG_SAVE_ACTS) != 0;
  if (!attn_only && (sex == nullptr || out->features == nullptr || out->logits == nullptr || out->y_prob == nullptr ||
                     out->y_hat == nullptr || out->site_logits == nullptr || out->site_prob == nullptr ||
                     out->site_hat == nullptr || out->softmax_stats == nullptr))
    return TOAD_ERR_ARG;
  const bool simt = (flags & TOAD_FLAG_SIMT_FP32) != 0;
  if (save && (saved == nullptr || !saved->a || !saved->b)) return TOAD_ERR_ARG;
  if (save && simt && (!saved->h1 || !saved->h)) return TOAD_ERR_ARG;
  if (save && !simt && (!saved->h1_hi || !saved->h1_lo || !saved->h_hi || !saved->h_lo)) return TOAD_ERR_ARG;
  if ((flags & TOAD_FLAG_DROPOUT) && (!save || saved->dropout_p < 0.f || saved->dropout_p >= 1.f)) return TOAD_ERR_ARG;
  const DropoutCfg drop = make_drop(saved, flags);
  FwdWs w = carve_fwd(d, n, flags, workspace, batch != nullptr ? batch->n_slides : 1);
  TOAD_TRY(check_ws(workspace, workspace_bytes, w.bytes));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int Hd = d->hid_dim, D = d->attn_dim, L = d->in_dim;
  TOAD_CUDA_TRY(cudaMemsetAsync(w.ticket, 0, 64 * tail::MAX_BATCH * sizeof(unsigned int), st));

  if (flags & TOAD_FLAG_SIMT_FP32) {
    float* h1 = save ? saved->h1 : w.h1;
    float* h = save ? saved->h : w.h;
    float* a = save ? saved->a : w.a;
    float* b = save ? saved->b : w.b;
    TOAD_TRY(prof_mark(prof, 0, st));
    TOAD_TRY(prof_mark(prof, 1, st));
    auto with_drop = [&](simt::SgemmParams s, uint32_t layer) { s.drop = drop; s.drop_layer = layer; return s; };
    TOAD_TRY((simt::launch_sgemm<true, true, simt::EPI_BIAS_RELU>(with_drop(linear_params(x, L, P->w1, P->b1, h1, n, Hd, L), DROP_H1), 1, st)));
    TOAD_TRY(prof_mark(prof, 2, st));
    TOAD_TRY((simt::launch_sgemm<true, true, simt::EPI_BIAS_RELU>(with_drop(linear_params(h1, Hd, P->w2, P->b2, h, n, Hd, Hd), DROP_H), 1, st)));
    TOAD_TRY(prof_mark(prof, 3, st));
    TOAD_TRY((simt::launch_sgemm<true, true, simt::EPI_BIAS_TANH>(with_drop(linear_params(h, Hd, P->wa, P->ba, a, n, D, Hd), DROP_A), 1, st)));
    TOAD_TRY((simt::launch_sgemm<true, true, simt::EPI_BIAS_SIGMOID>(with_drop(linear_params(h, Hd, P->wb, P->bb, b, n, D, Hd), DROP_B), 1, st)));
    TOAD_TRY(tail::launch_attn_c(a, b, P->wc, w.part, n, D, d->n_tasks, st));
    TOAD_TRY(prof_mark(prof, 4, st));
    TOAD_TRY(run_tail(d, P, n, sex, out, w, h, nullptr, nullptr, attn_only, st, batch));
    TOAD_TRY(prof_mark(prof, 5, st));
    if (prof != nullptr && prof->n < prof->max_calls) prof->n++;
    return 0;
  }

  // ---- tensor-core path: split weights, three tcgen05 GEMMs with fused epilogues, tail
  // Tile shapes.  Default: fc1 (fp32-fed) runs as CTA pairs on 256 x 512 tiles -- the whole hidden width in
  // one tile, so every x row is read and converted exactly once --, fc2 / gate as CTA pairs on 256 x 256
  // tiles.  Debug flags: SINGLE_CTA = 128 x 256 cta_group::1 tiles everywhere, PAIR_ALL = 256 x 256 pairs
  // everywhere.
  const bool cg1 = (flags & TOAD_FLAG_TC_SINGLE_CTA) != 0;
  const bool fc1_pair = (flags & TOAD_FLAG_TC_PAIR_ALL) != 0;
  // activation planes: workspace scratch for inference, the caller's saved buffers when the backward will need them
  bf16* h1_hi = save ? static_cast<bf16*>(saved->h1_hi) : w.h1_hi;
  bf16* h1_lo = save ? static_cast<bf16*>(saved->h1_lo) : w.h1_lo;
  bf16* h_hi = save ? static_cast<bf16*>(saved->h_hi) : w.h_hi;
  bf16* h_lo = save ? static_cast<bf16*>(saved->h_lo) : w.h_lo;
  TOAD_TRY(prof_mark(prof, 0, st));
  if (!(flags & TOAD_FLAG_REUSE_WEIGHT_PLANES)) {
    TOAD_TRY(tail::launch_split_planes(P->w1, w.w1_hi, w.w1_lo, static_cast<int64_t>(Hd) * L, st));
    TOAD_TRY(tail::launch_split_planes(P->w2, w.w2_hi, w.w2_lo, static_cast<int64_t>(Hd) * Hd, st));
    TOAD_TRY(tail::launch_split_gate_weights(P->wa, P->wb, w.wab_hi, w.wab_lo, D, Hd, kGateHalf, st));
  }
  TOAD_TRY(prof_mark(prof, 1, st));
  {
    tc::GemmTcParams g{};
    g.a_f32 = x; g.lda = L; g.M = n; g.N = Hd; g.K = L; g.bias = P->b1; g.relu = 1;
    g.drop = drop; g.drop_layer = DROP_H1;
    g.out_f32 = save ? saved->h1 : nullptr; g.ld_f32 = Hd;
    g.out_hi = h1_hi; g.out_lo = h1_lo; g.ld_split = Hd;
    if (cg1) TOAD_TRY((tc::launch_gemm<256, tc::A_F32, tc::EPI_LINEAR, 1>(g, nullptr, nullptr, w.w1_hi, w.w1_lo, st)));
    else if (fc1_pair) TOAD_TRY((tc::launch_gemm<256, tc::A_F32, tc::EPI_LINEAR, 2>(g, nullptr, nullptr, w.w1_hi, w.w1_lo, st)));
    else TOAD_TRY((tc::launch_gemm<512, tc::A_F32, tc::EPI_LINEAR, 2>(g, nullptr, nullptr, w.w1_hi, w.w1_lo, st)));
  }
  TOAD_TRY(prof_mark(prof, 2, st));
  {
    tc::GemmTcParams g{};
    g.M = n; g.N = Hd; g.K = Hd; g.bias = P->b2; g.relu = 1;
    g.drop = drop; g.drop_layer = DROP_H;
    g.out_f32 = save ? saved->h : nullptr; g.ld_f32 = Hd;
    g.out_hi = h_hi; g.out_lo = h_lo; g.ld_split = Hd;
    if (cg1) TOAD_TRY((tc::launch_gemm<256, tc::A_SPLIT, tc::EPI_LINEAR, 1>(g, h1_hi, h1_lo, w.w2_hi, w.w2_lo, st)));
    else if (flags & TOAD_FLAG_FC2_WIDE) TOAD_TRY((tc::launch_gemm<512, tc::A_SPLIT, tc::EPI_LINEAR, 2>(g, h1_hi, h1_lo, w.w2_hi, w.w2_lo, st)));
    else TOAD_TRY((tc::launch_gemm<256, tc::A_SPLIT, tc::EPI_LINEAR, 2>(g, h1_hi, h1_lo, w.w2_hi, w.w2_lo, st)));
  }
  TOAD_TRY(prof_mark(prof, 3, st));
  {
    tc::GemmTcParams g{};
    g.M = n; g.N = 2 * D; g.K = Hd;
    g.gate_ba = P->ba; g.gate_bb = P->bb; g.gate_wc = P->wc; g.gate_D = D; g.gate_ntasks = d->n_tasks;
    g.gate_part = w.part; g.gate_a = save ? saved->a : nullptr; g.gate_b = save ? saved->b : nullptr;
    g.drop = drop;
    if (cg1) TOAD_TRY((tc::launch_gemm<256, tc::A_SPLIT, tc::EPI_GATE, 1>(g, h_hi, h_lo, w.wab_hi, w.wab_lo, st)));
    else TOAD_TRY((tc::launch_gemm<256, tc::A_SPLIT, tc::EPI_GATE, 2>(g, h_hi, h_lo, w.wab_hi, w.wab_lo, st)));
  }
  TOAD_TRY(prof_mark(prof, 4, st));
  TOAD_TRY(run_tail(d, P, n, sex, out, w, nullptr, h_hi, h_lo, attn_only, st, batch));
  TOAD_TRY(prof_mark(prof, 5, st));
  if (prof != nullptr && prof->n < prof->max_calls) prof->n++;
  return 0;
}

extern "C" int toad_fwd(const toad_dims_t* d, const toad_params_t* P, const float* x, int64_t n, const float* sex,
                        const toad_fwd_out_t* out, const toad_saved_t* saved, void* workspace, size_t workspace_bytes,
                        uint32_t flags, toad_stream_t stream) {
  return fwd_impl(d, P, x, n, sex, out, saved, workspace, workspace_bytes, flags, stream, nullptr);
}

static int make_batch(const int64_t* offsets, int32_t n_slides, tail::TailBatch* tb) {
  if (offsets == nullptr || n_slides < 1 || n_slides > tail::MAX_BATCH || offsets[0] != 0) return TOAD_ERR_ARG;
  tb->n_slides = n_slides;
  for (int s = 0; s <= n_slides; ++s) tb->off[s] = offsets[s];
  for (int s = 0; s < n_slides; ++s)
    if (offsets[s + 1] <= offsets[s]) return TOAD_ERR_ARG;  // empty slide: softmax over zero patches
  return 0;
}

extern "C" int toad_fwd_batch_workspace_bytes(const toad_dims_t* d, int64_t n_total, int32_t n_slides, uint32_t flags,
                                              size_t* bytes) {
  TOAD_TRY(check_dims(d));
  if (bytes == nullptr || n_total <= 0 || n_slides < 1 || n_slides > tail::MAX_BATCH) return TOAD_ERR_ARG;
  *bytes = carve_fwd(d, n_total, flags, nullptr, n_slides).bytes;
  return 0;
}

extern "C" int toad_fwd_batch(const toad_dims_t* d, const toad_params_t* P, const float* x, const int64_t* offsets,
                              int32_t n_slides, const float* sex, const toad_fwd_out_t* out, void* workspace,
                              size_t workspace_bytes, uint32_t flags, toad_stream_t stream) {
  tail::TailBatch tb{};
  TOAD_TRY(make_batch(offsets, n_slides, &tb));
  if (flags & (TOAD_FLAG_SAVE_ACTS | TOAD_FLAG_DROPOUT | TOAD_FLAG_ATTENTION_ONLY)) return TOAD_ERR_UNSUPPORTED;  // eval only
  return fwd_impl(d, P, x, offsets[n_slides], sex, out, nullptr, workspace, workspace_bytes, flags, stream, nullptr, &tb);
}

extern "C" int toad_fwd_profiled(const toad_dims_t* d, const toad_params_t* P, const float* x, int64_t n,
                                 const float* sex, const toad_fwd_out_t* out, const toad_saved_t* saved,
                                 void* workspace, size_t workspace_bytes, uint32_t flags, toad_stream_t stream,
                                 void* prof) {
  return fwd_impl(d, P, x, n, sex, out, saved, workspace, workspace_bytes, flags, stream, static_cast<Prof*>(prof));
}

extern "C" int toad_profile_create(void** prof, int32_t max_calls) {
  if (prof == nullptr || max_calls <= 0 || max_calls > 65536) return TOAD_ERR_ARG;
  Prof* p = new (std::nothrow) Prof();
  if (p == nullptr) return TOAD_ERR_ARG;
  const int ne = max_calls * (TOAD_N_STAGES + 1);
  p->ev = new (std::nothrow) cudaEvent_t[ne];
  p->max_calls = max_calls;
  p->n = 0;
  if (p->ev == nullptr) { delete p; return TOAD_ERR_ARG; }
  for (int i = 0; i < ne; ++i) {
    cudaError_t e = cudaEventCreate(&p->ev[i]);
    if (e != cudaSuccess) {
      for (int j = 0; j < i; ++j) cudaEventDestroy(p->ev[j]);
      delete[] p->ev;
      delete p;
      return static_cast<int>(e);
    }
  }
  *prof = p;
  return 0;
}

extern "C" int toad_profile_destroy(void* prof) {
  Prof* p = static_cast<Prof*>(prof);
  if (p == nullptr) return TOAD_ERR_ARG;
  for (int i = 0; i < p->max_calls * (TOAD_N_STAGES + 1); ++i) cudaEventDestroy(p->ev[i]);
  delete[] p->ev;
  delete p;
  return 0;
}

extern "C" int toad_profile_read(void* prof, double stage_ms[TOAD_N_STAGES], int32_t* n_calls) {
  Prof* p = static_cast<Prof*>(prof);
  if (p == nullptr || stage_ms == nullptr || n_calls == nullptr) return TOAD_ERR_ARG;
  for (int s = 0; s < TOAD_N_STAGES; ++s) stage_ms[s] = 0.0;
  for (int c = 0; c < p->n; ++c) {
    cudaEvent_t* e = p->ev + c * (TOAD_N_STAGES + 1);
    TOAD_CUDA_TRY(cudaEventSynchronize(e[TOAD_N_STAGES]));
    for (int s = 0; s < TOAD_N_STAGES; ++s) {
      float ms = 0.f;
      TOAD_CUDA_TRY(cudaEventElapsedTime(&ms, e[s], e[s + 1]));
      stage_ms[s] += ms;
    }
  }
  *n_calls = p->n;
  p->n = 0;
  return 0;
}

// ------------------------------------------------------------------------------------------ backward
namespace {
struct BwdWs {
  float *dM, *sdot, *P, *dA, *dab, *dz2, *dz1, *splitk, *gate_part, *col_part;
  int splits, gate_blocks, col_blocks;
  int64_t k_chunk;
  size_t bytes;
};
BwdWs carve_bwd(const toad_dims_t* d, int64_t n, void* base) {
  BwdWs w{};
  Carver c(base);
  const int64_t Hd = d->hid_dim, D = d->attn_dim, L = d->in_dim;
  w.dM = c.take<float>(2 * Hd);
  w.sdot = c.take<float>(64);
  w.P = c.take<float>(n * 2);
  w.dA = c.take<float>(n * 2);
  w.dab = c.take<float>(n * 2 * D);
  w.dz2 = c.take<float>(n * Hd);
  w.dz1 = c.take<float>(n * Hd);
  int64_t s = (n + 1023) / 1024;
  if (s < 1) s = 1;
  if (s > 32) s = 32;
  w.splits = static_cast<int>(s);
  w.k_chunk = ((n + s - 1) / s + 15) / 16 * 16;
  int64_t big = Hd * L;
  if (2 * D * Hd > big) big = 2 * D * Hd;
  w.splitk = c.take<float>(static_cast<size_t>(w.splits) * big);
  int64_t gb = (n + 127) / 128;
  if (gb > 2 * kSMs) gb = 2 * kSMs;
  w.gate_blocks = static_cast<int>(gb);
  w.col_blocks = static_cast<int>(gb);
  w.gate_part = c.take<float>(static_cast<size_t>(gb) * (4 * D + 2));
  w.col_part = c.take<float>(static_cast<size_t>(gb) * Hd);
  w.bytes = align_up(c.off, 256);
  return w;
}
}  // namespace

namespace {
// ---- tensor-core backward workspace
struct BwdTcWs {
  float *dM, *sdot, *P, *dA, *splitk, *gate_part, *col_part;
  bf16 *dab_hi, *dab_lo, *dabT_hi, *dabT_lo, *hT_hi, *hT_lo, *h1T_hi, *h1T_lo, *xT_hi, *xT_lo;
  bf16 *dz2_hi, *dz2_lo, *dz2T_hi, *dz2T_lo, *dz1_hi, *dz1_lo, *dz1T_hi, *dz1T_lo;
  bf16 *w2T_hi, *w2T_lo, *wabT_hi, *wabT_lo;
  bf16 *xp_hi, *xp_lo;  // natural-layout planes of x
  int gate_blocks, col_blocks;
  int64_t ldT;
  size_t bytes;
};
BwdTcWs carve_bwd_tc(const toad_dims_t* d, int64_t n, void* base) {
  BwdTcWs w{};
  Carver c(base);
  const int64_t Hd = d->hid_dim, D = d->attn_dim, L = d->in_dim;
  w.ldT = (n + 63) / 64 * 64;
  w.dM = c.take<float>(2 * Hd);
  w.sdot = c.take<float>(64);
  w.P = c.take<float>(n * 2);
  w.dA = c.take<float>(n * 2);
  w.dab_hi = c.take<bf16>(n * 2 * D);  w.dab_lo = c.take<bf16>(n * 2 * D);
  w.dabT_hi = c.take<bf16>(2 * D * w.ldT); w.dabT_lo = c.take<bf16>(2 * D * w.ldT);
  w.hT_hi = c.take<bf16>(Hd * w.ldT);  w.hT_lo = c.take<bf16>(Hd * w.ldT);
  w.h1T_hi = c.take<bf16>(Hd * w.ldT); w.h1T_lo = c.take<bf16>(Hd * w.ldT);
  w.xT_hi = c.take<bf16>(L * w.ldT);   w.xT_lo = c.take<bf16>(L * w.ldT);
  w.dz2_hi = c.take<bf16>(n * Hd);     w.dz2_lo = c.take<bf16>(n * Hd);
  w.dz2T_hi = c.take<bf16>(Hd * w.ldT); w.dz2T_lo = c.take<bf16>(Hd * w.ldT);
  w.dz1_hi = c.take<bf16>(n * Hd);     w.dz1_lo = c.take<bf16>(n * Hd);
  w.dz1T_hi = c.take<bf16>(Hd * w.ldT); w.dz1T_lo = c.take<bf16>(Hd * w.ldT);
  w.w2T_hi = c.take<bf16>(Hd * Hd);    w.w2T_lo = c.take<bf16>(Hd * Hd);
  w.wabT_hi = c.take<bf16>(Hd * 2 * D); w.wabT_lo = c.take<bf16>(Hd * 2 * D);
  // the MN-major wgrads read natural-layout planes; they alias the (then unused) transposed buffers
  w.xp_hi = w.xT_hi; w.xp_lo = w.xT_lo;
  int64_t big = Hd * L;
  if (2 * D * Hd > big) big = 2 * D * Hd;
  // split-K partials: S slices of one [M_out, N_in] gradient.  S = (74 pairs) / (256x256 tiles) when the gradient has
  // fewer tiles than pairs -- at most one tile per pair in total --, and S = 1 (the whole matrix once) beyond that.
  const size_t pair_tiles = static_cast<size_t>(kSMs / 2) * 256 * 256;
  w.splitk = c.take<float>(pair_tiles > static_cast<size_t>(big) ? pair_tiles : static_cast<size_t>(big));
  int64_t gb = (n + 127) / 128;
  if (gb > 2 * kSMs) gb = 2 * kSMs;
  w.gate_blocks = static_cast<int>(gb);
  w.col_blocks = static_cast<int>(gb);
  w.gate_part = c.take<float>(static_cast<size_t>(gb) * (4 * D + 2));
  w.col_part = c.take<float>(static_cast<size_t>(gb) * Hd);
  w.bytes = align_up(c.off, 256);
  return w;
}

// dW[M_out, N_in] = A^T-planes [M_out, n] . (B^T-planes [N_in, n])^T with K = n patches, split-K over CTA pairs;
// the fp32 partial tiles are then summed in fixed order into `dst` (row stride N_in).
int wgrad_tc(const bf16* aT_hi, const bf16* aT_lo, const bf16* bT_hi, const bf16* bT_lo, int M_out, int N_in, int64_t n,
             int64_t ldT, float* splitk, float* dst, int64_t dst_rows_first, float* dst2, cudaStream_t st) {
  tc::GemmTcParams g{};
  g.M = M_out; g.N = N_in; g.K = static_cast<int32_t>(n); g.lda = ldT; g.ldb = ldT;
  const int num_kb = static_cast<int>((n + 63) / 64);
  const int units = ((M_out + 255) / 256) * (N_in / 256);
  int S = (kSMs / 2) / units;
  if (S < 1) S = 1;
  if (S > num_kb) S = num_kb;
  const int kb_per = (num_kb + S - 1) / S;
  S = (num_kb + kb_per - 1) / kb_per;  // no empty slices
  g.k_splits = S; g.kb_per_split = kb_per;
  g.out_f32 = splitk; g.ld_f32 = N_in;
  TOAD_TRY((tc::launch_gemm<256, tc::A_SPLIT, tc::EPI_LINEAR, 2>(g, aT_hi, aT_lo, bT_hi, bT_lo, st)));
  const int64_t stride = static_cast<int64_t>(M_out) * N_in;
  if (dst2 == nullptr) return bwd::launch_reduce_strided(splitk, dst, stride, stride, S, st);
  TOAD_TRY(bwd::launch_reduce_strided(splitk, dst, dst_rows_first * N_in, stride, S, st));
  return bwd::launch_reduce_strided(splitk + dst_rows_first * N_in, dst2, stride - dst_rows_first * N_in, stride, S, st);
}

// dW[M_out, N_in] = dY^T . X straight from the natural [patch, channel] planes (both operands MN-major), split-K.
int wgrad_mn(const bf16* dy_hi, const bf16* dy_lo, int64_t ld_dy, const bf16* x_hi, const bf16* x_lo, int64_t ld_x,
             int M_out, int N_in, int64_t n, float* splitk, float* dst, int64_t dst_rows_first, float* dst2,
             cudaStream_t st) {
  tc::GemmTcParams g{};
  g.M = M_out; g.N = N_in; g.K = static_cast<int32_t>(n); g.lda = ld_dy; g.ldb = ld_x;
  const int num_kb = static_cast<int>((n + 63) / 64);
  const int units = ((M_out + 255) / 256) * (N_in / 256);
  int S = (kSMs / 2) / units;
  if (S < 1) S = 1;
  if (S > num_kb) S = num_kb;
  const int kb_per = (num_kb + S - 1) / S;
  S = (num_kb + kb_per - 1) / kb_per;
  g.k_splits = S; g.kb_per_split = kb_per;
  g.out_f32 = splitk; g.ld_f32 = N_in;
  TOAD_TRY((tc::launch_gemm_mn<256, 2>(g, dy_hi, dy_lo, x_hi, x_lo, st)));
  const int64_t stride = static_cast<int64_t>(M_out) * N_in;
  if (dst2 == nullptr) return bwd::launch_reduce_strided(splitk, dst, stride, stride, S, st);
  TOAD_TRY(bwd::launch_reduce_strided(splitk, dst, dst_rows_first * N_in, stride, S, st));
  return bwd::launch_reduce_strided(splitk + dst_rows_first * N_in, dst2, stride - dst_rows_first * N_in, stride, S, st);
}

int bwd_tc(const toad_dims_t* d, const toad_params_t* P, const float* x, int64_t n, const toad_fwd_out_t* fo,
           const toad_saved_t* sv, const float* dlogits, const float* dsite, float* grad, void* workspace,
           size_t workspace_bytes, uint32_t flags, toad_stream_t stream) {
  // wgrads: MN-major operands straight from the natural-layout planes (default) or, with the debug flag, K-major
  // operands from explicitly transposed planes (the first implementation, kept as a cross-check)
  const bool mn = (flags & TOAD_FLAG_BWD_TRANSPOSED) == 0;
  BwdTcWs w = carve_bwd_tc(d, n, workspace);
  TOAD_TRY(check_ws(workspace, workspace_bytes, w.bytes));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int Hd = d->hid_dim, D = d->attn_dim, L = d->in_dim;
  if (sv->dropout_p < 0.f || sv->dropout_p >= 1.f) return TOAD_ERR_ARG;
  // the saved activations of the tensor-core path are the (hi, lo) planes the forward GEMMs consumed
  const bf16 *sh1_hi = static_cast<const bf16*>(sv->h1_hi), *sh1_lo = static_cast<const bf16*>(sv->h1_lo);
  const bf16 *sh_hi = static_cast<const bf16*>(sv->h_hi), *sh_lo = static_cast<const bf16*>(sv->h_lo);
  const float keep = 1.0f - sv->dropout_p, inv_keep = 1.0f / keep;
  int64_t off[15];
  toad_param_offsets(d, off);
  float *g_w1 = grad + off[0], *g_b1 = grad + off[1], *g_w2 = grad + off[2], *g_b2 = grad + off[3];
  float *g_wa = grad + off[4], *g_ba = grad + off[5], *g_wb = grad + off[6], *g_bb = grad + off[7];
  float *g_wc = grad + off[8], *g_bc = grad + off[9], *g_wcls = grad + off[10], *g_bcls = grad + off[11];
  float *g_wsite = grad + off[12], *g_bsite = grad + off[13];

  // heads, softmax-pooling and gate backward: small HBM-bound kernels (shared with the fp32 path)
  bwd::heads_bwd_kernel<<<1, 2 * bwd::H, 0, st>>>(dlogits, dsite, fo->features, P->wcls, P->wsite, d->n_classes, g_wcls,
                                                   g_bcls, g_wsite, g_bsite, w.dM, w.sdot);
  TOAD_CUDA_TRY(cudaGetLastError());
  {
    int64_t blocks = (n + 7) / 8;
    if (blocks > 8 * kSMs) blocks = 8 * kSMs;
    bwd::pool_bwd_kernel<true><<<static_cast<unsigned>(blocks), 256, 0, st>>>(nullptr, sh_hi, sh_lo, fo->a_raw, fo->softmax_stats,
                                                                               w.dM, w.sdot, w.P, w.dA, n);
    TOAD_CUDA_TRY(cudaGetLastError());
  }
  {
    const int rpb = static_cast<int>((n + w.gate_blocks - 1) / w.gate_blocks);
    bwd::gate_bwd_kernel<true><<<w.gate_blocks, D, bwd::gate_bwd_smem(D), st>>>(sv->a, sv->b, w.dA, P->wc, nullptr, w.dab_hi, w.dab_lo,
                                                             w.gate_part, n, D, rpb, keep);
    TOAD_CUDA_TRY(cudaGetLastError());
    bwd::ReduceSegs segs{};
    segs.dst[0] = g_wc; segs.dst[1] = g_ba; segs.dst[2] = g_bb; segs.dst[3] = g_bc;
    segs.begin[0] = 0; segs.begin[1] = 2 * D; segs.begin[2] = 3 * D; segs.begin[3] = 4 * D; segs.begin[4] = 4 * D + 2;
    TOAD_TRY(bwd::launch_reduce_segs(w.gate_part, segs, 4 * D + 2, 4 * D + 2, w.gate_blocks, st));
  }
  // operand preparation: (hi, lo) planes (dab already left gate_bwd as planes)
  if (mn) {
    TOAD_TRY(tail::launch_split_planes(x, w.xp_hi, w.xp_lo, n * L, st));
  } else {
    TOAD_TRY(bwd::launch_transpose_planes(w.dab_hi, w.dab_lo, n, 2 * D, w.dabT_hi, w.dabT_lo, w.ldT, st));
    TOAD_TRY(bwd::launch_transpose_planes(sh_hi, sh_lo, n, Hd, w.hT_hi, w.hT_lo, w.ldT, st));
    TOAD_TRY(bwd::launch_transpose_planes(sh1_hi, sh1_lo, n, Hd, w.h1T_hi, w.h1T_lo, w.ldT, st));
    TOAD_TRY(bwd::launch_transpose_split(x, n, L, L, w.xT_hi, w.xT_lo, w.ldT, st));
  }
  TOAD_TRY(bwd::launch_transpose_split(P->w2, Hd, Hd, Hd, w.w2T_hi, w.w2T_lo, Hd, st));                  // W2^T   [in, out]
  TOAD_TRY(bwd::launch_transpose_split(P->wa, D, Hd, Hd, w.wabT_hi, w.wabT_lo, 2 * D, st));              // [Wa;Wb]^T [hid, 2D]
  TOAD_TRY(bwd::launch_transpose_split(P->wb, D, Hd, Hd, w.wabT_hi + D, w.wabT_lo + D, 2 * D, st));

  // dWa | dWb = dab^T . h
  if (mn) TOAD_TRY(wgrad_mn(w.dab_hi, w.dab_lo, 2 * D, sh_hi, sh_lo, Hd, 2 * D, Hd, n, w.splitk, g_wa, D, g_wb, st));
  else TOAD_TRY(wgrad_tc(w.dabT_hi, w.dabT_lo, w.hT_hi, w.hT_lo, 2 * D, Hd, n, w.ldT, w.splitk, g_wa, D, g_wb, st));
  // dz2 = (dab . [Wa;Wb] + P0 dM0 + P1 dM1) * (h > 0) / keep      -> planes
  {
    tc::GemmTcParams g{};
    g.M = n; g.N = Hd; g.K = 2 * D;
    g.pool_p = w.P; g.pool_v = w.dM; g.mask_bf16 = sh_hi; g.ld_mask = Hd; g.out_scale = inv_keep;
    g.out_hi = w.dz2_hi; g.out_lo = w.dz2_lo; g.ld_split = Hd;
    TOAD_TRY((tc::launch_gemm<256, tc::A_SPLIT, tc::EPI_DGRAD, 2>(g, w.dab_hi, w.dab_lo, w.wabT_hi, w.wabT_lo, st)));
  }
  TOAD_TRY(bwd::launch_colsum_planes(w.dz2_hi, w.dz2_lo, w.col_part, n, Hd, w.col_blocks, st));
  TOAD_TRY(bwd::launch_reduce_strided(w.col_part, g_b2, Hd, Hd, w.col_blocks, st));
  // dW2 = dz2^T . h1
  if (mn) {
    TOAD_TRY(wgrad_mn(w.dz2_hi, w.dz2_lo, Hd, sh1_hi, sh1_lo, Hd, Hd, Hd, n, w.splitk, g_w2, 0, nullptr, st));
  } else {
    TOAD_TRY(bwd::launch_transpose_planes(w.dz2_hi, w.dz2_lo, n, Hd, w.dz2T_hi, w.dz2T_lo, w.ldT, st));
    TOAD_TRY(wgrad_tc(w.dz2T_hi, w.dz2T_lo, w.h1T_hi, w.h1T_lo, Hd, Hd, n, w.ldT, w.splitk, g_w2, 0, nullptr, st));
  }
  // dz1 = (dz2 . W2) * (h1 > 0) / keep
  {
    tc::GemmTcParams g{};
    g.M = n; g.N = Hd; g.K = Hd;
    g.mask_bf16 = sh1_hi; g.ld_mask = Hd; g.out_scale = inv_keep;
    g.out_hi = w.dz1_hi; g.out_lo = w.dz1_lo; g.ld_split = Hd;
    TOAD_TRY((tc::launch_gemm<256, tc::A_SPLIT, tc::EPI_DGRAD, 2>(g, w.dz2_hi, w.dz2_lo, w.w2T_hi, w.w2T_lo, st)));
  }
  TOAD_TRY(bwd::launch_colsum_planes(w.dz1_hi, w.dz1_lo, w.col_part, n, Hd, w.col_blocks, st));
  TOAD_TRY(bwd::launch_reduce_strided(w.col_part, g_b1, Hd, Hd, w.col_blocks, st));
  // dW1 = dz1^T . x
  if (mn) return wgrad_mn(w.dz1_hi, w.dz1_lo, Hd, w.xp_hi, w.xp_lo, L, Hd, L, n, w.splitk, g_w1, 0, nullptr, st);
  TOAD_TRY(bwd::launch_transpose_planes(w.dz1_hi, w.dz1_lo, n, Hd, w.dz1T_hi, w.dz1T_lo, w.ldT, st));
  return wgrad_tc(w.dz1T_hi, w.dz1T_lo, w.xT_hi, w.xT_lo, Hd, L, n, w.ldT, w.splitk, g_w1, 0, nullptr, st);
}
}  // namespace

// the tensor-core wgrads tile the [out, in] gradients in 256 x 256 blocks (launch_gemm_mn<256, 2>)
static int check_bwd_tc_dims(const toad_dims_t* d, uint32_t flags) {
  if (flags & TOAD_FLAG_SIMT_FP32) return 0;
  if (d->in_dim % 256 != 0 || d->hid_dim % 256 != 0) return TOAD_ERR_UNSUPPORTED;
  return 0;
}

extern "C" int toad_bwd_workspace_bytes(const toad_dims_t* d, int64_t n, uint32_t flags, size_t* bytes) {
  TOAD_TRY(check_dims(d));
  TOAD_TRY(check_bwd_tc_dims(d, flags));
  if (bytes == nullptr || n <= 0) return TOAD_ERR_ARG;
  *bytes = (flags & TOAD_FLAG_SIMT_FP32) ? carve_bwd(d, n, nullptr).bytes : carve_bwd_tc(d, n, nullptr).bytes;
  return 0;
}

static int bwd_simt(const toad_dims_t* d, const toad_params_t* P, const float* x, int64_t n, const toad_fwd_out_t* fo,
                    const toad_saved_t* sv, const float* dlogits, const float* dsite, float* grad, void* workspace,
                    size_t workspace_bytes, toad_stream_t stream) {
  TOAD_TRY(check_dims(d));
  if (!P || !x || !fo || !sv || !dlogits || !dsite || !grad || n <= 0) return TOAD_ERR_ARG;
  if (!fo->a_raw || !fo->features || !fo->softmax_stats || !sv->h1 || !sv->h || !sv->a || !sv->b) return TOAD_ERR_ARG;
  BwdWs w = carve_bwd(d, n, workspace);
  TOAD_TRY(check_ws(workspace, workspace_bytes, w.bytes));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int Hd = d->hid_dim, D = d->attn_dim, L = d->in_dim;
  if (sv->dropout_p < 0.f || sv->dropout_p >= 1.f) return TOAD_ERR_ARG;
  const float keep = 1.0f - sv->dropout_p;  // saved activations are post-dropout: undo / re-apply the 1/keep scale
  const float inv_keep = 1.0f / keep;
  int64_t off[15];
  toad_param_offsets(d, off);
  float *g_w1 = grad + off[0], *g_b1 = grad + off[1], *g_w2 = grad + off[2], *g_b2 = grad + off[3];
  float *g_wa = grad + off[4], *g_ba = grad + off[5], *g_wb = grad + off[6], *g_bb = grad + off[7];
  float *g_wc = grad + off[8], *g_bc = grad + off[9], *g_wcls = grad + off[10], *g_bcls = grad + off[11];
  float *g_wsite = grad + off[12], *g_bsite = grad + off[13];

  // 1. heads -> dM, sdot ; 2. softmax-pooling backward -> P, dA
  bwd::heads_bwd_kernel<<<1, 2 * bwd::H, 0, st>>>(dlogits, dsite, fo->features, P->wcls, P->wsite, d->n_classes, g_wcls,
                                                   g_bcls, g_wsite, g_bsite, w.dM, w.sdot);
  TOAD_CUDA_TRY(cudaGetLastError());
  {
    int64_t blocks = (n + 7) / 8;
    if (blocks > 8 * kSMs) blocks = 8 * kSMs;
    bwd::pool_bwd_kernel<false><<<static_cast<unsigned>(blocks), 256, 0, st>>>(sv->h, nullptr, nullptr, fo->a_raw, fo->softmax_stats, w.dM, w.sdot,
                                                                         w.P, w.dA, n);
    TOAD_CUDA_TRY(cudaGetLastError());
  }
  // 3. gate backward -> dab, partials of dWc/dba/dbb/dbc
  {
    const int rpb = static_cast<int>((n + w.gate_blocks - 1) / w.gate_blocks);
    bwd::gate_bwd_kernel<false><<<w.gate_blocks, D, bwd::gate_bwd_smem(D), st>>>(sv->a, sv->b, w.dA, P->wc, w.dab, nullptr, nullptr,
                                                              w.gate_part, n, D, rpb, keep);
    TOAD_CUDA_TRY(cudaGetLastError());
    const int64_t stride = 4 * D + 2;
    TOAD_TRY(bwd::launch_reduce_strided(w.gate_part, g_wc, 2 * D, stride, w.gate_blocks, st));
    TOAD_TRY(bwd::launch_reduce_strided(w.gate_part + 2 * D, g_ba, D, stride, w.gate_blocks, st));
    TOAD_TRY(bwd::launch_reduce_strided(w.gate_part + 3 * D, g_bb, D, stride, w.gate_blocks, st));
    TOAD_TRY(bwd::launch_reduce_strided(w.gate_part + 4 * D, g_bc, 2, stride, w.gate_blocks, st));
  }
  // 4. dWa | dWb = dab^T . h   (K = patches, split-K)
  {
    simt::SgemmParams s{};
    s.a = w.dab; s.a_rs = 1; s.a_ks = 2 * D;
    s.b = sv->h; s.b_rs = 1; s.b_ks = Hd;
    s.c = w.splitk; s.ldc = Hd; s.M = 2 * D; s.N = Hd; s.K = n; s.k_chunk = w.k_chunk;
    TOAD_TRY((simt::launch_sgemm<false, false, simt::EPI_STORE>(s, w.splits, st)));
    const int64_t stride = static_cast<int64_t>(2) * D * Hd;
    TOAD_TRY(bwd::launch_reduce_strided(w.splitk, g_wa, static_cast<int64_t>(D) * Hd, stride, w.splits, st));
    TOAD_TRY(bwd::launch_reduce_strided(w.splitk + static_cast<int64_t>(D) * Hd, g_wb, static_cast<int64_t>(D) * Hd, stride, w.splits, st));
  }
  // 5. dz2 = (da_pre.Wa + db_pre.Wb + P0 dM0 + P1 dM1) * (h > 0)
  {
    simt::SgemmParams s{};
    s.a = w.dab; s.a_rs = 2 * D; s.a_ks = 1;
    s.b = P->wa; s.b_rs = 1; s.b_ks = Hd;
    s.c = w.dz2; s.ldc = Hd; s.M = n; s.N = Hd; s.K = D; s.k_chunk = D;
    TOAD_TRY((simt::launch_sgemm<true, false, simt::EPI_STORE>(s, 1, st)));
    s.a = w.dab + D; s.b = P->wb;
    s.accumulate = 1; s.mask = sv->h; s.ldmask = Hd; s.out_scale = inv_keep;
    s.p0 = w.P; s.p1 = w.P + 1; s.p_stride = 2; s.v0 = w.dM; s.v1 = w.dM + Hd;
    TOAD_TRY((simt::launch_sgemm<true, false, simt::EPI_POOL_RELUMASK>(s, 1, st)));
  }
  // 6. db2, dW2 = dz2^T . h1
  {
    const int rpb = static_cast<int>((n + w.col_blocks - 1) / w.col_blocks);
    bwd::colsum_kernel<<<w.col_blocks, Hd, 0, st>>>(w.dz2, w.col_part, n, Hd, rpb);
    TOAD_CUDA_TRY(cudaGetLastError());
    TOAD_TRY(bwd::launch_reduce_strided(w.col_part, g_b2, Hd, Hd, w.col_blocks, st));
    simt::SgemmParams s{};
    s.a = w.dz2; s.a_rs = 1; s.a_ks = Hd;
    s.b = sv->h1; s.b_rs = 1; s.b_ks = Hd;
    s.c = w.splitk; s.ldc = Hd; s.M = Hd; s.N = Hd; s.K = n; s.k_chunk = w.k_chunk;
    TOAD_TRY((simt::launch_sgemm<false, false, simt::EPI_STORE>(s, w.splits, st)));
    TOAD_TRY(bwd::launch_reduce_strided(w.splitk, g_w2, static_cast<int64_t>(Hd) * Hd, static_cast<int64_t>(Hd) * Hd, w.splits, st));
  }
  // 7. dz1 = (dz2 . W2) * (h1 > 0)
  {
    simt::SgemmParams s{};
    s.a = w.dz2; s.a_rs = Hd; s.a_ks = 1;
    s.b = P->w2; s.b_rs = 1; s.b_ks = Hd;
    s.c = w.dz1; s.ldc = Hd; s.M = n; s.N = Hd; s.K = Hd; s.k_chunk = Hd;
    s.mask = sv->h1; s.ldmask = Hd; s.out_scale = inv_keep;
    TOAD_TRY((simt::launch_sgemm<true, false, simt::EPI_RELUMASK>(s, 1, st)));
  }
  // 8. db1, dW1 = dz1^T . x
  {
    const int rpb = static_cast<int>((n + w.col_blocks - 1) / w.col_blocks);
    bwd::colsum_kernel<<<w.col_blocks, Hd, 0, st>>>(w.dz1, w.col_part, n, Hd, rpb);
    TOAD_CUDA_TRY(cudaGetLastError());
    TOAD_TRY(bwd::launch_reduce_strided(w.col_part, g_b1, Hd, Hd, w.col_blocks, st));
    simt::SgemmParams s{};
    s.a = w.dz1; s.a_rs = 1; s.a_ks = Hd;
    s.b = x; s.b_rs = 1; s.b_ks = L;
    s.c = w.splitk; s.ldc = L; s.M = Hd; s.N = L; s.K = n; s.k_chunk = w.k_chunk;
    TOAD_TRY((simt::launch_sgemm<false, false, simt::EPI_STORE>(s, w.splits, st)));
    TOAD_TRY(bwd::launch_reduce_strided(w.splitk, g_w1, static_cast<int64_t>(Hd) * L, static_cast<int64_t>(Hd) * L, w.splits, st));
  }
  return 0;
}

extern "C" int toad_bwd(const toad_dims_t* d, const toad_params_t* P, const float* x, int64_t n, const toad_fwd_out_t* fo,
                        const toad_saved_t* sv, const float* dlogits, const float* dsite, float* grad, void* workspace,
                        size_t workspace_bytes, uint32_t flags, toad_stream_t stream) {
  TOAD_TRY(check_dims(d));
  TOAD_TRY(check_bwd_tc_dims(d, flags));
  if (!P || !x || !fo || !sv || !dlogits || !dsite || !grad || n <= 0) return TOAD_ERR_ARG;
  if (!fo->a_raw || !fo->features || !fo->softmax_stats || !sv->a || !sv->b) return TOAD_ERR_ARG;
  if (flags & TOAD_FLAG_SIMT_FP32) {
    if (!sv->h1 || !sv->h) return TOAD_ERR_ARG;
  } else if (!sv->h1_hi || !sv->h1_lo || !sv->h_hi || !sv->h_lo) {
    return TOAD_ERR_ARG;
  }
  if (flags & TOAD_FLAG_SIMT_FP32) return bwd_simt(d, P, x, n, fo, sv, dlogits, dsite, grad, workspace, workspace_bytes, stream);
  return bwd_tc(d, P, x, n, fo, sv, dlogits, dsite, grad, workspace, workspace_bytes, flags, stream);
}

// ------------------------------------------------------------------------------------------ loss + optimizer
extern "C" int toad_ce_loss_grad(const float* logits, const float* site_logits, int32_t n_classes, const int64_t* label,
                                 const int64_t* site, float w_cls, float w_site, float* loss3, float* dlogits,
                                 float* dsite_logits, toad_stream_t stream) {
  if (!logits || !site_logits || !label || !site || !loss3 || !dlogits || !dsite_logits) return TOAD_ERR_ARG;
  if (n_classes < 1 || n_classes > 1024) return TOAD_ERR_UNSUPPORTED;
  train::ce_loss_grad_kernel<<<1, 64, 0, static_cast<cudaStream_t>(stream)>>>(logits, site_logits, n_classes, label, site,
                                                                              w_cls, w_site, loss3, dlogits, dsite_logits);
  TOAD_CUDA_TRY(cudaGetLastError());
  return 0;
}

extern "C" int toad_adam_step(const toad_dims_t* d, const toad_params_t* P, const float* grad_flat, float* exp_avg,
                              float* exp_avg_sq, int64_t step, double lr, double beta1, double beta2, double eps,
                              double weight_decay, float grad_scale, toad_stream_t stream) {
  TOAD_TRY(check_dims(d));
  if (!P || !grad_flat || !exp_avg || !exp_avg_sq || step < 1) return TOAD_ERR_ARG;
  if (!(lr >= 0.0) || !(beta1 >= 0.0 && beta1 < 1.0) || !(beta2 >= 0.0 && beta2 < 1.0) || !(eps >= 0.0) ||
      !(weight_decay >= 0.0))
    return TOAD_ERR_ARG;
  train::AdamArgs a{};
  // the one entry point that WRITES the parameters (toad_params_t is const for every other caller)
  const float* const ptrs[14] = {P->w1, P->b1, P->w2, P->b2, P->wa, P->ba, P->wb, P->bb, P->wc, P->bc, P->wcls, P->bcls, P->wsite, P->bsite};
  for (int i = 0; i < 14; ++i) {
    if (ptrs[i] == nullptr) return TOAD_ERR_ARG;
    a.p[i] = const_cast<float*>(ptrs[i]);
  }
  toad_param_offsets(d, a.off);
  a.g = grad_flat; a.m = exp_avg; a.v = exp_avg_sq;
  // hyper-parameters arrive as doubles (python floats on the reference side); derived constants are formed in
  // double and rounded to fp32 once, which is what torch does with its python scalars
  a.grad_scale = grad_scale; a.wd = static_cast<float>(weight_decay);
  a.b1 = static_cast<float>(beta1); a.b2 = static_cast<float>(beta2); a.eps = static_cast<float>(eps);
  a.one_m_b1 = static_cast<float>(1.0 - beta1);
  a.one_m_b2 = static_cast<float>(1.0 - beta2);
  // bias corrections in double on the host, like torch's python scalars (torch/optim/adam.py _single_tensor_adam)
  const double bc1 = 1.0 - std::pow(beta1, static_cast<double>(step));
  const double bc2 = 1.0 - std::pow(beta2, static_cast<double>(step));
  a.step_size = static_cast<float>(lr / bc1);
  a.bc2_sqrt = static_cast<float>(std::sqrt(bc2));
  const int64_t total = a.off[14];
  int64_t blocks = (total + 255) / 256;
  if (blocks > 8 * kSMs) blocks = 8 * kSMs;
  train::adam_kernel<<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(a);
  TOAD_CUDA_TRY(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------------ Attn_Net_Gated
namespace {
struct AgWs {
  bf16 *w_hi, *w_lo;
  float *part, *a, *b;
  int n_parts;
  size_t bytes;
};
AgWs carve_ag(int L, int D, int nt, int64_t n, uint32_t flags, void* base) {
  AgWs w{};
  Carver c(base);
  const bool simt = (flags & TOAD_FLAG_SIMT_FP32) != 0;
  w.n_parts = simt ? 1 : D / kGateHalf;  // fp32-fed gate kernel: one epilogue warp set
  w.part = c.take<float>(static_cast<size_t>(w.n_parts) * n * nt);
  if (simt) {
    w.a = c.take<float>(n * D);
    w.b = c.take<float>(n * D);
  } else {
    w.w_hi = c.take<bf16>(static_cast<size_t>(2) * D * L);
    w.w_lo = c.take<bf16>(static_cast<size_t>(2) * D * L);
  }
  w.bytes = align_up(c.off, 256);
  return w;
}
int check_ag(int L, int D, int nt, int64_t n) {
  if (n <= 0) return TOAD_ERR_ARG;
  if (L <= 0 || L % 64 != 0 || D <= 0 || D % kGateHalf != 0 || nt < 1 || nt > 4) return TOAD_ERR_UNSUPPORTED;
  return 0;
}
}  // namespace

extern "C" int toad_attn_gated_workspace_bytes(int32_t L, int32_t D, int32_t nt, int64_t n, uint32_t flags, size_t* bytes) {
  TOAD_TRY(check_ag(L, D, nt, n));
  if (bytes == nullptr) return TOAD_ERR_ARG;
  *bytes = carve_ag(L, D, nt, n, flags, nullptr).bytes;
  return 0;
}

extern "C" int toad_attn_gated_fwd(int32_t L, int32_t D, int32_t nt, const float* wa, const float* ba, const float* wb,
                        const float* bb, const float* wc, const float* bc, const float* x, int64_t n, float* A_out,
                        const toad_attn_saved_t* saved, void* workspace, size_t workspace_bytes, uint32_t flags,
                        toad_stream_t stream) {
  TOAD_TRY(check_ag(L, D, nt, n));
  if (!wa || !ba || !wb || !bb || !wc || !bc || !x || !A_out) return TOAD_ERR_ARG;
  if (saved != nullptr && (!saved->a || !saved->b || (flags & TOAD_FLAG_SIMT_FP32))) return saved->a && saved->b ? TOAD_ERR_UNSUPPORTED : TOAD_ERR_ARG;
  if ((flags & TOAD_FLAG_DROPOUT) && (saved == nullptr || saved->dropout_p < 0.f || saved->dropout_p >= 1.f)) return TOAD_ERR_ARG;
  AgWs w = carve_ag(L, D, nt, n, flags, workspace);
  TOAD_TRY(check_ws(workspace, workspace_bytes, w.bytes));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (flags & TOAD_FLAG_SIMT_FP32) {
    TOAD_TRY((simt::launch_sgemm<true, true, simt::EPI_BIAS_TANH>(linear_params(x, L, wa, ba, w.a, n, D, L), 1, st)));
    TOAD_TRY((simt::launch_sgemm<true, true, simt::EPI_BIAS_SIGMOID>(linear_params(x, L, wb, bb, w.b, n, D, L), 1, st)));
    TOAD_TRY(tail::launch_attn_c(w.a, w.b, wc, w.part, n, D, nt, st));
  } else {
    TOAD_TRY(tail::launch_split_gate_weights(wa, wb, w.w_hi, w.w_lo, D, L, kGateHalf, st));
    tc::GemmTcParams g{};
    g.a_f32 = x; g.lda = L; g.M = n; g.N = 2 * D; g.K = L;
    g.gate_ba = ba; g.gate_bb = bb; g.gate_wc = wc; g.gate_D = D; g.gate_ntasks = nt; g.gate_part = w.part;
    if (saved != nullptr) {
      g.gate_a = saved->a; g.gate_b = saved->b;
      if ((flags & TOAD_FLAG_DROPOUT) && saved->dropout_p > 0.f) {
        g.drop.seed = saved->dropout_seed;
        const double t = static_cast<double>(saved->dropout_p) * 4294967296.0;
        g.drop.thresh = t >= 4294967295.0 ? 0xFFFFFFFFu : static_cast<uint32_t>(t);
        g.drop.scale = 1.0f / (1.0f - saved->dropout_p);
      }
    }
    if (flags & TOAD_FLAG_TC_SINGLE_CTA) TOAD_TRY((tc::launch_gemm<256, tc::A_F32, tc::EPI_GATE, 1>(g, nullptr, nullptr, w.w_hi, w.w_lo, st)));
    else TOAD_TRY((tc::launch_gemm<256, tc::A_F32, tc::EPI_GATE, 2>(g, nullptr, nullptr, w.w_hi, w.w_lo, st)));
  }
  return tail::launch_finish_scores(w.part, w.n_parts, bc, A_out, n, nt, st);
}

namespace {
struct AgBwdWs {
  float *dA2, *wc2, *gtmp, *splitk, *gate_part;
  bf16 *dab_hi, *dab_lo, *xp_hi, *xp_lo, *wabT_hi, *wabT_lo;
  int gate_blocks;
  size_t bytes;
};
AgBwdWs carve_ag_bwd(int L, int D, int64_t n, void* base) {
  AgBwdWs w{};
  Carver c(base);
  w.dA2 = c.take<float>(n * 2);
  w.wc2 = c.take<float>(2 * D);
  w.gtmp = c.take<float>(4 * D + 2);
  w.dab_hi = c.take<bf16>(n * 2 * D); w.dab_lo = c.take<bf16>(n * 2 * D);
  w.xp_hi = c.take<bf16>(n * L);      w.xp_lo = c.take<bf16>(n * L);
  w.wabT_hi = c.take<bf16>(static_cast<size_t>(L) * 2 * D); w.wabT_lo = c.take<bf16>(static_cast<size_t>(L) * 2 * D);
  const size_t pair_tiles = static_cast<size_t>(kSMs / 2) * 256 * 256, full = static_cast<size_t>(2) * D * L;
  w.splitk = c.take<float>(pair_tiles > full ? pair_tiles : full);
  int64_t gb = (n + 127) / 128;
  if (gb > 2 * kSMs) gb = 2 * kSMs;
  w.gate_blocks = static_cast<int>(gb);
  w.gate_part = c.take<float>(static_cast<size_t>(gb) * (4 * D + 2));
  w.bytes = align_up(c.off, 256);
  return w;
}
// [n, nt] -> [n, 2] (missing task columns zero): the gate backward kernels are written for the two tasks of TOAD
__global__ void pad_tasks_kernel(const float* __restrict__ src, float* __restrict__ dst, int64_t n, int nt) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  dst[2 * i] = src[i * nt];
  dst[2 * i + 1] = nt > 1 ? src[i * nt + 1] : 0.f;
}
int check_ag_bwd(int L, int D, int nt, int64_t n) {
  TOAD_TRY(check_ag(L, D, nt, n));
  if (nt > 2 || L % 256 != 0) return TOAD_ERR_UNSUPPORTED;
  return 0;
}
}  // namespace

extern "C" int toad_attn_gated_bwd_workspace_bytes(int32_t L, int32_t D, int32_t nt, int64_t n, size_t* bytes) {
  TOAD_TRY(check_ag_bwd(L, D, nt, n));
  if (bytes == nullptr) return TOAD_ERR_ARG;
  *bytes = carve_ag_bwd(L, D, n, nullptr).bytes;
  return 0;
}

extern "C" int toad_attn_gated_bwd(int32_t L, int32_t D, int32_t nt, const float* wa, const float* wb, const float* wc,
                                   const float* x, int64_t n, const toad_attn_saved_t* sv, const float* dA, float* d_wa,
                                   float* d_ba, float* d_wb, float* d_bb, float* d_wc, float* d_bc, float* dx,
                                   void* workspace, size_t workspace_bytes, toad_stream_t stream) {
  TOAD_TRY(check_ag_bwd(L, D, nt, n));
  if (!wa || !wb || !wc || !x || !sv || !sv->a || !sv->b || !dA || !d_wa || !d_ba || !d_wb || !d_bb || !d_wc || !d_bc) return TOAD_ERR_ARG;
  if (sv->dropout_p < 0.f || sv->dropout_p >= 1.f) return TOAD_ERR_ARG;
  AgBwdWs w = carve_ag_bwd(L, D, n, workspace);
  TOAD_TRY(check_ws(workspace, workspace_bytes, w.bytes));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const float keep = 1.0f - sv->dropout_p;
  // two-task views of dA and Wc (zero second task when n_tasks == 1)
  pad_tasks_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, st>>>(dA, w.dA2, n, nt);
  TOAD_CUDA_TRY(cudaGetLastError());
  TOAD_CUDA_TRY(cudaMemsetAsync(w.wc2, 0, 2 * D * sizeof(float), st));
  TOAD_CUDA_TRY(cudaMemcpyAsync(w.wc2, wc, static_cast<size_t>(nt) * D * sizeof(float), cudaMemcpyDeviceToDevice, st));
  // gate derivative: dab = (d tanh-branch pre-activation | d sigmoid-branch pre-activation) as (hi, lo) planes,
  // per-CTA partial sums of dWc, dba, dbb, dbc
  {
    const int rpb = static_cast<int>((n + w.gate_blocks - 1) / w.gate_blocks);
    bwd::gate_bwd_kernel<true><<<w.gate_blocks, D, bwd::gate_bwd_smem(D), st>>>(sv->a, sv->b, w.dA2, w.wc2, nullptr, w.dab_hi, w.dab_lo,
                                                                                 w.gate_part, n, D, rpb, keep);
    TOAD_CUDA_TRY(cudaGetLastError());
    bwd::ReduceSegs segs{};
    segs.dst[0] = w.gtmp; segs.dst[1] = d_ba; segs.dst[2] = d_bb; segs.dst[3] = w.gtmp + 2 * D;
    segs.begin[0] = 0; segs.begin[1] = 2 * D; segs.begin[2] = 3 * D; segs.begin[3] = 4 * D; segs.begin[4] = 4 * D + 2;
    TOAD_TRY(bwd::launch_reduce_segs(w.gate_part, segs, 4 * D + 2, 4 * D + 2, w.gate_blocks, st));
    TOAD_CUDA_TRY(cudaMemcpyAsync(d_wc, w.gtmp, static_cast<size_t>(nt) * D * sizeof(float), cudaMemcpyDeviceToDevice, st));
    TOAD_CUDA_TRY(cudaMemcpyAsync(d_bc, w.gtmp + 2 * D, static_cast<size_t>(nt) * sizeof(float), cudaMemcpyDeviceToDevice, st));
  }
  // dWa | dWb = dab^T . x   (both operands MN-major from their natural [patch, channel] planes, split-K)
  TOAD_TRY(tail::launch_split_planes(x, w.xp_hi, w.xp_lo, n * L, st));
  TOAD_TRY(wgrad_mn(w.dab_hi, w.dab_lo, 2 * D, w.xp_hi, w.xp_lo, L, 2 * D, L, n, w.splitk, d_wa, D, d_wb, st));
  if (dx != nullptr) {  // dx = dab . [Wa ; Wb]
    TOAD_TRY(bwd::launch_transpose_split(wa, D, L, L, w.wabT_hi, w.wabT_lo, 2 * D, st));
    TOAD_TRY(bwd::launch_transpose_split(wb, D, L, L, w.wabT_hi + D, w.wabT_lo + D, 2 * D, st));
    tc::GemmTcParams g{};
    g.M = n; g.N = L; g.K = 2 * D; g.out_f32 = dx; g.ld_f32 = L;
    TOAD_TRY((tc::launch_gemm<256, tc::A_SPLIT, tc::EPI_LINEAR, 2>(g, w.dab_hi, w.dab_lo, w.wabT_hi, w.wabT_lo, st)));
  }
  return 0;
}

// ------------------------------------------------------------------------------------------ top-k
extern "C" int toad_topk_workspace_bytes(int64_t n, int32_t k, size_t* bytes) {
  if (bytes == nullptr || n <= 0 || k <= 0) return TOAD_ERR_ARG;
  *bytes = topk::topk_workspace_bytes();  // global histograms, per-CTA tie counts, the k winners (independent of n, k)
  return 0;
}

extern "C" int toad_gather_rows(const void* table, int64_t n_rows, int32_t row_bytes, const int64_t* idx, int32_t k,
                                void* out, toad_stream_t stream) {
  if (!table || !idx || !out) return TOAD_ERR_ARG;
  return topk::launch_gather_rows(table, n_rows, row_bytes, idx, k, out, static_cast<cudaStream_t>(stream));
}

extern "C" int toad_topk(const float* scores, int64_t n, int32_t k, float* out_vals, int64_t* out_idx, void* workspace,
              size_t workspace_bytes, toad_stream_t stream) {
  if (!scores || !out_vals || !out_idx || n <= 0 || k <= 0) return TOAD_ERR_ARG;
  return topk::launch_topk(scores, n, k, out_vals, out_idx, workspace, workspace_bytes, tc::sm_count(), static_cast<cudaStream_t>(stream));
}

// ------------------------------------------------------------------------------------------ linear (test hook)
namespace {
struct LinWs { bf16 *w_hi, *w_lo, *x_hi, *x_lo; size_t bytes; };
LinWs carve_lin(int64_t m, int n, int k, void* base) {
  LinWs w{};
  Carver c(base);
  w.w_hi = c.take<bf16>(static_cast<size_t>(n) * k);
  w.w_lo = c.take<bf16>(static_cast<size_t>(n) * k);
  w.x_hi = c.take<bf16>(static_cast<size_t>(m) * k);
  w.x_lo = c.take<bf16>(static_cast<size_t>(m) * k);
  w.bytes = align_up(c.off, 256);
  return w;
}
template <int BN, int CG>
int run_linear(const tc::GemmTcParams& g, bool split_a, const LinWs& w, cudaStream_t st) {
  if (split_a) return tc::launch_gemm<BN, tc::A_SPLIT, tc::EPI_LINEAR, CG>(g, w.x_hi, w.x_lo, w.w_hi, w.w_lo, st);
  return tc::launch_gemm<BN, tc::A_F32, tc::EPI_LINEAR, CG>(g, nullptr, nullptr, w.w_hi, w.w_lo, st);
}
}  // namespace

extern "C" int toad_linear_workspace_bytes(int64_t m, int32_t n, int32_t k, size_t* bytes) {
  if (bytes == nullptr || m <= 0 || n <= 0 || k <= 0) return TOAD_ERR_ARG;
  *bytes = carve_lin(m, n, k, nullptr).bytes;
  return 0;
}

extern "C" int toad_linear_bf16x3(const float* x, const float* wgt, const float* bias, float* y, int64_t m, int32_t n, int32_t k,
                       int32_t relu, int32_t variant, void* workspace, size_t workspace_bytes, toad_stream_t stream) {
  if (!x || !wgt || !y || m <= 0 || n <= 0 || k <= 0) return TOAD_ERR_ARG;
  if (k % 64 != 0 || n % 64 != 0) return TOAD_ERR_UNSUPPORTED;
  LinWs w = carve_lin(m, n, k, workspace);
  TOAD_TRY(check_ws(workspace, workspace_bytes, w.bytes));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const bool split_a = (variant & 1) != 0;
  int bn = (variant >> 4) & 3;
  if (bn == 0) bn = (n % 256 == 0) ? 3 : (n % 128 == 0 ? 2 : 1);
  const int BN = bn == 3 ? 256 : (bn == 2 ? 128 : 64);
  if (n % BN != 0) return TOAD_ERR_UNSUPPORTED;
  TOAD_TRY(tail::launch_split_planes(wgt, w.w_hi, w.w_lo, static_cast<int64_t>(n) * k, st));
  if (split_a) TOAD_TRY(tail::launch_split_planes(x, w.x_hi, w.x_lo, m * k, st));
  tc::GemmTcParams g{};
  g.a_f32 = x; g.lda = k; g.M = m; g.N = n; g.K = k; g.bias = bias; g.relu = relu; g.out_f32 = y; g.ld_f32 = n;
  const bool pair = (variant & 2) != 0;  // cta_group::2 (CTA pair per 256-row tile)
  if (variant & 0x40) {              // 512-wide tiles (pairs only)
    if (!pair || n % 512 != 0) return TOAD_ERR_UNSUPPORTED;
    return run_linear<512, 2>(g, split_a, w, st);
  }
  if (BN == 256) return pair ? run_linear<256, 2>(g, split_a, w, st) : run_linear<256, 1>(g, split_a, w, st);
  if (BN == 128) return pair ? run_linear<128, 2>(g, split_a, w, st) : run_linear<128, 1>(g, split_a, w, st);
  return pair ? run_linear<64, 2>(g, split_a, w, st) : run_linear<64, 1>(g, split_a, w, st);
}


// ------------------------------------------------------------------------------------------ ResNet-50 trunk
namespace {

struct ConvSpec { int cin, cout, k, stride; };
constexpr int kResOutBufs = 2;
#ifndef TOAD_RESNET_CHUNK
#define TOAD_RESNET_CHUNK 256
#endif
// images per stem + layer1 pass.  Small chunks were meant to keep layer1's activations in L2; measured at batch 256
// the opposite holds (16: 21.3k, 32: 23.6k, 64: 24.8k, 128: 25.1k, 256: 25.7k patches/s): fuller waves win.
constexpr int kResChunk = TOAD_RESNET_CHUNK;

// the 43 convolutions in state_dict order (resnet_custom.py:57-94 with layers [3,4,6])
int build_specs(ConvSpec* specs) {
  int n = 0;
  specs[n++] = {3, 64, 7, 2};
  int inplanes = 64;
  const int planes[3] = {64, 128, 256}, blocks[3] = {3, 4, 6}, strides[3] = {1, 2, 2};
  for (int l = 0; l < 3; ++l) {
    for (int i = 0; i < blocks[l]; ++i) {
      const int s = i == 0 ? strides[l] : 1;
      specs[n++] = {inplanes, planes[l], 1, 1};
      specs[n++] = {planes[l], planes[l], 3, s};
      specs[n++] = {planes[l], planes[l] * 4, 1, 1};
      if (i == 0) specs[n++] = {inplanes, planes[l] * 4, 1, s};  // downsample
      inplanes = planes[l] * 4;
    }
  }
  return n;  // 43
}

inline int kpad_of(const ConvSpec& c) {
  const int k = c.k * c.k * c.cin;
  return (k + 63) / 64 * 64;
}

struct PreparedConv { bf16* hi; bf16* lo; float* bias; };
struct Prepared { PreparedConv conv[43]; size_t bytes; };

Prepared carve_prepared(void* base) {
  Prepared p{};
  Carver c(base);
  ConvSpec specs[43];
  build_specs(specs);
  for (int i = 0; i < 43; ++i) {
    const size_t n = static_cast<size_t>(specs[i].cout) * kpad_of(specs[i]);
    p.conv[i].hi = c.take<bf16>(n);
    p.conv[i].lo = c.take<bf16>(n);
    p.conv[i].bias = c.take<float>(specs[i].cout);
  }
  p.bytes = align_up(c.off, 256);
  return p;
}

struct ResWs {
  bf16 *col_hi, *col_lo, *stem_hi, *stem_lo, *pool_hi, *pool_lo;
  bf16 *buf_hi[5], *buf_lo[5];
  bf16* pad_hi;   // fp16 mode: zero-bordered plane feeding the halo 3x3 convolutions (conv3x3_halo.cuh)
  int stem_chunk;
  size_t bytes;
};

// images per stem + layer1 pass (TOAD_RESNET_CHUNK in the environment overrides the default: tuning aid)
int resnet_chunk() {
  static int v = []() {
    const char* e = getenv("TOAD_RESNET_CHUNK");
    const int c = e != nullptr ? atoi(e) : 0;
    return c > 0 ? c : kResChunk;
  }();
  return v;
}

// TOAD_RESNET_STEM_IM2COL=1: the fp16 mode's stem through the explicit im2col plane + GEMM too (cross-check / A-B aid)
bool stem_im2col_forced() {
  static bool v = []() {
    const char* e = getenv("TOAD_RESNET_STEM_IM2COL");
    return e != nullptr && e[0] == '1';
  }();
  return v;
}

// TOAD_RESNET_PDL=0: the trunk's launches fully serialised (A/B aid)
bool resnet_pdl() {
  static bool v = []() {
    const char* e = getenv("TOAD_RESNET_PDL");
    return !(e != nullptr && e[0] == '0');
  }();
  return v;
}

// TOAD_RESNET_STEM_POOL=0: fused stem without the fused max-pool (A/B and cross-check aid)
bool stem_pool_unfused() {
  static bool v = []() {
    const char* e = getenv("TOAD_RESNET_STEM_POOL");
    return e != nullptr && e[0] == '0';
  }();
  return v;
}

// TOAD_RESNET_HALO=0: the 3x3 convolutions of layer1 / layer2 as tap-by-tap implicit GEMMs (A/B and cross-check aid)
bool resnet_halo() {
  static bool v = []() {
    const char* e = getenv("TOAD_RESNET_HALO");
    return !(e != nullptr && e[0] == '0');
  }();
  return v;
}

// exact = (hi, lo) bf16 plane pairs (4 B / element); default = one fp16 plane (2 B / element, lo pointers stay null)
ResWs carve_resnet(int B, int H, int W, bool exact, void* base) {
  ResWs w{};
  Carver c(base);
  const int64_t H1 = H / 2, W1 = W / 2, H2 = H / 4, W2 = W / 4;
  w.stem_chunk = B < resnet_chunk() ? B : resnet_chunk();
  if (exact || stem_im2col_forced()) {  // (the fp16 mode's stem is an implicit GEMM: no im2col plane)
    const size_t col = static_cast<size_t>(w.stem_chunk) * H1 * W1 * resnet::STEM_KPAD;
    w.col_hi = c.take<bf16>(col);
    if (exact) w.col_lo = c.take<bf16>(col);
  }
  const size_t stem = static_cast<size_t>(w.stem_chunk) * H1 * W1 * 64;
  w.stem_hi = c.take<bf16>(stem);
  if (exact) w.stem_lo = c.take<bf16>(stem);
  w.pool_hi = c.take<bf16>(stem / 4);
  if (exact) w.pool_lo = c.take<bf16>(stem / 4);
  const size_t act = static_cast<size_t>(B) * H2 * W2 * 256;
  for (int i = 0; i < 5; ++i) {
    w.buf_hi[i] = c.take<bf16>(act);
    if (exact) w.buf_lo[i] = c.take<bf16>(act);
  }
  if (!exact) {  // the larger of layer1's (per chunk, 64 channels) and layer2's (whole batch, 128 channels) padded planes
    const size_t p1 = static_cast<size_t>(tc::pad_positions(w.stem_chunk, static_cast<int>(H2), static_cast<int>(W2))) * 64;
    const size_t p2 = static_cast<size_t>(tc::pad_positions(B, static_cast<int>(H2 / 2), static_cast<int>(W2 / 2))) * 128;
    w.pad_hi = c.take<bf16>(p1 > p2 ? p1 : p2);
  }
  w.bytes = align_up(c.off, 256);
  return w;
}

// Any H, W that are multiples of 16 up to 512 (the trunk halves the resolution four times; an M tile of the 3x3 /
// strided convolutions spans full output rows of at most 128 pixels).  224 x 224 and 256 x 256 are the usual patches.
int check_resnet_shape(int B, int H, int W) {
  if (B <= 0 || H <= 0 || W <= 0) return TOAD_ERR_ARG;
  if (H % 16 != 0 || W % 16 != 0 || W > 512 || H > 4096) return TOAD_ERR_UNSUPPORTED;
  return 0;
}

// out planes [M, Cout] = act(conv(in) + bias (+ residual)); in: NHWC planes [B, H, W, Cin]
template <int PREC>
int run_conv(const ConvSpec& cs, const PreparedConv& pc, const bf16* in_hi, const bf16* in_lo, int B, int H, int W,
             bf16* out_hi, bf16* out_lo, const bf16* res_hi, const bf16* res_lo, bool relu, cudaStream_t st,
             bool out_padded = false) {
  tc::GemmTcParams g{};
  if (out_padded) { g.out_pad_H = H; g.out_pad_W = W; }   // (1x1 / stride 1 only: checked by launch_gemm)
  g.N = cs.cout;
  g.bias = pc.bias;
  g.relu = relu ? 1 : 0;
  g.out_hi = out_hi; g.out_lo = out_lo; g.ld_split = cs.cout;
  g.res_hi = res_hi; g.res_lo = res_lo; g.ld_res = cs.cout;
  g.pdl = resnet_pdl() ? 1 : 0;   // the next convolution's prologue overlaps this one's last wave
  // OUT_BUFS = 2: the trunk's layers are store-/epilogue-bound (short K, wide outputs), so the TMA-store staging
  // is double buffered at the price of one operand stage.
  if (cs.k == 1 && cs.stride == 1) {  // plain GEMM over the NHWC plane
    g.M = static_cast<int64_t>(B) * H * W;
    g.K = cs.cin;
    if (cs.cout % 256 == 0) return tc::launch_gemm<256, tc::A_SPLIT, tc::EPI_LINEAR, 2, kResOutBufs, PREC>(g, in_hi, in_lo, pc.hi, pc.lo, st);
    if (cs.cout % 128 == 0) return tc::launch_gemm<128, tc::A_SPLIT, tc::EPI_LINEAR, 2, kResOutBufs, PREC>(g, in_hi, in_lo, pc.hi, pc.lo, st);
    return tc::launch_gemm<64, tc::A_SPLIT, tc::EPI_LINEAR, 2, kResOutBufs, PREC>(g, in_hi, in_lo, pc.hi, pc.lo, st);
  }
  if (cs.cout % 256 == 0) return tc::launch_conv_gemm<256, 2, kResOutBufs, PREC>(g, in_hi, in_lo, B, H, W, cs.cin, cs.k, cs.stride, pc.hi, pc.lo, st);
  if (cs.cout % 128 == 0) return tc::launch_conv_gemm<128, 2, kResOutBufs, PREC>(g, in_hi, in_lo, B, H, W, cs.cin, cs.k, cs.stride, pc.hi, pc.lo, st);
  return tc::launch_conv_gemm<64, 2, kResOutBufs, PREC>(g, in_hi, in_lo, B, H, W, cs.cin, cs.k, cs.stride, pc.hi, pc.lo, st);
}

template <int PREC>
int resnet_fwd_impl(const Prepared& P, const float* x, int B, int H, int W, float* out, const ResWs& w, cudaStream_t st) {
  constexpr bool HALF = PREC == tc::PREC_F16X2;
  ConvSpec specs[43];
  build_specs(specs);
  const int H1 = H / 2, W1 = W / 2, H2 = H / 4, W2 = W / 4;

  // One layer (a chain of bottleneck blocks, resnet_custom.py:35-55) over `bc` images: block 0 reads `in` and has
  // the downsample branch; intermediates rotate through the 4 scratch plane pairs; the last block writes `out`.
  struct Planes { bf16* hi; bf16* lo; };
  auto run_layer = [&](int l, int ci, int bc, int h, int wd, Planes in, const Planes* s, Planes out) -> int {
    const int blocks[3] = {3, 4, 6};
    Planes x = in;
    int xs = -1;  // scratch slot holding x (-1: external)
    bool borders_zeroed = false;
    for (int i = 0; i < blocks[l]; ++i) {
      const ConvSpec &c1 = specs[ci], &c2 = specs[ci + 1], &c3 = specs[ci + 2];
      const bool has_ds = i == 0, is_last = i == blocks[l] - 1;
      int free_slots[4], nf = 0;
      for (int q = 0; q < 4; ++q)
        if (q != xs) free_slots[nf++] = q;
      const Planes t1 = s[free_slots[0]], t2 = s[free_slots[1]];
      const int ys = free_slots[2];
      const Planes y = is_last ? out : s[ys];
      const int ho = h / c2.stride, wo = wd / c2.stride;
      // conv2 with halo reuse (conv3x3_halo.cuh) where it applies: conv1 then writes into the zero-bordered plane
      const int halo_kb = (HALF && resnet_halo() && c2.k == 3 && c2.stride == 1 && c1.k == 1 && c1.stride == 1)
                              ? (halo::halo_supported<1>(h, wd, c2.cin, c2.cout) ? 1 : (halo::halo_supported<2>(h, wd, c2.cin, c2.cout) ? 2 : 0))
                              : 0;
      if (halo_kb != 0) {
        if (!borders_zeroed) {  // (conv1 only ever writes the interior: once per layer call)
          TOAD_TRY(halo::launch_zero_borders(w.pad_hi, bc, h, wd, c2.cin, st));
          borders_zeroed = true;
        }
        TOAD_TRY(run_conv<PREC>(c1, P.conv[ci], x.hi, x.lo, bc, h, wd, w.pad_hi, nullptr, nullptr, nullptr, true, st, true));
        const PreparedConv& pc = P.conv[ci + 1];
        if (halo_kb == 1) TOAD_TRY(halo::launch_conv3x3_halo<1>(w.pad_hi, pc.hi, pc.lo, pc.bias, t2.hi, bc, h, wd, c2.cout, true, resnet_pdl(), st));
        else TOAD_TRY(halo::launch_conv3x3_halo<2>(w.pad_hi, pc.hi, pc.lo, pc.bias, t2.hi, bc, h, wd, c2.cout, true, resnet_pdl(), st));
      } else {
        TOAD_TRY(run_conv<PREC>(c1, P.conv[ci], x.hi, x.lo, bc, h, wd, t1.hi, t1.lo, nullptr, nullptr, true, st));
        TOAD_TRY(run_conv<PREC>(c2, P.conv[ci + 1], t1.hi, t1.lo, bc, h, wd, t2.hi, t2.lo, nullptr, nullptr, true, st));
      }
      Planes res = x;
      if (has_ds) {  // (xs == -1 here: all four scratch slots are free, the fourth holds the shortcut)
        res = s[free_slots[3]];
        TOAD_TRY(run_conv<PREC>(specs[ci + 3], P.conv[ci + 3], x.hi, x.lo, bc, h, wd, res.hi, res.lo, nullptr, nullptr, false, st));
      }
      TOAD_TRY(run_conv<PREC>(c3, P.conv[ci + 2], t2.hi, t2.lo, bc, ho, wo, y.hi, y.lo, res.hi, res.lo, true, st));
      x = y;
      xs = is_last ? -1 : ys;
      h = ho; wd = wo;
      ci += has_ds ? 4 : 3;
    }
    return 0;
  };
  const Planes bufs[5] = {{w.buf_hi[0], w.buf_lo[0]}, {w.buf_hi[1], w.buf_lo[1]}, {w.buf_hi[2], w.buf_lo[2]},
                          {w.buf_hi[3], w.buf_lo[3]}, {w.buf_hi[4], w.buf_lo[4]}};

  // ---- stem (conv1 7x7/s2 as im2col + GEMM, BN, ReLU; resnet_custom.py:97-99), maxpool 3x3/s2 (:100) and layer1,
  // in chunks of stem_chunk images (bounds the im2col scratch); layer1's output of every chunk lands in its slice
  // of bufs[4].
  const int64_t l1_img = static_cast<int64_t>(H2) * W2 * 256;  // elements per image of a layer1-sized plane
  for (int b0 = 0; b0 < B; b0 += w.stem_chunk) {
    const int nb = (B - b0) < w.stem_chunk ? (B - b0) : w.stem_chunk;
    const float* xb = x + static_cast<int64_t>(b0) * 3 * H * W;
    bool pooled = false;
    if (HALF && !stem_im2col_forced()) {
      // conv1 + bn1 + relu (+ the 3x3/s2 max-pool where the output rows fit one tile) as one implicit-GEMM kernel (stem.cuh)
      pooled = stem::stem_can_pool(H, W) && !stem_pool_unfused();
      TOAD_TRY(stem::launch_stem_fused(xb, P.conv[0].hi, P.conv[0].lo, P.conv[0].bias, pooled ? w.pool_hi : w.stem_hi, nb, H, W,
                                       pooled, st));
    } else {
      const int64_t rows = static_cast<int64_t>(nb) * H1 * W1;
      TOAD_TRY(resnet::launch_stem_im2col<HALF>(xb, w.col_hi, w.col_lo, nb, H, W, H1, W1, st));
      tc::GemmTcParams g{};
      g.M = rows; g.N = 64; g.K = resnet::STEM_KPAD; g.bias = P.conv[0].bias; g.relu = 1;
      g.out_hi = w.stem_hi; g.out_lo = w.stem_lo; g.ld_split = 64;
      TOAD_TRY((tc::launch_gemm<64, tc::A_SPLIT, tc::EPI_LINEAR, 2, kResOutBufs, PREC>(g, w.col_hi, w.col_lo, P.conv[0].hi, P.conv[0].lo, st)));
    }
    if (!pooled) {
      const int64_t threads = static_cast<int64_t>(nb) * H2 * W2 * (64 / 8);
      resnet::maxpool3x3s2_kernel<HALF><<<static_cast<unsigned>((threads + 255) / 256), 256, 0, st>>>(
          w.stem_hi, w.stem_lo, w.pool_hi, w.pool_lo, nb, H1, W1, 64);
      TOAD_CUDA_TRY(cudaGetLastError());
    }
    const Planes l1_out = {bufs[4].hi + b0 * l1_img, HALF ? nullptr : bufs[4].lo + b0 * l1_img};
    TOAD_TRY(run_layer(0, 1, nb, H2, W2, Planes{w.pool_hi, w.pool_lo}, bufs, l1_out));
  }
  // ---- layer2, layer3 over the whole batch
  // (a layer may write its output over its input: only block 0 reads `in`, only the last block writes `out`)
  TOAD_TRY(run_layer(1, 1 + 10, B, H2, W2, bufs[4], bufs, bufs[4]));
  TOAD_TRY(run_layer(2, 1 + 10 + 13, B, H2 / 2, W2 / 2, bufs[4], bufs, bufs[4]));
  const int cur = 4, h = H2 / 4, wd = W2 / 4;
  // ---- global average pool + flatten (resnet_custom.py:106-107)
  resnet::avgpool_kernel<HALF><<<dim3(B, 1024 / 256), 256, 0, st>>>(w.buf_hi[cur], w.buf_lo[cur], out, h * wd, 1024);
  TOAD_CUDA_TRY(cudaGetLastError());
  return 0;
}

}  // namespace

extern "C" int toad_resnet_prepared_bytes(size_t* bytes) {
  if (bytes == nullptr) return TOAD_ERR_ARG;
  *bytes = carve_prepared(nullptr).bytes;
  return 0;
}

extern "C" int toad_resnet_prepare(const float* const* tensors, int32_t n_tensors, void* prepared, size_t prepared_bytes,
                                   uint32_t flags, toad_stream_t stream) {
  if (tensors == nullptr || n_tensors != TOAD_RESNET_N_TENSORS) return TOAD_ERR_ARG;
  for (int i = 0; i < n_tensors; ++i)
    if (tensors[i] == nullptr) return TOAD_ERR_ARG;
  Prepared P = carve_prepared(prepared);
  TOAD_TRY(check_ws(prepared, prepared_bytes, P.bytes));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  ConvSpec specs[43];
  build_specs(specs);
  const bool exact = (flags & TOAD_RESNET_FLAG_EXACT) != 0;
  for (int i = 0; i < 43; ++i) {
    const float* const* t = tensors + 5 * i;
    if (exact)
      resnet::fold_conv_kernel<false><<<specs[i].cout, 256, 0, st>>>(t[0], t[1], t[2], t[3], t[4], 1e-5f, P.conv[i].hi, P.conv[i].lo,
                                                                   P.conv[i].bias, specs[i].cout, specs[i].cin, specs[i].k,
                                                                   specs[i].k, kpad_of(specs[i]), 0);
    else
      // (conv 0 of the fp16 mode feeds the fused stem kernel: its own K order, see stem.cuh)
      resnet::fold_conv_kernel<true><<<specs[i].cout, 256, 0, st>>>(t[0], t[1], t[2], t[3], t[4], 1e-5f, P.conv[i].hi, P.conv[i].lo,
                                                                  P.conv[i].bias, specs[i].cout, specs[i].cin, specs[i].k,
                                                                  specs[i].k, kpad_of(specs[i]), (i == 0 && !stem_im2col_forced()) ? 1 : 0);
    TOAD_CUDA_TRY(cudaGetLastError());
  }
  return 0;
}

extern "C" int toad_resnet_workspace_bytes(int32_t B, int32_t H, int32_t W, uint32_t flags, size_t* bytes) {
  TOAD_TRY(check_resnet_shape(B, H, W));
  if (bytes == nullptr) return TOAD_ERR_ARG;
  *bytes = carve_resnet(B, H, W, (flags & TOAD_RESNET_FLAG_EXACT) != 0, nullptr).bytes;
  return 0;
}

extern "C" int toad_resnet_fwd(const void* prepared, const float* x, int32_t B, int32_t H, int32_t W, float* out,
                               void* workspace, size_t workspace_bytes, uint32_t flags, toad_stream_t stream) {
  TOAD_TRY(check_resnet_shape(B, H, W));
  if (prepared == nullptr || x == nullptr || out == nullptr) return TOAD_ERR_ARG;
  if ((reinterpret_cast<uintptr_t>(prepared) & 255) != 0) return TOAD_ERR_WORKSPACE;
  const bool exact = (flags & TOAD_RESNET_FLAG_EXACT) != 0;
  Prepared P = carve_prepared(const_cast<void*>(prepared));
  ResWs w = carve_resnet(B, H, W, exact, workspace);
  TOAD_TRY(check_ws(workspace, workspace_bytes, w.bytes));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  return exact ? resnet_fwd_impl<tc::PREC_BF16X3>(P, x, B, H, W, out, w, st)
               : resnet_fwd_impl<tc::PREC_F16X2>(P, x, B, H, W, out, w, st);
}
