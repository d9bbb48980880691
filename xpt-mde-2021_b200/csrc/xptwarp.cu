// xptwarp.cu -- C-ABI of libxptwarp.so (see include/xptwarp.h).
// Host-side orchestration only: argument validation, scratch ownership, kernel
// launches on the caller's stream.  No CPU fallback: every entry point either
// launches the sm_100a kernels or returns an error.
#include "../../include/xptwarp.h"

#include <cuda_runtime.h>
#include <dlfcn.h>

#include <cstdarg>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <utility>
#include <vector>

#include "xpt_kernels.cuh"
#include "xpt_fused.cuh"
#include "xpt_strip.cuh"
#include "xpt_flow.cuh"
#include "xpt_minloss.cuh"
#include "xpt_minstrip.cuh"

using namespace xpt;

namespace {

thread_local char g_err[512] = "";

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

#define XPT_CUDA(call)                                                                         \
  do {                                                                                         \
    cudaError_t e_ = (call);                                                                   \
    if (e_ != cudaSuccess)                                                                     \
      return fail(XPT_CUDA_ERROR, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_),      \
                  __FILE__, __LINE__);                                                         \
  } while (0)

#define XPT_LAUNCH_CHECK(name)                                                                 \
  do {                                                                                         \
    cudaError_t e_ = cudaGetLastError();                                                       \
    if (e_ != cudaSuccess)                                                                     \
      return fail(XPT_CUDA_ERROR, "launch of %s failed: %s", name, cudaGetErrorString(e_));    \
    ++ctx->launches;                                                                           \
  } while (0)

inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

}  // namespace

constexpr int kMaxChunks = 8;   // pipeline depth of the host-buffer entry point

struct xpt_ctx {
  xpt_config cfg;
  int S, B, N, H, W;
  int h[kMaxScales], w[kMaxScales], s[kMaxScales];
  int tiles_x[kMaxScales], tiles_y[kMaxScales], first_tile[kMaxScales + 1];
  int ftiles_x[kMaxScales], ftiles_y[kMaxScales], ffirst_tile[kMaxScales + 1];   // 64x13 tiles (k_fused)
  int chunks[kMaxScales], first_chunk[kMaxScales + 1];       // 1024-pixel chunks (k_warp_bwd)
  int sm_chunks[kMaxScales], first_sm_chunk[kMaxScales + 1]; // 256-pixel chunks (k_smooth)
  int slots_per_b;
  size_t scratch_bytes;
  int launches;
  // device scratch
  float* geoK;
  float* geoT;
  float* src_pyr[kMaxScales];   // s > 1
  float* src4_pyr[kMaxScales];  // RGBx (16-byte) texels of every source level for k_fused's gathers; lazily allocated
  float* tgt_pyr[kMaxScales];   // s > 1
  float* loss_part;
  float* pose_part;
  double* loss_sum_b;           // [B][3] (k_epilogue)
  unsigned int* ticket;         // k_epilogue's last-block ticket (self-resetting)
  // lazily allocated
  float* synth_scr[kMaxScales];
  float* gsynth_scr[kMaxScales];
  float* dsrc_lvl[kMaxScales];  // s > 1
  float* dsrc4_lvl[kMaxScales]; // RGBx gradient levels of the fused path (all levels)
  float* tgt0_copy;             // unused unless a level-0 copy is wanted without a user buffer
  float* min_part;              // [B][S * full-res tiles] partial sums of xpt_photometric_min_loss
  float* l2_part;               // block partials (doubles) of xpt_l2_regularizer
  // streaming strip kernel (k_strip): static equal-cost partition of the (snippet, level, strip) rows over the grid
  StripPiece* strip_pieces; StripCta* strip_ctas;
  int strip_nctas, strip_slots, strip_ns; bool strip_ready;
  float* strip_loss_part; float* strip_pose_part;
  void* nccl_comm; bool nccl_owned;   // communicator of XPT_FLAG_ALLREDUCE / xpt_allreduce
  // peer-memory loss exchange (k_epilogue / loss_exchange): inbox of this rank, the table of every rank's inbox as
  // mapped into this process (CUDA IPC), the step counter and the error flag; p2p = all ranks agreed to use it
  bool p2p; int nranks, rank;
  float* inbox; float** peer_host; float** peer_table; unsigned int* exch_seq; unsigned int* exch_err;
  bool exchanged;                     // the last fused epilogue already summed the losses over the ranks
  bool host_call;                     // inside xpt_total_loss_host: the chunks' losses are rank-local (no collective per chunk)
  float** alloc_slot[160]; size_t alloc_floats[160]; int n_allocs;   // sizes of the lazily allocated scratch slots
  int geo_k_off, geo_t_off, geo_cap;   // this ctx's runs of K / [R|t] records in the constant bank (float offsets), snippets per launch
  bool geo_shared;              // no private run was free: the start of both regions, shared with other such contexts
  // staging for the host-buffer entry point
  float* st_frames; float* st_K; float* st_pose; float* st_losses; float* st_loss_batch; float* st_dpose;
  float* st_dsource;
  float* st_depth[kMaxScales]; float* st_disp[kMaxScales];
  float* st_ddepth[kMaxScales]; float* st_ddisp[kMaxScales];
  float* st_synth[kMaxScales]; float* st_mask[kMaxScales]; float* st_target[kMaxScales];
  // CUDA-graph cache of whole xpt_total_loss calls (XPT_FLAG_GRAPH), keyed by every argument
  struct GraphEntry { std::vector<uint64_t> key; cudaGraphExec_t exec; int launches; };
  std::vector<GraphEntry>* graphs;
  bool warm;                    // one eager call has run (all lazy scratch exists)
  // pipelined host entry point: a child ctx for one batch chunk, copy streams, events
  xpt_ctx* child;
  cudaStream_t s_in, s_in2, s_out;
  cudaEvent_t ev_small, ev_fork, ev_join;
  std::vector<GraphEntry>* host_graphs;   // whole xpt_total_loss_host calls (copies + compute), keyed by every pointer
  bool host_warm;
  cudaEvent_t ev_in[kMaxChunks], ev_done[kMaxChunks];
  float* h_losses;              // pinned [4][4]: a pageable destination would block the host per chunk
  bool host_pending = false, host_pending_async = false;   // xpt_total_loss_host_begin without its _end yet
  int host_pending_nc = 0;
  float* host_pending_losses = nullptr;
  cudaEvent_t host_done_ev = nullptr;
  // per-launch device timing of the dominant kernel (xpt_profile_*)
  std::vector<cudaEvent_t>* prof_events;
  int prof_count;
  int prof_on;                  // records still allowed
  int prof_kind;                // which kernel xpt_profile_* brackets: XPT_PROFILE_FUSED / XPT_PROFILE_PYRAMID
};

namespace {

// Lazily allocated scratch: a slot is allocated once; its size is remembered, and a later request for MORE than that
// (a call site that reuses the slot for another purpose) grows it instead of silently overrunning it.
int dev_alloc(xpt_ctx* ctx, float** p, size_t nfloats) {
  size_t* have = nullptr;
  for (int i = 0; i < ctx->n_allocs; ++i)
    if (ctx->alloc_slot[i] == p) { have = &ctx->alloc_floats[i]; break; }
  if (*p && have && *have >= nfloats) return XPT_OK;
  if (*p && !have) return XPT_OK;                  // allocated before tracking (xpt_create): sizes are fixed there
  if (*p) {                                        // grow: the old contents are scratch, nothing to preserve
    cudaFree(*p);
    ctx->scratch_bytes -= *have * sizeof(float);
    *p = nullptr;
  }
  void* q = nullptr;
  cudaError_t e = cudaMalloc(&q, nfloats * sizeof(float));
  if (e != cudaSuccess) {
    (void)cudaGetLastError();
    return fail(XPT_OUT_OF_MEMORY, "cudaMalloc of %zu bytes failed: %s", nfloats * sizeof(float),
                cudaGetErrorString(e));
  }
  *p = static_cast<float*>(q);
  ctx->scratch_bytes += nfloats * sizeof(float);
  if (have) *have = nfloats;
  else if (ctx->n_allocs < 160) { ctx->alloc_slot[ctx->n_allocs] = p; ctx->alloc_floats[ctx->n_allocs] = nfloats; ++ctx->n_allocs; }
  return XPT_OK;
}

#define XPT_TRY(expr)            \
  do {                           \
    int rc_ = (expr);            \
    if (rc_ != XPT_OK) return rc_; \
  } while (0)

size_t lvl_pix(const xpt_ctx* c, int l) { return (size_t)c->h[l] * c->w[l]; }

int check_frames(const xpt_ctx* ctx, const xpt_frames* f, bool need_target) {
  if (!f) return fail(XPT_BAD_ARGUMENT, "frames is NULL");
  if (!f->source) return fail(XPT_BAD_ARGUMENT, "frames.source is NULL");
  if (!f->intrinsic) return fail(XPT_BAD_ARGUMENT, "frames.intrinsic is NULL");
  if (need_target && !f->target) return fail(XPT_BAD_ARGUMENT, "frames.target is NULL");
  long long hw3 = (long long)ctx->H * ctx->W * 3;
  if (f->source_frame_stride < hw3) return fail(XPT_BAD_SHAPE, "source_frame_stride %lld < H*W*3", (long long)f->source_frame_stride);
  if (ctx->B > 1 && f->source_batch_stride < hw3) return fail(XPT_BAD_SHAPE, "source_batch_stride too small");
  if (need_target && ctx->B > 1 && f->target_batch_stride < hw3) return fail(XPT_BAD_SHAPE, "target_batch_stride too small");
  return XPT_OK;
}

int check_list(const xpt_ctx* ctx, const void* const* p, const char* name, bool all_required) {
  if (!p) return fail(XPT_BAD_ARGUMENT, "%s is NULL", name);
  if (all_required)
    for (int l = 0; l < ctx->S; ++l)
      if (!p[l]) return fail(XPT_BAD_ARGUMENT, "%s[%d] is NULL", name, l);
  return XPT_OK;
}

// level table with the source / target pointers for this call
LevelTable make_levels(const xpt_ctx* ctx, const xpt_frames* f, const float* const tgt_override[]) {
  LevelTable lt;
  memset(&lt, 0, sizeof(lt));
  lt.S = ctx->S;
  for (int l = 0; l < ctx->S; ++l) {
    Level& L = lt.lv[l];
    L.s = ctx->s[l]; L.H = ctx->h[l]; L.W = ctx->w[l];
    L.tiles_x = ctx->tiles_x[l]; L.tiles_y = ctx->tiles_y[l];
    L.slot_base = ctx->first_tile[l];
    if (f) {
      if (ctx->s[l] == 1) {
        L.src = f->source; L.src_bs = f->source_batch_stride; L.src_fs = f->source_frame_stride;
        L.tgt = f->target; L.tgt_bs = f->target_batch_stride;
      } else {
        L.src = ctx->src_pyr[l]; L.src_fs = (long long)lvl_pix(ctx, l) * 3; L.src_bs = L.src_fs * ctx->N;
        L.tgt = ctx->tgt_pyr[l]; L.tgt_bs = (long long)lvl_pix(ctx, l) * 3;
      }
    }
    if (tgt_override && tgt_override[l]) { L.tgt = tgt_override[l]; L.tgt_bs = (long long)lvl_pix(ctx, l) * 3; }
  }
  return lt;
}

GeoArgs make_geo(xpt_ctx* ctx, const float* pose, const float* intrinsic, float* matr_out) {
  GeoArgs g;
  memset(&g, 0, sizeof(g));
  g.pose = pose; g.intrinsic = intrinsic;
  g.geoK = intrinsic ? ctx->geoK : nullptr; g.geoT = pose ? ctx->geoT : nullptr; g.matr_out = matr_out;
  g.B = ctx->B; g.N = ctx->N; g.S = ctx->S;
  for (int l = 0; l < ctx->S; ++l) g.s[l] = ctx->s[l];
  return g;
}

int launch_geometry(xpt_ctx* ctx, const float* pose, const float* intrinsic, float* matr_out, cudaStream_t st) {
  int n = ctx->B * (ctx->N > ctx->S ? ctx->N : ctx->S);
  k_geometry<<<cdiv(n, 128), 128, 0, st>>>(make_geo(ctx, pose, intrinsic, matr_out));
  XPT_LAUNCH_CHECK("k_geometry");
  return XPT_OK;
}

// cuTensorMapEncodeTiled through the runtime (no link against libcuda): NULL when the driver has none
typedef CUresult (*TmapEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                 const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
TmapEncodeFn tmap_encoder() {
  static TmapEncodeFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr) == cudaSuccess && qr == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<TmapEncodeFn>(p);
    else
      (void)cudaGetLastError();
  });
  return getenv("XPT_NO_TMA") ? nullptr : fn;
}

// fp32 frames [outer...][H][W*3] as a TMA tensor: dim 0 = W*3 floats (dense), dim 1 = rows, then the frame / batch dims
bool make_frame_tmap(CUtensorMap* tm, const float* base, int W, int H, int rank, const long long outer_dims[],
                     const long long outer_strides_elems[]) {
  TmapEncodeFn enc = tmap_encoder();
  if (!enc) return false;
  cuuint64_t dims[5] = {(cuuint64_t)W * 3, (cuuint64_t)H, 1, 1, 1};
  cuuint64_t strides[4] = {(cuuint64_t)W * 3 * sizeof(float), 0, 0, 0};      // strides of dims 1.. in bytes
  cuuint32_t box[5] = {(cuuint32_t)kPyrTmaTW * 3, (cuuint32_t)kPyrTH, 1, 1, 1};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  for (int i = 0; i < rank - 2; ++i) {
    dims[2 + i] = (cuuint64_t)outer_dims[i];
    strides[1 + i] = (cuuint64_t)outer_strides_elems[i] * sizeof(float);
    if (outer_dims[i] > 1 && (strides[1 + i] % 16 || strides[1 + i] == 0)) return false;
  }
  // a dimension of extent 1 still needs a legal (multiple of 16, non-zero) stride
  for (int i = 0; i < rank - 2; ++i)
    if (strides[1 + i] == 0 || strides[1 + i] % 16) strides[1 + i] = (cuuint64_t)W * 3 * sizeof(float) * H;
  return enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, const_cast<float*>(base), dims, strides, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// source pyramid into the ctx (+ target pyramid into ctx or user buffers)
// rgba: write the source levels (incl. full resolution) as RGBx texels for the fused kernel INSTEAD of the 3-channel levels
int launch_pyramids(xpt_ctx* ctx, const xpt_frames* f, float* const target_ms[], bool want_source,
                    cudaStream_t st, const float* geo_pose = nullptr, bool rgba = false) {
  if (rgba)
    for (int l = 0; l < ctx->S; ++l)
      XPT_TRY(dev_alloc(ctx, &ctx->src4_pyr[l], (size_t)ctx->B * ctx->N * lvl_pix(ctx, l) * 4));
  // ---- fast path: scales within {1,2,4,8}, 16-byte aligned rows -> one tiled pass over the frames
  {
    bool ok = ctx->W % 8 == 0 && ctx->H % 8 == 0 && want_source;
    for (int l = 0; l < ctx->S; ++l) ok = ok && (ctx->s[l] == 1 || ctx->s[l] == 2 || ctx->s[l] == 4 || ctx->s[l] == 8);
    auto al16 = [](const void* p) { return ((uintptr_t)p & 15u) == 0; };
    ok = ok && al16(f->source) && f->source_batch_stride % 4 == 0 && f->source_frame_stride % 4 == 0;
    if (f->target) ok = ok && al16(f->target) && f->target_batch_stride % 4 == 0;
    if (ok) {
      PyramidTiledArgs t;
      memset(&t, 0, sizeof(t));
      t.source = f->source; t.src_bs = f->source_batch_stride; t.src_fs = f->source_frame_stride;
      t.target = f->target; t.tgt_bs = f->target_batch_stride;
      t.B = ctx->B; t.N = ctx->N; t.H = ctx->H; t.W = ctx->W;
      bool any = false;
      for (int l = 0; l < ctx->S; ++l) {
        const int sc = ctx->s[l];
        const int lg = sc == 1 ? 0 : (sc == 2 ? 1 : (sc == 4 ? 2 : 3));
        if (rgba) { t.src4_out[lg] = reinterpret_cast<float4*>(ctx->src4_pyr[l]); any = true; }
        if (sc == 1) continue;
        if (!rgba) t.src_out[lg] = ctx->src_pyr[l];
        if (f->target) t.tgt_out[lg] = ctx->tgt_pyr[l];
        any = true;
      }
      if (geo_pose) { t.with_geometry = 1; t.geo = make_geo(ctx, geo_pose, f->intrinsic, nullptr); }
      if (any || geo_pose) {
        int gx = cdiv(ctx->W, kPyrTW);
        const int need = cdiv((long long)ctx->B * (ctx->N > ctx->S ? ctx->N : ctx->S), kPyrThreads);
        if (geo_pose && need > gx) gx = need;
        dim3 grid(gx, ctx->H / kPyrTH, ctx->B * (ctx->N + 1) + 1);
        if (!any) grid = dim3(need, 1, ctx->B * (ctx->N + 1) + 1);
        const bool prof = ctx->prof_on > 0 && ctx->prof_count < ctx->prof_on && ctx->prof_kind == XPT_PROFILE_PYRAMID;
        if (prof) XPT_CUDA(cudaEventRecord((*ctx->prof_events)[2 * ctx->prof_count], st));
        // tile loads by the TMA unit (k_pyramid_tma) when tensor maps over the caller's frames can be encoded
        CUtensorMap tm_src, tm_tgt;
        // XPT_PYRAMID=tiled: the LDG/STS tile kernel instead (A/B switch for profiling, read once)
        static const bool force_tiled = [] { const char* e = getenv("XPT_PYRAMID"); return e && !strcmp(e, "tiled"); }();
        bool tma = any && !force_tiled;
        if (tma) {
          const long long sd[2] = {ctx->N, ctx->B}, ss[2] = {(long long)f->source_frame_stride, (long long)f->source_batch_stride};
          tma = make_frame_tmap(&tm_src, f->source, ctx->W, ctx->H, 4, sd, ss);
          if (tma && f->target) {
            const long long td[1] = {ctx->B}, ts[1] = {(long long)f->target_batch_stride};
            tma = make_frame_tmap(&tm_tgt, f->target, ctx->W, ctx->H, 3, td, ts);
          } else if (tma) {
            tm_tgt = tm_src;
          }
        }
        // XPT_PYRAMID=tma_persistent: the persistent, double-buffered form (k_pyramid_tma_p) instead of one CTA per tile
        // pair.  Measured slower (18.9 vs 17.9 us at config 2, 94.6 vs 86.0 us at config 3: profiles/r02_pyramid_ab.txt),
        // so it is opt-in.
        static const bool persistent = [] { const char* e = getenv("XPT_PYRAMID"); return e && !strcmp(e, "tma_persistent"); }();
        if (tma && persistent) {
          const int ntx = cdiv(ctx->W, kPyrTmaTW * kPyrTmaBoxes), nty = ctx->H / kPyrTH;
          const long long total = (long long)ntx * nty * ctx->B * (ctx->N + 1);
          int dev_sms = 148;
          cudaDeviceGetAttribute(&dev_sms, cudaDevAttrMultiProcessorCount, ctx->cfg.device);
          long long gp = (long long)dev_sms * XPT_PYR_MINB;
          if (gp > total) gp = total;
          if (geo_pose && need > gp) gp = need;
          k_pyramid_tma_p<<<(unsigned)gp, kPyrThreads, 0, st>>>(t, tm_src, tm_tgt, ntx, nty, (int)total);
          XPT_LAUNCH_CHECK("k_pyramid_tma_p");
        } else if (tma) {
          int gxt = cdiv(ctx->W, kPyrTmaTW * kPyrTmaBoxes);
          if (geo_pose && need > gxt) gxt = need;
          dim3 gridt(gxt, ctx->H / kPyrTH, ctx->B * (ctx->N + 1) + 1);
          k_pyramid_tma<<<gridt, kPyrThreads, 0, st>>>(t, tm_src, tm_tgt);
          XPT_LAUNCH_CHECK("k_pyramid_tma");
        } else {
          k_pyramid_tiled<<<grid, kPyrThreads, 0, st>>>(t);
          XPT_LAUNCH_CHECK("k_pyramid_tiled");
        }
        if (prof) { XPT_CUDA(cudaEventRecord((*ctx->prof_events)[2 * ctx->prof_count + 1], st)); ++ctx->prof_count; }
      }
      if (target_ms && f->target)
        for (int l = 0; l < ctx->S; ++l) {
          if (!target_ms[l]) continue;
          const size_t row = lvl_pix(ctx, l) * 3 * sizeof(float);
          if (ctx->s[l] > 1)
            XPT_CUDA(cudaMemcpyAsync(target_ms[l], ctx->tgt_pyr[l], (size_t)ctx->B * row, cudaMemcpyDeviceToDevice, st));
          else
            XPT_CUDA(cudaMemcpy2DAsync(target_ms[l], row, f->target, f->target_batch_stride * sizeof(float), row, ctx->B,
                                       cudaMemcpyDeviceToDevice, st));
        }
      return XPT_OK;
    }
  }
  PyramidArgs a;
  memset(&a, 0, sizeof(a));
  if (geo_pose) { a.with_geometry = 1; a.geo = make_geo(ctx, geo_pose, f->intrinsic, nullptr); }
  a.source = f->source; a.src_bs = f->source_batch_stride; a.src_fs = f->source_frame_stride;
  a.target = f->target; a.tgt_bs = f->target_batch_stride;
  a.B = ctx->B; a.N = ctx->N; a.H = ctx->H; a.W = ctx->W; a.S = ctx->S;
  long long maxcount = 0;
  for (int l = 0; l < ctx->S; ++l) {
    a.s[l] = ctx->s[l];
    bool work = false;
    if (rgba && want_source) { a.src4_out[l] = reinterpret_cast<float4*>(ctx->src4_pyr[l]); work = true; }
    if (ctx->s[l] > 1) {
      if (want_source && !rgba) { a.src_out[l] = ctx->src_pyr[l]; work = true; }
      if (f->target) { a.tgt_out[l] = ctx->tgt_pyr[l]; work = true; }
    } else if (f->target && target_ms && target_ms[l]) {
      a.tgt_out[l] = target_ms[l];      // level-0 copy for the caller
      work = true;
    }
    if (work) {
      long long c = (long long)ctx->B * (ctx->N + 1) * lvl_pix(ctx, l);
      if (c > maxcount) maxcount = c;
    }
  }
  if (a.with_geometry) {
    long long g = (long long)ctx->B * (ctx->N > ctx->S ? ctx->N : ctx->S);
    if (g > maxcount) maxcount = g;
  }
  if (maxcount == 0) return XPT_OK;
  dim3 grid(cdiv(maxcount, 256), ctx->S + (a.with_geometry ? 1 : 0));
  k_pyramid<<<grid, 256, 0, st>>>(a);
  XPT_LAUNCH_CHECK("k_pyramid");
  // user copies of the target pyramid (augm_data["target_ms"])
  if (target_ms && f->target)
    for (int l = 0; l < ctx->S; ++l)
      if (ctx->s[l] > 1 && target_ms[l])
        XPT_CUDA(cudaMemcpyAsync(target_ms[l], ctx->tgt_pyr[l], (size_t)ctx->B * lvl_pix(ctx, l) * 3 * sizeof(float),
                                 cudaMemcpyDeviceToDevice, st));
  return XPT_OK;
}

int launch_warp_fwd(xpt_ctx* ctx, const LevelTable& lt, const float* const depth_ms[], float* const synth_ms[],
                    float* const mask_ms[], cudaStream_t st) {
  WarpFwdArgs a;
  memset(&a, 0, sizeof(a));
  a.lt = lt; a.B = ctx->B; a.N = ctx->N; a.geoK = ctx->geoK; a.geoT = ctx->geoT;
  for (int l = 0; l < ctx->S; ++l) {
    a.depth[l] = depth_ms[l]; a.synth[l] = synth_ms[l]; a.mask[l] = mask_ms ? mask_ms[l] : nullptr;
  }
  dim3 grid(cdiv((long long)ctx->B * lvl_pix(ctx, 0), 256), ctx->S);
  for (int l = 1; l < ctx->S; ++l) {
    int g = cdiv((long long)ctx->B * lvl_pix(ctx, l), 256);
    if ((unsigned)g > grid.x) grid.x = g;
  }
  k_warp_fwd<<<grid, 256, 0, st>>>(a);
  XPT_LAUNCH_CHECK("k_warp_fwd");
  return XPT_OK;
}

// prepares per-level dL/dsource buffers; level with s == 1 accumulates straight into d_source
int prepare_dsource(xpt_ctx* ctx, float* d_source, float* d_src[], long long bs[], long long fs[], cudaStream_t st) {
  for (int l = 0; l < ctx->S; ++l) { d_src[l] = nullptr; bs[l] = 0; fs[l] = 0; }
  if (!d_source) return XPT_OK;
  size_t full = (size_t)ctx->B * ctx->N * ctx->H * ctx->W * 3;
  XPT_CUDA(cudaMemsetAsync(d_source, 0, full * sizeof(float), st));
  for (int l = 0; l < ctx->S; ++l) {
    size_t n = (size_t)ctx->B * ctx->N * lvl_pix(ctx, l) * 3;
    if (ctx->s[l] == 1) {
      d_src[l] = d_source;
    } else {
      XPT_TRY(dev_alloc(ctx, &ctx->dsrc_lvl[l], n));
      XPT_CUDA(cudaMemsetAsync(ctx->dsrc_lvl[l], 0, n * sizeof(float), st));
      d_src[l] = ctx->dsrc_lvl[l];
    }
    fs[l] = (long long)lvl_pix(ctx, l) * 3;
    bs[l] = fs[l] * ctx->N;
  }
  return XPT_OK;
}

int finish_dsource(xpt_ctx* ctx, float* d_source, cudaStream_t st) {
  if (!d_source) return XPT_OK;
  PyramidAdjArgs a;
  memset(&a, 0, sizeof(a));
  a.d_source = d_source; a.BN = ctx->B * ctx->N; a.H = ctx->H; a.W = ctx->W; a.S = ctx->S;
  bool any = false;
  for (int l = 0; l < ctx->S; ++l) {
    a.s[l] = ctx->s[l];
    if (ctx->s[l] > 1) { a.d_level[l] = ctx->dsrc_lvl[l]; any = true; }
  }
  if (!any) return XPT_OK;
  long long total = (long long)a.BN * a.H * a.W;
  k_pyramid_adjoint<<<cdiv(total, 256), 256, 0, st>>>(a);
  XPT_LAUNCH_CHECK("k_pyramid_adjoint");
  return XPT_OK;
}

// fused path: RGBx gradient levels (zeroed), folded into the dense d_source by k_dsource_finish
int prepare_dsource4(xpt_ctx* ctx, float4* d_src4[], cudaStream_t st) {
  for (int l = 0; l < ctx->S; ++l) {
    const size_t n = (size_t)ctx->B * ctx->N * lvl_pix(ctx, l) * 4;
    XPT_TRY(dev_alloc(ctx, &ctx->dsrc4_lvl[l], n));
    XPT_CUDA(cudaMemsetAsync(ctx->dsrc4_lvl[l], 0, n * sizeof(float), st));
    d_src4[l] = reinterpret_cast<float4*>(ctx->dsrc4_lvl[l]);
  }
  return XPT_OK;
}

int finish_dsource4(xpt_ctx* ctx, float* d_source, cudaStream_t st) {
  DsourceFinishArgs a;
  memset(&a, 0, sizeof(a));
  a.d_source = d_source; a.BN = ctx->B * ctx->N; a.H = ctx->H; a.W = ctx->W; a.S = ctx->S;
  bool have0 = false;
  for (int l = 0; l < ctx->S; ++l) {
    a.s[l] = ctx->s[l];
    a.d_level4[l] = reinterpret_cast<const float4*>(ctx->dsrc4_lvl[l]);
    have0 = have0 || ctx->s[l] == 1;
  }
  (void)have0;       // without a level-0 scale the pass still writes every element (zeros where no level touches it)
  if (a.H > 65535 || a.BN > 65535) return fail(XPT_BAD_SHAPE, "d_source: H or B*N exceeds the grid limit 65535");
  k_dsource_finish<<<dim3(cdiv(a.W, 256), a.H, a.BN), 256, 0, st>>>(a);
  XPT_LAUNCH_CHECK("k_dsource_finish");
  return XPT_OK;
}

int launch_warp_bwd(xpt_ctx* ctx, const LevelTable& lt, const float* const depth_ms[], const float* const gsynth[],
                    float* const d_depth_ms[], float* d_source, const float* pose, float* d_pose, float scale,
                    cudaStream_t st) {
  WarpBwdArgs a;
  memset(&a, 0, sizeof(a));
  a.lt = lt; a.B = ctx->B; a.N = ctx->N; a.geoK = ctx->geoK; a.geoT = ctx->geoT;
  a.pose_part = ctx->pose_part; a.slots_per_b = ctx->slots_per_b;
  XPT_TRY(prepare_dsource(ctx, d_source, a.d_src, a.d_src_bs, a.d_src_fs, st));
  int maxchunks = 0;
  for (int l = 0; l < ctx->S; ++l) {
    a.depth[l] = depth_ms[l]; a.gsynth[l] = gsynth[l]; a.d_depth[l] = d_depth_ms ? d_depth_ms[l] : nullptr;
    a.chunk_base[l] = ctx->first_chunk[l];
    if (ctx->chunks[l] > maxchunks) maxchunks = ctx->chunks[l];
  }
  dim3 grid(maxchunks, ctx->B, ctx->S);
  k_warp_bwd<<<grid, kWarpBwdThreads, 0, st>>>(a);
  XPT_LAUNCH_CHECK("k_warp_bwd");
  if (d_pose) {
    k_pose_epilogue<<<ctx->B * ctx->N, 128, 0, st>>>(ctx->pose_part, ctx->slots_per_b, ctx->first_chunk[ctx->S],
                                                     pose, d_pose, ctx->N, scale);
    XPT_LAUNCH_CHECK("k_pose_epilogue");
  }
  XPT_TRY(finish_dsource(ctx, d_source, st));
  return XPT_OK;
}

void fill_photo_norms(const xpt_ctx* ctx, PhotoArgs& a) {
  a.B = ctx->B; a.N = ctx->N;
  a.tiles_per_b = ctx->first_tile[ctx->S];
  for (int l = 0; l <= ctx->S; ++l) a.first_tile[l] = ctx->first_tile[l];
  a.slots_per_b = ctx->slots_per_b;
  a.loss_part = ctx->loss_part;
  a.pose_part = ctx->pose_part;
  a.grad_factor = ctx->cfg.img_grad_factor;
  for (int l = 0; l < ctx->S; ++l) {
    double hw = (double)ctx->h[l] * ctx->w[l];
    double sw = ctx->cfg.scale_weights[l];
    a.norm_photo[l] = (float)(sw / ((double)ctx->N * hw * 3.0));
    // losses.py:401-402: each scale's smoothness divided by scale = orig_width / width
    double scale = (double)ctx->w[0] / (double)ctx->w[l];   // orig_width = width of target_ms[0] (the FIRST level, losses.py:399)
    a.norm_sm_x[l] = (float)(sw * 0.5 / ((double)ctx->h[l] * (ctx->w[l] - 1)) / scale);
    a.norm_sm_y[l] = (float)(sw * 0.5 / ((double)(ctx->h[l] - 1) * ctx->w[l]) / scale);
  }
}

// opt a kernel in to its dynamic shared-memory size once per DEVICE (the attribute is per context; `done` is the
// call site's bit mask over device ordinals, so several ranks / devices in one process each get it)
template <typename F>
int ensure_dyn_smem(F* func, size_t bytes, int device, unsigned long long* done) {
  const unsigned long long bit = 1ull << (device & 63);
  if (__atomic_load_n(done, __ATOMIC_ACQUIRE) & bit) return XPT_OK;
  XPT_CUDA(cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  __atomic_fetch_or(done, bit, __ATOMIC_RELEASE);
  return XPT_OK;
}

template <bool FUSED, bool GRAD>
int launch_photo(xpt_ctx* ctx, const PhotoArgs& a, cudaStream_t st) {
  static unsigned long long attr_done = 0;
  size_t smem = PhotoSmem<GRAD>::kBytes;
  XPT_TRY(ensure_dyn_smem(k_photo<FUSED, GRAD>, smem, ctx->cfg.device, &attr_done));
  dim3 grid(a.tiles_per_b, ctx->B);
  k_photo<FUSED, GRAD><<<grid, kPhotoThreads, smem, st>>>(a);
  XPT_LAUNCH_CHECK(FUSED ? "k_photo<fused>" : "k_photo<tensor>");
  return XPT_OK;
}

// Constant-bank slots: k_fused reads the camera geometry through uniform-register operands, i.e. from the 60 KB
// __constant__ block c_geo: a K region (records of S x 18 floats) and a [R|t] region (records of N x 12 floats).
// Every ctx owns a private run of records in both (first fit, per device), so launches of DIFFERENT contexts on
// different streams never touch each other's geometry.  A ctx that finds no free run shares the start of both
// regions (geo_shared) and must then be ordered against the other contexts of the device by its caller.
struct GeoSlotTable { std::mutex m; std::vector<std::pair<int, int>> used_k, used_t; };     // (offset, length) in floats, sorted
GeoSlotTable g_geo_slots[64];

// first fit of `want` floats inside [lo, hi) with the offset (relative to lo) a multiple of `align`; -1 = none
int geo_first_fit(std::vector<std::pair<int, int>>& used, int lo, int hi, int want, int align, size_t* at) {
  int pos = lo;
  for (size_t i = 0; i <= used.size(); ++i) {
    const int end = i < used.size() ? used[i].first : hi;
    const int p = lo + ((pos - lo + align - 1) / align) * align;
    if (end - p >= want) { *at = i; return p; }
    if (i < used.size()) pos = used[i].first + used[i].second;
  }
  return -1;
}

void geo_slot_acquire(xpt_ctx* ctx) {
  const int krec = ctx->S * kGeoK, trec = ctx->N * kGeoT;
  int cap = kGeoKRegion / krec;
  if ((kGeoConstFloats - kGeoKRegion) / trec < cap) cap = (kGeoConstFloats - kGeoKRegion) / trec;
  const int nb = ctx->B < cap ? ctx->B : cap;             // snippets per launch; larger batches launch in chunks
  GeoSlotTable& t = g_geo_slots[ctx->cfg.device & 63];
  std::lock_guard<std::mutex> lk(t.m);
  size_t ik = 0, it = 0;
  const int pk = geo_first_fit(t.used_k, 0, kGeoKRegion, nb * krec, krec, &ik);
  const int pt = geo_first_fit(t.used_t, kGeoKRegion, kGeoConstFloats, nb * trec, 1, &it);
  ctx->geo_cap = nb;
  if (pk >= 0 && pt >= 0) {
    t.used_k.insert(t.used_k.begin() + ik, std::make_pair(pk, nb * krec));
    t.used_t.insert(t.used_t.begin() + it, std::make_pair(pt, nb * trec));
    ctx->geo_k_off = pk; ctx->geo_t_off = pt; ctx->geo_shared = false;
  } else {
    ctx->geo_k_off = 0; ctx->geo_t_off = kGeoKRegion; ctx->geo_shared = true;
  }
}

void geo_slot_release(xpt_ctx* ctx) {
  if (ctx->geo_shared || ctx->geo_cap == 0) return;
  GeoSlotTable& t = g_geo_slots[ctx->cfg.device & 63];
  std::lock_guard<std::mutex> lk(t.m);
  for (size_t i = 0; i < t.used_k.size(); ++i)
    if (t.used_k[i].first == ctx->geo_k_off) { t.used_k.erase(t.used_k.begin() + i); break; }
  for (size_t i = 0; i < t.used_t.size(); ++i)
    if (t.used_t[i].first == ctx->geo_t_off) { t.used_t.erase(t.used_t.begin() + i); break; }
  ctx->geo_cap = 0;
}

template <bool GRAD, bool OUT, bool DSRC, int DERIVE>
int launch_fused(xpt_ctx* ctx, FusedArgs& a, cudaStream_t st) {
  static unsigned long long attr_done = 0;
  const size_t smem = FusedSmem<GRAD>::kBytes;
  XPT_TRY(ensure_dyn_smem(k_fused<GRAD, OUT, DSRC, DERIVE>, smem, ctx->cfg.device, &attr_done));
  // the geometry of up to `cap` snippets fits this ctx's constant-bank records; larger batches are launched in chunks
  if (ctx->geo_cap == 0) geo_slot_acquire(ctx);        // taken at the first fused launch: other entry points need none
  const int cap = ctx->geo_cap;
  const int krec = ctx->S * kGeoK, trec = ctx->N * kGeoT;
  const int k_rec0 = ctx->geo_k_off / krec;            // blockIdx.x of the first snippet of a launch
  const bool prof = ctx->prof_on > 0 && ctx->prof_count < ctx->prof_on && ctx->prof_kind == XPT_PROFILE_FUSED;
  if (prof) XPT_CUDA(cudaEventRecord((*ctx->prof_events)[2 * ctx->prof_count], st));
  for (int b0 = 0; b0 < ctx->B; b0 += cap) {
    const int bc = ctx->B - b0 < cap ? ctx->B - b0 : cap;
    XPT_CUDA(cudaMemcpyToSymbolAsync(c_geo, ctx->geoK + (size_t)b0 * krec, (size_t)bc * krec * sizeof(float),
                                     (size_t)ctx->geo_k_off * sizeof(float), cudaMemcpyDeviceToDevice, st));
    XPT_CUDA(cudaMemcpyToSymbolAsync(c_geo, ctx->geoT + (size_t)b0 * trec, (size_t)bc * trec * sizeof(float),
                                     (size_t)ctx->geo_t_off * sizeof(float), cudaMemcpyDeviceToDevice, st));
    a.k_rec0 = k_rec0;
    a.b_off = b0 - k_rec0;                               // snippet of blockIdx.x = 0
    a.geo_t_off = ctx->geo_t_off - k_rec0 * trec;        // [R|t] of (blockIdx.x, n) at geo_t_off + (blockIdx.x * N + n) * 12
    // tile-major once the grid is more than ~4 waves of the device (XPT_GRID_TILE_MAJOR = 0 / 1 forces the A/B)
    a.tile_major = XPT_GRID_TILE_MAJOR == 2 ? ((long long)bc * a.tiles_per_b > 4 * 296) : XPT_GRID_TILE_MAJOR;
    dim3 grid = a.tile_major ? dim3(a.tiles_per_b, k_rec0 + bc) : dim3(k_rec0 + bc, a.tiles_per_b);
    k_fused<GRAD, OUT, DSRC, DERIVE><<<grid, kFThreads, smem, st>>>(a);
    XPT_LAUNCH_CHECK("k_fused");
  }
  if (prof) { XPT_CUDA(cudaEventRecord((*ctx->prof_events)[2 * ctx->prof_count + 1], st)); ++ctx->prof_count; }
  return XPT_OK;
}

// ---- streaming strip kernel ---------------------------------------------------------------------------------
// Static partition: the rows of every (level, snippet, strip) are laid end to end (large levels first) and cut into
// `nctas` runs of equal cost, cost = row chunks of 2 rows incl. the 4 halo rows a piece re-warps.  A cut inside a
// strip makes two pieces.  Partial-sum slots are numbered per snippet; unused slots stay zero for the ctx's life.
template <int NS, bool DERIVE>
int strip_grid(int device, int* ctas_per_sm) {
  static unsigned long long attr_done = 0;
  XPT_TRY(ensure_dyn_smem(k_strip<NS, DERIVE>, StripSmem<NS>::kBytes, device, &attr_done));
  int n = 0;
  XPT_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_strip<NS, DERIVE>, StripSmem<NS>::kThreads, StripSmem<NS>::kBytes));
  if (n < 1) return fail(XPT_CUDA_ERROR, "k_strip<%d> does not fit an SM", NS);
  *ctas_per_sm = n;
  return XPT_OK;
}

int prepare_strip(xpt_ctx* ctx) {
  if (ctx->strip_ready) return XPT_OK;
  const int NS = ctx->N == 1 ? 1 : (ctx->N == 2 ? 2 : 4);
  int per_sm = 1;
  if (NS == 1) XPT_TRY((strip_grid<1, false>(ctx->cfg.device, &per_sm)));
  else if (NS == 2) XPT_TRY((strip_grid<2, false>(ctx->cfg.device, &per_sm)));
  else XPT_TRY((strip_grid<4, false>(ctx->cfg.device, &per_sm)));
  cudaDeviceProp prop;
  XPT_CUDA(cudaGetDeviceProperties(&prop, ctx->cfg.device));
  int nctas = prop.multiProcessorCount * per_sm;
  if (const char* e = getenv("XPT_STRIP_CTAS")) { const int v = atoi(e); if (v > 0) nctas = v; }
  // The strips of ONE snippet, large levels first.  Every snippet is cut at the same rows (so a snippet's losses and
  // gradients do not depend on its position in the batch) into R runs of equal cost; run r of snippet b goes to
  // CTA (r * B + b) mod nctas.  R is chosen to minimise the largest per-CTA load.
  struct Unit { int l, x0, cw, rows; };
  std::vector<Unit> units;
  long long snip = 0;
  for (int l = 0; l < ctx->S; ++l) {
    const int W = ctx->w[l], ns = cdiv(W, kSCWMax);
    int cw = 2 * cdiv(W, 2 * ns);
    if (cw > kSCWMax) cw = kSCWMax;
    for (int k = 0; k < ns; ++k) {
      const int x0 = k * cw;
      if (x0 >= W) break;
      units.push_back({l, x0, (W - x0 < cw ? W - x0 : cw), ctx->h[l]});
      snip += (ctx->h[l] + 4 + 1) / 2;
    }
  }
  const int kMinRows = 8;
  struct Tpl { std::vector<StripPiece> pieces; std::vector<int> run_first; std::vector<int> run_cost; };
  auto build = [&](int R, Tpl& t) {
    t.pieces.clear(); t.run_first.assign(1, 0); t.run_cost.assign(1, 0);
    long long target = (snip + 2LL * R + R - 1) / R;          // every cut adds two chunks of halo
    if (target < 8) target = 8;
    for (const Unit& u : units) {
      int ya = 0;
      while (ya < u.rows) {
        const bool last = (int)t.run_cost.size() == R;
        const long long left = target - t.run_cost.back();
        if (!last && left < (kMinRows + 4) / 2) { t.run_first.push_back((int)t.pieces.size()); t.run_cost.push_back(0); continue; }
        int take = u.rows - ya;
        if (!last) {
          const long long max_rows = 2 * left - 4;
          if (take > max_rows) take = (int)max_rows;
          if (u.rows - ya - take > 0 && u.rows - ya - take < kMinRows) take = u.rows - ya;     // no sliver pieces
        }
        StripPiece p;
        p.b = 0; p.l = u.l; p.x0 = u.x0; p.cw = u.cw; p.ya = ya; p.yb = ya + take;
        p.slot = (int)t.pieces.size(); p.nch = (take + 4 + 1) / 2;
        t.pieces.push_back(p);
        t.run_cost.back() += p.nch;
        ya += take;
      }
    }
    t.run_first.push_back((int)t.pieces.size());
  };
  auto makespan = [&](const Tpl& t) {
    std::vector<long long> load(nctas, 0);
    const int R = (int)t.run_cost.size();
    long long mx = 0;
    for (int r = 0; r < R; ++r)
      for (int b = 0; b < ctx->B; ++b) {
        long long& v = load[((long long)r * ctx->B + b) % nctas];
        v += t.run_cost[r];
        if (v > mx) mx = v;
      }
    return mx;
  };
  Tpl best, cand;
  long long best_ms = -1;
  int Rmax = (int)(snip / 6) + 1;
  if (Rmax > 2 * nctas) Rmax = 2 * nctas;
  if (const char* e = getenv("XPT_STRIP_RUNS")) { const int v = atoi(e); if (v > 0) { build(v, best); best_ms = makespan(best); Rmax = 0; } }
  for (int R = 1; R <= Rmax; ++R) {
    build(R, cand);
    const long long ms = makespan(cand);
    if (best_ms < 0 || ms < best_ms) { best_ms = ms; best = cand; }
  }
  const int R = (int)best.run_cost.size();
  std::vector<StripPiece> pieces;
  std::vector<StripCta> ctas(nctas, StripCta{0, 0, 0, 0});
  for (int c = 0; c < nctas; ++c) {
    ctas[c].first = (int)pieces.size();
    for (long long i = c; i < (long long)R * ctx->B; i += nctas) {
      const int r = (int)(i / ctx->B), b = (int)(i % ctx->B);
      for (int k = best.run_first[r]; k < best.run_first[r + 1]; ++k) {
        StripPiece p = best.pieces[k];
        p.b = b;
        pieces.push_back(p);
        ctas[c].count += 1; ctas[c].chunks += p.nch;
      }
    }
  }
  const int slots = (int)best.pieces.size() > 0 ? (int)best.pieces.size() : 1;
  ctx->strip_nctas = nctas; ctx->strip_slots = slots; ctx->strip_ns = NS;
  {
    float* tmp = nullptr;
    XPT_TRY(dev_alloc(ctx, &tmp, (pieces.size() + 1) * sizeof(StripPiece) / sizeof(float)));
    ctx->strip_pieces = reinterpret_cast<StripPiece*>(tmp);
    tmp = nullptr;
    XPT_TRY(dev_alloc(ctx, &tmp, ctas.size() * sizeof(StripCta) / sizeof(float)));
    ctx->strip_ctas = reinterpret_cast<StripCta*>(tmp);
  }
  XPT_CUDA(cudaMemcpy(ctx->strip_pieces, pieces.data(), pieces.size() * sizeof(StripPiece), cudaMemcpyHostToDevice));
  XPT_CUDA(cudaMemcpy(ctx->strip_ctas, ctas.data(), ctas.size() * sizeof(StripCta), cudaMemcpyHostToDevice));
  XPT_TRY(dev_alloc(ctx, &ctx->strip_loss_part, (size_t)ctx->B * slots * 3));
  XPT_TRY(dev_alloc(ctx, &ctx->strip_pose_part, (size_t)ctx->B * slots * ctx->N * 12));
  XPT_CUDA(cudaMemset(ctx->strip_loss_part, 0, (size_t)ctx->B * slots * 3 * sizeof(float)));
  XPT_CUDA(cudaMemset(ctx->strip_pose_part, 0, (size_t)ctx->B * slots * ctx->N * 12 * sizeof(float)));
  if (getenv("XPT_STRIP_VERBOSE"))
    fprintf(stderr, "xptwarp: strip partition B=%d: %d runs x %d snippets on %d CTAs, %zu pieces, %lld chunks per snippet, largest CTA %lld chunks (ideal %.1f)\n",
            ctx->B, R, ctx->B, nctas, pieces.size(), snip, best_ms, (double)snip * ctx->B / nctas);
  ctx->strip_ready = true;
  return XPT_OK;
}

template <int NS, bool DERIVE>
int launch_strip_t(xpt_ctx* ctx, const StripArgs& a, cudaStream_t st) {
  int per_sm = 0;
  XPT_TRY((strip_grid<NS, DERIVE>(ctx->cfg.device, &per_sm)));       // sets the dynamic shared-memory attribute once
  const bool prof = ctx->prof_on > 0 && ctx->prof_count < ctx->prof_on && ctx->prof_kind == XPT_PROFILE_FUSED;
  if (prof) XPT_CUDA(cudaEventRecord((*ctx->prof_events)[2 * ctx->prof_count], st));
  k_strip<NS, DERIVE><<<ctx->strip_nctas, StripSmem<NS>::kThreads, StripSmem<NS>::kBytes, st>>>(a);
  XPT_LAUNCH_CHECK("k_strip");
  if (prof) { XPT_CUDA(cudaEventRecord((*ctx->prof_events)[2 * ctx->prof_count + 1], st)); ++ctx->prof_count; }
  return XPT_OK;
}

int launch_strip(xpt_ctx* ctx, const StripArgs& a, bool derive, cudaStream_t st) {
  switch (ctx->strip_ns * 2 + (derive ? 1 : 0)) {
    case 2: return launch_strip_t<1, false>(ctx, a, st);
    case 3: return launch_strip_t<1, true>(ctx, a, st);
    case 4: return launch_strip_t<2, false>(ctx, a, st);
    case 5: return launch_strip_t<2, true>(ctx, a, st);
    case 8: return launch_strip_t<4, false>(ctx, a, st);
    default: return launch_strip_t<4, true>(ctx, a, st);
  }
}

int launch_smooth(xpt_ctx* ctx, const float* const disp_ms[], const LevelTable& lt, const float* gbatch,
                  float gcoef, float* const d_disp_ms[], int slot_offset, cudaStream_t st) {
  SmoothArgs a;
  memset(&a, 0, sizeof(a));
  PhotoArgs nrm;
  memset(&nrm, 0, sizeof(nrm));
  fill_photo_norms(ctx, nrm);
  a.S = ctx->S; a.B = ctx->B; a.grad_factor = ctx->cfg.img_grad_factor; a.gbatch = gbatch;
  a.loss_part = ctx->loss_part; a.slots_per_b = ctx->slots_per_b;
  int maxchunks = 0;
  for (int l = 0; l < ctx->S; ++l) {
    a.H[l] = ctx->h[l]; a.W[l] = ctx->w[l];
    a.disp[l] = disp_ms[l]; a.tgt[l] = lt.lv[l].tgt; a.tgt_bs[l] = lt.lv[l].tgt_bs;
    // gcoef folds dTotal/d(loss) into the per-level normaliser of the backward only
    a.norm_x[l] = nrm.norm_sm_x[l]; a.norm_y[l] = nrm.norm_sm_y[l];
    a.d_disp[l] = d_disp_ms ? d_disp_ms[l] : nullptr;
    a.chunk_base[l] = slot_offset + ctx->first_sm_chunk[l];
    if (ctx->sm_chunks[l] > maxchunks) maxchunks = ctx->sm_chunks[l];
  }
  (void)gcoef;
  dim3 grid(maxchunks, ctx->B, ctx->S);
  k_smooth<<<grid, 256, 0, st>>>(a);
  XPT_LAUNCH_CHECK("k_smooth");
  return XPT_OK;
}

int launch_loss_epilogue(xpt_ctx* ctx, int slots_used, float w0, float w1, float w2, float* losses,
                         float* loss_batch, cudaStream_t st) {
  int gb = ctx->cfg.global_batch > 0 ? ctx->cfg.global_batch : ctx->B;
  k_loss_epilogue<<<1, 256, 0, st>>>(ctx->loss_part, ctx->slots_per_b, slots_used, ctx->B, 1.0f / (float)gb, w0, w1,
                                     w2, losses, loss_batch);
  XPT_LAUNCH_CHECK("k_loss_epilogue");
  return XPT_OK;
}

}  // namespace

// ---- NCCL, bound at run time ---------------------------------------------------------------------------------
struct NcclId { char b[128]; };          // ncclUniqueId (passed BY VALUE to ncclCommInitRank)
struct NcclApi {
  void* so = nullptr;
  int (*GetUniqueId)(void*) = nullptr;
  int (*CommInitRank)(void**, int, NcclId, int) = nullptr;
  int (*CommDestroy)(void*) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, void*, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
};
typedef decltype(NcclApi::CommInitRank) nccl_init_fn;

NcclApi* nccl_api() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    // the copy already loaded by the host process (PyTorch ships libnccl.so.2) wins, then the system one
    void* so = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
    if (!so) so = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!so) so = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!so) return;
    api.GetUniqueId = reinterpret_cast<int (*)(void*)>(dlsym(so, "ncclGetUniqueId"));
    api.CommInitRank = reinterpret_cast<nccl_init_fn>(dlsym(so, "ncclCommInitRank"));
    api.CommDestroy = reinterpret_cast<int (*)(void*)>(dlsym(so, "ncclCommDestroy"));
    api.AllReduce = reinterpret_cast<int (*)(const void*, void*, size_t, int, int, void*, cudaStream_t)>(dlsym(so, "ncclAllReduce"));
    api.AllGather = reinterpret_cast<int (*)(const void*, void*, size_t, int, void*, cudaStream_t)>(dlsym(so, "ncclAllGather"));
    api.GroupStart = reinterpret_cast<int (*)()>(dlsym(so, "ncclGroupStart"));
    api.GroupEnd = reinterpret_cast<int (*)()>(dlsym(so, "ncclGroupEnd"));
    api.GetErrorString = reinterpret_cast<const char* (*)(int)>(dlsym(so, "ncclGetErrorString"));
    if (api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.AllReduce && api.GroupStart && api.GroupEnd) api.so = so;
  });
  return api.so ? &api : nullptr;
}

int nccl_fail(NcclApi* n, int rc, const char* what) {
  return fail(XPT_NCCL_ERROR, "%s failed: %s", what, (n && n->GetErrorString) ? n->GetErrorString(rc) : "NCCL error");
}
constexpr int kNcclFloat32 = 7, kNcclSum = 0;

int allreduce_on(xpt_ctx* ctx, float* const bufs[], const int64_t counts[], int num, cudaStream_t st) {
  NcclApi* n = nccl_api();
  if (!n) return fail(XPT_NCCL_ERROR, "libnccl.so.2 not found (dlopen)");
  if (!ctx->nccl_comm) return fail(XPT_BAD_ARGUMENT, "no communicator bound: call xpt_comm_init or xpt_comm_attach first");
  int rc = n->GroupStart();
  if (rc != 0) return nccl_fail(n, rc, "ncclGroupStart");
  for (int i = 0; i < num; ++i) {
    if (!bufs[i] || counts[i] < 0) { n->GroupEnd(); return fail(XPT_BAD_ARGUMENT, "xpt_allreduce: bufs[%d] is NULL or has a negative count", i); }
    if (counts[i] == 0) continue;
    rc = n->AllReduce(bufs[i], bufs[i], (size_t)counts[i], kNcclFloat32, kNcclSum, ctx->nccl_comm, st);
    if (rc != 0) { n->GroupEnd(); return nccl_fail(n, rc, "ncclAllReduce"); }
  }
  rc = n->GroupEnd();
  if (rc != 0) return nccl_fail(n, rc, "ncclGroupEnd");
  return XPT_OK;
}

// exchange fields of the fused epilogue: the losses are summed over the ranks inside k_epilogue when the ctx asks for
// the collective (XPT_FLAG_ALLREDUCE), the peer inboxes are mapped, and the call is not a chunk of the host entry point
void set_exchange(xpt_ctx* ctx, EpilogueArgs& ea) {
  ctx->exchanged = false;
  if (!(ctx->cfg.flags & XPT_FLAG_ALLREDUCE) || ctx->host_call || !ctx->p2p || ctx->nranks < 2) return;
  ea.nranks = ctx->nranks; ea.rank = ctx->rank;
  ea.peer_inbox = ctx->peer_table; ea.seq = ctx->exch_seq; ea.exchange_error = ctx->exch_err;
  ea.spin_budget = 4000000000LL;           // ~2 s at 1.9 GHz
  ctx->exchanged = true;
}

struct ScaleArgs { const float* src[16]; float* dst[16]; long long count[16]; int n; const float* scale; };
__global__ void __launch_bounds__(256) k_scale_tensors(ScaleArgs a) {
  const int seg = blockIdx.y;
  if (seg >= a.n) return;
  const float sc = __ldg(a.scale);
  const float* __restrict__ src = a.src[seg];
  float* __restrict__ dst = a.dst[seg];
  const long long n = a.count[seg], stride = (long long)gridDim.x * blockDim.x, t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if ((((uintptr_t)src | (uintptr_t)dst) & 15u) == 0) {
    const long long n4 = n >> 2;
    for (long long i = t; i < n4; i += stride) {
      float4 v = __ldg(reinterpret_cast<const float4*>(src) + i);
      v.x *= sc; v.y *= sc; v.z *= sc; v.w *= sc;
      reinterpret_cast<float4*>(dst)[i] = v;
    }
    for (long long i = (n4 << 2) + t; i < n; i += stride) dst[i] = src[i] * sc;
  } else {
    for (long long i = t; i < n; i += stride) dst[i] = src[i] * sc;
  }
}

// dlpack.h v0.8 (ABI-stable): just enough of DLManagedTensor to validate a tensor
struct DlTensorView {
  void* data; int32_t device_type, device_id; int32_t ndim; uint8_t code, bits; uint16_t lanes;
  int64_t* shape; int64_t* strides; uint64_t byte_offset;
};

// ===========================================================================
// C-ABI
// ===========================================================================
extern "C" {

int xpt_version(void) { return XPT_VERSION; }
#ifdef XPT_STRIP_PROF
// profiling build only (profiles/strip_roles_time.py): busy cycles of every warp of the last k_strip launch
__attribute__((visibility("default"))) int xpt_debug_strip_busy(long long* out, int n) {
  return cudaMemcpyFromSymbol(out, g_strip_busy, sizeof(long long) * n) == cudaSuccess ? 0 : -1;
}
#endif
const char* xpt_last_error(void) { return g_err; }

const char* xpt_status_string(int status) {
  switch (status) {
    case XPT_OK: return "XPT_OK";
    case XPT_BAD_ARGUMENT: return "XPT_BAD_ARGUMENT";
    case XPT_BAD_SHAPE: return "XPT_BAD_SHAPE";
    case XPT_CUDA_ERROR: return "XPT_CUDA_ERROR";
    case XPT_NO_DEVICE: return "XPT_NO_DEVICE";
    case XPT_OUT_OF_MEMORY: return "XPT_OUT_OF_MEMORY";
    case XPT_BAD_DTYPE: return "XPT_BAD_DTYPE";
    case XPT_BAD_DEVICE: return "XPT_BAD_DEVICE";
    case XPT_NOT_CONTIGUOUS: return "XPT_NOT_CONTIGUOUS";
    case XPT_NCCL_ERROR: return "XPT_NCCL_ERROR";
    default: return "XPT_UNKNOWN";
  }
}

int xpt_create(xpt_ctx** out, const xpt_config* cfg) {
  if (!out || !cfg) return fail(XPT_BAD_ARGUMENT, "xpt_create: NULL argument");
  *out = nullptr;
  if (cfg->batch < 1 || cfg->num_src < 1 || cfg->num_src > kMaxSrc)
    return fail(XPT_BAD_SHAPE, "batch=%d num_src=%d out of range (num_src <= %d)", cfg->batch, cfg->num_src, kMaxSrc);
  if (cfg->num_scales < 1 || cfg->num_scales > XPT_MAX_SCALES)
    return fail(XPT_BAD_SHAPE, "num_scales=%d out of range", cfg->num_scales);
  for (int l = 0; l < cfg->num_scales; ++l) {
    int s = cfg->scales[l];
    if (s < 1 || cfg->height % s || cfg->width % s)
      return fail(XPT_BAD_SHAPE, "scale %d does not divide %dx%d", s, cfg->height, cfg->width);
    if (cfg->height / s < 2 || cfg->width / s < 2)
      return fail(XPT_BAD_SHAPE, "level %d (%dx%d) smaller than 2x2", l, cfg->height / s, cfg->width / s);
  }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    (void)cudaGetLastError();
    return fail(XPT_NO_DEVICE, "no CUDA device visible (libxptwarp has no CPU fallback)");
  }
  if (cfg->device < 0 || cfg->device >= ndev) return fail(XPT_NO_DEVICE, "device %d out of range (%d visible)", cfg->device, ndev);
  cudaDeviceProp prop;
  XPT_CUDA(cudaGetDeviceProperties(&prop, cfg->device));
  if (prop.major != 10)
    return fail(XPT_NO_DEVICE, "device %d is sm_%d%d; libxptwarp is built for sm_100a only", cfg->device, prop.major, prop.minor);
  XPT_CUDA(cudaSetDevice(cfg->device));

  xpt_ctx* ctx = new (std::nothrow) xpt_ctx;
  if (!ctx) return fail(XPT_OUT_OF_MEMORY, "host allocation failed");
  memset(ctx, 0, sizeof(*ctx));
  ctx->cfg = *cfg;
  if (ctx->cfg.global_batch <= 0) ctx->cfg.global_batch = cfg->batch;
  ctx->S = cfg->num_scales; ctx->B = cfg->batch; ctx->N = cfg->num_src; ctx->H = cfg->height; ctx->W = cfg->width;
  ctx->first_tile[0] = ctx->first_chunk[0] = ctx->first_sm_chunk[0] = ctx->ffirst_tile[0] = 0;
  for (int l = 0; l < ctx->S; ++l) {
    ctx->s[l] = cfg->scales[l];
    ctx->h[l] = ctx->H / ctx->s[l]; ctx->w[l] = ctx->W / ctx->s[l];
    ctx->tiles_x[l] = cdiv(ctx->w[l], kTW); ctx->tiles_y[l] = cdiv(ctx->h[l], kTH);
    ctx->first_tile[l + 1] = ctx->first_tile[l] + ctx->tiles_x[l] * ctx->tiles_y[l];
    ctx->ftiles_x[l] = cdiv(ctx->w[l], kFCW); ctx->ftiles_y[l] = cdiv(ctx->h[l], kFCH);
    ctx->ffirst_tile[l + 1] = ctx->ffirst_tile[l] + ctx->ftiles_x[l] * ctx->ftiles_y[l];
    ctx->chunks[l] = cdiv((long long)ctx->h[l] * ctx->w[l], kWarpBwdChunk);
    ctx->first_chunk[l + 1] = ctx->first_chunk[l] + ctx->chunks[l];
    ctx->sm_chunks[l] = cdiv((long long)ctx->h[l] * ctx->w[l], 256);
    ctx->first_sm_chunk[l + 1] = ctx->first_sm_chunk[l] + ctx->sm_chunks[l];
  }
  ctx->slots_per_b = ctx->first_tile[ctx->S] + ctx->first_sm_chunk[ctx->S];
  if (ctx->first_chunk[ctx->S] > ctx->slots_per_b) ctx->slots_per_b = ctx->first_chunk[ctx->S];
  if (ctx->ffirst_tile[ctx->S] > ctx->slots_per_b) ctx->slots_per_b = ctx->ffirst_tile[ctx->S];

  int rc = XPT_OK;
  auto A = [&](float** p, size_t n) { if (rc == XPT_OK) rc = dev_alloc(ctx, p, n); };
  A(&ctx->geoK, (size_t)ctx->B * ctx->S * kGeoK + (size_t)ctx->B * ctx->N * kGeoT);   // K block, then [R|t] block
  if (rc == XPT_OK) ctx->geoT = ctx->geoK + (size_t)ctx->B * ctx->S * kGeoK;
  {
    float* tmp = nullptr;
    A(&tmp, (size_t)ctx->B * 6 + 2);          // [B][3] doubles + ticket
    if (rc == XPT_OK) {
      ctx->loss_sum_b = reinterpret_cast<double*>(tmp);
      ctx->ticket = reinterpret_cast<unsigned int*>(tmp + (size_t)ctx->B * 6);
      if (cudaMemset(ctx->ticket, 0, sizeof(unsigned int)) != cudaSuccess) rc = fail(XPT_CUDA_ERROR, "cudaMemset failed");
    }
  }
  for (int l = 0; l < ctx->S; ++l)
    if (ctx->s[l] > 1) {
      A(&ctx->src_pyr[l], (size_t)ctx->B * ctx->N * lvl_pix(ctx, l) * 3);
      A(&ctx->tgt_pyr[l], (size_t)ctx->B * lvl_pix(ctx, l) * 3);
    }
  A(&ctx->loss_part, (size_t)ctx->B * ctx->slots_per_b * 3);
  A(&ctx->pose_part, (size_t)ctx->B * ctx->slots_per_b * ctx->N * 12);
  if (rc != XPT_OK) { xpt_destroy(ctx); return rc; }
  *out = ctx;
  return XPT_OK;
}

// unmap the peers' inboxes, free this rank's, destroy an owned communicator
static void comm_release(xpt_ctx* ctx) {
  if (ctx->peer_host) {
    for (int r = 0; r < ctx->nranks; ++r)
      if (r != ctx->rank && ctx->peer_host[r]) cudaIpcCloseMemHandle(ctx->peer_host[r]);
    delete[] ctx->peer_host; ctx->peer_host = nullptr;
  }
  if (ctx->inbox) { cudaFree(ctx->inbox); ctx->inbox = nullptr; }
  if (ctx->peer_table) { cudaFree(ctx->peer_table); ctx->peer_table = nullptr; }
  if (ctx->exch_seq) { cudaFree(ctx->exch_seq); ctx->exch_seq = nullptr; ctx->exch_err = nullptr; }
  ctx->p2p = false;
  if (ctx->nccl_comm && ctx->nccl_owned) { NcclApi* n = nccl_api(); if (n) n->CommDestroy(ctx->nccl_comm); }
  ctx->nccl_comm = nullptr; ctx->nccl_owned = false;
}

void xpt_destroy(xpt_ctx* ctx) {
  if (!ctx) return;
  geo_slot_release(ctx);
  cudaSetDevice(ctx->cfg.device);
  // captured steps hold references on the communicator (NCCL waits for them in ncclCommDestroy): graphs go first
  if (ctx->graphs) {
    for (auto& e : *ctx->graphs) cudaGraphExecDestroy(e.exec);
    delete ctx->graphs; ctx->graphs = nullptr;
  }
  if (ctx->host_graphs) {
    for (auto& e : *ctx->host_graphs) cudaGraphExecDestroy(e.exec);
    delete ctx->host_graphs; ctx->host_graphs = nullptr;
  }
  comm_release(ctx);
  auto F = [](float* p) { if (p) cudaFree(p); };
  F(ctx->geoK); F(reinterpret_cast<float*>(ctx->strip_pieces)); F(reinterpret_cast<float*>(ctx->strip_ctas)); F(ctx->strip_loss_part); F(ctx->strip_pose_part); F(reinterpret_cast<float*>(ctx->loss_sum_b)); F(ctx->loss_part); F(ctx->pose_part); F(ctx->tgt0_copy); F(ctx->min_part); F(ctx->l2_part);
  F(ctx->st_frames); F(ctx->st_K); F(ctx->st_pose); F(ctx->st_losses); F(ctx->st_loss_batch); F(ctx->st_dpose);
  F(ctx->st_dsource);
  for (int l = 0; l < kMaxScales; ++l) {
    F(ctx->src_pyr[l]); F(ctx->src4_pyr[l]); F(ctx->tgt_pyr[l]); F(ctx->synth_scr[l]); F(ctx->gsynth_scr[l]); F(ctx->dsrc_lvl[l]); F(ctx->dsrc4_lvl[l]);
    F(ctx->st_depth[l]); F(ctx->st_disp[l]); F(ctx->st_ddepth[l]); F(ctx->st_ddisp[l]);
    F(ctx->st_synth[l]); F(ctx->st_mask[l]); F(ctx->st_target[l]);
  }
  if (ctx->child) xpt_destroy(ctx->child);
  if (ctx->h_losses) cudaFreeHost(ctx->h_losses);
  if (ctx->host_done_ev) cudaEventDestroy(ctx->host_done_ev);
  if (ctx->s_in) {
    cudaStreamDestroy(ctx->s_in); cudaStreamDestroy(ctx->s_out); cudaStreamDestroy(ctx->s_in2); cudaEventDestroy(ctx->ev_small);
    cudaEventDestroy(ctx->ev_fork); cudaEventDestroy(ctx->ev_join);
    for (int k = 0; k < kMaxChunks; ++k) { cudaEventDestroy(ctx->ev_in[k]); cudaEventDestroy(ctx->ev_done[k]); }
  }
  if (ctx->prof_events) {
    for (auto e : *ctx->prof_events) cudaEventDestroy(e);
    delete ctx->prof_events;
  }
  delete ctx;
}

int xpt_get_config(const xpt_ctx* ctx, xpt_config* out) {
  if (!ctx || !out) return fail(XPT_BAD_ARGUMENT, "xpt_get_config: NULL argument");
  *out = ctx->cfg;
  return XPT_OK;
}

size_t xpt_scratch_bytes(const xpt_ctx* ctx) { return ctx ? ctx->scratch_bytes : 0; }
int xpt_last_launch_count(const xpt_ctx* ctx) { return ctx ? ctx->launches : 0; }
int xpt_geometry_slot_shared(const xpt_ctx* ctx) { return (ctx && ctx->geo_shared) ? 1 : 0; }

int xpt_check_dlpack(const void* dl_managed_tensor, int device, int dense_from_dim) {
  if (!dl_managed_tensor) return fail(XPT_BAD_ARGUMENT, "xpt_check_dlpack: NULL tensor");
  const DlTensorView* t = static_cast<const DlTensorView*>(dl_managed_tensor);
  if (t->code != 2 /* kDLFloat */ || t->bits != 32 || t->lanes != 1)
    return fail(XPT_BAD_DTYPE, "tensor is not float32 (DLPack dtype code=%d bits=%d lanes=%d)", t->code, t->bits, t->lanes);
  if (t->device_type != 2 /* kDLCUDA */) return fail(XPT_BAD_DEVICE, "tensor is not on a CUDA device (DLPack device_type %d): there is no CPU path", t->device_type);
  if (device >= 0 && t->device_id != device) return fail(XPT_BAD_DEVICE, "tensor lives on cuda:%d, the ctx on cuda:%d", t->device_id, device);
  if (t->strides) {
    int64_t expect = 1;
    for (int d = t->ndim - 1; d >= (dense_from_dim < 0 ? 0 : dense_from_dim); --d) {
      if (t->shape[d] != 1 && t->strides[d] != expect)
        return fail(XPT_NOT_CONTIGUOUS, "dimension %d has stride %lld, dense layout needs %lld", d, (long long)t->strides[d], (long long)expect);
      expect *= t->shape[d];
    }
  }
  return XPT_OK;
}

int xpt_scale_tensors(int device, const float* const src[], float* const dst[], const int64_t counts[], int num,
                      const float* scale, void* stream) {
  if (num < 0 || (num > 0 && (!src || !dst || !counts)) || !scale) return fail(XPT_BAD_ARGUMENT, "xpt_scale_tensors: NULL argument");
  XPT_CUDA(cudaSetDevice(device));
  for (int i0 = 0; i0 < num; i0 += 16) {
    ScaleArgs a;
    memset(&a, 0, sizeof(a));
    a.scale = scale;
    long long mx = 0;
    for (int i = i0; i < num && i < i0 + 16; ++i) {
      if (!src[i] || !dst[i] || counts[i] < 0) return fail(XPT_BAD_ARGUMENT, "xpt_scale_tensors: tensor %d is NULL or has a negative count", i);
      a.src[a.n] = src[i]; a.dst[a.n] = dst[i]; a.count[a.n] = counts[i];
      if (counts[i] > mx) mx = counts[i];
      ++a.n;
    }
    if (mx == 0) continue;
    int gx = cdiv(mx, 256 * 8);
    if (gx > 1184) gx = 1184;
    k_scale_tensors<<<dim3(gx, a.n), 256, 0, (cudaStream_t)stream>>>(a);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(XPT_CUDA_ERROR, "launch of k_scale_tensors failed: %s", cudaGetErrorString(e));
  }
  return XPT_OK;
}

// Map every rank's loss inbox into this process (CUDA IPC over NVLink peer access).  The 64-byte handles travel through
// one ncclAllGather on the new communicator; a second tiny all-reduce makes the ranks agree: either all of them use
// the peer-memory exchange of k_epilogue, or all fall back to ncclAllReduce behind the epilogue (XPT_NO_P2P=1 forces it).
static int p2p_setup(xpt_ctx* ctx, NcclApi* n) {
  const int R = ctx->nranks, me = ctx->rank;
  ctx->p2p = false;
  if (R < 2) return XPT_OK;
  if (R > kMaxRanks) return fail(XPT_BAD_ARGUMENT, "xpt_comm_init: more than %d ranks", kMaxRanks);
  float ok = (n->AllGather && !getenv("XPT_NO_P2P")) ? 1.f : 0.f;
  const size_t inbox_floats = (size_t)2 * R * kInboxSlot;
  unsigned char* hbuf = nullptr;          // device: [R][64] handles + 1 float of agreement
  cudaIpcMemHandle_t mine;
  memset(&mine, 0, sizeof(mine));
  XPT_CUDA(cudaMalloc(&ctx->inbox, inbox_floats * sizeof(float)));
  XPT_CUDA(cudaMemset(ctx->inbox, 0, inbox_floats * sizeof(float)));
  XPT_CUDA(cudaMalloc(&ctx->exch_seq, 2 * sizeof(unsigned)));
  XPT_CUDA(cudaMemset(ctx->exch_seq, 0, 2 * sizeof(unsigned)));
  ctx->exch_err = ctx->exch_seq + 1;
  if (cudaIpcGetMemHandle(&mine, ctx->inbox) != cudaSuccess) { (void)cudaGetLastError(); ok = 0.f; }
  XPT_CUDA(cudaMalloc(&hbuf, (size_t)R * 64 + 16));
  std::vector<cudaIpcMemHandle_t> all(R);
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle size");
  if (n->AllGather) {
    XPT_CUDA(cudaMemcpy(hbuf + (size_t)me * 64, &mine, 64, cudaMemcpyHostToDevice));
    int rc = n->AllGather(hbuf + (size_t)me * 64, hbuf, 64, /*ncclUint8*/ 1, ctx->nccl_comm, nullptr);
    if (rc != 0) { cudaFree(hbuf); return nccl_fail(n, rc, "ncclAllGather"); }
    XPT_CUDA(cudaStreamSynchronize(nullptr));
    XPT_CUDA(cudaMemcpy(all.data(), hbuf, (size_t)R * 64, cudaMemcpyDeviceToHost));
  }
  ctx->peer_host = new float*[R]();
  for (int r = 0; r < R && ok != 0.f; ++r) {
    if (r == me) { ctx->peer_host[r] = ctx->inbox; continue; }
    void* p = nullptr;
    if (cudaIpcOpenMemHandle(&p, all[r], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { (void)cudaGetLastError(); ok = 0.f; break; }
    ctx->peer_host[r] = static_cast<float*>(p);
  }
  // agreement: min over the ranks (sum of R ones == R)
  float* agree = reinterpret_cast<float*>(hbuf + (size_t)R * 64);
  XPT_CUDA(cudaMemcpy(agree, &ok, sizeof(float), cudaMemcpyHostToDevice));
  int rc = n->AllReduce(agree, agree, 1, kNcclFloat32, kNcclSum, ctx->nccl_comm, nullptr);
  if (rc != 0) { cudaFree(hbuf); return nccl_fail(n, rc, "ncclAllReduce"); }
  XPT_CUDA(cudaStreamSynchronize(nullptr));
  float sum = 0.f;
  XPT_CUDA(cudaMemcpy(&sum, agree, sizeof(float), cudaMemcpyDeviceToHost));
  cudaFree(hbuf);
  if (sum == (float)R) {
    XPT_CUDA(cudaMalloc(&ctx->peer_table, (size_t)R * sizeof(float*)));
    XPT_CUDA(cudaMemcpy(ctx->peer_table, ctx->peer_host, (size_t)R * sizeof(float*), cudaMemcpyHostToDevice));
    ctx->p2p = true;
  }
  return XPT_OK;
}

int xpt_comm_unique_id(unsigned char id[128]) {
  if (!id) return fail(XPT_BAD_ARGUMENT, "xpt_comm_unique_id: NULL id");
  NcclApi* n = nccl_api();
  if (!n) return fail(XPT_NCCL_ERROR, "libnccl.so.2 not found (dlopen)");
  const int rc = n->GetUniqueId(id);
  return rc == 0 ? XPT_OK : nccl_fail(n, rc, "ncclGetUniqueId");
}

int xpt_comm_init(xpt_ctx* ctx, const unsigned char id[128], int nranks, int rank) {
  if (!ctx || !id || nranks < 1 || rank < 0 || rank >= nranks) return fail(XPT_BAD_ARGUMENT, "xpt_comm_init: bad argument");
  NcclApi* n = nccl_api();
  if (!n) return fail(XPT_NCCL_ERROR, "libnccl.so.2 not found (dlopen)");
  XPT_CUDA(cudaSetDevice(ctx->cfg.device));
  comm_release(ctx);
  NcclId uid;
  memcpy(uid.b, id, 128);
  void* comm = nullptr;
  const int rc = n->CommInitRank(&comm, nranks, uid, rank);
  if (rc != 0) return nccl_fail(n, rc, "ncclCommInitRank");
  ctx->nccl_comm = comm; ctx->nccl_owned = true;
  ctx->nranks = nranks; ctx->rank = rank;
  return p2p_setup(ctx, n);
}

int xpt_comm_destroy(xpt_ctx* ctx) {
  if (!ctx) return fail(XPT_BAD_ARGUMENT, "xpt_comm_destroy: NULL ctx");
  XPT_CUDA(cudaSetDevice(ctx->cfg.device));
  XPT_CUDA(cudaDeviceSynchronize());
  // graphs that captured the collective reference the communicator
  if (ctx->graphs) { for (auto& e : *ctx->graphs) cudaGraphExecDestroy(e.exec); ctx->graphs->clear(); }
  if (ctx->host_graphs) { for (auto& e : *ctx->host_graphs) cudaGraphExecDestroy(e.exec); ctx->host_graphs->clear(); }
  comm_release(ctx);
  return XPT_OK;
}

int xpt_comm_status(xpt_ctx* ctx, int* uses_peer_memory, int* exchange_error) {
  if (!ctx) return fail(XPT_BAD_ARGUMENT, "xpt_comm_status: NULL ctx");
  if (uses_peer_memory) *uses_peer_memory = ctx->p2p ? 1 : 0;
  if (exchange_error) {
    *exchange_error = 0;
    if (ctx->exch_err) {
      XPT_CUDA(cudaSetDevice(ctx->cfg.device));
      unsigned v = 0;
      XPT_CUDA(cudaMemcpy(&v, ctx->exch_err, sizeof(v), cudaMemcpyDeviceToHost));
      *exchange_error = (int)v;
    }
  }
  return XPT_OK;
}

int xpt_comm_attach(xpt_ctx* ctx, void* nccl_comm) {
  if (!ctx) return fail(XPT_BAD_ARGUMENT, "xpt_comm_attach: NULL ctx");
  cudaSetDevice(ctx->cfg.device);
  comm_release(ctx);                 // an attached communicator carries no rank table: the losses go through ncclAllReduce
  ctx->nccl_comm = nccl_comm; ctx->nccl_owned = false;
  return XPT_OK;
}

int xpt_allreduce(xpt_ctx* ctx, float* const bufs[], const int64_t counts[], int num, void* stream) {
  if (!ctx || num < 0 || (num > 0 && (!bufs || !counts))) return fail(XPT_BAD_ARGUMENT, "xpt_allreduce: NULL argument");
  XPT_CUDA(cudaSetDevice(ctx->cfg.device));
  return allreduce_on(ctx, bufs, counts, num, (cudaStream_t)stream);
}

int xpt_pose_matr2rvec(int device, const float* matr, int count, int invert, float* rvec, void* stream) {
  if (!matr || !rvec || count < 0) return fail(XPT_BAD_ARGUMENT, "xpt_pose_matr2rvec: bad argument");
  XPT_CUDA(cudaSetDevice(device));
  if (count == 0) return XPT_OK;
  k_pose_matr2rvec<<<cdiv(count, 128), 128, 0, (cudaStream_t)stream>>>(matr, count, invert, rvec);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(XPT_CUDA_ERROR, "launch of k_pose_matr2rvec failed: %s", cudaGetErrorString(e));
  return XPT_OK;
}

int xpt_stereo_pose_loss(int device, const float* stereo_T_LR, const float* pose_lr, const float* pose_rl, int batch,
                         int num, float* loss_batch, const float* grad_loss_batch, float* d_pose_lr, float* d_pose_rl,
                         void* stream) {
  if (!stereo_T_LR || !pose_lr || !pose_rl || batch <= 0 || num <= 0)
    return fail(XPT_BAD_ARGUMENT, "xpt_stereo_pose_loss: bad argument");
  XPT_CUDA(cudaSetDevice(device));
  k_stereo_pose_loss<<<cdiv(batch, 64), 64, 0, (cudaStream_t)stream>>>(stereo_T_LR, pose_lr, pose_rl, batch, num, loss_batch,
                                                                        grad_loss_batch, d_pose_lr, d_pose_rl);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(XPT_CUDA_ERROR, "launch of k_stereo_pose_loss failed: %s", cudaGetErrorString(e));
  return XPT_OK;
}

int xpt_pose_rvec2matr(xpt_ctx* ctx, const float* pose, float* matr, void* stream) {
  if (!ctx || !pose || !matr) return fail(XPT_BAD_ARGUMENT, "xpt_pose_rvec2matr: NULL argument");
  XPT_CUDA(cudaSetDevice(ctx->cfg.device));
  ctx->launches = 0;
  return launch_geometry(ctx, pose, nullptr, matr, (cudaStream_t)stream);
}

int xpt_build_pyramids(xpt_ctx* ctx, const xpt_frames* frames, float* const target_ms[], void* stream) {
  if (!ctx) return fail(XPT_BAD_ARGUMENT, "ctx is NULL");
  XPT_TRY(check_frames(ctx, frames, false));
  XPT_CUDA(cudaSetDevice(ctx->cfg.device));
  ctx->launches = 0;
  return launch_pyramids(ctx, frames, target_ms, true, (cudaStream_t)stream);
}

int xpt_synthesize(xpt_ctx* ctx, const xpt_frames* frames, const float* const depth_ms[], const float* pose,
                   float* const synth_ms[], float* const mask_ms[], void* stream) {
  if (!ctx || !pose) return fail(XPT_BAD_ARGUMENT, "xpt_synthesize: NULL argument");
  XPT_TRY(check_frames(ctx, frames, false));
  XPT_TRY(check_list(ctx, (const void* const*)depth_ms, "depth_ms", true));
  XPT_TRY(check_list(ctx, (const void* const*)synth_ms, "synth_ms", true));
  cudaStream_t st = (cudaStream_t)stream;
  XPT_CUDA(cudaSetDevice(ctx->cfg.device));
  ctx->launches = 0;
  xpt_frames f = *frames;
  f.target = nullptr;                       // synthesis does not touch the target frame
  XPT_TRY(launch_geometry(ctx, pose, f.intrinsic, nullptr, st));
  XPT_TRY(launch_pyramids(ctx, &f, nullptr, true, st));
  LevelTable lt = make_levels(ctx, &f, nullptr);
  return launch_warp_fwd(ctx, lt, depth_ms, synth_ms, mask_ms, st);
}

int xpt_synthesize_backward(xpt_ctx* ctx, const xpt_frames* frames, const float* const depth_ms[], const float* pose,
                            const float* const grad_synth_ms[], float* const d_depth_ms[], float* d_pose,
                            float* d_source, void* stream) {
  if (!ctx || !pose) return fail(XPT_BAD_ARGUMENT, "xpt_synthesize_backward: NULL argument");
  XPT_TRY(check_frames(ctx, frames, false));
  XPT_TRY(check_list(ctx, (const void* const*)depth_ms, "depth_ms", true));
  XPT_TRY(check_list(ctx, (const void* const*)grad_synth_ms, "grad_synth_ms", true));
  cudaStream_t st = (cudaStream_t)stream;
  XPT_CUDA(cudaSetDevice(ctx->cfg.device));
  ctx->launches = 0;
  xpt_frames f = *frames;
  f.target = nullptr;
  XPT_TRY(launch_geometry(ctx, pose, f.intrinsic, nullptr, st));
  XPT_TRY(launch_pyramids(ctx, &f, nullptr, true, st));
  LevelTable lt = make_levels(ctx, &f, nullptr);
  return launch_warp_bwd(ctx, lt, depth_ms, grad_synth_ms, d_depth_ms, d_source, pose, d_pose, 1.0f, st);
}

int xpt_photometric_loss(xpt_ctx* ctx, int method, const float* const synth_ms[], const float* const target_ms[],
                         float* loss_batch, const float* grad_loss_batch, float* const d_synth_ms[], void* stream) {
  if (!ctx || !loss_batch) return fail(XPT_BAD_ARGUMENT, "xpt_photometric_loss: NULL argument");
  if (method != XPT_PHOTO_L1 && method != XPT_PHOTO_L2 && method != XPT_PHOTO_SSIM)
    return fail(XPT_BAD_ARGUMENT, "unknown photometric method %d", method);
  XPT_TRY(check_list(ctx, (const void* const*)synth_ms, "synth_ms", true));
  XPT_TRY(check_list(ctx, (const void* const*)target_ms, "target_ms", true));
  if (d_synth_ms) XPT_TRY(check_list(ctx, (const void* const*)d_synth_ms, "d_synth_ms", true));
  cudaStream_t st = (cudaStream_t)stream;
  XPT_CUDA(cudaSetDevice(ctx->cfg.device));
  ctx->launches = 0;
  PhotoArgs a;
  memset(&a, 0, sizeof(a));
  a.lt = make_levels(ctx, nullptr, target_ms);
  fill_photo_norms(ctx, a);
  a.l1_kind = method == XPT_PHOTO_L1 ? 1 : (method == XPT_PHOTO_L2 ? 2 : 0);
  a.do_ssim = method == XPT_PHOTO_SSIM;
  a.gcoef_l1 = 1.f; a.gcoef_ssim = 1.f; a.gbatch = grad_loss_batch;
  for (int l = 0; l < ctx->S; ++l) { a.synth[l] = synth_ms[l]; a.gsynth[l] = d_synth_ms ? d_synth_ms[l] : nullptr; }
  if (d_synth_ms) XPT_TRY((launch_photo<false, true>(ctx, a, st)));
  else XPT_TRY((launch_photo<false, false>(ctx, a, st)));
  // column 0 holds L1/L2, column 1 SSIM: select the requested one into loss_batch[B]
  float* lb3 = nullptr;
  XPT_TRY(dev_alloc(ctx, &ctx->st_loss_batch, (size_t)3 * ctx->B));
  lb3 = ctx->st_loss_batch;
  XPT_TRY(launch_loss_epilogue(ctx, ctx->first_tile[ctx->S], 0.f, 0.f, 0.f, nullptr, lb3, st));
  int col = method == XPT_PHOTO_SSIM ? 1 : 0;
  XPT_CUDA(cudaMemcpyAsync(loss_batch, lb3 + (size_t)col * ctx->B, ctx->B * sizeof(float), cudaMemcpyDeviceToDevice, st));
  return XPT_OK;
}

static int min_loss_impl(xpt_ctx* ctx, int method, const float* const synth_ms[], const float* const stereo_synth_ms[],
                         const float* cmb_flow, int cmb_h, int cmb_w,
                         const float* target, int64_t target_batch_stride, float* loss_batch,
                         const float* grad_loss_batch, float* const d_synth_ms[], float* const d_stereo_synth_ms[],
                         void* stream, float* loss_batch_pair_ssim = nullptr, float pair_c_l1 = 0.f, float pair_c_ssim = 0.f) {
  // loss_batch_pair_ssim != NULL: the L1 + SSIM pair in one launch (loss_batch receives L1); method is ignored
  const bool pair = loss_batch_pair_ssim != nullptr;
  if (!ctx || !loss_batch || !target) return fail(XPT_BAD_ARGUMENT, "xpt_photometric_min_loss: NULL argument");
  if (method != XPT_PHOTO_L1 && method != XPT_PHOTO_L2 && method != XPT_PHOTO_SSIM)
    return fail(XPT_BAD_ARGUMENT, "unknown photometric method %d", method);
  XPT_TRY(check_list(ctx, (const void* const*)synth_ms, "synth_ms", true));
  if (stereo_synth_ms) XPT_TRY(check_list(ctx, (const void* const*)stereo_synth_ms, "stereo_synth_ms", true));
  if (d_synth_ms) XPT_TRY(check_list(ctx, (const void* const*)d_synth_ms, "d_synth_ms", true));
  if (d_synth_ms && stereo_synth_ms) XPT_TRY(check_list(ctx, (const void* const*)d_stereo_synth_ms, "d_stereo_synth_ms", true));
  if (ctx->B > 1 && target_batch_stride < (int64_t)ctx->H * ctx->W * 3) return fail(XPT_BAD_SHAPE, "target_batch_stride too small");
  cudaStream_t st = (cudaStream_t)stream;
  XPT_CUDA(cudaSetDevice(ctx->cfg.device));
  ctx->launches = 0;
  MinLossArgs a;
  memset(&a, 0, sizeof(a));
  a.B = ctx->B; a.N = ctx->N; a.NS = stereo_synth_ms ? 1 : 0; a.H = ctx->H; a.W = ctx->W; a.S = ctx->S;
  a.target = target; a.tgt_bs = target_batch_stride;
  a.method = method == XPT_PHOTO_L1 ? 0 : (method == XPT_PHOTO_L2 ? 1 : 2);
  a.gbatch = grad_loss_batch;
  a.cmb_flow = cmb_flow; a.cmb_h = cmb_h; a.cmb_w = cmb_w;
  // the strip kernel (64x13 tiles) serves both modes; XPT_FLAG_MIN_TILES (A/B) and levels between full and half
  // resolution (footprint wider than the strip kernel's buffer) keep the 32x16 tile kernel
  bool strip = !(ctx->cfg.flags & XPT_FLAG_MIN_TILES);
  for (int l = 0; l < ctx->S; ++l)
    if (!((ctx->h[l] == ctx->H && ctx->w[l] == ctx->W) || (2 * ctx->w[l] <= ctx->W && 2 * ctx->h[l] <= ctx->H))) strip = false;
  a.tiles_x = strip ? cdiv(ctx->W, kFCW) : cdiv(ctx->W, kTW);
  a.tiles = a.tiles_x * (strip ? cdiv(ctx->H, kFCH) : cdiv(ctx->H, kTH));
  if (pair && (ctx->cfg.flags & XPT_FLAG_MIN_TILES))
    return fail(XPT_BAD_ARGUMENT, "the pair launches run on the strip kernel only: ctx created with XPT_FLAG_MIN_TILES");
  if (pair && !strip) return fail(XPT_BAD_SHAPE, "pair launch: every level must be at full or at most half resolution");
  const size_t n_part = (size_t)ctx->B * ctx->S * a.tiles;
  XPT_TRY(dev_alloc(ctx, &ctx->min_part, 2 * n_part));
  a.loss_part = ctx->min_part;
  a.loss_part2 = ctx->min_part + n_part;
  a.pair_c_l1 = pair_c_l1; a.pair_c_ssim = pair_c_ssim;
  for (int l = 0; l < ctx->S; ++l) {
    a.h[l] = ctx->h[l]; a.w[l] = ctx->w[l];
    a.synth[l] = synth_ms[l]; a.stereo[l] = stereo_synth_ms ? stereo_synth_ms[l] : nullptr;
    // min over sources: mean over [H,W,3]; combined loss: mean over [N,H,W,3] (losses.py:273)
    a.norm[l] = (float)((double)ctx->cfg.scale_weights[l] / ((double)ctx->H * ctx->W * 3.0 * (cmb_flow ? ctx->N : 1)));
    if (d_synth_ms) {
      a.gsynth[l] = d_synth_ms[l];
      XPT_CUDA(cudaMemsetAsync(d_synth_ms[l], 0, (size_t)ctx->B * ctx->N * lvl_pix(ctx, l) * 3 * sizeof(float), st));
      if (stereo_synth_ms) {
        a.gstereo[l] = d_stereo_synth_ms[l];
        XPT_CUDA(cudaMemsetAsync(d_stereo_synth_ms[l], 0, (size_t)ctx->B * lvl_pix(ctx, l) * 3 * sizeof(float), st));
      }
    }
  }
  dim3 grid(ctx->S * a.tiles, ctx->B);
  if (strip) {
    const dim3 sgrid(ctx->B, ctx->S * a.tiles);
    const bool grad = d_synth_ms != nullptr, cmb = cmb_flow != nullptr;
    // one instantiation per (gradient, pair, combined)
#define XPT_MIN_STRIP(G, P, C)                                                                              \
    do {                                                                                                    \
      static unsigned long long attr_done = 0;                                                              \
      const size_t smem = MinStripSmem<G, P>::kBytes;                                                       \
      XPT_TRY(ensure_dyn_smem(k_min_strip<G, P, C>, smem, ctx->cfg.device, &attr_done));                    \
      k_min_strip<G, P, C><<<sgrid, kFThreads, smem, st>>>(a);                                              \
    } while (0)
    if (grad) {
      if (pair) { if (cmb) XPT_MIN_STRIP(true, true, true); else XPT_MIN_STRIP(true, true, false); }
      else { if (cmb) XPT_MIN_STRIP(true, false, true); else XPT_MIN_STRIP(true, false, false); }
    } else {
      if (pair) { if (cmb) XPT_MIN_STRIP(false, true, true); else XPT_MIN_STRIP(false, true, false); }
      else { if (cmb) XPT_MIN_STRIP(false, false, true); else XPT_MIN_STRIP(false, false, false); }
    }
#undef XPT_MIN_STRIP
    XPT_LAUNCH_CHECK("k_min_strip");
  } else if (d_synth_ms) {
    static unsigned long long attr_done = 0;
    const size_t smem = MinLossSmem<true>::kBytes;
    XPT_TRY(ensure_dyn_smem(k_photo_min<true>, smem, ctx->cfg.device, &attr_done));
    k_photo_min<true><<<grid, kPhotoThreads, smem, st>>>(a);
  } else {
    static unsigned long long attr_done = 0;
    const size_t smem = MinLossSmem<false>::kBytes;
    XPT_TRY(ensure_dyn_smem(k_photo_min<false>, smem, ctx->cfg.device, &attr_done));
    k_photo_min<false><<<grid, kPhotoThreads, smem, st>>>(a);
  }
  if (!strip) XPT_LAUNCH_CHECK("k_photo_min");
  k_sum_slots<<<ctx->B, 128, 0, st>>>(ctx->min_part, ctx->S * a.tiles, loss_batch);
  XPT_LAUNCH_CHECK("k_sum_slots");
  if (pair) {
    k_sum_slots<<<ctx->B, 128, 0, st>>>(ctx->min_part + n_part, ctx->S * a.tiles, loss_batch_pair_ssim);
    XPT_LAUNCH_CHECK("k_sum_slots");
  }
  return XPT_OK;
}

int xpt_photometric_min_pair_loss(xpt_ctx* ctx, const float* const synth_ms[], const float* const stereo_synth_ms[],
                                  const float* target, int64_t target_batch_stride, float* loss_batch_l1,
                                  float* loss_batch_ssim, float grad_l1, float grad_ssim, float* const d_synth_ms[],
                                  float* const d_stereo_synth_ms[], void* stream) {
  if (!loss_batch_ssim) return fail(XPT_BAD_ARGUMENT, "xpt_photometric_min_pair_loss: NULL argument");
  return min_loss_impl(ctx, XPT_PHOTO_L1, synth_ms, stereo_synth_ms, nullptr, 0, 0, target, target_batch_stride, loss_batch_l1,
                       nullptr, d_synth_ms, d_stereo_synth_ms, stream, loss_batch_ssim, grad_l1, grad_ssim);
}

int xpt_photometric_min_loss(xpt_ctx* ctx, int method, const float* const synth_ms[], const float* const stereo_synth_ms[],
                             const float* target, int64_t target_batch_stride, float* loss_batch,
                             const float* grad_loss_batch, float* const d_synth_ms[], float* const d_stereo_synth_ms[],
                             void* stream) {
  return min_loss_impl(ctx, method, synth_ms, stereo_synth_ms, nullptr, 0, 0, target, target_batch_stride, loss_batch,
                       grad_loss_batch, d_synth_ms, d_stereo_synth_ms, stream);
}

int xpt_photometric_cmb_loss(xpt_ctx* ctx, int method, const float* const synth_ms[], const float* warped,
                             int warped_height, int warped_width, const float* target, int64_t target_batch_stride,
                             float* loss_batch, const float* grad_loss_batch, float* const d_synth_ms[], void* stream) {
  if (!warped) return fail(XPT_BAD_ARGUMENT, "xpt_photometric_cmb_loss: warped is NULL");
  if (warped_height < 1 || warped_width < 1) return fail(XPT_BAD_SHAPE, "warped view %dx%d", warped_height, warped_width);
  return min_loss_impl(ctx, method, synth_ms, nullptr, warped, warped_height, warped_width, target, target_batch_stride,
                       loss_batch, grad_loss_batch, d_synth_ms, nullptr, stream);
}

int xpt_photometric_cmb_pair_loss(xpt_ctx* ctx, const float* const synth_ms[], const float* warped, int warped_height,
                                  int warped_width, const float* target, int64_t target_batch_stride, float* loss_batch_l1,
                                  float* loss_batch_ssim, float grad_l1, float grad_ssim, float* const d_synth_ms[],
                                  void* stream) {
  if (!warped || !loss_batch_ssim) return fail(XPT_BAD_ARGUMENT, "xpt_photometric_cmb_pair_loss: NULL argument");
  if (warped_height < 1 || warped_width < 1) return fail(XPT_BAD_SHAPE, "warped view %dx%d", warped_height, warped_width);
  return min_loss_impl(ctx, XPT_PHOTO_L1, synth_ms, nullptr, warped, warped_height, warped_width, target, target_batch_stride,
                       loss_batch_l1, nullptr, d_synth_ms, nullptr, stream, loss_batch_ssim, grad_l1, grad_ssim);
}

// ---- optical-flow warping (flow_warping.py) -------------------------------------------------------------
static int flow_warp_impl(xpt_ctx* ctx, const xpt_frames* frames, const float* const flow_ms[], float* const warped_ms[],
                          float* const mask_ms[], const float* const grad_warped_ms[], float* const d_flow_ms[],
                          float* d_source, void* stream) {
  const bool bwd = grad_warped_ms != nullptr;
  if (!ctx) return fail(XPT_BAD_ARGUMENT, "xpt_flow_warp: ctx is NULL");
  if (!frames || !frames->source) return fail(XPT_BAD_ARGUMENT, "frames.source is NULL");
  const long long hw3 = (long long)ctx->H * ctx->W * 3;
  if (frames->source_frame_stride < hw3) return fail(XPT_BAD_SHAPE, "source_frame_stride %lld < H*W*3", (long long)frames->source_frame_stride);
  if (ctx->B > 1 && frames->source_batch_stride < hw3) return fail(XPT_BAD_SHAPE, "source_batch_stride too small");
  XPT_TRY(check_list(ctx, (const void* const*)flow_ms, "flow_ms", true));
  if (bwd) {
    XPT_TRY(check_list(ctx, (const void* const*)grad_warped_ms, "grad_warped_ms", true));
    if (!d_flow_ms && !d_source) return fail(XPT_BAD_ARGUMENT, "xpt_flow_warp_backward: no gradient output requested");
  } else {
    XPT_TRY(check_list(ctx, (const void* const*)warped_ms, "warped_ms", true));
  }
  cudaStream_t st = (cudaStream_t)stream;
  XPT_CUDA(cudaSetDevice(ctx->cfg.device));
  ctx->launches = 0;
  xpt_frames f = *frames;
  f.target = nullptr;                       // the flow warp reads the source frames only
  XPT_TRY(launch_pyramids(ctx, &f, nullptr, true, st));
  FlowWarpArgs a;
  memset(&a, 0, sizeof(a));
  a.lt = make_levels(ctx, &f, nullptr);
  a.B = ctx->B; a.N = ctx->N;
  if (bwd) XPT_TRY(prepare_dsource(ctx, d_source, a.d_src, a.d_src_bs, a.d_src_fs, st));
  unsigned gx = 1;
  for (int l = 0; l < ctx->S; ++l) {
    a.flow[l] = flow_ms[l];
    if (bwd) { a.gwarped[l] = grad_warped_ms[l]; a.d_flow[l] = d_flow_ms ? d_flow_ms[l] : nullptr; }
    else { a.warped[l] = warped_ms[l]; a.mask[l] = mask_ms ? mask_ms[l] : nullptr; }
    const unsigned g = (unsigned)cdiv((long long)ctx->B * ctx->N * lvl_pix(ctx, l), 256);
    if (g > gx) gx = g;
  }
  dim3 grid(gx, ctx->S);
  if (bwd) k_flow_warp<true><<<grid, 256, 0, st>>>(a);
  else k_flow_warp<false><<<grid, 256, 0, st>>>(a);
  XPT_LAUNCH_CHECK("k_flow_warp");
  if (bwd) XPT_TRY(finish_dsource(ctx, d_source, st));
  return XPT_OK;
}

int xpt_flow_warp(xpt_ctx* ctx, const xpt_frames* frames, const float* const flow_ms[], float* const warped_ms[],
                  float* const mask_ms[], void* stream) {
  return flow_warp_impl(ctx, frames, flow_ms, warped_ms, mask_ms, nullptr, nullptr, nullptr, stream);
}

int xpt_flow_warp_backward(xpt_ctx* ctx, const xpt_frames* frames, const float* const flow_ms[],
                           const float* const grad_warped_ms[], float* const d_flow_ms[], float* d_source, void* stream) {
  if (!grad_warped_ms) return fail(XPT_BAD_ARGUMENT, "grad_warped_ms is NULL");
  return flow_warp_impl(ctx, frames, flow_ms, nullptr, nullptr, grad_warped_ms, d_flow_ms, d_source, stream);
}

int xpt_l2_regularizer(xpt_ctx* ctx, const float* const weights[], const int64_t counts[], int num, float* loss,
                       const float* grad_loss, float* const d_weights[], void* stream) {
  if (!ctx || !loss || (num > 0 && (!weights || !counts))) return fail(XPT_BAD_ARGUMENT, "xpt_l2_regularizer: NULL argument");
  if (d_weights && !grad_loss) return fail(XPT_BAD_ARGUMENT, "xpt_l2_regularizer: d_weights needs grad_loss");
  cudaStream_t st = (cudaStream_t)stream;
  XPT_CUDA(cudaSetDevice(ctx->cfg.device));
  ctx->launches = 0;
  constexpr int kBlocks = 296;
  XPT_TRY(dev_alloc(ctx, &ctx->l2_part, 2 * kBlocks));
  XPT_CUDA(cudaMemsetAsync(loss, 0, sizeof(float), st));
  for (int i = 0; i < num; ++i) {
    if (!weights[i] || counts[i] < 0) return fail(XPT_BAD_ARGUMENT, "weights[%d] is NULL or has a negative count", i);
    if (counts[i] == 0) continue;
    const int blocks = (int)(counts[i] < (int64_t)kBlocks * 256 ? cdiv(counts[i], 256) : kBlocks);
    k_l2_partial<<<blocks, 256, 0, st>>>(weights[i], counts[i], reinterpret_cast<double*>(ctx->l2_part));
    XPT_LAUNCH_CHECK("k_l2_partial");
    k_l2_finish<<<1, 32, 0, st>>>(reinterpret_cast<const double*>(ctx->l2_part), blocks, loss, 1);
    XPT_LAUNCH_CHECK("k_l2_finish");
    if (d_weights && d_weights[i]) {
      k_scale_by<<<blocks, 256, 0, st>>>(weights[i], counts[i], grad_loss, d_weights[i]);
      XPT_LAUNCH_CHECK("k_scale_by");
    }
  }
  return XPT_OK;
}

int xpt_smoothness_loss(xpt_ctx* ctx, const float* const disp_ms[], const float* const target_ms[], float* loss_batch,
                        const float* grad_loss_batch, float* const d_disp_ms[], void* stream) {
  if (!ctx || !loss_batch) return fail(XPT_BAD_ARGUMENT, "xpt_smoothness_loss: NULL argument");
  XPT_TRY(check_list(ctx, (const void* const*)disp_ms, "disp_ms", true));
  XPT_TRY(check_list(ctx, (const void* const*)target_ms, "target_ms", true));
  if (d_disp_ms) XPT_TRY(check_list(ctx, (const void* const*)d_disp_ms, "d_disp_ms", true));
  cudaStream_t st = (cudaStream_t)stream;
  XPT_CUDA(cudaSetDevice(ctx->cfg.device));
  ctx->launches = 0;
  LevelTable lt = make_levels(ctx, nullptr, target_ms);
  XPT_TRY(launch_smooth(ctx, disp_ms, lt, grad_loss_batch, 1.f, d_disp_ms, 0, st));
  XPT_TRY(dev_alloc(ctx, &ctx->st_loss_batch, (size_t)3 * ctx->B));
  XPT_TRY(launch_loss_epilogue(ctx, ctx->first_sm_chunk[ctx->S], 0.f, 0.f, 0.f, nullptr, ctx->st_loss_batch, st));
  XPT_CUDA(cudaMemcpyAsync(loss_batch, ctx->st_loss_batch + (size_t)2 * ctx->B, ctx->B * sizeof(float),
                           cudaMemcpyDeviceToDevice, st));
  return XPT_OK;
}

static int total_loss_body(xpt_ctx* ctx, const xpt_frames* frames, const float* const depth_ms[],
                           const float* const disp_ms[], const float* pose, const xpt_loss_outputs* out,
                           void* stream) {
  if (!ctx || !pose || !out) return fail(XPT_BAD_ARGUMENT, "xpt_total_loss: NULL argument");
  if (!out->losses) return fail(XPT_BAD_ARGUMENT, "out->losses is NULL");
  XPT_TRY(check_frames(ctx, frames, true));
  XPT_TRY(check_list(ctx, (const void* const*)depth_ms, "depth_ms", true));
  const xpt_config& c = ctx->cfg;
  const bool do_smooth = c.w_smooth != 0.f;
  // disp_ms == NULL with a smoothness weight: the disparity is safe_reciprocal_number(depth_ms), evaluated in the
  // fused kernel (utils/util_funcs.py:146-160 + model_wrappers.py:47-48); its gradient lands in d_depth_ms
  const bool derive_disp = do_smooth && disp_ms == nullptr;
  if (do_smooth && !derive_disp) XPT_TRY(check_list(ctx, (const void* const*)disp_ms, "disp_ms", true));
  if (derive_disp && (c.flags & XPT_FLAG_UNFUSED))
    return fail(XPT_BAD_ARGUMENT, "disp_ms == NULL (disparity derived from depth) needs the fused path");
  if ((c.flags & XPT_FLAG_DEPTH_LOGIT) && (c.flags & XPT_FLAG_UNFUSED))
    return fail(XPT_BAD_ARGUMENT, "XPT_FLAG_DEPTH_LOGIT (depth activation in the kernel) needs the fused path");
  cudaStream_t st = (cudaStream_t)stream;
  XPT_CUDA(cudaSetDevice(c.device));
  ctx->launches = 0;

  bool grad = out->d_pose || out->d_source;
  for (int l = 0; l < ctx->S; ++l) grad = grad || out->d_depth_ms[l] || out->d_disp_ms[l];
  const float gs = out->grad_scale;
  const float inv_gb = 1.0f / (float)c.global_batch;

  const bool fused = !(c.flags & XPT_FLAG_UNFUSED);
  const int rc_pyr = launch_pyramids(ctx, frames, out->target_ms, true, st, pose, fused);   // + camera geometry in the same launch
  if (rc_pyr != XPT_OK) return rc_pyr;
  LevelTable lt = make_levels(ctx, frames, nullptr);

  PhotoArgs a;
  memset(&a, 0, sizeof(a));
  a.lt = lt;
  fill_photo_norms(ctx, a);
  a.geoK = ctx->geoK; a.geoT = ctx->geoT;
  a.l1_kind = c.w_l1 != 0.f ? 1 : 0;
  a.do_ssim = c.w_ssim != 0.f;
  a.do_smooth = do_smooth;
  a.gcoef_l1 = c.w_l1 * inv_gb * gs;
  a.gcoef_ssim = c.w_ssim * inv_gb * gs;
  a.gcoef_smooth = c.w_smooth * inv_gb * gs;
  const bool fused_path = !(c.flags & XPT_FLAG_UNFUSED);
  float* d_src[kMaxScales]; long long dbs[kMaxScales], dfs[kMaxScales];
  XPT_TRY(prepare_dsource(ctx, (grad && !fused_path) ? out->d_source : nullptr, d_src, dbs, dfs, st));
  for (int l = 0; l < ctx->S; ++l) {
    a.depth[l] = depth_ms[l];
    a.disp[l] = (do_smooth && !derive_disp) ? disp_ms[l] : nullptr;
    a.synth_out[l] = out->synth_ms[l]; a.mask_out[l] = out->mask_ms[l];
    a.d_depth[l] = out->d_depth_ms[l]; a.d_disp[l] = out->d_disp_ms[l];
    a.d_src[l] = d_src[l]; a.d_src_bs[l] = dbs[l]; a.d_src_fs[l] = dfs[l];
  }
  const int tiles = ctx->first_tile[ctx->S];
  if (!do_smooth || derive_disp)      // no smoothness term / no disparity tensor: that gradient is identically zero
    for (int l = 0; l < ctx->S; ++l)
      if (out->d_disp_ms[l])
        XPT_CUDA(cudaMemsetAsync(out->d_disp_ms[l], 0, (size_t)ctx->B * lvl_pix(ctx, l) * sizeof(float), st));

  if (!(c.flags & XPT_FLAG_UNFUSED)) {
    // ---- fused path: warp + L1 + SSIM + smoothness (+ all gradients) in one kernel
    FusedArgs fa;
    memset(&fa, 0, sizeof(fa));
    fa.lt = lt;
    for (int l = 0; l < ctx->S; ++l) {
      fa.lt.lv[l].tiles_x = ctx->ftiles_x[l]; fa.lt.lv[l].tiles_y = ctx->ftiles_y[l];
      fa.lt.lv[l].slot_base = ctx->ffirst_tile[l];
      fa.first_tile[l] = ctx->ffirst_tile[l];
      fa.depth[l] = a.depth[l]; fa.disp[l] = a.disp[l];
      fa.src4[l] = reinterpret_cast<const float4*>(ctx->src4_pyr[l]);
      fa.norm_photo[l] = a.norm_photo[l]; fa.norm_sm_x[l] = a.norm_sm_x[l]; fa.norm_sm_y[l] = a.norm_sm_y[l];
      fa.synth_out[l] = a.synth_out[l]; fa.mask_out[l] = a.mask_out[l];
      fa.d_depth[l] = a.d_depth[l]; fa.d_disp[l] = a.d_disp[l];
    }
    const bool dsrc = grad && out->d_source;
    bool want_out0 = false;
    for (int l = 0; l < ctx->S; ++l) want_out0 = want_out0 || out->synth_ms[l] || out->mask_ms[l];
    // training step (gradients, no synthesis tensors, no dL/dsource, N <= 4): the streaming strip kernel
    const bool use_strip = grad && !dsrc && !want_out0 && ctx->N <= 4 && (c.flags & XPT_FLAG_STRIP);
    if (use_strip) {
      XPT_TRY(prepare_strip(ctx));
      StripArgs sa;
      memset(&sa, 0, sizeof(sa));
      sa.lt = lt;
      sa.B = ctx->B; sa.N = ctx->N; sa.S = ctx->S;
      sa.pieces = ctx->strip_pieces; sa.ctas = ctx->strip_ctas;
      sa.geoK = ctx->geoK; sa.geoT = ctx->geoT;
      for (int l = 0; l < ctx->S; ++l) {
        sa.depth[l] = a.depth[l]; sa.disp[l] = a.disp[l];
        sa.src4[l] = reinterpret_cast<const float4*>(ctx->src4_pyr[l]);
        sa.norm_photo[l] = a.norm_photo[l]; sa.norm_sm_x[l] = a.norm_sm_x[l]; sa.norm_sm_y[l] = a.norm_sm_y[l];
        sa.d_depth[l] = a.d_depth[l]; sa.d_disp[l] = a.d_disp[l];
      }
      sa.do_l1 = a.l1_kind != 0; sa.do_ssim = a.do_ssim; sa.do_smooth = a.do_smooth;
      sa.logit = (c.flags & XPT_FLAG_DEPTH_LOGIT) ? 1 : 0;
      sa.grad_factor = a.grad_factor;
      sa.gcoef_l1 = a.gcoef_l1; sa.gcoef_ssim = a.gcoef_ssim; sa.gcoef_smooth = a.gcoef_smooth;
      sa.loss_part = ctx->strip_loss_part; sa.slots_per_b = ctx->strip_slots; sa.pose_part = ctx->strip_pose_part;
      XPT_TRY(launch_strip(ctx, sa, derive_disp, st));
      EpilogueArgs ea;
      memset(&ea, 0, sizeof(ea));
      ea.pose_part = ctx->strip_pose_part; ea.loss_part = ctx->strip_loss_part;
      ea.slots_per_b = ctx->strip_slots; ea.slots_used = ctx->strip_slots; ea.B = ctx->B; ea.N = ctx->N;
      ea.pose = pose; ea.d_pose = out->d_pose;
      ea.inv_global_batch = inv_gb; ea.w0 = c.w_l1; ea.w1 = c.w_ssim; ea.w2 = c.w_smooth;
      ea.losses = out->losses; ea.loss_batch = out->loss_batch;
      ea.loss_sum_b = ctx->loss_sum_b; ea.ticket = ctx->ticket;
      set_exchange(ctx, ea);
      k_epilogue<<<(ea.d_pose ? ctx->B * ctx->N : 0) + ctx->B, 128, 0, st>>>(ea);
      XPT_LAUNCH_CHECK("k_epilogue");
      return XPT_OK;
    }
    if (dsrc) XPT_TRY(prepare_dsource4(ctx, fa.d_src4, st));
    const int ftiles = ctx->ffirst_tile[ctx->S];
    fa.first_tile[ctx->S] = ftiles;
    fa.tiles_per_b = ftiles;
    fa.B = ctx->B; fa.N = ctx->N;
    fa.do_l1 = a.l1_kind != 0; fa.do_ssim = a.do_ssim; fa.do_smooth = a.do_smooth;
    fa.grad_factor = a.grad_factor;
    fa.gcoef_l1 = a.gcoef_l1; fa.gcoef_ssim = a.gcoef_ssim; fa.gcoef_smooth = a.gcoef_smooth;
    fa.loss_part = ctx->loss_part; fa.slots_per_b = ctx->slots_per_b; fa.pose_part = ctx->pose_part;
    bool want_out = false;
    for (int l = 0; l < ctx->S; ++l) want_out = want_out || out->synth_ms[l] || out->mask_ms[l];
    // template dispatch: (GRAD, OUT, DSRC, DERIVE mode); the logit mode is compiled for training steps only
    const bool logit = (c.flags & XPT_FLAG_DEPTH_LOGIT) != 0;
    if (logit && do_smooth && !derive_disp)
      return fail(XPT_BAD_ARGUMENT, "XPT_FLAG_DEPTH_LOGIT: pass disp_ms == NULL (the disparity is formed from the activated depth)");
    const int sel = (grad ? 8 : 0) | (want_out ? 4 : 0) | (dsrc ? 2 : 0) | (derive_disp ? 1 : 0);
    if (logit) {
      switch ((grad ? 4 : 0) | (want_out ? 2 : 0) | (dsrc ? 1 : 0)) {
        case 0: XPT_TRY((launch_fused<false, false, false, 2>(ctx, fa, st))); break;
        case 2: XPT_TRY((launch_fused<false, true, false, 2>(ctx, fa, st))); break;
        case 4: XPT_TRY((launch_fused<true, false, false, 2>(ctx, fa, st))); break;
        case 5: XPT_TRY((launch_fused<true, false, true, 2>(ctx, fa, st))); break;
        case 6: XPT_TRY((launch_fused<true, true, false, 2>(ctx, fa, st))); break;
        case 7: XPT_TRY((launch_fused<true, true, true, 2>(ctx, fa, st))); break;
        default: return fail(XPT_BAD_ARGUMENT, "dL/dsource without other gradients is not served");
      }
    } else switch (sel) {
      case 0: XPT_TRY((launch_fused<false, false, false, 0>(ctx, fa, st))); break;
      case 1: XPT_TRY((launch_fused<false, false, false, 1>(ctx, fa, st))); break;
      case 4: XPT_TRY((launch_fused<false, true, false, 0>(ctx, fa, st))); break;
      case 5: XPT_TRY((launch_fused<false, true, false, 1>(ctx, fa, st))); break;
      case 8: XPT_TRY((launch_fused<true, false, false, 0>(ctx, fa, st))); break;
      case 9: XPT_TRY((launch_fused<true, false, false, 1>(ctx, fa, st))); break;
      case 10: XPT_TRY((launch_fused<true, false, true, 0>(ctx, fa, st))); break;
      case 11: XPT_TRY((launch_fused<true, false, true, 1>(ctx, fa, st))); break;
      case 12: XPT_TRY((launch_fused<true, true, false, 0>(ctx, fa, st))); break;
      case 13: XPT_TRY((launch_fused<true, true, false, 1>(ctx, fa, st))); break;
      case 14: XPT_TRY((launch_fused<true, true, true, 0>(ctx, fa, st))); break;
      default: XPT_TRY((launch_fused<true, true, true, 1>(ctx, fa, st))); break;
    }
    EpilogueArgs ea;
    memset(&ea, 0, sizeof(ea));
    ea.pose_part = ctx->pose_part; ea.loss_part = ctx->loss_part;
    ea.slots_per_b = ctx->slots_per_b; ea.slots_used = ftiles; ea.B = ctx->B; ea.N = ctx->N;
    ea.pose = pose; ea.d_pose = grad ? out->d_pose : nullptr;
    ea.inv_global_batch = inv_gb; ea.w0 = c.w_l1; ea.w1 = c.w_ssim; ea.w2 = c.w_smooth;
    ea.losses = out->losses; ea.loss_batch = out->loss_batch;
    ea.loss_sum_b = ctx->loss_sum_b; ea.ticket = ctx->ticket;
    set_exchange(ctx, ea);
    k_epilogue<<<(ea.d_pose ? ctx->B * ctx->N : 0) + ctx->B, 128, 0, st>>>(ea);
    XPT_LAUNCH_CHECK("k_epilogue");
    if (dsrc) XPT_TRY(finish_dsource4(ctx, out->d_source, st));
  } else {
    // ---- unfused path (flags bit 0): one kernel per reference stage, tensors through HBM
    float* synth[kMaxScales]; float* gsyn[kMaxScales];
    for (int l = 0; l < ctx->S; ++l) {
      size_t n = (size_t)ctx->B * ctx->N * lvl_pix(ctx, l) * 3;
      if (out->synth_ms[l]) synth[l] = out->synth_ms[l];
      else { XPT_TRY(dev_alloc(ctx, &ctx->synth_scr[l], n)); synth[l] = ctx->synth_scr[l]; }
      gsyn[l] = nullptr;
      if (grad) { XPT_TRY(dev_alloc(ctx, &ctx->gsynth_scr[l], n)); gsyn[l] = ctx->gsynth_scr[l]; }
      a.synth[l] = synth[l]; a.gsynth[l] = gsyn[l];
    }
    XPT_TRY(launch_warp_fwd(ctx, lt, depth_ms, synth, out->mask_ms, st));
    a.do_smooth = 0;
    if (grad) XPT_TRY((launch_photo<false, true>(ctx, a, st)));
    else XPT_TRY((launch_photo<false, false>(ctx, a, st)));
    int used = tiles;
    if (do_smooth) {
      // k_smooth applies gbatch * norm; fold dTotal/dloss in through a per-snippet vector
      float* gvec = nullptr;
      if (grad) {
        XPT_TRY(dev_alloc(ctx, &ctx->st_dpose, (size_t)ctx->B * ctx->N * 6 + ctx->B));
        gvec = ctx->st_dpose + (size_t)ctx->B * ctx->N * 6;
        k_fill<<<cdiv(ctx->B, 128), 128, 0, st>>>(gvec, ctx->B, a.gcoef_smooth);
        XPT_LAUNCH_CHECK("k_fill");
      }
      XPT_TRY(launch_smooth(ctx, disp_ms, lt, gvec, 1.f, grad ? out->d_disp_ms : nullptr, tiles, st));
      used = tiles + ctx->first_sm_chunk[ctx->S];
    }
    XPT_TRY(launch_loss_epilogue(ctx, used, c.w_l1, c.w_ssim, c.w_smooth, out->losses, out->loss_batch, st));
    if (grad) {
      // reuse the standalone warp backward; d_source was prepared above, so pass NULL and finish here
      WarpBwdArgs w;
      memset(&w, 0, sizeof(w));
      w.lt = lt; w.B = ctx->B; w.N = ctx->N; w.geoK = ctx->geoK; w.geoT = ctx->geoT;
      w.pose_part = ctx->pose_part; w.slots_per_b = ctx->slots_per_b;
      int maxchunks = 0;
      for (int l = 0; l < ctx->S; ++l) {
        w.depth[l] = depth_ms[l]; w.gsynth[l] = gsyn[l]; w.d_depth[l] = out->d_depth_ms[l];
        w.d_src[l] = d_src[l]; w.d_src_bs[l] = dbs[l]; w.d_src_fs[l] = dfs[l];
        w.chunk_base[l] = ctx->first_chunk[l];
        if (ctx->chunks[l] > maxchunks) maxchunks = ctx->chunks[l];
      }
      dim3 grid(maxchunks, ctx->B, ctx->S);
      k_warp_bwd<<<grid, kWarpBwdThreads, 0, st>>>(w);
      XPT_LAUNCH_CHECK("k_warp_bwd");
      if (out->d_pose) {
        k_pose_epilogue<<<ctx->B * ctx->N, 128, 0, st>>>(ctx->pose_part, ctx->slots_per_b, ctx->first_chunk[ctx->S],
                                                         pose, out->d_pose, ctx->N, 1.0f);
        XPT_LAUNCH_CHECK("k_pose_epilogue");
      }
    }
  }
  if (grad && !fused_path) XPT_TRY(finish_dsource(ctx, out->d_source, st));
  return XPT_OK;
}

static int total_loss_impl(xpt_ctx* ctx, const xpt_frames* frames, const float* const depth_ms[],
                           const float* const disp_ms[], const float* pose, const xpt_loss_outputs* out, void* stream) {
  if (ctx) ctx->exchanged = false;
  int rc = total_loss_body(ctx, frames, depth_ms, disp_ms, pose, out, stream);
  if (rc == XPT_OK && ctx && (ctx->cfg.flags & XPT_FLAG_ALLREDUCE) && !ctx->host_call && !ctx->exchanged) {
    // (the fused path sums the losses inside k_epilogue through peer memory when the inboxes are mapped: no launch here)
    // reference distributer.py:93-110: the replicas' loss scalars are summed; here on the step's own stream, so the
    // collective is one more node of the step's CUDA graph
    float* bufs[1] = {out->losses};
    const int64_t counts[1] = {4};
    rc = allreduce_on(ctx, bufs, counts, 1, (cudaStream_t)stream);
  }
  return rc;
}

int xpt_total_loss(xpt_ctx* ctx, const xpt_frames* frames, const float* const depth_ms[],
                   const float* const disp_ms[], const float* pose, const xpt_loss_outputs* out, void* stream) {
  if (!ctx || !frames || !pose || !out || !depth_ms) return fail(XPT_BAD_ARGUMENT, "xpt_total_loss: NULL argument");
  cudaStream_t st = (cudaStream_t)stream;
  if (!(ctx->cfg.flags & XPT_FLAG_GRAPH)) return total_loss_impl(ctx, frames, depth_ms, disp_ms, pose, out, stream);
  XPT_CUDA(cudaSetDevice(ctx->cfg.device));
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  XPT_CUDA(cudaStreamIsCapturing(st, &cs));
  const bool legacy = st == nullptr || st == cudaStreamLegacy;       // the NULL stream cannot be captured
  if (cs != cudaStreamCaptureStatusNone || !ctx->warm || legacy) {
    // already inside someone else's capture, or first call (lazy scratch is allocated eagerly once)
    int rc = total_loss_impl(ctx, frames, depth_ms, disp_ms, pose, out, stream);
    if (rc == XPT_OK && cs == cudaStreamCaptureStatusNone) ctx->warm = true;
    return rc;
  }
  // key = every argument that ends up in a kernel parameter
  std::vector<uint64_t> key;
  key.reserve(16 + 7 * kMaxScales);
  auto K = [&](const void* p) { key.push_back((uint64_t)(uintptr_t)p); };
  K(frames->source); key.push_back((uint64_t)frames->source_batch_stride); key.push_back((uint64_t)frames->source_frame_stride);
  K(frames->target); key.push_back((uint64_t)frames->target_batch_stride); K(frames->intrinsic);
  K(pose); K(stream); K(out->losses); K(out->loss_batch); K(out->d_pose); K(out->d_source);
  key.push_back(ctx->host_call ? 1u : 0u);
  uint32_t gsbits; memcpy(&gsbits, &out->grad_scale, 4); key.push_back(gsbits);
  for (int l = 0; l < ctx->S; ++l) {
    K(depth_ms[l]); K(disp_ms ? disp_ms[l] : nullptr); K(out->synth_ms[l]); K(out->mask_ms[l]);
    K(out->target_ms[l]); K(out->d_depth_ms[l]); K(out->d_disp_ms[l]);
  }
  if (!ctx->graphs) ctx->graphs = new std::vector<xpt_ctx::GraphEntry>();
  for (auto& e : *ctx->graphs)
    if (e.key == key) {
      XPT_CUDA(cudaGraphLaunch(e.exec, st));
      ctx->launches = e.launches;
      return XPT_OK;
    }
  XPT_CUDA(cudaStreamBeginCapture(st, cudaStreamCaptureModeRelaxed));
  int rc = total_loss_impl(ctx, frames, depth_ms, disp_ms, pose, out, stream);
  cudaGraph_t graph = nullptr;
  cudaError_t ce = cudaStreamEndCapture(st, &graph);
  if (rc != XPT_OK) { if (graph) cudaGraphDestroy(graph); return rc; }
  if (ce != cudaSuccess) return fail(XPT_CUDA_ERROR, "cudaStreamEndCapture failed: %s", cudaGetErrorString(ce));
  cudaGraphExec_t exec = nullptr;
  ce = cudaGraphInstantiate(&exec, graph, 0);
  cudaGraphDestroy(graph);
  if (ce != cudaSuccess) return fail(XPT_CUDA_ERROR, "cudaGraphInstantiate failed: %s", cudaGetErrorString(ce));
  if (ctx->graphs->size() >= 64) {           // bounded cache: drop the oldest
    cudaGraphExecDestroy(ctx->graphs->front().exec);
    ctx->graphs->erase(ctx->graphs->begin());
  }
  ctx->graphs->push_back({key, exec, ctx->launches});
  XPT_CUDA(cudaGraphLaunch(exec, st));
  return XPT_OK;
}

int xpt_profile_select(xpt_ctx* ctx, int kernel) {
  if (!ctx || (kernel != XPT_PROFILE_FUSED && kernel != XPT_PROFILE_PYRAMID))
    return fail(XPT_BAD_ARGUMENT, "xpt_profile_select: bad argument");
  ctx->prof_kind = kernel;
  return XPT_OK;
}

int xpt_profile_begin(xpt_ctx* ctx, int max_records) {
  if (!ctx || max_records < 0) return fail(XPT_BAD_ARGUMENT, "xpt_profile_begin: bad argument");
  XPT_CUDA(cudaSetDevice(ctx->cfg.device));
  if (!ctx->prof_events) ctx->prof_events = new std::vector<cudaEvent_t>();
  while ((int)ctx->prof_events->size() < 2 * max_records) {
    cudaEvent_t e;
    XPT_CUDA(cudaEventCreate(&e));
    ctx->prof_events->push_back(e);
  }
  ctx->prof_count = 0;
  ctx->prof_on = max_records;
  return XPT_OK;
}

int xpt_profile_end(xpt_ctx* ctx, float* ms_out, int capacity) {
  if (!ctx) return fail(XPT_BAD_ARGUMENT, "xpt_profile_end: NULL ctx");
  int n = ctx->prof_count;
  ctx->prof_on = 0;
  if (n > capacity) n = capacity;
  for (int i = 0; i < n; ++i) {
    XPT_CUDA(cudaEventSynchronize((*ctx->prof_events)[2 * i + 1]));
    XPT_CUDA(cudaEventElapsedTime(ms_out + i, (*ctx->prof_events)[2 * i], (*ctx->prof_events)[2 * i + 1]));
  }
  return n;
}

// Enqueues one whole host-buffer call (copies in, chunked compute, copies out) on `stream` and the ctx's copy
// streams; every side stream is forked from and joined back into `stream`, so the sequence can be captured
// into ONE CUDA graph.  Returns the number of chunks in *nc_out.  No synchronisation in here.
// XPT_HOST_TRACE=1: per-stage device timestamps of the (eager) host-buffer pipeline, printed to stderr
struct HostTrace { bool on; cudaEvent_t t0, small_in, in[kMaxChunks], done[kMaxChunks], out_end; };
static HostTrace g_trace = {};
static bool host_trace_on() {
  static int state = -1;
  if (state < 0) {
    const char* e = getenv("XPT_HOST_TRACE");
    state = (e && e[0] == '1') ? 1 : 0;
    if (state) {
      cudaEventCreate(&g_trace.t0); cudaEventCreate(&g_trace.small_in); cudaEventCreate(&g_trace.out_end);
      for (int k = 0; k < kMaxChunks; ++k) { cudaEventCreate(&g_trace.in[k]); cudaEventCreate(&g_trace.done[k]); }
    }
  }
  return state == 1;
}

static int host_enqueue(xpt_ctx* ctx, const xpt_frames* frames, const float* const depth_ms[],
                        const float* const disp_ms[], const float* pose, const xpt_loss_outputs* out, void* stream,
                        int* nc_out) {
  if (!ctx || !pose || !out) return fail(XPT_BAD_ARGUMENT, "xpt_total_loss_host: NULL argument");
  if (!out->losses) return fail(XPT_BAD_ARGUMENT, "out->losses is NULL");
  XPT_TRY(check_frames(ctx, frames, true));
  XPT_TRY(check_list(ctx, (const void* const*)depth_ms, "depth_ms", true));
  const bool do_smooth = ctx->cfg.w_smooth != 0.f && disp_ms != nullptr;     // NULL: derived from depth in the kernel
  if (do_smooth) XPT_TRY(check_list(ctx, (const void* const*)disp_ms, "disp_ms", true));
  const long long hw3 = (long long)ctx->H * ctx->W * 3;
  if (frames->source_frame_stride != hw3)
    return fail(XPT_BAD_SHAPE, "host entry point needs dense frames (source_frame_stride == H*W*3)");
  cudaStream_t st = (cudaStream_t)stream;
  XPT_CUDA(cudaSetDevice(ctx->cfg.device));
  const int B = ctx->B, N = ctx->N, S = ctx->S;
  const size_t fb = sizeof(float);

  // ---- pipeline: the batch is cut into chunks; chunk k+1's host->device copies run on a copy-in
  // stream while chunk k computes on the caller's stream and chunk k-1's results drain on a copy-out
  // stream (snippets are independent; every chunk normalises by the global batch, so chunk losses add).
  // The link is the bottleneck (the frames are ~70 B per pixel), so what is exposed is the LAST chunk's compute
  // and drain: use the deepest pipeline whose chunks still hold >= 64 Ki pixels (a smaller chunk no longer fills the GPU: its compute takes as long as its copy and the pipeline turns compute-bound).  Chunk arguments are stable
  // (ctx-owned staging), so the child replays one captured CUDA graph per chunk.
  int nc = 1;
  if (!(ctx->cfg.flags & XPT_FLAG_NO_PIPELINE))
    for (int c = kMaxChunks; c >= 2; --c)
      if (B % c == 0 && (long long)(B / c) * ctx->H * ctx->W >= 65536) { nc = c; break; }
  const int Bc = B / nc;
  xpt_ctx* child = ctx;
  if (nc > 1) {
    if (!ctx->child) {
      xpt_config cc = ctx->cfg;
      cc.batch = Bc;
      cc.flags |= XPT_FLAG_NO_PIPELINE | XPT_FLAG_GRAPH;
      XPT_TRY(xpt_create(&ctx->child, &cc));
    }
    child = ctx->child;
    if (!ctx->s_in) {
      XPT_CUDA(cudaStreamCreateWithFlags(&ctx->s_in, cudaStreamNonBlocking));
      XPT_CUDA(cudaStreamCreateWithFlags(&ctx->s_out, cudaStreamNonBlocking));
      XPT_CUDA(cudaStreamCreateWithFlags(&ctx->s_in2, cudaStreamNonBlocking));
      XPT_CUDA(cudaEventCreateWithFlags(&ctx->ev_small, cudaEventDisableTiming));
      XPT_CUDA(cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming));
      XPT_CUDA(cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming));
      for (int k = 0; k < kMaxChunks; ++k) {
        XPT_CUDA(cudaEventCreateWithFlags(&ctx->ev_in[k], cudaEventDisableTiming));
        XPT_CUDA(cudaEventCreateWithFlags(&ctx->ev_done[k], cudaEventDisableTiming));
      }
    }
  }
  cudaStream_t s_in = nc > 1 ? ctx->s_in : st, s_out = nc > 1 ? ctx->s_out : st;
  cudaStream_t s_in2 = s_in;          // small inputs first, then the frame chunks: one FIFO keeps the link saturated

  // ---- device staging (sized for the whole batch, owned by the parent ctx) ---------------------------
  const long long snip = (long long)(N + 1) * hw3;
  XPT_TRY(dev_alloc(ctx, &ctx->st_frames, (size_t)B * snip));
  XPT_TRY(dev_alloc(ctx, &ctx->st_K, (size_t)B * 9));
  XPT_TRY(dev_alloc(ctx, &ctx->st_pose, (size_t)B * N * 6));
  XPT_TRY(dev_alloc(ctx, &ctx->st_losses, 4 * kMaxChunks));
  if (out->loss_batch) XPT_TRY(dev_alloc(ctx, &ctx->st_loss_batch, (size_t)3 * B));
  if (out->d_pose) XPT_TRY(dev_alloc(ctx, &ctx->st_dpose, (size_t)B * N * 6 + B));
  if (out->d_source) XPT_TRY(dev_alloc(ctx, &ctx->st_dsource, (size_t)B * N * hw3));
  for (int l = 0; l < S; ++l) {
    const size_t n = (size_t)B * lvl_pix(ctx, l);
    XPT_TRY(dev_alloc(ctx, &ctx->st_depth[l], n));
    if (do_smooth) XPT_TRY(dev_alloc(ctx, &ctx->st_disp[l], n));
    if (out->d_depth_ms[l]) XPT_TRY(dev_alloc(ctx, &ctx->st_ddepth[l], n));
    if (out->d_disp_ms[l]) XPT_TRY(dev_alloc(ctx, &ctx->st_ddisp[l], n));
    if (out->synth_ms[l]) XPT_TRY(dev_alloc(ctx, &ctx->st_synth[l], n * N * 3));
    if (out->mask_ms[l]) XPT_TRY(dev_alloc(ctx, &ctx->st_mask[l], n * N));
    if (out->target_ms[l]) XPT_TRY(dev_alloc(ctx, &ctx->st_target[l], n * 3));
  }
  const bool one_block = frames->target == frames->source + (long long)N * hw3 &&
                         frames->source_batch_stride == snip && frames->target_batch_stride == snip;
  if (!ctx->h_losses) XPT_CUDA(cudaMallocHost(reinterpret_cast<void**>(&ctx->h_losses), 4 * kMaxChunks * sizeof(float)));

  // Small outputs whose host buffer is pinned (device-mapped) are written by the epilogue kernel THROUGH the
  // mapping -- no staging and no copy-out latency; pageable buffers take the staged device->host copies.
  auto mapped = [](float* host) -> float* {
    if (!host) return nullptr;
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, host) != cudaSuccess) { (void)cudaGetLastError(); return nullptr; }
    return at.type == cudaMemoryTypeHost ? static_cast<float*>(at.devicePointer) : nullptr;
  };
  float* m_dpose = mapped(out->d_pose);
  float* m_losses = mapped(ctx->h_losses);
  float* m_ddepth[kMaxScales]; float* m_ddisp[kMaxScales];
  // Only the tiny outputs go this way: the fused kernel's 8-byte gradient stores would become one PCIe
  // transaction each (measured: chunk compute 50 -> 150 us), so the gradient maps stay staged.
  for (int l = 0; l < S; ++l) { m_ddepth[l] = nullptr; m_ddisp[l] = nullptr; }
  const bool trace = host_trace_on() && nc > 1;
  if (trace) cudaEventRecord(g_trace.t0, st);
  if (nc > 1) {      // fork the copy streams from the caller's stream
    XPT_CUDA(cudaEventRecord(ctx->ev_fork, st));
    XPT_CUDA(cudaStreamWaitEvent(s_in, ctx->ev_fork, 0));
    XPT_CUDA(cudaStreamWaitEvent(s_out, ctx->ev_fork, 0));
  }
  // small inputs (K, pose, depth_ms, disp_ms: ~15 % of the bytes): whole batch at once, ahead of the first frame chunk
  XPT_CUDA(cudaMemcpyAsync(ctx->st_K, frames->intrinsic, (size_t)B * 9 * fb, cudaMemcpyHostToDevice, s_in2));
  XPT_CUDA(cudaMemcpyAsync(ctx->st_pose, pose, (size_t)B * N * 6 * fb, cudaMemcpyHostToDevice, s_in2));
  for (int l = 0; l < S; ++l) {
    const size_t n = (size_t)B * lvl_pix(ctx, l);
    XPT_CUDA(cudaMemcpyAsync(ctx->st_depth[l], depth_ms[l], n * fb, cudaMemcpyHostToDevice, s_in2));
    if (do_smooth) XPT_CUDA(cudaMemcpyAsync(ctx->st_disp[l], disp_ms[l], n * fb, cudaMemcpyHostToDevice, s_in2));
  }
  if (nc > 1) {
    XPT_CUDA(cudaEventRecord(ctx->ev_small, s_in2));
    XPT_CUDA(cudaStreamWaitEvent(st, ctx->ev_small, 0));
  }
  if (trace) cudaEventRecord(g_trace.small_in, s_in2);

  for (int k = 0; k < nc; ++k) {
    const int b0 = k * Bc;
    // -- host -> device, chunk k: the frames
    // ONE copy-in stream: a second one only splits the link and delays every chunk (profiles/microbench/h2d_pipe.cu)
    float* dfr = ctx->st_frames + (long long)b0 * snip;
    if (one_block) {
      XPT_CUDA(cudaMemcpyAsync(dfr, frames->source + (long long)b0 * snip, (size_t)Bc * snip * fb, cudaMemcpyHostToDevice, s_in));
    } else {
      XPT_CUDA(cudaMemcpy2DAsync(dfr, snip * fb, frames->source + (long long)b0 * frames->source_batch_stride,
                                 frames->source_batch_stride * fb, (size_t)N * hw3 * fb, Bc, cudaMemcpyHostToDevice, s_in));
      XPT_CUDA(cudaMemcpy2DAsync(dfr + (long long)N * hw3, snip * fb, frames->target + (long long)b0 * frames->target_batch_stride,
                                 frames->target_batch_stride * fb, (size_t)hw3 * fb, Bc, cudaMemcpyHostToDevice, s_in));
    }
    const float* cdepth[kMaxScales]; const float* cdisp[kMaxScales];
    for (int l = 0; l < S; ++l) {
      const size_t off = (size_t)b0 * lvl_pix(ctx, l);
      cdepth[l] = ctx->st_depth[l] + off;
      cdisp[l] = do_smooth ? ctx->st_disp[l] + off : nullptr;
    }
    if (nc > 1) {
      XPT_CUDA(cudaEventRecord(ctx->ev_in[k], s_in));
      XPT_CUDA(cudaStreamWaitEvent(st, ctx->ev_in[k], 0));
    }
    if (trace) cudaEventRecord(g_trace.in[k], s_in);
    // -- compute, chunk k (caller's stream)
    xpt_frames df;
    df.source = dfr; df.source_batch_stride = snip; df.source_frame_stride = hw3;
    df.target = dfr + (long long)N * hw3; df.target_batch_stride = snip;
    df.intrinsic = ctx->st_K + (size_t)b0 * 9;
    xpt_loss_outputs dout;
    memset(&dout, 0, sizeof(dout));
    dout.grad_scale = out->grad_scale;
    dout.losses = m_losses ? m_losses + 4 * k : ctx->st_losses + 4 * k;
    if (out->loss_batch) dout.loss_batch = ctx->st_loss_batch + (size_t)3 * b0;     // chunk-major [k][3][Bc]
    if (out->d_pose) dout.d_pose = m_dpose ? m_dpose + (size_t)b0 * N * 6 : ctx->st_dpose + (size_t)b0 * N * 6;
    if (out->d_source) dout.d_source = ctx->st_dsource + (size_t)b0 * N * hw3;
    for (int l = 0; l < S; ++l) {
      const size_t off = (size_t)b0 * lvl_pix(ctx, l);
      if (out->d_depth_ms[l]) dout.d_depth_ms[l] = m_ddepth[l] ? m_ddepth[l] + off : ctx->st_ddepth[l] + off;
      if (out->d_disp_ms[l]) dout.d_disp_ms[l] = m_ddisp[l] ? m_ddisp[l] + off : ctx->st_ddisp[l] + off;
      if (out->synth_ms[l]) dout.synth_ms[l] = ctx->st_synth[l] + off * N * 3;
      if (out->mask_ms[l]) dout.mask_ms[l] = ctx->st_mask[l] + off * N;
      if (out->target_ms[l]) dout.target_ms[l] = ctx->st_target[l] + off * 3;
    }
    child->host_call = true;          // rank-local losses: a collective per chunk would be N x nc tiny all-reduces
    const int crc = xpt_total_loss(child, &df, cdepth, do_smooth ? cdisp : nullptr, ctx->st_pose + (size_t)b0 * N * 6, &dout, stream);
    child->host_call = false;
    XPT_TRY(crc);
    if (nc > 1) {
      XPT_CUDA(cudaEventRecord(ctx->ev_done[k], st));
      XPT_CUDA(cudaStreamWaitEvent(s_out, ctx->ev_done[k], 0));
    }
    if (trace) cudaEventRecord(g_trace.done[k], st);
    // -- device -> host, chunk k: the large per-chunk outputs; small ones leave once at the end
    if (out->d_source) XPT_CUDA(cudaMemcpyAsync(out->d_source + (size_t)b0 * N * hw3, dout.d_source, (size_t)Bc * N * hw3 * fb, cudaMemcpyDeviceToHost, s_out));
    for (int l = 0; l < S; ++l) {
      const size_t off = (size_t)b0 * lvl_pix(ctx, l), n = (size_t)Bc * lvl_pix(ctx, l);
      if (l == 0 && out->d_depth_ms[l] && !m_ddepth[l]) XPT_CUDA(cudaMemcpyAsync(out->d_depth_ms[l] + off, dout.d_depth_ms[l], n * fb, cudaMemcpyDeviceToHost, s_out));
      if (l == 0 && out->d_disp_ms[l] && !m_ddisp[l]) XPT_CUDA(cudaMemcpyAsync(out->d_disp_ms[l] + off, dout.d_disp_ms[l], n * fb, cudaMemcpyDeviceToHost, s_out));
      if (out->synth_ms[l]) XPT_CUDA(cudaMemcpyAsync(out->synth_ms[l] + off * N * 3, dout.synth_ms[l], n * N * 3 * fb, cudaMemcpyDeviceToHost, s_out));
      if (out->mask_ms[l]) XPT_CUDA(cudaMemcpyAsync(out->mask_ms[l] + off * N, dout.mask_ms[l], n * N * fb, cudaMemcpyDeviceToHost, s_out));
      if (out->target_ms[l]) XPT_CUDA(cudaMemcpyAsync(out->target_ms[l] + off * 3, dout.target_ms[l], n * 3 * fb, cudaMemcpyDeviceToHost, s_out));
    }
  }
  // small outputs, whole batch (s_out already waits for the last chunk's compute)
  if (!m_losses) XPT_CUDA(cudaMemcpyAsync(ctx->h_losses, ctx->st_losses, (size_t)4 * nc * fb, cudaMemcpyDeviceToHost, s_out));
  if (out->loss_batch)
    for (int k = 0; k < nc; ++k)
      for (int r = 0; r < 3; ++r)
        XPT_CUDA(cudaMemcpyAsync(out->loss_batch + (size_t)r * B + (size_t)k * Bc, ctx->st_loss_batch + (size_t)3 * k * Bc + (size_t)r * Bc,
                                 Bc * fb, cudaMemcpyDeviceToHost, s_out));
  if (out->d_pose && !m_dpose) XPT_CUDA(cudaMemcpyAsync(out->d_pose, ctx->st_dpose, (size_t)B * N * 6 * fb, cudaMemcpyDeviceToHost, s_out));
  // levels >= 1 of the gradient maps: one kernel with wide coalesced stores through the pinned mapping instead of
  // six copies of ~7 us latency each in the tail of the call; pageable destinations take the copies
  {
    DrainArgs da;
    memset(&da, 0, sizeof(da));
    for (int l = 1; l < S; ++l) {
      const long long n = (long long)B * lvl_pix(ctx, l);
      float* md = mapped(out->d_depth_ms[l]);
      float* ms = mapped(out->d_disp_ms[l]);
      if (md && da.n < 16) { da.src[da.n] = ctx->st_ddepth[l]; da.dst[da.n] = md; da.count[da.n] = n; ++da.n; m_ddepth[l] = md; }
      if (ms && da.n < 16) { da.src[da.n] = ctx->st_ddisp[l]; da.dst[da.n] = ms; da.count[da.n] = n; ++da.n; m_ddisp[l] = ms; }
    }
    if (da.n > 0) {
      k_drain<<<dim3(32, da.n), 256, 0, s_out>>>(da);
      XPT_LAUNCH_CHECK("k_drain");
    }
  }
  for (int l = 1; l < S; ++l) {
    const size_t n = (size_t)B * lvl_pix(ctx, l);
    if (out->d_depth_ms[l] && !m_ddepth[l]) XPT_CUDA(cudaMemcpyAsync(out->d_depth_ms[l], ctx->st_ddepth[l], n * fb, cudaMemcpyDeviceToHost, s_out));
    if (out->d_disp_ms[l] && !m_ddisp[l]) XPT_CUDA(cudaMemcpyAsync(out->d_disp_ms[l], ctx->st_ddisp[l], n * fb, cudaMemcpyDeviceToHost, s_out));
  }
  if (trace) cudaEventRecord(g_trace.out_end, s_out);
  if (nc > 1) {      // join: the caller's stream completes when the last copy-out has landed
    XPT_CUDA(cudaEventRecord(ctx->ev_join, s_out));
    XPT_CUDA(cudaStreamWaitEvent(st, ctx->ev_join, 0));
  }
  *nc_out = nc;
  return XPT_OK;
}

static bool is_pinned(const void* p) {
  if (!p) return true;
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { (void)cudaGetLastError(); return false; }
  return at.type == cudaMemoryTypeHost;
}

int xpt_total_loss_host_begin(xpt_ctx* ctx, const xpt_frames* frames, const float* const depth_ms[],
                              const float* const disp_ms[], const float* pose, const xpt_loss_outputs* out, void* stream) {
  if (!ctx || !frames || !pose || !out || !depth_ms) return fail(XPT_BAD_ARGUMENT, "xpt_total_loss_host: NULL argument");
  if (ctx->host_pending) return fail(XPT_BAD_ARGUMENT, "xpt_total_loss_host_begin: the previous call on this ctx was not ended");
  cudaStream_t st = (cudaStream_t)stream;
  XPT_CUDA(cudaSetDevice(ctx->cfg.device));
  int nc = 1, rc = XPT_OK;
  // Replay path: once the ctx is warm, a call whose buffers are all pinned and whose stream can be captured is
  // recorded ONCE (copies + chunked compute + copies out, ~100 API calls) and replayed as one graph launch.
  bool can_graph = ctx->host_warm && !(ctx->cfg.flags & XPT_FLAG_NO_PIPELINE) && st != nullptr && st != cudaStreamLegacy &&
                   !host_trace_on();
  if (can_graph) {
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    XPT_CUDA(cudaStreamIsCapturing(st, &cs));
    can_graph = cs == cudaStreamCaptureStatusNone;
  }
  std::vector<uint64_t> key;
  if (can_graph) {
    auto K = [&](const void* p) { key.push_back((uint64_t)(uintptr_t)p); can_graph = can_graph && is_pinned(p); };
    K(frames->source); K(frames->target); K(frames->intrinsic); K(pose);
    key.push_back((uint64_t)frames->source_batch_stride); key.push_back((uint64_t)frames->source_frame_stride);
    key.push_back((uint64_t)frames->target_batch_stride); key.push_back((uint64_t)(uintptr_t)stream);
    K(out->losses); K(out->loss_batch); K(out->d_pose); K(out->d_source);
    uint32_t gsbits; memcpy(&gsbits, &out->grad_scale, 4); key.push_back(gsbits);
    for (int l = 0; l < ctx->S; ++l) {
      K(depth_ms[l]); K(disp_ms ? disp_ms[l] : nullptr); K(out->synth_ms[l]); K(out->mask_ms[l]);
      K(out->target_ms[l]); K(out->d_depth_ms[l]); K(out->d_disp_ms[l]);
    }
  }
  if (!can_graph) {
    rc = host_enqueue(ctx, frames, depth_ms, disp_ms, pose, out, stream, &nc);
    if (rc != XPT_OK) return rc;
    XPT_CUDA(cudaStreamSynchronize(st));
    if (host_trace_on() && nc > 1 && ctx->host_warm) {
      float ms;
      auto T = [&](cudaEvent_t e) { cudaEventElapsedTime(&ms, g_trace.t0, e); return ms * 1e3f; };
      fprintf(stderr, "[xpt host trace] small-in %.0f us |", T(g_trace.small_in));
      for (int k = 0; k < nc; ++k) { fprintf(stderr, " in%d %.0f", k, T(g_trace.in[k])); fprintf(stderr, " done%d %.0f |", k, T(g_trace.done[k])); }
      fprintf(stderr, " out-end %.0f us\n", T(g_trace.out_end));
    }
    ctx->host_warm = true;
  } else {
    if (!ctx->host_graphs) ctx->host_graphs = new std::vector<xpt_ctx::GraphEntry>();
    cudaGraphExec_t exec = nullptr;
    for (auto& e : *ctx->host_graphs)
      if (e.key == key) { exec = e.exec; nc = e.launches; break; }
    if (!exec) {
      XPT_CUDA(cudaStreamBeginCapture(st, cudaStreamCaptureModeRelaxed));
      rc = host_enqueue(ctx, frames, depth_ms, disp_ms, pose, out, stream, &nc);
      cudaGraph_t graph = nullptr;
      cudaError_t ce = cudaStreamEndCapture(st, &graph);
      if (rc != XPT_OK) { if (graph) cudaGraphDestroy(graph); return rc; }
      if (ce != cudaSuccess) return fail(XPT_CUDA_ERROR, "cudaStreamEndCapture (host call) failed: %s", cudaGetErrorString(ce));
      ce = cudaGraphInstantiate(&exec, graph, 0);
      cudaGraphDestroy(graph);
      if (ce != cudaSuccess) return fail(XPT_CUDA_ERROR, "cudaGraphInstantiate (host call) failed: %s", cudaGetErrorString(ce));
      if (ctx->host_graphs->size() >= 16) {
        cudaGraphExecDestroy(ctx->host_graphs->front().exec);
        ctx->host_graphs->erase(ctx->host_graphs->begin());
      }
      ctx->host_graphs->push_back({key, exec, nc});      // `launches` carries the chunk count here
    }
    XPT_CUDA(cudaGraphLaunch(exec, st));
    // the replayed call is left in flight: xpt_total_loss_host_end waits for this event (not for the whole stream)
    if (!ctx->host_done_ev) XPT_CUDA(cudaEventCreateWithFlags(&ctx->host_done_ev, cudaEventDisableTiming));
    XPT_CUDA(cudaEventRecord(ctx->host_done_ev, st));
    ctx->host_pending_async = true;
  }
  ctx->host_pending = true;
  ctx->host_pending_nc = nc;
  ctx->host_pending_losses = out->losses;
  return XPT_OK;
}

int xpt_total_loss_host_end(xpt_ctx* ctx) {
  if (!ctx) return fail(XPT_BAD_ARGUMENT, "xpt_total_loss_host_end: ctx is NULL");
  if (!ctx->host_pending) return fail(XPT_BAD_ARGUMENT, "xpt_total_loss_host_end: no call in flight on this ctx");
  ctx->host_pending = false;
  if (ctx->host_pending_async) {
    ctx->host_pending_async = false;
    XPT_CUDA(cudaEventSynchronize(ctx->host_done_ev));
  }
  float (*h_losses)[4] = reinterpret_cast<float (*)[4]>(ctx->h_losses);
  for (int j = 0; j < 4; ++j) {
    double v = 0.0;
    for (int k = 0; k < ctx->host_pending_nc; ++k) v += (double)h_losses[k][j];
    ctx->host_pending_losses[j] = (float)v;
  }
  return XPT_OK;
}

int xpt_total_loss_host(xpt_ctx* ctx, const xpt_frames* frames, const float* const depth_ms[],
                        const float* const disp_ms[], const float* pose, const xpt_loss_outputs* out, void* stream) {
  XPT_TRY(xpt_total_loss_host_begin(ctx, frames, depth_ms, disp_ms, pose, out, stream));
  return xpt_total_loss_host_end(ctx);
}

}  // extern "C"
