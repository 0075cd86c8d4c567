// api.cu — the C ABI (include/ppca_b200.h) and the host-side orchestration of the EM engine.
//
// One context = one device + one stream.  The dataset stays resident on the device; per call only the
// model parameters (d k + d + 1 doubles) cross the bus in, and the new model (or per-sample outputs) out.
// Chunked over samples so that the E-step intermediates (packed Gram / second-moment rows) are O(chunk).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <new>
#include <thread>

#include "common.cuh"

namespace ppca {

static thread_local std::string g_last_error;
void set_error(const std::string &msg) { g_last_error = msg; }

enum { FAM_KSYM = 0, FAM_GRAM, FAM_PROJ, FAM_SOLVE, FAM_MOMENT, FAM_CROSS, FAM_FINISH, FAM_SLICE, FAM_COUNT };

}  // namespace ppca

using namespace ppca;

struct ppca_b200_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  int sms = 148;
  int64_t launches = 0;
  int64_t variants[V_COUNT] = {0};
  int64_t chunk = 0;  // 0 = automatic
  // arithmetic of the two masked-Gram contractions IN USE by the current pass: 0 = DMMA, 1 = int8-sliced on mma.sync
  // (ibitgemm.cu), 2 = int8-sliced on tcgen05 (tbitgemm.cu).  base_* is what the caller configured; a pass starts at
  // rung `rung` of the precision ladder (base -> 8 digit planes -> DMMA) and climbs when the guard fires (run_guarded).
  int gemm_mode = 2;
  int slices = 6;
  int base_mode = 2, base_slices = 6;
  bool guard = true;     // PPCA_B200_GUARD=0 disables the precision guard (and with it the ladder)
  int guard_bits = 40;   // eps = 2^-guard_bits: accepted perturbation of M_n / A_i relative to their diagonals
  int rung = 0, rung_ttl = 0;
  int64_t em_rows = 0;   // samples accumulated since em_begin (term count of the M-step guard)
  DevBuf<int8_t> ZQ;             // digit planes of w Z (second M-step contraction, Mask^T (w Z))
  DevBuf<double> ZScale, MZ, part_mz;
  DevBuf<unsigned long long> zcolmax;
  DevBuf<double> dv, part_rx;    // identity-form residual norms of a chunk; partial slots of the exact residual pass
  DevBuf<int> rflag;
  DevBuf<double> rscratch;       // block totals + completion counter of solve_reduce_kernel (zeroed at allocation)
  DevBuf<unsigned int> unsafe;   // [0] E-step guard violations (solve kernels), [1] spare
  DevBuf<double> WScaleMax;      // running maximum over chunks of the W column scales
  DevBuf<int8_t> KsymQ, WQ;
  DevBuf<double> KsymScale, WScale;
  DevBuf<unsigned long long> colmax;
  // model staging
  DevBuf<double> Cdense, mudense, Cpad, mupad, Ksym, logw;
  // chunk workspaces
  DevBuf<double> GW, YZ, WZ, nx, llk, tn, part_bg, part_cr, part_solve, stats, Cnew, cov, rbuf;
  DevBuf<double> mixLP, mixLlk, mixMax, mixSum, mixStats;  // mixture workspaces (grow-only, no per-call cudaMalloc)
  DevBuf<double> gen_ws, gen_Ct, gen_mut;                  // regenerated-chunk EM: generator workspace and truth tables (grow-only)
  DevBuf<double> mixArena;                                 // single-pass mixture EM: per-component chunk buffers
  DevBuf<int8_t> mixArenaQ;                                //   ... and their digit planes
  DevBuf<int> flags;
  DevBuf<double> priorP, priorV;  // mean prior: total precision (d x d) and [m0 | mu_hat | rhs]
  // sample-sharded EM: NCCL communicator of this rank (ppca_b200_comm_init), null = single process
  void *comm = nullptr;
  int comm_rank = 0, comm_world = 1;
  // pinned staging
  double *pinned = nullptr;
  size_t pinned_count = 0;
  double *upload_pin[2] = {nullptr, nullptr};  // double-buffered dataset upload (kept for the context's lifetime)
  size_t upload_count = 0;
  DevBuf<double> upload_raw[2];
  // out-of-core streaming (ppca_b200_iterate_host): copy stream, raw block buffers, block stores
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t ev_copied[2] = {nullptr, nullptr}, ev_free[2] = {nullptr, nullptr};
  DevBuf<double> s_raw[2], s_wraw[2], s_w;
  DevBuf<int64_t> s_rowptr[2];   // compact host format: row offsets and mask words of the block in flight
  DevBuf<uint32_t> s_maskw[2];
  std::shared_ptr<SampleStore> s_store, s_tail;
  // out-of-core inference (ppca_b200_reconstruct_host): D2H stream and double-buffered outputs
  cudaStream_t out_stream = nullptr;
  cudaEvent_t ev_done[2] = {nullptr, nullptr}, ev_outfree[2] = {nullptr, nullptr};
  DevBuf<double> s_out[2], s_llk[2];
  // CUDA graphs of the mixture chunk loop (mix_em_pass): the loop is ~30 small launches per component and chunk — at
  // M = 32 the HOST launch rate, not the GPU, bounded the step on slow hosts (measured 106 ... 336 ms for the same work).
  // A pass with a key seen before (same dataset, shapes, arithmetic, buffers) is captured once and replayed.
  struct MixGraph {
    std::vector<uint64_t> key;
    cudaGraphExec_t exec = nullptr;
    int64_t launches = 0;
    int64_t variants[V_COUNT] = {0};
  };
  std::vector<MixGraph> mix_graphs;
  std::vector<std::vector<uint64_t>> mix_seen;
  bool use_graphs = true;  // PPCA_B200_GRAPHS=0 disables
  int64_t graph_replays = 0;
  // profiling
  bool profiling = false;
  std::vector<cudaEvent_t> ev_pool;
  size_t ev_used = 0;
  struct Span { int fam; cudaEvent_t a, b; };
  std::vector<Span> spans;
  double last_profile[FAM_COUNT] = {0, 0, 0, 0, 0, 0, 0, 0};

  Launcher L() { return Launcher{stream, &launches, sms, variants}; }
  // The pinned staging buffer is written by the host and read by asynchronous copies: instead of draining the whole stream
  // before reusing it, an event marks the last copy that read it (pin_mark) and the next writer waits for that alone.
  cudaEvent_t pin_ev = nullptr;
  bool pin_pending = false;
  void pin_wait() {
    if (pin_pending) CUDA_CHECK(cudaEventSynchronize(pin_ev));
    pin_pending = false;
  }
  void pin_mark() {
    if (!pin_ev) CUDA_CHECK(cudaEventCreateWithFlags(&pin_ev, cudaEventDisableTiming));
    CUDA_CHECK(cudaEventRecord(pin_ev, stream));
    pin_pending = true;
  }
  double *pin(size_t count) {
    if (count > pinned_count) {
      pin_wait();
      if (pinned) cudaFreeHost(pinned);
      pinned = nullptr;
      pinned_count = 0;
      CUDA_CHECK(cudaMallocHost((void **)&pinned, count * sizeof(double)));
      pinned_count = count;
    }
    return pinned;
  }
  cudaEvent_t next_event() {
    if (ev_used == ev_pool.size()) {
      cudaEvent_t e;
      CUDA_CHECK(cudaEventCreate(&e));
      ev_pool.push_back(e);
    }
    return ev_pool[ev_used++];
  }
  void span_begin(int fam) {
    if (!profiling) return;
    Span s{fam, next_event(), nullptr};
    CUDA_CHECK(cudaEventRecord(s.a, stream));
    spans.push_back(s);
  }
  void span_end() {
    if (!profiling) return;
    spans.back().b = next_event();
    CUDA_CHECK(cudaEventRecord(spans.back().b, stream));
  }
  void profile_reset(bool zero = true) {
    spans.clear();
    ev_used = 0;
    if (zero)
      for (int i = 0; i < FAM_COUNT; ++i) last_profile[i] = 0.0;
  }
};

namespace {

struct DeviceGuard {
  int prev = 0;
  explicit DeviceGuard(int dev) {
    cudaGetDevice(&prev);
    if (prev != dev) CUDA_CHECK(cudaSetDevice(dev));
    else prev = -1;
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

template <class F>
int32_t guarded(F &&f) {
  try {
    f();
    return PPCA_OK;
  } catch (const Error &e) {
    set_error(e.msg);
    return e.code;
  } catch (const std::bad_alloc &) {
    set_error("out of host memory");
    return PPCA_ERR_INVALID;
  } catch (...) {
    set_error("unknown error");
    return PPCA_ERR_INVALID;
  }
}

__global__ void fill_value_kernel(double *p, int64_t n, double v) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = v;
}

std::shared_ptr<SampleStore> make_store(ppca_b200_ctx *ctx, int64_t n, int d) {
  auto st = std::make_shared<SampleStore>();
  st->n = n;
  st->d = d;
  st->ldx = (int)round_up(d > 0 ? d : 1, 4);
  st->dw = (int)((d + 31) / 32);
  if (st->dw == 0) st->dw = 1;
  st->n_pad = round_up(n > 0 ? n : 1, 256);
  st->nwT = st->n_pad / 32;
  st->d_pad = (int)round_up(d > 0 ? d : 1, 256);
  st->X.alloc((size_t)st->n_pad * st->ldx);
  st->mask.alloc((size_t)st->n_pad * st->dw);
  st->maskT.alloc((size_t)st->d_pad * st->nwT);
  st->dn.alloc((size_t)st->n_pad);
  CUDA_CHECK(cudaMemsetAsync(st->X.p, 0, sizeof(double) * st->n_pad * st->ldx, ctx->stream));
  CUDA_CHECK(cudaMemsetAsync(st->mask.p, 0, sizeof(uint32_t) * st->n_pad * st->dw, ctx->stream));
  CUDA_CHECK(cudaMemsetAsync(st->maskT.p, 0, sizeof(uint32_t) * st->d_pad * st->nwT, ctx->stream));
  CUDA_CHECK(cudaMemsetAsync(st->dn.p, 0, sizeof(int) * st->n_pad, ctx->stream));
  return st;
}

ppca_b200_dataset *make_dataset(ppca_b200_ctx *ctx, std::shared_ptr<SampleStore> st, const double *w_host) {
  std::unique_ptr<ppca_b200_dataset> ds(new ppca_b200_dataset());
  ds->store = st;
  ds->device = ctx->device;
  ds->w.alloc((size_t)st->n_pad);
  CUDA_CHECK(cudaMemsetAsync(ds->w.p, 0, sizeof(double) * st->n_pad, ctx->stream));
  if (st->n > 0) {
    if (w_host) {
      CUDA_CHECK(cudaMemcpyAsync(ds->w.p, w_host, sizeof(double) * st->n, cudaMemcpyHostToDevice, ctx->stream));
      double mn = std::numeric_limits<double>::infinity();
      for (int64_t i = 0; i < st->n; ++i) mn = w_host[i] < mn ? w_host[i] : (w_host[i] != w_host[i] ? -1.0 : mn);
      ds->min_w = mn;
    } else {
      const int64_t want = (st->n + 255) / 256;
      const int blocks = (int)(want < (int64_t)ctx->sms * 8 ? want : (int64_t)ctx->sms * 8);
      fill_value_kernel<<<blocks, 256, 0, ctx->stream>>>(ds->w.p, st->n, 1.0);
      CUDA_CHECK(cudaGetLastError());
      ++ctx->launches;
      ds->min_w = 1.0;
    }
  }
  CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  return ds.release();
}

// Samples per E/M-step chunk.  Measured on B200 (profiles/r01_summary.md, chunk sweep): bigger is better all the way to
// one chunk per dataset (c2: 5.47 ms at 4 waves of 128-row tiles, 4.17 ms with the whole 1 M rows in one chunk) — the
// per-launch tails and the split-K/slab reductions outweigh any L2 residency of a small chunk.  So: as many rows as
// 8 GiB of chunk workspace holds, at most 2 M (grid.y limits of the 64-row reconstruction tiles; int32-exact digit
// sums need K < 2^24 rows per split-K slab), split evenly over the dataset.
int64_t pick_chunk(ppca_b200_ctx *ctx, int64_t n_pad, const Shape &s) {
  int64_t chunk = ctx->chunk;
  if (chunk <= 0) {
    const int64_t per_row = (int64_t)(s.kkp + 2 * s.kp + 4) * 8 + (int64_t)s.kkp * ctx->slices;
    int64_t cap = ((int64_t)8 << 30) / per_row;
    if (cap > ((int64_t)2 << 20)) cap = (int64_t)2 << 20;
    const int64_t wave = (int64_t)ctx->sms * 128;  // one full wave of 128-row tiles
    if (cap < wave) cap = wave;
    const int64_t nchunks = (n_pad + cap - 1) / cap;
    chunk = (n_pad + nchunks - 1) / nchunks;
  }
  chunk = round_up(chunk, 256);
  if (chunk > n_pad) chunk = n_pad;
  return chunk;
}

// Rows per uploaded block of the out-of-core paths: small enough that the first H2D and the last block's kernels
// (the parts that cannot overlap) are a small fraction of the pass, large enough to keep the kernels efficient.
int64_t pick_stream_block(ppca_b200_ctx *ctx, int64_t n_pad, const Shape &s) {
  int64_t blk = pick_chunk(ctx, n_pad, s);
  const int64_t four_waves = (int64_t)ctx->sms * 128 * 4;
  if (ctx->chunk <= 0 && blk > four_waves) blk = four_waves;
  return blk;
}

// ---- precision guard and ladder -----------------------------------------------------------------------------------
// The int8-sliced contractions keep every term to 8T-2 bits below the maximum of its COLUMN; a sample (dimension) whose
// own terms are much smaller than that maximum loses relative accuracy.  The solve kernels (E-step) and
// mstep_guard_kernel (M-step) bound that loss against the diagonal of the matrix it perturbs and count violations;
// when any is counted the whole pass is repeated one rung up: T -> 8 digit planes -> FP64 DMMA (which has the
// reference's own FP64 semantics).  The rung that was needed is remembered for the next 16 passes of the context.
struct Rung {
  int mode, slices;
};
int ladder(const ppca_b200_ctx *ctx, Rung out[3]) {
  int n = 0;
  out[n++] = Rung{ctx->base_mode, ctx->base_slices};
  if (ctx->guard && ctx->base_mode != 0 && ctx->base_slices != 4) {  // 4 planes = the opt-in FP32-class fast path
    if (ctx->base_slices < 8) out[n++] = Rung{ctx->base_mode, 8};
    out[n++] = Rung{0, ctx->base_slices};
  }
  return n;
}
bool guard_on(const ppca_b200_ctx *ctx) { return ctx->guard && ctx->gemm_mode != 0 && ctx->slices != 4; }
double guard_coef(const ppca_b200_ctx *ctx) { return std::ldexp(1.0, ctx->guard_bits + 1 - 8 * ctx->slices); }
void guard_reset(ppca_b200_ctx *ctx) {
  ctx->unsafe.reserve(2);
  CUDA_CHECK(cudaMemsetAsync(ctx->unsafe.p, 0, 2 * sizeof(unsigned int), ctx->stream));
}
// selects the arithmetic of the next pass from the remembered rung
void begin_pass(ppca_b200_ctx *ctx) {
  Rung r[3];
  const int n = ladder(ctx, r);
  if (ctx->rung >= n) ctx->rung = n - 1;
  ctx->gemm_mode = r[ctx->rung].mode;
  ctx->slices = r[ctx->rung].slices;
  guard_reset(ctx);
}
// after a pass: true = accept; false = the caller must repeat it (the context has moved one rung up)
bool end_pass(ppca_b200_ctx *ctx, double violations) {
  Rung r[3];
  const int n = ladder(ctx, r);
  if (violations > 0.0 && ctx->rung + 1 < n) {
    ++ctx->rung;
    ctx->rung_ttl = 16;
    ++ctx->variants[V_PRECISION_RETRY];
    return false;
  }
  if (ctx->rung > 0 && --ctx->rung_ttl <= 0) ctx->rung = 0;
  return true;
}
// E-step guard violations counted so far in this pass (synchronises the stream)
double read_unsafe(ppca_b200_ctx *ctx) {
  if (!guard_on(ctx)) return 0.0;
  unsigned int h[2] = {0u, 0u};
  CUDA_CHECK(cudaMemcpyAsync(h, ctx->unsafe.p, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  return (double)h[0] + (double)h[1];
}
// body() runs one complete pass and returns its violation count
template <class Body>
void run_guarded(ppca_b200_ctx *ctx, Body &&body) {
  for (;;) {
    begin_pass(ctx);
    if (end_pass(ctx, body())) return;
  }
}

// uploads (C, mu) and builds the padded model + Ksym table (+ its digit planes) into the given device buffers
DevModel stage_model_into(ppca_b200_ctx *ctx, int d, int k, const double *C, const double *mu, double sigma, double *Cpad,
                          double *mupad, double *Ksym, int8_t *KsymQ, double *KsymScale, unsigned long long *colmax) {
  REQUIRE(k >= 1, "state_size must be >= 1 (got %d)", k);
  REQUIRE(C && mu, "null model parameters");
  REQUIRE(sigma > 0.0 && std::isfinite(sigma), "isotropic_noise must be positive and finite");
  Shape s(d, k);
  ctx->Cdense.reserve((size_t)d * k);
  ctx->mudense.reserve((size_t)d);
  double *stage = ctx->pin((size_t)d * k + d);
  ctx->pin_wait();  // the staging buffer may still be the source of an earlier copy
  memcpy(stage, C, sizeof(double) * d * k);
  memcpy(stage + (size_t)d * k, mu, sizeof(double) * d);
  CUDA_CHECK(cudaMemcpyAsync(ctx->Cdense.p, stage, sizeof(double) * d * k, cudaMemcpyHostToDevice, ctx->stream));
  CUDA_CHECK(cudaMemcpyAsync(ctx->mudense.p, stage + (size_t)d * k, sizeof(double) * d, cudaMemcpyHostToDevice,
                             ctx->stream));
  ctx->pin_mark();
  ctx->span_begin(FAM_KSYM);
  launch_prepare_model(ctx->L(), ctx->Cdense.p, ctx->mudense.p, d, k, Cpad, mupad, Ksym);
  const int kblocks = s.d32 / 32;
  if (ctx->gemm_mode == 2)
    launch_slice_tc(ctx->L(), Ksym, s.kkp, d, s.kkp, kblocks, ctx->slices, KsymQ, KsymScale, colmax);
  else if (ctx->gemm_mode == 1)
    launch_slice(ctx->L(), Ksym, s.kkp, d, s.kkp, kblocks, ctx->slices, KsymQ, KsymScale, colmax);
  ctx->span_end();
  DevModel m;
  m.s = s;
  m.sigma = sigma;
  m.C = Cpad;
  m.mu = mupad;
  m.Ksym = Ksym;
  return m;
}

size_t ksym_planes_bytes(const ppca_b200_ctx *ctx, const Shape &s) {
  const int kblocks = s.d32 / 32;
  return ctx->gemm_mode == 2 ? sliced_tc_bytes(kblocks, s.kkp, ctx->slices)
                             : (ctx->gemm_mode == 1 ? sliced_bytes(kblocks, s.kkp, ctx->slices) : 0);
}

// single model: into the context's own buffers
DevModel stage_model(ppca_b200_ctx *ctx, int d, int k, const double *C, const double *mu, double sigma,
                     bool need_ksym = true) {
  REQUIRE(k >= 1, "state_size must be >= 1 (got %d)", k);
  Shape s(d, k);
  ctx->Cpad.reserve((size_t)s.d32 * s.kp);
  ctx->mupad.reserve((size_t)s.d32);
  ctx->Ksym.reserve((size_t)s.d32 * s.kkp);
  if (ctx->gemm_mode != 0) {
    ctx->KsymQ.reserve(ksym_planes_bytes(ctx, s));
    ctx->KsymScale.reserve((size_t)s.kkp);
    ctx->colmax.reserve((size_t)s.kkp);
  }
  return stage_model_into(ctx, d, k, C, mu, sigma, ctx->Cpad.p, ctx->mupad.p, ctx->Ksym.p, ctx->KsymQ.p,
                          ctx->KsymScale.p, ctx->colmax.p);
}

// the buffers a model's chunk kernels work on: its own set (mixture pass) or the context's
ModelWs resolve_ws(ppca_b200_ctx *ctx, const DevModel &m) {
  if (m.ws) return *m.ws;
  ModelWs w;
  w.GW = ctx->GW.p;
  w.YZ = ctx->YZ.p;
  w.WZ = ctx->WZ.p;
  w.nx = ctx->nx.p;
  w.llk = ctx->llk.p;
  w.tn = ctx->tn.p;
  w.KsymQ = ctx->KsymQ.p;
  w.KsymScale = ctx->KsymScale.p;
  w.WQ = ctx->WQ.p;
  w.WScale = ctx->WScale.p;
  w.WScaleMax = ctx->WScaleMax.p;
  w.colmax = ctx->colmax.p;
  w.part_bg = ctx->part_bg.p;
  w.part_cr = ctx->part_cr.p;
  w.part_solve = ctx->part_solve.p;
  w.ZQ = ctx->ZQ.p;
  w.ZScale = ctx->ZScale.p;
  w.MZ = ctx->MZ.p;
  w.part_mz = ctx->part_mz.p;
  w.zcolmax = ctx->zcolmax.p;
  w.dv = ctx->dv.p;
  w.rflag = ctx->rflag.p;
  w.part_rx = ctx->part_rx.p;
  w.rscratch = ctx->rscratch.p;
  return w;
}

void ensure_rscratch(ppca_b200_ctx *ctx) {
  if (ctx->rscratch.p) return;
  ctx->rscratch.alloc((size_t)SOLVE_SCRATCH);
  CUDA_CHECK(cudaMemsetAsync(ctx->rscratch.p, 0, sizeof(double) * SOLVE_SCRATCH, ctx->stream));
}

void reserve_chunk_ws(ppca_b200_ctx *ctx, int64_t chunk, const Shape &s) {
  ctx->GW.reserve((size_t)chunk * s.kkp);
  ctx->YZ.reserve((size_t)chunk * s.kp);
  ctx->WZ.reserve((size_t)chunk * s.kp);
  ctx->nx.reserve((size_t)chunk);
  ctx->llk.reserve((size_t)chunk);
  ctx->tn.reserve((size_t)chunk);
  ctx->dv.reserve((size_t)chunk);
  ctx->rflag.reserve(8);
  ensure_rscratch(ctx);
}

// E-step of one chunk: Gram contraction, projection, per-sample solve
void e_step_chunk(ppca_b200_ctx *ctx, const SampleStore &st, const double *w, int64_t row0, int rows,
                  const DevModel &m, int mode, double *llk_out, double *cov_out, double *solve_part,
                  unsigned long long *w_colmax = nullptr) {
  const Launcher L = ctx->L();
  const ModelWs ws = resolve_ws(ctx, m);
  const int rows_pad = (int)round_up(rows, 256);
  ctx->span_begin(FAM_GRAM);
  if (ctx->gemm_mode == 2) {
    launch_tbitgemm(L, st.mask.p + row0 * st.dw, st.dw, st.dw, ws.KsymQ, ws.KsymScale, ctx->slices, ws.GW,
                    m.s.kkp, rows, m.s.kkp, (m.s.d32 / 32 + 3) / 4, 0, nullptr, 1, 0);
  } else if (ctx->gemm_mode == 1) {
    IBitGemmArgs g;
    g.bits = st.mask.p + row0 * st.dw;
    g.ldbits = st.dw;
    g.Bq = ws.KsymQ;
    g.scale = ws.KsymScale;
    g.T = ctx->slices;
    g.Out = ws.GW;
    g.ldo = m.s.kkp;
    g.M = rows;
    g.Nq = m.s.kkp;
    g.kblocks = m.s.d32 / 32;
    g.accumulate = 0;
    g.partials = nullptr;
    g.splitk = 1;
    g.defer_reduce = 0;
    launch_ibitgemm(L, g);
  } else {
    BitGemmArgs g;
    g.bits = st.mask.p + row0 * st.dw;
    g.ldbits = st.dw;
    g.Bmat = m.Ksym;
    g.ldb = m.s.kkp;
    g.Out = ws.GW;
    g.ldo = m.s.kkp;
    g.M = rows;
    g.Nq = m.s.kkp;
    g.kblocks = m.s.d32 / 32;
    g.kcols = m.s.d;
    g.accumulate = 0;
    g.partials = nullptr;
    g.splitk = 1;
    g.defer_reduce = 0;
    launch_bitgemm(L, g);
  }
  ctx->span_end();
  ctx->span_begin(FAM_PROJ);
  launch_proj(L, st, row0, rows, m, ws.YZ, ws.nx);
  ctx->span_end();
  SolveArgs sa;
  sa.s = m.s;
  sa.sigma = m.sigma;
  sa.sigma_dev = m.sigma_dev;
  sa.rows = rows;
  sa.rows_pad = rows_pad;
  sa.GW = ws.GW;
  sa.YZ = ws.YZ;
  sa.WZ = (mode == 2 && w) ? ws.WZ : nullptr;
  sa.nx = ws.nx;
  sa.dn = st.dn.p + row0;
  sa.w = w ? w + row0 : nullptr;
  sa.llk = llk_out;
  sa.tn = mode == 2 ? ws.tn : nullptr;
  sa.dv = mode == 2 ? ws.dv : nullptr;
  sa.cov = cov_out;
  sa.part = solve_part;
  sa.rscratch = ws.rscratch;
  sa.mode = mode;
  sa.colmax = w_colmax;
  if (guard_on(ctx)) {
    sa.gscale = ws.KsymScale;
    sa.guard_coef = guard_coef(ctx);
    sa.unsafe = ctx->unsafe.p;
  }
  if (w_colmax) CUDA_CHECK(cudaMemsetAsync(w_colmax, 0, sizeof(unsigned long long) * m.s.kkp, ctx->stream));
  ctx->span_begin(FAM_SOLVE);
  launch_solve(L, sa);
  ctx->span_end();
}

// E-step + local M-step statistics, chunk by chunk.  em_begin zeroes the statistics and the partial slots (owned by
// fixed CTAs and accumulated across chunks), em_chunk runs one chunk of `rows` samples starting at row0 of `st`,
// em_end reduces the partial slots once, in fixed order.  The resident path (em_stats_impl) walks one store; the
// out-of-core path (ppca_b200_iterate_host) feeds it one freshly uploaded block store after another.
struct EmPlan {
  int64_t chunk = 0;
  int splitk = 1;
  size_t bglen = 0, crlen = 0;
  int slabs = 1;
  int splitk_mz = 1;  // split-K of the second contraction Mask^T (w Z) (d x kp)
  size_t mzlen = 0;
  size_t rxlen = 0;   // partial slots of the exact residual pass
};

size_t planes_bytes(const ppca_b200_ctx *ctx, int64_t chunk, int Nq) {
  const int kb_chunk = (int)(chunk / 32);
  return ctx->gemm_mode == 2 ? sliced_tc_bytes(kb_chunk, Nq, ctx->slices)
                             : (ctx->gemm_mode == 1 ? sliced_bytes(kb_chunk, Nq, ctx->slices) : 0);
}
size_t wq_bytes(const ppca_b200_ctx *ctx, int64_t chunk, const Shape &s) { return planes_bytes(ctx, chunk, s.kkp); }
size_t zq_bytes(const ppca_b200_ctx *ctx, int64_t chunk, const Shape &s) { return planes_bytes(ctx, chunk, s.kp); }

// sizes of the split-K / slab partial slots for chunks of `chunk` rows
EmPlan em_plan(const ppca_b200_ctx *ctx, int64_t chunk, const Shape &s) {
  EmPlan p;
  p.chunk = chunk;
  const int kb_chunk = (int)(p.chunk / 32);
  p.splitk = ctx->gemm_mode == 2   ? tbitgemm_pick_splitk(s.d, s.kkp, kb_chunk / 4, ctx->sms)
             : ctx->gemm_mode == 1 ? ibitgemm_pick_splitk(s.d, s.kkp, kb_chunk, ctx->sms)
                                   : bitgemm_pick_splitk(s.d, s.kkp, kb_chunk, ctx->sms);
  p.bglen = bitgemm_partials_len(s.d, s.kkp, p.splitk);
  p.splitk_mz = ctx->gemm_mode == 2   ? tbitgemm_pick_splitk(s.d, s.kp, kb_chunk / 4, ctx->sms)
                : ctx->gemm_mode == 1 ? ibitgemm_pick_splitk(s.d, s.kp, kb_chunk, ctx->sms)
                                      : bitgemm_pick_splitk(s.d, s.kp, kb_chunk, ctx->sms);
  p.mzlen = bitgemm_partials_len(s.d, s.kp, p.splitk_mz);
  p.slabs = cross_resid_slabs(s.d, s.k, (int)p.chunk, ctx->sms);
  p.crlen = cross_resid_partials_len(s.d, s.k, p.slabs);
  p.rxlen = resid_exact_partials_len(s.d, p.slabs);
  return p;
}

// zeroes the statistics and the partial slots of one model
void em_zero(ppca_b200_ctx *ctx, const ModelWs &ws, const EmPlan &p, const Shape &s, double *stats_dev) {
  CUDA_CHECK(cudaMemsetAsync(stats_dev, 0, sizeof(double) * StatsLayout(s.d, s.k).len, ctx->stream));
  if (guard_on(ctx)) CUDA_CHECK(cudaMemsetAsync(ws.WScaleMax, 0, sizeof(double) * s.kkp, ctx->stream));
  if (p.bglen) CUDA_CHECK(cudaMemsetAsync(ws.part_bg, 0, sizeof(double) * p.bglen, ctx->stream));
  CUDA_CHECK(cudaMemsetAsync(ws.part_cr, 0, sizeof(double) * p.crlen, ctx->stream));
  CUDA_CHECK(cudaMemsetAsync(ws.part_solve, 0, sizeof(double) * SOLVE_SLOTS * 4, ctx->stream));
  CUDA_CHECK(cudaMemsetAsync(ws.part_rx, 0, sizeof(double) * p.rxlen, ctx->stream));
  CUDA_CHECK(cudaMemsetAsync(ws.rflag, 0, sizeof(int), ctx->stream));
  CUDA_CHECK(cudaMemsetAsync(ws.MZ, 0, sizeof(double) * (size_t)s.d * s.kp, ctx->stream));
  if (p.mzlen) CUDA_CHECK(cudaMemsetAsync(ws.part_mz, 0, sizeof(double) * p.mzlen, ctx->stream));
}

EmPlan em_begin(ppca_b200_ctx *ctx, int64_t chunk, const DevModel &m, double *stats_dev) {
  const EmPlan p = em_plan(ctx, chunk, m.s);
  reserve_chunk_ws(ctx, p.chunk, m.s);
  if (ctx->gemm_mode != 0) {
    ctx->WQ.reserve(wq_bytes(ctx, p.chunk, m.s));
    ctx->WScale.reserve((size_t)m.s.kkp);
    ctx->colmax.reserve((size_t)m.s.kkp);
  }
  ctx->em_rows = 0;
  if (guard_on(ctx)) ctx->WScaleMax.reserve((size_t)m.s.kkp);
  ctx->part_bg.reserve(p.bglen);
  ctx->part_cr.reserve(p.crlen);
  ctx->part_solve.reserve((size_t)SOLVE_SLOTS * 4);
  if (ctx->gemm_mode != 0) {
    ctx->ZQ.reserve(zq_bytes(ctx, p.chunk, m.s));
    ctx->ZScale.reserve((size_t)m.s.kp);
    ctx->zcolmax.reserve((size_t)m.s.kp);
  }
  ctx->MZ.reserve((size_t)m.s.d * m.s.kp);
  ctx->part_mz.reserve(p.mzlen);
  ctx->part_rx.reserve(p.rxlen);
  em_zero(ctx, resolve_ws(ctx, m), p, m.s, stats_dev);
  return p;
}

// M-step statistics of one chunk whose second moments W (ws.GW), states (ws.YZ) and weighted states (ws.WZ) are in place;
// wloc = the chunk's weights (index 0 = row0); have_colmax: ws.colmax already holds max_n |W[n][q]|
void m_step_chunk(ppca_b200_ctx *ctx, const SampleStore &st, const double *wloc, int64_t row0, int rows, const DevModel &m,
                  double *stats_dev, const EmPlan &p, bool have_colmax) {
  const Launcher L = ctx->L();
  const ModelWs ws = resolve_ws(ctx, m);
  const StatsLayout lay(m.s.d, m.s.k);
  const int kblocks = (int)(round_up(rows, 32) / 32);
  // Out[d x Nq] += Mask^T[d x rows] * Bsrc[rows x Nq] on the configured arithmetic, split-K partials deferred to em_end
  auto contract = [&](const double *Bsrc, int Nq, int8_t *Q, double *Scale, unsigned long long *cmax, bool have_cmax,
                      double *Out, double *partials, int splitk, double *scale_max) {
    int sk = splitk < kblocks ? splitk : (kblocks > 0 ? kblocks : 1);
    if (splitk > 1 && sk < 2) sk = 2 <= kblocks ? 2 : 1;
    const bool direct = splitk > 1 && sk == 1;  // degenerate last chunk: accumulate straight into the output
    if (ctx->gemm_mode == 2) {
      ctx->span_begin(FAM_SLICE);
      launch_slice_tc(L, Bsrc, Nq, rows, Nq, kblocks, ctx->slices, Q, Scale, cmax, have_cmax);
      ctx->span_end();
    } else if (ctx->gemm_mode == 1) {
      ctx->span_begin(FAM_SLICE);
      launch_slice(L, Bsrc, Nq, rows, Nq, kblocks, ctx->slices, Q, Scale, cmax);
      ctx->span_end();
    }
    if (scale_max && guard_on(ctx)) launch_scale_max(L, Scale, Nq, scale_max);
    ctx->span_begin(FAM_MOMENT);
    if (ctx->gemm_mode == 2) {
      const int ksteps = (kblocks + 3) / 4;
      int skt = splitk < ksteps ? splitk : ksteps;
      if (skt < 1) skt = 1;
      {  // no empty slabs
        const int per = (ksteps + skt - 1) / skt;
        skt = (ksteps + per - 1) / per;
      }
      const bool to_partials = splitk > 1 && skt > 1;
      launch_tbitgemm(L, st.maskT.p + row0 / 32, st.nwT, 4 * ksteps, Q, Scale, ctx->slices, Out, Nq, m.s.d, Nq, ksteps, 1,
                      to_partials ? partials : nullptr, skt, to_partials ? 1 : 0);
    } else if (ctx->gemm_mode == 1) {
      IBitGemmArgs g;
      g.bits = st.maskT.p + row0 / 32;
      g.ldbits = st.nwT;
      g.Bq = Q;
      g.scale = Scale;
      g.T = ctx->slices;
      g.Out = Out;
      g.ldo = Nq;
      g.M = m.s.d;
      g.Nq = Nq;
      g.kblocks = kblocks;
      g.accumulate = 1;
      g.splitk = sk;
      g.partials = (splitk > 1 && !direct) ? partials : nullptr;
      g.defer_reduce = direct ? 0 : 1;
      launch_ibitgemm(L, g);
    } else {
      BitGemmArgs g;
      g.bits = st.maskT.p + row0 / 32;
      g.ldbits = st.nwT;
      g.Bmat = Bsrc;
      g.ldb = Nq;
      g.Out = Out;
      g.ldo = Nq;
      g.M = m.s.d;
      g.Nq = Nq;
      g.kblocks = kblocks;
      g.kcols = 32 * kblocks;
      g.accumulate = 1;
      g.splitk = sk;
      g.partials = (splitk > 1 && !direct) ? partials : nullptr;
      g.defer_reduce = direct ? 0 : 1;
      launch_bitgemm(L, g);
    }
    ctx->span_end();
  };
  // A += Mask^T W   (second moments, ppca_model.rs:297-306)
  contract(ws.GW, m.s.kkp, ws.WQ, ws.WScale, ws.colmax, have_colmax, stats_dev + lay.offA, ws.part_bg, p.splitk,
           ws.WScaleMax);
  // MZ += Mask^T (w Z)   (the part of total_deviation that needs the mask, ppca_model.rs:338-347)
  contract(ws.WZ, m.s.kp, ws.ZQ, ws.ZScale, ws.zcolmax, false, ws.MZ, ws.part_mz, p.splitk_mz, nullptr);
  ctx->span_begin(FAM_CROSS);
  launch_cross_resid(L, st, row0, rows, m, ws.WZ, wloc, ws.part_cr, p.slabs);
  // exact residual norms for a chunk whose identity-form ones were flagged (returns at once otherwise)
  launch_resid_exact(L, st, row0, rows, m, ws.YZ, wloc, ws.rflag, ws.part_rx, p.slabs);
  ctx->span_end();
}

void em_chunk(ppca_b200_ctx *ctx, const SampleStore &st, const double *w, int64_t row0, int rows, const DevModel &m,
              double *stats_dev, const EmPlan &p) {
  const ModelWs ws = resolve_ws(ctx, m);
  // the solve kernels (state_size <= 64) leave the column maxima of W behind for the tcgen05 digit planes
  const bool fused_colmax = ctx->gemm_mode == 2 && m.s.k <= 64;
  e_step_chunk(ctx, st, w, row0, rows, m, 2, ws.llk, nullptr, nullptr, fused_colmax ? ws.colmax : nullptr);
  ctx->span_begin(FAM_SOLVE);
  launch_solve_reduce(ctx->L(), rows, ws.llk, ws.tn, ws.dv, ws.nx, st.dn.p + row0, w + row0, ws.part_solve, ws.rscratch,
                      ws.rflag);
  ctx->span_end();
  ctx->em_rows += rows;
  m_step_chunk(ctx, st, w + row0, row0, rows, m, stats_dev, p, fused_colmax);
}

void em_end(ppca_b200_ctx *ctx, const DevModel &m, double *stats_dev, const EmPlan &p, int64_t rows_total = -1) {
  const Launcher L = ctx->L();
  const ModelWs ws = resolve_ws(ctx, m);
  const StatsLayout lay(m.s.d, m.s.k);
  if (rows_total < 0) rows_total = ctx->em_rows;
  ctx->span_begin(FAM_SOLVE);
  launch_solve_finish(L, ws.part_solve, stats_dev + lay.offScalars);
  ctx->span_end();
  ctx->span_begin(FAM_MOMENT);
  launch_bitgemm_reduce(L, ws.part_bg, p.splitk, m.s.d, m.s.kkp, stats_dev + lay.offA, m.s.kkp, 1);
  launch_bitgemm_reduce(L, ws.part_mz, p.splitk_mz, m.s.d, m.s.kp, ws.MZ, m.s.kp, 1);
  ctx->span_end();
  ctx->span_begin(FAM_CROSS);
  launch_cross_resid_finish(L, m.s.d, m.s.k, ws.part_cr, p.slabs, m.C, ws.MZ, stats_dev + lay.offB,
                            stats_dev + lay.offTdev, stats_dev + lay.offTotals);
  launch_resid_exact_finish(L, m.s.d, ws.part_rx, p.slabs, stats_dev + lay.offScalars);
  ctx->span_end();
  if (guard_on(ctx)) {  // scalars 5, 6: guard violations of this shard's E- and M-step contractions
    ctx->span_begin(FAM_FINISH);
    launch_mstep_guard(L, m.s.d, m.s.k, stats_dev + lay.offA, stats_dev + lay.offTotals, ws.WScaleMax,
                       guard_terms((double)rows_total) * guard_coef(ctx), ctx->unsafe.p, stats_dev + lay.offScalars);
    ctx->span_end();
  }
}

// E-step + local M-step statistics of a whole (local, resident) dataset into stats_dev
void em_stats_impl(ppca_b200_ctx *ctx, const SampleStore &st, const double *w, const DevModel &m, double *stats_dev) {
  if (st.n == 0) {
    CUDA_CHECK(cudaMemsetAsync(stats_dev, 0, sizeof(double) * StatsLayout(m.s.d, m.s.k).len, ctx->stream));
    return;
  }
  const EmPlan p = em_begin(ctx, pick_chunk(ctx, st.n_pad, m.s), m, stats_dev);
  for (int64_t row0 = 0; row0 < st.n; row0 += p.chunk) {
    const int rows = (int)((st.n - row0) < p.chunk ? (st.n - row0) : p.chunk);
    em_chunk(ctx, st, w, row0, rows, m, stats_dev, p);
  }
  em_end(ctx, m, stats_dev, p);
}

// Householder QR solve on the host (prior.rs:97-110 smooth_mean uses total_precision.qr().solve)
bool host_qr_solve(std::vector<double> &A, int n, std::vector<double> &b) {
  std::vector<double> diag(n);
  for (int i = 0; i < n; ++i) {
    double sq = 0.0;
    for (int r = i; r < n; ++r) sq += A[(size_t)r * n + i] * A[(size_t)r * n + i];
    const double norm = std::sqrt(sq);
    const double a0 = A[(size_t)i * n + i];
    const double sgn = std::signbit(a0) ? -1.0 : 1.0;
    const double signed_norm = sgn * norm;
    const double factor = (sq + std::fabs(a0) * norm) * 2.0;
    A[(size_t)i * n + i] = a0 + signed_norm;
    if (factor != 0.0) {
      const double s = std::sqrt(factor);
      for (int r = i; r < n; ++r) A[(size_t)r * n + i] /= s;
      diag[i] = -signed_norm;
      for (int c = i + 1; c < n; ++c) {
        double dot = 0.0;
        for (int r = i; r < n; ++r) dot += A[(size_t)r * n + i] * A[(size_t)r * n + c];
        dot *= 2.0;
        for (int r = i; r < n; ++r) A[(size_t)r * n + c] -= dot * A[(size_t)r * n + i];
      }
      double dot = 0.0;
      for (int r = i; r < n; ++r) dot += A[(size_t)r * n + i] * b[r];
      dot *= 2.0;
      for (int r = i; r < n; ++r) b[r] -= dot * A[(size_t)r * n + i];
    } else {
      diag[i] = signed_norm;
    }
  }
  for (int i = n - 1; i >= 0; --i) {
    if (diag[i] == 0.0) return false;
    double v = b[i];
    for (int c = i + 1; c < n; ++c) v -= A[(size_t)i * n + c] * b[c];
    b[i] = v / diag[i];
  }
  return true;
}

// M-step finish from (reduced) statistics
// returns the precision-guard violations carried by the (reduced) statistics
double em_finish_impl(ppca_b200_ctx *ctx, int d, int k, const double *C, const double *mu, double sigma,
                      const ppca_b200_prior *prior, const double *stats_dev, double *C_out, double *mu_out,
                      double *sigma_out, double *llk_in, double *sumw_out, double *viol_em = nullptr) {
  REQUIRE(C && mu && C_out && mu_out && sigma_out, "null model parameters");
  const Shape s(d, k);
  const StatsLayout lay(d, k);
  // old transform, padded (fallback rows)
  ctx->Cdense.reserve((size_t)d * k);
  ctx->mudense.reserve((size_t)d);
  ctx->Cpad.reserve((size_t)s.d32 * s.kp);
  ctx->mupad.reserve((size_t)s.d32);
  ctx->Ksym.reserve((size_t)s.d32 * s.kkp);
  ctx->Cnew.reserve((size_t)d * k);
  ctx->flags.reserve((size_t)d);
  const size_t tail = (size_t)2 * d + 8;
  double *stage = ctx->pin((size_t)d * k + tail + (size_t)d * k);
  ctx->pin_wait();  // no stream drain here: everything below is enqueued behind the statistics kernels still running
  memcpy(stage, C, sizeof(double) * d * k);
  CUDA_CHECK(cudaMemcpyAsync(ctx->Cdense.p, stage, sizeof(double) * d * k, cudaMemcpyHostToDevice, ctx->stream));
  // only the padded old transform is needed on the device here (fallback rows); mu stays on the host
  CUDA_CHECK(cudaMemsetAsync(ctx->mudense.p, 0, sizeof(double) * d, ctx->stream));
  ctx->span_begin(FAM_FINISH);
  launch_prepare_model(ctx->L(), ctx->Cdense.p, ctx->mudense.p, d, k, ctx->Cpad.p, ctx->mupad.p, ctx->Ksym.p);
  const double tau = prior ? prior->transformation_precision : 0.0;
  launch_row_solve(ctx->L(), d, k, stats_dev + lay.offA, stats_dev + lay.offB, tau, ctx->Cpad.p, ctx->Cnew.p,
                   ctx->flags.p);
  ctx->span_end();
  double *h_tail = stage + (size_t)d * k;
  double *h_C = h_tail + tail;
  CUDA_CHECK(cudaMemcpyAsync(h_tail, stats_dev + lay.offTdev, sizeof(double) * tail, cudaMemcpyDeviceToHost,
                             ctx->stream));
  CUDA_CHECK(cudaMemcpyAsync(h_C, ctx->Cnew.p, sizeof(double) * d * k, cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_CHECK(cudaStreamSynchronize(ctx->stream));  // the one synchronisation of an EM step
  ctx->pin_pending = false;
  const double *tdev = h_tail, *totals = h_tail + d, *sc = h_tail + 2 * d;
  if (!(sc[SC_NONEMPTY] > 0.0)) PPCA_THROW(PPCA_ERR_EMPTY, "non-empty dataset required (ppca_model.rs:358)");
  memcpy(C_out, h_C, sizeof(double) * d * k);
  double tot_sum = 0.0;
  for (int i = 0; i < d; ++i) tot_sum += totals[i];
  double noise_sq;
  if (prior && prior->has_isotropic_noise_prior)  // ppca_model.rs:360-368
    noise_sq = ((sc[SC_SQERR] + sc[SC_DEV2]) / 2.0 + prior->isotropic_noise_beta) /
               (tot_sum / 2.0 + prior->isotropic_noise_alpha + 1.0);
  else
    noise_sq = (sc[SC_SQERR] + sc[SC_DEV2]) / tot_sum;  // :370
  for (int i = 0; i < d; ++i) mu_out[i] = (totals[i] > 0.0 ? tdev[i] / totals[i] : 0.0) + mu[i];  // :373-377
  if (prior && prior->has_mean_prior) {  // :379-384 ; prior.rs:97-110
    REQUIRE(prior->mean && prior->mean_precision, "mean prior needs mean and mean_precision");
    // device path: blocked Cholesky of the (SPD) total precision; see finish.cu
    bool solved = false;
    {
      ctx->priorP.reserve((size_t)d * d);
      ctx->priorV.reserve((size_t)3 * d);
      ctx->flags.reserve((size_t)d + 1);
      double *m0_dev = ctx->priorV.p, *muhat_dev = ctx->priorV.p + d, *rhs_dev = ctx->priorV.p + 2 * (size_t)d;
      CUDA_CHECK(cudaMemcpyAsync(ctx->priorP.p, prior->mean_precision, sizeof(double) * (size_t)d * d,
                                 cudaMemcpyHostToDevice, ctx->stream));
      CUDA_CHECK(cudaMemcpyAsync(m0_dev, prior->mean, sizeof(double) * d, cudaMemcpyHostToDevice, ctx->stream));
      CUDA_CHECK(cudaMemcpyAsync(muhat_dev, mu_out, sizeof(double) * d, cudaMemcpyHostToDevice, ctx->stream));
      ctx->span_begin(FAM_FINISH);
      launch_mean_prior_solve(ctx->L(), ctx->priorP.p, d, m0_dev, stats_dev + lay.offTotals, muhat_dev, noise_sq, rhs_dev,
                              ctx->flags.p + d);
      ctx->span_end();
      std::vector<double> sol((size_t)d);
      int fail = 0;
      CUDA_CHECK(cudaMemcpyAsync(sol.data(), rhs_dev, sizeof(double) * d, cudaMemcpyDeviceToHost, ctx->stream));
      CUDA_CHECK(cudaMemcpyAsync(&fail, ctx->flags.p + d, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
      CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
      if (!fail) {
        bool finite = true;
        for (int i = 0; i < d; ++i) finite = finite && std::isfinite(sol[i]);
        if (finite) {
          for (int i = 0; i < d; ++i) mu_out[i] = sol[i];
          solved = true;
        }
      }
    }
    if (!solved) {  // precision matrix not positive definite: the reference's QR route, on the host
      std::vector<double> P((size_t)d * d), num(d);
      for (int i = 0; i < d; ++i) {
        double acc = 0.0;
        for (int j = 0; j < d; ++j) {
          const double pm = prior->mean_precision[(size_t)i * d + j];
          P[(size_t)i * d + j] = pm + (i == j ? totals[i] / noise_sq : 0.0);
          acc += pm * prior->mean[j];
        }
        num[i] = acc + (totals[i] / noise_sq) * mu_out[i];
      }
      if (!host_qr_solve(P, d, num))
        PPCA_THROW(PPCA_ERR_NUMERIC, "total precision matrix is always invertible (prior.rs:109)");
      for (int i = 0; i < d; ++i) mu_out[i] = num[i];
    }
  }
  *sigma_out = std::sqrt(noise_sq);  // :389
  if (llk_in) *llk_in = sc[SC_LLK];
  if (sumw_out) *sumw_out = sc[SC_SUMW];
  if (viol_em) {
    viol_em[0] = sc[SC_UNSAFE_E];
    viol_em[1] = sc[SC_UNSAFE_M];
  }
  return sc[SC_UNSAFE_E] + sc[SC_UNSAFE_M];
}

void check_ds(const ppca_b200_ctx *ctx, const ppca_b200_dataset *ds) {
  REQUIRE(ctx != nullptr, "null context");
  REQUIRE(ds != nullptr && ds->store, "null dataset");
  REQUIRE(ds->device == ctx->device, "dataset lives on device %d, context on %d", ds->device, ctx->device);
}

// llks of one model for all samples into llk_dev (device, n entries, optionally strided)
void llks_impl(ppca_b200_ctx *ctx, const SampleStore &st, const DevModel &m, double *llk_dev) {
  if (st.n == 0) return;
  const int64_t chunk = pick_chunk(ctx, st.n_pad, m.s);
  reserve_chunk_ws(ctx, chunk, m.s);
  for (int64_t row0 = 0; row0 < st.n; row0 += chunk) {
    const int rows = (int)((st.n - row0) < chunk ? (st.n - row0) : chunk);
    e_step_chunk(ctx, st, nullptr, row0, rows, m, 0, llk_dev + row0, nullptr, nullptr);
  }
}

__global__ void fill_ones_mask_kernel(uint32_t *mask, int dw, int d, int64_t n, int *dn) {
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < n * dw;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int j = (int)(idx % dw);
    const int bits = d - 32 * j;
    mask[idx] = bits >= 32 ? 0xffffffffu : ((1u << bits) - 1u);
    if (j == 0) dn[idx / dw] = d;
  }
}

__global__ void strided_copy_kernel(const double *src, int64_t n, int cols, int64_t lds, double *dst, int64_t ldd) {
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < n * cols;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = idx / cols;
    const int c = (int)(idx % cols);
    dst[r * ldd + c] = src[r * lds + c];
  }
}

// smooth / extrapolate into a fresh all-observed store; scale/accumulate for mixtures
void reconstruct_impl(ppca_b200_ctx *ctx, const SampleStore &st, const DevModel &m, int extrapolate,
                      const double *scale, int64_t scale_ld, int accumulate, SampleStore &out,
                      double *llk_dev = nullptr) {
  if (st.n == 0) return;
  const int64_t chunk = pick_chunk(ctx, st.n_pad, m.s);
  reserve_chunk_ws(ctx, chunk, m.s);
  for (int64_t row0 = 0; row0 < st.n; row0 += chunk) {
    const int rows = (int)((st.n - row0) < chunk ? (st.n - row0) : chunk);
    // the per-sample log-likelihoods are a by-product of the same E-step (llk_dev: n entries, nullable)
    e_step_chunk(ctx, st, nullptr, row0, rows, m, 1, llk_dev ? llk_dev + row0 : nullptr, nullptr, nullptr);
    launch_reconstruct(ctx->L(), st, row0, rows, m, ctx->YZ.p, extrapolate, scale ? scale + row0 * scale_ld : nullptr,
                       scale_ld, accumulate, out.X.p + row0 * out.ldx, out.ldx);
  }
}

void finalize_full_store(ppca_b200_ctx *ctx, SampleStore &out) {
  if (out.n == 0) return;
  const int64_t total = out.n * out.dw;
  const int blocks = (int)((total + 255) / 256 < (int64_t)ctx->sms * 8 ? (total + 255) / 256 : (int64_t)ctx->sms * 8);
  fill_ones_mask_kernel<<<blocks, 256, 0, ctx->stream>>>(out.mask.p, out.dw, out.d, out.n, out.dn.p);
  CUDA_CHECK(cudaGetLastError());
  ++ctx->launches;
  launch_transpose_mask(ctx->L(), out);
}


// ---- covariance diagonals (InferredMasked::smoothed_ / extrapolated_covariance_diagonal) ----------------------------
// S2[n][q(a,b)] = (a == b ? 1 : 2) Sigma_n[a][b] (packed upper, zero padded): diag_i = sigma^2 + sum_q S2[n][q] Ksym[i][q]
__global__ void pack_cov_kernel(const double *__restrict__ cov, int64_t rows, int64_t rows_pad, int k, int kk, int kkp,
                                double *S2) {
  const int64_t total = rows_pad * kkp;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t n = idx / kkp;
    const int q = (int)(idx % kkp);
    double v = 0.0;
    if (n < rows && q < kk) {
      int a = 0, off = q;  // q -> (a, b) of the packed upper triangle by rows
      while (off >= k - a) {
        off -= k - a;
        ++a;
      }
      const int b = a + off;
      v = cov[n * k * k + (int64_t)a * k + b] * (a == b ? 1.0 : 2.0);
    }
    S2[idx] = v;
  }
}

// KT[q][i] = Ksym[i][q] for i < d, q < kk (zero elsewhere): kq32 x d8
__global__ void transpose_ksym_kernel(const double *__restrict__ Ksym, int d, int kk, int kkp, int kq32, int d8, double *KT) {
  const int64_t total = (int64_t)kq32 * d8;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int q = (int)(idx / d8), i = (int)(idx % d8);
    KT[idx] = (i < d && q < kk) ? Ksym[(int64_t)i * kkp + q] : 0.0;
  }
}

__global__ void cov_diag_finish_kernel(const double *__restrict__ Y, int d8, int64_t rows, int d, double s2,
                                       const uint32_t *__restrict__ mask, int dw, int64_t mrow0, double *outX, int ldx) {
  const int64_t total = rows * d;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t n = idx / d;
    const int i = (int)(idx % d);
    const bool observed = mask && ((mask[(mrow0 + n) * dw + (i >> 5)] >> (i & 31)) & 1u);
    outX[n * ldx + i] = observed ? 0.0 : Y[n * d8 + i] + s2;  // expand() leaves observed slots at 0 (ppca_model.rs:576)
  }
}

struct MixView {
  int m;
  const int32_t *ks;
  const double *Cs, *mus, *sigmas, *logw;
  int d;
  const double *C(int j) const {
    size_t off = 0;
    for (int l = 0; l < j; ++l) off += (size_t)d * ks[l];
    return Cs + off;
  }
  const double *mu(int j) const { return mus + (size_t)j * d; }
};

void check_mix(const MixView &mv) {
  REQUIRE(mv.m >= 1, "a mixture needs at least one model (mix.rs:51)");
  REQUIRE(mv.ks && mv.Cs && mv.mus && mv.sigmas && mv.logw, "null mixture parameters");
}

// component log-likelihoods into LP (n x m, device), then in-place log-softmax with the prior log-weights
void mix_posteriors_impl(ppca_b200_ctx *ctx, const ppca_b200_dataset *ds, const MixView &mv, double *LP_dev,
                         double *mix_llk_dev, double *comp_max_dev, double *llk_sum_dev) {
  const SampleStore &st = *ds->store;
  ctx->rbuf.reserve((size_t)st.n_pad);
  for (int j = 0; j < mv.m; ++j) {
    DevModel m = stage_model(ctx, st.d, mv.ks[j], mv.C(j), mv.mu(j), mv.sigmas[j]);
    llks_impl(ctx, st, m, ctx->rbuf.p);
    if (st.n > 0) {
      const int64_t total = st.n;
      const int blocks = (int)((total + 255) / 256 < (int64_t)ctx->sms * 8 ? (total + 255) / 256 : (int64_t)ctx->sms * 8);
      strided_copy_kernel<<<blocks, 256, 0, ctx->stream>>>(ctx->rbuf.p, st.n, 1, 1, LP_dev + j, mv.m);
      CUDA_CHECK(cudaGetLastError());
      ++ctx->launches;
    }
  }
  ctx->logw.reserve((size_t)mv.m);
  CUDA_CHECK(cudaMemcpyAsync(ctx->logw.p, mv.logw, sizeof(double) * mv.m, cudaMemcpyHostToDevice, ctx->stream));
  launch_log_softmax_rows(ctx->L(), LP_dev, st.n, mv.m, ctx->logw.p, ds->w.p, mix_llk_dev, comp_max_dev, llk_sum_dev);
}

}  // namespace

// =================================================================================================
// C ABI
// =================================================================================================
extern "C" {

int32_t ppca_b200_abi_version(void) { return PPCA_B200_ABI_VERSION; }
const char *ppca_b200_last_error(void) { return g_last_error.c_str(); }

int32_t ppca_b200_device_count(int32_t *out) {
  return guarded([&] {
    REQUIRE(out != nullptr, "null output");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
      cudaGetLastError();
      n = 0;
    }
    *out = n;
  });
}

int32_t ppca_b200_ctx_create(int32_t device, void *cuda_stream, ppca_b200_ctx **out) {
  return guarded([&] {
    REQUIRE(out != nullptr, "null output");
    int n = 0;
    CUDA_CHECK(cudaGetDeviceCount(&n));
    REQUIRE(device >= 0 && device < n, "device %d out of range (have %d)", device, n);
    CUDA_CHECK(cudaSetDevice(device));
    cudaDeviceProp prop;
    CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
      PPCA_THROW(PPCA_ERR_CUDA, "this engine is built for sm_100a only; device %d is sm_%d%d", device, prop.major,
                 prop.minor);
    std::unique_ptr<ppca_b200_ctx> ctx(new ppca_b200_ctx());
    ctx->device = device;
    ctx->sms = prop.multiProcessorCount;
    if (const char *e = getenv("PPCA_B200_GEMM"))
      ctx->gemm_mode = (strcmp(e, "int8") == 0) ? 1 : (strcmp(e, "dmma") == 0 ? 0 : 2);
    if (const char *e = getenv("PPCA_B200_SLICES")) {
      const int t = atoi(e);
      if ((t >= 6 && t <= 8) || (t == 4 && ctx->gemm_mode == 2)) ctx->slices = t;
    }
    ctx->base_mode = ctx->gemm_mode;
    ctx->base_slices = ctx->slices;
    if (const char *e = getenv("PPCA_B200_GUARD")) ctx->guard = atoi(e) != 0;
    if (const char *e = getenv("PPCA_B200_GUARD_BITS")) {
      const int b = atoi(e);
      if (b >= 8 && b <= 52) ctx->guard_bits = b;
    }
    if (cuda_stream) {
      ctx->stream = (cudaStream_t)cuda_stream;
      ctx->own_stream = false;
    } else {
      CUDA_CHECK(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
      ctx->own_stream = true;
    }
    *out = ctx.release();
  });
}

int32_t ppca_b200_ctx_destroy(ppca_b200_ctx *ctx) {
  return guarded([&] {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    for (auto e : ctx->ev_pool) cudaEventDestroy(e);
    for (auto &g : ctx->mix_graphs)
      if (g.exec) cudaGraphExecDestroy(g.exec);
    if (ctx->comm) {
      try {
        comm_destroy(ctx->comm);
      } catch (...) {
      }
      ctx->comm = nullptr;
    }
    if (ctx->pinned) cudaFreeHost(ctx->pinned);
    if (ctx->pin_ev) cudaEventDestroy(ctx->pin_ev);
    for (int b = 0; b < 2; ++b)
      if (ctx->upload_pin[b]) cudaFreeHost(ctx->upload_pin[b]);
    for (int b = 0; b < 2; ++b) {
      if (ctx->ev_copied[b]) cudaEventDestroy(ctx->ev_copied[b]);
      if (ctx->ev_free[b]) cudaEventDestroy(ctx->ev_free[b]);
    }
    for (int b = 0; b < 2; ++b) {
      if (ctx->ev_done[b]) cudaEventDestroy(ctx->ev_done[b]);
      if (ctx->ev_outfree[b]) cudaEventDestroy(ctx->ev_outfree[b]);
    }
    if (ctx->out_stream) cudaStreamDestroy(ctx->out_stream);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
  });
}

int32_t ppca_b200_ctx_synchronize(ppca_b200_ctx *ctx) {
  return guarded([&] {
    REQUIRE(ctx != nullptr, "null context");
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  });
}

int32_t ppca_b200_ctx_stream(ppca_b200_ctx *ctx, void **out) {
  return guarded([&] {
    REQUIRE(ctx != nullptr && out != nullptr, "null argument");
    *out = (void *)ctx->stream;
  });
}

int32_t ppca_b200_ctx_set_chunk(ppca_b200_ctx *ctx, int64_t chunk_samples) {
  return guarded([&] {
    REQUIRE(ctx != nullptr, "null context");
    REQUIRE(chunk_samples >= 0 && chunk_samples <= ((int64_t)1 << 30), "chunk out of range");
    ctx->chunk = chunk_samples;
  });
}

int32_t ppca_b200_ctx_set_gemm(ppca_b200_ctx *ctx, int32_t mode, int32_t slices) {
  return guarded([&] {
    REQUIRE(ctx != nullptr, "null context");
    REQUIRE(mode >= 0 && mode <= 2, "gemm mode must be 0 (dmma), 1 (int8 on mma.sync) or 2 (int8 on tcgen05)");
    REQUIRE((slices >= 6 && slices <= 8) || (slices == 4 && mode == 2),
            "slices must be 6, 7 or 8 (or 4, the FP32-class fast path, with the tcgen05 mode)");
    ctx->gemm_mode = ctx->base_mode = mode;
    ctx->slices = ctx->base_slices = slices;
    ctx->rung = ctx->rung_ttl = 0;
  });
}

int32_t ppca_b200_ctx_set_guard(ppca_b200_ctx *ctx, int32_t enabled, int32_t eps_bits) {
  return guarded([&] {
    REQUIRE(ctx != nullptr, "null context");
    REQUIRE(eps_bits == 0 || (eps_bits >= 8 && eps_bits <= 52), "eps_bits must be 0 (keep) or in [8, 52]");
    ctx->guard = enabled != 0;
    if (eps_bits) ctx->guard_bits = eps_bits;
    ctx->rung = ctx->rung_ttl = 0;
  });
}

int32_t ppca_b200_ctx_launch_count(ppca_b200_ctx *ctx, int64_t *out) {
  return guarded([&] {
    REQUIRE(ctx != nullptr && out != nullptr, "null argument");
    *out = ctx->launches;
  });
}

int32_t ppca_b200_ctx_variant_counts(ppca_b200_ctx *ctx, int64_t *out16) {
  return guarded([&] {
    REQUIRE(ctx != nullptr && out16 != nullptr, "null argument");
    for (int i = 0; i < V_COUNT; ++i) out16[i] = ctx->variants[i];
  });
}

int32_t ppca_b200_ctx_set_profiling(ppca_b200_ctx *ctx, int32_t enabled) {
  return guarded([&] {
    REQUIRE(ctx != nullptr, "null context");
    ctx->profiling = enabled != 0;
    ctx->profile_reset(true);  // spans accumulate from here until ppca_b200_ctx_last_profile reads them
  });
}

int32_t ppca_b200_ctx_last_profile(ppca_b200_ctx *ctx, double *out8) {
  return guarded([&] {
    REQUIRE(ctx != nullptr && out8 != nullptr, "null argument");
    DeviceGuard g(ctx->device);
    if (ctx->profiling) {  // one synchronisation here, none inside the profiled calls
      CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
      for (auto &sp : ctx->spans) {
        float ms = 0.f;
        CUDA_CHECK(cudaEventElapsedTime(&ms, sp.a, sp.b));
        ctx->last_profile[sp.fam] += ms;
      }
      ctx->spans.clear();
      ctx->ev_used = 0;
    }
    for (int i = 0; i < FAM_COUNT; ++i) out8[i] = ctx->last_profile[i];
  });
}

// ---- datasets -----------------------------------------------------------------------------------
int32_t ppca_b200_dataset_from_host(ppca_b200_ctx *ctx, const double *x, int64_t n, int32_t d, const double *weights,
                                    ppca_b200_dataset **out) {
  return guarded([&] {
    REQUIRE(ctx != nullptr && out != nullptr, "null argument");
    REQUIRE(n >= 0 && d >= 1, "bad dataset shape %lld x %d", (long long)n, d);
    REQUIRE(n == 0 || x != nullptr, "null data");
    DeviceGuard g(ctx->device);
    auto st = make_store(ctx, n, d);
    if (n > 0) {
      // Double-buffered upload: a few host threads copy the next block of rows into pinned memory while the
      // previous block is on the bus (pageable cudaMemcpy alone tops out near 6 GB/s).
      const int64_t rows_per = std::max<int64_t>(1, ((int64_t)32 << 20) / ((int64_t)d * 8));
      const size_t blk = (size_t)rows_per * d;
      if (ctx->upload_count < blk) {
        for (int b = 0; b < 2; ++b) {
          if (ctx->upload_pin[b]) cudaFreeHost(ctx->upload_pin[b]);
          ctx->upload_pin[b] = nullptr;
        }
        ctx->upload_count = 0;
        for (int b = 0; b < 2; ++b) CUDA_CHECK(cudaMallocHost((void **)&ctx->upload_pin[b], blk * sizeof(double)));
        ctx->upload_count = blk;
      }
      DevBuf<double> *raw = ctx->upload_raw;
      double **pin = ctx->upload_pin;
      cudaEvent_t done[2];
      for (int b = 0; b < 2; ++b) {
        raw[b].reserve(blk);
        CUDA_CHECK(cudaEventCreateWithFlags(&done[b], cudaEventDisableTiming));
      }
      auto cleanup = [&] {
        for (int b = 0; b < 2; ++b) cudaEventDestroy(done[b]);
      };
      // caller memory that is already page-locked (cudaHostRegister / cudaHostAlloc) is DMA-ed directly
      bool src_pinned = false;
      {
        cudaPointerAttributes pa;
        if (cudaPointerGetAttributes(&pa, x) == cudaSuccess) src_pinned = pa.type == cudaMemoryTypeHost;
        else cudaGetLastError();
      }
      try {
        int i = 0;
        for (int64_t r0 = 0; r0 < n; r0 += rows_per, ++i) {
          const int b = i & 1;
          const int64_t rows = std::min<int64_t>(rows_per, n - r0);
          if (i >= 2) CUDA_CHECK(cudaEventSynchronize(done[b]));  // the H2D that last used this pinned block
          const size_t bytes = (size_t)rows * d * sizeof(double);
          const char *src = reinterpret_cast<const char *>(x + r0 * d);
          if (src_pinned) {
            CUDA_CHECK(cudaMemcpyAsync(raw[b].p, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
            CUDA_CHECK(cudaEventRecord(done[b], ctx->stream));
            launch_ingest(ctx->L(), raw[b].p, rows, d, r0, *st);
            continue;
          }
          char *dst = reinterpret_cast<char *>(pin[b]);
          const int nthr = bytes >= ((size_t)4 << 20) ? 4 : 1;
          if (nthr == 1) {
            memcpy(dst, src, bytes);
          } else {
            std::vector<std::thread> th;
            const size_t per = (bytes / nthr + 63) & ~(size_t)63;
            for (int t = 0; t < nthr; ++t) {
              const size_t lo = std::min(bytes, per * t), hi = std::min(bytes, per * (t + 1));
              if (hi > lo) th.emplace_back([=] { memcpy(dst + lo, src + lo, hi - lo); });
            }
            for (auto &t : th) t.join();
          }
          CUDA_CHECK(cudaMemcpyAsync(raw[b].p, pin[b], bytes, cudaMemcpyHostToDevice, ctx->stream));
          CUDA_CHECK(cudaEventRecord(done[b], ctx->stream));
          launch_ingest(ctx->L(), raw[b].p, rows, d, r0, *st);
        }
        launch_transpose_mask(ctx->L(), *st);
        CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
      } catch (...) {
        cudaStreamSynchronize(ctx->stream);
        cleanup();
        throw;
      }
      cleanup();
    }
    *out = make_dataset(ctx, st, weights);
  });
}

// smallest weight of a device array, NaN -> -1 (mixture EM rejects non-positive weights, mix.rs:304-309)
__global__ void min_weight_kernel(const double *__restrict__ w, int64_t n, unsigned long long *out) {
  double mn = __longlong_as_double(0x7ff0000000000000LL);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const double v = w[i];
    mn = (v != v) ? -1.0 : fmin(mn, v);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o));
  // order-preserving map of a double onto an unsigned integer so atomicMin works for negative values too
  if ((threadIdx.x & 31) == 0) {
    long long b = __double_as_longlong(mn);
    unsigned long long key = b < 0 ? ~(unsigned long long)b : ((unsigned long long)b | 0x8000000000000000ULL);
    atomicMin(out, key);
  }
}

static double device_min_weight(ppca_b200_ctx *ctx, const double *w_dev, int64_t n) {
  if (n <= 0) return std::numeric_limits<double>::infinity();
  DevBuf<unsigned long long> key;
  key.alloc(1);
  CUDA_CHECK(cudaMemsetAsync(key.p, 0xff, sizeof(unsigned long long), ctx->stream));
  const int64_t want = (n + 255) / 256;
  const int blocks = (int)std::min<int64_t>(want, (int64_t)ctx->sms * 8);
  min_weight_kernel<<<blocks, 256, 0, ctx->stream>>>(w_dev, n, key.p);
  CUDA_CHECK(cudaGetLastError());
  ++ctx->launches;
  unsigned long long h = 0;
  CUDA_CHECK(cudaMemcpyAsync(&h, key.p, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  const unsigned long long bits = (h & 0x8000000000000000ULL) ? (h & 0x7fffffffffffffffULL) : ~h;
  double v;
  memcpy(&v, &bits, sizeof(v));
  return v;
}

// Dataset::new from a matrix that already lives on the device (zero-copy ingestion for DLPack / __cuda_array_interface__
// producers; the reference copies numpy -> Rust element by element, src/python_bindings.rs:41-54).  `x` is row-major with
// `row_stride` doubles between rows; `producer_stream` is the stream the caller last wrote x / weights on (the ingest is
// ordered after it by an event; 0 = legacy default stream).
int32_t ppca_b200_dataset_from_device(ppca_b200_ctx *ctx, const double *x, int64_t n, int32_t d, int64_t row_stride,
                                      const double *weights, void *producer_stream, ppca_b200_dataset **out) {
  return guarded([&] {
    REQUIRE(ctx != nullptr && out != nullptr, "null argument");
    REQUIRE(n >= 0 && d >= 1, "bad dataset shape %lld x %d", (long long)n, d);
    REQUIRE(n == 0 || x != nullptr, "null data");
    REQUIRE(row_stride >= d, "row stride %lld is shorter than a row of %d values", (long long)row_stride, d);
    DeviceGuard g(ctx->device);
    auto on_device = [&](const void *p, const char *what) {
      cudaPointerAttributes pa;
      const cudaError_t e = cudaPointerGetAttributes(&pa, p);
      if (e != cudaSuccess) cudaGetLastError();
      REQUIRE(e == cudaSuccess && (pa.type == cudaMemoryTypeDevice || pa.type == cudaMemoryTypeManaged),
              "%s is not a device pointer", what);
      REQUIRE(pa.type == cudaMemoryTypeManaged || pa.device == ctx->device, "%s lives on device %d, the context on %d", what,
              pa.device, ctx->device);
    };
    if (n > 0) on_device(x, "data");
    if (n > 0 && weights) on_device(weights, "weights");
    {  // order the ingest after the producer's work
      cudaEvent_t ev;
      CUDA_CHECK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
      cudaError_t e = cudaEventRecord(ev, (cudaStream_t)producer_stream);
      if (e == cudaSuccess) e = cudaStreamWaitEvent(ctx->stream, ev, 0);
      cudaEventDestroy(ev);
      CUDA_CHECK(e);
    }
    auto st = make_store(ctx, n, d);
    if (n > 0) {
      if (row_stride == d) {
        launch_ingest(ctx->L(), x, n, d, 0, *st);
      } else {  // strided rows: compact block by block, then ingest
        const int64_t rows_per = std::max<int64_t>(1, ((int64_t)64 << 20) / ((int64_t)d * 8));
        DevBuf<double> buf;
        buf.alloc((size_t)std::min<int64_t>(rows_per, n) * d);
        for (int64_t r0 = 0; r0 < n; r0 += rows_per) {
          const int64_t rows = std::min<int64_t>(rows_per, n - r0);
          CUDA_CHECK(cudaMemcpy2DAsync(buf.p, sizeof(double) * d, x + r0 * row_stride, sizeof(double) * row_stride,
                                       sizeof(double) * d, (size_t)rows, cudaMemcpyDeviceToDevice, ctx->stream));
          launch_ingest(ctx->L(), buf.p, rows, d, r0, *st);
        }
        CUDA_CHECK(cudaStreamSynchronize(ctx->stream));  // buf is released at the end of this scope
      }
      launch_transpose_mask(ctx->L(), *st);
    }
    std::unique_ptr<ppca_b200_dataset> nd(make_dataset(ctx, st, nullptr));
    if (n > 0 && weights) {
      CUDA_CHECK(cudaMemcpyAsync(nd->w.p, weights, sizeof(double) * n, cudaMemcpyDeviceToDevice, ctx->stream));
      nd->min_w = device_min_weight(ctx, nd->w.p, n);
    }
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    *out = nd.release();
  });
}

// Dataset::numpy into caller-provided DEVICE memory (n x d doubles, row-major, NaN at the masked slots): the export side of
// the zero-copy path.  Ordered on the context's stream; returns after the copy has completed.
int32_t ppca_b200_dataset_to_device(ppca_b200_ctx *ctx, const ppca_b200_dataset *ds, int64_t row0, int64_t nrows,
                                    double *out_dev) {
  return guarded([&] {
    check_ds(ctx, ds);
    const SampleStore &st = *ds->store;
    REQUIRE(row0 >= 0 && nrows >= 0 && row0 + nrows <= st.n, "row range out of bounds");
    if (nrows == 0) return;
    REQUIRE(out_dev != nullptr, "null output");
    DeviceGuard g(ctx->device);
    launch_export(ctx->L(), st, row0, nrows, out_dev);
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  });
}

static void dataset_synthetic_impl(ppca_b200_ctx *ctx, int64_t row_begin, int64_t n, int32_t d, int32_t k_true,
                                   double sigma_true, double mask_prob, int32_t n_components, uint64_t seed,
                                   ppca_b200_dataset **out) {
  REQUIRE(ctx != nullptr && out != nullptr, "null argument");
  REQUIRE(n >= 1 && d >= 1 && k_true >= 1 && n_components >= 1 && row_begin >= 0, "bad synthetic shape");
  REQUIRE(mask_prob >= 0.0 && mask_prob <= 1.0, "invalid mask probability");
  DeviceGuard g(ctx->device);
  auto st = make_store(ctx, n, d);
  launch_synthetic(ctx->L(), *st, k_true, sigma_true, mask_prob, n_components, seed, row_begin);
  launch_transpose_mask(ctx->L(), *st);
  CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  *out = make_dataset(ctx, st, nullptr);
}

int32_t ppca_b200_dataset_synthetic(ppca_b200_ctx *ctx, int64_t n, int32_t d, int32_t k_true, double sigma_true,
                                    double mask_prob, int32_t n_components, uint64_t seed, ppca_b200_dataset **out) {
  return guarded([&] { dataset_synthetic_impl(ctx, 0, n, d, k_true, sigma_true, mask_prob, n_components, seed, out); });
}

// rows [row_begin, row_begin + n) of the synthetic dataset: every counter of the generator is keyed by the global row, so
// the ranks of a sharded job hold disjoint row ranges of ONE dataset (same truth) when they pass the same seed
int32_t ppca_b200_dataset_synthetic_rows(ppca_b200_ctx *ctx, int64_t row_begin, int64_t n, int32_t d, int32_t k_true,
                                         double sigma_true, double mask_prob, int32_t n_components, uint64_t seed,
                                         ppca_b200_dataset **out) {
  return guarded([&] { dataset_synthetic_impl(ctx, row_begin, n, d, k_true, sigma_true, mask_prob, n_components, seed, out); });
}

int32_t ppca_b200_model_sample(ppca_b200_ctx *ctx, int64_t n, int32_t d, int32_t k, const double *C, const double *mu,
                                double sigma, double mask_prob, uint64_t seed, ppca_b200_dataset **out) {
  return guarded([&] {
    REQUIRE(ctx != nullptr && out != nullptr && C != nullptr && mu != nullptr, "null argument");
    REQUIRE(n >= 0 && d >= 1 && k >= 1, "bad sample shape");
    REQUIRE(mask_prob >= 0.0 && mask_prob <= 1.0, "invalid mask probability");  // ppca_model.rs:165
    REQUIRE(sigma >= 0.0 && std::isfinite(sigma), "isotropic_noise must be finite and non-negative");
    DeviceGuard g(ctx->device);
    auto st = make_store(ctx, n, d);
    if (n > 0) {
      ctx->Cdense.reserve((size_t)d * k);
      ctx->mudense.reserve((size_t)d);
      double *stage = ctx->pin((size_t)d * k + d);
      ctx->pin_wait();
      memcpy(stage, C, sizeof(double) * d * k);
      memcpy(stage + (size_t)d * k, mu, sizeof(double) * d);
      CUDA_CHECK(cudaMemcpyAsync(ctx->Cdense.p, stage, sizeof(double) * d * k, cudaMemcpyHostToDevice, ctx->stream));
      CUDA_CHECK(cudaMemcpyAsync(ctx->mudense.p, stage + (size_t)d * k, sizeof(double) * d, cudaMemcpyHostToDevice,
                                 ctx->stream));
      launch_model_sample(ctx->L(), *st, k, ctx->Cdense.p, ctx->mudense.p, sigma, mask_prob, seed);
      launch_transpose_mask(ctx->L(), *st);
    }
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    *out = make_dataset(ctx, st, nullptr);
  });
}

// Shared body of ppca_b200_mix_sample / ppca_b200_posterior_sample: stage the models, run sample_general_kernel.
static void sample_general_impl(ppca_b200_ctx *ctx, int64_t n, int32_t d, int32_t m, const int32_t *ks, const double *Cs,
                                const double *mus, const double *sigmas, const double *cdf_host, const double *post_host,
                                const double *const *states_host, const double *const *covs_host, double mask_prob,
                                uint64_t seed, ppca_b200_dataset **out) {
  REQUIRE(ctx != nullptr && out != nullptr && ks && Cs && mus && sigmas, "null argument");
  REQUIRE(n >= 0 && d >= 1 && m >= 1, "bad sample shape");
  REQUIRE(mask_prob >= 0.0 && mask_prob <= 1.0, "invalid mask probability");  // ppca_model.rs:165
  int kmax = 1;
  std::vector<int64_t> coff((size_t)m);
  int64_t ctot = 0;
  for (int j = 0; j < m; ++j) {
    REQUIRE(ks[j] >= 1, "state_size must be >= 1 (got %d)", ks[j]);
    REQUIRE(sigmas[j] >= 0.0 && std::isfinite(sigmas[j]), "isotropic_noise must be finite and non-negative");
    kmax = std::max(kmax, (int)ks[j]);
    coff[j] = ctot;
    ctot += (int64_t)d * ks[j];
  }
  const bool post_mode = states_host != nullptr;
  REQUIRE(!post_mode || covs_host != nullptr, "posterior sampling needs states and covariances");
  DeviceGuard g(ctx->device);
  auto st = make_store(ctx, n, d);
  if (n > 0) {
    DevBuf<double> dC, dmu, dsig, dcdf, dpost;
    DevBuf<int> dks, dfail;
    DevBuf<int64_t> dcoff;
    DevBuf<const double *> dstates, dcovs;
    std::vector<std::unique_ptr<DevBuf<double>>> sbuf, cbuf;
    auto up = [&](auto &buf, const auto *src, size_t count) {
      buf.alloc(count);
      CUDA_CHECK(cudaMemcpyAsync(buf.p, src, sizeof(*src) * count, cudaMemcpyHostToDevice, ctx->stream));
    };
    up(dC, Cs, (size_t)ctot);
    up(dmu, mus, (size_t)m * d);
    up(dsig, sigmas, (size_t)m);
    up(dks, ks, (size_t)m);
    up(dcoff, coff.data(), (size_t)m);
    if (cdf_host) up(dcdf, cdf_host, (size_t)m);
    if (post_host) up(dpost, post_host, (size_t)n * m);
    std::vector<const double *> sp((size_t)m, nullptr), cp((size_t)m, nullptr);
    if (post_mode) {
      for (int j = 0; j < m; ++j) {
        REQUIRE(states_host[j] && covs_host[j], "null states / covariances of component %d", j);
        sbuf.emplace_back(new DevBuf<double>());
        cbuf.emplace_back(new DevBuf<double>());
        up(*sbuf[j], states_host[j], (size_t)n * ks[j]);
        up(*cbuf[j], covs_host[j], (size_t)n * ks[j] * ks[j]);
        sp[j] = sbuf[j]->p;
        cp[j] = cbuf[j]->p;
      }
      up(dstates, sp.data(), (size_t)m);
      up(dcovs, cp.data(), (size_t)m);
    }
    dfail.alloc(1);
    CUDA_CHECK(cudaMemsetAsync(dfail.p, 0, sizeof(int), ctx->stream));
    SamplerArgs a;
    a.n = n;
    a.d = d;
    a.m = m;
    a.kmax = kmax;
    a.ks = dks.p;
    a.coff = dcoff.p;
    a.Cs = dC.p;
    a.mus = dmu.p;
    a.sigmas = dsig.p;
    a.cdf = cdf_host ? dcdf.p : nullptr;
    a.post = post_host ? dpost.p : nullptr;
    a.post_mode = post_mode ? 1 : 0;
    a.states = post_mode ? dstates.p : nullptr;
    a.covs = post_mode ? dcovs.p : nullptr;
    a.mask_prob = mask_prob;
    a.seed = seed;
    a.fail = dfail.p;
    launch_sample_general(ctx->L(), *st, a);
    launch_transpose_mask(ctx->L(), *st);
    int fail = 0;
    CUDA_CHECK(cudaMemcpyAsync(&fail, dfail.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));  // the staging buffers above are released on return
    if (fail) PPCA_THROW(PPCA_ERR_NUMERIC, "Cholesky decomposition failed (ppca_model.rs:582-586)");
  }
  *out = make_dataset(ctx, st, nullptr);
}

int32_t ppca_b200_mix_sample(ppca_b200_ctx *ctx, int64_t n, int32_t d, int32_t m, const int32_t *ks, const double *Cs,
                             const double *mus, const double *sigmas, const double *log_weights, double mask_prob,
                             uint64_t seed, ppca_b200_dataset **out) {
  return guarded([&] {
    REQUIRE(log_weights != nullptr && m >= 1, "null mixture parameters");
    // WeightedIndex over exp(log_weights) (mix.rs:124-126): cumulative distribution, last entry exactly 1
    std::vector<double> cdf((size_t)m);
    double mx = log_weights[0], tot = 0.0;
    for (int j = 1; j < m; ++j) mx = std::max(mx, log_weights[j]);
    for (int j = 0; j < m; ++j) tot += std::exp(log_weights[j] - mx);
    REQUIRE(tot > 0.0 && std::isfinite(tot), "can create WeightedIndex from distribution (mix.rs:125)");
    double acc = 0.0;
    for (int j = 0; j < m; ++j) {
      acc += std::exp(log_weights[j] - mx) / tot;
      cdf[j] = acc;
    }
    cdf[(size_t)m - 1] = 1.0;
    sample_general_impl(ctx, n, d, m, ks, Cs, mus, sigmas, cdf.data(), nullptr, nullptr, nullptr, mask_prob, seed, out);
  });
}

int32_t ppca_b200_posterior_sample(ppca_b200_ctx *ctx, int64_t n, int32_t d, int32_t m, const int32_t *ks, const double *Cs,
                                   const double *mus, const double *sigmas, const double *posteriors,
                                   const double *const *states, const double *const *covariances, uint64_t seed,
                                   ppca_b200_dataset **out) {
  return guarded([&] {
    REQUIRE(states != nullptr && covariances != nullptr, "null states / covariances");
    REQUIRE(m == 1 || posteriors != nullptr, "a mixture posterior sampler needs the posterior probabilities");
    sample_general_impl(ctx, n, d, m, ks, Cs, mus, sigmas, nullptr, m > 1 ? posteriors : nullptr, states, covariances, 0.0,
                        seed, out);
  });
}

int32_t ppca_b200_dataset_with_weights(ppca_b200_ctx *ctx, const ppca_b200_dataset *ds, const double *weights,
                                       ppca_b200_dataset **out) {
  return guarded([&] {
    check_ds(ctx, ds);
    REQUIRE(out != nullptr && (weights != nullptr || ds->store->n == 0), "null argument");
    DeviceGuard g(ctx->device);
    *out = make_dataset(ctx, ds->store, weights);
  });
}

int32_t ppca_b200_dataset_len(const ppca_b200_dataset *ds, int64_t *out) {
  return guarded([&] {
    REQUIRE(ds != nullptr && out != nullptr, "null argument");
    *out = ds->store->n;
  });
}

int32_t ppca_b200_dataset_output_size(const ppca_b200_dataset *ds, int32_t *out) {
  return guarded([&] {
    REQUIRE(ds != nullptr && out != nullptr, "null argument");
    *out = ds->store->d;
  });
}

int32_t ppca_b200_dataset_to_host(ppca_b200_ctx *ctx, const ppca_b200_dataset *ds, int64_t row0, int64_t nrows,
                                  double *out) {
  return guarded([&] {
    check_ds(ctx, ds);
    const SampleStore &st = *ds->store;
    REQUIRE(row0 >= 0 && nrows >= 0 && row0 + nrows <= st.n, "row range out of bounds");
    if (nrows == 0) return;
    REQUIRE(out != nullptr, "null output");
    DeviceGuard g(ctx->device);
    const int64_t rows_per = std::max<int64_t>(1, ((int64_t)64 << 20) / ((int64_t)st.d * 8));
    DevBuf<double> buf;
    buf.alloc((size_t)std::min<int64_t>(rows_per, nrows) * st.d);
    for (int64_t r0 = 0; r0 < nrows; r0 += rows_per) {
      const int64_t rows = std::min<int64_t>(rows_per, nrows - r0);
      launch_export(ctx->L(), st, row0 + r0, rows, buf.p);
      CUDA_CHECK(cudaMemcpyAsync(out + r0 * st.d, buf.p, sizeof(double) * rows * st.d, cudaMemcpyDeviceToHost,
                                 ctx->stream));
      CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    }
  });
}

int32_t ppca_b200_dataset_weights(ppca_b200_ctx *ctx, const ppca_b200_dataset *ds, double *out) {
  return guarded([&] {
    check_ds(ctx, ds);
    if (ds->store->n == 0) return;
    REQUIRE(out != nullptr, "null output");
    DeviceGuard g(ctx->device);
    CUDA_CHECK(cudaMemcpyAsync(out, ds->w.p, sizeof(double) * ds->store->n, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  });
}

int32_t ppca_b200_dataset_empty_dimensions(ppca_b200_ctx *ctx, const ppca_b200_dataset *ds, uint8_t *out) {
  return guarded([&] {
    check_ds(ctx, ds);
    REQUIRE(out != nullptr, "null output");
    const SampleStore &st = *ds->store;
    if (st.n == 0) {  // dataset.rs:195-197: no first sample -> empty list
      memset(out, 0, st.d);
      return;
    }
    DeviceGuard g(ctx->device);
    DevBuf<uint8_t> buf;
    buf.alloc(st.d);
    launch_empty_dims(ctx->L(), st, buf.p);
    CUDA_CHECK(cudaMemcpyAsync(out, buf.p, st.d, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  });
}

int32_t ppca_b200_dataset_slice(ppca_b200_ctx *ctx, const ppca_b200_dataset *ds, int64_t row0, int64_t nrows,
                                ppca_b200_dataset **out) {
  return guarded([&] {
    check_ds(ctx, ds);
    const SampleStore &src = *ds->store;
    REQUIRE(out != nullptr, "null output");
    REQUIRE(row0 >= 0 && nrows >= 0 && row0 + nrows <= src.n, "row range out of bounds");
    DeviceGuard g(ctx->device);
    auto st = make_store(ctx, nrows, src.d);
    launch_copy_rows(ctx->L(), src, row0, nrows, *st, 0);
    launch_transpose_mask(ctx->L(), *st);
    std::unique_ptr<ppca_b200_dataset> nd(make_dataset(ctx, st, nullptr));
    if (nrows > 0)
      CUDA_CHECK(cudaMemcpyAsync(nd->w.p, ds->w.p + row0, sizeof(double) * nrows, cudaMemcpyDeviceToDevice, ctx->stream));
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    // smallest weight OF THE SLICE (mixture EM rejects non-positive weights, mix.rs:304-309): the parent's minimum
    // may sit outside this row range
    double mn = std::numeric_limits<double>::infinity();
    if (nrows > 0) {
      std::vector<double> hw((size_t)nrows);
      CUDA_CHECK(cudaMemcpy(hw.data(), ds->w.p + row0, sizeof(double) * nrows, cudaMemcpyDeviceToHost));
      for (int64_t i = 0; i < nrows; ++i) mn = hw[i] < mn ? hw[i] : (hw[i] != hw[i] ? -1.0 : mn);
    }
    nd->min_w = mn;
    *out = nd.release();
  });
}

int32_t ppca_b200_dataset_concat(ppca_b200_ctx *ctx, const ppca_b200_dataset *const *list, int32_t count,
                                 ppca_b200_dataset **out) {
  return guarded([&] {
    REQUIRE(ctx != nullptr && out != nullptr, "null argument");
    REQUIRE(count >= 1 && list != nullptr, "concat needs at least one dataset");
    int64_t total = 0;
    const int d = list[0] && list[0]->store ? list[0]->store->d : 0;
    for (int i = 0; i < count; ++i) {
      check_ds(ctx, list[i]);
      REQUIRE(list[i]->store->d == d, "output sizes differ in concat");
      total += list[i]->store->n;
    }
    DeviceGuard g(ctx->device);
    auto st = make_store(ctx, total, d);
    std::unique_ptr<ppca_b200_dataset> nd(make_dataset(ctx, st, nullptr));
    int64_t pos = 0;
    double mn = std::numeric_limits<double>::infinity();
    for (int i = 0; i < count; ++i) {
      const SampleStore &src = *list[i]->store;
      launch_copy_rows(ctx->L(), src, 0, src.n, *st, pos);
      if (src.n > 0)
        CUDA_CHECK(cudaMemcpyAsync(nd->w.p + pos, list[i]->w.p, sizeof(double) * src.n, cudaMemcpyDeviceToDevice,
                                   ctx->stream));
      if (src.n > 0 && list[i]->min_w < mn) mn = list[i]->min_w;
      pos += src.n;
    }
    nd->min_w = mn;
    launch_transpose_mask(ctx->L(), *st);
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    *out = nd.release();
  });
}

int32_t ppca_b200_dataset_destroy(ppca_b200_dataset *ds) {
  return guarded([&] {
    if (!ds) return;
    cudaSetDevice(ds->device);
    delete ds;
  });
}

// ---- PPCAModel ------------------------------------------------------------------------------------
int32_t ppca_b200_llks(ppca_b200_ctx *ctx, const ppca_b200_dataset *ds, int32_t k, const double *C, const double *mu,
                       double sigma, double *out) {
  return guarded([&] {
    check_ds(ctx, ds);
    const SampleStore &st = *ds->store;
    if (st.n == 0) return;
    REQUIRE(out != nullptr, "null output");
    DeviceGuard g(ctx->device);
    run_guarded(ctx, [&] {
      DevModel m = stage_model(ctx, st.d, k, C, mu, sigma);
      ctx->rbuf.reserve((size_t)st.n_pad);
      llks_impl(ctx, st, m, ctx->rbuf.p);
      CUDA_CHECK(cudaMemcpyAsync(out, ctx->rbuf.p, sizeof(double) * st.n, cudaMemcpyDeviceToHost, ctx->stream));
      CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
      return read_unsafe(ctx);
    });
  });
}

int32_t ppca_b200_llk(ppca_b200_ctx *ctx, const ppca_b200_dataset *ds, int32_t k, const double *C, const double *mu,
                      double sigma, double *out) {
  return guarded([&] {
    check_ds(ctx, ds);
    REQUIRE(out != nullptr, "null output");
    const SampleStore &st = *ds->store;
    if (st.n == 0) {
      *out = 0.0;
      return;
    }
    DeviceGuard g(ctx->device);
    run_guarded(ctx, [&] {
      DevModel m = stage_model(ctx, st.d, k, C, mu, sigma);
      const int64_t chunk = pick_chunk(ctx, st.n_pad, m.s);
      reserve_chunk_ws(ctx, chunk, m.s);
      ctx->stats.reserve(8);
      ctx->part_solve.reserve((size_t)SOLVE_SLOTS * 4);
      CUDA_CHECK(cudaMemsetAsync(ctx->stats.p, 0, sizeof(double) * 8, ctx->stream));
      CUDA_CHECK(cudaMemsetAsync(ctx->part_solve.p, 0, sizeof(double) * SOLVE_SLOTS * 4, ctx->stream));
      for (int64_t row0 = 0; row0 < st.n; row0 += chunk) {
        const int rows = (int)((st.n - row0) < chunk ? (st.n - row0) : chunk);
        e_step_chunk(ctx, st, ds->w.p, row0, rows, m, 0, ctx->llk.p, nullptr, ctx->part_solve.p);
      }
      launch_solve_finish(ctx->L(), ctx->part_solve.p, ctx->stats.p);
      double h[8];
      CUDA_CHECK(cudaMemcpyAsync(h, ctx->stats.p, sizeof(double) * 8, cudaMemcpyDeviceToHost, ctx->stream));
      CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
      *out = h[SC_LLK];
      return read_unsafe(ctx);
    });
  });
}

int32_t ppca_b200_infer(ppca_b200_ctx *ctx, const ppca_b200_dataset *ds, int32_t k, const double *C, const double *mu,
                        double sigma, double *states, double *covariances) {
  return guarded([&] {
    check_ds(ctx, ds);
    const SampleStore &st = *ds->store;
    if (st.n == 0) return;
    REQUIRE(states != nullptr, "null output");
    DeviceGuard g(ctx->device);
    run_guarded(ctx, [&] {
    DevModel m = stage_model(ctx, st.d, k, C, mu, sigma);
    int64_t chunk = pick_chunk(ctx, st.n_pad, m.s);
    if (covariances) {  // bound the k x k covariance staging buffer to ~1 GiB
      const int64_t cap = round_up(std::max<int64_t>(256, ((int64_t)1 << 30) / ((int64_t)k * k * 8)), 256);
      if (chunk > cap) chunk = cap;
    }
    reserve_chunk_ws(ctx, chunk, m.s);
    if (covariances) ctx->cov.reserve((size_t)chunk * k * k);
    ctx->rbuf.reserve((size_t)chunk * k);
    for (int64_t row0 = 0; row0 < st.n; row0 += chunk) {
      const int rows = (int)((st.n - row0) < chunk ? (st.n - row0) : chunk);
      e_step_chunk(ctx, st, nullptr, row0, rows, m, 1, nullptr, covariances ? ctx->cov.p : nullptr, nullptr);
      const int64_t total = (int64_t)rows * k;
      const int blocks = (int)((total + 255) / 256 < (int64_t)ctx->sms * 8 ? (total + 255) / 256 : (int64_t)ctx->sms * 8);
      strided_copy_kernel<<<blocks, 256, 0, ctx->stream>>>(ctx->YZ.p, rows, k, m.s.kp, ctx->rbuf.p, k);
      CUDA_CHECK(cudaGetLastError());
      ++ctx->launches;
      CUDA_CHECK(cudaMemcpyAsync(states + row0 * k, ctx->rbuf.p, sizeof(double) * rows * k, cudaMemcpyDeviceToHost,
                                 ctx->stream));
      if (covariances)
        CUDA_CHECK(cudaMemcpyAsync(covariances + row0 * k * k, ctx->cov.p, sizeof(double) * rows * k * k,
                                   cudaMemcpyDeviceToHost, ctx->stream));
      CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    }
    return read_unsafe(ctx);
    });
  });
}

static int32_t smooth_or_extrapolate(ppca_b200_ctx *ctx, const ppca_b200_dataset *ds, int32_t k, const double *C,
                                     const double *mu, double sigma, int extrapolate, ppca_b200_dataset **out) {
  return guarded([&] {
    check_ds(ctx, ds);
    REQUIRE(out != nullptr, "null output");
    const SampleStore &st = *ds->store;
    DeviceGuard g(ctx->device);
    auto ost = make_store(ctx, st.n, st.d);
    if (st.n > 0) {
      run_guarded(ctx, [&] {
        DevModel m = stage_model(ctx, st.d, k, C, mu, sigma);
        reconstruct_impl(ctx, st, m, extrapolate, nullptr, 0, 0, *ost);
        return read_unsafe(ctx);
      });
      finalize_full_store(ctx, *ost);
    }
    ost->full = true;
    std::unique_ptr<ppca_b200_dataset> nd(make_dataset(ctx, ost, nullptr));
    if (st.n > 0)  // weights are carried through (ppca_model.rs:242,259)
      CUDA_CHECK(cudaMemcpyAsync(nd->w.p, ds->w.p, sizeof(double) * st.n, cudaMemcpyDeviceToDevice, ctx->stream));
    nd->min_w = ds->min_w;
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    *out = nd.release();
  });
}

int32_t ppca_b200_reconstruct(ppca_b200_ctx *ctx, const ppca_b200_dataset *ds, int32_t k, const double *C,
                              const double *mu, double sigma, int32_t extrapolate, ppca_b200_dataset *reuse,
                              double *llks, ppca_b200_dataset **out) {
  return guarded([&] {
    check_ds(ctx, ds);
    REQUIRE(out != nullptr, "null output");
    const SampleStore &st = *ds->store;
    if (reuse) {
      check_ds(ctx, reuse);
      REQUIRE(reuse != ds && reuse->store != ds->store, "the output dataset must not share samples with the input");
      REQUIRE(reuse->store->n == st.n && reuse->store->d == st.d, "the reused output dataset has another shape");
      REQUIRE(reuse->store->full, "the reused output must be a dataset produced by smooth / extrapolate (all observed)");
    }
    DeviceGuard g(ctx->device);
    std::shared_ptr<SampleStore> ost = reuse ? reuse->store : make_store(ctx, st.n, st.d);
    if (st.n > 0) {
      if (llks) ctx->rbuf.reserve((size_t)st.n_pad);
      run_guarded(ctx, [&] {
        DevModel m = stage_model(ctx, st.d, k, C, mu, sigma);
        reconstruct_impl(ctx, st, m, extrapolate, nullptr, 0, 0, *ost, llks ? ctx->rbuf.p : nullptr);
        return read_unsafe(ctx);
      });
      if (!reuse) finalize_full_store(ctx, *ost);
      if (llks) CUDA_CHECK(cudaMemcpyAsync(llks, ctx->rbuf.p, sizeof(double) * st.n, cudaMemcpyDeviceToHost, ctx->stream));
    }
    ost->full = true;
    ppca_b200_dataset *nd = reuse ? reuse : make_dataset(ctx, ost, nullptr);
    if (st.n > 0)  // weights are carried through (ppca_model.rs:242,259)
      CUDA_CHECK(cudaMemcpyAsync(nd->w.p, ds->w.p, sizeof(double) * st.n, cudaMemcpyDeviceToDevice, ctx->stream));
    nd->min_w = ds->min_w;
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    *out = nd;
  });
}

int32_t ppca_b200_smooth(ppca_b200_ctx *ctx, const ppca_b200_dataset *ds, int32_t k, const double *C, const double *mu,
                         double sigma, ppca_b200_dataset **out) {
  return smooth_or_extrapolate(ctx, ds, k, C, mu, sigma, 0, out);
}
int32_t ppca_b200_extrapolate(ppca_b200_ctx *ctx, const ppca_b200_dataset *ds, int32_t k, const double *C,
                              const double *mu, double sigma, ppca_b200_dataset **out) {
  return smooth_or_extrapolate(ctx, ds, k, C, mu, sigma, 1, out);
}

int64_t ppca_b200_em_stats_len(int32_t d, int32_t k) {
  if (d < 1 || k < 1) return 0;
  return StatsLayout(d, k).len;
}

int32_t ppca_b200_em_stats(ppca_b200_ctx *ctx, const ppca_b200_dataset *ds, int32_t k, const double *C,
                           const double *mu, double sigma, double *stats_dev) {
  return guarded([&] {
    check_ds(ctx, ds);
    REQUIRE(stats_dev != nullptr, "null statistics buffer");
    DeviceGuard g(ctx->device);
    begin_pass(ctx);
    DevModel m = stage_model(ctx, ds->store->d, k, C, mu, sigma);
    em_stats_impl(ctx, *ds->store, ds->w.p, m, stats_dev);
  });
}

int32_t ppca_b200_em_finish(ppca_b200_ctx *ctx, int32_t d, int32_t k, const double *C, const double *mu, double sigma,
                            const ppca_b200_prior *prior, const double *stats_dev, double *C_out, double *mu_out,
                            double *sigma_out, double *llk_in) {
  return guarded([&] {
    REQUIRE(ctx != nullptr && stats_dev != nullptr, "null argument");
    REQUIRE(d >= 1 && k >= 1, "bad shape");
    DeviceGuard g(ctx->device);
    const double viol = em_finish_impl(ctx, d, k, C, mu, sigma, prior, stats_dev, C_out, mu_out, sigma_out, llk_in, nullptr);
    if (!end_pass(ctx, viol))
      PPCA_THROW(PPCA_ERR_PRECISION,
                 "precision guard: %.0f entries of the int8-sliced contractions lost accuracy against their column "
                 "scale; the context now uses a wider arithmetic - repeat em_stats, the all-reduce and em_finish",
                 viol);
  });
}

int32_t ppca_b200_iterate(ppca_b200_ctx *ctx, const ppca_b200_dataset *ds, int32_t k, const double *C,
                          const double *mu, double sigma, const ppca_b200_prior *prior, double *C_out, double *mu_out,
                          double *sigma_out, double *llk_in) {
  return guarded([&] {
    check_ds(ctx, ds);
    const SampleStore &st = *ds->store;
    if (st.n == 0) PPCA_THROW(PPCA_ERR_EMPTY, "non-empty dataset required (ppca_model.rs:358)");
    DeviceGuard g(ctx->device);
    run_guarded(ctx, [&] {
      DevModel m = stage_model(ctx, st.d, k, C, mu, sigma);
      ctx->stats.reserve((size_t)StatsLayout(st.d, k).len);
      em_stats_impl(ctx, st, ds->w.p, m, ctx->stats.p);
      return em_finish_impl(ctx, st.d, k, C, mu, sigma, prior, ctx->stats.p, C_out, mu_out, sigma_out, llk_in, nullptr);
    });
  });
}

// ---- out-of-core EM step: samples stay in host memory -------------------------------------------------
int32_t ppca_b200_host_register(const void *p, uint64_t bytes) {
  return guarded([&] {
    REQUIRE(p != nullptr && bytes > 0, "null host range");
    const cudaError_t e = cudaHostRegister(const_cast<void *>(p), (size_t)bytes, cudaHostRegisterPortable);
    if (e != cudaSuccess) {
      cudaGetLastError();  // not sticky: clear it so the next launch check does not see a stale error
      PPCA_THROW(PPCA_ERR_CUDA, "cudaHostRegister failed: %s (the range stays pageable)", cudaGetErrorString(e));
    }
  });
}

int32_t ppca_b200_host_unregister(const void *p) {
  return guarded([&] {
    REQUIRE(p != nullptr, "null host range");
    CUDA_CHECK(cudaHostUnregister(const_cast<void *>(p)));
  });
}

}  // extern "C"

namespace {
// streams x (host) through the device block by block and accumulates the statistics into stats_dev
// compact host format of a shard (see ppca_b200_iterate_packed_host); null = the plain n x d matrix `x`
struct PackedHost {
  const double *vals = nullptr;
  const int64_t *rowptr = nullptr;
  const uint32_t *maskw = nullptr;
};

void em_stats_host_impl(ppca_b200_ctx *ctx, const double *x, int64_t n, int d, const double *weights, const DevModel &m,
                        double *stats_dev, const PackedHost *pk = nullptr) {
  if (!ctx->copy_stream) {
    CUDA_CHECK(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
    for (int b = 0; b < 2; ++b) {
      CUDA_CHECK(cudaEventCreateWithFlags(&ctx->ev_copied[b], cudaEventDisableTiming));
      CUDA_CHECK(cudaEventCreateWithFlags(&ctx->ev_free[b], cudaEventDisableTiming));
    }
  }
  const EmPlan p = em_begin(ctx, pick_stream_block(ctx, round_up(n, 256), m.s), m, stats_dev);
  const int64_t blk = p.chunk;  // one uploaded block = one E/M-step chunk
  const int64_t tail = n % blk;
  if (n >= blk && (!ctx->s_store || ctx->s_store->n != blk || ctx->s_store->d != d)) ctx->s_store = make_store(ctx, blk, d);
  if (tail && (!ctx->s_tail || ctx->s_tail->n != tail || ctx->s_tail->d != d)) ctx->s_tail = make_store(ctx, tail, d);
  const int64_t brows = n < blk ? n : blk;
  const int dw = (d + 31) / 32;
  size_t max_nnz = 0;  // most observed values of any block (packed source)
  if (pk)
    for (int64_t r0 = 0; r0 < n; r0 += blk) {
      const int64_t r1 = std::min<int64_t>(n, r0 + blk);
      max_nnz = std::max(max_nnz, (size_t)(pk->rowptr[r1] - pk->rowptr[r0]));
    }
  for (int b = 0; b < 2; ++b) {
    if (pk) {
      ctx->s_raw[b].reserve(max_nnz ? max_nnz : 1);
      ctx->s_rowptr[b].reserve((size_t)brows + 1);
      ctx->s_maskw[b].reserve((size_t)brows * dw);
    } else {
      ctx->s_raw[b].reserve((size_t)brows * d);
    }
    if (weights) ctx->s_wraw[b].reserve((size_t)brows);
  }
  ctx->s_w.reserve((size_t)round_up(brows, 256));
  // the copy stream must not run ahead of work already queued on the compute stream that still reads the raw blocks
  CUDA_CHECK(cudaEventRecord(ctx->ev_free[0], ctx->stream));
  CUDA_CHECK(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_free[0], 0));
  const Launcher L = ctx->L();
  int i = 0;
  for (int64_t r0 = 0; r0 < n; r0 += blk, ++i) {
    const int b = i & 1;
    const int64_t rows = (n - r0) < blk ? (n - r0) : blk;
    SampleStore &st = rows == blk ? *ctx->s_store : *ctx->s_tail;
    // H2D of block i on the copy stream, overlapped with the kernels of block i-1 on the compute stream
    if (i >= 2) CUDA_CHECK(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_free[b], 0));
    if (pk) {
      const int64_t nnz = pk->rowptr[r0 + rows] - pk->rowptr[r0];
      if (nnz > 0)
        CUDA_CHECK(cudaMemcpyAsync(ctx->s_raw[b].p, pk->vals + pk->rowptr[r0], sizeof(double) * nnz, cudaMemcpyHostToDevice,
                                   ctx->copy_stream));
      CUDA_CHECK(cudaMemcpyAsync(ctx->s_rowptr[b].p, pk->rowptr + r0, sizeof(int64_t) * (rows + 1), cudaMemcpyHostToDevice,
                                 ctx->copy_stream));
      CUDA_CHECK(cudaMemcpyAsync(ctx->s_maskw[b].p, pk->maskw + r0 * dw, sizeof(uint32_t) * rows * dw,
                                 cudaMemcpyHostToDevice, ctx->copy_stream));
    } else {
      CUDA_CHECK(cudaMemcpyAsync(ctx->s_raw[b].p, x + r0 * d, sizeof(double) * rows * d, cudaMemcpyHostToDevice,
                                 ctx->copy_stream));
    }
    if (weights)
      CUDA_CHECK(cudaMemcpyAsync(ctx->s_wraw[b].p, weights + r0, sizeof(double) * rows, cudaMemcpyHostToDevice,
                                 ctx->copy_stream));
    CUDA_CHECK(cudaEventRecord(ctx->ev_copied[b], ctx->copy_stream));
    CUDA_CHECK(cudaStreamWaitEvent(ctx->stream, ctx->ev_copied[b], 0));
    if (pk) launch_unpack(L, ctx->s_raw[b].p, ctx->s_rowptr[b].p, ctx->s_maskw[b].p, rows, d, 0, st);
    else launch_ingest(L, ctx->s_raw[b].p, rows, d, 0, st);
    CUDA_CHECK(cudaMemsetAsync(ctx->s_w.p, 0, sizeof(double) * st.n_pad, ctx->stream));
    if (weights) {
      CUDA_CHECK(cudaMemcpyAsync(ctx->s_w.p, ctx->s_wraw[b].p, sizeof(double) * rows, cudaMemcpyDeviceToDevice,
                                 ctx->stream));
    } else {
      const int64_t want = (rows + 255) / 256;
      const int blocks = (int)(want < (int64_t)ctx->sms * 8 ? want : (int64_t)ctx->sms * 8);
      fill_value_kernel<<<blocks, 256, 0, ctx->stream>>>(ctx->s_w.p, rows, 1.0);
      CUDA_CHECK(cudaGetLastError());
      ++ctx->launches;
    }
    CUDA_CHECK(cudaEventRecord(ctx->ev_free[b], ctx->stream));
    launch_transpose_mask(L, st);
    em_chunk(ctx, st, ctx->s_w.p, 0, (int)rows, m, stats_dev, p);
  }
  em_end(ctx, m, stats_dev, p);
}
}  // namespace

extern "C" {

int32_t ppca_b200_em_stats_host(ppca_b200_ctx *ctx, const double *x, int64_t n, int32_t d, const double *weights,
                                int32_t k, const double *C, const double *mu, double sigma, double *stats_dev) {
  return guarded([&] {
    REQUIRE(ctx != nullptr && stats_dev != nullptr, "null argument");
    REQUIRE(n >= 0 && d >= 1, "bad dataset shape %lld x %d", (long long)n, d);
    REQUIRE(n == 0 || x != nullptr, "null data");
    DeviceGuard g(ctx->device);
    begin_pass(ctx);
    DevModel m = stage_model(ctx, d, k, C, mu, sigma);
    if (n == 0) {
      CUDA_CHECK(cudaMemsetAsync(stats_dev, 0, sizeof(double) * StatsLayout(d, k).len, ctx->stream));
      return;
    }
    em_stats_host_impl(ctx, x, n, d, weights, m, stats_dev);
  });
}

int32_t ppca_b200_iterate_host(ppca_b200_ctx *ctx, const double *x, int64_t n, int32_t d, const double *weights,
                               int32_t k, const double *C, const double *mu, double sigma,
                               const ppca_b200_prior *prior, double *C_out, double *mu_out, double *sigma_out,
                               double *llk_in) {
  return guarded([&] {
    REQUIRE(ctx != nullptr, "null context");
    REQUIRE(n >= 0 && d >= 1, "bad dataset shape %lld x %d", (long long)n, d);
    if (n == 0) PPCA_THROW(PPCA_ERR_EMPTY, "non-empty dataset required (ppca_model.rs:358)");
    REQUIRE(x != nullptr, "null data");
    DeviceGuard g(ctx->device);
    run_guarded(ctx, [&] {
      DevModel m = stage_model(ctx, d, k, C, mu, sigma);
      ctx->stats.reserve((size_t)StatsLayout(d, k).len);
      em_stats_host_impl(ctx, x, n, d, weights, m, ctx->stats.p);
      return em_finish_impl(ctx, d, k, C, mu, sigma, prior, ctx->stats.p, C_out, mu_out, sigma_out, llk_in, nullptr);
    });
  });
}

// ---- out-of-core inference: smooth / extrapolate / llks over samples that stay in host memory ------------
int32_t ppca_b200_reconstruct_host(ppca_b200_ctx *ctx, const double *x, int64_t n, int32_t d, int32_t k,
                                   const double *C, const double *mu, double sigma, int32_t extrapolate, double *out,
                                   double *llks) {
  return guarded([&] {
    REQUIRE(ctx != nullptr, "null context");
    REQUIRE(n >= 0 && d >= 1, "bad dataset shape %lld x %d", (long long)n, d);
    REQUIRE(out != nullptr || llks != nullptr, "nothing to compute: both outputs are null");
    if (n == 0) return;
    REQUIRE(x != nullptr, "null data");
    DeviceGuard g(ctx->device);
    if (!ctx->copy_stream) {
      CUDA_CHECK(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
      for (int b = 0; b < 2; ++b) {
        CUDA_CHECK(cudaEventCreateWithFlags(&ctx->ev_copied[b], cudaEventDisableTiming));
        CUDA_CHECK(cudaEventCreateWithFlags(&ctx->ev_free[b], cudaEventDisableTiming));
      }
    }
    if (!ctx->out_stream) {
      CUDA_CHECK(cudaStreamCreateWithFlags(&ctx->out_stream, cudaStreamNonBlocking));
      for (int b = 0; b < 2; ++b) {
        CUDA_CHECK(cudaEventCreateWithFlags(&ctx->ev_done[b], cudaEventDisableTiming));
        CUDA_CHECK(cudaEventCreateWithFlags(&ctx->ev_outfree[b], cudaEventDisableTiming));
      }
    }
    run_guarded(ctx, [&] {
    DevModel m = stage_model(ctx, d, k, C, mu, sigma);
    const int64_t blk = pick_stream_block(ctx, round_up(n, 256), m.s);
    reserve_chunk_ws(ctx, blk, m.s);
    const int64_t tail = n % blk;
    if (n >= blk && (!ctx->s_store || ctx->s_store->n != blk || ctx->s_store->d != d)) ctx->s_store = make_store(ctx, blk, d);
    if (tail && (!ctx->s_tail || ctx->s_tail->n != tail || ctx->s_tail->d != d)) ctx->s_tail = make_store(ctx, tail, d);
    const int64_t brows = n < blk ? n : blk;
    for (int b = 0; b < 2; ++b) {
      ctx->s_raw[b].reserve((size_t)brows * d);
      if (out) ctx->s_out[b].reserve((size_t)brows * d);
      if (llks) ctx->s_llk[b].reserve((size_t)round_up(brows, 256));
    }
    CUDA_CHECK(cudaEventRecord(ctx->ev_free[0], ctx->stream));  // neither side stream may run ahead of queued work
    CUDA_CHECK(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_free[0], 0));
    CUDA_CHECK(cudaStreamWaitEvent(ctx->out_stream, ctx->ev_free[0], 0));
    const Launcher L = ctx->L();
    int i = 0;
    for (int64_t r0 = 0; r0 < n; r0 += blk, ++i) {
      const int b = i & 1;
      const int64_t rows = (n - r0) < blk ? (n - r0) : blk;
      SampleStore &st = rows == blk ? *ctx->s_store : *ctx->s_tail;
      // H2D of block i (copy stream) || kernels of block i-1 (compute stream) || D2H of block i-2 (out stream)
      if (i >= 2) CUDA_CHECK(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_free[b], 0));
      CUDA_CHECK(cudaMemcpyAsync(ctx->s_raw[b].p, x + r0 * d, sizeof(double) * rows * d, cudaMemcpyHostToDevice,
                                 ctx->copy_stream));
      CUDA_CHECK(cudaEventRecord(ctx->ev_copied[b], ctx->copy_stream));
      CUDA_CHECK(cudaStreamWaitEvent(ctx->stream, ctx->ev_copied[b], 0));
      launch_ingest(L, ctx->s_raw[b].p, rows, d, 0, st);
      CUDA_CHECK(cudaEventRecord(ctx->ev_free[b], ctx->stream));
      if (i >= 2) CUDA_CHECK(cudaStreamWaitEvent(ctx->stream, ctx->ev_outfree[b], 0));  // outputs of block i-2 left
      e_step_chunk(ctx, st, nullptr, 0, (int)rows, m, out ? 1 : 0, llks ? ctx->s_llk[b].p : nullptr, nullptr, nullptr);
      if (out)
        launch_reconstruct(L, st, 0, (int)rows, m, ctx->YZ.p, extrapolate, nullptr, 0, 0, ctx->s_out[b].p, d);
      CUDA_CHECK(cudaEventRecord(ctx->ev_done[b], ctx->stream));
      CUDA_CHECK(cudaStreamWaitEvent(ctx->out_stream, ctx->ev_done[b], 0));
      if (out)
        CUDA_CHECK(cudaMemcpyAsync(out + r0 * d, ctx->s_out[b].p, sizeof(double) * rows * d, cudaMemcpyDeviceToHost,
                                   ctx->out_stream));
      if (llks)
        CUDA_CHECK(cudaMemcpyAsync(llks + r0, ctx->s_llk[b].p, sizeof(double) * rows, cudaMemcpyDeviceToHost,
                                   ctx->out_stream));
      CUDA_CHECK(cudaEventRecord(ctx->ev_outfree[b], ctx->out_stream));
    }
    CUDA_CHECK(cudaStreamSynchronize(ctx->out_stream));
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    return read_unsafe(ctx);
    });
  });
}

// ---- covariance diagonals ------------------------------------------------------------------------------------------
// InferredMasked::smoothed_covariance / extrapolated_covariance (ppca_model.rs:471-477, 517-534): per sample the full
// d x d matrix  sigma^2 I + C Sigma_n C^T, with the rows and columns of the dimensions the sample observed zeroed when a
// dataset is given (negative.expand_matrix; all zeros when nothing is missing).  One CTA per (sample, 32 x 32 output tile):
// T = C_tile Sigma_n (32 x k) and the column tile of C sit in shared memory with an odd pitch.
__global__ void __launch_bounds__(256) cov_full_kernel(const double *__restrict__ C, int d, int k, double s2,
                                                       const double *__restrict__ covs, int64_t rows,
                                                       const uint32_t *__restrict__ mask, int dw, int64_t mask_row0,
                                                       double *__restrict__ out) {
  extern __shared__ double sm_cov[];
  const int pitch = k | 1;
  double *T = sm_cov;               // 32 x pitch
  double *Cj = T + 32 * pitch;      // 32 x pitch
  double *Sg = Cj + 32 * pitch;     // k x k
  const int tid = threadIdx.x, lane = tid & 31, wi = tid >> 5;
  const int i0 = blockIdx.y * 32, j0 = blockIdx.x * 32;
  for (int64_t row = blockIdx.z; row < rows; row += gridDim.z) {
    const double *S = covs + row * (int64_t)k * k;
    for (int q = tid; q < k * k; q += 256) Sg[q] = S[q];
    for (int q = tid; q < 32 * k; q += 256) {
      const int r = q / k, a = q % k;
      Cj[r * pitch + a] = (j0 + r < d) ? C[(int64_t)(j0 + r) * k + a] : 0.0;
    }
    __syncthreads();
    for (int q = tid; q < 32 * k; q += 256) {  // T[r][b] = sum_a C[i0 + r][a] Sigma[a][b]
      const int r = q / k, b = q % k;
      double acc = 0.0;
      if (i0 + r < d) {
        const double *ci = C + (int64_t)(i0 + r) * k;
        for (int a = 0; a < k; ++a) acc = fma(ci[a], Sg[a * k + b], acc);
      }
      T[r * pitch + b] = acc;
    }
    __syncthreads();
    const int j = j0 + lane;
    bool obs_j = false;
    if (mask && j < d) obs_j = (mask[(mask_row0 + row) * dw + (j >> 5)] >> (j & 31)) & 1u;
    for (int r = wi; r < 32; r += 8) {
      const int i = i0 + r;
      if (i >= d || j >= d) continue;
      double acc = (i == j) ? s2 : 0.0;
      for (int a = 0; a < k; ++a) acc = fma(T[r * pitch + a], Cj[lane * pitch + a], acc);
      if (mask) {
        const bool obs_i = (mask[(mask_row0 + row) * dw + (i >> 5)] >> (i & 31)) & 1u;
        if (obs_i || obs_j) acc = 0.0;
      }
      out[(row * d + i) * (int64_t)d + j] = acc;
    }
    __syncthreads();
  }
}

int32_t ppca_b200_covariance_full(ppca_b200_ctx *ctx, int64_t n, int32_t d, int32_t k, const double *C, double sigma,
                                  const double *covariances, const ppca_b200_dataset *masked_by, double *out) {
  return guarded([&] {
    REQUIRE(ctx != nullptr && C != nullptr, "null argument");
    REQUIRE(n >= 0 && d >= 1 && k >= 1, "bad shape");
    REQUIRE(n == 0 || (covariances != nullptr && out != nullptr), "null covariances / output");
    if (masked_by) {
      check_ds(ctx, masked_by);
      REQUIRE(masked_by->store->n == n && masked_by->store->d == d, "dataset shape does not match the inferred batch");
    }
    if (n == 0) return;
    DeviceGuard g(ctx->device);
    const size_t smem = sizeof(double) * ((size_t)64 * (k | 1) + (size_t)k * k);
    REQUIRE(smem <= 200 * 1024, "state_size %d too large for the full-covariance kernel", k);
    static PerDeviceOnce configured;
    if (configured.need())
      CUDA_CHECK(cudaFuncSetAttribute(cov_full_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    DevBuf<double> Cd, covd, outd;
    Cd.alloc((size_t)d * k);
    CUDA_CHECK(cudaMemcpyAsync(Cd.p, C, sizeof(double) * d * k, cudaMemcpyHostToDevice, ctx->stream));
    // rows per pass: result staging <= 512 MiB
    int64_t rows_per = std::max<int64_t>(1, ((int64_t)512 << 20) / ((int64_t)d * d * 8));
    rows_per = std::min<int64_t>(rows_per, n);
    covd.alloc((size_t)rows_per * k * k);
    outd.alloc((size_t)rows_per * d * d);
    const int tiles = (d + 31) / 32;
    for (int64_t r0 = 0; r0 < n; r0 += rows_per) {
      const int64_t rows = std::min<int64_t>(rows_per, n - r0);
      CUDA_CHECK(cudaMemcpyAsync(covd.p, covariances + r0 * k * k, sizeof(double) * rows * k * k, cudaMemcpyHostToDevice,
                                 ctx->stream));
      dim3 grid((unsigned)tiles, (unsigned)tiles, (unsigned)std::min<int64_t>(rows, 32768));
      cov_full_kernel<<<grid, 256, smem, ctx->stream>>>(Cd.p, d, k, sigma * sigma, covd.p, rows,
                                                        masked_by ? masked_by->store->mask.p : nullptr,
                                                        masked_by ? masked_by->store->dw : 0, r0, outd.p);
      CUDA_CHECK(cudaGetLastError());
      ++ctx->launches;
      CUDA_CHECK(cudaMemcpyAsync(out + r0 * d * d, outd.p, sizeof(double) * rows * d * d, cudaMemcpyDeviceToHost,
                                 ctx->stream));
      CUDA_CHECK(cudaStreamSynchronize(ctx->stream));  // covd / outd are reused by the next pass
    }
  });
}

int32_t ppca_b200_covariance_diagonal(ppca_b200_ctx *ctx, int64_t n, int32_t d, int32_t k, const double *C, double sigma,
                                      const double *covariances, const ppca_b200_dataset *masked_by,
                                      ppca_b200_dataset **out) {
  return guarded([&] {
    REQUIRE(ctx != nullptr && out != nullptr && C != nullptr, "null argument");
    REQUIRE(n >= 0 && d >= 1 && k >= 1, "bad shape");
    REQUIRE(n == 0 || covariances != nullptr, "null covariances");
    if (masked_by) {
      check_ds(ctx, masked_by);
      REQUIRE(masked_by->store->n == n && masked_by->store->d == d, "dataset shape does not match the inferred batch");
    }
    DeviceGuard g(ctx->device);
    auto ost = make_store(ctx, n, d);
    if (n > 0) {
      std::vector<double> zero_mu((size_t)d, 0.0);
      DevModel m = stage_model(ctx, d, k, C, zero_mu.data(), sigma > 0.0 ? sigma : 1.0, true);
      const int kk = m.s.kk, kkp = m.s.kkp, kq32 = (int)round_up(kkp, 32), d8 = (int)round_up(d, 8);
      DevBuf<double> KT, zeros, S2, Y, nxs, covd;
      DevBuf<uint32_t> ones;
      KT.alloc((size_t)kq32 * d8);
      zeros.alloc((size_t)kq32);
      ones.alloc((size_t)kq32 / 32);
      CUDA_CHECK(cudaMemsetAsync(zeros.p, 0, sizeof(double) * kq32, ctx->stream));
      CUDA_CHECK(cudaMemsetAsync(ones.p, 0xff, sizeof(uint32_t) * (kq32 / 32), ctx->stream));
      const int tb = (int)std::min<int64_t>((int64_t)ctx->sms * 8, ((int64_t)kq32 * d8 + 255) / 256);
      transpose_ksym_kernel<<<tb, 256, 0, ctx->stream>>>(m.Ksym, d, kk, kkp, kq32, d8, KT.p);
      CUDA_CHECK(cudaGetLastError());
      ++ctx->launches;
      // rows per pass: covariance staging <= 256 MiB, result staging <= 512 MiB
      int64_t rows_per = std::min<int64_t>(((int64_t)256 << 20) / ((int64_t)k * k * 8), ((int64_t)512 << 20) / ((int64_t)d8 * 8));
      rows_per = std::max<int64_t>(128, rows_per / 128 * 128);
      if (rows_per > round_up(n, 128)) rows_per = round_up(n, 128);
      covd.alloc((size_t)rows_per * k * k);
      S2.alloc((size_t)rows_per * kkp);
      Y.alloc((size_t)rows_per * d8);
      nxs.alloc((size_t)rows_per);
      for (int64_t r0 = 0; r0 < n; r0 += rows_per) {
        const int64_t rows = std::min<int64_t>(rows_per, n - r0), rows_pad = round_up(rows, 128);
        CUDA_CHECK(cudaMemcpyAsync(covd.p, covariances + r0 * k * k, sizeof(double) * rows * k * k, cudaMemcpyHostToDevice,
                                   ctx->stream));
        const int pb = (int)std::min<int64_t>((int64_t)ctx->sms * 16, (rows_pad * kkp + 255) / 256);
        pack_cov_kernel<<<pb, 256, 0, ctx->stream>>>(covd.p, rows, rows_pad, k, kk, kkp, S2.p);
        CUDA_CHECK(cudaGetLastError());
        ++ctx->launches;
        launch_rowgemm(ctx->L(), S2.p, kkp, (int)rows_pad, kkp, KT.p, d8, ones.p, zeros.p, Y.p, nxs.p);
        const int fb = (int)std::min<int64_t>((int64_t)ctx->sms * 16, (rows * d + 255) / 256);
        cov_diag_finish_kernel<<<fb, 256, 0, ctx->stream>>>(Y.p, d8, rows, d, sigma * sigma,
                                                            masked_by ? masked_by->store->mask.p : nullptr,
                                                            masked_by ? masked_by->store->dw : 0, r0,
                                                            ost->X.p + r0 * ost->ldx, ost->ldx);
        CUDA_CHECK(cudaGetLastError());
        ++ctx->launches;
        CUDA_CHECK(cudaStreamSynchronize(ctx->stream));  // covd / S2 / Y are reused by the next pass
      }
      finalize_full_store(ctx, *ost);
    }
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    *out = make_dataset(ctx, ost, nullptr);
  });
}

// ---- PPCAMix --------------------------------------------------------------------------------------
int32_t ppca_b200_mix_llks(ppca_b200_ctx *ctx, const ppca_b200_dataset *ds, int32_t m, const int32_t *ks,
                           const double *Cs, const double *mus, const double *sigmas, const double *log_weights,
                           double *out) {
  return guarded([&] {
    check_ds(ctx, ds);
    MixView mv{m, ks, Cs, mus, sigmas, log_weights, ds->store->d};
    check_mix(mv);
    const SampleStore &st = *ds->store;
    if (st.n == 0) return;
    REQUIRE(out != nullptr, "null output");
    DeviceGuard g(ctx->device);
    DevBuf<double> &LP = ctx->mixLP, &ml = ctx->mixLlk;
    LP.reserve((size_t)st.n * m);
    ml.reserve((size_t)st.n);
    run_guarded(ctx, [&] {
      mix_posteriors_impl(ctx, ds, mv, LP.p, ml.p, nullptr, nullptr);
      CUDA_CHECK(cudaMemcpyAsync(out, ml.p, sizeof(double) * st.n, cudaMemcpyDeviceToHost, ctx->stream));
      CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
      return read_unsafe(ctx);
    });
  });
}

int32_t ppca_b200_mix_llk(ppca_b200_ctx *ctx, const ppca_b200_dataset *ds, int32_t m, const int32_t *ks,
                          const double *Cs, const double *mus, const double *sigmas, const double *log_weights,
                          double *out) {
  return guarded([&] {
    check_ds(ctx, ds);
    MixView mv{m, ks, Cs, mus, sigmas, log_weights, ds->store->d};
    check_mix(mv);
    REQUIRE(out != nullptr, "null output");
    const SampleStore &st = *ds->store;
    if (st.n == 0) {  // mix.rs:164-166
      *out = 0.0;
      return;
    }
    DeviceGuard g(ctx->device);
    DevBuf<double> &LP = ctx->mixLP, &ml = ctx->mixLlk, &sum = ctx->mixSum;
    LP.reserve((size_t)st.n * m);
    ml.reserve((size_t)st.n);
    sum.reserve(1);
    run_guarded(ctx, [&] {
      mix_posteriors_impl(ctx, ds, mv, LP.p, ml.p, nullptr, sum.p);
      CUDA_CHECK(cudaMemcpyAsync(out, sum.p, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
      CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
      return read_unsafe(ctx);
    });
  });
}

int32_t ppca_b200_mix_infer_cluster(ppca_b200_ctx *ctx, const ppca_b200_dataset *ds, int32_t m, const int32_t *ks,
                                    const double *Cs, const double *mus, const double *sigmas,
                                    const double *log_weights, double *out) {
  return guarded([&] {
    check_ds(ctx, ds);
    MixView mv{m, ks, Cs, mus, sigmas, log_weights, ds->store->d};
    check_mix(mv);
    const SampleStore &st = *ds->store;
    if (st.n == 0) return;
    REQUIRE(out != nullptr, "null output");
    DeviceGuard g(ctx->device);
    DevBuf<double> &LP = ctx->mixLP;
    LP.reserve((size_t)st.n * m);
    run_guarded(ctx, [&] {
      mix_posteriors_impl(ctx, ds, mv, LP.p, nullptr, nullptr, nullptr);
      CUDA_CHECK(cudaMemcpyAsync(out, LP.p, sizeof(double) * st.n * m, cudaMemcpyDeviceToHost, ctx->stream));
      CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
      return read_unsafe(ctx);
    });
  });
}

__global__ void exp_inplace_kernel(double *p, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    p[i] = exp(p[i]);
}

static int32_t mix_smooth_or_extrapolate(ppca_b200_ctx *ctx, const ppca_b200_dataset *ds, int32_t m,
                                         const int32_t *ks, const double *Cs, const double *mus, const double *sigmas,
                                         const double *log_weights, int extrapolate, ppca_b200_dataset **out) {
  return guarded([&] {
    check_ds(ctx, ds);
    MixView mv{m, ks, Cs, mus, sigmas, log_weights, ds->store->d};
    check_mix(mv);
    REQUIRE(out != nullptr, "null output");
    const SampleStore &st = *ds->store;
    DeviceGuard g(ctx->device);
    auto ost = make_store(ctx, st.n, st.d);
    if (st.n > 0) {
      DevBuf<double> &LP = ctx->mixLP;
      LP.reserve((size_t)st.n * m);
      run_guarded(ctx, [&] {
        mix_posteriors_impl(ctx, ds, mv, LP.p, nullptr, nullptr, nullptr);
        const int64_t total = st.n * m;
        const int blocks = (int)((total + 255) / 256 < (int64_t)ctx->sms * 8 ? (total + 255) / 256 : (int64_t)ctx->sms * 8);
        exp_inplace_kernel<<<blocks, 256, 0, ctx->stream>>>(LP.p, total);  // posterior() (mix.rs:366-368)
        CUDA_CHECK(cudaGetLastError());
        ++ctx->launches;
        for (int j = 0; j < m; ++j) {  // sum_j posterior_j * (smoothed | extrapolated)_j  (mix.rs:397-414)
          DevModel dm = stage_model(ctx, st.d, ks[j], mv.C(j), mv.mu(j), sigmas[j]);
          reconstruct_impl(ctx, st, dm, extrapolate, LP.p + j, m, j > 0 ? 1 : 0, *ost);
        }
        return read_unsafe(ctx);
      });
      finalize_full_store(ctx, *ost);
    }
    *out = make_dataset(ctx, ost, nullptr);  // weights reset to 1 (mix.rs:245-265)
  });
}

int32_t ppca_b200_mix_smooth(ppca_b200_ctx *ctx, const ppca_b200_dataset *ds, int32_t m, const int32_t *ks,
                             const double *Cs, const double *mus, const double *sigmas, const double *log_weights,
                             ppca_b200_dataset **out) {
  return mix_smooth_or_extrapolate(ctx, ds, m, ks, Cs, mus, sigmas, log_weights, 0, out);
}
int32_t ppca_b200_mix_extrapolate(ppca_b200_ctx *ctx, const ppca_b200_dataset *ds, int32_t m, const int32_t *ks,
                                  const double *Cs, const double *mus, const double *sigmas, const double *log_weights,
                                  ppca_b200_dataset **out) {
  return mix_smooth_or_extrapolate(ctx, ds, m, ks, Cs, mus, sigmas, log_weights, 1, out);
}

int32_t ppca_b200_mix_posteriors(ppca_b200_ctx *ctx, const ppca_b200_dataset *ds, int32_t m, const int32_t *ks,
                                 const double *Cs, const double *mus, const double *sigmas, const double *log_weights,
                                 double *logpost_dev, double *comp_max, double *llk_in) {
  return guarded([&] {
    check_ds(ctx, ds);
    MixView mv{m, ks, Cs, mus, sigmas, log_weights, ds->store->d};
    check_mix(mv);
    REQUIRE(logpost_dev != nullptr && comp_max != nullptr, "null output");
    const SampleStore &st = *ds->store;
    if (st.n > 0 && !(ds->min_w > 0.0))
      PPCA_THROW(PPCA_ERR_WEIGHTS, "mixture EM needs strictly positive weights (mix.rs:304-309,326)");
    DeviceGuard g(ctx->device);
    DevBuf<double> &ml = ctx->mixLlk, &cm = ctx->mixMax, &sum = ctx->mixSum;
    ml.reserve((size_t)std::max<int64_t>(st.n, 1));
    cm.reserve((size_t)m);
    sum.reserve(1);
    run_guarded(ctx, [&] {
      CUDA_CHECK(cudaMemsetAsync(sum.p, 0, sizeof(double), ctx->stream));
      mix_posteriors_impl(ctx, ds, mv, logpost_dev, ml.p, cm.p, st.n > 0 ? sum.p : nullptr);
      double h = 0.0;
      CUDA_CHECK(cudaMemcpyAsync(comp_max, cm.p, sizeof(double) * m, cudaMemcpyDeviceToHost, ctx->stream));
      CUDA_CHECK(cudaMemcpyAsync(&h, sum.p, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
      CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
      if (llk_in) *llk_in = h;
      return read_unsafe(ctx);
    });
  });
}

int32_t ppca_b200_mix_em_stats(ppca_b200_ctx *ctx, const ppca_b200_dataset *ds, int32_t m, int32_t j, int32_t k,
                               const double *C, const double *mu, double sigma, const double *logpost_dev,
                               double comp_max_j, double *stats_dev) {
  return guarded([&] {
    check_ds(ctx, ds);
    REQUIRE(logpost_dev != nullptr && stats_dev != nullptr, "null argument");
    REQUIRE(j >= 0 && j < m, "component index out of range");
    const SampleStore &st = *ds->store;
    DeviceGuard g(ctx->device);
    begin_pass(ctx);
    ctx->rbuf.reserve((size_t)st.n_pad);
    CUDA_CHECK(cudaMemsetAsync(ctx->rbuf.p, 0, sizeof(double) * st.n_pad, ctx->stream));
    launch_responsibilities(ctx->L(), logpost_dev, st.n, m, j, ds->w.p, comp_max_j, ctx->rbuf.p);
    DevModel dm = stage_model(ctx, st.d, k, C, mu, sigma);
    em_stats_impl(ctx, st, ctx->rbuf.p, dm, stats_dev);
  });
}

}  // extern "C"

namespace {

// ---- single-pass mixture EM (mix.rs:281-337) ---------------------------------------------------------------------
// The reference runs the E-step of every component twice per iteration: once for the posteriors (infer_cluster,
// mix.rs:297-303) and once more inside each component's weighted iterate (:326).  Here ONE E-step per component and chunk
// serves both: all components' unweighted second moments V = z z^T + Sigma, states and trace terms of a chunk stay on the
// device while the chunk's log-posteriors are normalised, then each component weighs its V by its responsibilities and
// runs its M-step contractions.  The responsibilities are scaled by exp(-max) of a RUNNING per-component maximum; the
// accumulators of a component are rescaled when a later chunk raises it (see mix.cu), so the result is the reference's
// exp(lp - max over the dataset) weighting up to rounding.
struct MixComp {
  DevModel m;
  ModelWs ws;
  EmPlan plan;
  double *stats = nullptr;
  int64_t stats_len = 0;
};

struct MixPass {
  std::vector<MixComp> comps;
  double *stats_all = nullptr;  // [stats_0 | ... | stats_{M-1} | llk_sum]   (one all-reduce)
  int64_t stats_total = 0;      // doubles, including the trailing llk_sum
  double *run_max = nullptr;    // M
};

struct Carver {
  double *p;
  size_t used = 0;
  explicit Carver(double *base) : p(base) {}
  double *take(size_t n) {
    n = (n + 3) & ~(size_t)3;  // keep 32-byte alignment
    double *r = p ? p + used : nullptr;
    used += n;
    return r;
  }
};

void mix_em_pass(ppca_b200_ctx *ctx, const ppca_b200_dataset *ds, const MixView &mv, MixPass &out, double *LP_full) {
  const SampleStore &st = *ds->store;
  const Launcher L = ctx->L();
  const int M = mv.m, d = st.d;
  std::vector<Shape> shp(M);
  int kp_max = 8, kkp_max = 8;
  int64_t per_row = 0, stats_total = 0;
  for (int j = 0; j < M; ++j) {
    REQUIRE(mv.ks[j] >= 1, "state_size must be >= 1 (got %d)", mv.ks[j]);
    shp[j] = Shape(d, mv.ks[j]);
    kp_max = std::max(kp_max, shp[j].kp);
    kkp_max = std::max(kkp_max, shp[j].kkp);
    per_row += (int64_t)(shp[j].kkp + shp[j].kp + 3) * 8;
    stats_total += round_up(StatsLayout(d, mv.ks[j]).len, 4);  // every component's buffer stays 32-byte aligned
  }
  per_row += (int64_t)(2 * M + kp_max + 4) * 8 + (int64_t)kkp_max * ctx->slices;
  // rows per chunk: all components' V / Z / t of a chunk live together (at most 16 GiB), at least one wave of row tiles
  int64_t chunk = ctx->chunk;
  if (chunk <= 0) {
    int64_t cap = ((int64_t)16 << 30) / per_row;
    cap = std::min<int64_t>(cap, (int64_t)1 << 20);
    cap = std::max<int64_t>(cap, (int64_t)ctx->sms * 128);
    const int64_t nchunks = (st.n_pad + cap - 1) / cap;
    chunk = (st.n_pad + nchunks - 1) / nchunks;
  }
  chunk = std::min<int64_t>(round_up(chunk, 256), st.n_pad);

  ensure_rscratch(ctx);
  // ---- carve the arenas (two passes: size, then pointers)
  out.comps.assign(M, MixComp());
  size_t need = 0, need_q = 0;
  double *sh_nx = nullptr, *sh_llk = nullptr, *sh_WZ = nullptr, *sh_r = nullptr, *LPc = nullptr, *mixllk = nullptr,
         *chunk_max = nullptr, *factor = nullptr, *sigmas_dev = nullptr;
  int8_t *sh_WQ = nullptr, *sh_ZQ = nullptr;
  for (int pass = 0; pass < 2; ++pass) {
    if (pass == 1) {
      ctx->mixArena.reserve(need);
      ctx->mixArenaQ.reserve(need_q ? need_q : 1);
      ctx->mixStats.reserve((size_t)stats_total + 1);
    }
    Carver cv(pass ? ctx->mixArena.p : nullptr);
    size_t qoff = 0;
    sh_nx = cv.take((size_t)chunk);
    sh_llk = cv.take((size_t)chunk);
    sh_WZ = cv.take((size_t)chunk * kp_max);
    sh_r = cv.take((size_t)chunk);
    LPc = cv.take((size_t)chunk * M);
    mixllk = cv.take((size_t)chunk);
    out.run_max = cv.take((size_t)M);
    chunk_max = cv.take((size_t)M);
    factor = cv.take((size_t)M);
    sigmas_dev = cv.take((size_t)M);
    sh_WQ = pass ? ctx->mixArenaQ.p + qoff : nullptr;
    {
      size_t wq = 0;
      for (int j = 0; j < M; ++j) wq = std::max(wq, wq_bytes(ctx, chunk, shp[j]));
      qoff += (wq + 1023) & ~(size_t)1023;
    }
    sh_ZQ = pass ? ctx->mixArenaQ.p + qoff : nullptr;
    {
      size_t zq = 0;
      for (int j = 0; j < M; ++j) zq = std::max(zq, zq_bytes(ctx, chunk, shp[j]));
      qoff += (zq + 1023) & ~(size_t)1023;
    }
    int64_t soff = 0;
    for (int j = 0; j < M; ++j) {
      MixComp &c = out.comps[j];
      const Shape &s = shp[j];
      c.plan = em_plan(ctx, chunk, s);
      double *Cpad = cv.take((size_t)s.d32 * s.kp), *mupad = cv.take((size_t)s.d32), *Ksym = cv.take((size_t)s.d32 * s.kkp);
      c.ws.KsymScale = cv.take((size_t)s.kkp);
      c.ws.WScale = cv.take((size_t)s.kkp);
      c.ws.WScaleMax = cv.take((size_t)s.kkp);
      c.ws.colmax = reinterpret_cast<unsigned long long *>(cv.take((size_t)s.kkp));
      c.ws.GW = cv.take((size_t)chunk * s.kkp);
      c.ws.YZ = cv.take((size_t)chunk * s.kp);
      c.ws.tn = cv.take((size_t)chunk);
      c.ws.part_bg = cv.take(c.plan.bglen);
      c.ws.part_cr = cv.take(c.plan.crlen);
      c.ws.part_solve = cv.take((size_t)SOLVE_SLOTS * 4);
      c.ws.ZScale = cv.take((size_t)s.kp);
      c.ws.zcolmax = reinterpret_cast<unsigned long long *>(cv.take((size_t)s.kp));
      c.ws.MZ = cv.take((size_t)s.d * s.kp);
      c.ws.part_mz = cv.take(c.plan.mzlen);
      c.ws.ZQ = sh_ZQ;
      c.ws.nx = cv.take((size_t)chunk);   // |x~|^2 depends on the component's mean
      c.ws.dv = cv.take((size_t)chunk);
      c.ws.part_rx = cv.take(c.plan.rxlen);
      c.ws.rflag = reinterpret_cast<int *>(cv.take(4));
      c.ws.rscratch = ctx->rscratch.p;
      c.ws.llk = sh_llk;
      c.ws.WZ = sh_WZ;
      c.ws.WQ = sh_WQ;
      c.ws.KsymQ = pass ? ctx->mixArenaQ.p + qoff : nullptr;
      qoff += (ksym_planes_bytes(ctx, s) + 1023) & ~(size_t)1023;
      c.stats_len = StatsLayout(d, mv.ks[j]).len;
      c.stats = pass ? ctx->mixStats.p + soff : nullptr;
      soff += round_up(c.stats_len, 4);
      if (pass) {
        c.m = stage_model_into(ctx, d, mv.ks[j], mv.C(j), mv.mu(j), mv.sigmas[j], Cpad, mupad, Ksym, c.ws.KsymQ,
                               c.ws.KsymScale, c.ws.colmax);
        c.m.ws = &c.ws;
        c.m.sigma_dev = sigmas_dev + j;
        em_zero(ctx, c.ws, c.plan, s, c.stats);
      }
    }
    need = cv.used;
    need_q = qoff;
  }
  out.stats_all = ctx->mixStats.p;
  out.stats_total = stats_total + 1;
  CUDA_CHECK(cudaMemsetAsync(ctx->mixStats.p, 0, sizeof(double) * (stats_total + 1), ctx->stream));  // alignment gaps too
  double *llk_sum = ctx->mixStats.p + stats_total;
  CUDA_CHECK(cudaMemsetAsync(llk_sum, 0, sizeof(double), ctx->stream));
  {
    const double ninf = -std::numeric_limits<double>::infinity();
    std::vector<double> init((size_t)M, ninf);
    // small synchronous upload (M doubles) before any kernel of the pass reads it
    CUDA_CHECK(cudaMemcpyAsync(out.run_max, init.data(), sizeof(double) * M, cudaMemcpyHostToDevice, ctx->stream));
    CUDA_CHECK(cudaMemcpyAsync(sigmas_dev, mv.sigmas, sizeof(double) * M, cudaMemcpyHostToDevice, ctx->stream));
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  }
  ctx->logw.reserve((size_t)M);
  CUDA_CHECK(cudaMemcpyAsync(ctx->logw.p, mv.logw, sizeof(double) * M, cudaMemcpyHostToDevice, ctx->stream));

  // ---- the chunk loop: everything below runs on the context's stream from device-resident inputs only, with launch
  // parameters that depend on shapes alone (the models' sigma is read from sigmas_dev), so it can be replayed from a graph
  auto run_loop = [&] {
    int64_t rows_total = 0;
    for (int64_t row0 = 0; row0 < st.n; row0 += chunk) {
      const int rows = (int)std::min<int64_t>(chunk, st.n - row0);
      const int rows_pad = (int)round_up(rows, 256);
      rows_total += rows;
      // E-step of every component on this chunk (unweighted: W = V = z z^T + Sigma)
      for (int j = 0; j < M; ++j) {
        MixComp &c = out.comps[j];
        e_step_chunk(ctx, st, nullptr, row0, rows, c.m, 2, sh_llk, nullptr, nullptr, nullptr);
        const int blocks = (int)std::min<int64_t>((int64_t)ctx->sms * 8, (rows + 255) / 256);
        strided_copy_kernel<<<blocks, 256, 0, ctx->stream>>>(sh_llk, rows, 1, 1, LPc + j, M);
        CUDA_CHECK(cudaGetLastError());
        ++ctx->launches;
      }
      // log-posteriors of the chunk (mix.rs:179-189), mixture log-likelihood (:162-174), chunk maxima of ln w + lp (:312-318)
      launch_log_softmax_rows(L, LPc, rows, M, ctx->logw.p, ds->w.p + row0, mixllk, chunk_max, nullptr);
      launch_weighted_sum(L, mixllk, ds->w.p + row0, rows, llk_sum, 1);
      if (LP_full)  // kept for the per-component repeats of the precision ladder
        CUDA_CHECK(cudaMemcpyAsync(LP_full + row0 * M, LPc, sizeof(double) * (size_t)rows * M, cudaMemcpyDeviceToDevice,
                                   ctx->stream));
      launch_mix_update_max(L, M, out.run_max, chunk_max, factor);
      for (int j = 0; j < M; ++j) {
        MixComp &c = out.comps[j];
        const Shape &s = shp[j];
        if (row0 > 0) {  // bring what this component has accumulated so far onto the new maximum
          const StatsLayout lay(d, s.k);
          launch_scale_by(L, c.stats, lay.offScalars, factor + j, 0);
          launch_scale_by(L, c.ws.part_bg, (int64_t)c.plan.bglen, factor + j, 0);
          launch_scale_by(L, c.ws.part_cr, (int64_t)c.plan.crlen, factor + j, 0);
          launch_scale_by(L, c.ws.part_solve, (int64_t)SOLVE_SLOTS * 4, factor + j, 1);
          launch_scale_by(L, c.ws.MZ, (int64_t)s.d * s.kp, factor + j, 0);
          launch_scale_by(L, c.ws.part_mz, (int64_t)c.plan.mzlen, factor + j, 0);
          launch_scale_by(L, c.ws.part_rx, (int64_t)c.plan.rxlen, factor + j, 0);
        }
        // column maxima of W fused into the weighting kernel (its shared-memory scratch holds 8 rows of W)
        const bool tc = ctx->gemm_mode == 2 && (size_t)8 * s.kkp * sizeof(double) <= 200 * 1024;
        if (tc) CUDA_CHECK(cudaMemsetAsync(c.ws.colmax, 0, sizeof(unsigned long long) * s.kkp, ctx->stream));
        ctx->span_begin(FAM_SOLVE);
        launch_mix_weight(L, LPc, M, j, ds->w.p + row0, out.run_max, rows, rows_pad, s.kkp, s.kp, c.ws.GW, c.ws.YZ, sh_WZ,
                          sh_r, tc ? c.ws.colmax : nullptr);
        launch_solve_reduce(L, rows, nullptr, c.ws.tn, c.ws.dv, c.ws.nx, st.dn.p + row0, sh_r, c.ws.part_solve, c.ws.rscratch,
                            c.ws.rflag);
        ctx->span_end();
        m_step_chunk(ctx, st, sh_r, row0, rows, c.m, c.stats, c.plan, tc);
      }
    }
    for (int j = 0; j < M; ++j) em_end(ctx, out.comps[j].m, out.comps[j].stats, out.comps[j].plan, rows_total);
  };
  static const bool graphs_env = !(getenv("PPCA_B200_GRAPHS") && !strcmp(getenv("PPCA_B200_GRAPHS"), "0"));
  const bool graphable = graphs_env && ctx->use_graphs && !ctx->profiling;
  if (!graphable) {
    run_loop();
    return;
  }
  // key: every pointer and scalar the captured launches bake in
  std::vector<uint64_t> key;
  auto kp = [&](const void *p) { key.push_back((uint64_t)(uintptr_t)p); };
  auto kv = [&](int64_t v) { key.push_back((uint64_t)v); };
  kp(st.X.p); kp(st.mask.p); kp(st.maskT.p); kp(st.dn.p); kp(ds->w.p); kp(ctx->mixArena.p); kp(ctx->mixArenaQ.p);
  kp(ctx->mixStats.p); kp(LP_full); kp(ctx->logw.p); kp(ctx->unsafe.p); kp(ctx->rscratch.p); kp(ctx->stream);
  kv(st.n); kv(st.d); kv(M); kv(chunk); kv(ctx->gemm_mode); kv(ctx->slices); kv(guard_on(ctx) ? ctx->guard_bits : -1);
  kv((int64_t)need); kv((int64_t)need_q);
  for (int j = 0; j < M; ++j) kv(mv.ks[j]);
  for (auto &g : ctx->mix_graphs)
    if (g.key == key) {
      CUDA_CHECK(cudaGraphLaunch(g.exec, ctx->stream));
      ctx->launches += g.launches;
      for (int v = 0; v < V_COUNT; ++v) ctx->variants[v] += g.variants[v];
      ++ctx->graph_replays;
      ++ctx->variants[V_GRAPH_REPLAYS];
      return;
    }
  bool seen = false;
  for (auto &k2 : ctx->mix_seen) seen = seen || k2 == key;
  if (!seen) {  // first pass with this key runs eagerly (one-time kernel attributes, buffer growth), the next one is captured
    if (ctx->mix_seen.size() >= 8) ctx->mix_seen.erase(ctx->mix_seen.begin());
    ctx->mix_seen.push_back(key);
    run_loop();
    return;
  }
  const int64_t launches0 = ctx->launches;
  int64_t variants0[V_COUNT];
  for (int v = 0; v < V_COUNT; ++v) variants0[v] = ctx->variants[v];
  cudaGraph_t graph = nullptr;
  ppca_b200_ctx::MixGraph g;
  g.key = key;
  bool captured = false;
  if (cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeRelaxed) == cudaSuccess) {
    try {
      run_loop();
      captured = true;
    } catch (...) {
    }
    const cudaError_t ee = cudaStreamEndCapture(ctx->stream, &graph);
    captured = captured && ee == cudaSuccess && graph != nullptr;
    if (captured) captured = cudaGraphInstantiate(&g.exec, graph, 0) == cudaSuccess;
    if (graph) cudaGraphDestroy(graph);
  }
  if (!captured) {  // something in the loop is not capturable on this driver: run it the ordinary way from now on
    cudaGetLastError();
    ctx->launches = launches0;
    for (int v = 0; v < V_COUNT; ++v) ctx->variants[v] = variants0[v];
    ctx->use_graphs = false;
    run_loop();
    return;
  }
  g.launches = ctx->launches - launches0;
  for (int v = 0; v < V_COUNT; ++v) g.variants[v] = ctx->variants[v] - variants0[v];
  CUDA_CHECK(cudaGraphLaunch(g.exec, ctx->stream));
  ++ctx->graph_replays;
  ++ctx->variants[V_GRAPH_REPLAYS];
  if (ctx->mix_graphs.size() >= 4) {
    // the oldest graph may still be executing on the stream: drain before destroying it
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    cudaGraphExecDestroy(ctx->mix_graphs.front().exec);
    ctx->mix_graphs.erase(ctx->mix_graphs.begin());
  }
  ctx->mix_graphs.push_back(std::move(g));
}

void mix_iterate_impl(ppca_b200_ctx *ctx, const ppca_b200_dataset *ds, int32_t m, const int32_t *ks, const double *Cs,
                      const double *mus, const double *sigmas, const double *log_weights, const ppca_b200_prior *prior,
                      double *Cs_out, double *mus_out, double *sigmas_out, double *log_weights_out, double *llk_in,
                      bool sharded) {
  check_ds(ctx, ds);
  MixView mv{m, ks, Cs, mus, sigmas, log_weights, ds->store->d};
  check_mix(mv);
  REQUIRE(Cs_out && mus_out && sigmas_out && log_weights_out, "null output");
  REQUIRE(!sharded || ctx->comm != nullptr, "no communicator: call ppca_b200_comm_init first");
  const SampleStore &st = *ds->store;
  if (st.n == 0 && !sharded) PPCA_THROW(PPCA_ERR_EMPTY, "dataset not empty (mix.rs:315)");
  if (st.n > 0 && !(ds->min_w > 0.0))
    PPCA_THROW(PPCA_ERR_WEIGHTS, "mixture EM needs strictly positive weights (mix.rs:304-309,326)");
  DeviceGuard g(ctx->device);
  std::vector<double> cmax(m), sumw(m);
  double llk = 0.0;
  run_guarded(ctx, [&] {
    Rung rungs[3];
    const int nrungs = ladder(ctx, rungs);
    const bool can_climb = guard_on(ctx) && ctx->rung + 1 < nrungs;
    double *LP_full = nullptr;
    if (can_climb) {
      ctx->mixLP.reserve((size_t)std::max<int64_t>(st.n, 1) * m);
      LP_full = ctx->mixLP.p;
    }
    MixPass pass;
    mix_em_pass(ctx, ds, mv, pass, LP_full);
    ctx->mixMax.reserve((size_t)2 * m);
    double *gmax = ctx->mixMax.p;  // global maxima (= local ones in a single process)
    CUDA_CHECK(cudaMemcpyAsync(gmax, pass.run_max, sizeof(double) * m, cudaMemcpyDeviceToDevice, ctx->stream));
    if (sharded) {
      comm_allreduce(ctx->comm, gmax, m, 1, ctx->stream);
      for (int j = 0; j < m; ++j) {
        const StatsLayout lay(st.d, ks[j]);
        launch_mix_rescale_stats(ctx->L(), pass.comps[j].stats, lay.offScalars, pass.comps[j].stats + lay.offScalars,
                                 pass.run_max, gmax, j);
      }
      comm_allreduce(ctx->comm, pass.stats_all, pass.stats_total, 0, ctx->stream);  // every component + the llk at once
    }
    CUDA_CHECK(cudaMemcpyAsync(cmax.data(), gmax, sizeof(double) * m, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_CHECK(cudaMemcpyAsync(&llk, pass.stats_all + pass.stats_total - 1, sizeof(double), cudaMemcpyDeviceToHost,
                               ctx->stream));
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    double viol_e = 0.0, viol_m = 0.0;
    std::vector<double> vm(m, 0.0);
    for (int j = 0; j < m; ++j) {
      const size_t off = (size_t)(mv.C(j) - Cs);
      double v2[2] = {0.0, 0.0};
      em_finish_impl(ctx, st.d, ks[j], mv.C(j), mv.mu(j), sigmas[j], prior, pass.comps[j].stats, Cs_out + off,
                     mus_out + (size_t)j * st.d, sigmas_out + j, nullptr, &sumw[j], v2);
      viol_e = std::max(viol_e, v2[0]);  // one counter for the whole pass (the solve kernels of every component)
      vm[j] = v2[1];
      viol_m += v2[1];
    }
    if (getenv("PPCA_B200_DEBUG")) {
      int nflag = 0;
      for (int j = 0; j < m; ++j) nflag += vm[j] > 0.0;
      fprintf(stderr, "[ppca_b200] mixture pass at rung %d: E-step violations %.0f, components with M-step violations %d/%d\n",
              ctx->rung, viol_e, nflag, m);
    }
    if (viol_e > 0.0 || !can_climb) return viol_e + viol_m;  // the whole pass climbs (or nothing can)
    // Only M-step violations: a component whose second moments sit far below their column scales (few effective samples).
    // Repeat THAT component alone, as a weighted single-model pass from the kept log-posteriors, one rung up at a time.
    const int base_mode = ctx->gemm_mode, base_slices = ctx->slices;
    for (int j = 0; j < m; ++j) {
      if (!(vm[j] > 0.0)) continue;
      const size_t off = (size_t)(mv.C(j) - Cs);
      const int64_t slen = StatsLayout(st.d, ks[j]).len;
      for (int r = ctx->rung + 1; r < nrungs; ++r) {
        ctx->gemm_mode = rungs[r].mode;
        ctx->slices = rungs[r].slices;
        guard_reset(ctx);
        ++ctx->variants[V_PRECISION_RETRY];
        ctx->rbuf.reserve((size_t)st.n_pad);
        CUDA_CHECK(cudaMemsetAsync(ctx->rbuf.p, 0, sizeof(double) * st.n_pad, ctx->stream));
        launch_responsibilities(ctx->L(), LP_full, st.n, m, j, ds->w.p, cmax[j], ctx->rbuf.p);
        DevModel dm = stage_model(ctx, st.d, ks[j], mv.C(j), mv.mu(j), sigmas[j]);
        ctx->stats.reserve((size_t)slen);
        em_stats_impl(ctx, st, ctx->rbuf.p, dm, ctx->stats.p);
        if (sharded) comm_allreduce(ctx->comm, ctx->stats.p, slen, 0, ctx->stream);
        const double v = em_finish_impl(ctx, st.d, ks[j], mv.C(j), mv.mu(j), sigmas[j], prior, ctx->stats.p, Cs_out + off,
                                        mus_out + (size_t)j * st.d, sigmas_out + j, nullptr, &sumw[j]);
        if (!(v > 0.0)) break;
      }
    }
    ctx->gemm_mode = base_mode;
    ctx->slices = base_slices;
    return 0.0;
  });
  if (llk_in) *llk_in = llk;
  std::vector<double> logsum(m);
  for (int j = 0; j < m; ++j) logsum[j] = std::log(sumw[j]) + cmax[j];  // mix.rs:323-324
  // robust_log_softmax (mix.rs:14-18, :335)
  double mx = logsum[0];
  for (int j = 1; j < m; ++j) mx = logsum[j] > mx ? logsum[j] : mx;
  double sm = 0.0;
  for (int j = 0; j < m; ++j) sm += std::exp(logsum[j] - mx);
  const double ln = std::log(sm);
  for (int j = 0; j < m; ++j) log_weights_out[j] = logsum[j] - mx - ln;
}

}  // namespace

extern "C" {

int32_t ppca_b200_mix_iterate(ppca_b200_ctx *ctx, const ppca_b200_dataset *ds, int32_t m, const int32_t *ks,
                              const double *Cs, const double *mus, const double *sigmas, const double *log_weights,
                              const ppca_b200_prior *prior, double *Cs_out, double *mus_out, double *sigmas_out,
                              double *log_weights_out, double *llk_in) {
  return guarded([&] {
    mix_iterate_impl(ctx, ds, m, ks, Cs, mus, sigmas, log_weights, prior, Cs_out, mus_out, sigmas_out, log_weights_out,
                     llk_in, false);
  });
}

// ---- sample-sharded EM with the collective inside the library (one process per GPU) ----------------------------------
int32_t ppca_b200_comm_unique_id(uint8_t *out) {
  return guarded([&] {
    REQUIRE(out != nullptr, "null output");
    comm_unique_id(out);
  });
}

int32_t ppca_b200_comm_init(ppca_b200_ctx *ctx, const uint8_t *unique_id, int32_t rank, int32_t world) {
  return guarded([&] {
    REQUIRE(ctx != nullptr && unique_id != nullptr, "null argument");
    REQUIRE(world >= 1 && rank >= 0 && rank < world, "bad rank %d of %d", rank, world);
    REQUIRE(ctx->comm == nullptr, "the context already has a communicator");
    DeviceGuard g(ctx->device);
    ctx->comm = comm_create(unique_id, rank, world);
    ctx->comm_rank = rank;
    ctx->comm_world = world;
  });
}

int32_t ppca_b200_comm_destroy(ppca_b200_ctx *ctx) {
  return guarded([&] {
    REQUIRE(ctx != nullptr, "null context");
    if (!ctx->comm) return;
    DeviceGuard g(ctx->device);
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    comm_destroy(ctx->comm);
    ctx->comm = nullptr;
    ctx->comm_rank = 0;
    ctx->comm_world = 1;
  });
}

int32_t ppca_b200_comm_allreduce(ppca_b200_ctx *ctx, double *buf_dev, int64_t count, int32_t op) {
  return guarded([&] {
    REQUIRE(ctx != nullptr && ctx->comm != nullptr, "no communicator: call ppca_b200_comm_init first");
    REQUIRE(buf_dev != nullptr && count >= 0 && (op == 0 || op == 1), "bad all-reduce arguments");
    DeviceGuard g(ctx->device);
    comm_allreduce(ctx->comm, buf_dev, count, op, ctx->stream);
  });
}

int32_t ppca_b200_iterate_sharded(ppca_b200_ctx *ctx, const ppca_b200_dataset *ds, int32_t k, const double *C,
                                  const double *mu, double sigma, const ppca_b200_prior *prior, double *C_out,
                                  double *mu_out, double *sigma_out, double *llk_in) {
  return guarded([&] {
    check_ds(ctx, ds);
    REQUIRE(ctx->comm != nullptr, "no communicator: call ppca_b200_comm_init first");
    const SampleStore &st = *ds->store;
    DeviceGuard g(ctx->device);
    const int64_t slen = StatsLayout(st.d, k).len;
    run_guarded(ctx, [&] {  // the guard counters ride in the reduced buffer: every rank repeats (or accepts) together
      DevModel m = stage_model(ctx, st.d, k, C, mu, sigma);
      ctx->stats.reserve((size_t)slen);
      em_stats_impl(ctx, st, ds->w.p, m, ctx->stats.p);
      comm_allreduce(ctx->comm, ctx->stats.p, slen, 0, ctx->stream);
      return em_finish_impl(ctx, st.d, k, C, mu, sigma, prior, ctx->stats.p, C_out, mu_out, sigma_out, llk_in, nullptr);
    });
  });
}

int32_t ppca_b200_iterate_host_sharded(ppca_b200_ctx *ctx, const double *x, int64_t n, int32_t d, const double *weights,
                                       int32_t k, const double *C, const double *mu, double sigma,
                                       const ppca_b200_prior *prior, double *C_out, double *mu_out, double *sigma_out,
                                       double *llk_in) {
  return guarded([&] {
    REQUIRE(ctx != nullptr && ctx->comm != nullptr, "no communicator: call ppca_b200_comm_init first");
    REQUIRE(n >= 0 && d >= 1, "bad dataset shape %lld x %d", (long long)n, d);
    REQUIRE(n == 0 || x != nullptr, "null data");
    DeviceGuard g(ctx->device);
    const int64_t slen = StatsLayout(d, k).len;
    run_guarded(ctx, [&] {
      DevModel m = stage_model(ctx, d, k, C, mu, sigma);
      ctx->stats.reserve((size_t)slen);
      if (n == 0) CUDA_CHECK(cudaMemsetAsync(ctx->stats.p, 0, sizeof(double) * slen, ctx->stream));
      else em_stats_host_impl(ctx, x, n, d, weights, m, ctx->stats.p);
      comm_allreduce(ctx->comm, ctx->stats.p, slen, 0, ctx->stream);
      return em_finish_impl(ctx, d, k, C, mu, sigma, prior, ctx->stats.p, C_out, mu_out, sigma_out, llk_in, nullptr);
    });
  });
}

// EM step over rows [row_begin, row_begin + nrows) of the synthetic dataset (ppca_b200_dataset_synthetic with the same
// d, k_true, sigma_true, mask_prob, seed and a single component) WITHOUT storing them: every chunk is regenerated on the
// device (counter-based RNG keyed by the global row), ingested into a chunk-sized store and consumed.  This is how a job
// larger than HBM (BASELINE configs[2]: N = 100 M x d = 2048 = 1.6 TB) runs at kernel speed on any number of GPUs:
// `sharded` != 0 all-reduces the statistics over the context's communicator before the finish, as
// ppca_b200_iterate_sharded does.  The result equals ppca_b200_iterate on the stored rows up to summation order.
int32_t ppca_b200_iterate_generated(ppca_b200_ctx *ctx, int64_t row_begin, int64_t nrows, int32_t d, int32_t k_true,
                                    double sigma_true, double mask_prob, uint64_t seed, int32_t k, const double *C,
                                    const double *mu, double sigma, const ppca_b200_prior *prior, int32_t sharded,
                                    double *C_out, double *mu_out, double *sigma_out, double *llk_in) {
  return guarded([&] {
    REQUIRE(ctx != nullptr, "null context");
    REQUIRE(row_begin >= 0 && nrows >= 0 && d >= 1 && k_true >= 1, "bad synthetic shape");
    REQUIRE(mask_prob >= 0.0 && mask_prob <= 1.0, "invalid mask probability");
    REQUIRE(!sharded || ctx->comm != nullptr, "no communicator: call ppca_b200_comm_init first");
    if (nrows == 0 && !sharded) PPCA_THROW(PPCA_ERR_EMPTY, "non-empty dataset required (ppca_model.rs:358)");
    DeviceGuard g(ctx->device);
    const int64_t slen = StatsLayout(d, k).len;
    DevBuf<double> &Ct = ctx->gen_Ct, &mut = ctx->gen_mut;
    Ct.reserve((size_t)d * k_true);
    mut.reserve((size_t)d);
    launch_synth_truth(ctx->L(), d, k_true, 1, seed, Ct.p, mut.p);
    run_guarded(ctx, [&] {
      DevModel m = stage_model(ctx, d, k, C, mu, sigma);
      ctx->stats.reserve((size_t)slen);
      if (nrows == 0) {
        CUDA_CHECK(cudaMemsetAsync(ctx->stats.p, 0, sizeof(double) * slen, ctx->stream));
      } else {
        const EmPlan p = em_begin(ctx, pick_chunk(ctx, round_up(nrows, 256), m.s), m, ctx->stats.p);
        const int64_t blk = p.chunk, tail = nrows % blk;
        if (nrows >= blk && (!ctx->s_store || ctx->s_store->n != blk || ctx->s_store->d != d)) ctx->s_store = make_store(ctx, blk, d);
        if (tail && (!ctx->s_tail || ctx->s_tail->n != tail || ctx->s_tail->d != d)) ctx->s_tail = make_store(ctx, tail, d);
        ctx->s_w.reserve((size_t)round_up(std::min(nrows, blk), 256));
        DevBuf<double> &gws = ctx->gen_ws;
        gws.reserve(synth_ws_doubles(std::min(nrows, blk), d, k_true));
        const Launcher L = ctx->L();
        for (int64_t r0 = 0; r0 < nrows; r0 += blk) {
          const int64_t rows = std::min<int64_t>(blk, nrows - r0);
          SampleStore &st = rows == blk ? *ctx->s_store : *ctx->s_tail;
          launch_generate_block(L, st, rows, row_begin + r0, k_true, Ct.p, mut.p, sigma_true, mask_prob, seed, gws.p);
          launch_transpose_mask(L, st);
          CUDA_CHECK(cudaMemsetAsync(ctx->s_w.p, 0, sizeof(double) * st.n_pad, ctx->stream));
          const int64_t want = (rows + 255) / 256;
          const int blocks = (int)(want < (int64_t)ctx->sms * 8 ? want : (int64_t)ctx->sms * 8);
          fill_value_kernel<<<blocks, 256, 0, ctx->stream>>>(ctx->s_w.p, rows, 1.0);
          CUDA_CHECK(cudaGetLastError());
          ++ctx->launches;
          em_chunk(ctx, st, ctx->s_w.p, 0, (int)rows, m, ctx->stats.p, p);
        }
        em_end(ctx, m, ctx->stats.p, p);
      }
      if (sharded) comm_allreduce(ctx->comm, ctx->stats.p, slen, 0, ctx->stream);
      return em_finish_impl(ctx, d, k, C, mu, sigma, prior, ctx->stats.p, C_out, mu_out, sigma_out, llk_in, nullptr);
    });
  });
}

int32_t ppca_b200_pack_host(const double *x, int64_t n, int32_t d, double *vals, int64_t *rowptr, uint32_t *maskw) {
  return guarded([&] {
    REQUIRE(n >= 0 && d >= 1 && rowptr != nullptr && (n == 0 || x != nullptr), "bad arguments");
    const int dw = (d + 31) / 32;
    // pass 1: per-row counts (threads over row ranges), pass 2: exclusive scan, pass 3: scatter (vals / maskw nullable)
    const int nthr = (int)std::max<int64_t>(1, std::min<int64_t>(16, n / 4096));
    rowptr[0] = 0;
    auto rows_of = [&](int t, int64_t &lo, int64_t &hi) {
      lo = n * t / nthr;
      hi = n * (t + 1) / nthr;
    };
    {
      std::vector<std::thread> th;
      for (int t = 0; t < nthr; ++t)
        th.emplace_back([&, t] {
          int64_t lo, hi;
          rows_of(t, lo, hi);
          for (int64_t r = lo; r < hi; ++r) {
            int64_t c = 0;
            const double *row = x + r * d;
            for (int i = 0; i < d; ++i) c += std::isfinite(row[i]) ? 1 : 0;
            rowptr[r + 1] = c;
          }
        });
      for (auto &t : th) t.join();
    }
    for (int64_t r = 0; r < n; ++r) rowptr[r + 1] += rowptr[r];
    if (!vals || !maskw) return;  // counting call: rowptr[n] = number of observed values
    {
      std::vector<std::thread> th;
      for (int t = 0; t < nthr; ++t)
        th.emplace_back([&, t] {
          int64_t lo, hi;
          rows_of(t, lo, hi);
          for (int64_t r = lo; r < hi; ++r) {
            const double *row = x + r * d;
            double *out = vals + rowptr[r];
            uint32_t *mw = maskw + r * dw;
            for (int j = 0; j < dw; ++j) mw[j] = 0u;
            for (int i = 0; i < d; ++i)
              if (std::isfinite(row[i])) {
                *out++ = row[i];
                mw[i >> 5] |= 1u << (i & 31);
              }
          }
        });
      for (auto &t : th) t.join();
    }
  });
}

static int32_t iterate_packed_host(ppca_b200_ctx *ctx, const double *vals, const int64_t *rowptr, const uint32_t *maskw,
                                   int64_t n, int32_t d, const double *weights, int32_t k, const double *C,
                                   const double *mu, double sigma, const ppca_b200_prior *prior, double *C_out,
                                   double *mu_out, double *sigma_out, double *llk_in, bool sharded) {
  return guarded([&] {
    REQUIRE(ctx != nullptr, "null context");
    REQUIRE(n >= 0 && d >= 1, "bad dataset shape %lld x %d", (long long)n, d);
    REQUIRE(!sharded || ctx->comm != nullptr, "no communicator: call ppca_b200_comm_init first");
    if (n == 0 && !sharded) PPCA_THROW(PPCA_ERR_EMPTY, "non-empty dataset required (ppca_model.rs:358)");
    REQUIRE(n == 0 || (rowptr != nullptr && maskw != nullptr && (vals != nullptr || rowptr[n] == 0)), "null data");
    DeviceGuard g(ctx->device);
    const int64_t slen = StatsLayout(d, k).len;
    PackedHost pk;
    pk.vals = vals;
    pk.rowptr = rowptr;
    pk.maskw = maskw;
    run_guarded(ctx, [&] {
      DevModel m = stage_model(ctx, d, k, C, mu, sigma);
      ctx->stats.reserve((size_t)slen);
      if (n == 0) CUDA_CHECK(cudaMemsetAsync(ctx->stats.p, 0, sizeof(double) * slen, ctx->stream));
      else em_stats_host_impl(ctx, nullptr, n, d, weights, m, ctx->stats.p, &pk);
      if (sharded) comm_allreduce(ctx->comm, ctx->stats.p, slen, 0, ctx->stream);
      return em_finish_impl(ctx, d, k, C, mu, sigma, prior, ctx->stats.p, C_out, mu_out, sigma_out, llk_in, nullptr);
    });
  });
}

int32_t ppca_b200_iterate_packed_host(ppca_b200_ctx *ctx, const double *vals, const int64_t *rowptr,
                                      const uint32_t *maskw, int64_t n, int32_t d, const double *weights, int32_t k,
                                      const double *C, const double *mu, double sigma, const ppca_b200_prior *prior,
                                      double *C_out, double *mu_out, double *sigma_out, double *llk_in) {
  return iterate_packed_host(ctx, vals, rowptr, maskw, n, d, weights, k, C, mu, sigma, prior, C_out, mu_out, sigma_out,
                             llk_in, false);
}

int32_t ppca_b200_iterate_packed_host_sharded(ppca_b200_ctx *ctx, const double *vals, const int64_t *rowptr,
                                              const uint32_t *maskw, int64_t n, int32_t d, const double *weights,
                                              int32_t k, const double *C, const double *mu, double sigma,
                                              const ppca_b200_prior *prior, double *C_out, double *mu_out,
                                              double *sigma_out, double *llk_in) {
  return iterate_packed_host(ctx, vals, rowptr, maskw, n, d, weights, k, C, mu, sigma, prior, C_out, mu_out, sigma_out,
                             llk_in, true);
}

int32_t ppca_b200_mix_iterate_sharded(ppca_b200_ctx *ctx, const ppca_b200_dataset *ds, int32_t m, const int32_t *ks,
                                      const double *Cs, const double *mus, const double *sigmas,
                                      const double *log_weights, const ppca_b200_prior *prior, double *Cs_out,
                                      double *mus_out, double *sigmas_out, double *log_weights_out, double *llk_in) {
  return guarded([&] {
    mix_iterate_impl(ctx, ds, m, ks, Cs, mus, sigmas, log_weights, prior, Cs_out, mus_out, sigmas_out, log_weights_out,
                     llk_in, true);
  });
}

}  // extern "C"
