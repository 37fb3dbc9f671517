// libchi: C ABI (include/chi.h) over the sm_100a kernels in kernels.cuh.
// Host-side logic only: argument checks, device workspace, launch geometry,
// host<->device copies.  There is no CPU compute path in this file.
#include <cuda_runtime.h>
#include <sched.h>
#include <sys/syscall.h>
#include <unistd.h>
#include <cctype>
#include <cstring>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <vector>

#include "../../include/chi.h"
#include "kernels.cuh"
#include "launch.cuh"
#include "moments_ea.cuh"
#include "nuts.cuh"
#include "comm.cuh"
#include "nuts_warp.cuh"

#define CHI_STAGE_RING 64

using namespace chi;
namespace chi {
int launch_nuts_chain(int NL, int threads, bool pv, const MapCfg& cfg, const SmallArgs& a, const NutsState& s,
                      const NutsAsync& A, int max_ticks, size_t dyn_smem, cudaStream_t st);  // launch_chain.cu
int launch_nuts_chain_cl(int NL, bool pv, const MapCfg& cfg, const SmallArgs& a, const NutsState& s, const NutsAsync& A,
                         int max_ticks, size_t gb_bytes, cudaStream_t st);
}

namespace {

std::string g_create_error;

struct Buf {
  void* p = nullptr;
  size_t cap = 0;
  cudaError_t ensure(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    size_t want = bytes + bytes / 8 + 256;
    cudaError_t e = cudaMalloc(&p, want);
    if (e == cudaSuccess) cap = want;
    return e;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
  double* d() const { return (double*)p; }
};

// scope guards of the entry points that own temporaries (chi_nuts_sample): every early return
// (CK, fail) releases them
struct ScopedBuf : Buf {
  ~ScopedBuf() { release(); }
};
struct ScopedPinned {
  int* p = nullptr;
  ~ScopedPinned() { if (p) cudaFreeHost(p); }
};
struct ScopedGraph {
  cudaGraph_t g = nullptr;
  cudaGraphExec_t x = nullptr;
  ~ScopedGraph() {
    if (x) cudaGraphExecDestroy(x);
    if (g) cudaGraphDestroy(g);
  }
};

static RowMap identity_rows() {
  RowMap m;
  m.n_rows = CHI_NSLOT;
  for (int s = 0; s < CHI_NSLOT; ++s) m.row[s] = (signed char)s;
  return m;
}

#define CHI_TICKETS 8
#define CHI_SUMS_MAXPIECES 64

struct DevDataset {
  int S = 0, nb = 0;
  bool has_offset = false;
  Buf vax, chan, basis, offset;  // chan: packed v / p_0 / y / w planes (ChanArgs::chan)
  Buf vax32;                     // (hi, lo) float pair per channel: the single-precision spectra kernels
  void release() { vax.release(); chan.release(); basis.release(); offset.release(); vax32.release(); }
};

// One pipeline lane: a stream plus the device workspace a batch needs.  Lane 0 uses the
// handle's public stream; the host entry points rotate sub-batches over all lanes so that
// host->device copies, kernels and device->host copies of consecutive sub-batches overlap.
#define CHI_LANES 3
struct Lane {
  cudaStream_t st = nullptr;
  // small batches run the absorption kernel on `aux`, concurrently with the emission kernel
  cudaStream_t aux = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  // small batches: the Jacobian of the parameter maps (k_jacobian) runs on `jac` next to the channel kernels
  cudaStream_t jac = nullptr;
  cudaEvent_t ev_jac = nullptr;
  Buf jacbuf;
  Buf sums_partial, sums_counter, sums_out;  // per-piece shard sums of the host "sums" entry points
  Buf cnt;            // small batches: blocks-done counter per draw (k_eval_small), zero between launches
  size_t cnt_zeroed = 0;
  // pipelined host path: the epilogue and the downloads of a piece run on a high-priority stream, so
  // that they are not starved by the channel kernels of the following pieces (which were queued
  // first and would otherwise keep every SM until they drain) and the lane is free again early
  cudaStream_t hi = nullptr;
  cudaEvent_t ev_mom = nullptr, ev_tail = nullptr;
  bool tail_pending = false;
  Buf cc, mom_e, mom_a;                                           // kernel workspace
  Buf xcon, prior;                                                // constrained values, per-cloud priors
  Buf w_theta, w_ce, w_ca, w_sl, w_logp, w_grad, w_gce, w_gca;    // staging of the host entry points
  void release() {
    cnt_zeroed = 0;
    Buf* bufs[] = {&cc, &mom_e, &mom_a, &xcon, &prior, &jacbuf, &cnt, &sums_partial, &sums_counter, &sums_out, &w_theta, &w_ce, &w_ca, &w_sl, &w_logp, &w_grad, &w_gce, &w_gca};
    for (Buf* b : bufs) b->release();
  }
};

}  // namespace

struct chi_handle {
  chi_config cfg;
  MapCfg map;
  int device = 0;
  int n_sm = 148;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  // ring of stage-boundary events (map | em | abs | epilogue) of the last CHI_STAGE_RING batches
  cudaEvent_t ring[CHI_STAGE_RING][5];
  int64_t stage_calls = 0;
  bool ev_pending = false;
  float kernel_ms = 0.f;
  bool has_e = false, has_a = false, spectra_set = false;
  int n_sl = 0;
  DevDataset e, a;
  Buf lconst;
  Lane lanes[CHI_LANES];
  // further staging of the host entry points (single-lane paths)
  Buf w_gbe, w_gba, w_mue, w_mua, w_misc[8];
  RowMap rm;                 // layout of the caller's theta / grad blocks for the call in flight
  // asynchronous host calls: completion events of the last CHI_TICKETS calls, two per lane
  cudaEvent_t tickets[CHI_TICKETS][2 * CHI_LANES] = {};
  int64_t next_ticket = 0;
  bool shared_axis = false;  // both datasets on the same spectral axis: fused kernel
  Buf chan;                  // its packed per-channel data [n_sl][S][8]
  bool priors_set = false;
  bool capturing = false;  // inside a stream capture: no timing events
  bool graph_path = false; // the evaluations of this call are replayed from a CUDA graph (the sampler)
  double bsig_e[CHI_MAX_BASELINE], bsig_a[CHI_MAX_BASELINE];
  int64_t call_draws = 0;        // draws of the host call in flight (0 outside the pipelined host path)
  bool epilogue_on_hi = false;   // pipelined host path: epilogue + downloads on the lane's priority stream
  int map_ahead = 0;             // PHASE_MAP: ring slots between the piece being mapped and stage_calls
  int64_t call_first_stage = 0;  // stage_calls at the start of the current model-level call
  int64_t launches = 0;
  int64_t opt[CHI_OPT_COUNT] = {};  // chi_set_option
  int64_t dev_calls = 0;            // CHI_OPT_DEV_OVERLAP: device-resident calls issued (lane rotation)
  // two-launch forms of the large-batch kernels (moments_ea.cuh, MODE): the EA_FIXUP launch counts the draws it
  // evaluated; the count is copied to pinned memory behind it and read, without waiting, by later calls -- a
  // handle whose batches are full of marked draws goes back to the one-launch form (see two_pass_ok)
  Buf marked_dev;                                 // [CHI_LANES] counters
  unsigned long long* marked_host = nullptr;      // pinned [CHI_LANES]
  int64_t marked_B[CHI_LANES] = {};               // draws of the call whose count is (being) copied there
  int64_t dense_calls = 0;
  cudaEvent_t ev_dev_in = nullptr;
  // multi-GPU (comm.cuh): communicator, its stream, and the workspace of k_shard_sums
  NcclComm comm = nullptr;
  int comm_rank = 0, comm_size = 1;
  cudaStream_t comm_stream = nullptr;
  cudaEvent_t ev_comm_in = nullptr, ev_comm_out = nullptr;
  Buf sums_partial, sums_counter;
  // host "sums" entry points: per-piece sums land in pinned staging; the host adds them in piece order
  // when the call (or its ticket) completes
  double* sums_stage = nullptr;  // [CHI_TICKETS + 1][CHI_SUMS_MAXPIECES][CHI_SUMS_MAXW + 2], pinned
  struct PendingSums { double* out = nullptr; int pieces = 0, width = 0; } pending[CHI_TICKETS + 1];
  std::string err;
};

namespace {

bool make_ring(chi_handle* h);

// ends a stream capture that an error path would otherwise leave open (the stream and the handle
// would be unusable afterwards)
struct CaptureGuard {
  chi_handle* h;
  cudaStream_t st;
  bool active = false;
  CaptureGuard(chi_handle* h_, cudaStream_t st_) : h(h_), st(st_) {}
  cudaError_t begin() {
    cudaError_t e = cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal);
    if (e == cudaSuccess) { active = true; h->capturing = true; }
    return e;
  }
  cudaError_t end(cudaGraph_t* g) {
    active = false;
    h->capturing = false;
    return cudaStreamEndCapture(st, g);
  }
  ~CaptureGuard() {
    if (!active) return;
    cudaGraph_t g = nullptr;
    cudaStreamEndCapture(st, &g);
    if (g) cudaGraphDestroy(g);
    h->capturing = false;
  }
};

int fail(chi_handle* h, int code, const std::string& msg) {
  if (h) h->err = msg; else g_create_error = msg;
  return code;
}

#define CK(call)                                                                          \
  do {                                                                                    \
    cudaError_t _e = (call);                                                              \
    if (_e != cudaSuccess)                                                                \
      return fail(h, _e == cudaErrorMemoryAllocation ? CHI_ENOMEM : CHI_ECUDA,            \
                  std::string(#call) + ": " + cudaGetErrorString(_e));                    \
  } while (0)

bool make_ring(chi_handle* h) {
  for (auto& row : h->ring)
    for (cudaEvent_t& e : row) e = nullptr;
  for (auto& row : h->ring)
    for (cudaEvent_t& e : row)
      if (cudaEventCreate(&e) != cudaSuccess) return false;
  return true;
}

bool kind_has_emission(int m) {
  return m == CHI_MODEL_EMISSION || m == CHI_MODEL_EMISSION_ABSORPTION ||
         m == CHI_MODEL_EMISSION_PHYSICAL || m == CHI_MODEL_EMISSION_ABSORPTION_PHYSICAL;
}
bool kind_has_absorption(int m) {
  return m == CHI_MODEL_ABSORPTION || m == CHI_MODEL_EMISSION_ABSORPTION ||
         m == CHI_MODEL_ABSORPTION_PHYSICAL || m == CHI_MODEL_EMISSION_ABSORPTION_PHYSICAL;
}

// moment records are sized for the (possibly padded) instantiation that runs
int em_record(const chi_handle* h) {
  return 1 + h->e.nb + padded_clouds(h->cfg.n_clouds) * (h->cfg.lorentzian ? 8 : 5);
}
int ab_record(const chi_handle* h) {
  return 1 + h->a.nb + padded_clouds(h->cfg.n_clouds) * (h->cfg.lorentzian ? 6 : 3);
}

// channel chunking: enough warps to fill the machine when B is small
void pick_chunks(const chi_handle* h, int64_t B, int S, int* n_chunks, int* chunk_len) {
  const int64_t target = (int64_t)h->n_sm * 16;  // warp items wanted per launch
  int64_t want = (target + B - 1) / B;
  static const int min_len = [] {  // tuning knob: shortest chunk, in channels
    const char* e = getenv("CHI_CHUNK_MIN");
    return e ? std::max(32, atoi(e)) : 128;
  }();
  int max_chunks = std::max(1, S / min_len);  // at least 4 channel iterations per lane
  int nc = (int)std::min<int64_t>(std::max<int64_t>(want, 1), max_chunks);
  int len = (S + nc - 1) / nc;
  len = (len + 31) / 32 * 32;
  *chunk_len = len;
  *n_chunks = (S + len - 1) / len;
}

#define MODEL_SWITCH(model, STMT)                                                                 \
  switch (model) {                                                                                \
    case CHI_MODEL_PHYSICS: { constexpr int M = CHI_MODEL_PHYSICS; STMT; } break;                 \
    case CHI_MODEL_ABSORPTION: { constexpr int M = CHI_MODEL_ABSORPTION; STMT; } break;           \
    case CHI_MODEL_EMISSION: { constexpr int M = CHI_MODEL_EMISSION; STMT; } break;               \
    case CHI_MODEL_EMISSION_ABSORPTION: { constexpr int M = CHI_MODEL_EMISSION_ABSORPTION; STMT; } break; \
    case CHI_MODEL_ABSORPTION_PHYSICAL: { constexpr int M = CHI_MODEL_ABSORPTION_PHYSICAL; STMT; } break; \
    case CHI_MODEL_EMISSION_PHYSICAL: { constexpr int M = CHI_MODEL_EMISSION_PHYSICAL; STMT; } break;     \
    case CHI_MODEL_EMISSION_ABSORPTION_PHYSICAL: { constexpr int M = CHI_MODEL_EMISSION_ABSORPTION_PHYSICAL; STMT; } break; \
  }

int launch_cloud_map(chi_handle* h, Lane& L, int64_t B, const double* theta, double* constrained = nullptr) {
  const int N = h->cfg.n_clouds;
  CK(L.cc.ensure((size_t)B * N * CHI_CC_STRIDE * sizeof(double)));
  unsigned grid = (unsigned)((B * N + 127) / 128);
  MODEL_SWITCH(h->cfg.model, (k_cloud_map<M><<<grid, 128, 0, L.st>>>(h->map, h->rm, B, N, theta, L.cc.d(), constrained)));
  h->launches++;
  CK(cudaGetLastError());
  return CHI_OK;
}

// ---- small batches: one launch per evaluation (small.cuh) ----------------------------------------
// A batch is "small" when its evaluation is bound by dependent latency rather than throughput (a few
// dozen chains).  The choice follows the size of the caller's batch, not of a pipeline piece.
bool batch_is_small(const chi_handle* h, int64_t B) {
  const int N = h->cfg.n_clouds;
  const int K = h->cfg.model == CHI_MODEL_PHYSICS ? 7 : CHI_NSLOT;
  const int64_t B_call = std::max<int64_t>(B, h->call_draws);
  return B_call * N * K <= (int64_t)h->n_sm * 1024 && N * K <= 128;
}
// k_eval_small (one launch: map + channels + Jacobian + contraction) against the pipeline
// (map -> channel kernel || Jacobian -> contraction), measured on B200 (tools/small_ab.py,
// profiles/r2_small_batch.txt):
//   * more than 8 clouds: the pipeline runs on the 12- / 16-cloud instantiations (255 registers,
//     kilobytes of spills); one launch is 2.4x faster (57.6 -> 23.9 us at 32 chains x 16 clouds);
//   * launched from the host one call at a time (the Op / Engine path): a call costs what its
//     launches cost the host, one launch instead of four is 44.0 against 53.6 us at 64 chains x
//     8 clouds x 2048+2048 channels, 25.9 against 38.8 us at 8 chains x 2 clouds;
//   * replayed from a CUDA graph (the sampler's ticks): launches are free and the pipeline's channel
//     kernel executes fewer instructions per channel (every cloud in one thread); with records and a
//     counter in global memory the one launch lost there (31.6 against 29.2 us); as a cluster it is level,
//     but not inside the sampler (see the end of this function).
// CHI_OPT_SMALL_PATH overrides (tests run both paths).
bool use_eval_small(const chi_handle* h, int64_t B, bool vjp) {
  if (vjp || h->epilogue_on_hi || h->cfg.n_clouds * CHI_NSLOT > CHI_SM_THREADS - 96) return false;
  const int64_t mode = h->opt[CHI_OPT_SMALL_PATH];
  if (mode == 1) return false;
  if (mode == 2) return true;
  if (!batch_is_small(h, B)) return false;
  if (h->cfg.n_clouds > 8) return true;
  // (graph_path = the sampler's graph of kernels per tick.  With the blocks of a draw meeting as a cluster the
  // one launch is level with the pipeline for a LIKELIHOOD evaluation replayed from a graph -- 28.8 against 28.9 us
  // at configs[2]'s shape, 11.2 against 13.9 at 8 chains x 2 clouds -- but the sampler evaluates the MODEL logp,
  // whose prior terms lengthen the Jacobian warps, and its tick kernel maps the point for the pipeline: measured
  // inside the sampler the pipeline stays ahead, 29 against 25 it/s at configs[2]'s size, 124 against 116 at 8
  // chains x 2 clouds.)
  return !h->graph_path && std::max<int64_t>(B, h->call_draws) <= 128;
}

// argument block of k_eval_small / k_nuts_chain for a batch; force_chunks > 0 fixes the blocks per draw
int small_args(chi_handle* h, Lane& L, int64_t B, const double* theta, const double* ce, const double* ca,
               const int32_t* sl, double* logp, double* grad, double* gce, double* gca, bool model, int force_chunks,
               SmallArgs* out, int* NL_out, size_t* dyn_smem_out, int threads = CHI_SM_THREADS) {
  const int N = h->cfg.n_clouds;
  const bool pv = h->cfg.lorentzian != 0;
  SmallArgs& a = *out;
  memset(&a, 0, sizeof(a));
  a.B = B; a.N = N; a.K = h->cfg.model == CHI_MODEL_PHYSICS ? 7 : CHI_NSLOT;
  a.rm = h->rm; a.prior = model ? 1 : 0; a.want_grad = grad ? 1 : 0;
  a.fused = h->shared_axis ? 1 : 0;
  a.has_e = h->has_e ? 1 : 0; a.has_a = h->has_a ? 1 : 0;
  a.S_e = h->has_e ? h->e.S : 0; a.S_a = h->has_a ? h->a.S : 0;
  a.nb_e = h->has_e ? h->e.nb : 0; a.nb_a = h->has_a ? h->a.nb : 0;
  a.bg = h->cfg.bg_temp;
  // clouds per thread (NL) and lane groups (2^lgG) of a channel: one cloud per thread, two for 5..8
  // clouds (measured: 31.6 us against 36.1 at configs[2] shape; more clouds per thread need more
  // registers than two resident blocks per SM leave)
  const int force_nl = (int)h->opt[CHI_OPT_SMALL_NL];
  const int S_max = std::max(a.S_e, a.S_a);
  int NL = (N >= 5 && N <= 8) ? 2 : 1;
  if (force_nl >= 1 && force_nl <= 2 && force_nl * 16 >= N) NL = force_nl;
  int lgG = 0;
  while ((NL << lgG) < N) ++lgG;
  a.lgG = lgG;
  const int NP = NL << lgG, cpw = 32 >> lgG;
  // blocks per draw: fill the resident block slots once, at least one warp tile per channel warp
  if (force_chunks <= 0) force_chunks = (int)h->opt[CHI_OPT_SMALL_CHUNKS];
  int64_t want = std::max<int64_t>(1, ((int64_t)h->n_sm * eval_small_blocks_per_sm(NL)) / B);
  if (force_chunks > 0) want = force_chunks;
  const int max_chunks = std::max(1, S_max / (cpw * 7));
  int nc = (int)std::min<int64_t>(want, max_chunks);
  // up to 8 blocks per draw meet as a thread-block cluster (small.cuh, SM_CLUSTER): a little over is not worth
  // the trip through global memory
  if (nc > 8 && nc <= 10 && force_chunks <= 0 && h->opt[CHI_OPT_SMALL_CLUSTER] != 1) nc = 8;
  auto chunk_len = [&](int S) { int len = (S + nc - 1) / nc; return std::max(cpw, (len + cpw - 1) / cpw * cpw); };
  a.chunk_e = a.S_e ? chunk_len(a.S_e) : 0;
  a.chunk_a = a.S_a ? chunk_len(a.S_a) : 0;
  nc = std::max(a.S_e ? (a.S_e + a.chunk_e - 1) / a.chunk_e : 1, a.S_a ? (a.S_a + a.chunk_a - 1) / a.chunk_a : 1);
  a.n_chunks = nc;
  // (the same warps are set aside whether or not this call wants the gradient: a logp-only call
  // then sums its channels in the same order and returns the same bits)
  a.jac_warps = ((N * a.K + nc - 1) / nc + 31) / 32;
  const int per_cloud = a.fused ? (pv ? 10 : 6) : (a.has_e ? (pv ? 8 : 5) : 0) + (a.has_a ? (pv ? 6 : 3) : 0);
  a.stride = 1 + a.nb_e + a.nb_a + NP * per_cloud;
  CK(L.mom_e.ensure((size_t)B * nc * a.stride * sizeof(double)));
  CK(L.jacbuf.ensure((size_t)B * N * a.K * CHI_JAC_W * sizeof(double)));
  CK(L.prior.ensure((size_t)B * N * sizeof(double)));
  CK(L.cnt.ensure((size_t)B * sizeof(unsigned)));
  if (L.cnt_zeroed != L.cnt.cap) {  // fresh allocation: the kernel leaves the counters at zero afterwards
    CK(cudaMemsetAsync(L.cnt.p, 0, L.cnt.cap, L.st));
    L.cnt_zeroed = L.cnt.cap;
  }
  a.chan = h->chan.d(); a.chan_e = h->e.chan.d(); a.chan_a = h->a.chan.d();
  a.basis_e = h->e.basis.d(); a.basis_a = h->a.basis.d();
  a.theta = theta; a.coeff_e = ce; a.coeff_a = ca; a.sl = sl; a.lconst = h->lconst.d();
  a.mom = L.mom_e.d(); a.jac = L.jacbuf.d(); a.prior_buf = L.prior.d(); a.counter = (unsigned*)L.cnt.p;
  a.logp = logp; a.grad = grad; a.gce = gce; a.gca = gca;
  memcpy(a.bsig_e, h->bsig_e, sizeof(a.bsig_e));
  memcpy(a.bsig_a, h->bsig_a, sizeof(a.bsig_a));
  a.stamps = (long long*)(uintptr_t)h->opt[CHI_OPT_SMALL_STAMPS];
  for (int i = 0; i < a.nb_e && model; ++i) a.bprior_const += -log(h->bsig_e[i]) - CHI_HALF_LOG_2PI;
  for (int i = 0; i < a.nb_a && model; ++i) a.bprior_const += -log(h->bsig_a[i]) - CHI_HALF_LOG_2PI;
  const int n_gb = (a.nb_e > 1 ? a.nb_e - 1 : 0) + (a.nb_a > 1 ? a.nb_a - 1 : 0);
  *NL_out = NL;
  *dyn_smem_out = (size_t)n_gb * threads * sizeof(double);
  return CHI_OK;
}

// The sampler's per-chain kernel (nuts_chain.cuh) against the graph of kernels per tick: one block evaluates a
// whole draw per tick, which wins while a tick is bound by its dependent latency and loses once a draw's channel
// sums want more than one SM.  Measured on B200 (profiles/r2_nuts_chain.txt), us per tick:
//   2 clouds x 334+166 channels, 8 chains:                      graph 26   chain kernel 14.5
//   8 clouds x 2048+2048 channels, pseudo-Voigt, 64 chains:     graph 43   chain kernel 86
// i.e. chain ~ 10 + 3.1 w, graph ~ 26 + 0.7 w with w = clouds x channels (x 1.5 pseudo-Voigt) in thousands: the
// crossover is near w = 6.5.
// With 512 threads per chain (one block per SM, <= 8 clouds) a tick has twice the channel warps: the chain line
// becomes ~ 9 + 1.7 w.
// CHI_OPT_NUTS_TICK: 0 = by size, 1 = graph with the round-1 tick, 2 = graph with the staged tick, 3 = the chain
// kernel with 256 threads whenever it can run (bit-identical to 2 + the one-launch evaluation with one block per
// draw), 4 = the chain kernel with 512 threads where it can (else 256), 5 = the cluster form, 6 = the speculative
// form (a ninth warp runs the state machine next to the evaluation; same draws as 3).
#define CHI_CHAIN_TICKS 2048
// blocks per chain of the chain kernel's CLUSTER form (k_nuts_chain_cl), 0 = not that form: every block of every
// chain has to be resident (two blocks per SM), and a cluster has at most 8 blocks
int nuts_chain_cluster(const chi_handle* h, int64_t B) {
  const int64_t mode = h->opt[CHI_OPT_NUTS_TICK];
  const int N = h->cfg.n_clouds;
  if (N * CHI_NSLOT > CHI_SM_THREADS - 96 || getenv("CHI_NUTS_THREAD")) return 0;
  const int nc = (int)std::min<int64_t>(8, (2 * (int64_t)h->n_sm) / std::max<int64_t>(B, 1));
  if (nc < 2) return 0;
  if (mode == 5) return nc;
  if (mode != 0) return 0;
  // by size: past the single-block forms (nuts_chain_threads) and while a block's share of the channels stays
  // small enough for the per-tick latency to matter (measured: profiles/r2_nuts_chain.txt)
  const int64_t S = std::max(h->has_e ? h->e.S : 0, h->has_a ? h->a.S : 0);
  const int64_t work = (int64_t)N * S * (h->cfg.lorentzian ? 3 : 2) / 2;
  const bool single = (N <= 8 && B <= (int64_t)h->n_sm && work <= 14000) || work <= 6500;
  return (!single && work / nc <= 12000) ? nc : 0;
}
// threads per chain of the chain kernel, 0 = the graph of kernels runs instead
int nuts_chain_threads(const chi_handle* h, int64_t B) {
  const int64_t mode = h->opt[CHI_OPT_NUTS_TICK];
  if (mode == 1 || mode == 2 || mode == 5) return 0;
  if (getenv("CHI_NUTS_THREAD")) return 0;
  const int N = h->cfg.n_clouds;
  if (N * CHI_NSLOT > CHI_SM_THREADS - 96) return 0;
  const bool wide_ok = N <= 8 && B <= (int64_t)h->n_sm;
  if (mode == 3) return CHI_SM_THREADS;
  if (mode == 4) return wide_ok ? 512 : CHI_SM_THREADS;
  if (mode == 6) return CHI_SM_THREADS + 32;  // the speculative form (nuts_chain.cuh: k_nuts_chain_spec)
  const int64_t S = std::max(h->has_e ? h->e.S : 0, h->has_a ? h->a.S : 0);
  const int64_t work = (int64_t)N * S * (h->cfg.lorentzian ? 3 : 2) / 2;
  if (B > 8 * (int64_t)h->n_sm) return 0;
  if (work <= 6500 && B <= (int64_t)h->n_sm) return CHI_SM_THREADS + 32;  // one block per SM: the speculative form
  if (wide_ok && work <= 14000) return 512;
  return work <= 6500 ? CHI_SM_THREADS : 0;
}

int run_small(chi_handle* h, Lane& L, int64_t B, const double* theta, const double* ce, const double* ca,
              const int32_t* sl, double* logp, double* grad, double* gce, double* gca, bool model) {
  cudaEvent_t* evs = h->ring[h->stage_calls % CHI_STAGE_RING];
  SmallArgs a;
  int NL = 1;
  size_t dyn = 0;
  int rc = small_args(h, L, B, theta, ce, ca, sl, logp, grad, gce, gca, model, 0, &a, &NL, &dyn);
  if (rc) return rc;
  if (!h->capturing) {
    CK(cudaEventRecord(evs[0], L.st));
    CK(cudaEventRecord(evs[1], L.st));
  }
  // several blocks per draw: as one thread-block cluster that exchanges through distributed shared memory
  // (small.cuh, SM_CLUSTER) unless CHI_OPT_SMALL_CLUSTER says otherwise
  const bool cluster = a.n_chunks >= 2 && a.n_chunks <= 8 && h->opt[CHI_OPT_SMALL_CLUSTER] != 1;
  int lrc = -1;
  if (cluster) {
    lrc = launch_eval_small_cl(NL, h->cfg.lorentzian != 0, h->map, a, dyn, L.st);
    if (lrc) cudaGetLastError();  // (a cluster that cannot be placed: the global-memory form computes the same)
  }
  if (lrc) lrc = launch_eval_small(NL, h->cfg.lorentzian != 0, h->map, a, dyn, L.st);
  if (lrc) return fail(h, CHI_EINVAL, "unsupported small-batch configuration");
  h->launches++;
  CK(cudaGetLastError());
  if (!h->capturing) {
    CK(cudaEventRecord(evs[2], L.st));
    CK(cudaEventRecord(evs[3], L.st));
    CK(cudaEventRecord(evs[4], L.st));
    h->stage_calls++;
  }
  return CHI_OK;
}

// The two-launch forms pay off while marked draws are rare: a marked draw is one warp's work for ~34 us, and
// EA_FIXUP's static partition balances a batch full of them badly (measured: every draw marked 2.45 against 1.47 ms, 1 % marked 1.46 against
// 1.47, none 1.41 against 1.47; tools/marked_batch_bench.py).  A recent COMPLETED call's count of marked draws above
// 0.5 % of its batch sends the handle back to the one-launch form; every 64th call tries the two launches again.
bool two_pass_ok(chi_handle* h) {
  if (h->opt[CHI_OPT_EA_ONE_PASS]) return false;
  bool dense = false;
  if (h->marked_host)
    for (int i = 0; i < CHI_LANES; ++i)
      if (h->marked_B[i] > 0 && (int64_t)h->marked_host[i] * 200 > h->marked_B[i]) dense = true;
  if (!dense) return true;
  return (++h->dense_calls % 64) == 0;
}
// the counter of lane L for an EA_FIXUP launch on `st`, zeroed there; null if it cannot be had
unsigned long long* marked_counter(chi_handle* h, Lane& L, cudaStream_t st) {
  if (!h->marked_host || !h->marked_dev.p) return nullptr;
  unsigned long long* c = (unsigned long long*)h->marked_dev.p + (&L - h->lanes);
  if (cudaMemsetAsync(c, 0, sizeof(unsigned long long), st) != cudaSuccess) return nullptr;
  return c;
}
// behind the EA_FIXUP launch: its count to pinned memory (not inside a capture: a replay would not update marked_B)
void marked_readback(chi_handle* h, Lane& L, unsigned long long* c, int64_t B, cudaStream_t st) {
  if (!c || h->capturing) return;
  const int i = (int)(&L - h->lanes);
  if (cudaMemcpyAsync(h->marked_host + i, c, sizeof(unsigned long long), cudaMemcpyDeviceToHost, st) == cudaSuccess) h->marked_B[i] = B;
}

// moments (+ optional logp) + epilogue; gbar_* != null selects the VJP mode
// `model` = true: theta holds unconstrained values; priors and Jacobians are added (chi_model_logp_grad)
enum { PHASE_ALL = 0, PHASE_MAP = 1, PHASE_REST = 2 };

int run_backward(chi_handle* h, Lane& L, int64_t B, const double* theta, const double* ce, const double* ca,
                 const int32_t* sl, const double* gbar_e, const double* gbar_a, bool vjp, double* logp,
                 double* grad, double* gce, double* gca, bool model = false, int phase = PHASE_ALL) {
  // phase: the pipelined host path issues the cloud map of sub-batch k+1 (PHASE_MAP, `ahead` = 1
  // stage-ring slot further on) before the channel kernels of sub-batch k (PHASE_REST), so that the
  // map is not queued behind the previous sub-batch's channel kernel and the channel kernels of
  // consecutive sub-batches run back to back
  if (use_eval_small(h, B, vjp)) {
    if (phase == PHASE_MAP) return CHI_OK;  // the map is part of the one launch
    return run_small(h, L, B, theta, ce, ca, sl, logp, grad, gce, gca, model);
  }
  const int N = h->cfg.n_clouds;
  const bool pv = h->cfg.lorentzian != 0;
  const int ahead = (phase == PHASE_MAP) ? h->map_ahead : 0;
  cudaEvent_t* evs = h->ring[(h->stage_calls + ahead) % CHI_STAGE_RING];
  const double* u = nullptr;
  if (model) {
    CK(L.xcon.ensure((size_t)B * CHI_NSLOT * N * sizeof(double)));
    u = theta;
  }
  if (phase != PHASE_REST) {
    if (!h->capturing) CK(cudaEventRecord(evs[0], L.st));
    int rc = launch_cloud_map(h, L, B, theta, model ? L.xcon.d() : nullptr);
    if (rc) return rc;
    if (!h->capturing) CK(cudaEventRecord(evs[1], L.st));
    CK(cudaEventRecord(L.ev_fork, L.st));  // the cloud constants are ready from here on
    if (phase == PHASE_MAP) return CHI_OK;
  }
  if (model) theta = L.xcon.d();  // everything downstream works on the constrained values

  EpiArgs ea;
  memset(&ea, 0, sizeof(ea));
  ea.B = B; ea.N = N; ea.pv = pv; ea.rm = h->rm;
  ea.theta = theta; ea.sl = sl; ea.logp = logp; ea.grad = grad; ea.gcoeff_e = gce; ea.gcoeff_a = gca;
  ea.lconst = vjp ? nullptr : h->lconst.d();
  if (model) {
    ea.u = u; ea.coeff_e = ce; ea.coeff_a = ca;
    memcpy(ea.bsig_e, h->bsig_e, sizeof(ea.bsig_e));
    memcpy(ea.bsig_a, h->bsig_a, sizeof(ea.bsig_a));
  }

  // in VJP mode a dataset without cotangent contributes nothing
  const bool do_e = h->has_e && (!vjp || gbar_e);
  const bool do_a = h->has_a && (!vjp || gbar_a);
  bool fused_done = false;

  // Small batches are bound by dependent latency, not throughput: the Jacobian of the per-cloud maps
  // (forward-mode duals, needs theta only) runs on the lane's `jac` stream NEXT TO the channel
  // kernels, and a short contraction kernel replaces the epilogue behind them.  (The choice follows
  // the size of the caller's batch, not of a pipeline piece: every draw of a call goes through the
  // same arithmetic whatever piece it lands in.)
  const int K = h->cfg.model == CHI_MODEL_PHYSICS ? 7 : CHI_NSLOT;
  const bool small = batch_is_small(h, B);
  static const bool old_split = getenv("CHI_OLD_SPLIT") != nullptr;  // A/B knob: split epilogue on the main stream
  const bool side_jac = small && !old_split && !h->epilogue_on_hi;
  if (side_jac && (grad || model)) {
    const bool fused = h->shared_axis && (do_e || do_a);
    JacArgs ja;
    memset(&ja, 0, sizeof(ja));
    ja.B = B; ja.N = N; ja.rm = h->rm; ja.has_e = (fused || do_e) ? 1 : 0; ja.has_a = (fused || do_a) ? 1 : 0;
    ja.theta = theta; ja.u = u;
    CK(L.jacbuf.ensure((size_t)B * N * K * CHI_JAC_W * sizeof(double)));
    CK(L.prior.ensure((size_t)B * N * sizeof(double)));
    ja.jac = L.jacbuf.d(); ja.prior = L.prior.d();
    CK(cudaStreamWaitEvent(L.jac, L.ev_fork, 0));  // theta (and, in model mode, the constrained values) are ready
    const unsigned gj = (unsigned)((B * N * K + 127) / 128);
    if (model) { MODEL_SWITCH(h->cfg.model, (k_jacobian<M, true><<<gj, 128, 0, L.jac>>>(h->map, ja))); }
    else { MODEL_SWITCH(h->cfg.model, (k_jacobian<M, false><<<gj, 128, 0, L.jac>>>(h->map, ja))); }
    h->launches++;
    CK(cudaGetLastError());
    CK(cudaEventRecord(L.ev_jac, L.jac));
  }
  if (h->shared_axis && (do_e || do_a)) {
    // both datasets on one axis: one pass evaluates exp(-k d^2) once for the two of them
    FusedArgs a;
    memset(&a, 0, sizeof(a));
    a.B = B; a.S = h->e.S; a.nb_e = h->e.nb; a.nb_a = h->a.nb; a.bg = h->cfg.bg_temp;
    pick_chunks(h, std::max<int64_t>(B, h->call_draws), a.S, &a.n_chunks, &a.chunk_len);
    a.n_real = N;
    a.staged = (h->n_sl > 1 && !getenv("CHI_NO_STAGED")) ? 1 : 0;
    a.mom_stride = 1 + a.nb_e + a.nb_a + padded_clouds(N) * (pv ? 10 : 6);
    CK(L.mom_e.ensure((size_t)B * a.n_chunks * a.mom_stride * sizeof(double)));
    a.chan = h->chan.d();
    a.basis_e = h->e.basis.d(); a.coeff_e = ce;
    a.basis_a = h->a.basis.d(); a.coeff_a = ca;
    a.vjp = vjp ? 1 : 0; a.gbar_e = vjp ? gbar_e : nullptr; a.gbar_a = vjp ? gbar_a : nullptr;
    a.sl = sl; a.cc = L.cc.d(); a.mom = L.mom_e.d();
    if (!small && moments_ea_two_pass(N, pv, a) && two_pass_ok(h)) {
      // the marked draws (a negative or huge optical depth; normally none) on a wave of resident blocks next to
      // the big grid, which then runs without the range selects of exp(-tau) (moments_ea.cuh, MODE)
      CK(cudaStreamWaitEvent(L.aux, L.ev_fork, 0));
      a.n_marked = marked_counter(h, L, L.aux);
      if (launch_moments_ea_fixup(N, a, h->n_sm, L.aux)) return fail(h, CHI_EINVAL, "unsupported n_clouds");
      marked_readback(h, L, a.n_marked, B, L.aux);
      CK(cudaEventRecord(L.ev_join, L.aux));
      if (launch_moments_ea_tame(N, a, L.st)) return fail(h, CHI_EINVAL, "unsupported n_clouds");
      CK(cudaStreamWaitEvent(L.st, L.ev_join, 0));
      h->launches += 2;
    } else {
      if (launch_moments_ea(N, pv, a, L.st)) return fail(h, CHI_EINVAL, "unsupported n_clouds");
      h->launches++;
    }
    CK(cudaGetLastError());
    ea.fused = 1; ea.has_e = 1; ea.has_a = 1; ea.nb_e = a.nb_e; ea.nb_a = a.nb_a;
    ea.chunks_e = a.n_chunks; ea.stride_e = a.mom_stride; ea.mom_e = L.mom_e.d();
    if (!h->capturing) {
      CK(cudaEventRecord(evs[2], L.st));
      CK(cudaEventRecord(evs[3], L.st));
    }
    fused_done = true;
  }
  if (do_e && !fused_done) {
    ChanArgs a;
    memset(&a, 0, sizeof(a));
    a.B = B; a.S = h->e.S; a.n_base = h->e.nb; a.n_real = N; a.bg = h->cfg.bg_temp;
    a.staged = (h->n_sl > 1 && !getenv("CHI_NO_STAGED")) ? 1 : 0;
    pick_chunks(h, std::max<int64_t>(B, h->call_draws), a.S, &a.n_chunks, &a.chunk_len);
    a.mom_stride = em_record(h);
    CK(L.mom_e.ensure((size_t)B * a.n_chunks * a.mom_stride * sizeof(double)));
    a.chan = h->e.chan.d(); a.basis = h->e.basis.d();
    a.coeff = ce; a.sl = sl; a.gbar = vjp ? gbar_e : nullptr; a.cc = L.cc.d(); a.mom = L.mom_e.d();
    if (!small && moments_em_two_pass(N, a) && two_pass_ok(h)) {
      // as for the fused kernel: the marked draws on a wave of resident blocks next to the big grid
      CK(cudaStreamWaitEvent(L.aux, L.ev_fork, 0));
      a.n_marked = marked_counter(h, L, L.aux);
      if (launch_moments_em_fixup(N, pv, a, h->n_sm, L.aux)) return fail(h, CHI_EINVAL, "unsupported n_clouds");
      marked_readback(h, L, a.n_marked, B, L.aux);
      CK(cudaEventRecord(L.ev_join, L.aux));
      if (launch_moments_em_tame(N, pv, a, L.st)) return fail(h, CHI_EINVAL, "unsupported n_clouds");
      CK(cudaStreamWaitEvent(L.st, L.ev_join, 0));
      h->launches += 2;
    } else {
      if (launch_moments_em(N, pv, a, L.st)) return fail(h, CHI_EINVAL, "unsupported n_clouds");
      h->launches++;
    }
    CK(cudaGetLastError());
    ea.has_e = 1; ea.nb_e = a.n_base; ea.chunks_e = a.n_chunks; ea.stride_e = a.mom_stride;
    ea.mom_e = L.mom_e.d();
  }
  if (!h->capturing) CK(cudaEventRecord(evs[2], L.st));
  // under-filled launches (few chains): absorption runs next to emission on the aux stream
  const bool fork = !fused_done && do_e && do_a && B * 8 < (int64_t)h->n_sm * 16;
  cudaStream_t st_a = fork ? L.aux : L.st;
  if (fork) CK(cudaStreamWaitEvent(L.aux, L.ev_fork, 0));
  if (do_a && !fused_done) {
    ChanArgs a;
    memset(&a, 0, sizeof(a));
    a.B = B; a.S = h->a.S; a.n_base = h->a.nb; a.n_real = N; a.bg = 0.0;
    a.staged = (h->n_sl > 1 && !getenv("CHI_NO_STAGED")) ? 1 : 0;
    pick_chunks(h, std::max<int64_t>(B, h->call_draws), a.S, &a.n_chunks, &a.chunk_len);
    a.mom_stride = ab_record(h);
    CK(L.mom_a.ensure((size_t)B * a.n_chunks * a.mom_stride * sizeof(double)));
    a.chan = h->a.chan.d(); a.basis = h->a.basis.d();
    a.coeff = ca; a.sl = sl; a.gbar = vjp ? gbar_a : nullptr; a.cc = L.cc.d(); a.mom = L.mom_a.d();
    if (launch_moments_abs(N, pv, a, st_a)) return fail(h, CHI_EINVAL, "unsupported n_clouds");
    h->launches++;
    CK(cudaGetLastError());
    ea.has_a = 1; ea.nb_a = a.n_base; ea.chunks_a = a.n_chunks; ea.stride_a = a.mom_stride;
    ea.mom_a = L.mom_a.d();
  }
  if (fork) {
    CK(cudaEventRecord(L.ev_join, L.aux));
    CK(cudaStreamWaitEvent(L.st, L.ev_join, 0));
  }
  if (!h->capturing) CK(cudaEventRecord(evs[3], L.st));
  cudaStream_t es = L.st;  // stream of the epilogue
  if (h->epilogue_on_hi) {
    CK(cudaEventRecord(L.ev_mom, L.st));
    CK(cudaStreamWaitEvent(L.hi, L.ev_mom, 0));
    es = L.hi;
  }
  // baseline-coefficient gradients of a dataset that did not run are zero
  if (gce && !do_e && h->e.nb > 0) CK(cudaMemsetAsync(gce, 0, (size_t)B * h->e.nb * sizeof(double), es));
  if (gca && !do_a && h->a.nb > 0) CK(cudaMemsetAsync(gca, 0, (size_t)B * h->a.nb * sizeof(double), es));

  if (side_jac) {
    ConArgs ca_;
    memset(&ca_, 0, sizeof(ca_));
    ca_.e = ea; ca_.K = K; ca_.cc = L.cc.d();
    ca_.jac = (grad || model) ? L.jacbuf.d() : nullptr; ca_.prior = L.prior.d();
    if (grad || model) CK(cudaStreamWaitEvent(es, L.ev_jac, 0));
    if (model) k_contract<true><<<(unsigned)B, 128, 0, es>>>(ca_);
    else k_contract<false><<<(unsigned)B, 128, 0, es>>>(ca_);
  } else
  // split epilogue (kept as the A/B reference of the path above; N * K > 128 also lands here)
  if (small) {
    unsigned grid = (unsigned)((B * N * K + 127) / 128);
    if (model) {  // whole draws per block
      const int per = N * K, bd = (128 / per) * per;
      const unsigned gp = (unsigned)((B + bd / per - 1) / (bd / per));
      MODEL_SWITCH(h->cfg.model, (k_epilogue<M, true, true><<<gp, bd, 0, es>>>(h->map, ea)));
    }
    else { MODEL_SWITCH(h->cfg.model, (k_epilogue<M, true, false><<<grid, 128, 0, es>>>(h->map, ea))); }
  } else {
    unsigned grid = (unsigned)((B * N + 127) / 128);
    if (model) {  // whole draws per block
      const int bd = (128 / N) * N;
      const unsigned gp = (unsigned)((B + bd / N - 1) / (bd / N));
      MODEL_SWITCH(h->cfg.model, (k_epilogue<M, false, true><<<gp, bd, 0, es>>>(h->map, ea)));
    }
    else { MODEL_SWITCH(h->cfg.model, (k_epilogue<M, false, false><<<grid, 128, 0, es>>>(h->map, ea))); }
  }

  h->launches++;
  CK(cudaGetLastError());
  if (!h->capturing) CK(cudaEventRecord(evs[4], es));
  if (!h->capturing) h->stage_calls++;
  return CHI_OK;
}

int run_spectra(chi_handle* h, int64_t B, const double* theta, const double* ce, const double* ca,
                double* mu_e, double* mu_a) {
  const int N = h->cfg.n_clouds;
  h->call_first_stage = h->stage_calls;  // no logp/VJP stage events belong to this call
  CK(cudaEventRecord(h->ev0, h->stream));
  int rc = launch_cloud_map(h, h->lanes[0], B, theta);
  if (rc) return rc;
  SpecArgs a;
  memset(&a, 0, sizeof(a));
  a.B = B; a.N = N; a.pv = h->cfg.lorentzian != 0; a.cc = h->lanes[0].cc.d();
  if (mu_e && mu_a && h->has_e && h->has_a && h->shared_axis) {
    // one pass for both spectra of a shared axis
    a.S = h->e.S; a.bg = h->cfg.bg_temp; a.vax = h->e.vax.d(); a.emission = 1;
    a.n_base = h->e.nb; a.basis = h->e.basis.d(); a.offset = h->e.has_offset ? h->e.offset.d() : nullptr;
    a.coeff = ce; a.mu = mu_e;
    a.n_base_a = h->a.nb; a.basis_a = h->a.basis.d(); a.offset_a = h->a.has_offset ? h->a.offset.d() : nullptr;
    a.coeff_a = ca; a.mu_a = mu_a;
    if (launch_spectra_ea(N, a.pv != 0, a, h->stream)) return fail(h, CHI_EINVAL, "unsupported n_clouds");
    h->launches++;
    CK(cudaGetLastError());
    CK(cudaEventRecord(h->ev1, h->stream));
    h->ev_pending = true;
    return CHI_OK;
  }
  if (mu_e && h->has_e) {
    a.S = h->e.S; a.n_base = h->e.nb; a.bg = h->cfg.bg_temp; a.vax = h->e.vax.d();
    a.basis = h->e.basis.d(); a.offset = h->e.has_offset ? h->e.offset.d() : nullptr;
    a.coeff = ce; a.mu = mu_e; a.emission = 1;
    if (launch_spectra(N, a.pv != 0, a, h->stream)) return fail(h, CHI_EINVAL, "unsupported n_clouds");
    h->launches++;
    CK(cudaGetLastError());
  }
  if (mu_a && h->has_a) {
    a.S = h->a.S; a.n_base = h->a.nb; a.bg = 0.0; a.vax = h->a.vax.d();
    a.basis = h->a.basis.d(); a.offset = h->a.has_offset ? h->a.offset.d() : nullptr;
    a.coeff = ca; a.mu = mu_a; a.emission = 0;
    if (launch_spectra(N, a.pv != 0, a, h->stream)) return fail(h, CHI_EINVAL, "unsupported n_clouds");
    h->launches++;
    CK(cudaGetLastError());
  }
  CK(cudaEventRecord(h->ev1, h->stream));
  h->ev_pending = true;
  return CHI_OK;
}

// single-precision outputs (launch_spectra_f32.cu); the cloud map is the fp64 one
int run_spectra32(chi_handle* h, int64_t B, const double* theta, const double* ce, const double* ca,
                  float* mu_e, float* mu_a) {
  const int N = h->cfg.n_clouds;
  h->call_first_stage = h->stage_calls;
  CK(cudaEventRecord(h->ev0, h->stream));
  int rc = launch_cloud_map(h, h->lanes[0], B, theta);
  if (rc) return rc;
  Spec32Args a;
  memset(&a, 0, sizeof(a));
  a.B = B; a.N = N; a.cc = h->lanes[0].cc.d();
  const bool pv = h->cfg.lorentzian != 0;
  const bool do_e = mu_e && h->has_e, do_a = mu_a && h->has_a;
  if (do_e) {
    a.S = h->e.S; a.bg = (float)h->cfg.bg_temp; a.vax32 = (const float*)h->e.vax32.p;
    a.n_base = h->e.nb; a.basis = h->e.basis.d(); a.offset = h->e.has_offset ? h->e.offset.d() : nullptr;
    a.coeff = ce; a.mu = mu_e; a.mode = 0;
    if (do_a && h->shared_axis) {  // one pass for both spectra of a shared axis
      a.mode = 2;
      a.n_base_a = h->a.nb; a.basis_a = h->a.basis.d(); a.offset_a = h->a.has_offset ? h->a.offset.d() : nullptr;
      a.coeff_a = ca; a.mu_a = mu_a;
    }
    if (launch_spectra32(N, pv, a, h->stream)) return fail(h, CHI_EINVAL, "unsupported n_clouds");
    h->launches++;
    CK(cudaGetLastError());
  }
  if (do_a && !(do_e && h->shared_axis)) {
    a.S = h->a.S; a.bg = 0.f; a.vax32 = (const float*)h->a.vax32.p;
    a.n_base = h->a.nb; a.basis = h->a.basis.d(); a.offset = h->a.has_offset ? h->a.offset.d() : nullptr;
    a.coeff = ca; a.mu = mu_a; a.mode = 1;
    a.n_base_a = 0; a.basis_a = nullptr; a.offset_a = nullptr; a.coeff_a = nullptr; a.mu_a = nullptr;
    if (launch_spectra32(N, pv, a, h->stream)) return fail(h, CHI_EINVAL, "unsupported n_clouds");
    h->launches++;
    CK(cudaGetLastError());
  }
  CK(cudaEventRecord(h->ev1, h->stream));
  h->ev_pending = true;
  return CHI_OK;
}

int check_model_call(chi_handle* h, int64_t B, const double* theta) {
  if (!h) return CHI_EINVAL;
  if (B < 0 || (B > 0 && !theta)) return fail(h, CHI_EINVAL, "theta is NULL or B < 0");
  if (!h->spectra_set) return fail(h, CHI_ESTATE, "chi_set_spectra has not been called");
  return CHI_OK;
}

int check_prior_call(chi_handle* h) {
  if (h->cfg.model == CHI_MODEL_PHYSICS)
    return fail(h, CHI_EINVAL, "chi_model_logp_grad: not available for CHI_MODEL_PHYSICS");
  if (!h->priors_set) return fail(h, CHI_ESTATE, "chi_set_priors has not been called");
  return CHI_OK;
}

int upload_dataset(chi_handle* h, DevDataset& d, const chi_dataset* src, int n_sl, double* lconst_host,
                   const char* name) {
  const int S = src->n_channels;
  if (S < 2) return fail(h, CHI_EINVAL, std::string(name) + ": need at least 2 channels (physics.py:294)");
  if (!src->spectral || !src->brightness || !src->noise)
    return fail(h, CHI_EINVAL, std::string(name) + ": spectral/brightness/noise must be non-NULL");
  if (src->n_baseline < 0 || src->n_baseline > CHI_MAX_BASELINE)
    return fail(h, CHI_EINVAL, std::string(name) + ": n_baseline out of range");
  if (src->n_baseline > 0 && !src->basis) return fail(h, CHI_EINVAL, std::string(name) + ": basis is NULL");
  d.S = S;
  d.nb = src->n_baseline;
  d.has_offset = src->offset != nullptr;
  const size_t groups = ((size_t)S + 31) / 32;
  std::vector<double> chan((size_t)n_sl * groups * 128, 0.0);  // [sightline][group][v, p_0, y, w][32]
  const double half_log_2pi = 0.5 * log(2.0 * M_PI);
  for (int i = 0; i < n_sl; ++i) {
    double lc = 0.0;
    for (int s = 0; s < S; ++s) {
      const double sig = src->noise[(size_t)i * S + s];
      if (!(sig > 0.0)) return fail(h, CHI_EINVAL, std::string(name) + ": noise must be > 0");
      double* c = &chan[((size_t)i * groups + s / 32) * 128 + s % 32];
      c[0] = src->spectral[s];
      c[32] = src->n_baseline > 0 ? src->basis[s] : 0.0;
      c[64] = src->brightness[(size_t)i * S + s] - (src->offset ? src->offset[s] : 0.0);
      c[96] = 1.0 / (sig * sig);
      lc += -log(sig) - half_log_2pi;  // pm.Normal logp normalisation
    }
    lconst_host[i] += lc;
  }
  CK(d.vax.ensure(S * sizeof(double)));
  CK(d.chan.ensure(chan.size() * sizeof(double)));
  CK(d.basis.ensure(std::max<size_t>(1, (size_t)d.nb * S) * sizeof(double)));
  CK(d.offset.ensure(S * sizeof(double)));
  CK(cudaMemcpy(d.vax.p, src->spectral, S * sizeof(double), cudaMemcpyHostToDevice));
  {
    std::vector<float> v32(((size_t)S + 1) / 2 * 4, 0.f);
    for (int s = 0; s < S; ++s) {
      const float hi = (float)src->spectral[s];
      v32[2 * s] = hi;
      v32[2 * s + 1] = (float)(src->spectral[s] - (double)hi);
    }
    CK(d.vax32.ensure(v32.size() * sizeof(float)));
    CK(cudaMemcpy(d.vax32.p, v32.data(), v32.size() * sizeof(float), cudaMemcpyHostToDevice));
  }
  CK(cudaMemcpy(d.chan.p, chan.data(), chan.size() * sizeof(double), cudaMemcpyHostToDevice));
  if (d.nb > 0) CK(cudaMemcpy(d.basis.p, src->basis, (size_t)d.nb * S * sizeof(double), cudaMemcpyHostToDevice));
  if (src->offset) CK(cudaMemcpy(d.offset.p, src->offset, S * sizeof(double), cudaMemcpyHostToDevice));
  return CHI_OK;
}

double channel_fwhm2(const double* v) {
  // physics.py:294-295
  const double channel_size = fabs(v[1] - v[0]);
  const double channel_fwhm = 4.0 * log(2.0) * channel_size / M_PI;
  return channel_fwhm * channel_fwhm;
}

// sub-batch size of the host entry points
int64_t host_sub_batch(const chi_handle* h, size_t bytes_per_draw) {
  const size_t budget = (size_t)1 << 30;
  int64_t n = (int64_t)(budget / std::max<size_t>(bytes_per_draw, 1));
  return std::max<int64_t>(1, std::min<int64_t>(n, (int64_t)1 << 18));
}

int finish_timing(chi_handle* h) {
  if (h->ev_pending) {
    CK(cudaEventSynchronize(h->ev1));
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
    h->kernel_ms += ms;
    h->ev_pending = false;
  }
  return CHI_OK;
}

}  // namespace

namespace {
inline int run_spectra_any(chi_handle* h, int64_t B, const double* th, const double* ce, const double* ca, double* e, double* a) {
  return run_spectra(h, B, th, ce, ca, e, a);
}
inline int run_spectra_any(chi_handle* h, int64_t B, const double* th, const double* ce, const double* ca, float* e, float* a) {
  return run_spectra32(h, B, th, ce, ca, e, a);
}

// predictive draws: y = mu + sigma z on top of spectra already in mu_e / mu_a (device), same stream
struct NoiseSpec {
  unsigned long long seed;
  int64_t draw_offset;
  const int32_t* sl;  // device, per draw, or null
};
template <typename T>
int add_noise(chi_handle* h, int64_t B, const NoiseSpec& ns, T* mu_e, T* mu_a) {
  if (mu_e && h->has_e) {
    const int64_t n = B * ((h->e.S + 1) / 2);
    k_add_noise<T><<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>(B, h->e.S, mu_e, h->e.chan.d(), ns.sl, ns.seed,
                                                                        ns.draw_offset, 0);
    h->launches++;
  }
  if (mu_a && h->has_a) {
    const int64_t n = B * ((h->a.S + 1) / 2);
    k_add_noise<T><<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>(B, h->a.S, mu_a, h->a.chan.d(), ns.sl, ns.seed,
                                                                        ns.draw_offset, 1);
    h->launches++;
  }
  CK(cudaGetLastError());
  CK(cudaEventRecord(h->ev1, h->stream));  // the timed region of the call ends after the noise
  return CHI_OK;
}

// host entry point of the predicted spectra, T = double (chi_spectra) or float (chi_spectra_f32);
// with `noise` the predictive draws of the observed nodes (chi_predictive[_f32]), `sl` = host sightlines
template <typename T>
int host_spectra(chi_handle* h, int64_t B, const double* theta, const double* ce, const double* ca, T* mu_e, T* mu_a,
                 const NoiseSpec* noise = nullptr, const int32_t* sl = nullptr) {
  int rc = check_model_call(h, B, theta);
  if (rc) return rc;
  if (noise && sl)
    for (int64_t i = 0; i < B; ++i)
      if (sl[i] < 0 || sl[i] >= h->n_sl) return fail(h, CHI_EINVAL, "chi_predictive: sightline index out of range");
  CK(cudaSetDevice(h->device));
  const int N = h->cfg.n_clouds;
  const size_t th_b = (size_t)CHI_NSLOT * N * sizeof(double);
  const int nbe = h->has_e ? h->e.nb : 0, nba = h->has_a ? h->a.nb : 0;
  const int Se = h->has_e ? h->e.S : 0, Sa = h->has_a ? h->a.S : 0;
  const int64_t sub = host_sub_batch(h, (size_t)(Se + Sa) * sizeof(T) + 4096);
  h->kernel_ms = 0.f;
  for (int64_t b0 = 0; b0 < B; b0 += sub) {
    const int64_t nb = std::min(sub, B - b0);
    CK(h->lanes[0].w_theta.ensure(nb * th_b));
    CK(cudaMemcpyAsync(h->lanes[0].w_theta.p, theta + b0 * CHI_NSLOT * N, nb * th_b, cudaMemcpyHostToDevice, h->stream));
    double *d_ce = nullptr, *d_ca = nullptr;
    T *d_me = nullptr, *d_ma = nullptr;
    if (ce && nbe) {
      CK(h->lanes[0].w_ce.ensure(nb * nbe * sizeof(double)));
      CK(cudaMemcpyAsync(h->lanes[0].w_ce.p, ce + b0 * nbe, nb * nbe * sizeof(double), cudaMemcpyHostToDevice, h->stream));
      d_ce = h->lanes[0].w_ce.d();
    }
    if (ca && nba) {
      CK(h->lanes[0].w_ca.ensure(nb * nba * sizeof(double)));
      CK(cudaMemcpyAsync(h->lanes[0].w_ca.p, ca + b0 * nba, nb * nba * sizeof(double), cudaMemcpyHostToDevice, h->stream));
      d_ca = h->lanes[0].w_ca.d();
    }
    if (mu_e && Se) { CK(h->w_mue.ensure((size_t)nb * Se * sizeof(T))); d_me = (T*)h->w_mue.p; }
    if (mu_a && Sa) { CK(h->w_mua.ensure((size_t)nb * Sa * sizeof(T))); d_ma = (T*)h->w_mua.p; }
    rc = run_spectra_any(h, nb, h->lanes[0].w_theta.d(), d_ce, d_ca, d_me, d_ma);
    if (rc) return rc;
    if (noise) {
      NoiseSpec ns = *noise;
      ns.draw_offset += b0;
      ns.sl = nullptr;
      if (sl) {
        CK(h->lanes[0].w_sl.ensure(nb * sizeof(int32_t)));
        CK(cudaMemcpyAsync(h->lanes[0].w_sl.p, sl + b0, nb * sizeof(int32_t), cudaMemcpyHostToDevice, h->stream));
        ns.sl = (const int32_t*)h->lanes[0].w_sl.p;
      }
      rc = add_noise<T>(h, nb, ns, d_me, d_ma);
      if (rc) return rc;
    }
    if (d_me) CK(cudaMemcpyAsync(mu_e + b0 * Se, d_me, (size_t)nb * Se * sizeof(T), cudaMemcpyDeviceToHost, h->stream));
    if (d_ma) CK(cudaMemcpyAsync(mu_a + b0 * Sa, d_ma, (size_t)nb * Sa * sizeof(T), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    rc = finish_timing(h);
    if (rc) return rc;
  }
  return CHI_OK;
}
}  // namespace

extern "C" {

int chi_version(void) { return CHI_VERSION; }

int chi_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
  int ok = 0;
  for (int i = 0; i < n; ++i) {
    cudaDeviceProp p;
    if (cudaGetDeviceProperties(&p, i) == cudaSuccess && p.major == 10) ok++;
  }
  return ok;
}

const char* chi_last_error(const chi_handle* h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int chi_create(chi_handle** out, int device, const chi_config* cfg) {
  chi_handle* h = nullptr;
  if (!out || !cfg) return fail(h, CHI_EINVAL, "chi_create: NULL argument");
  *out = nullptr;
  if (cfg->model < CHI_MODEL_PHYSICS || cfg->model > CHI_MODEL_EMISSION_ABSORPTION_PHYSICAL)
    return fail(h, CHI_EINVAL, "chi_create: unknown model kind");
  if (cfg->n_clouds < 1 || cfg->n_clouds > CHI_MAX_CLOUDS)
    return fail(h, CHI_EINVAL, "chi_create: n_clouds must be in 1..CHI_MAX_CLOUDS");
  if (cfg->dtype != 0) return fail(h, CHI_EINVAL, "chi_create: only dtype 0 (fp64) is implemented");
  if (cfg->model >= CHI_MODEL_ABSORPTION_PHYSICAL && !(cfg->depth_nth_fwhm_power != 0.0))
    return fail(h, CHI_EINVAL, "chi_create: depth_nth_fwhm_power must be non-zero");
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0)
    return fail(h, CHI_ENODEV, std::string("chi_create: no CUDA device (") + cudaGetErrorString(e) +
                                   "); libchi has no CPU fallback");
  if (device < 0 || device >= n) return fail(h, CHI_EINVAL, "chi_create: device index out of range");
  cudaDeviceProp prop;
  e = cudaGetDeviceProperties(&prop, device);
  if (e != cudaSuccess) return fail(h, CHI_ECUDA, cudaGetErrorString(e));
  if (prop.major != 10)
    return fail(h, CHI_ENODEV, std::string("chi_create: device '") + prop.name +
                                   "' is not sm_100; libchi is built for sm_100a only and has no fallback");
  e = cudaSetDevice(device);
  if (e != cudaSuccess) return fail(h, CHI_ECUDA, cudaGetErrorString(e));
  h = new chi_handle();
  h->rm = identity_rows();
  h->cfg = *cfg;
  h->device = device;
  h->n_sm = prop.multiProcessorCount;
  // (allocated here, not at first use: a first use may be inside a stream capture)
  if (h->marked_dev.ensure(CHI_LANES * sizeof(unsigned long long)) != cudaSuccess ||
      cudaMallocHost(&h->marked_host, CHI_LANES * sizeof(unsigned long long)) != cudaSuccess) {
    h->marked_host = nullptr;
    cudaGetLastError();
  } else {
    memset(h->marked_host, 0, CHI_LANES * sizeof(unsigned long long));
  }
  const int m = cfg->model;
  h->has_e = (m == CHI_MODEL_PHYSICS) ? cfg->has_emission != 0 : kind_has_emission(m);
  h->has_a = (m == CHI_MODEL_PHYSICS) ? cfg->has_absorption != 0 : kind_has_absorption(m);
  MapCfg& mc = h->map;
  memset(&mc, 0, sizeof(mc));
  mc.model = m;
  mc.lorentzian = cfg->lorentzian != 0;
  mc.free_weight = cfg->free_weight != 0;
  mc.has_e = h->has_e;
  mc.has_a = h->has_a;
  mc.bg_temp = cfg->bg_temp;
  mc.inv_power = (cfg->depth_nth_fwhm_power != 0.0) ? 1.0 / cfg->depth_nth_fwhm_power : 0.0;  // physics.py:188
  mc.prior_fwhm2 = cfg->prior_fwhm2;
  mc.pv0 = cfg->prior_velocity[0]; mc.pv1 = cfg->prior_velocity[1];
  mc.pn0 = cfg->prior_log10_nHI[0]; mc.pn1 = cfg->prior_log10_nHI[1];
  mc.pa0 = cfg->prior_log10_n_alpha[0]; mc.pa1 = cfg->prior_log10_n_alpha[1];
  mc.prior_fwhm_L = cfg->prior_fwhm_L;
  mc.prior_amp = cfg->prior_amplitude;
  mc.const_G = sqrt(2.0 * M_PI) / (2.0 * sqrt(2.0 * log(2.0)));  // emission_model.py:127
  bool ok = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) == cudaSuccess &&
            cudaEventCreate(&h->ev0) == cudaSuccess && cudaEventCreate(&h->ev1) == cudaSuccess && make_ring(h);
  h->lanes[0].st = h->stream;
  for (int i = 1; ok && i < CHI_LANES; ++i)
    ok = cudaStreamCreateWithFlags(&h->lanes[i].st, cudaStreamNonBlocking) == cudaSuccess;
  for (int i = 0; ok && i < CHI_LANES; ++i)
    ok = cudaStreamCreateWithFlags(&h->lanes[i].aux, cudaStreamNonBlocking) == cudaSuccess &&
         cudaEventCreateWithFlags(&h->lanes[i].ev_fork, cudaEventDisableTiming) == cudaSuccess &&
         cudaEventCreateWithFlags(&h->lanes[i].ev_join, cudaEventDisableTiming) == cudaSuccess &&
         cudaStreamCreateWithFlags(&h->lanes[i].jac, cudaStreamNonBlocking) == cudaSuccess &&
         cudaEventCreateWithFlags(&h->lanes[i].ev_jac, cudaEventDisableTiming) == cudaSuccess;
  {
    int least = 0, greatest = 0;
    cudaDeviceGetStreamPriorityRange(&least, &greatest);
    for (int i = 0; ok && i < CHI_LANES; ++i)
      ok = cudaStreamCreateWithPriority(&h->lanes[i].hi, cudaStreamNonBlocking, greatest) == cudaSuccess &&
           cudaEventCreateWithFlags(&h->lanes[i].ev_mom, cudaEventDisableTiming) == cudaSuccess &&
           cudaEventCreateWithFlags(&h->lanes[i].ev_tail, cudaEventDisableTiming) == cudaSuccess;
  }
  if (!ok) {
    delete h;
    return fail(nullptr, CHI_ECUDA, "chi_create: stream/event creation failed");
  }
  *out = h;
  return CHI_OK;
}

int chi_comm_destroy(chi_handle* h);

void chi_destroy(chi_handle* h) {
  if (!h) return;
  cudaSetDevice(h->device);
  chi_comm_destroy(h);
  if (h->comm_stream) {
    cudaStreamSynchronize(h->comm_stream);
    cudaStreamDestroy(h->comm_stream);
    cudaEventDestroy(h->ev_comm_in);
    cudaEventDestroy(h->ev_comm_out);
  }
  h->sums_partial.release();
  h->marked_dev.release();
  h->sums_counter.release();
  if (h->sums_stage) cudaFreeHost(h->sums_stage);
  if (h->marked_host) cudaFreeHost(h->marked_host);
  for (Lane& L : h->lanes) {
    if (L.st) cudaStreamSynchronize(L.st);
    if (L.hi) cudaStreamSynchronize(L.hi);
  }
  h->e.release(); h->a.release(); h->lconst.release(); h->chan.release();
  for (Lane& L : h->lanes) L.release();
  Buf* bufs[] = {&h->w_gbe, &h->w_gba, &h->w_mue, &h->w_mua};
  for (Buf* b : bufs) b->release();
  for (Buf& b : h->w_misc) b.release();
  if (h->ev_dev_in) cudaEventDestroy(h->ev_dev_in);
  if (h->ev0) cudaEventDestroy(h->ev0);
  if (h->ev1) cudaEventDestroy(h->ev1);
  for (auto& row : h->ring)
    for (cudaEvent_t e : row)
      if (e) cudaEventDestroy(e);
  for (auto& row : h->tickets)
    for (cudaEvent_t e : row)
      if (e) cudaEventDestroy(e);
  for (Lane& L : h->lanes) {
    if (L.ev_fork) cudaEventDestroy(L.ev_fork);
    if (L.ev_join) cudaEventDestroy(L.ev_join);
    if (L.ev_jac) cudaEventDestroy(L.ev_jac);
    if (L.jac) cudaStreamDestroy(L.jac);
    if (L.ev_mom) cudaEventDestroy(L.ev_mom);
    if (L.ev_tail) cudaEventDestroy(L.ev_tail);
    if (L.hi) cudaStreamDestroy(L.hi);
    if (L.aux) cudaStreamDestroy(L.aux);
    if (L.st) cudaStreamDestroy(L.st);
  }
  delete h;
}

void* chi_stream(chi_handle* h) { return h ? (void*)h->stream : nullptr; }

static void finish_pending_sums(chi_handle* h, int slot);

int chi_synchronize(chi_handle* h) {
  if (!h) return CHI_EINVAL;
  CK(cudaSetDevice(h->device));
  for (Lane& L : h->lanes) {
    CK(cudaStreamSynchronize(L.st));
    if (L.hi) CK(cudaStreamSynchronize(L.hi));
    L.tail_pending = false;
  }
  if (h->sums_stage)
    for (int i = 0; i <= CHI_TICKETS; ++i) finish_pending_sums(h, i);
  return CHI_OK;
}

int chi_wait(chi_handle* h, int64_t ticket) {
  if (!h) return CHI_EINVAL;
  if (ticket < 0) return chi_synchronize(h);
  if (ticket >= h->next_ticket) return fail(h, CHI_EINVAL, "chi_wait: unknown ticket");
  // a ticket that has left the ring: the oldest one still in it was queued later on the same streams,
  // so its completion implies the older one's
  ticket = std::max<int64_t>(ticket, h->next_ticket - CHI_TICKETS);
  CK(cudaSetDevice(h->device));
  for (cudaEvent_t e : h->tickets[ticket % CHI_TICKETS]) CK(cudaEventSynchronize(e));
  // calls complete in issue order: every older ticket's sums are in the staging buffer too
  for (int64_t t = std::max<int64_t>(0, h->next_ticket - CHI_TICKETS); t <= ticket; ++t)
    finish_pending_sums(h, (int)(t % CHI_TICKETS));
  return CHI_OK;
}

// CPUs that are local to the current device's PCIe root (sysfs), intersected with the CPUs this
// process may use; empty when the topology cannot be read.
static bool gpu_local_cpus(cpu_set_t* set) {
  int dev = 0;
  char id[32] = {0};
  if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetPCIBusId(id, sizeof(id), dev) != cudaSuccess) return false;
  for (char* c = id; *c; ++c) *c = (char)tolower(*c);
  const std::string path = std::string("/sys/bus/pci/devices/") + id + "/local_cpulist";
  FILE* f = fopen(path.c_str(), "r");
  if (!f) return false;
  char buf[4096] = {0};
  const bool got = fgets(buf, sizeof(buf), f) != nullptr;
  fclose(f);
  if (!got) return false;
  cpu_set_t allowed;
  CPU_ZERO(&allowed);
  if (sched_getaffinity(0, sizeof(allowed), &allowed) != 0) return false;
  CPU_ZERO(set);
  int n = 0;
  for (char* tok = strtok(buf, ",\n"); tok; tok = strtok(nullptr, ",\n")) {
    int a = 0, b = 0;
    const int k = sscanf(tok, "%d-%d", &a, &b);
    if (k < 1) continue;
    if (k == 1) b = a;
    for (int c = a; c <= b && c < CPU_SETSIZE; ++c)
      if (CPU_ISSET(c, &allowed)) { CPU_SET(c, set); ++n; }
  }
  return n > 0;
}

// NUMA node of the current device (sysfs), -1 when unknown
static int gpu_numa_node() {
  int dev = 0;
  char id[32] = {0};
  if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetPCIBusId(id, sizeof(id), dev) != cudaSuccess) return -1;
  for (char* c = id; *c; ++c) *c = (char)tolower(*c);
  const std::string path = std::string("/sys/bus/pci/devices/") + id + "/numa_node";
  FILE* f = fopen(path.c_str(), "r");
  if (!f) return -1;
  int node = -1;
  if (fscanf(f, "%d", &node) != 1) node = -1;
  fclose(f);
  return node;
}

// Page-locked host memory on the NUMA node of the current device, so that H2D reads and D2H writes do
// not cross the socket interconnect (measured on the bench box: 23 GB/s remote, 55 GB/s local).  Two
// mechanisms, both undone before returning: a PREFERRED memory policy for the GPU's node while the
// pages are allocated and first touched (works even when the process may not run on that node's CPUs
// -- a container pinned to the other socket, the case of ranks 4..7 of an 8-GPU job), and, where those
// CPUs are allowed, running the first touch on them.
int chi_host_alloc(void** out, size_t bytes) {
  if (!out) return CHI_EINVAL;
  cpu_set_t old_set, local;
  const bool numa = !getenv("CHI_NO_NUMA");
  const bool have_old = sched_getaffinity(0, sizeof(old_set), &old_set) == 0;
  const int node = numa ? gpu_numa_node() : -1;
  bool policy = false;
  if (node >= 0 && node < 64) {
    unsigned long mask = 1UL << node;
    policy = syscall(SYS_set_mempolicy, 1 /* MPOL_PREFERRED */, &mask, 65UL) == 0;
  }
  const bool moved = have_old && numa && gpu_local_cpus(&local) && sched_setaffinity(0, sizeof(local), &local) == 0;
  cudaError_t e = cudaMallocHost(out, bytes ? bytes : 1);
  if (e == cudaSuccess && (moved || policy)) memset(*out, 0, bytes ? bytes : 1);
  if (moved) sched_setaffinity(0, sizeof(old_set), &old_set);
  if (policy) syscall(SYS_set_mempolicy, 0 /* MPOL_DEFAULT */, nullptr, 0UL);
  if (e != cudaSuccess) {
    g_create_error = std::string("chi_host_alloc: ") + cudaGetErrorString(e);
    return CHI_ENOMEM;
  }
  return CHI_OK;
}

int chi_host_free(void* p) {
  if (!p) return CHI_OK;
  return cudaFreeHost(p) == cudaSuccess ? CHI_OK : CHI_ECUDA;
}

int chi_set_spectra(chi_handle* h, int n_sightlines, const chi_dataset* emission,
                    const chi_dataset* absorption) {
  if (!h) return CHI_EINVAL;
  if (n_sightlines < 1) return fail(h, CHI_EINVAL, "chi_set_spectra: n_sightlines must be >= 1");
  CK(cudaSetDevice(h->device));
  const bool got_e = emission && emission->n_channels > 0;
  const bool got_a = absorption && absorption->n_channels > 0;
  if (h->has_e && !got_e) return fail(h, CHI_EINVAL, "chi_set_spectra: model needs an emission dataset");
  if (h->has_a && !got_a) return fail(h, CHI_EINVAL, "chi_set_spectra: model needs an absorption dataset");
  // every lane and priority stream: kernels of outstanding asynchronous tickets may still read the
  // buffers replaced below (the tickets are complete when this returns)
  {
    int rc = chi_synchronize(h);
    if (rc) return rc;
  }
  std::vector<double> lconst(n_sightlines, 0.0);
  if (h->has_e) {
    int rc = upload_dataset(h, h->e, emission, n_sightlines, lconst.data(), "emission");
    if (rc) return rc;
    h->map.wch2_e = channel_fwhm2(emission->spectral);
  }
  if (h->has_a) {
    int rc = upload_dataset(h, h->a, absorption, n_sightlines, lconst.data(), "absorption");
    if (rc) return rc;
    h->map.wch2_a = channel_fwhm2(absorption->spectral);
  }
  h->shared_axis = false;
  if (h->has_e && h->has_a && emission->n_channels == absorption->n_channels && !getenv("CHI_NO_FUSED")) {
    h->shared_axis = memcmp(emission->spectral, absorption->spectral, emission->n_channels * sizeof(double)) == 0;
  }
  if (h->shared_axis) {
    // packed channel data of the fused kernel, [sightline][group of 32 channels][plane][32]:
    // planes v, p_e0, p_a0, -, y_e, w_e, y_a, w_a
    const int S = emission->n_channels;
    const size_t groups = ((size_t)S + 31) / 32;
    std::vector<double> chan((size_t)n_sightlines * groups * 256, 0.0);
    for (int i = 0; i < n_sightlines; ++i)
      for (int s = 0; s < S; ++s) {
        double* c = &chan[((size_t)i * groups + s / 32) * 256 + s % 32];
        const size_t k = (size_t)i * S + s;
        const double sig_e = emission->noise[k], sig_a = absorption->noise[k];
        c[0] = emission->spectral[s];
        c[32] = emission->n_baseline > 0 ? emission->basis[s] : 0.0;
        c[64] = absorption->n_baseline > 0 ? absorption->basis[s] : 0.0;
        c[128] = emission->brightness[k] - (emission->offset ? emission->offset[s] : 0.0);
        c[160] = 1.0 / (sig_e * sig_e);
        c[192] = absorption->brightness[k] - (absorption->offset ? absorption->offset[s] : 0.0);
        c[224] = 1.0 / (sig_a * sig_a);
      }
    CK(h->chan.ensure(chan.size() * sizeof(double)));
    CK(cudaMemcpy(h->chan.p, chan.data(), chan.size() * sizeof(double), cudaMemcpyHostToDevice));
  }
  CK(h->lconst.ensure(n_sightlines * sizeof(double)));
  CK(cudaMemcpy(h->lconst.p, lconst.data(), n_sightlines * sizeof(double), cudaMemcpyHostToDevice));
  h->n_sl = n_sightlines;
  h->spectra_set = true;
  return CHI_OK;
}

// ---- physics-level ------------------------------------------------------------
int chi_spin_temp(chi_handle* h, int64_t n, const double* tk, const double* dens, const double* na,
                  double* out) {
  if (!h) return CHI_EINVAL;
  if (n < 0 || (n > 0 && (!tk || !dens || !na || !out))) return fail(h, CHI_EINVAL, "chi_spin_temp: NULL argument");
  if (n == 0) return CHI_OK;
  CK(cudaSetDevice(h->device));
  const size_t nb = (size_t)n * sizeof(double);
  for (int i = 0; i < 4; ++i) CK(h->w_misc[i].ensure(nb));
  CK(cudaMemcpyAsync(h->w_misc[0].p, tk, nb, cudaMemcpyHostToDevice, h->stream));
  CK(cudaMemcpyAsync(h->w_misc[1].p, dens, nb, cudaMemcpyHostToDevice, h->stream));
  CK(cudaMemcpyAsync(h->w_misc[2].p, na, nb, cudaMemcpyHostToDevice, h->stream));
  k_spin_temp<<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>(n, h->w_misc[0].d(), h->w_misc[1].d(),
                                                                 h->w_misc[2].d(), h->w_misc[3].d());
  h->launches++;
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(out, h->w_misc[3].p, nb, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return CHI_OK;
}

int chi_spin_temp_vjp(chi_handle* h, int64_t n, const double* tk, const double* dens, const double* na,
                      const double* gbar, double* g_tk, double* g_dens, double* g_na) {
  if (!h) return CHI_EINVAL;
  if (n < 0 || (n > 0 && (!tk || !dens || !na || !gbar || !g_tk || !g_dens || !g_na)))
    return fail(h, CHI_EINVAL, "chi_spin_temp_vjp: NULL argument");
  if (n == 0) return CHI_OK;
  CK(cudaSetDevice(h->device));
  const size_t nb = (size_t)n * sizeof(double);
  for (int i = 0; i < 7; ++i) CK(h->w_misc[i].ensure(nb));
  CK(cudaMemcpyAsync(h->w_misc[0].p, tk, nb, cudaMemcpyHostToDevice, h->stream));
  CK(cudaMemcpyAsync(h->w_misc[1].p, dens, nb, cudaMemcpyHostToDevice, h->stream));
  CK(cudaMemcpyAsync(h->w_misc[2].p, na, nb, cudaMemcpyHostToDevice, h->stream));
  CK(cudaMemcpyAsync(h->w_misc[3].p, gbar, nb, cudaMemcpyHostToDevice, h->stream));
  k_spin_temp_vjp<<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>(
      n, h->w_misc[0].d(), h->w_misc[1].d(), h->w_misc[2].d(), h->w_misc[3].d(), h->w_misc[4].d(),
      h->w_misc[5].d(), h->w_misc[6].d());
  h->launches++;
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(g_tk, h->w_misc[4].p, nb, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaMemcpyAsync(g_dens, h->w_misc[5].p, nb, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaMemcpyAsync(g_na, h->w_misc[6].p, nb, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return CHI_OK;
}

int chi_elementwise(chi_handle* h, int fn, int64_t n, const double* a, const double* b, const double* c,
                    double* out) {
  if (!h) return CHI_EINVAL;
  if (fn < CHI_FN_GAUSSIAN || fn > CHI_FN_LOG10_NONTHERMAL_PRESSURE)
    return fail(h, CHI_EINVAL, "chi_elementwise: unknown function id");
  const bool three = fn <= CHI_FN_LORENTZIAN;
  if (n < 0 || (n > 0 && (!a || !b || !out || (three && !c))))
    return fail(h, CHI_EINVAL, "chi_elementwise: NULL argument");
  if (n == 0) return CHI_OK;
  CK(cudaSetDevice(h->device));
  const size_t nb = (size_t)n * sizeof(double);
  for (int i = 0; i < 4; ++i) CK(h->w_misc[i].ensure(nb));
  CK(cudaMemcpyAsync(h->w_misc[0].p, a, nb, cudaMemcpyHostToDevice, h->stream));
  CK(cudaMemcpyAsync(h->w_misc[1].p, b, nb, cudaMemcpyHostToDevice, h->stream));
  if (three) CK(cudaMemcpyAsync(h->w_misc[2].p, c, nb, cudaMemcpyHostToDevice, h->stream));
  k_elementwise<<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>(fn, n, h->w_misc[0].d(), h->w_misc[1].d(),
                                                                   h->w_misc[2].d(), h->w_misc[3].d());
  h->launches++;
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(out, h->w_misc[3].p, nb, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return CHI_OK;
}

int chi_pseudo_voigt(chi_handle* h, int S, int N, const double* vax, const double* velocity,
                     const double* fwhm2, const double* fwhm_L, double* out) {
  if (!h) return CHI_EINVAL;
  if (S < 2) return fail(h, CHI_EINVAL, "chi_pseudo_voigt: need at least 2 channels (physics.py:294)");
  if (N < 1 || !vax || !velocity || !fwhm2 || !fwhm_L || !out)
    return fail(h, CHI_EINVAL, "chi_pseudo_voigt: bad argument");
  CK(cudaSetDevice(h->device));
  CK(h->w_misc[0].ensure(S * sizeof(double)));
  for (int i = 1; i < 4; ++i) CK(h->w_misc[i].ensure(N * sizeof(double)));
  CK(h->w_misc[4].ensure((size_t)S * N * sizeof(double)));
  CK(cudaMemcpyAsync(h->w_misc[0].p, vax, S * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  CK(cudaMemcpyAsync(h->w_misc[1].p, velocity, N * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  CK(cudaMemcpyAsync(h->w_misc[2].p, fwhm2, N * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  CK(cudaMemcpyAsync(h->w_misc[3].p, fwhm_L, N * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  int64_t tot = (int64_t)S * N;
  k_pseudo_voigt<<<(unsigned)((tot + 255) / 256), 256, 0, h->stream>>>(
      S, N, h->w_misc[0].d(), h->w_misc[1].d(), h->w_misc[2].d(), h->w_misc[3].d(), h->w_misc[4].d());
  h->launches++;
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(out, h->w_misc[4].p, (size_t)S * N * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return CHI_OK;
}

int chi_radiative_transfer(chi_handle* h, int S, int N, const double* tau, const double* tspin,
                           const double* ff, double bg_temp, double* out) {
  if (!h) return CHI_EINVAL;
  if (S < 1 || N < 1 || !tau || !tspin || !ff || !out)
    return fail(h, CHI_EINVAL, "chi_radiative_transfer: bad argument");
  CK(cudaSetDevice(h->device));
  CK(h->w_misc[0].ensure((size_t)S * N * sizeof(double)));
  CK(h->w_misc[1].ensure(N * sizeof(double)));
  CK(h->w_misc[2].ensure(N * sizeof(double)));
  CK(h->w_misc[3].ensure(S * sizeof(double)));
  CK(cudaMemcpyAsync(h->w_misc[0].p, tau, (size_t)S * N * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  CK(cudaMemcpyAsync(h->w_misc[1].p, tspin, N * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  CK(cudaMemcpyAsync(h->w_misc[2].p, ff, N * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  k_radiative_transfer<<<(S + 127) / 128, 128, 0, h->stream>>>(S, N, h->w_misc[0].d(), h->w_misc[1].d(),
                                                              h->w_misc[2].d(), bg_temp, h->w_misc[3].d());
  h->launches++;
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(out, h->w_misc[3].p, S * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return CHI_OK;
}

int chi_pseudo_voigt_vjp(chi_handle* h, int S, int N, const double* vax, const double* velocity,
                         const double* fwhm2, const double* fwhm_L, const double* gbar, double* g_velocity,
                         double* g_fwhm2, double* g_fwhm_L) {
  if (!h) return CHI_EINVAL;
  if (S < 2) return fail(h, CHI_EINVAL, "chi_pseudo_voigt_vjp: need at least 2 channels (physics.py:294)");
  if (N < 1 || !vax || !velocity || !fwhm2 || !fwhm_L || !gbar || !g_velocity || !g_fwhm2 || !g_fwhm_L)
    return fail(h, CHI_EINVAL, "chi_pseudo_voigt_vjp: bad argument");
  CK(cudaSetDevice(h->device));
  const size_t sb = S * sizeof(double), nb = N * sizeof(double), snb = (size_t)S * N * sizeof(double);
  CK(h->w_misc[0].ensure(sb));
  for (int i = 1; i < 4; ++i) CK(h->w_misc[i].ensure(nb));
  CK(h->w_misc[4].ensure(snb));
  CK(h->w_misc[5].ensure(3 * nb));
  CK(cudaMemcpyAsync(h->w_misc[0].p, vax, sb, cudaMemcpyHostToDevice, h->stream));
  CK(cudaMemcpyAsync(h->w_misc[1].p, velocity, nb, cudaMemcpyHostToDevice, h->stream));
  CK(cudaMemcpyAsync(h->w_misc[2].p, fwhm2, nb, cudaMemcpyHostToDevice, h->stream));
  CK(cudaMemcpyAsync(h->w_misc[3].p, fwhm_L, nb, cudaMemcpyHostToDevice, h->stream));
  CK(cudaMemcpyAsync(h->w_misc[4].p, gbar, snb, cudaMemcpyHostToDevice, h->stream));
  double* g = h->w_misc[5].d();
  k_pseudo_voigt_vjp<<<N, 256, 0, h->stream>>>(S, N, h->w_misc[0].d(), h->w_misc[1].d(), h->w_misc[2].d(),
                                               h->w_misc[3].d(), h->w_misc[4].d(), g, g + N, g + 2 * N);
  h->launches++;
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(g_velocity, g, nb, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaMemcpyAsync(g_fwhm2, g + N, nb, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaMemcpyAsync(g_fwhm_L, g + 2 * N, nb, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return CHI_OK;
}

int chi_radiative_transfer_vjp(chi_handle* h, int S, int N, const double* tau, const double* tspin,
                               const double* ff, double bg_temp, const double* gbar, double* g_tau,
                               double* g_tspin, double* g_ff) {
  if (!h) return CHI_EINVAL;
  if (S < 1 || N < 1 || N > CHI_RT_MAX_CLOUDS || !tau || !tspin || !ff || !gbar || !g_tau || !g_tspin || !g_ff)
    return fail(h, CHI_EINVAL, "chi_radiative_transfer_vjp: bad argument (N must be <= 64)");
  CK(cudaSetDevice(h->device));
  const size_t sb = S * sizeof(double), nb = N * sizeof(double), snb = (size_t)S * N * sizeof(double);
  CK(h->w_misc[0].ensure(snb));
  CK(h->w_misc[1].ensure(nb));
  CK(h->w_misc[2].ensure(nb));
  CK(h->w_misc[3].ensure(sb));
  CK(h->w_misc[4].ensure(snb));
  CK(h->w_misc[5].ensure(2 * nb));
  CK(cudaMemcpyAsync(h->w_misc[0].p, tau, snb, cudaMemcpyHostToDevice, h->stream));
  CK(cudaMemcpyAsync(h->w_misc[1].p, tspin, nb, cudaMemcpyHostToDevice, h->stream));
  CK(cudaMemcpyAsync(h->w_misc[2].p, ff, nb, cudaMemcpyHostToDevice, h->stream));
  CK(cudaMemcpyAsync(h->w_misc[3].p, gbar, sb, cudaMemcpyHostToDevice, h->stream));
  CK(cudaMemsetAsync(h->w_misc[5].p, 0, 2 * nb, h->stream));
  double* g = h->w_misc[5].d();
  k_radiative_transfer_vjp<<<(S + 127) / 128, 128, 0, h->stream>>>(S, N, h->w_misc[0].d(), h->w_misc[1].d(),
                                                                  h->w_misc[2].d(), bg_temp, h->w_misc[3].d(),
                                                                  h->w_misc[4].d(), g, g + N);
  h->launches++;
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(g_tau, h->w_misc[4].p, snb, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaMemcpyAsync(g_tspin, g, nb, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaMemcpyAsync(g_ff, g + N, nb, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return CHI_OK;
}

// ---- model-level, device pointers ------------------------------------------------
// The *_dev entry points may be called while the caller captures the handle's stream into a CUDA
// graph (after one un-captured call of the same size, which allocates the workspace): no timing
// events are recorded then.
struct CaptureAware {
  chi_handle* h;
  bool was;
  explicit CaptureAware(chi_handle* h_) : h(h_), was(h_->capturing) {
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(h->stream, &cs) == cudaSuccess && cs == cudaStreamCaptureStatusActive) h->capturing = true;
  }
  ~CaptureAware() { h->capturing = was; }
};

int chi_logp_grad_dev(chi_handle* h, int64_t B, const double* theta, const double* ce, const double* ca,
                      const int32_t* sl, double* logp, double* grad, double* gce, double* gca) {
  int rc = check_model_call(h, B, theta);
  if (rc) return rc;
  if (!logp) return fail(h, CHI_EINVAL, "chi_logp_grad: logp_out is NULL");
  if (B == 0) return CHI_OK;
  CK(cudaSetDevice(h->device));
  CaptureAware cap_aware(h);
  h->call_first_stage = h->stage_calls;
  h->kernel_ms = 0.f;
  if (h->opt[CHI_OPT_DEV_OVERLAP] && !h->capturing && !batch_is_small(h, B)) {
    // Relaxed stream contract (opt-in): consecutive device-resident calls alternate between two lanes,
    // so the cloud map of call k+1 and the epilogue of call k run next to the other call's channel
    // kernel instead of in front of / behind it.  Each call starts after everything enqueued on the
    // handle's stream so far; its outputs are ordered into the handle's stream by chi_dev_join (or
    // chi_synchronize), not by the call itself.
    // The small kernels (map, epilogue) go to the lane's HIGH-PRIORITY stream: their blocks are placed
    // ahead of the other call's channel-kernel blocks as slots free up, so the map of call k+1 is done
    // when the channel kernel of call k drains and the two channel kernels run back to back.
    Lane& L = h->lanes[1 + (h->dev_calls++ & 1)];
    cudaEvent_t& ev = h->ev_dev_in;
    if (!ev) CK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    CK(cudaEventRecord(ev, h->stream));
    CK(cudaStreamWaitEvent(L.hi, ev, 0));
    std::swap(L.st, L.hi);  // the map phase launches on what run_backward sees as the lane's stream
    rc = run_backward(h, L, B, theta, ce, ca, sl, nullptr, nullptr, false, logp, grad, gce, gca, false, PHASE_MAP);
    std::swap(L.st, L.hi);
    if (rc) return rc;
    CK(cudaStreamWaitEvent(L.st, L.ev_fork, 0));
    h->epilogue_on_hi = true;
    rc = run_backward(h, L, B, theta, ce, ca, sl, nullptr, nullptr, false, logp, grad, gce, gca, false, PHASE_REST);
    h->epilogue_on_hi = false;
    return rc;
  }
  return run_backward(h, h->lanes[0], B, theta, ce, ca, sl, nullptr, nullptr, false, logp, grad, gce, gca);
}

// The handle's stream waits for the device-resident calls issued in CHI_OPT_DEV_OVERLAP mode.
int chi_dev_join(chi_handle* h) {
  if (!h) return CHI_EINVAL;
  CK(cudaSetDevice(h->device));
  for (int i = 1; i < CHI_LANES; ++i) {
    Lane& L = h->lanes[i];
    CK(cudaEventRecord(L.ev_join, L.st));
    CK(cudaStreamWaitEvent(h->stream, L.ev_join, 0));
    CK(cudaEventRecord(L.ev_tail, L.hi));
    CK(cudaStreamWaitEvent(h->stream, L.ev_tail, 0));
  }
  return CHI_OK;
}

int chi_spectra_dev(chi_handle* h, int64_t B, const double* theta, const double* ce, const double* ca,
                    double* mu_e, double* mu_a) {
  int rc = check_model_call(h, B, theta);
  if (rc) return rc;
  if (B == 0) return CHI_OK;
  CK(cudaSetDevice(h->device));
  h->kernel_ms = 0.f;
  return run_spectra(h, B, theta, ce, ca, mu_e, mu_a);
}

int chi_spectra_f32_dev(chi_handle* h, int64_t B, const double* theta, const double* ce, const double* ca,
                        float* mu_e, float* mu_a) {
  int rc = check_model_call(h, B, theta);
  if (rc) return rc;
  if (B == 0) return CHI_OK;
  CK(cudaSetDevice(h->device));
  h->kernel_ms = 0.f;
  return run_spectra32(h, B, theta, ce, ca, mu_e, mu_a);
}

int chi_spectra_vjp_dev(chi_handle* h, int64_t B, const double* theta, const double* ce, const double* ca,
                        const double* gbar_e, const double* gbar_a, double* grad, double* gce, double* gca) {
  int rc = check_model_call(h, B, theta);
  if (rc) return rc;
  if (!grad) return fail(h, CHI_EINVAL, "chi_spectra_vjp: grad_out is NULL");
  if (B == 0) return CHI_OK;
  CK(cudaSetDevice(h->device));
  h->call_first_stage = h->stage_calls;
  h->kernel_ms = 0.f;
  return run_backward(h, h->lanes[0], B, theta, ce, ca, nullptr, gbar_e, gbar_a, true, nullptr, grad, gce, gca);
}

// ---- model-level, host pointers ---------------------------------------------------
// k_shard_sums (comm.cuh) of B draws on stream `st` with the given workspace: out[width + 2] on the device
static int launch_shard_sums(chi_handle* h, cudaStream_t st, Buf& partial, Buf& counter, int64_t B, int width,
                             const double* logp, const double* grad, double* out) {
  if (B == 0) {
    CK(cudaMemsetAsync(out, 0, (size_t)(width + 2) * sizeof(double), st));
    return CHI_OK;
  }
  // enough blocks to stream at HBM speed, few enough that the last block's pass over the partials is short
  const int grid = (int)std::min<int64_t>((B + 63) / 64, (int64_t)h->n_sm * 2);
  const int per = (int)((B + grid - 1) / grid);
  const int grid2 = (int)((B + per - 1) / per);
  CK(partial.ensure((size_t)grid2 * (width + 2) * sizeof(double)));
  if (!counter.p) {
    CK(counter.ensure(sizeof(unsigned)));
    CK(cudaMemsetAsync(counter.p, 0, sizeof(unsigned), st));
  }
  k_shard_sums<<<grid2, CHI_SUMS_THREADS, 0, st>>>(B, width, per, logp, grad, partial.d(), (unsigned*)counter.p, out);
  h->launches++;
  CK(cudaGetLastError());
  return CHI_OK;
}

// adds the per-piece sums of a completed "sums" call in piece order (deterministic) into the caller's array
static void finish_pending_sums(chi_handle* h, int slot) {
  chi_handle::PendingSums& ps = h->pending[slot];
  if (!ps.out) return;
  const int w2 = ps.width + 2;
  const double* st = h->sums_stage + (size_t)slot * CHI_SUMS_MAXPIECES * (CHI_SUMS_MAXW + 2);
  for (int t = 0; t < w2; ++t) {
    double v = 0.0;
    for (int k = 0; k < ps.pieces; ++k) v += st[(size_t)k * (CHI_SUMS_MAXW + 2) + t];
    ps.out[t] = v;
  }
  ps.out = nullptr;
}

// `sums` != null: the per-draw gradient stays on the device; what comes back is logp (if asked for)
// and sums[2 + n_rows N] = {sum logp, sum of every gradient entry over the finite draws, number of
// non-finite draws} (chi_logp_sums_rows)
static int host_logp_grad(chi_handle* h, int64_t B, const double* theta, const double* ce, const double* ca,
                          const int32_t* sl, double* logp, double* grad, double* gce, double* gca, bool model,
                          int64_t* ticket_out = nullptr, double* sums = nullptr) {
  int rc = check_model_call(h, B, theta);
  if (rc) return rc;
  if (model) {
    rc = check_prior_call(h);
    if (rc) return rc;
  }
  if (!logp && !sums) return fail(h, CHI_EINVAL, "chi_logp_grad: logp_out is NULL");
  CK(cudaSetDevice(h->device));
  if (sl)
    for (int64_t i = 0; i < B; ++i)
      if (sl[i] < 0 || sl[i] >= h->n_sl) return fail(h, CHI_EINVAL, "chi_logp_grad: sightline index out of range");
  const int N = h->cfg.n_clouds;
  const int n_rows = h->rm.n_rows;
  const int sums_w = n_rows * N;
  if (sums) {
    if (sums_w > CHI_SUMS_MAXW) return fail(h, CHI_EINVAL, "chi_logp_sums: n_rows * n_clouds > 256");
    if (!h->sums_stage)
      CK(cudaMallocHost(&h->sums_stage, sizeof(double) * (CHI_TICKETS + 1) * CHI_SUMS_MAXPIECES * (CHI_SUMS_MAXW + 2)));
  }
  const size_t th_b = (size_t)n_rows * N * sizeof(double);
  const int nbe = h->has_e ? h->e.nb : 0, nba = h->has_a ? h->a.nb : 0;
  // Sub-batches rotate over the lanes: while one lane computes, the next one uploads its
  // parameters and the previous one downloads its results (stream order keeps each lane's
  // buffers safe).  Large batches are cut into at least 2*CHI_LANES pieces.
  // Piece schedule: small pieces at both ends keep the pipeline's fill (first upload) and drain
  // (last download) short, large ones in the middle keep the kernels efficient:
  // B/24, B/12, B/8, 3 x B/6, B/8, B/12, B/24 -- each capped by the workspace budget.
  const int64_t cap = host_sub_batch(h, 4096);
  std::vector<int64_t> sizes;
  int pieces = 0;
  if (const char* e = getenv("CHI_PIPE_PIECES")) pieces = std::max(1, atoi(e));  // tuning knob: equal pieces
  if (pieces > 0 || B < 8192) {
    int64_t sub = cap;
    if (pieces > 0 && B >= 8192) sub = std::min<int64_t>(cap, std::max<int64_t>(4096, (B + pieces - 1) / pieces));
    for (int64_t b0 = 0; b0 < B; b0 += sub) sizes.push_back(std::min(sub, B - b0));
  } else {
    std::vector<double> den = {24, 12, 8, 6, 6, 6, 8, 12, 24};
    // asynchronous calls overlap with their neighbours: fill and drain are hidden anyway, three equal
    // pieces keep the kernels efficient (measured: 1.69 ms per 1e5-draw call against 1.87 with the ramp)
    if (ticket_out) den = {3, 3, 3};
    if (const char* e = getenv("CHI_PIPE_RAMP")) {  // tuning knob: comma-separated denominators
      den.clear();
      for (const char* c = e; *c;) {
        char* end = nullptr;
        const double v = strtod(c, &end);
        if (end == c) break;
        if (v >= 1.0) den.push_back(v);
        c = (*end == ',') ? end + 1 : end;
      }
      if (den.empty()) den.push_back(1.0);
    }
    const int nd = (int)den.size();
    int64_t left = B;
    for (int i = 0; i < nd && left > 0; ++i) {
      int64_t want = (i == nd - 1) ? left : std::min<int64_t>(left, ((int64_t)(B / den[i]) + 255) / 256 * 256);
      while (want > 0) {  // honour the workspace cap
        const int64_t nb = std::min(want, cap);
        sizes.push_back(nb);
        want -= nb;
        left -= nb;
      }
    }
  }
  h->call_first_stage = h->stage_calls;
  h->kernel_ms = 0.f;
  // upload + cloud map of piece k+1 are issued before the channel kernels of piece k
  struct Piece {
    int64_t b0, nb;
    double *d_ce, *d_ca, *d_grad, *d_gce, *d_gca;
    int32_t* d_sl;
  };
  std::vector<Piece> P(sizes.size());
  auto issue_upload_and_map = [&](size_t k) -> int {
    Lane& L = h->lanes[k % CHI_LANES];
    Piece& p = P[k];
    const int64_t nb = p.nb, b0 = p.b0;
    if (L.tail_pending) {  // the lane's previous piece must have left its buffers
      CK(cudaStreamWaitEvent(L.st, L.ev_tail, 0));
      L.tail_pending = false;
    }
    CK(L.w_theta.ensure(nb * th_b));
    CK(L.w_logp.ensure(nb * sizeof(double)));
    CK(cudaMemcpyAsync(L.w_theta.p, theta + b0 * n_rows * N, nb * th_b, cudaMemcpyHostToDevice, L.st));
    p.d_ce = p.d_ca = p.d_grad = p.d_gce = p.d_gca = nullptr;
    p.d_sl = nullptr;
    if (ce && nbe) {
      CK(L.w_ce.ensure(nb * nbe * sizeof(double)));
      CK(cudaMemcpyAsync(L.w_ce.p, ce + b0 * nbe, nb * nbe * sizeof(double), cudaMemcpyHostToDevice, L.st));
      p.d_ce = L.w_ce.d();
    }
    if (ca && nba) {
      CK(L.w_ca.ensure(nb * nba * sizeof(double)));
      CK(cudaMemcpyAsync(L.w_ca.p, ca + b0 * nba, nb * nba * sizeof(double), cudaMemcpyHostToDevice, L.st));
      p.d_ca = L.w_ca.d();
    }
    if (sl) {
      CK(L.w_sl.ensure(nb * sizeof(int32_t)));
      CK(cudaMemcpyAsync(L.w_sl.p, sl + b0, nb * sizeof(int32_t), cudaMemcpyHostToDevice, L.st));
      p.d_sl = (int32_t*)L.w_sl.p;
    }
    if (grad || sums) { CK(L.w_grad.ensure(nb * th_b)); p.d_grad = L.w_grad.d(); }
    if (gce && nbe) { CK(L.w_gce.ensure(nb * nbe * sizeof(double))); p.d_gce = L.w_gce.d(); }
    if (gca && nba) { CK(L.w_gca.ensure(nb * nba * sizeof(double))); p.d_gca = L.w_gca.d(); }
    return run_backward(h, L, nb, L.w_theta.d(), p.d_ce, p.d_ca, p.d_sl, nullptr, nullptr, false, L.w_logp.d(),
                        p.d_grad, p.d_gce, p.d_gca, model, PHASE_MAP);
  };
  {
    int64_t b0 = 0;
    for (size_t k = 0; k < sizes.size(); ++k) { P[k].b0 = b0; P[k].nb = sizes[k]; b0 += sizes[k]; }
  }
  if (sizes.empty()) {  // B == 0
    if (sums) memset(sums, 0, sizeof(double) * (sums_w + 2));
    return CHI_OK;
  }
  if (sums && sizes.size() > CHI_SUMS_MAXPIECES) return fail(h, CHI_EINVAL, "chi_logp_sums: batch needs too many pieces");
  // staging slot of this call's per-piece sums: the ticket's for asynchronous calls, the last one otherwise
  const int sums_slot = ticket_out ? (int)(h->next_ticket % CHI_TICKETS) : CHI_TICKETS;
  if (sums) finish_pending_sums(h, sums_slot);  // a ticket that left the ring without a wait: complete it now
  struct CallScope {  // h->call_draws is only meaningful while this call issues work
    chi_handle* h;
    ~CallScope() { h->call_draws = 0; }
  } scope{h};
  h->call_draws = B;
  const bool use_hi = sizes.size() > 1 && !getenv("CHI_NO_HI");
  h->map_ahead = 0;
  rc = issue_upload_and_map(0);
  if (rc) return rc;
  for (size_t k = 0; k < sizes.size(); ++k) {
    if (k + 1 < sizes.size()) {
      h->map_ahead = 1;
      rc = issue_upload_and_map(k + 1);
      h->map_ahead = 0;
      if (rc) return rc;
    }
    Lane& L = h->lanes[k % CHI_LANES];
    const Piece& p = P[k];
    const int64_t nb = p.nb, b0 = p.b0;
    h->epilogue_on_hi = use_hi;
    rc = run_backward(h, L, nb, L.w_theta.d(), p.d_ce, p.d_ca, p.d_sl, nullptr, nullptr, false, L.w_logp.d(), p.d_grad,
                      p.d_gce, p.d_gca, model, PHASE_REST);
    h->epilogue_on_hi = false;
    if (rc) return rc;
    cudaStream_t ds = use_hi ? L.hi : L.st;
    if (logp) CK(cudaMemcpyAsync(logp + b0, L.w_logp.p, nb * sizeof(double), cudaMemcpyDeviceToHost, ds));
    if (sums) {
      CK(L.sums_out.ensure(sizeof(double) * (CHI_SUMS_MAXW + 2)));
      rc = launch_shard_sums(h, ds, L.sums_partial, L.sums_counter, nb, sums_w, L.w_logp.d(), p.d_grad, L.sums_out.d());
      if (rc) return rc;
      double* dst = h->sums_stage + ((size_t)sums_slot * CHI_SUMS_MAXPIECES + k) * (CHI_SUMS_MAXW + 2);
      CK(cudaMemcpyAsync(dst, L.sums_out.p, sizeof(double) * (sums_w + 2), cudaMemcpyDeviceToHost, ds));
    }
    if (p.d_grad && grad) CK(cudaMemcpyAsync(grad + b0 * n_rows * N, p.d_grad, nb * th_b, cudaMemcpyDeviceToHost, ds));
    if (p.d_gce) CK(cudaMemcpyAsync(gce + b0 * nbe, p.d_gce, nb * nbe * sizeof(double), cudaMemcpyDeviceToHost, ds));
    if (p.d_gca) CK(cudaMemcpyAsync(gca + b0 * nba, p.d_gca, nb * nba * sizeof(double), cudaMemcpyDeviceToHost, ds));
    if (use_hi) {
      CK(cudaEventRecord(L.ev_tail, L.hi));
      L.tail_pending = true;
    }
  }
  if (sums) {
    h->pending[sums_slot].out = sums;
    h->pending[sums_slot].pieces = (int)sizes.size();
    h->pending[sums_slot].width = sums_w;
  }
  if (ticket_out) {
    // asynchronous call: completion = everything queued on the lanes' streams so far
    const int64_t t = h->next_ticket++;
    cudaEvent_t* ev = h->tickets[t % CHI_TICKETS];
    for (int i = 0; i < CHI_LANES; ++i) {
      for (int j = 0; j < 2; ++j)
        if (!ev[2 * i + j]) CK(cudaEventCreateWithFlags(&ev[2 * i + j], cudaEventDisableTiming));
      CK(cudaEventRecord(ev[2 * i], h->lanes[i].st));
      CK(cudaEventRecord(ev[2 * i + 1], h->lanes[i].hi));
    }
    *ticket_out = t;
    return CHI_OK;
  }
  for (Lane& L : h->lanes) {
    CK(cudaStreamSynchronize(L.st));
    CK(cudaStreamSynchronize(L.hi));
    L.tail_pending = false;
  }
  if (sums) finish_pending_sums(h, sums_slot);
  if (getenv("CHI_TIMELINE")) {  // development aid: when each piece's stages ran, relative to the first
    const int first = h->call_first_stage, n = h->stage_calls - first;
    cudaEvent_t t0 = h->ring[first % CHI_STAGE_RING][0];
    for (int k = 0; k < n && k < CHI_STAGE_RING; ++k) {
      cudaEvent_t* evs = h->ring[(first + k) % CHI_STAGE_RING];
      float t[5];
      for (int i = 0; i < 5; ++i) cudaEventElapsedTime(&t[i], t0, evs[i]);
      fprintf(stderr, "piece %d (%lld draws): start %.3f map-end %.3f moments-end %.3f %.3f epilogue-end %.3f ms\n", k,
              (long long)sizes[k], t[0], t[1], t[2], t[3], t[4]);
    }
  }
  return CHI_OK;
}

int chi_logp_grad(chi_handle* h, int64_t B, const double* theta, const double* ce, const double* ca,
                  const int32_t* sl, double* logp, double* grad, double* gce, double* gca) {
  return host_logp_grad(h, B, theta, ce, ca, sl, logp, grad, gce, gca, false);
}

static int make_row_map(chi_handle* h, int n_rows, const int32_t* slot_of_row, RowMap* out) {
  if (n_rows < 1 || n_rows > CHI_NSLOT || !slot_of_row)
    return fail(h, CHI_EINVAL, "chi_logp_grad_rows: n_rows must be in 1..CHI_NSLOT and slot_of_row non-NULL");
  out->n_rows = n_rows;
  for (int s = 0; s < CHI_NSLOT; ++s) out->row[s] = -1;
  for (int r = 0; r < n_rows; ++r) {
    const int s = slot_of_row[r];
    if (s < 0 || s >= CHI_NSLOT || out->row[s] >= 0)
      return fail(h, CHI_EINVAL, "chi_logp_grad_rows: slot_of_row must hold distinct slots in 0..CHI_NSLOT-1");
    out->row[s] = (signed char)r;
  }
  return CHI_OK;
}

int chi_logp_grad_rows(chi_handle* h, int64_t B, int32_t n_rows, const int32_t* slot_of_row, const double* rows,
                       const double* ce, const double* ca, const int32_t* sl, double* logp, double* grad_rows,
                       double* gce, double* gca) {
  if (!h) return CHI_EINVAL;
  RowMap m;
  int rc = make_row_map(h, n_rows, slot_of_row, &m);
  if (rc) return rc;
  h->rm = m;
  rc = host_logp_grad(h, B, rows, ce, ca, sl, logp, grad_rows, gce, gca, false);
  h->rm = identity_rows();
  return rc;
}

int chi_logp_grad_rows_async(chi_handle* h, int64_t B, int32_t n_rows, const int32_t* slot_of_row, const double* rows,
                             const double* ce, const double* ca, const int32_t* sl, double* logp, double* grad_rows,
                             double* gce, double* gca, int64_t* ticket_out) {
  if (!h || !ticket_out) return CHI_EINVAL;
  RowMap m;
  int rc = make_row_map(h, n_rows, slot_of_row, &m);
  if (rc) return rc;
  if (B == 0) {  // nothing to enqueue: a ticket that is already complete
    *ticket_out = -1;
    return CHI_OK;
  }
  h->rm = m;
  rc = host_logp_grad(h, B, rows, ce, ca, sl, logp, grad_rows, gce, gca, false, ticket_out);
  h->rm = identity_rows();
  return rc;
}

int chi_logp_sums_rows(chi_handle* h, int64_t B, int32_t n_rows, const int32_t* slot_of_row, const double* rows,
                       const double* ce, const double* ca, const int32_t* sl, double* logp, double* sums) {
  if (!h) return CHI_EINVAL;
  if (!sums) return fail(h, CHI_EINVAL, "chi_logp_sums_rows: sums_out is NULL");
  RowMap m;
  int rc = make_row_map(h, n_rows, slot_of_row, &m);
  if (rc) return rc;
  h->rm = m;
  rc = host_logp_grad(h, B, rows, ce, ca, sl, logp, nullptr, nullptr, nullptr, false, nullptr, sums);
  h->rm = identity_rows();
  return rc;
}
int chi_logp_sums_rows_async(chi_handle* h, int64_t B, int32_t n_rows, const int32_t* slot_of_row, const double* rows,
                             const double* ce, const double* ca, const int32_t* sl, double* logp, double* sums,
                             int64_t* ticket_out) {
  if (!h) return CHI_EINVAL;
  if (!sums || !ticket_out) return fail(h, CHI_EINVAL, "chi_logp_sums_rows_async: sums_out / ticket_out is NULL");
  RowMap m;
  int rc = make_row_map(h, n_rows, slot_of_row, &m);
  if (rc) return rc;
  h->rm = m;
  rc = host_logp_grad(h, B, rows, ce, ca, sl, logp, nullptr, nullptr, nullptr, false, ticket_out, sums);
  h->rm = identity_rows();
  return rc;
}
int chi_logp_grad_rows_dev(chi_handle* h, int64_t B, int32_t n_rows, const int32_t* slot_of_row, const double* rows,
                           const double* ce, const double* ca, const int32_t* sl, double* logp, double* grad_rows,
                           double* gce, double* gca) {
  if (!h) return CHI_EINVAL;
  RowMap m;
  int rc = make_row_map(h, n_rows, slot_of_row, &m);
  if (rc) return rc;
  h->rm = m;
  rc = chi_logp_grad_dev(h, B, rows, ce, ca, sl, logp, grad_rows, gce, gca);
  h->rm = identity_rows();
  return rc;
}

int chi_model_logp_grad(chi_handle* h, int64_t B, const double* u, const double* ce, const double* ca,
                        const int32_t* sl, double* logp, double* grad, double* gce, double* gca) {
  return host_logp_grad(h, B, u, ce, ca, sl, logp, grad, gce, gca, true);
}

int chi_model_logp_grad_dev(chi_handle* h, int64_t B, const double* u, const double* ce, const double* ca,
                            const int32_t* sl, double* logp, double* grad, double* gce, double* gca) {
  int rc = check_model_call(h, B, u);
  if (rc) return rc;
  rc = check_prior_call(h);
  if (rc) return rc;
  if (!logp) return fail(h, CHI_EINVAL, "chi_model_logp_grad: logp_out is NULL");
  if (B == 0) return CHI_OK;
  CK(cudaSetDevice(h->device));
  CaptureAware cap_aware(h);
  h->call_first_stage = h->stage_calls;
  h->kernel_ms = 0.f;
  return run_backward(h, h->lanes[0], B, u, ce, ca, sl, nullptr, nullptr, false, logp, grad, gce, gca, true);
}

int chi_set_priors(chi_handle* h, const chi_priors* p) {
  if (!h || !p) return CHI_EINVAL;
  if (h->cfg.model == CHI_MODEL_PHYSICS)
    return fail(h, CHI_EINVAL, "chi_set_priors: CHI_MODEL_PHYSICS has no free RVs with priors");
  if (!(p->beta[0] > 0.0 && p->beta[1] > 0.0)) return fail(h, CHI_EINVAL, "chi_set_priors: Beta parameters must be > 0");
  if (!(p->nth_fwhm_1pc[1] > 0.0)) return fail(h, CHI_EINVAL, "chi_set_priors: nth_fwhm_1pc sigma must be > 0");
  const bool coupled = h->cfg.model == CHI_MODEL_EMISSION_ABSORPTION_PHYSICAL ||
                       (h->cfg.model == CHI_MODEL_EMISSION_ABSORPTION && h->cfg.free_weight);
  if (coupled && !(p->sigma_log10_NHI > 0.0))
    return fail(h, CHI_EINVAL, "chi_set_priors: sigma_log10_NHI must be > 0 for a free absorption weight");
  for (int i = 0; i < CHI_MAX_BASELINE; ++i) {
    h->bsig_e[i] = p->baseline_sigma_e[i] > 0.0 ? p->baseline_sigma_e[i] : 1.0;
    h->bsig_a[i] = p->baseline_sigma_a[i] > 0.0 ? p->baseline_sigma_a[i] : 1.0;
  }
  MapCfg& mc = h->map;
  const double half_log_2pi = 0.5 * log(2.0 * M_PI);
  mc.beta_a = p->beta[0];
  mc.beta_b = p->beta[1];
  mc.beta_lnB = lgamma(mc.beta_a) + lgamma(mc.beta_b) - lgamma(mc.beta_a + mc.beta_b);
  mc.nth_mu = p->nth_fwhm_1pc[0];
  mc.nth_sigma = p->nth_fwhm_1pc[1];
  const double alpha = (0.0 - mc.nth_mu) / mc.nth_sigma;  // lower bound 0, hi_physical_model.py:142
  mc.nth_const = -log(mc.nth_sigma) - half_log_2pi - log(0.5 * erfc(alpha / sqrt(2.0)));
  mc.wt_sigma = coupled ? p->sigma_log10_NHI : 1.0;
  mc.wt_mu0 = -log(10.0) * mc.wt_sigma * mc.wt_sigma / 2.0;  // emission_absorption_model.py:50
  mc.wt_const = -log(mc.wt_sigma) - half_log_2pi;
  mc.fwhm_L_weight = h->cfg.model >= CHI_MODEL_ABSORPTION_PHYSICAL ? 1.0 / h->cfg.n_clouds : 1.0;
  h->priors_set = true;
  return CHI_OK;
}

int chi_transform(chi_handle* h, int64_t B, int direction, const double* in, double* out) {
  if (!h) return CHI_EINVAL;
  if (B < 0 || (B > 0 && (!in || !out)) || (direction != 1 && direction != -1))
    return fail(h, CHI_EINVAL, "chi_transform: bad argument");
  if (B == 0) return CHI_OK;
  CK(cudaSetDevice(h->device));
  const int N = h->cfg.n_clouds;
  const int64_t total = B * CHI_NSLOT * N;
  CK(h->w_misc[0].ensure(total * sizeof(double)));
  CK(h->w_misc[1].ensure(total * sizeof(double)));
  CK(cudaMemcpyAsync(h->w_misc[0].p, in, total * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  unsigned grid = (unsigned)((total + 255) / 256);
  MODEL_SWITCH(h->cfg.model, (k_transform<M><<<grid, 256, 0, h->stream>>>(h->map, total, N, direction,
                                                                        h->w_misc[0].d(), h->w_misc[1].d())));
  h->launches++;
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(out, h->w_misc[1].p, total * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return CHI_OK;
}

int chi_spectra(chi_handle* h, int64_t B, const double* theta, const double* ce, const double* ca,
                double* mu_e, double* mu_a) {
  return host_spectra<double>(h, B, theta, ce, ca, mu_e, mu_a);
}

int chi_spectra_f32(chi_handle* h, int64_t B, const double* theta, const double* ce, const double* ca,
                    float* mu_e, float* mu_a) {
  return host_spectra<float>(h, B, theta, ce, ca, mu_e, mu_a);
}

int chi_predictive(chi_handle* h, int64_t B, const double* theta, const double* ce, const double* ca,
                   const int32_t* sl, uint64_t seed, int64_t draw_offset, double* y_e, double* y_a) {
  NoiseSpec ns{seed, draw_offset, nullptr};
  return host_spectra<double>(h, B, theta, ce, ca, y_e, y_a, &ns, sl);
}

int chi_predictive_f32(chi_handle* h, int64_t B, const double* theta, const double* ce, const double* ca,
                       const int32_t* sl, uint64_t seed, int64_t draw_offset, float* y_e, float* y_a) {
  NoiseSpec ns{seed, draw_offset, nullptr};
  return host_spectra<float>(h, B, theta, ce, ca, y_e, y_a, &ns, sl);
}

int chi_predictive_dev(chi_handle* h, int64_t B, const double* theta, const double* ce, const double* ca,
                       const int32_t* sl, uint64_t seed, int64_t draw_offset, double* y_e, double* y_a) {
  int rc = chi_spectra_dev(h, B, theta, ce, ca, y_e, y_a);
  if (rc || B == 0) return rc;
  NoiseSpec ns{seed, draw_offset, sl};
  return add_noise<double>(h, B, ns, y_e, y_a);
}

int chi_predictive_f32_dev(chi_handle* h, int64_t B, const double* theta, const double* ce, const double* ca,
                           const int32_t* sl, uint64_t seed, int64_t draw_offset, float* y_e, float* y_a) {
  int rc = chi_spectra_f32_dev(h, B, theta, ce, ca, y_e, y_a);
  if (rc || B == 0) return rc;
  NoiseSpec ns{seed, draw_offset, sl};
  return add_noise<float>(h, B, ns, y_e, y_a);
}

int chi_spectra_vjp(chi_handle* h, int64_t B, const double* theta, const double* ce, const double* ca,
                    const double* gbar_e, const double* gbar_a, double* grad, double* gce, double* gca) {
  int rc = check_model_call(h, B, theta);
  if (rc) return rc;
  if (!grad) return fail(h, CHI_EINVAL, "chi_spectra_vjp: grad_out is NULL");
  CK(cudaSetDevice(h->device));
  const int N = h->cfg.n_clouds;
  const size_t th_b = (size_t)CHI_NSLOT * N * sizeof(double);
  const int nbe = h->has_e ? h->e.nb : 0, nba = h->has_a ? h->a.nb : 0;
  const int Se = h->has_e ? h->e.S : 0, Sa = h->has_a ? h->a.S : 0;
  const int64_t sub = host_sub_batch(h, (size_t)(Se + Sa) * sizeof(double) + 4096);
  h->kernel_ms = 0.f;
  for (int64_t b0 = 0; b0 < B; b0 += sub) {
    const int64_t nb = std::min(sub, B - b0);
    CK(h->lanes[0].w_theta.ensure(nb * th_b));
    CK(h->lanes[0].w_grad.ensure(nb * th_b));
    CK(cudaMemcpyAsync(h->lanes[0].w_theta.p, theta + b0 * CHI_NSLOT * N, nb * th_b, cudaMemcpyHostToDevice, h->stream));
    double *d_ce = nullptr, *d_ca = nullptr, *d_ge = nullptr, *d_ga = nullptr, *d_gce = nullptr, *d_gca = nullptr;
    if (ce && nbe) {
      CK(h->lanes[0].w_ce.ensure(nb * nbe * sizeof(double)));
      CK(cudaMemcpyAsync(h->lanes[0].w_ce.p, ce + b0 * nbe, nb * nbe * sizeof(double), cudaMemcpyHostToDevice, h->stream));
      d_ce = h->lanes[0].w_ce.d();
    }
    if (ca && nba) {
      CK(h->lanes[0].w_ca.ensure(nb * nba * sizeof(double)));
      CK(cudaMemcpyAsync(h->lanes[0].w_ca.p, ca + b0 * nba, nb * nba * sizeof(double), cudaMemcpyHostToDevice, h->stream));
      d_ca = h->lanes[0].w_ca.d();
    }
    if (gbar_e && Se) {
      CK(h->w_gbe.ensure((size_t)nb * Se * sizeof(double)));
      CK(cudaMemcpyAsync(h->w_gbe.p, gbar_e + b0 * Se, (size_t)nb * Se * sizeof(double), cudaMemcpyHostToDevice, h->stream));
      d_ge = h->w_gbe.d();
    }
    if (gbar_a && Sa) {
      CK(h->w_gba.ensure((size_t)nb * Sa * sizeof(double)));
      CK(cudaMemcpyAsync(h->w_gba.p, gbar_a + b0 * Sa, (size_t)nb * Sa * sizeof(double), cudaMemcpyHostToDevice, h->stream));
      d_ga = h->w_gba.d();
    }
    if (gce && nbe) { CK(h->lanes[0].w_gce.ensure(nb * nbe * sizeof(double))); d_gce = h->lanes[0].w_gce.d(); }
    if (gca && nba) { CK(h->lanes[0].w_gca.ensure(nb * nba * sizeof(double))); d_gca = h->lanes[0].w_gca.d(); }
    rc = run_backward(h, h->lanes[0], nb, h->lanes[0].w_theta.d(), d_ce, d_ca, nullptr, d_ge, d_ga, true, nullptr, h->lanes[0].w_grad.d(), d_gce, d_gca);
    if (rc) return rc;
    CK(cudaMemcpyAsync(grad + b0 * CHI_NSLOT * N, h->lanes[0].w_grad.p, nb * th_b, cudaMemcpyDeviceToHost, h->stream));
    if (d_gce) CK(cudaMemcpyAsync(gce + b0 * nbe, d_gce, nb * nbe * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    if (d_gca) CK(cudaMemcpyAsync(gca + b0 * nba, d_gca, nb * nba * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    rc = finish_timing(h);
    if (rc) return rc;
  }
  return CHI_OK;
}

int chi_deterministics(chi_handle* h, int64_t B, const double* theta, double* det_out) {
  if (!h) return CHI_EINVAL;
  if (B < 0 || (B > 0 && (!theta || !det_out))) return fail(h, CHI_EINVAL, "chi_deterministics: NULL argument");
  if (B == 0) return CHI_OK;
  CK(cudaSetDevice(h->device));
  const int N = h->cfg.n_clouds;
  const size_t th_b = (size_t)CHI_NSLOT * N * sizeof(double), det_b = (size_t)CHI_NDET * N * sizeof(double);
  const int64_t sub = host_sub_batch(h, 4096);
  for (int64_t b0 = 0; b0 < B; b0 += sub) {
    const int64_t nb = std::min(sub, B - b0);
    CK(h->lanes[0].w_theta.ensure(nb * th_b));
    CK(h->w_misc[0].ensure(nb * det_b));
    CK(cudaMemcpyAsync(h->lanes[0].w_theta.p, theta + b0 * CHI_NSLOT * N, nb * th_b, cudaMemcpyHostToDevice, h->stream));
    unsigned grid = (unsigned)((nb * N + 127) / 128);
    MODEL_SWITCH(h->cfg.model, (k_deterministics<M><<<grid, 128, 0, h->stream>>>(h->map, nb, N, h->lanes[0].w_theta.d(),
                                                                                 h->w_misc[0].d())));
    h->launches++;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(det_out + b0 * CHI_NDET * N, h->w_misc[0].p, nb * det_b, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
  }
  return CHI_OK;
}

// ---- measurement ---------------------------------------------------------------
int chi_last_kernel_ms(chi_handle* h, float* ms_out) {
  if (!h || !ms_out) return CHI_EINVAL;
  CK(cudaSetDevice(h->device));
  int rc = finish_timing(h);  // spectra path (ev0/ev1 on the public stream)
  if (rc) return rc;
  float total = h->kernel_ms;
  // logp / VJP path: sum of the stage events of the sub-batches of the last call
  int64_t first = std::max(h->call_first_stage, h->stage_calls - CHI_STAGE_RING);
  for (int64_t k = first; k < h->stage_calls; ++k) {
    cudaEvent_t* evs = h->ring[k % CHI_STAGE_RING];
    CK(cudaEventSynchronize(evs[4]));
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, evs[0], evs[4]));
    total += ms;
  }
  *ms_out = total;
  return CHI_OK;
}

int chi_stage_ms(chi_handle* h, int n, float* ms_out) {
  if (!h || !ms_out || n < 1) return CHI_EINVAL;
  if (n > CHI_STAGE_RING || n > h->stage_calls)
    return fail(h, CHI_ESTATE, "chi_stage_ms: fewer than n logp/vjp batches recorded (ring holds 64)");
  CK(cudaSetDevice(h->device));
  for (int k = 0; k < n; ++k) {
    cudaEvent_t* evs = h->ring[(h->stage_calls - n + k) % CHI_STAGE_RING];
    CK(cudaEventSynchronize(evs[4]));
    for (int i = 0; i < 4; ++i) CK(cudaEventElapsedTime(&ms_out[k * 4 + i], evs[i], evs[i + 1]));
  }
  return CHI_OK;
}

int64_t chi_launch_count(const chi_handle* h) { return h ? h->launches : 0; }

// ---- multi-GPU: shard sums and the NCCL collectives that gather them (comm.cuh) --------------------
static NcclApi g_nccl;

static int nccl_fail(chi_handle* h, const char* what, int rc) {
  return fail(h, CHI_ECUDA, std::string(what) + ": " + (g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "NCCL error"));
}

int chi_comm_unique_id(void* id_out) {
  if (!id_out) return CHI_EINVAL;
  const char* why = "";
  if (!g_nccl.load(&why)) { g_create_error = why; return CHI_ESTATE; }
  NcclId id;
  const int rc = g_nccl.GetUniqueId(&id);
  if (rc) { g_create_error = std::string("ncclGetUniqueId: ") + g_nccl.GetErrorString(rc); return CHI_ECUDA; }
  memcpy(id_out, &id, sizeof(id));
  return CHI_OK;
}

int chi_comm_init(chi_handle* h, const void* id, int rank, int n_ranks) {
  if (!h) return CHI_EINVAL;
  if (!id || n_ranks < 1 || rank < 0 || rank >= n_ranks) return fail(h, CHI_EINVAL, "chi_comm_init: bad rank / n_ranks / id");
  if (h->comm) return fail(h, CHI_ESTATE, "chi_comm_init: the handle already has a communicator");
  const char* why = "";
  if (!g_nccl.load(&why)) return fail(h, CHI_ESTATE, why);
  CK(cudaSetDevice(h->device));
  NcclId nid;
  memcpy(&nid, id, sizeof(nid));
  const int rc = g_nccl.CommInitRank(&h->comm, n_ranks, nid, rank);
  if (rc) { h->comm = nullptr; return nccl_fail(h, "ncclCommInitRank", rc); }
  h->comm_rank = rank;
  h->comm_size = n_ranks;
  return CHI_OK;
}

int chi_comm_destroy(chi_handle* h) {
  if (!h) return CHI_EINVAL;
  if (h->comm) {
    cudaSetDevice(h->device);
    if (h->comm_stream) cudaStreamSynchronize(h->comm_stream);
    g_nccl.CommDestroy(h->comm);
    h->comm = nullptr;
  }
  h->comm_rank = 0;
  h->comm_size = 1;
  return CHI_OK;
}

int chi_comm_rank(const chi_handle* h) { return h ? h->comm_rank : 0; }
int chi_comm_size(const chi_handle* h) { return h ? h->comm_size : 1; }

// stream the shard sums and the collectives run on: the communication stream once a communicator (or a
// previous call) created it, else the handle's stream
static int comm_stream_of(chi_handle* h, cudaStream_t* st) {
  if (!h->comm_stream) {
    int lo = 0, hi = 0;
    CK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    CK(cudaStreamCreateWithPriority(&h->comm_stream, cudaStreamNonBlocking, hi));
    CK(cudaEventCreateWithFlags(&h->ev_comm_in, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&h->ev_comm_out, cudaEventDisableTiming));
  }
  // everything enqueued on the handle's stream so far is visible to the communication stream
  CK(cudaEventRecord(h->ev_comm_in, h->stream));
  CK(cudaStreamWaitEvent(h->comm_stream, h->ev_comm_in, 0));
  *st = h->comm_stream;
  return CHI_OK;
}

int chi_shard_sums_dev(chi_handle* h, int64_t B, int32_t width, const double* logp, const double* grad, double* sums_out) {
  if (!h) return CHI_EINVAL;
  if (B < 0 || width < 1 || width > CHI_SUMS_MAXW || !logp || !grad || !sums_out)
    return fail(h, CHI_EINVAL, "chi_shard_sums: bad arguments (1 <= width <= 256)");
  CK(cudaSetDevice(h->device));
  cudaStream_t st;
  int rc = comm_stream_of(h, &st);
  if (rc) return rc;
  return launch_shard_sums(h, st, h->sums_partial, h->sums_counter, B, width, logp, grad, sums_out);
}

int chi_allgather_dev(chi_handle* h, const double* send, double* recv, int64_t count) {
  if (!h) return CHI_EINVAL;
  if (!send || !recv || count < 1) return fail(h, CHI_EINVAL, "chi_allgather: bad arguments");
  CK(cudaSetDevice(h->device));
  cudaStream_t st;
  int rc = comm_stream_of(h, &st);
  if (rc) return rc;
  if (!h->comm) {  // one rank: the gather is a copy
    if (recv != send) CK(cudaMemcpyAsync(recv, send, (size_t)count * sizeof(double), cudaMemcpyDeviceToDevice, st));
    return CHI_OK;
  }
  rc = g_nccl.AllGather(send, recv, (size_t)count, CHI_NCCL_FLOAT64, h->comm, st);
  if (rc) return nccl_fail(h, "ncclAllGather", rc);
  return CHI_OK;
}

int chi_allreduce_sum_dev(chi_handle* h, double* buf, int64_t count) {
  if (!h) return CHI_EINVAL;
  if (!buf || count < 1) return fail(h, CHI_EINVAL, "chi_allreduce_sum: bad arguments");
  CK(cudaSetDevice(h->device));
  cudaStream_t st;
  int rc = comm_stream_of(h, &st);
  if (rc) return rc;
  if (!h->comm) return CHI_OK;
  rc = g_nccl.AllReduce(buf, buf, (size_t)count, CHI_NCCL_FLOAT64, CHI_NCCL_SUM, h->comm, st);
  if (rc) return nccl_fail(h, "ncclAllReduce", rc);
  return CHI_OK;
}

// the handle's stream waits for the communication stream (before the caller reuses a buffer a
// collective still reads, or reads one it writes)
int chi_comm_join(chi_handle* h) {
  if (!h) return CHI_EINVAL;
  if (!h->comm_stream) return CHI_OK;
  CK(cudaSetDevice(h->device));
  CK(cudaEventRecord(h->ev_comm_out, h->comm_stream));
  CK(cudaStreamWaitEvent(h->stream, h->ev_comm_out, 0));
  return CHI_OK;
}

// the communication stream (cudaStream_t as void*; created on first use) for callers that order their
// own work against the collectives with events
void* chi_comm_stream(chi_handle* h) {
  if (!h) return nullptr;
  if (!h->comm_stream) {
    cudaStream_t st;
    if (comm_stream_of(h, &st) != CHI_OK) return nullptr;
  }
  return (void*)h->comm_stream;
}

// host-blocking wait for the communication stream
int chi_comm_synchronize(chi_handle* h) {
  if (!h) return CHI_EINVAL;
  if (!h->comm_stream) return CHI_OK;
  CK(cudaSetDevice(h->device));
  CK(cudaStreamSynchronize(h->comm_stream));
  return CHI_OK;
}

int chi_set_option(chi_handle* h, int option, int64_t value) {
  if (!h) return CHI_EINVAL;
  if (option < 0 || option >= CHI_OPT_COUNT) return fail(h, CHI_EINVAL, "chi_set_option: unknown option");
  h->opt[option] = value;
  return CHI_OK;
}
int64_t chi_stage_count(const chi_handle* h) { return h ? h->stage_calls : 0; }

int chi_microbench_fp64(chi_handle* h, double* tflops_out) {
  if (!h || !tflops_out) return CHI_EINVAL;
  CK(cudaSetDevice(h->device));
  CK(h->w_misc[7].ensure(256));
  const int iters = 4096, blocks = h->n_sm * 8, threads = 256;
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a));
  CK(cudaEventCreate(&b));
  float best = 1e30f;
  for (int rep = 0; rep < 5; ++rep) {
    CK(cudaEventRecord(a, h->stream));
    k_dfma_peak<<<blocks, threads, 0, h->stream>>>(h->w_misc[7].d(), iters, 1.0000001, 1e-9);
    h->launches++;
    CK(cudaEventRecord(b, h->stream));
    CK(cudaEventSynchronize(b));
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, a, b));
    if (rep > 0) best = std::min(best, ms);
  }
  cudaEventDestroy(a);
  cudaEventDestroy(b);
  const double flops = 2.0 * 8.0 * (double)iters * blocks * threads;
  *tflops_out = flops / (best * 1e-3) / 1e12;
  return CHI_OK;
}

// FP32 FMA peak (TFLOP/s) and SFU peak (tera MUFU.EX2 per second) of the handle's device
int chi_microbench_fp32(chi_handle* h, double* tflops_out, double* tera_sfu_out) {
  if (!h || !tflops_out || !tera_sfu_out) return CHI_EINVAL;
  CK(cudaSetDevice(h->device));
  CK(h->w_misc[7].ensure(256));
  const int iters = 4096, blocks = h->n_sm * 8, threads = 256;
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a));
  CK(cudaEventCreate(&b));
  float best[2] = {1e30f, 1e30f};
  for (int which = 0; which < 2; ++which)
    for (int rep = 0; rep < 5; ++rep) {
      CK(cudaEventRecord(a, h->stream));
      if (which == 0) k_ffma_peak<<<blocks, threads, 0, h->stream>>>((float*)h->w_misc[7].p, iters, 1.0000001f, 1e-9f);
      else k_sfu_peak<<<blocks, threads, 0, h->stream>>>((float*)h->w_misc[7].p, iters);
      h->launches++;
      CK(cudaEventRecord(b, h->stream));
      CK(cudaEventSynchronize(b));
      float ms = 0.f;
      CK(cudaEventElapsedTime(&ms, a, b));
      if (rep > 0) best[which] = std::min(best[which], ms);
    }
  cudaEventDestroy(a);
  cudaEventDestroy(b);
  *tflops_out = 2.0 * 16.0 * (double)iters * blocks * threads / (best[0] * 1e-3) / 1e12;
  *tera_sfu_out = 8.0 * (double)iters * blocks * threads / (best[1] * 1e-3) / 1e12;
  return CHI_OK;
}

// ---- chain-batched NUTS ---------------------------------------------------
int chi_nuts_sample(chi_handle* h, const chi_nuts_config* cfg, const double* init_u, const double* init_ce,
                    const double* init_ca, const int32_t* sightline, double* draws_u, double* draws_ce,
                    double* draws_ca, double* stat_accept, int32_t* stat_depth, int32_t* stat_diverged,
                    double* stat_logp, double* stat_step) {
  int rc = check_model_call(h, cfg ? cfg->chains : -1, init_u);
  if (rc) return rc;
  rc = check_prior_call(h);
  if (rc) return rc;
  if (!cfg || cfg->chains < 1 || cfg->n_tune < 0 || cfg->n_draws < 1 || !draws_u)
    return fail(h, CHI_EINVAL, "chi_nuts_sample: bad configuration");
  const int max_depth = cfg->max_treedepth > 0 ? std::min(cfg->max_treedepth, CHI_NUTS_MAX_DEPTH) : 10;
  const bool pool_mass = cfg->pool_mass != 0 && !sightline && cfg->chains > 1;
  const bool lock_step = cfg->lock_step != 0 || pool_mass;  // pooled windows close for all chains at once
  const double target = cfg->target_accept > 0.0 ? cfg->target_accept : 0.8;
  CK(cudaSetDevice(h->device));
  struct GraphPath {  // every evaluation of the sampler takes the path that is best inside a graph
    chi_handle* h;
    explicit GraphPath(chi_handle* h_) : h(h_) { h->graph_path = true; }
    ~GraphPath() { h->graph_path = false; }
  } graph_path(h);
  const int64_t B = cfg->chains;
  const int N = h->cfg.n_clouds, nu = CHI_NSLOT * N;
  const int nbe = h->has_e ? h->e.nb : 0, nba = h->has_a ? h->a.nb : 0;
  const int D = nu + nbe + nba;
  cudaStream_t st = h->stream;

  // roles of the position vector's dimensions
  std::vector<int> role(D, 0);
  int n_free = 0;
  for (int s = 0; s < CHI_NSLOT; ++s) {
    int dist = PD_NONE;
    MODEL_SWITCH(h->cfg.model, (dist = slot_dist<M>(h->map, s)));
    if (dist == PD_NONE) continue;
    const bool tied = s == CHI_FR_FWHM_L_NORM && h->cfg.model >= CHI_MODEL_ABSORPTION_PHYSICAL;
    for (int n = 0; n < N; ++n) {
      role[s * N + n] = (tied && n > 0) ? 2 : 1;
      n_free += role[s * N + n] == 1;
    }
  }
  for (int i = 0; i < nbe + nba; ++i) { role[nu + i] = 1; ++n_free; }

  // one arena for the whole sampler state
  const size_t vecs = 22, scal_d = 11, scal_i = 10;
  const size_t n_vec = (size_t)B * D, n_ck = (size_t)B * CHI_NUTS_MAX_DEPTH * D;
  const size_t bytes_d = (vecs * n_vec + 2 * n_ck + scal_d * B + 2 * (size_t)B * nu + 2 * (size_t)B * (nbe + nba) + B) * sizeof(double);
  const size_t bytes_i = (scal_i * B + D + 2) * sizeof(int) + (size_t)B * sizeof(unsigned long long);
  const int64_t n_keep = cfg->n_draws;
  const size_t bytes_out = (size_t)n_keep * B * (D + 3) * sizeof(double) + (size_t)n_keep * B * 2 * sizeof(int);
  if (sightline)
    for (int64_t i = 0; i < B; ++i)
      if (sightline[i] < 0 || sightline[i] >= h->n_sl) return fail(h, CHI_EINVAL, "chi_nuts_sample: sightline out of range");
  ScopedBuf arena, slbuf, abuf;
  ScopedPinned flags;
  CK(arena.ensure(bytes_d + bytes_i + bytes_out + 4096));
  CK(cudaMemsetAsync(arena.p, 0, bytes_d + bytes_i + 1024, st));
  char* cur = (char*)arena.p;
  auto take_d = [&](size_t n) { double* p = (double*)cur; cur += n * sizeof(double); return p; };
  NutsState s;
  memset(&s, 0, sizeof(s));
  s.B = B; s.N = N; s.D = D; s.nbe = nbe; s.nba = nba;
  s.eu = take_d((size_t)B * nu); s.egu = take_d((size_t)B * nu);
  s.ece = take_d((size_t)B * std::max(nbe, 1)); s.egce = take_d((size_t)B * std::max(nbe, 1));
  s.eca = take_d((size_t)B * std::max(nba, 1)); s.egca = take_d((size_t)B * std::max(nba, 1));
  s.elogp = take_d(B);
  double** vec_ptrs[] = {&s.q, &s.g, &s.ql, &s.pl, &s.gl, &s.qr, &s.pr, &s.gr, &s.qp, &s.gp, &s.qs, &s.gs,
                         &s.qc, &s.pc, &s.gc, &s.ph, &s.rsum, &s.rsum_s, &s.inv_mass, &s.w_mean, &s.w_m2};
  for (double** pp : vec_ptrs) *pp = take_d(n_vec);
  s.ck_r = take_d(n_ck); s.ck_rs = take_d(n_ck);
  double** sc_ptrs[] = {&s.logp, &s.logp_p, &s.logp_s, &s.energy0, &s.lw_tree, &s.lw_sub, &s.sum_acc, &s.eps,
                        &s.da_mu, &s.da_hbar, &s.da_logeps_bar};
  for (double** pp : sc_ptrs) *pp = take_d(B);
  s.rng = (unsigned long long*)cur; cur += (size_t)B * sizeof(unsigned long long);
  auto take_i = [&](size_t n) { int* p = (int*)cur; cur += n * sizeof(int); return p; };
  int** int_ptrs[] = {&s.n_leap, &s.dir, &s.active, &s.sub_stop, &s.diverged, &s.depth, &s.da_count, &s.w_count, &s.leaf};
  for (int** pp : int_ptrs) *pp = take_i(B);
  int* d_role = take_i(D);
  s.role = d_role;
  s.n_active = take_i(1);
  s.n_stopped = take_i(1);
  cur = (char*)(((uintptr_t)cur + 255) & ~(uintptr_t)255);
  double* d_draws = (double*)cur; cur += (size_t)n_keep * B * D * sizeof(double);
  double* d_acc = (double*)cur; cur += (size_t)n_keep * B * sizeof(double);
  double* d_lp = (double*)cur; cur += (size_t)n_keep * B * sizeof(double);
  double* d_eps = (double*)cur; cur += (size_t)n_keep * B * sizeof(double);
  int* d_depth = (int*)cur; cur += (size_t)n_keep * B * sizeof(int);
  int* d_div = (int*)cur; cur += (size_t)n_keep * B * sizeof(int);
  int32_t* d_sl = nullptr;
  if (sightline) {
    CK(slbuf.ensure(B * sizeof(int32_t)));
    CK(cudaMemcpyAsync(slbuf.p, sightline, B * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    d_sl = (int32_t*)slbuf.p;
  }
  CK(cudaMemcpyAsync(d_role, role.data(), D * sizeof(int), cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(s.eu, init_u, (size_t)B * nu * sizeof(double), cudaMemcpyHostToDevice, st));
  if (nbe && init_ce) CK(cudaMemcpyAsync(s.ece, init_ce, (size_t)B * nbe * sizeof(double), cudaMemcpyHostToDevice, st));
  if (nba && init_ca) CK(cudaMemcpyAsync(s.eca, init_ca, (size_t)B * nba * sizeof(double), cudaMemcpyHostToDevice, st));

  const unsigned grid = (unsigned)((B + 63) / 64);
  auto evaluate = [&]() -> int {
    h->call_first_stage = h->stage_calls;
    return run_backward(h, h->lanes[0], B, s.eu, nbe ? s.ece : nullptr, nba ? s.eca : nullptr, d_sl, nullptr, nullptr,
                        false, s.elogp, s.egu, nbe ? s.egce : nullptr, nba ? s.egca : nullptr, true);
  };
  const double eps0 = cfg->init_step > 0.0 ? cfg->init_step : 0.25 / pow((double)std::max(n_free, 1), 0.25);
  k_nuts_init<<<grid, 64, 0, st>>>(s, (unsigned long long)cfg->seed, eps0);
  // the per-chain kernel (nuts_chain.cuh) evaluates a draw with one block; the starting point goes through the
  // same arithmetic (k_eval_small, one block per draw)
  int chain_cluster = lock_step ? 0 : nuts_chain_cluster(h, B);
  int chain_threads = (lock_step || chain_cluster) ? 0 : nuts_chain_threads(h, B);
  SmallArgs chain_args;
  int chain_nl = 1;
  size_t chain_dyn = 0;
  if (chain_cluster) {
    rc = small_args(h, h->lanes[0], B, s.eu, nbe ? s.ece : nullptr, nba ? s.eca : nullptr, d_sl, s.elogp, s.egu,
                    nbe ? s.egce : nullptr, nba ? s.egca : nullptr, true, chain_cluster, &chain_args, &chain_nl, &chain_dyn);
    if (rc) return rc;
    chain_cluster = chain_args.n_chunks;  // (what the channel count allows)
    if (chain_cluster < 2 || chain_cluster > 8) {  // a model this short per block: the single-block form
      chain_cluster = 0;
      chain_threads = CHI_SM_THREADS;
    }
  }
  const bool chain_kernel = chain_threads != 0 || chain_cluster != 0;
  if (chain_cluster) {
    if (launch_eval_small_cl(chain_nl, h->cfg.lorentzian != 0, h->map, chain_args, chain_dyn, st))
      return fail(h, CHI_EINVAL, "unsupported configuration of the chain kernel");
    h->launches++;
    CK(cudaGetLastError());
  } else if (chain_kernel) {
    rc = small_args(h, h->lanes[0], B, s.eu, nbe ? s.ece : nullptr, nba ? s.eca : nullptr, d_sl, s.elogp, s.egu,
                    nbe ? s.egce : nullptr, nba ? s.egca : nullptr, true, 1, &chain_args, &chain_nl, &chain_dyn,
                    chain_threads == 512 ? 512 : CHI_SM_THREADS);
    if (rc) return rc;
    if (chain_args.n_chunks != 1) return fail(h, CHI_ECUDA, "chi_nuts_sample: the chain kernel needs one block per draw");
    if (launch_eval_small(chain_nl, h->cfg.lorentzian != 0, h->map, chain_args, chain_dyn, st))
      return fail(h, CHI_EINVAL, "unsupported configuration of the chain kernel");
    h->launches++;
    CK(cudaGetLastError());
  } else {
    rc = evaluate();
  }
  if (rc) return rc;
  k_nuts_adopt<<<grid, 64, 0, st>>>(s);
  h->launches += 2;

  // Stan-style warm-up schedule: fast buffer | doubling slow windows | fast buffer
  const int n_tune = cfg->n_tune;
  int init_buf = 75, term_buf = 50, base_win = 25;
  if (n_tune < 150) { init_buf = (int)(0.15 * n_tune); term_buf = (int)(0.1 * n_tune); base_win = n_tune - init_buf - term_buf; }
  std::vector<char> closes(std::max(n_tune, 1), 0);
  {
    int start = init_buf, win = base_win;
    const int slow_end = n_tune - term_buf;
    while (win > 0 && start < slow_end) {
      int end = start + win;
      if (end + 2 * win > slow_end) end = slow_end;  // the last window absorbs the remainder
      closes[end - 1] = 1;
      start = end;
      win *= 2;
    }
  }
  // one leapfrog step (half kick + drift, batched logp+grad, half kick + tree bookkeeping) as a
  // CUDA graph: ~8 dependent launches per step would otherwise be pure launch latency for a few
  // dozen chains
  CK(cudaMallocHost(&flags.p, 2 * sizeof(int)));
  int* h_flags = flags.p;
  const int total = n_tune + cfg->n_draws;
  if (!lock_step) {
    // asynchronous chains (nuts.cuh: k_nuts_tick): one graph = tick + batched evaluation, replayed
    // until every chain has done its n_tune + n_draws iterations; the host looks at one counter
    // every 32 ticks
    const size_t ab = (2 * (size_t)B + 4) * sizeof(int) + (size_t)std::max(n_tune, 1);
    CK(abuf.ensure(ab));
    CK(cudaMemsetAsync(abuf.p, 0, ab, st));
    NutsAsync A;
    memset(&A, 0, sizeof(A));
    A.n_tune = n_tune; A.n_total = total; A.init_buf = init_buf; A.term_buf = term_buf; A.max_depth = max_depth;
    A.target_accept = target;
    A.iter = (int*)abuf.p; A.started = A.iter + B; A.n_done = A.started + B;
    unsigned char* d_closes = (unsigned char*)(A.n_done + 4);
    if (n_tune > 0) CK(cudaMemcpyAsync(d_closes, closes.data(), n_tune, cudaMemcpyHostToDevice, st));
    A.closes = d_closes;
    A.draws = d_draws; A.stat_accept = d_acc; A.stat_logp = d_lp; A.stat_eps = d_eps;
    A.stat_depth = d_depth; A.stat_diverged = d_div;
    // every chain needs at most total * (2^max_depth + 1) ticks
    const int64_t tick_cap = (int64_t)total * ((1LL << max_depth) + 1) + 64;
    if (chain_kernel) {
      // one resident block per chain runs whole leapfrog steps tick after tick (nuts_chain.cuh); a launch is
      // bounded by CHI_CHAIN_TICKS ticks per chain, the host looks at the counter between launches
      int64_t ticks = 0;
      h_flags[0] = 0;
      while (h_flags[0] < B && ticks < tick_cap) {
        if (chain_cluster ? launch_nuts_chain_cl(chain_nl, h->cfg.lorentzian != 0, h->map, chain_args, s, A, CHI_CHAIN_TICKS, chain_dyn, st)
                          : launch_nuts_chain(chain_nl, chain_threads, h->cfg.lorentzian != 0, h->map, chain_args, s, A,
                                              CHI_CHAIN_TICKS, chain_dyn, st))
          return fail(h, CHI_EINVAL, "unsupported configuration of the chain kernel");
        CK(cudaGetLastError());
        ticks += CHI_CHAIN_TICKS;
        h->launches++;
        CK(cudaMemcpyAsync(h_flags, A.n_done, sizeof(int), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
      }
      if (h_flags[0] < B) return fail(h, CHI_ECUDA, "chi_nuts_sample: chains did not finish within the tick budget");
    } else {
    ScopedGraph tgs;
    cudaGraph_t& tg = tgs.g;
    cudaGraphExec_t& tgx = tgs.x;
    size_t tn = 0;
    // workspace of the tick is allocated before the capture starts
    CK(h->lanes[0].xcon.ensure((size_t)B * CHI_NSLOT * N * sizeof(double)));
    CK(h->lanes[0].cc.ensure((size_t)B * N * CHI_CC_STRIDE * sizeof(double)));
    CaptureGuard cap(h, st);
    CK(cap.begin());
    if (getenv("CHI_NUTS_THREAD")) {  // A/B: one thread per chain, separate map kernel
      k_nuts_tick<<<grid, 64, 0, st>>>(s, A);
      rc = evaluate();
    } else {  // one warp per chain; the tick also maps the new evaluation point
      Lane& L0 = h->lanes[0];
      double* tick_cc = use_eval_small(h, B, false) ? nullptr : L0.cc.d();
      if (h->opt[CHI_OPT_NUTS_TICK] == 1) {  // A/B and regression reference: every phase through global memory
        MODEL_SWITCH(h->cfg.model, (k_nuts_tick_w<M, false><<<(unsigned)((B + 3) / 4), 128, 0, st>>>(
                                       s, A, h->map, h->rm, tick_cc, L0.xcon.d())));
      } else {  // the common tick staged in shared memory (nuts_warp.cuh: nw_leaf_staged)
        MODEL_SWITCH(h->cfg.model, (k_nuts_tick_w<M, true><<<(unsigned)((B + 3) / 4), 128, 0, st>>>(
                                       s, A, h->map, h->rm, tick_cc, L0.xcon.d())));
      }
      CK(cudaEventRecord(L0.ev_fork, L0.st));  // the cloud constants are ready from here on
      h->call_first_stage = h->stage_calls;
      rc = run_backward(h, L0, B, s.eu, nbe ? s.ece : nullptr, nba ? s.eca : nullptr, d_sl, nullptr, nullptr, false,
                        s.elogp, s.egu, nbe ? s.egce : nullptr, nba ? s.egca : nullptr, true, PHASE_REST);
    }
    cudaError_t cap_err = cap.end(&tg);
    if (rc) return rc;
    CK(cap_err);
    CK(cudaGraphInstantiate(&tgx, tg, 0));
    CK(cudaGraphGetNodes(tg, nullptr, &tn));
    int64_t ticks = 0;
    h_flags[0] = 0;
    while (h_flags[0] < B && ticks < tick_cap) {
      for (int i = 0; i < 32; ++i) CK(cudaGraphLaunch(tgx, st));
      ticks += 32;
      h->launches += 32 * (int64_t)tn;
      CK(cudaMemcpyAsync(h_flags, A.n_done, sizeof(int), cudaMemcpyDeviceToHost, st));
      CK(cudaStreamSynchronize(st));
    }
    if (h_flags[0] < B) {
      return fail(h, CHI_ECUDA, "chi_nuts_sample: chains did not finish within the tick budget");
    }
    }  // graph of kernels per tick
  }
  ScopedGraph lgs[2];
  cudaGraphExec_t gexec[2] = {nullptr, nullptr};
  size_t n_nodes[2] = {0, 0};
  for (int g = 0; g < 2 && lock_step; ++g) {  // g = 0: first leaf of a doubling (no pending half step), 1: the others
    CaptureGuard cap(h, st);
    CK(cap.begin());
    k_nuts_leap<<<grid, 64, 0, st>>>(s, g, 1);
    rc = evaluate();
    cudaError_t cap_err = cap.end(&lgs[g].g);
    if (rc) return rc;
    CK(cap_err);
    CK(cudaGraphInstantiate(&lgs[g].x, lgs[g].g, 0));
    gexec[g] = lgs[g].x;
    CK(cudaGraphGetNodes(lgs[g].g, nullptr, &n_nodes[g]));
  }
  for (int it = 0; it < total && lock_step; ++it) {
    k_nuts_begin_iter<<<grid, 64, 0, st>>>(s);
    h->launches++;
    for (int depth = 0; depth < max_depth; ++depth) {
      CK(cudaMemsetAsync(s.n_stopped, 0, sizeof(int), st));
      k_nuts_begin_subtree<<<grid, 64, 0, st>>>(s);
      h->launches++;
      const int leaves = 1 << depth;
      for (int leaf = 0; leaf < leaves; ++leaf) {
        CK(cudaGraphLaunch(gexec[leaf ? 1 : 0], st));
        h->launches += (int64_t)n_nodes[leaf ? 1 : 0];
        if (depth >= 4 && (leaf & 7) == 7 && leaf + 1 < leaves) {  // every chain's sub-tree stopped early?
          CK(cudaMemcpyAsync(h_flags + 1, s.n_stopped, sizeof(int), cudaMemcpyDeviceToHost, st));
          CK(cudaStreamSynchronize(st));
          if (h_flags[1] >= B) break;
        }
      }
      k_nuts_leap<<<grid, 64, 0, st>>>(s, 1, 0);  // second half of the last leaf
      k_nuts_end_subtree<<<grid, 64, 0, st>>>(s, max_depth);
      h->launches += 2;
      CK(cudaMemcpyAsync(h_flags, s.n_active, sizeof(int), cudaMemcpyDeviceToHost, st));
      CK(cudaStreamSynchronize(st));
      if (h_flags[0] <= 0) break;
    }
    NutsAdapt a;
    memset(&a, 0, sizeof(a));
    a.target_accept = target;
    const bool tuning = it < n_tune;
    a.tuning = tuning;
    a.adapt_step = tuning;
    a.collect_mass = tuning && it >= init_buf && it < n_tune - term_buf;
    const bool close = tuning && closes[it];
    a.close_window = close && !pool_mass;
    // the averaged step size is adopted at the end of warm-up; with a pooled window that also closes
    // there (never the case for Stan's schedule: the terminal buffer follows the last window)
    a.finalize = tuning && it == n_tune - 1;
    const int64_t k = it - n_tune;
    k_nuts_end_iter<<<grid, 64, 0, st>>>(s, a, tuning ? nullptr : d_draws, d_acc, d_depth, d_div, d_lp, d_eps, tuning ? 0 : k);
    h->launches++;
    if (close && pool_mass) {
      k_nuts_pool_mass<<<(unsigned)D, 128, 0, st>>>(s);
      k_nuts_restart_window<<<grid, 64, 0, st>>>(s);
      h->launches += 2;
    }
    CK(cudaGetLastError());
  }
  // draws back to the host in the layout of the API blocks
  CK(cudaStreamSynchronize(st));
  std::vector<double> tmp((size_t)n_keep * B * D);
  CK(cudaMemcpy(tmp.data(), d_draws, tmp.size() * sizeof(double), cudaMemcpyDeviceToHost));
  for (int64_t i = 0; i < n_keep * B; ++i) {
    const double* src = tmp.data() + i * D;
    memcpy(draws_u + i * nu, src, nu * sizeof(double));
    // replicas of a per-draw scalar mirror cloud 0
    for (int d = 0; d < nu; ++d)
      if (role[d] == 2) draws_u[i * nu + d] = src[d - (d % N)];
    if (draws_ce && nbe) memcpy(draws_ce + i * nbe, src + nu, nbe * sizeof(double));
    if (draws_ca && nba) memcpy(draws_ca + i * nba, src + nu + nbe, nba * sizeof(double));
  }
  if (stat_accept) CK(cudaMemcpy(stat_accept, d_acc, (size_t)n_keep * B * sizeof(double), cudaMemcpyDeviceToHost));
  if (stat_logp) CK(cudaMemcpy(stat_logp, d_lp, (size_t)n_keep * B * sizeof(double), cudaMemcpyDeviceToHost));
  if (stat_step) CK(cudaMemcpy(stat_step, d_eps, (size_t)n_keep * B * sizeof(double), cudaMemcpyDeviceToHost));
  if (stat_depth) CK(cudaMemcpy(stat_depth, d_depth, (size_t)n_keep * B * sizeof(int), cudaMemcpyDeviceToHost));
  if (stat_diverged) CK(cudaMemcpy(stat_diverged, d_div, (size_t)n_keep * B * sizeof(int), cudaMemcpyDeviceToHost));
  return CHI_OK;
}

}  // extern "C"
