// NUTS with one resident BLOCK per chain: the whole leapfrog step -- the chain's state machine (nuts_warp.cuh:
// nw_tick, warp 0), the parameter map, the channel sums, the Jacobian and the contraction (small.cuh: sm_eval,
// all warps) -- runs inside one kernel, tick after tick, without a launch or a grid-wide dependency in between.
//
// Why: a few chains on a small model (BASELINE configs[0]: 8 chains x 2 clouds x 334+166 channels) are bound
// by the DEPENDENT latency of a tick, and the graph of two to four kernels per tick pays a launch boundary per
// kernel, a trip through global memory for the chunk records and their counter, and the wait for the slowest
// chain's blocks in every kernel.  Here the chains are independent of each other (each adapts its own step size
// and mass matrix, as PyMC's processes do), so a block never waits for another block: the evaluation point is
// handed over in shared memory and the block's record is the draw's record.
// (Keeping the chain's state vectors in shared memory for the duration of a launch was tried too -- the
// per-chain functions ran on rebased pointers -- and changed nothing: 4.04 against 4.05 us per tick for the
// state machine; its time is its own arithmetic and shuffles, not memory.  profiles/r2_nuts_chain.txt)
//
// The arithmetic of a tick is that of k_nuts_tick_w followed by k_eval_small with one block per draw; the draws
// are bit-identical to that pair of kernels (tests/test_gpu_nuts.py).  A launch runs at most `max_ticks` ticks
// per chain and returns; the host relaunches until every chain is done (all state lives in global memory, as
// with the graph), which bounds the run time of one launch.
#pragma once
#define CHI_DEVICE_ONLY
#include "nuts_warp.cuh"
#include "small.cuh"

namespace chi {

// NT = 256: two blocks per SM, any model; NT = 512: one block per SM with twice the channel warps (a tick of a
// 2-cloud model is 3 rounds of channel tiles instead of 5), at most 8 clouds.  The partition of the channels over
// the warps differs, so the two forms agree to rounding, not bit for bit.
template <int NL, bool PV, int NT>
__global__ void __launch_bounds__(NT, NT == CHI_SM_THREADS ? 2 : 1) k_nuts_chain(MapCfg cfg, SmallArgs a, NutsState s, NutsAsync A, int max_ticks) {
  __shared__ SmSmemT<NT> S;
  __shared__ double s_stage[5][CHI_NUTS_MAXD];
  __shared__ double s_point[CHI_NUTS_MAXD];  // the evaluation point: theta block [CHI_NSLOT][N], then the coefficients
  __shared__ int s_go;
  extern __shared__ double s_gb_dyn[];
  const int tid = threadIdx.x, lane = tid & 31;
  const int64_t b = blockIdx.x;
  if (A.iter[b] >= A.n_total) return;  // finished in an earlier launch
  const int nu = CHI_NSLOT * s.N;
  sm_init(S, cfg, a.rm);
  __syncthreads();
  long long c_tick = 0, c_eval = 0, n_tick = 0;  // development aid (CHI_OPT_SMALL_STAMPS): block 0's cycles per phase
  for (int t = 0; t < max_ticks; ++t) {
    const long long t0 = a.stamps ? clock64() : 0;
    if (tid < 32) {
      // (the model kind only selects the map of the evaluation point, which sm_eval does here: cc == null)
      nw_tick<CHI_MODEL_PHYSICS, true>(s, A, S.cfg, S.rm, nullptr, nullptr, b, lane, s_stage);
      __syncwarp();
      const bool go = A.iter[b] < A.n_total;
      // every lane reads back the dimensions it wrote itself
      for (int d = lane; d < s.D; d += 32) s_point[d] = nuts_ev(s.eu, s.ece, s.eca, s, b, d);
      if (lane == 0) s_go = go;
    }
    __syncthreads();
    const long long t1 = a.stamps ? clock64() : 0;
    if (!s_go) break;
    sm_eval<NL, PV, SM_CHAIN, NT>(S, s_gb_dyn, a, b, 0, s_point, s.nbe ? s_point + nu : nullptr,
                          s.nba ? s_point + nu + s.nbe : nullptr);
    __syncthreads();  // logp and the gradient of the point are in global memory for warp 0's next tick
    if (a.stamps) { c_tick += t1 - t0; c_eval += clock64() - t1; ++n_tick; }
  }
  if (a.stamps && b == 0 && tid == 0) {
    atomicAdd((unsigned long long*)a.stamps + 12, (unsigned long long)c_tick);
    atomicAdd((unsigned long long*)a.stamps + 13, (unsigned long long)c_eval);
    atomicAdd((unsigned long long*)a.stamps + 14, (unsigned long long)n_tick);
  }
}

// ---- the speculative form: the state machine NEXT TO the evaluation ---------------------------------------------
// In the common tick (a leaf inside a sub-tree) the next evaluation point depends on the gradient that has just
// arrived and on nothing else: x' = x + h M^-1 (p + h g); everything else the tick does for the finished leaf --
// energy, acceptance, multinomial choice, the U-turn checks against the checkpoints -- only decides whether the
// sub-tree goes on.  So a ninth warp runs the state machine: it computes and publishes the next point FIRST, the
// eight evaluating warps start on it at once, and the bookkeeping of the finished leaf runs next to the
// evaluation instead of in front of it (3.8 of a 14 us tick).  When the bookkeeping says the sub-tree has ended
// (a U-turn, a divergence: a handful of leaves per iteration) the evaluation in flight is of a point nobody
// wants: its result is dropped, the general path computes the real next point and that is evaluated.  Same
// arithmetic in the same order for everything that is kept: bit-identical draws.
#define CHI_SPEC_THREADS (CHI_SM_THREADS + 32)
#define CHI_FOR_D(d) for (int d = lane; d < s.D; d += 32)

__device__ __forceinline__ void chain_bar(int id) { asm volatile("bar.sync %0, %1;" ::"r"(id), "n"(CHI_SPEC_THREADS) : "memory"); }

// nw_leaf_staged with the drift moved to the front: `point` receives the next evaluation point and `go()` (called
// by every lane) releases the evaluating warps, both only if the sub-tree can go on (*released says so).  Returns
// what nw_leaf_staged returns.
// `cached`: the previous tick of this chain was such a leaf and the sub-tree went on, so the position of the leaf
// now finished (= the point just evaluated, still in `point`), its half-step momentum (wph), the inverse mass and
// the sub-tree's momentum sum are in shared memory from that tick: only the gradient comes from global memory.
// The leaf's scalars travel in registers from one such tick to the next (LeafRegs); `role` is the block's copy of
// s.role in shared memory.
// nuts_grad with the roles from shared memory
__device__ __forceinline__ double chain_grad(const NutsState& s, const int* role, int64_t b, int d) {
  double g = nuts_ev(s.egu, s.egce, s.egca, s, b, d);
  if (d < CHI_NSLOT * s.N && (d % s.N) == 0) {
    for (int n = 1; n < s.N; ++n)
      if (role[d + n] == 2) g += nuts_ev(s.egu, s.egce, s.egca, s, b, d + n);
  }
  return g;
}
struct LeafRegs {
  int leaf, depth;
  double h, energy0, lw_sub;
};
template <class Go>
__device__ __forceinline__ bool nw_leaf_spec(const NutsState& s, int64_t b, int lane, const NwStage& w, double* point,
                                             double* wph, const int* role, LeafRegs& c, const bool cached, bool* released,
                                             Go go) {
  if (!cached) {
    c.leaf = s.leaf[b];
    c.depth = s.depth[b];
    c.h = s.dir[b] * s.eps[b];
    c.energy0 = s.energy0[b];
    c.lw_sub = s.lw_sub[b];
  }
  const int leaf = c.leaf;
  const int depth = c.depth;
  const int64_t o = b * s.D;
  const double h = c.h;
  const double energy0 = c.energy0, lw_sub = c.lw_sub;
  double kin = 0.0;
  CHI_FOR_D(d) {
    const double x = cached ? point[d] : nuts_ev(s.eu, s.ece, s.eca, s, b, d);
    const double im = cached ? w.im[d] : s.inv_mass[o + d];
    double g = 0.0, p = 0.0;
    if (role[d] == 1) {
      g = chain_grad(s, role, b, d);
      p = (cached ? wph[d] : s.ph[o + d]) + 0.5 * h * g;
      kin += 0.5 * im * p * p;
    }
    const double rs = (cached ? w.rs[d] : s.rsum_s[o + d]) + p;
    w.x[d] = x; w.g[d] = g; w.p[d] = p; w.im[d] = im; w.rs[d] = rs;
    s.qc[o + d] = x;
    s.gc[o + d] = g;
    s.pc[o + d] = p;
    s.rsum_s[o + d] = rs;
  }
  *released = leaf + 1 < (1 << depth);
  if (*released) {
    // ---- first half of the next leapfrog step, ahead of the bookkeeping (nw_leaf_staged does it last) ----
    __syncwarp();
    CHI_FOR_D(d) {
      double x = w.x[d];
      if (role[d] == 1) {
        const double ph = w.p[d] + 0.5 * h * w.g[d];
        s.ph[o + d] = ph;
        wph[d] = ph;
        x += h * w.im[d] * ph;
      }
      point[d] = x;
    }
    __syncwarp();
    CHI_FOR_D(d) {
      if (role[d] == 2) point[d] = point[d - (d % s.N)];
    }
    __syncwarp();
    CHI_FOR_D(d) nuts_ev(s.eu, s.ece, s.eca, s, b, d) = point[d];
    go();
  }
  const double lp = s.elogp[b];
  kin = nw_sum(kin);
  double energy = -lp + kin;
  if (!(energy == energy) || isinf(energy)) energy = INFINITY;
  const double delta = energy - energy0;
  const bool diverging = !(delta <= 1000.0);
  const double lw_leaf = -delta;
  const double lw_new = nuts_logaddexp(lw_sub, lw_leaf);
  int take = 0;
  if (lane == 0) {
    unsigned long long st = s.rng[b];
    take = nuts_uniform(st) < exp(lw_leaf - lw_new);
    s.rng[b] = st;
  }
  take = __shfl_sync(0xffffffffu, take, 0);
  if (lane == 0) {
    s.leaf[b] = leaf + 1;
    s.sum_acc[b] += (delta == delta) ? fmin(1.0, exp(-delta)) : 0.0;
    s.n_leap[b] += 1;
    s.lw_sub[b] = lw_new;
    if (take) s.logp_s[b] = lp;
    if (diverging) s.diverged[b] = 1;
  }
  if (take) {
    CHI_FOR_D(d) {
      s.qs[o + d] = w.x[d];
      s.gs[o + d] = w.g[d];
    }
  }
  bool turning = false;
  const int idx_max = __popc((unsigned)leaf >> 1);
  double* ckr = s.ck_r + (b * CHI_NUTS_MAX_DEPTH) * s.D;
  double* ckrs = s.ck_rs + (b * CHI_NUTS_MAX_DEPTH) * s.D;
  if ((leaf & 1) == 0) {
    CHI_FOR_D(d) {
      ckr[idx_max * s.D + d] = w.p[d];
      ckrs[idx_max * s.D + d] = w.rs[d];
    }
  } else {
    const int n_sub = __ffs(~(unsigned)leaf) - 1;
    const int idx_min = idx_max - n_sub + 1;
    for (int i = idx_max; i >= idx_min && !turning; --i) {
      const double* ca = ckr + i * s.D;
      const double* cs = ckrs + i * s.D;
      double da = 0.0, db = 0.0;
      CHI_FOR_D(d) {
        if (role[d] != 1) continue;
        const double a = ca[d], bb = w.p[d];
        const double rsum = (w.rs[d] - cs[d] + a) - 0.5 * (a + bb);
        da += w.im[d] * a * rsum;
        db += w.im[d] * bb * rsum;
      }
      da = nw_sum(da);
      db = nw_sum(db);
      turning = da <= 0.0 || db <= 0.0;
    }
  }
  int stop = 0;
  if (turning || diverging) stop = turning ? 2 : 3;
  if (lane == 0 && stop) s.sub_stop[b] = stop;
  __syncwarp();
  c.leaf = leaf + 1;
  c.lw_sub = lw_new;
  return !(stop != 0 || leaf + 1 >= (1 << depth));
}
#undef CHI_FOR_D

// barriers between the state-machine warp and the evaluating warps (all CHI_SPEC_THREADS threads):
//   3 "a point is published (s_go = 1) or the kernel ends (s_go = 0)",  4 "its evaluation is complete"
template <int NL, bool PV>
__global__ void __launch_bounds__(CHI_SPEC_THREADS, 1) k_nuts_chain_spec(MapCfg cfg, SmallArgs a, NutsState s, NutsAsync A, int max_ticks) {
  __shared__ SmSmem S;
  __shared__ double s_stage[6][CHI_NUTS_MAXD];
  __shared__ double s_point[CHI_NUTS_MAXD];
  __shared__ int s_go;
  __shared__ int s_role[CHI_NUTS_MAXD];
  extern __shared__ double s_gb_dyn[];
  const int tid = threadIdx.x, lane = tid & 31;
  const int64_t b = blockIdx.x;
  if (A.iter[b] >= A.n_total) return;  // finished in an earlier launch
  const int nu = CHI_NSLOT * s.N;
  if (tid < CHI_SM_THREADS) sm_init(S, cfg, a.rm);
  for (int d = tid; d < s.D; d += CHI_SPEC_THREADS) s_role[d] = s.role[d];
  __syncthreads();
  if (tid < CHI_SM_THREADS) {
    // ---- the evaluating warps ----
    for (;;) {
      chain_bar(3);
      if (!s_go) break;
      sm_eval<NL, PV, SM_CHAIN, CHI_SM_THREADS, true>(S, s_gb_dyn, a, b, 0, s_point, s.nbe ? s_point + nu : nullptr,
                                                      s.nba ? s_point + nu + s.nbe : nullptr);
      sm_sync<true, CHI_SM_THREADS>();  // logp and gradient are written
      chain_bar(4);
    }
    return;
  }
  // ---- the state-machine warp ----
  const NwStage w = {s_stage[0], s_stage[1], s_stage[2], s_stage[3], s_stage[4]};
  auto publish = [&]() {  // the evaluation point as the general path left it in global memory
    for (int d = lane; d < s.D; d += 32) s_point[d] = nuts_ev(s.eu, s.ece, s.eca, s, b, d);
    if (lane == 0) s_go = 1;
    __syncwarp();
    chain_bar(3);
  };
  bool done = false;
  LeafRegs regs = {0, 0, 0.0, 0.0, 0.0};
  for (int t = 0; t < max_ticks; ++t) {
    // here the logp / gradient of the point last published are in global memory
    const int started = A.started[b];
    __syncwarp();
    const bool cached = done;  // the previous tick was a leaf and its sub-tree went on
    done = false;
    if (started) {
      bool released = false;
      done = nw_leaf_spec(s, b, lane, w, s_point, s_stage[5], s_role, regs, cached, &released, [&]() {
        if (lane == 0) s_go = 1;
        __syncwarp();
        chain_bar(3);
      });
      if (!done) {
        nw_tick_rest<CHI_MODEL_PHYSICS>(s, A, S.cfg, S.rm, nullptr, nullptr, b, lane, 1, true);
        if (released) chain_bar(4);  // the evaluation in flight is of a point nobody wants: let it finish, drop it
      }
    } else {
      nw_tick_rest<CHI_MODEL_PHYSICS>(s, A, S.cfg, S.rm, nullptr, nullptr, b, lane, 0, false);
    }
    __syncwarp();
    if (!done) {
      if (A.iter[b] >= A.n_total) break;  // the chain has finished
      publish();
    }
    chain_bar(4);  // the evaluation of the published point is complete
  }
  if (lane == 0) s_go = 0;
  __syncwarp();
  chain_bar(3);
}

// The same with the n_chunks blocks of a chain as one thread-block CLUSTER (small.cuh, SM_CLUSTER), for models
// whose channel sums want more than one SM per chain (configs[2]: 8 clouds x 2048+2048 channels): the cluster's
// first block runs the chain's state machine and publishes the evaluation point in its shared memory; after a
// cluster barrier every block copies the point through distributed shared memory and evaluates its chunk of the
// channels; records and Jacobian records meet in the first block (sm_eval), which contracts them and goes on
// to the next tick.  Three cluster barriers per tick, no launch, no trip through global memory.  Bit-identical
// to the graph of {staged tick, k_eval_small_cl with the same blocks per draw}.
template <int NL, bool PV>
__global__ void __launch_bounds__(CHI_SM_THREADS, 2) k_nuts_chain_cl(MapCfg cfg, SmallArgs a, NutsState s, NutsAsync A, int max_ticks,
                                                                     unsigned x_off) {
  namespace cg = cooperative_groups;
  __shared__ SmSmem S;
  __shared__ double s_stage[5][CHI_NUTS_MAXD];
  __shared__ double s_point[CHI_NUTS_MAXD + 1];  // the evaluation point, then 1.0 / 0.0: evaluate it / the chain is done
  extern __shared__ double s_gb_dyn[];
  cg::cluster_group cluster = cg::this_cluster();
  const int tid = threadIdx.x, lane = tid & 31;
  const int64_t b = blockIdx.x / a.n_chunks;
  const int rank = (int)(blockIdx.x - b * a.n_chunks);  // == cluster.block_rank()
  if (A.iter[b] >= A.n_total) return;  // finished in an earlier launch (the same answer in every block of the cluster)
  const int nu = CHI_NSLOT * s.N, D = s.D;
  double* const xbuf = reinterpret_cast<double*>(reinterpret_cast<char*>(s_gb_dyn) + x_off);
  const double* const lead_point = cluster.map_shared_rank(s_point, 0);
  sm_init(S, cfg, a.rm);
  __syncthreads();
  for (int t = 0; t < max_ticks; ++t) {
    if (rank == 0 && tid < 32) {
      nw_tick<CHI_MODEL_PHYSICS, true>(s, A, S.cfg, S.rm, nullptr, nullptr, b, lane, s_stage);
      __syncwarp();
      const bool go = A.iter[b] < A.n_total;
      for (int d = lane; d < D; d += 32) s_point[d] = nuts_ev(s.eu, s.ece, s.eca, s, b, d);
      if (lane == 0) s_point[D] = go ? 1.0 : 0.0;
    }
    cluster.sync();  // the point is published; also: every block is done with the previous tick
    if (rank != 0)
      for (int d = tid; d <= D; d += CHI_SM_THREADS) s_point[d] = lead_point[d];
    __syncthreads();
    if (s_point[D] == 0.0) {
      cluster.sync();  // nobody leaves while its shared memory may still be read
      break;
    }
    sm_eval<NL, PV, SM_CLUSTER>(S, s_gb_dyn, a, b, rank, s_point, s.nbe ? s_point + nu : nullptr,
                                s.nba ? s_point + nu + s.nbe : nullptr, xbuf);
    __syncthreads();
  }
}

}  // namespace chi
