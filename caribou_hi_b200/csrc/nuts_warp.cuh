// Asynchronous chains, one WARP per chain (nuts.cuh holds the algorithm and the thread-per-chain
// kernels of the lock-step mode; this file is the same per-chain logic with the loops over the D
// dimensions of the position vector spread over the 32 lanes).
//
// Why: in the asynchronous scheme some chain finishes a sub-tree / an iteration in almost every
// tick, so the "heavy" path (merge the sub-tree, adapt, record the draw, draw a fresh momentum,
// start the next doubling) sits on the critical path of every tick; with one thread per chain
// that path is ~1000 dependent global-memory operations (~25 us), with one lane per dimension it is
// a few dozen.
//
// Conventions: every lane of the warp calls every function; per-chain scalars are read by all
// lanes and written by lane 0, with a __syncwarp() between a write and the next read; random
// numbers of the chain come from lane 0's stream and are broadcast, the momentum of dimension d
// from a stream derived from (chain stream, d).
#pragma once
#include "kernels.cuh"
#include "nuts.cuh"

namespace chi {

#define CHI_FOR_D(d) for (int d = lane; d < s.D; d += 32)

__device__ __forceinline__ double nw_sum(double x) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
  return x;
}
__device__ __forceinline__ int nw_any(bool p) { return __any_sync(0xffffffffu, p); }

// generalised U-turn criterion; r_a, r_b, r_sum are given per dimension by the functors
template <class FA, class FB, class FS>
__device__ __forceinline__ bool nw_turning(const NutsState& s, int64_t o, int lane, FA r_a, FB r_b, FS r_sum) {
  double da = 0.0, db = 0.0;
  CHI_FOR_D(d) {
    if (s.role[d] != 1) continue;
    const double a = r_a(d), b = r_b(d);
    const double rs = r_sum(d) - 0.5 * (a + b);
    da += s.inv_mass[o + d] * a * rs;
    db += s.inv_mass[o + d] * b * rs;
  }
  da = nw_sum(da);
  db = nw_sum(db);
  return da <= 0.0 || db <= 0.0;
}

// evaluation point <- position vector x (global, offset o); replicas mirror cloud 0
__device__ __forceinline__ void nw_set_eval(const NutsState& s, int64_t b, int lane, const double* x) {
  const int64_t o = b * s.D;
  CHI_FOR_D(d) {
    const int src = (s.role[d] == 2) ? d - (d % s.N) : d;
    nuts_ev(s.eu, s.ece, s.eca, s, b, d) = x[o + src];
  }
}

__device__ __forceinline__ void nw_begin_iter(const NutsState& s, int64_t b, int lane) {
  const int64_t o = b * s.D;
  const unsigned long long st0 = s.rng[b];
  double kin = 0.0;
  CHI_FOR_D(d) {
    double p = 0.0;
    if (s.role[d] == 1) {
      unsigned long long sd = st0 ^ (0x9E3779B97F4A7C15ULL * (unsigned long long)(d + 1));
      nuts_next(sd);
      p = nuts_normal(sd) / sqrt(s.inv_mass[o + d]);
      kin += 0.5 * s.inv_mass[o + d] * p * p;
    }
    s.pl[o + d] = s.pr[o + d] = s.rsum[o + d] = p;
    s.ql[o + d] = s.qr[o + d] = s.qp[o + d] = s.q[o + d];
    s.gl[o + d] = s.gr[o + d] = s.gp[o + d] = s.g[o + d];
  }
  kin = nw_sum(kin);
  if (lane == 0) {
    unsigned long long st = st0;
    nuts_next(st);  // the chain's stream moves on by one step per momentum refresh
    s.rng[b] = st;
    s.logp_p[b] = s.logp[b];
    s.energy0[b] = -s.logp[b] + kin;
    s.lw_tree[b] = 0.0;
    s.sum_acc[b] = 0.0;
    s.n_leap[b] = 0;
    s.depth[b] = 0;
    s.diverged[b] = 0;
    s.active[b] = 1;
  }
  __syncwarp();
}

__device__ __forceinline__ void nw_begin_subtree(const NutsState& s, int64_t b, int lane) {
  const int64_t o = b * s.D;
  int dir = 0;
  if (lane == 0) {
    unsigned long long st = s.rng[b];
    dir = (nuts_next(st) & 1ULL) ? 1 : -1;
    s.rng[b] = st;
    s.dir[b] = dir;
    s.sub_stop[b] = 0;
    s.lw_sub[b] = -INFINITY;
    s.leaf[b] = 0;
  }
  dir = __shfl_sync(0xffffffffu, dir, 0);
  const double *eq = dir > 0 ? s.qr : s.ql, *ep = dir > 0 ? s.pr : s.pl, *eg = dir > 0 ? s.gr : s.gl;
  CHI_FOR_D(d) {
    s.qc[o + d] = eq[o + d];
    s.pc[o + d] = ep[o + d];
    s.gc[o + d] = eg[o + d];
    s.rsum_s[o + d] = 0.0;
  }
  __syncwarp();
}

// leapfrog, first half: half kick + drift; the new position becomes the evaluation point
__device__ __forceinline__ void nw_leap_a(const NutsState& s, int64_t b, int lane) {
  const int64_t o = b * s.D;
  const double h = s.dir[b] * s.eps[b];
  CHI_FOR_D(d) {
    double x = s.qc[o + d];
    if (s.role[d] == 1) {
      const double ph = s.pc[o + d] + 0.5 * h * s.gc[o + d];
      s.ph[o + d] = ph;
      x += h * s.inv_mass[o + d] * ph;
    }
    if (s.role[d] != 2) nuts_ev(s.eu, s.ece, s.eca, s, b, d) = x;
  }
  __syncwarp();
  CHI_FOR_D(d) {
    if (s.role[d] == 2) nuts_ev(s.eu, s.ece, s.eca, s, b, d) = nuts_ev(s.eu, s.ece, s.eca, s, b, d - (d % s.N));
  }
  __syncwarp();
}

// leapfrog, second half + sub-tree bookkeeping of the chain's next leaf; returns true when the
// sub-tree is finished (all its leaves done, a U-turn inside it, or a divergence)
__device__ __forceinline__ bool nw_leap_b(const NutsState& s, int64_t b, int lane) {
  const int leaf = s.leaf[b];
  const int depth = s.depth[b];
  const int64_t o = b * s.D;
  const double h = s.dir[b] * s.eps[b];
  const double lp = s.elogp[b];
  double kin = 0.0;
  CHI_FOR_D(d) {
    s.qc[o + d] = nuts_ev(s.eu, s.ece, s.eca, s, b, d);
    double g = 0.0, p = 0.0;
    if (s.role[d] == 1) {
      g = nuts_grad(s, b, d);
      p = s.ph[o + d] + 0.5 * h * g;
      kin += 0.5 * s.inv_mass[o + d] * p * p;
    }
    s.gc[o + d] = g;
    s.pc[o + d] = p;
  }
  kin = nw_sum(kin);
  double energy = -lp + kin;
  if (!(energy == energy) || isinf(energy)) energy = INFINITY;
  const double delta = energy - s.energy0[b];
  const bool diverging = !(delta <= 1000.0);
  const double lw_leaf = -delta;
  const double lw_new = nuts_logaddexp(s.lw_sub[b], lw_leaf);
  int take = 0;
  if (lane == 0) {
    unsigned long long st = s.rng[b];
    take = nuts_uniform(st) < exp(lw_leaf - lw_new);
    s.rng[b] = st;
  }
  take = __shfl_sync(0xffffffffu, take, 0);
  __syncwarp();  // every lane has read lw_sub / leaf / energy0 before lane 0 updates the scalars
  if (lane == 0) {
    s.leaf[b] = leaf + 1;
    s.sum_acc[b] += (delta == delta) ? fmin(1.0, exp(-delta)) : 0.0;
    s.n_leap[b] += 1;
    s.lw_sub[b] = lw_new;
    if (take) s.logp_s[b] = lp;
    if (diverging) s.diverged[b] = 1;
  }
  CHI_FOR_D(d) {
    s.rsum_s[o + d] += s.pc[o + d];
    if (take) {
      s.qs[o + d] = s.qc[o + d];
      s.gs[o + d] = s.gc[o + d];
    }
  }
  // iterative U-turn test through the checkpoints (Phan et al. 2019); a lane only ever reads the
  // checkpoint entries it wrote itself (same d), so no warp barrier is needed around them
  bool turning = false;
  const int idx_max = __popc((unsigned)leaf >> 1);
  double* ckr = s.ck_r + (b * CHI_NUTS_MAX_DEPTH) * s.D;
  double* ckrs = s.ck_rs + (b * CHI_NUTS_MAX_DEPTH) * s.D;
  if ((leaf & 1) == 0) {
    CHI_FOR_D(d) {
      ckr[idx_max * s.D + d] = s.pc[o + d];
      ckrs[idx_max * s.D + d] = s.rsum_s[o + d];
    }
  } else {
    const int n_sub = __ffs(~(unsigned)leaf) - 1;  // trailing ones of leaf
    const int idx_min = idx_max - n_sub + 1;
    for (int i = idx_max; i >= idx_min && !turning; --i) {
      const double* ca = ckr + i * s.D;
      const double* cs = ckrs + i * s.D;
      turning = nw_turning(
          s, o, lane, [&](int d) { return ca[d]; }, [&](int d) { return s.pc[o + d]; },
          [&](int d) { return s.rsum_s[o + d] - cs[d] + ca[d]; });
    }
  }
  int stop = 0;
  if (turning || diverging) stop = turning ? 2 : 3;  // 2: sub-tree turned, 3: diverged
  if (lane == 0 && stop) s.sub_stop[b] = stop;
  __syncwarp();
  return stop != 0 || leaf + 1 >= (1 << depth);
}

// end of a doubling: merge the sub-tree into the tree; returns true while the tree keeps growing
__device__ __forceinline__ bool nw_end_subtree(const NutsState& s, int64_t b, int lane, int max_depth) {
  const int64_t o = b * s.D;
  const int dir = s.dir[b];
  const bool sub_bad = s.sub_stop[b] >= 2;
  const double lw_sub = s.lw_sub[b], lw_tree = s.lw_tree[b];
  const int depth = s.depth[b] + 1;
  double *eq = dir > 0 ? s.qr : s.ql, *ep = dir > 0 ? s.pr : s.pl, *eg = dir > 0 ? s.gr : s.gl;
  int take = 0;
  if (lane == 0) {
    unsigned long long st = s.rng[b];
    // biased progressive sampling between the old tree and the new sub-tree
    take = !sub_bad && (nuts_uniform(st) < fmin(1.0, exp(lw_sub - lw_tree)));
    s.rng[b] = st;
  }
  take = __shfl_sync(0xffffffffu, take, 0);
  CHI_FOR_D(d) {
    eq[o + d] = s.qc[o + d];
    ep[o + d] = s.pc[o + d];
    eg[o + d] = s.gc[o + d];
    s.rsum[o + d] += s.rsum_s[o + d];
    if (take) {
      s.qp[o + d] = s.qs[o + d];
      s.gp[o + d] = s.gs[o + d];
    }
  }
  __syncwarp();  // the edge that moved (pl or pr) is read across lanes only through own d: still, order the scalars
  const bool turning =
      sub_bad || nw_turning(
                     s, o, lane, [&](int d) { return s.pl[o + d]; }, [&](int d) { return s.pr[o + d]; },
                     [&](int d) { return s.rsum[o + d]; });
  const bool grow = !(turning || depth >= max_depth);
  if (lane == 0) {
    if (take) s.logp_p[b] = s.logp_s[b];
    s.lw_tree[b] = nuts_logaddexp(lw_tree, lw_sub);
    s.depth[b] = depth;
    s.active[b] = grow ? 1 : 0;
  }
  __syncwarp();
  return grow;
}

// end of the iteration: move to the proposal, adapt, record
__device__ __forceinline__ void nw_end_iter(const NutsState& s, const NutsAdapt& a, int64_t b, int lane, double* draws,
                                            double* stat_accept, int* stat_depth, int* stat_diverged,
                                            double* stat_logp, double* stat_eps, int64_t draw_index) {
  const int64_t o = b * s.D;
  const double acc = s.n_leap[b] > 0 ? s.sum_acc[b] / s.n_leap[b] : 0.0;
  const int wc = s.w_count[b];
  const int c_mass = (a.tuning && a.collect_mass) ? wc + 1 : wc;
  const bool close = a.tuning && a.close_window && c_mass > 1;
  CHI_FOR_D(d) {
    const double x = s.qp[o + d];
    s.q[o + d] = x;
    s.g[o + d] = s.gp[o + d];
    if (a.tuning && s.role[d] == 1) {
      double mean = s.w_mean[o + d], m2 = s.w_m2[o + d];
      if (a.collect_mass) {
        const double dm = x - mean;
        mean += dm / c_mass;
        m2 += dm * (x - mean);
      }
      if (close) {
        const double n = c_mass, var = m2 / (n - 1.0);
        s.inv_mass[o + d] = (n / (n + 5.0)) * var + 1e-3 * (5.0 / (n + 5.0));  // Stan's regularisation
        mean = 0.0;
        m2 = 0.0;
      }
      s.w_mean[o + d] = mean;
      s.w_m2[o + d] = m2;
    }
    if (draws) draws[(draw_index * s.B + b) * s.D + d] = x;
  }
  if (lane == 0) {
    const double lp = s.logp_p[b];
    s.logp[b] = lp;
    double eps = s.eps[b];
    if (a.tuning) {
      if (a.adapt_step) {  // Nesterov dual averaging (Hoffman & Gelman 2014), gamma 0.05, t0 10, kappa 0.75
        const int c = ++s.da_count[b];
        const double w = 1.0 / (c + 10.0);
        s.da_hbar[b] = (1.0 - w) * s.da_hbar[b] + w * (a.target_accept - acc);
        const double log_eps = s.da_mu[b] - sqrt((double)c) / 0.05 * s.da_hbar[b];
        const double eta = pow((double)c, -0.75);
        s.da_logeps_bar[b] = eta * log_eps + (1.0 - eta) * s.da_logeps_bar[b];
        eps = exp(log_eps);
      }
      s.w_count[b] = close ? 0 : c_mass;
      if (close) {
        s.da_mu[b] = log(10.0 * eps);
        s.da_hbar[b] = 0.0;
        s.da_logeps_bar[b] = 0.0;
        s.da_count[b] = 0;
      }
      if (a.finalize) eps = exp(s.da_logeps_bar[b]);
      s.eps[b] = eps;
    }
    if (draws) {
      stat_accept[draw_index * s.B + b] = acc;
      stat_depth[draw_index * s.B + b] = s.depth[b];
      stat_diverged[draw_index * s.B + b] = s.diverged[b];
      stat_logp[draw_index * s.B + b] = lp;
      stat_eps[draw_index * s.B + b] = eps;
    }
  }
  __syncwarp();
}

// ---- the common tick, staged in shared memory ------------------------------------------------------
// Most ticks of a chain are "a leaf inside a sub-tree": finish the leapfrog step whose gradient has
// arrived (nw_leap_b), and since the sub-tree goes on, start the next one (nw_leap_a) and map the new
// point.  Written with the functions above that is a chain of five store -> load round trips through
// global memory (qc, pc, gc written by one phase and read back by the next; the evaluation point
// written and read back by the map), ~1 us each for a kernel whose whole job is latency.  Here the
// chain's working vectors live in shared memory for the tick: everything the later phases need is
// read from there, the global copies are still written (later ticks and the rare paths read them) but
// nothing waits for them.  Same operations in the same order as nw_leap_b + nw_leap_a + the map: the
// draws are bit-identical (tests/test_gpu_nuts.py::test_tick_variants_agree).
// Returns true when the tick is complete; false when the sub-tree ended with this leaf (all global
// state is as nw_leap_b leaves it) and the caller continues with the general path.
#define CHI_NUTS_MAXD (CHI_NSLOT * CHI_MAX_CLOUDS + 2 * CHI_MAX_BASELINE)
struct NwStage {
  double *x, *g, *p, *im, *rs;  // [D]: position, gradient, momentum of the finished leaf; inverse mass; sub-tree momentum sum
};

template <int MODEL>
__device__ __forceinline__ bool nw_leaf_staged(const NutsState& s, int64_t b, int lane, const NwStage& w, const MapCfg& cfg,
                                               const RowMap& rm, double* __restrict__ cc, double* __restrict__ xcon) {
  const int leaf = s.leaf[b];
  const int depth = s.depth[b];
  const int64_t o = b * s.D;
  const double h = s.dir[b] * s.eps[b];
  const double lp = s.elogp[b];
  const double energy0 = s.energy0[b], lw_sub = s.lw_sub[b];
  double kin = 0.0;
  CHI_FOR_D(d) {
    const double x = nuts_ev(s.eu, s.ece, s.eca, s, b, d);
    const double im = s.inv_mass[o + d];
    double g = 0.0, p = 0.0;
    if (s.role[d] == 1) {
      g = nuts_grad(s, b, d);
      p = s.ph[o + d] + 0.5 * h * g;
      kin += 0.5 * im * p * p;
    }
    const double rs = s.rsum_s[o + d] + p;
    w.x[d] = x; w.g[d] = g; w.p[d] = p; w.im[d] = im; w.rs[d] = rs;
    s.qc[o + d] = x;
    s.gc[o + d] = g;
    s.pc[o + d] = p;
    s.rsum_s[o + d] = rs;
  }
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
  // iterative U-turn test through the checkpoints; a lane only touches the entries of its own dimensions
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
      // (same arithmetic as nw_turning, with the inverse mass, momentum and momentum sum from shared memory)
      double da = 0.0, db = 0.0;
      CHI_FOR_D(d) {
        if (s.role[d] != 1) continue;
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
  if (stop != 0 || leaf + 1 >= (1 << depth)) {
    __syncwarp();  // the general path reads what lane 0 and the other lanes wrote
    return false;
  }
  // ---- the sub-tree goes on: first half of the next leapfrog step, straight from shared memory ----
  const int nu = CHI_NSLOT * s.N;
  CHI_FOR_D(d) {
    double x = w.x[d];
    if (s.role[d] == 1) {
      const double ph = w.p[d] + 0.5 * h * w.g[d];
      s.ph[o + d] = ph;
      x += h * w.im[d] * ph;
    }
    w.x[d] = x;
  }
  __syncwarp();
  CHI_FOR_D(d) {
    const double x = (s.role[d] == 2) ? w.x[d - (d % s.N)] : w.x[d];
    w.g[d] = x;  // the evaluation point, replicas resolved (w.g is free now)
    nuts_ev(s.eu, s.ece, s.eca, s, b, d) = x;
  }
  __syncwarp();
  (void)nu;
  if (cc && lane < s.N)
    cloud_map_draw<MODEL>(cfg, rm, s.N, w.g, cc + (b * s.N) * CHI_CC_STRIDE,
                          xcon ? xcon + (b * CHI_NSLOT) * s.N : nullptr, lane);
  return true;
}

// the rest of a tick once the leaf in flight is accounted for (`finished`: the sub-tree ended with it) or the
// chain has not started yet: merge the sub-tree, maybe end the iteration (adapt, record, next momentum), start
// the next sub-tree, and the first half of the next leapfrog step
template <int MODEL>
__device__ __forceinline__ void nw_tick_rest(const NutsState& s, const NutsAsync& A, const MapCfg& cfg, const RowMap& rm,
                                             double* __restrict__ cc, double* __restrict__ xcon, const int64_t b, const int lane,
                                             const int started, const bool finished) {
  int it = A.iter[b];
  if (started) {
    if (finished) {  // the logp / gradient of the leaf in flight has arrived
      if (!nw_end_subtree(s, b, lane, A.max_depth)) {
        NutsAdapt a;
        a.target_accept = A.target_accept;
        a.tuning = it < A.n_tune;
        a.adapt_step = a.tuning;
        a.collect_mass = a.tuning && it >= A.init_buf && it < A.n_tune - A.term_buf;
        a.close_window = a.tuning && A.closes[it];
        a.finalize = a.tuning && it == A.n_tune - 1;
        nw_end_iter(s, a, b, lane, a.tuning ? nullptr : A.draws, A.stat_accept, A.stat_depth, A.stat_diverged,
                    A.stat_logp, A.stat_eps, a.tuning ? 0 : it - A.n_tune);
        ++it;
        if (lane == 0) A.iter[b] = it;
        if (it >= A.n_total) {
          nw_set_eval(s, b, lane, s.q);
          if (lane == 0) atomicAdd(A.n_done, 1);
          return;
        }
        nw_begin_iter(s, b, lane);
      }
      nw_begin_subtree(s, b, lane);
    }
  } else {
    if (lane == 0) A.started[b] = 1;
    nw_begin_iter(s, b, lane);
    nw_begin_subtree(s, b, lane);
  }
  nw_leap_a(s, b, lane);
  if (cc && lane < s.N) cloud_map_one<MODEL>(cfg, rm, s.N, s.eu, cc, xcon, b, lane);
}

// one tick of chain b's state machine (see NutsAsync in nuts.cuh), by one warp.  `st`: the warp's
// [5][CHI_NUTS_MAXD] staging area in shared memory (STAGED).  The cloud map of the new evaluation point
// (k_cloud_map's work: transforms, parameter map, profile constants) is done here by lanes 0..N-1 when cc is
// given, which takes one kernel out of the dependent chain of a tick; cc == null: the evaluation maps the
// point itself (k_eval_small, k_nuts_chain).
template <int MODEL, bool STAGED>
__device__ __forceinline__ void nw_tick(const NutsState& s, const NutsAsync& A, const MapCfg& cfg, const RowMap& rm,
                                        double* __restrict__ cc, double* __restrict__ xcon, const int64_t b, const int lane,
                                        double (*st)[STAGED ? CHI_NUTS_MAXD : 1]) {
  int it = A.iter[b];
  if (it >= A.n_total) return;  // finished: its evaluation point stays at the last state
  const int started = A.started[b];
  __syncwarp();
  if (started && STAGED) {
    const NwStage w = {st[0], st[1], st[2], st[3], st[4]};
    if (nw_leaf_staged<MODEL>(s, b, lane, w, cfg, rm, cc, xcon)) return;  // the common tick: done
  }
  nw_tick_rest<MODEL>(s, A, cfg, rm, cc, xcon, b, lane, started, started ? (STAGED ? true : nw_leap_b(s, b, lane)) : false);
}

// one tick of every chain; 4 chains per block
template <int MODEL, bool STAGED>
__global__ void __launch_bounds__(128) k_nuts_tick_w(NutsState s, NutsAsync A, MapCfg cfg, RowMap rm,
                                                     double* __restrict__ cc, double* __restrict__ xcon) {
  __shared__ double s_stage[STAGED ? 4 : 1][5][STAGED ? CHI_NUTS_MAXD : 1];
  const int lane = threadIdx.x & 31;
  const int64_t b = (int64_t)blockIdx.x * 4 + (threadIdx.x >> 5);
  if (b >= s.B) return;
  nw_tick<MODEL, STAGED>(s, A, cfg, rm, cc, xcon, b, lane, s_stage[STAGED ? (threadIdx.x >> 5) : 0]);
}

#undef CHI_FOR_D

}  // namespace chi
