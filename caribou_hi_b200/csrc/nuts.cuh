// Chain-batched No-U-Turn sampler state machine.
//
// The reference samples with PyMC's NUTS, one CPU process per chain, one logp+grad call per
// leapfrog step (optimization.ipynb cell 12).  Here all chains advance in lock step on the
// device: one batched chi_model_logp_grad_dev evaluation per leapfrog step serves every chain,
// and the per-chain tree logic runs in the small kernels below (one thread per chain), so no
// parameter or gradient ever crosses the PCIe bus during sampling.
//
// Algorithm: multinomial NUTS with the iterative (checkpointed) sub-tree U-turn test --
// the scheme of Hoffman & Gelman 2014 as refined by Betancourt 2017 and implemented
// iteratively by Phan et al. 2019 (NumPyro): tree doubling in a random direction, uniform
// sampling inside a sub-tree, biased progressive sampling between sub-trees, generalised
// U-turn criterion on momentum sums, divergence when the energy error exceeds 1000; dual
// averaging of the step size and Stan-style windowed diagonal mass adaptation during tuning.
//
// Position vector of a chain: the unconstrained parameter block u[10][N] followed by the baseline
// coefficients (emission, absorption): D = 10 N + nb_e + nb_a.  role[d]: 0 = not a parameter of
// the model (frozen), 1 = free, 2 = replica of the same slot's cloud 0 (the physical models'
// fwhm_L_norm is one RV per draw replicated over clouds; its gradient is the sum over clouds).
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

namespace chi {

#define CHI_NUTS_MAX_DEPTH 12

struct NutsState {
  int64_t B;
  int N, D, nbe, nba;
  const int* role;  // [D]
  // evaluation point and its results, in the layout of chi_model_logp_grad_dev
  double *eu, *ece, *eca;      // [B][10N], [B][nbe], [B][nba]
  double *egu, *egce, *egca;   // gradients
  double* elogp;               // [B]
  // chain vectors [B][D]
  double *q, *g;               // current state
  double *ql, *pl, *gl;        // left edge of the tree
  double *qr, *pr, *gr;        // right edge
  double *qp, *gp;             // tree proposal
  double *qs, *gs;             // sub-tree proposal
  double *qc, *pc, *gc;        // cursor: the moving edge while a sub-tree is built
  double *ph;                  // half-step momentum of the leaf in flight
  double *rsum, *rsum_s;       // momentum sums of tree / sub-tree
  double *ck_r, *ck_rs;        // [B][MAX_DEPTH][D] sub-tree checkpoints
  double *inv_mass;            // [B][D]
  double *w_mean, *w_m2;       // Welford accumulators of the mass adaptation
  // chain scalars [B]
  double *logp, *logp_p, *logp_s, *energy0, *lw_tree, *lw_sub, *sum_acc, *eps;
  double *da_mu, *da_hbar, *da_logeps_bar;
  int *n_leap, *dir, *active, *sub_stop, *diverged, *depth, *da_count, *w_count;
  int* leaf;                   // [B] index of the next leaf of the sub-tree being built
  unsigned long long* rng;
  int* n_active;               // [1] chains whose tree is still growing
  int* n_stopped;              // [1] chains idle in the current doubling (inactive or sub-tree stopped)
};

__device__ __forceinline__ double& nuts_ev(double* u, double* ce, double* ca, const NutsState& s, int64_t b, int d) {
  const int nu = CHI_NSLOT * s.N;
  if (d < nu) return u[b * nu + d];
  d -= nu;
  if (d < s.nbe) return ce[b * s.nbe + d];
  return ca[b * s.nba + (d - s.nbe)];
}

__device__ __forceinline__ unsigned long long nuts_next(unsigned long long& st) {  // splitmix64
  unsigned long long z = (st += 0x9E3779B97F4A7C15ULL);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  return z ^ (z >> 31);
}
__device__ __forceinline__ double nuts_uniform(unsigned long long& st) {  // (0, 1)
  return ((nuts_next(st) >> 11) + 0.5) * (1.0 / 9007199254740992.0);
}
__device__ __forceinline__ double nuts_normal(unsigned long long& st) {
  const double u1 = nuts_uniform(st), u2 = nuts_uniform(st);
  return sqrt(-2.0 * log(u1)) * cospi(2.0 * u2);
}
__device__ __forceinline__ double nuts_logaddexp(double a, double b) {
  if (a == -INFINITY) return b;
  if (b == -INFINITY) return a;
  const double m = fmax(a, b);
  return m + log1p(exp(-fabs(a - b)));
}
// generalised U-turn criterion on momentum sums (symmetric in the two ends)
__device__ __forceinline__ bool nuts_turning(const NutsState& s, const double* im, const double* r_a,
                                             const double* r_b, const double* r_sum) {
  double da = 0.0, db = 0.0;
  for (int d = 0; d < s.D; ++d) {
    if (s.role[d] != 1) continue;
    const double rs = r_sum[d] - 0.5 * (r_a[d] + r_b[d]);
    da += im[d] * r_a[d] * rs;
    db += im[d] * r_b[d] * rs;
  }
  return da <= 0.0 || db <= 0.0;
}

// gradient of the evaluation point for dimension d (tied replicas folded onto cloud 0)
__device__ __forceinline__ double nuts_grad(const NutsState& s, int64_t b, int d) {
  double g = nuts_ev(s.egu, s.egce, s.egca, s, b, d);
  if (d < CHI_NSLOT * s.N && (d % s.N) == 0) {
    for (int n = 1; n < s.N; ++n)
      if (s.role[d + n] == 2) g += nuts_ev(s.egu, s.egce, s.egca, s, b, d + n);
  }
  return g;
}

// writes position vector x[D] of chain b as the evaluation point (replicas mirror cloud 0)
__device__ __forceinline__ void nuts_set_eval(const NutsState& s, int64_t b, const double* x) {
  for (int d = 0; d < s.D; ++d) {
    const int src = (s.role[d] == 2) ? d - (d % s.N) : d;
    nuts_ev(s.eu, s.ece, s.eca, s, b, d) = x[src];
  }
}

// ---- initial evaluation: take logp / gradient of the evaluation point as the current state ----
#ifndef CHI_DEVICE_ONLY
__global__ void k_nuts_adopt(NutsState s) {
  int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= s.B) return;
  double* q = s.q + b * s.D;
  double* g = s.g + b * s.D;
  for (int d = 0; d < s.D; ++d) {
    q[d] = nuts_ev(s.eu, s.ece, s.eca, s, b, d);
    g[d] = (s.role[d] == 1) ? nuts_grad(s, b, d) : 0.0;
  }
  s.logp[b] = s.elogp[b];
}
#endif

// ---- start of a NUTS iteration: fresh momentum, single-node tree --------------------------
__device__ __forceinline__ void nuts_begin_iter(const NutsState& s, int64_t b) {
  const int64_t o = b * s.D;
  unsigned long long st = s.rng[b];
  double kin = 0.0;
  for (int d = 0; d < s.D; ++d) {
    double p = 0.0;
    if (s.role[d] == 1) {
      p = nuts_normal(st) / sqrt(s.inv_mass[o + d]);
      kin += 0.5 * s.inv_mass[o + d] * p * p;
    }
    s.pl[o + d] = s.pr[o + d] = s.rsum[o + d] = p;
    s.ql[o + d] = s.qr[o + d] = s.qp[o + d] = s.q[o + d];
    s.gl[o + d] = s.gr[o + d] = s.gp[o + d] = s.g[o + d];
  }
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
#ifndef CHI_DEVICE_ONLY
__global__ void k_nuts_begin_iter(NutsState s) {
  int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= s.B) return;
  nuts_begin_iter(s, b);
  if (b == 0) *s.n_active = (int)s.B;
}
#endif

// ---- start of a doubling: pick the direction, put the cursor on that edge -----------------
__device__ __forceinline__ void nuts_begin_subtree(const NutsState& s, int64_t b) {
  const int64_t o = b * s.D;
  s.sub_stop[b] = s.active[b] ? 0 : 1;
  if (!s.active[b]) {
    nuts_set_eval(s, b, s.q + o);  // keep a valid evaluation point for idle chains
    atomicAdd(s.n_stopped, 1);
    return;
  }
  unsigned long long st = s.rng[b];
  const int dir = (nuts_next(st) & 1ULL) ? 1 : -1;
  s.rng[b] = st;
  s.dir[b] = dir;
  const double *eq = dir > 0 ? s.qr : s.ql, *ep = dir > 0 ? s.pr : s.pl, *eg = dir > 0 ? s.gr : s.gl;
  for (int d = 0; d < s.D; ++d) {
    s.qc[o + d] = eq[o + d];
    s.pc[o + d] = ep[o + d];
    s.gc[o + d] = eg[o + d];
    s.rsum_s[o + d] = 0.0;
  }
  s.lw_sub[b] = -INFINITY;
  s.leaf[b] = 0;
}
#ifndef CHI_DEVICE_ONLY
__global__ void k_nuts_begin_subtree(NutsState s) {
  int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= s.B) return;
  nuts_begin_subtree(s, b);
}
#endif

// ---- leapfrog, first half: half kick + drift; the new position becomes the evaluation point ---
__device__ __forceinline__ void nuts_leap_a(const NutsState& s, int64_t b) {
  const int64_t o = b * s.D;
  const double h = s.dir[b] * s.eps[b];
  double x[CHI_NSLOT * CHI_MAX_CLOUDS + 2 * CHI_MAX_BASELINE];
  for (int d = 0; d < s.D; ++d) {
    x[d] = s.qc[o + d];
    if (s.role[d] == 1) {
      const double ph = s.pc[o + d] + 0.5 * h * s.gc[o + d];
      s.ph[o + d] = ph;
      x[d] += h * s.inv_mass[o + d] * ph;
    }
  }
  nuts_set_eval(s, b, x);
}

// ---- leapfrog, second half + sub-tree bookkeeping of the chain's next leaf -------------------
// (the leaf index lives on the device so that one captured CUDA graph serves every leapfrog step)
__device__ __forceinline__ void nuts_leap_b(const NutsState& s, int64_t b) {
  const int leaf = s.leaf[b]++;
  const int64_t o = b * s.D;
  const double h = s.dir[b] * s.eps[b];
  const double lp = s.elogp[b];
  double kin = 0.0;
  for (int d = 0; d < s.D; ++d) {
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
  double energy = -lp + kin;
  if (!(energy == energy) || isinf(energy)) energy = INFINITY;
  const double delta = energy - s.energy0[b];
  const bool diverging = !(delta <= 1000.0);
  s.sum_acc[b] += (delta == delta) ? fmin(1.0, exp(-delta)) : 0.0;
  s.n_leap[b] += 1;

  // uniform (multinomial) transition inside the sub-tree
  unsigned long long st = s.rng[b];
  const double lw_leaf = -delta;
  const double lw_new = nuts_logaddexp(s.lw_sub[b], lw_leaf);
  const bool take = nuts_uniform(st) < exp(lw_leaf - lw_new);
  s.rng[b] = st;
  s.lw_sub[b] = lw_new;
  for (int d = 0; d < s.D; ++d) {
    s.rsum_s[o + d] += s.pc[o + d];
    if (take) {
      s.qs[o + d] = s.qc[o + d];
      s.gs[o + d] = s.gc[o + d];
    }
  }
  if (take) s.logp_s[b] = lp;

  // iterative U-turn test through the checkpoints (Phan et al. 2019)
  bool turning = false;
  const int idx_max = __popc((unsigned)leaf >> 1);
  double* ckr = s.ck_r + (b * CHI_NUTS_MAX_DEPTH) * s.D;
  double* ckrs = s.ck_rs + (b * CHI_NUTS_MAX_DEPTH) * s.D;
  if ((leaf & 1) == 0) {
    for (int d = 0; d < s.D; ++d) {
      ckr[idx_max * s.D + d] = s.pc[o + d];
      ckrs[idx_max * s.D + d] = s.rsum_s[o + d];
    }
  } else {
    const int n_sub = __ffs(~(unsigned)leaf) - 1;  // trailing ones of leaf
    const int idx_min = idx_max - n_sub + 1;
    double sub[CHI_NSLOT * CHI_MAX_CLOUDS + 2 * CHI_MAX_BASELINE];
    for (int i = idx_max; i >= idx_min && !turning; --i) {
      for (int d = 0; d < s.D; ++d) sub[d] = s.rsum_s[o + d] - ckrs[i * s.D + d] + ckr[i * s.D + d];
      turning = nuts_turning(s, s.inv_mass + o, ckr + i * s.D, s.pc + o, sub);
    }
  }
  if (diverging) s.diverged[b] = 1;
  if (turning || diverging) {
    s.sub_stop[b] = turning ? 2 : 3;  // 2: sub-tree turned, 3: diverged
    atomicAdd(s.n_stopped, 1);
  }
}

// One kernel per leapfrog boundary: finish the leaf whose logp/gradient just arrived (do_b) and,
// unless the chain's sub-tree stopped, start the next one (do_a).  The first launch of a
// doubling has do_b = 0, the last one do_a = 0.
#ifndef CHI_DEVICE_ONLY
__global__ void k_nuts_leap(NutsState s, int do_b, int do_a) {
  int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= s.B || s.sub_stop[b]) return;
  if (do_b) nuts_leap_b(s, b);
  if (do_a && !s.sub_stop[b]) nuts_leap_a(s, b);
}
#endif

// ---- end of a doubling: merge the sub-tree into the tree ------------------------------------
__device__ __forceinline__ void nuts_end_subtree(const NutsState& s, int64_t b, int max_depth) {
  const int64_t o = b * s.D;
  const int dir = s.dir[b];
  const bool sub_bad = s.sub_stop[b] >= 2;
  double *eq = dir > 0 ? s.qr : s.ql, *ep = dir > 0 ? s.pr : s.pl, *eg = dir > 0 ? s.gr : s.gl;
  unsigned long long st = s.rng[b];
  // biased progressive sampling between the old tree and the new sub-tree
  const bool take = !sub_bad && (nuts_uniform(st) < fmin(1.0, exp(s.lw_sub[b] - s.lw_tree[b])));
  s.rng[b] = st;
  for (int d = 0; d < s.D; ++d) {
    eq[o + d] = s.qc[o + d];
    ep[o + d] = s.pc[o + d];
    eg[o + d] = s.gc[o + d];
    s.rsum[o + d] += s.rsum_s[o + d];
    if (take) {
      s.qp[o + d] = s.qs[o + d];
      s.gp[o + d] = s.gs[o + d];
    }
  }
  if (take) s.logp_p[b] = s.logp_s[b];
  s.lw_tree[b] = nuts_logaddexp(s.lw_tree[b], s.lw_sub[b]);
  s.depth[b] += 1;
  const bool turning = sub_bad || nuts_turning(s, s.inv_mass + o, s.pl + o, s.pr + o, s.rsum + o);
  if (turning || s.depth[b] >= max_depth) {
    s.active[b] = 0;
    atomicSub(s.n_active, 1);
  }
}
#ifndef CHI_DEVICE_ONLY
__global__ void k_nuts_end_subtree(NutsState s, int max_depth) {
  int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= s.B || !s.active[b]) return;
  nuts_end_subtree(s, b, max_depth);
}
#endif

// ---- end of the iteration: move to the proposal, adapt, record -------------------------------
struct NutsAdapt {
  int tuning;          // 1 during warm-up
  int adapt_step;      // dual averaging of the step size
  int collect_mass;    // accumulate Welford statistics of the position
  int close_window;    // set inv_mass from the window and restart the step-size averaging
  int finalize;        // last tuning step: switch to the averaged step size
  double target_accept;
};

// Pooled mass adaptation (chains that share one posterior): at the end of an adaptation window the
// variance of every free dimension is estimated from ALL chains' window samples (their Welford
// statistics are merged), which needs B times fewer iterations per window than PyMC's per-chain
// estimate for the same accuracy.  One block per dimension; then every chain restarts its window
// and its step-size averaging (k_nuts_restart_window).
#ifndef CHI_DEVICE_ONLY
__global__ void __launch_bounds__(128) k_nuts_pool_mass(NutsState s) {
  const int d = blockIdx.x;
  if (s.role[d] != 1) return;
  __shared__ double sh[128];
  auto block_sum = [&](double x) {
    sh[threadIdx.x] = x;
    __syncthreads();
    for (int o = 64; o > 0; o >>= 1) {
      if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
      __syncthreads();
    }
    const double r = sh[0];
    __syncthreads();
    return r;
  };
  double cnt = 0.0, sm = 0.0;
  for (int64_t b = threadIdx.x; b < s.B; b += blockDim.x) {
    const double n = s.w_count[b];
    cnt += n;
    sm += n * s.w_mean[b * s.D + d];
  }
  cnt = block_sum(cnt);
  sm = block_sum(sm);
  if (cnt < 2.0) return;
  const double mean = sm / cnt;
  double m2 = 0.0;
  for (int64_t b = threadIdx.x; b < s.B; b += blockDim.x) {
    const double n = s.w_count[b], dm = s.w_mean[b * s.D + d] - mean;
    m2 += s.w_m2[b * s.D + d] + n * dm * dm;
  }
  m2 = block_sum(m2);
  const double var = m2 / (cnt - 1.0);
  const double inv = (cnt / (cnt + 5.0)) * var + 1e-3 * (5.0 / (cnt + 5.0));  // Stan's regularisation
  for (int64_t b = threadIdx.x; b < s.B; b += blockDim.x) s.inv_mass[b * s.D + d] = inv;
}
#endif

#ifndef CHI_DEVICE_ONLY
__global__ void k_nuts_restart_window(NutsState s) {
  int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= s.B) return;
  const int64_t o = b * s.D;
  for (int d = 0; d < s.D; ++d) {
    s.w_mean[o + d] = 0.0;
    s.w_m2[o + d] = 0.0;
  }
  s.w_count[b] = 0;
  s.da_mu[b] = log(10.0 * s.eps[b]);
  s.da_hbar[b] = 0.0;
  s.da_logeps_bar[b] = 0.0;
  s.da_count[b] = 0;
}
#endif

__device__ __forceinline__ void nuts_end_iter(const NutsState& s, const NutsAdapt& a, int64_t b, double* draws,
                                              double* stat_accept, int* stat_depth, int* stat_diverged,
                                              double* stat_logp, double* stat_eps, int64_t draw_index) {
  const int64_t o = b * s.D;
  for (int d = 0; d < s.D; ++d) {
    s.q[o + d] = s.qp[o + d];
    s.g[o + d] = s.gp[o + d];
  }
  s.logp[b] = s.logp_p[b];
  const double acc = s.n_leap[b] > 0 ? s.sum_acc[b] / s.n_leap[b] : 0.0;

  if (a.tuning) {
    if (a.adapt_step) {  // Nesterov dual averaging (Hoffman & Gelman 2014), gamma 0.05, t0 10, kappa 0.75
      const int c = ++s.da_count[b];
      const double w = 1.0 / (c + 10.0);
      s.da_hbar[b] = (1.0 - w) * s.da_hbar[b] + w * (a.target_accept - acc);
      const double log_eps = s.da_mu[b] - sqrt((double)c) / 0.05 * s.da_hbar[b];
      const double eta = pow((double)c, -0.75);
      s.da_logeps_bar[b] = eta * log_eps + (1.0 - eta) * s.da_logeps_bar[b];
      s.eps[b] = exp(log_eps);
    }
    if (a.collect_mass) {
      const int c = ++s.w_count[b];
      for (int d = 0; d < s.D; ++d) {
        if (s.role[d] != 1) continue;
        const double x = s.q[o + d], dm = x - s.w_mean[o + d];
        s.w_mean[o + d] += dm / c;
        s.w_m2[o + d] += dm * (x - s.w_mean[o + d]);
      }
    }
    if (a.close_window && s.w_count[b] > 1) {
      const double n = s.w_count[b];
      for (int d = 0; d < s.D; ++d) {
        if (s.role[d] != 1) continue;
        const double var = s.w_m2[o + d] / (n - 1.0);
        s.inv_mass[o + d] = (n / (n + 5.0)) * var + 1e-3 * (5.0 / (n + 5.0));  // Stan's regularisation
        s.w_mean[o + d] = 0.0;
        s.w_m2[o + d] = 0.0;
      }
      s.w_count[b] = 0;
      s.da_mu[b] = log(10.0 * s.eps[b]);
      s.da_hbar[b] = 0.0;
      s.da_logeps_bar[b] = 0.0;
      s.da_count[b] = 0;
    }
    if (a.finalize) s.eps[b] = exp(s.da_logeps_bar[b]);
  }
  if (draws) {
    double* out = draws + (draw_index * s.B + b) * s.D;
    for (int d = 0; d < s.D; ++d) out[d] = s.q[o + d];
    stat_accept[draw_index * s.B + b] = acc;
    stat_depth[draw_index * s.B + b] = s.depth[b];
    stat_diverged[draw_index * s.B + b] = s.diverged[b];
    stat_logp[draw_index * s.B + b] = s.logp[b];
    stat_eps[draw_index * s.B + b] = s.eps[b];
  }
}
#ifndef CHI_DEVICE_ONLY
__global__ void k_nuts_end_iter(NutsState s, NutsAdapt a, double* draws, double* stat_accept, int* stat_depth,
                                int* stat_diverged, double* stat_logp, double* stat_eps, int64_t draw_index) {
  int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= s.B) return;
  nuts_end_iter(s, a, b, draws, stat_accept, stat_depth, stat_diverged, stat_logp, stat_eps, draw_index);
}
#endif

// ---- asynchronous chains ----------------------------------------------------------------------
// In lock step every chain waits, in every iteration, for the deepest tree among the chains, so the
// cost of an iteration is 2^(max depth over chains) leapfrog steps.  Here each chain runs its own
// state machine: one tick = one batched logp+grad evaluation that serves the leaf in flight of
// every chain, whatever iteration / doubling that chain is in; a chain that finishes its tree
// records the draw, adapts (its own iteration index drives the warm-up schedule) and starts the next
// iteration in the same tick.  Total ticks = the largest per-chain sum of tree sizes, not the sum
// of per-iteration maxima.
struct NutsAsync {
  int n_tune, n_total, init_buf, term_buf, max_depth;
  double target_accept;
  const unsigned char* closes;  // [n_tune] 1 where an adaptation window closes
  int* iter;                    // [B] iteration the chain is in
  int* started;                 // [B] 0 before the chain's first leaf
  int* n_done;                  // [1] chains that completed n_total iterations
  double *draws, *stat_accept, *stat_logp, *stat_eps;
  int *stat_depth, *stat_diverged;
};

#ifndef CHI_DEVICE_ONLY
__global__ void k_nuts_tick(NutsState s, NutsAsync A) {
  int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= s.B) return;
  int it = A.iter[b];
  if (it >= A.n_total) return;  // finished: its evaluation point stays at the last state
  if (A.started[b]) {
    nuts_leap_b(s, b);  // the logp / gradient of the leaf in flight has arrived
    if (s.sub_stop[b] || s.leaf[b] >= (1 << s.depth[b])) {
      nuts_end_subtree(s, b, A.max_depth);
      if (!s.active[b]) {
        NutsAdapt a;
        a.target_accept = A.target_accept;
        a.tuning = it < A.n_tune;
        a.adapt_step = a.tuning;
        a.collect_mass = a.tuning && it >= A.init_buf && it < A.n_tune - A.term_buf;
        a.close_window = a.tuning && A.closes[it];
        a.finalize = a.tuning && it == A.n_tune - 1;
        nuts_end_iter(s, a, b, a.tuning ? nullptr : A.draws, A.stat_accept, A.stat_depth, A.stat_diverged,
                      A.stat_logp, A.stat_eps, a.tuning ? 0 : it - A.n_tune);
        A.iter[b] = ++it;
        if (it >= A.n_total) {
          nuts_set_eval(s, b, s.q + b * s.D);
          atomicAdd(A.n_done, 1);
          return;
        }
        nuts_begin_iter(s, b);
      }
      nuts_begin_subtree(s, b);
    }
  } else {
    A.started[b] = 1;
    nuts_begin_iter(s, b);
    nuts_begin_subtree(s, b);
  }
  nuts_leap_a(s, b);
}
#endif

// initial per-chain state of the adaptation
#ifndef CHI_DEVICE_ONLY
__global__ void k_nuts_init(NutsState s, unsigned long long seed, double eps0) {
  int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= s.B) return;
  const int64_t o = b * s.D;
  unsigned long long st = seed ^ (0xD1B54A32D192ED03ULL * (unsigned long long)(b + 1));
  nuts_next(st);
  s.rng[b] = st;
  s.eps[b] = eps0;
  s.da_mu[b] = log(10.0 * eps0);
  s.da_hbar[b] = 0.0;
  s.da_logeps_bar[b] = 0.0;
  s.da_count[b] = 0;
  s.w_count[b] = 0;
  for (int d = 0; d < s.D; ++d) {
    s.inv_mass[o + d] = 1.0;
    s.w_mean[o + d] = 0.0;
    s.w_m2[o + d] = 0.0;
  }
}
#endif

}  // namespace chi
