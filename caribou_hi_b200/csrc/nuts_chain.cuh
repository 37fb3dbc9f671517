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
