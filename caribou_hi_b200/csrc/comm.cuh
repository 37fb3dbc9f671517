// Multi-GPU part of the library: the arithmetic of the path has no exchange step -- draws,
// chains and sightlines are independent -- so NCCL over NVLink only COLLECTS: the per-shard sums of logp
// and of the gradient (k_shard_sums below, then an all-gather or all-reduce of 2 + n_rows N doubles),
// or rows of per-draw results.  The handle owns the communicator and a communication stream; a
// collective waits for what the handle's stream has enqueued so far and runs NEXT TO whatever the
// caller enqueues afterwards (chi_comm_join orders the handle's stream behind it again).
//
// NCCL is loaded at run time (dlopen "libnccl.so.2": the copy a hosting PyTorch process has already
// loaded, else the system's) through the seven entry points declared here, so the library has no
// link-time dependency on it and single-GPU users never touch it.
#pragma once
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <stdint.h>

namespace chi {

#define CHI_NCCL_ID_BYTES 128
struct NcclId { char internal[CHI_NCCL_ID_BYTES]; };  // ncclUniqueId
typedef void* NcclComm;
enum { CHI_NCCL_SUM = 0, CHI_NCCL_FLOAT64 = 8 };       // ncclSum, ncclFloat64

struct NcclApi {
  void* lib = nullptr;
  int (*GetUniqueId)(NcclId*) = nullptr;
  int (*CommInitRank)(NcclComm*, int, NcclId, int) = nullptr;
  int (*CommDestroy)(NcclComm) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, NcclComm, cudaStream_t) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  int (*GetVersion)(int*) = nullptr;
  bool load(const char** why) {
    if (lib) return true;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
      lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (lib) break;
    }
    if (!lib) { *why = "libnccl.so.2 not found (dlopen)"; return false; }
    bool ok = true;
    auto sym = [&](const char* name) { void* p = dlsym(lib, name); if (!p) ok = false; return p; };
    GetUniqueId = (int (*)(NcclId*))sym("ncclGetUniqueId");
    CommInitRank = (int (*)(NcclComm*, int, NcclId, int))sym("ncclCommInitRank");
    CommDestroy = (int (*)(NcclComm))sym("ncclCommDestroy");
    AllGather = (int (*)(const void*, void*, size_t, int, NcclComm, cudaStream_t))sym("ncclAllGather");
    AllReduce = (int (*)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t))sym("ncclAllReduce");
    GetErrorString = (const char* (*)(int))sym("ncclGetErrorString");
    GetVersion = (int (*)(int*))sym("ncclGetVersion");
    if (!ok) { *why = "libnccl.so.2 lacks an expected symbol"; dlclose(lib); lib = nullptr; return false; }
    return true;
  }
};

// ---------------------------------------------------------------------------------------------------
// k_shard_sums: sums over the FINITE draws of a shard of logp and of every gradient entry, and the number
// of non-finite draws (a NaN draw -- 31 % of the prior draws of the physical models, where the reference
// is NaN too -- would otherwise poison the totals; the count tells the caller).  One pass over
// logp[B] and grad[B][W] (W = n_rows N, plus the baseline-coefficient gradients when given), fixed
// summation order (deterministic), no atomics on data: every block writes its partial sums, the block
// that finishes last adds the partials in block order.
//   out[0] = sum logp, out[1 .. W] = sum grad, out[1 + W] = number of non-finite draws
// ---------------------------------------------------------------------------------------------------
#define CHI_SUMS_THREADS 256
#define CHI_SUMS_MAXW 256
__global__ void __launch_bounds__(CHI_SUMS_THREADS) k_shard_sums(int64_t B, int W, int draws_per_block,
                                                                 const double* __restrict__ logp,
                                                                 const double* __restrict__ grad,
                                                                 double* __restrict__ partial,  // [grid][W + 2]
                                                                 unsigned* __restrict__ counter,
                                                                 double* __restrict__ out) {
  __shared__ double s_acc[CHI_SUMS_THREADS / 32][CHI_SUMS_MAXW + 2];
  __shared__ int s_last;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, nw = CHI_SUMS_THREADS / 32;
  const int64_t b0 = (int64_t)blockIdx.x * draws_per_block;
  const int64_t b1 = min(B, b0 + draws_per_block);
  // a warp walks draws b0 + warp, b0 + warp + nw, ...; lanes walk the W gradient entries of a draw
  double acc[CHI_SUMS_MAXW / 32];
#pragma unroll
  for (int i = 0; i < CHI_SUMS_MAXW / 32; ++i) acc[i] = 0.0;
  double lp_sum = 0.0, bad = 0.0;
  for (int64_t b = b0 + warp; b < b1; b += nw) {
    const double lp = logp[b];
    const bool ok = isfinite(lp);
    if (ok) lp_sum += lp; else bad += 1.0;
    const double* g = grad + b * W;
#pragma unroll
    for (int i = 0; i < CHI_SUMS_MAXW / 32; ++i) {
      const int w = lane + 32 * i;
      if (w < W) {
        const double v = g[w];
        if (ok) acc[i] += v;
      }
    }
  }
#pragma unroll
  for (int i = 0; i < CHI_SUMS_MAXW / 32; ++i)
    if (lane + 32 * i < W) s_acc[warp][1 + lane + 32 * i] = acc[i];
  if (lane == 0) { s_acc[warp][0] = lp_sum; s_acc[warp][W + 1] = bad; }
  __syncthreads();
  double* mine = partial + (int64_t)blockIdx.x * (W + 2);
  for (int t = tid; t < W + 2; t += CHI_SUMS_THREADS) {
    double v = 0.0;
    for (int w = 0; w < nw; ++w) v += s_acc[w][t];
    mine[t] = v;
  }
  __threadfence();
  __syncthreads();
  if (tid == 0) {
    const unsigned done = atomicAdd(counter, 1u);
    s_last = (done == gridDim.x - 1);
    if (s_last) *counter = 0;
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  for (int t = tid; t < W + 2; t += CHI_SUMS_THREADS) {
    double v = 0.0;
    for (unsigned k = 0; k < gridDim.x; ++k) v += __ldcg(partial + (int64_t)k * (W + 2) + t);
    out[t] = v;
  }
}

}  // namespace chi
