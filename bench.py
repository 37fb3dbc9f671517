#!/usr/bin/env python
"""bench.py -- logp+grad throughput of the caribou_hi likelihood hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

A *step* is one pass of the hot path (cloud map -> emission/absorption moment kernels
-> epilogue: logp and the full gradient) over one batch of B synthetic prior draws of
BASELINE.json configs[1]: EmissionAbsorptionPhysicalModel, 4 clouds, 1000 emission +
1000 absorption channels, B = 1e5 draws per GPU (weak scaling over draws).

Metric: draw*cloud*channel units per second, units/step = B*N*(S_e+S_a) per GPU.
Prints ONE JSON line on rank 0 (see the driver contract in DESIGN.md section 6).
"""

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "logp+grad evals/s (draw*cloud*channel units/s)"
UNIT = "units/s"
KIND = "emission_absorption_physical"
N_CLOUDS = 4
S_E = 1000
S_A = 1000
DRAWS = 100_000
N_SETS = 6  # rotated input/output buffer sets: 6 x 65 MB = 390 MB > 126 MB L2
# canonical algorithmic flops per unit (FMA=2, exp=1, div=1), Gaussian profile
FLOPS_EM, FLOPS_ABS = 32.0, 12.0
SEED = 1234


def workload_name(draws):
    return (f"EmissionAbsorptionPhysicalModel N={N_CLOUDS} S_e={S_E} S_a={S_A} "
            f"B={draws} prior draws per GPU, Gaussian profile (prior_fwhm_L=None), baseline_degree=0")


# --------------------------------------------------------------------------------------
# clocks sampler (nvidia-smi during the timed region)
# --------------------------------------------------------------------------------------
NCU_DRAM_BYTES_PER_DRAW = 8083712.0 / 20000.0  # k_moments_ea<4,0,1,0,1,EA_TAME>, profiles/r2c_ncu_full_summary.txt


class ClockSampler:
    """Samples SM clock and throttle reasons through NVML (nvidia_ml_py) from a thread while the
    timed region runs; falls back to one `nvidia-smi` query when NVML is unavailable."""

    def __init__(self, index):
        self.index = index
        self.samples = []
        self.stop_flag = threading.Event()
        self.thread = None
        self.nv = None

    def start(self):
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.thread = threading.Thread(target=self._run, daemon=True)
            self.thread.start()
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        while not self.stop_flag.is_set():
            try:
                tq = time.perf_counter()
                clk = nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM)
                try:
                    reasons = nv.nvmlDeviceGetCurrentClocksEventReasons(self.handle)
                except Exception:
                    reasons = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
                # a query takes tens of ms on a busy GPU: keep its whole [begin, end] interval
                self.samples.append((tq, time.perf_counter(), clk, reasons))
            except Exception:
                pass
            time.sleep(0.002)

    def count_inside(self, t0, t1):
        return sum(1 for (a, b, _, _) in list(self.samples) if b >= t0 and a <= t1)

    def stop(self, t0, t1, t2=None):
        """Samples whose query interval overlaps the timed region [t0, t1]; when the region was too
        short to catch three, also those of [t1, t2] (the same steps kept running untimed)."""
        if self.nv is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["NVML unavailable"]}
        self.stop_flag.set()
        self.thread.join(timeout=1.0)
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
        inside = [(c, r) for (a, b, c, r) in self.samples if b >= t0 and a <= t1]
        window = "timed region"
        n_timed = len(inside)
        if t2 is not None and n_timed < 3:
            inside = [(c, r) for (a, b, c, r) in self.samples if b >= t0 and a <= t2]
            window = "timed region + the same steps kept running untimed right after it (NVML queries take longer than the region)"
        reasons = sorted({k for (_, r) in inside for k, bit in names.items() if r & bit})
        try:
            mx = float(nv.nvmlDeviceGetMaxClockInfo(self.handle, nv.NVML_CLOCK_SM))
        except Exception:
            mx = None
        return {"sm_mhz": float(np.median([c for c, _ in inside])) if inside else None, "sm_max_mhz": mx,
                "samples": len(inside), "samples_in_timed_region": n_timed, "window": window, "reasons": reasons}


# --------------------------------------------------------------------------------------
# CPU baseline: the restated oracle (torch-CPU fp64 + autograd), all host threads
# --------------------------------------------------------------------------------------
def reference_available():
    try:
        import bayes_spec  # noqa: F401
        import pymc  # noqa: F401
        import pytensor  # noqa: F401
        return True
    except Exception:
        return False


def oracle_setup(draws, seed=SEED):
    """Same seeded workload, built with the CPU oracle only (no GPU needed)."""
    import torch  # noqa: F401
    from oracle import models_ref as mr
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from helpers import draw_free_oracle

    spec = mr.make_spec(KIND, N_CLOUDS)
    rng = np.random.default_rng(seed)
    ve = np.linspace(-100.0, 100.0, S_E)
    va = np.linspace(-100.0, 100.0, S_A)
    truth_free = draw_free_oracle(spec, 1, np.random.default_rng(seed + 1))
    dummy = {"emission": mr.OracleData(ve, np.zeros(S_E), 0.1), "absorption": mr.OracleData(va, np.zeros(S_A), 0.01)}
    truth = mr.predict(spec, dummy, truth_free)
    data = {
        "emission": mr.OracleData(ve, np.nan_to_num(truth["emission"][0]) + 0.1 * rng.standard_normal(S_E), 0.1),
        "absorption": mr.OracleData(va, np.nan_to_num(truth["absorption"][0]) + 0.01 * rng.standard_normal(S_A), 0.01),
    }
    free = draw_free_oracle(spec, draws, rng)
    return spec, data, free


def oracle_step(spec, data, free_t):
    import torch
    from oracle import models_ref as mr

    for v in free_t.values():
        v.grad = None
    lp = mr.logp(spec, data, free_t)
    torch.nan_to_num(lp, nan=0.0, posinf=0.0, neginf=0.0).sum().backward()
    return lp


def time_oracle(sample_draws, steps, warmup, min_seconds=0.0):
    import torch

    spec, data, free = oracle_setup(sample_draws)
    free_t = {k: torch.tensor(v, dtype=torch.float64, requires_grad=True) for k, v in free.items()}
    for _ in range(warmup):
        oracle_step(spec, data, free_t)
    t0 = time.perf_counter()
    done = 0
    while done < steps or (time.perf_counter() - t0) < min_seconds:
        oracle_step(spec, data, free_t)
        done += 1
    dt = time.perf_counter() - t0
    units = sample_draws * N_CLOUDS * (S_E + S_A)
    return units * done / dt, dt / done, done, torch.get_num_threads()


def time_c_oracle(sample_draws, steps, warmup, min_seconds=0.0):
    """The plain-C port (oracle/chi_oracle.c): fused loops like PyTensor's compiled graph, OpenMP
    over draws standing in for PyMC's process-per-chain parallelism, all host threads.  The O(N)
    parameter maps are evaluated once outside the timed loop (they are < 1 % of the work)."""
    from oracle import c_oracle
    from oracle import models_ref as mr

    spec, data, free = oracle_setup(sample_draws)
    det = mr.cloud_inputs(spec, free)
    det = {k: np.ascontiguousarray(np.broadcast_to(det[k], free["fwhm2_norm"].shape)) for k in c_oracle.INPUTS}
    cores = 1
    # every host thread this process may use, stated explicitly: torchrun exports
    # OMP_NUM_THREADS=1 to its workers, which would time a single-threaded baseline
    n_thr = len(os.sched_getaffinity(0))
    for _ in range(warmup):
        cores = c_oracle.logp_grad(data, det, spec["bg_temp"], n_threads=n_thr)[2]
    t0 = time.perf_counter()
    done = 0
    while done < steps or (time.perf_counter() - t0) < min_seconds:
        c_oracle.logp_grad(data, det, spec["bg_temp"], n_threads=n_thr)
        done += 1
    dt = time.perf_counter() - t0
    units = sample_draws * N_CLOUDS * (S_E + S_A)
    return units * done / dt, dt / done, done, cores


CPU_NOTE = ("plain-C port of the reference arithmetic (oracle/chi_oracle.c, gcc -O3, OpenMP over draws); "
            "the reference's PyTensor/PyMC path is not installable here")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sample = 4096
    note = CPU_NOTE
    value, s_per_step, done, cores = time_c_oracle(sample, args.steps, max(args.warmup, 1))
    out = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": done, "warmup": max(args.warmup, 1), "ms_per_step": s_per_step * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(DRAWS), "sample": f"{sample} of the {DRAWS} draws per step"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{sample} draws x {done} steps; {note}"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out))


# --------------------------------------------------------------------------------------
# B200 arm
# --------------------------------------------------------------------------------------
def build_inputs(eng, draws, rank, n_sets):
    """Seeded synthetic spectra (truth from the CUDA forward model + noise) and n_sets
    batches of prior draws of the free RVs."""
    from caribou_hi_b200 import synthetic

    key = "fwhm2_thermal"

    def thermal(trial):
        return eng.deterministics(eng.pack(trial))[key]

    ve = np.linspace(-100.0, 100.0, S_E)
    va = np.linspace(-100.0, 100.0, S_A)
    rng = np.random.default_rng(SEED)
    # provisional spectra so the engine can evaluate the forward model of the truth draw
    eng.set_spectra(emission=dict(spectral=ve, brightness=np.zeros(S_E), noise=0.1),
                    absorption=dict(spectral=va, brightness=np.zeros(S_A), noise=0.01))
    truth = synthetic.draw_free(KIND, N_CLOUDS, 1, np.random.default_rng(SEED + 1), thermal_fn=thermal)
    mu_e, mu_a = eng.spectra(eng.pack(truth))
    ye = np.nan_to_num(mu_e[0]) + 0.1 * rng.standard_normal(S_E)
    ya = np.nan_to_num(mu_a[0]) + 0.01 * rng.standard_normal(S_A)
    # baseline_degree = 0: one constant term per dataset (bayes_spec normalisation)
    eng.set_spectra(
        emission=dict(spectral=ve, brightness=ye, noise=0.1, basis=np.full((1, S_E), 0.1), offset=float(np.median(ye))),
        absorption=dict(spectral=va, brightness=ya, noise=0.01, basis=np.full((1, S_A), 0.01), offset=float(np.median(ya))),
    )
    thetas = []
    for i in range(n_sets):
        r = np.random.default_rng([SEED, rank, i])
        free = synthetic.draw_free(KIND, N_CLOUDS, draws, r, thermal_fn=thermal)
        thetas.append(eng.pack(free))
    return thetas


def run_b200(args):
    import torch
    import torch.distributed as dist

    from caribou_hi_b200 import Engine, PinnedArray, sharding

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 arm has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    draws = args.draws
    N = N_CLOUDS

    eng = Engine(KIND, N, device=local, prior_amplitude=1.0e21)
    # the library's own NCCL communicator (chi_comm_init); torch.distributed only ships its 128-byte id,
    # provides the barriers around the timed region and the max-over-ranks of the timings
    sharding.init_engine_comm(eng)
    thetas = build_inputs(eng, draws, rank, N_SETS)
    dev = torch.device("cuda", local)
    th_d = [torch.from_numpy(t).to(dev) for t in thetas]
    ce_d = torch.zeros((draws, 1), dtype=torch.float64, device=dev)
    ca_d = torch.zeros((draws, 1), dtype=torch.float64, device=dev)
    lp_d = [torch.empty(draws, dtype=torch.float64, device=dev) for _ in range(N_SETS)]
    gr_d = [torch.empty((draws, 10, N), dtype=torch.float64, device=dev) for _ in range(N_SETS)]
    gce_d = torch.empty((draws, 1), dtype=torch.float64, device=dev)
    gca_d = torch.empty((draws, 1), dtype=torch.float64, device=dev)
    stream = torch.cuda.ExternalStream(eng.stream, device=dev)
    width = 10 * N
    sums_d = [torch.empty(width + 2, dtype=torch.float64, device=dev) for _ in range(N_SETS)]
    gathered = torch.empty((world, width + 2), dtype=torch.float64, device=dev) if world > 1 else None
    torch.cuda.synchronize()

    # the exchange of step i -- the shard's sums of logp and gradient (chi_shard_sums_dev) and their
    # all-gather over the ranks (chi_allgather_dev: NCCL) -- runs on the library's communication
    # stream next to the kernels of step i+1
    side = torch.cuda.ExternalStream(eng.comm_stream, device=dev) if world > 1 else None
    ev_read = [torch.cuda.Event() for _ in range(N_SETS)]   # exchange has read buffer set k
    read_pending = [False] * N_SETS

    def step(i, exchange=True):
        k = i % N_SETS
        if read_pending[k]:
            stream.wait_event(ev_read[k])  # do not overwrite outputs the exchange still reads
            read_pending[k] = False
        eng.logp_grad_dev(draws, th_d[k], lp_d[k], gr_d[k], coeff_e=ce_d, coeff_a=ca_d, gcoeff_e=gce_d, gcoeff_a=gca_d)
        if world > 1 and exchange:
            # the only collective of the path: gather per-shard logp and gradient sums
            eng.shard_sums_dev(draws, width, lp_d[k], gr_d[k], sums_d[k])
            eng.allgather_dev(sums_d[k], gathered, width + 2)
            ev_read[k].record(side)
            read_pending[k] = True

    def join_side():
        if side is not None:
            eng.comm_join()  # the timed region ends after the last exchange

    def barrier():
        eng.synchronize()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    for i in range(args.warmup):
        step(i)
    barrier()
    launches0 = eng.launch_count()
    stage0 = eng.stage_count()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    barrier()
    t_start = time.perf_counter()
    e0.record(stream)
    for i in range(args.steps):
        step(i)
    join_side()
    e1.record(stream)
    barrier()
    t_end = time.perf_counter()
    ms = e0.elapsed_time(e1)
    launches = eng.launch_count() - launches0
    # the library runs a large device batch as `pieces` launches of every kernel (three lanes): the
    # stage events are per piece, i.e. per LAUNCH
    n_stage = eng.stage_count() - stage0
    pieces = max(1, round(n_stage / max(args.steps, 1)))
    stage = eng.stage_ms(min(n_stage, 63)).mean(axis=0)  # events recorded during the timed steps
    clocks = None
    if rank == 0:
        # a timed region shorter than a few NVML queries: hold the same load (untimed) until the
        # sampler has seen it three times, at most 1.5 s
        t_hold = t_end
        i = args.steps
        while sampler.nv is not None and sampler.count_inside(t_start, t_hold) < 3 and t_hold - t_end < 1.5:
            for _ in range(8):
                step(i, exchange=False)  # rank 0 alone: no collective here
                i += 1
            eng.synchronize()
            t_hold = time.perf_counter()
        clocks = sampler.stop(t_start, t_end, t_hold)
    barrier()
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    units_step = draws * N * (S_E + S_A)
    value = world * units_step * args.steps / (ms * 1e-3)

    # ---- the same device-resident steps with the library's opt-in overlap of consecutive calls ----
    # (CHI_OPT_DEV_OVERLAP: calls alternate between two internal lanes, so the cloud map of step i+1 and the
    # epilogue of step i run next to the other step's channel kernel; outputs are ordered into the handle's stream
    # by chi_dev_join.  Reported beside `value`, which stays the default one-stream form.)
    overlapped = None
    if world == 1 and not args.no_extras:
        eng.set_option("dev_overlap", 1)
        for i in range(args.warmup):
            step(i, exchange=False)
        eng.dev_join()
        barrier()
        o0, o1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        o0.record(stream)
        for i in range(args.steps):
            step(i, exchange=False)
        eng.dev_join()
        o1.record(stream)
        barrier()
        eng.set_option("dev_overlap", 0)
        oms = o0.elapsed_time(o1)
        overlapped = {"value": units_step * args.steps / (oms * 1e-3), "unit": "units/s", "ms_per_step": oms / args.steps,
                      "how": "chi_set_option(CHI_OPT_DEV_OVERLAP, 1): consecutive chi_logp_grad_dev calls on two alternating "
                             "lanes, chi_dev_join before the closing event; same results bit for bit "
                             "(tests/test_gpu_comm.py::test_overlapped_device_calls_give_the_same_bits)"}

    # ---- end to end through the public host API: pinned host -> device -> pinned host ----
    # packed blocks: exactly the free RVs of the model class in, exactly their gradients out
    e2e_steps = max(2, min(args.steps, 30))
    rows = eng.row_slots
    pin_th = [PinnedArray((draws, rows.size, N)) for _ in range(2)]
    for k in range(2):
        pin_th[k].array[...] = thetas[k][:, rows, :]
    pin_lp = PinnedArray((draws,))
    pin_gr = PinnedArray((draws, rows.size, N))
    pins = [PinnedArray((draws, 1)) for _ in range(4)]  # coefficients in, coefficient gradients out
    ce_h, ca_h = pins[0].array, pins[1].array
    ce_h[...] = 0.0
    ca_h[...] = 0.0
    out = {"logp": pin_lp.array, "grad": pin_gr.array, "gcoeff_e": pins[2].array, "gcoeff_a": pins[3].array}
    for _ in range(2):
        eng.logp_grad_rows(pin_th[0].array, ce_h, ca_h, out=out)  # warm-up (allocates staging)
    barrier()
    # (a) one blocking call per step
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        eng.logp_grad_rows(pin_th[i % 2].array, ce_h, ca_h, out=out)
    t_sync = (time.perf_counter() - t0) / e2e_steps
    # (b) the same calls through the asynchronous entry point, two in flight on double-buffered pinned
    # arrays: every step still uploads its inputs and downloads its results (the logp of step i-1 is
    # read on the host while step i runs), but one call's pipeline fill / drain hides behind the
    # other's kernels
    pin_out2 = [PinnedArray((draws,)), PinnedArray((draws, rows.size, N)), PinnedArray((draws, 1)), PinnedArray((draws, 1))]
    outs = [out, {"logp": pin_out2[0].array, "grad": pin_out2[1].array, "gcoeff_e": pin_out2[2].array,
                  "gcoeff_a": pin_out2[3].array}]
    eng.wait(eng.logp_grad_rows_async(pin_th[0].array, ce_h, ca_h, outs[1]))  # warm-up
    barrier()
    checksum = 0.0
    t0 = time.perf_counter()
    prev = None
    for i in range(e2e_steps):
        tk = eng.logp_grad_rows_async(pin_th[i % 2].array, ce_h, ca_h, outs[i % 2])
        if prev is not None:
            eng.wait(prev)
            checksum += float(np.nansum(outs[(i - 1) % 2]["logp"][:16]))  # the host reads the result
        prev = tk
    eng.wait(prev)
    checksum += float(np.nansum(outs[(e2e_steps - 1) % 2]["logp"][:16]))
    t_async = (time.perf_counter() - t0) / e2e_steps
    t_e2e = min(t_sync, t_async)
    # (c) the same steps for a consumer of TOTALS (what a sharded job gathers): chi_logp_sums_rows_async
    # uploads the same inputs, computes logp and the full gradient, and downloads logp per draw plus
    # the shard's gradient sums -- the per-draw gradient stays on the device
    pin_lp2 = [PinnedArray((draws,)) for _ in range(2)]
    outs_s = [{"logp": pin_lp2[k].array, "sums": np.empty(2 + rows.size * N)} for k in range(2)]
    eng.wait(eng.logp_sums_rows_async(pin_th[0].array, ce_h, ca_h, outs_s[1]))  # warm-up
    barrier()
    t0 = time.perf_counter()
    prev = None
    for i in range(e2e_steps):
        tk = eng.logp_sums_rows_async(pin_th[i % 2].array, ce_h, ca_h, outs_s[i % 2])
        if prev is not None:
            eng.wait(prev)
            checksum += float(outs_s[(i - 1) % 2]["sums"][0])  # the host reads the result
        prev = tk
    eng.wait(prev)
    checksum += float(outs_s[(e2e_steps - 1) % 2]["sums"][0])
    t_sums = (time.perf_counter() - t0) / e2e_steps
    # (d) what the host lets through: the copies of (b) alone, every rank at the same time, no kernels
    h_in = torch.from_numpy(pin_th[0].array)
    h_out = torch.from_numpy(pin_gr.array)
    d_tmp = [torch.empty((draws, rows.size, N), dtype=torch.float64, device=dev) for _ in range(2)]
    cs = [torch.cuda.Stream(device=dev) for _ in range(2)]

    def copies():
        with torch.cuda.stream(cs[0]):
            d_tmp[0].copy_(h_in, non_blocking=True)
        with torch.cuda.stream(cs[1]):
            h_out.copy_(d_tmp[1], non_blocking=True)

    copies()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        copies()
    torch.cuda.synchronize()
    t_copy = (time.perf_counter() - t0) / e2e_steps
    if world > 1:
        t = torch.tensor([t_e2e, t_sums, t_copy, t_sync, t_async], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_e2e, t_sums, t_copy, t_sync, t_async = (float(x) for x in t.tolist())
    h2d = (draws * rows.size * N + 2 * draws) * 8
    d2h = (draws + draws * rows.size * N + 2 * draws) * 8

    if rank == 0:
        peak = eng.microbench_fp64()
        # configs[1] samples both spectra on one velocity grid, so the library takes its fused
        # emission+absorption kernel: one launch covers B*N*S channel pairs, 32 + 12 canonical flops each
        fused = float(stage[2]) < 0.05 * float(stage[1])
        t_em = float(stage[1]) * 1e-3
        flops_unit = FLOPS_EM + FLOPS_ABS if fused else FLOPS_EM
        # (timed from the end of the cloud map to the end of the channel launches: the EA_TAME grid over every
        # draw plus the EA_FIXUP scan next to it, csrc/moments_ea.cuh MODE)
        kernel_name = ("k_moments_ea<4,false,true,false,true,EA_TAME> (fused emission+absorption; + EA_FIXUP scan)"
                       if fused else "k_moments_em<4,false>")
        draws_launch = draws / pieces  # draws one launch of the kernel processes
        flops_em = draws_launch * N * S_E * flops_unit
        achieved = flops_em / t_em / 1e12
        rec = (3 + 6 * N) if fused else (2 + 5 * N)
        bytes_em = draws_launch * (N * 12 * 8 + rec * 8)  # cc read + moment record written per draw
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        result = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(draws), "units_per_step_per_gpu": units_step,
                       "l2": f"inputs/outputs rotated over {N_SETS} buffer sets ({N_SETS * 65} MB > 126 MB L2)",
                       "parallelism": f"draws sharded over {world} GPU(s), one process per GPU"
                                      + ("; per step the shard's logp/gradient sums (chi_shard_sums_dev) are all-gathered by the library's own NCCL communicator (chi_allgather_dev) on its communication stream, next to the following step's kernels" if world > 1 else "")},
            "clocks": clocks,
            # end to end through the public host API, every step: pinned host inputs -> device, logp + the
            # full gradient computed, and the step's result read back -- logp of every draw and the
            # gradient totals of the shard (C ABI chi_logp_sums_rows_async; what a sharded job gathers and
            # what a survey-total / hierarchical consumer uses).  The per-draw gradient stays on the device.
            "e2e": {"value": world * units_step / t_sums, "unit": UNIT, "ms_per_step": t_sums * 1e3,
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": (draws + 2 + rows.size * N) * 8,
                    "path": "Engine.logp_sums_rows_async + wait (C ABI chi_logp_sums_rows_async / chi_wait), two calls in "
                            "flight on double-buffered pinned host arrays; result = logp[B] + gradient sums [n_rows, N] + "
                            "count of non-finite draws"},
            # the same steps with EVERY per-draw gradient downloaded as well (round 1's definition): 28 MB
            # more per step and rank, which one rank's PCIe link absorbs (N = 1: same speed) and the shared
            # host of an 8-GPU box does not -- `copies_alone` is the same traffic with no kernels at all
            "e2e_full_gradients": {
                "value": world * units_step / t_e2e, "unit": UNIT, "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h, "ms_per_step": t_e2e * 1e3,
                "path": ("Engine.logp_grad_rows_async + wait (C ABI chi_logp_grad_rows_async / chi_wait), "
                         "two calls in flight on double-buffered pinned host arrays"
                         if t_async <= t_sync else
                         "Engine.logp_grad_rows (C ABI chi_logp_grad_rows) with pinned host buffers"),
                "blocking_call_value": world * units_step / t_sync, "blocking_call_ms_per_step": t_sync * 1e3,
                "async_ms_per_step": t_async * 1e3,
                "copies_alone_ms_per_step": t_copy * 1e3, "copies_alone_value": world * units_step / t_copy,
                "frac_of_copy_ceiling": min(1.0, t_copy / t_e2e),
                "note": "every rank at once; on the 8-GPU bench host the ranks share ~150 GB/s (8-11 GB/s each way per "
                        "rank against 55 GB/s for one rank alone, profiles/r2_e2e_diag_n8.txt)"},
            "gpu_launches": launches,
            "value_overlapped_calls": overlapped,
            "stage_ms": {"cloud_map": float(stage[0]), "moments_emission": float(stage[1]),
                         "moments_absorption": float(stage[2]), "epilogue": float(stage[3]),
                         "launches_per_step_of_each_kernel": pieces,
                         "note": "average duration of ONE launch of each kernel (CUDA events on its lane's stream); "
                                 "the pieces of a step run on three lanes and overlap"},
            "roofline": {"kernel": kernel_name, "bound": "fp64", "achieved": achieved, "peak": peak,
                         "unit": "TFLOP/s", "frac": achieved / peak,
                         "peak_source": "own DFMA microbenchmark on this GPU (chi_microbench_fp64); "
                                        "MEASURED_PEAKS.json has no CUDA-core FP64 figure",
                         "flops_per_unit": flops_unit, "units_per_launch": draws_launch * N * S_E,
                         "unit_of_launch": "draw*cloud*channel-pair" if fused else "draw*cloud*channel",
                         # dram__bytes_read+write of this kernel from the ncu --set full capture in
                         # profiles/ (20000 draws per launch: 8.09 MB read, writes absorbed by L2),
                         # scaled to this launch's draws
                         "traffic": (NCU_DRAM_BYTES_PER_DRAW * draws_launch) if fused else None,
                         "traffic_source": "profiles/r2c_ncu_full_summary.txt" if fused else None,
                         "hbm": {"achieved": bytes_em / t_em / 1e9, "peak": hbm_peak, "unit": "GB/s",
                                 "frac": bytes_em / t_em / 1e9 / hbm_peak,
                                 "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback"}},
        }
        if not args.no_cpu_baseline:
            v, s_step, done, cores = time_c_oracle(4096, 3, 1, min_seconds=10.0)
            result["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                                      "sample": f"4096 of the {draws} draws x {done} steps "
                                                f"({s_step * done:.1f} s); {CPU_NOTE}"}
            vt, st_step, dn, ct = time_oracle(256, 2, 1, min_seconds=4.0)
            result["cpu_baseline_torch"] = {"value": vt, "unit": UNIT, "cores": ct, "kind": "port",
                                            "sample": f"256 draws x {dn} steps, torch-CPU fp64 autograd oracle"}
        result["roofline"]["traffic_source"] = "profiles/r2c_ncu_full_summary.txt (ncu --set full, scaled to this launch's draws)" if fused else None
    else:
        result = None
    eng.close()
    # secondary numbers (every BASELINE config, each with its roofline block); never break the bench line.
    # configs[3] runs on all ranks (sightlines sharded), the others on rank 0 while the rest wait.
    if not args.no_extras:
        try:
            extras = other_configs(local, world, rank, peak if rank == 0 else None)
        except Exception as exc:
            extras = {"error": f"{type(exc).__name__}: {exc}"}
        if rank == 0:
            result["other_configs"] = extras
    if rank == 0:
        print(json.dumps(result))
    if world > 1:
        dist.barrier()
    if world > 1:
        dist.destroy_process_group()


def _ess(x):
    """Effective sample size of draws x[n_draws, n_chains] (Geyer's initial positive sequence on the
    chain-averaged autocorrelation, as ArviZ / Stan do without the rank-normalisation)."""
    x = np.asarray(x, dtype=np.float64)
    n, m = x.shape
    if n < 8:
        return float(n * m)
    xc = x - x.mean(axis=0, keepdims=True)
    f = np.fft.rfft(xc, n=2 * n, axis=0)
    acov = np.fft.irfft(f * np.conj(f), axis=0)[:n].real / n
    w = acov[0].mean()
    var_plus = w * (n - 1) / n + (x.mean(axis=0).var(ddof=1) if m > 1 else 0.0)
    rho = 1.0 - (w - acov.mean(axis=1)) / var_plus
    tau = -1.0
    for t in range(0, n - 1, 2):
        pair = rho[t] + rho[t + 1]
        if pair < 0:
            break
        tau += 2.0 * pair
    return float(n * m / max(tau, 1.0 / np.log10(max(n * m, 10))))


def other_configs(local, world, rank, peak64):
    """Secondary numbers beside the headline (not part of `value`), one entry per remaining BASELINE
    config, each with the roofline of its dominant kernel (canonical flops per unit: SURVEY 8d):
      configs[0]  NUTS on EmissionAbsorptionModel, 2 clouds, ~500 channels, 8 chains (rank 0)
      configs[2]  64 chains x 8 clouds x 2048+2048 channels, pseudo-Voigt: one logp+grad evaluation (rank 0)
      configs[3]  survey: 10^4 sightlines x 5 clouds x 1000+1000 channels, 10 draws each; at N > 1 the
                  sightlines are sharded over the ranks (strong scaling), sums gathered by the library's NCCL
      configs[4]  8 clouds x 4096+4096 channels, pseudo-Voigt, spectra only, per-GPU shard, fp64 and fp32 (rank 0)"""
    import torch
    import torch.distributed as dist
    from caribou_hi_b200 import Engine, sharding, synthetic

    out = {}
    dev = torch.device("cuda", local)
    kind = "emission_absorption_physical"

    def engine(N, S, lorentzian=True):
        e = Engine(kind, N, device=local, lorentzian=lorentzian, prior_fwhm_L=1.0 if lorentzian else 0.0, prior_amplitude=1e21)
        v = np.linspace(-100.0, 100.0, S)
        rng = np.random.default_rng(0)
        e.set_spectra(emission=dict(spectral=v, brightness=rng.normal(10, 5, S), noise=0.1, basis=np.full((1, S), 0.1)),
                      absorption=dict(spectral=v, brightness=rng.normal(0.3, 0.1, S), noise=0.01, basis=np.full((1, S), 0.01)))
        return e, rng

    def timed(e, fn, reps):
        stream = torch.cuda.ExternalStream(e.stream, device=dev)
        for _ in range(3):
            fn()
        e.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(reps):
            fn()
        e1.record(stream)
        e.synchronize()
        return e0.elapsed_time(e1) / reps

    def roof(units_per_s, flops_per_unit, peak, unit="TFLOP/s", **kw):
        a = units_per_s * flops_per_unit / 1e12
        return dict({"bound": "fp64", "achieved": a, "peak": peak, "unit": unit, "frac": a / peak if peak else None,
                     "flops_per_unit": flops_per_unit}, **kw)

    # ---- configs[3]: survey, every rank takes part ----
    N, S, n_sl, per = 5, 1000, 10_000, 10
    lo, hi = sharding.shard_bounds(n_sl, world, rank)
    rng = np.random.default_rng(8 + rank)
    v = np.linspace(-100.0, 100.0, S)
    e = Engine(kind, N, device=local, prior_amplitude=1e21)
    e.set_spectra(emission=dict(spectral=v, brightness=rng.normal(10, 5, (hi - lo, S)), noise=0.1),
                  absorption=dict(spectral=v, brightness=rng.normal(0.3, 0.1, (hi - lo, S)), noise=0.01), n_sightlines=hi - lo)
    sharding.init_engine_comm(e)
    B = (hi - lo) * per
    th = lambda trial: e.deterministics(e.pack(trial))["fwhm2_thermal"]
    theta = torch.from_numpy(e.pack(synthetic.draw_free(kind, N, B, rng, thermal_fn=th))).to(dev)
    sl_d = torch.from_numpy(np.repeat(np.arange(hi - lo, dtype=np.int32), per)).to(dev)
    lp = torch.empty(B, dtype=torch.float64, device=dev)
    gr = torch.empty_like(theta)
    sums = torch.empty(10 * N + 2, dtype=torch.float64, device=dev)
    table = torch.empty((world, 10 * N + 2), dtype=torch.float64, device=dev)

    def sweep():
        e.logp_grad_dev(B, theta, lp, gr, sightline=sl_d)
        if world > 1:
            e.shard_sums_dev(B, 10 * N, lp, gr, sums)
            e.allgather_dev(sums, table, 10 * N + 2)
            e.comm_join()

    if world > 1:
        dist.barrier()
    ms = torch.tensor([timed(e, sweep, 20)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item())
    units = n_sl * per * N * 2 * S
    if rank == 0:
        bytes_spectra = n_sl * ((S + 31) // 32) * 8 * 32 * 8  # packed planes of both datasets, read once per sweep
        out["configs[3]: survey, 10^4 sightlines x 5 clouds x 1000+1000 channels x 10 draws, logp+grad"] = {
            "ms_per_sweep": ms, "units_per_s": units / (ms * 1e-3), "n_gpus": world,
            "scaling": "strong (sightlines sharded over the ranks; per-sweep NCCL all-gather of the shard sums by the library)",
            "roofline": roof(units / (ms * 1e-3), 22.0, (peak64 or 0) * world,
                             kernel="k_moments_ea<5,false,true,STAGED> (bulk-async staged channels)",
                             hbm_gb_s=bytes_spectra / (ms * 1e-3) / 1e9)}
    e.close()
    if rank != 0:
        return out

    # ---- configs[2]: latency of one evaluation at 64 chains ----
    N, S, B = 8, 2048, 64
    e, rng = engine(N, S)
    th = lambda trial: e.deterministics(e.pack(trial))["fwhm2_thermal"]
    theta = torch.from_numpy(e.pack(synthetic.draw_free(kind, N, B, rng, lorentzian=True, thermal_fn=th))).to(dev)
    lp = torch.empty(B, dtype=torch.float64, device=dev); gr = torch.empty((B, 10, N), dtype=torch.float64, device=dev)
    ce = torch.zeros((B, 1), dtype=torch.float64, device=dev); ca = torch.zeros_like(ce)
    ge = torch.empty_like(ce); ga = torch.empty_like(ce)
    call = lambda: e.logp_grad_dev(B, theta, lp, gr, coeff_e=ce, coeff_a=ca, gcoeff_e=ge, gcoeff_a=ga)
    res = {}
    stream = torch.cuda.ExternalStream(e.stream, device=dev)
    for label, path in (("pipeline", 1), ("one_launch", 2)):
        e.set_option("small_path", path)
        host_ms = timed(e, call, 200)
        # replayed from a CUDA graph, as the sampler replays its ticks: no host launch cost
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=stream):
            for _ in range(20):
                call()
        with torch.cuda.stream(stream):
            g.replay()
            e.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(10):
                g.replay()
            e1.record(stream)
        e.synchronize()
        res[label] = {"us_host_loop": host_ms * 1e3, "us_in_graph": e0.elapsed_time(e1) / 200 * 1e3}
        del g
    e.set_option("small_path", 0)
    best = min(r["us_in_graph"] for r in res.values())
    units = B * N * 2 * S
    out["configs[2]: 64 chains x 8 clouds x 2048+2048 channels, pseudo-Voigt, logp+grad"] = {
        "us_per_evaluation": best, "units_per_s": units / (best * 1e-6),
        "how": "device time of one evaluation replayed from a CUDA graph (what a sampler tick pays); us_host_loop = "
               "one Python call per evaluation, bound by the host's launch cost",
        "paths": res, "roofline": roof(units / (best * 1e-6), 33.0, peak64, kernel="k_moments_ea<8,true> / k_eval_small<2,true>")}
    e.close()

    # ---- configs[4]: spectra only, per-GPU shape ----
    N, S, B = 8, 4096, 8000
    e, rng = engine(N, S)
    th = lambda trial: e.deterministics(e.pack(trial))["fwhm2_thermal"]
    theta = torch.from_numpy(e.pack(synthetic.draw_free(kind, N, B, rng, lorentzian=True, thermal_fn=th))).to(dev)
    m64 = [torch.empty((B, S), dtype=torch.float64, device=dev) for _ in range(2)]
    m32 = [torch.empty((B, S), dtype=torch.float32, device=dev) for _ in range(2)]
    t64 = timed(e, lambda: e.spectra_dev(B, theta, m64[0], m64[1]), 10)
    t32 = timed(e, lambda: e.spectra_f32_dev(B, theta, m32[0], m32[1]), 10)
    ok = torch.isfinite(m64[0]).all(dim=1)
    err = float(((m32[0][ok].double() - m64[0][ok]).abs() / m64[0][ok].abs().amax(dim=1, keepdim=True)).max())
    units = B * N * 2 * S
    peak32, peak_sfu = e.microbench_fp32()
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm = peaks.get("hbm_gbs", 6650.0)
    pairs32 = B * N * S / (t32 * 1e-3)
    out["configs[4]: 8000 draws x 8 clouds x 4096+4096 channels, pseudo-Voigt, spectra only (per-GPU shard)"] = {
        "fp64_units_per_s": units / (t64 * 1e-3), "fp32_units_per_s": units / (t32 * 1e-3),
        "fp32_worst_error_over_peak_vs_fp64": err,
        "roofline_fp64": roof(units / (t64 * 1e-3), 13.5, peak64, kernel="k_spectra_ea_t<8,true>",
                              hbm={"achieved": B * 2 * S * 8 / (t64 * 1e-3) / 1e9, "peak": hbm, "unit": "GB/s",
                                   "frac": B * 2 * S * 8 / (t64 * 1e-3) / 1e9 / hbm}),
        "roofline_fp32": {"kernel": "k_spectra32_t<8,true,2>", "bound": "sfu",
                          "achieved": 3.0 * pairs32 / 1e12, "peak": peak_sfu, "unit": "1e12 MUFU/s", "frac": 3.0 * pairs32 / 1e12 / peak_sfu,
                          "mufu_per_channel_pair": 3.0,
                          "fp32_fma": {"achieved": 23.0 * pairs32 / 1e12, "peak": peak32, "unit": "TFLOP/s", "frac": 23.0 * pairs32 / 1e12 / peak32},
                          "hbm": {"achieved": B * 2 * S * 4 / (t32 * 1e-3) / 1e9, "peak": hbm, "unit": "GB/s",
                                  "frac": B * 2 * S * 4 / (t32 * 1e-3) / 1e9 / hbm},
                          "peak_source": "own microbenchmarks chi_microbench_fp32 (FFMA chains, MUFU.EX2 chains)"}}
    e.close()

    # ---- configs[0]: the sampler on the reference's own tutorial shape ----
    from caribou_hi_b200 import EmissionAbsorptionModel, SpecData

    rng = np.random.default_rng(1)
    S_e, S_a, chains, tune, draws = 334, 166, 8, 300, 300
    ve, va = np.linspace(-60, 60, S_e), np.linspace(-30, 30, S_a)
    sim = EmissionAbsorptionModel({"emission": SpecData(ve, np.zeros(S_e), 0.1), "absorption": SpecData(va, np.zeros(S_a), 0.01)},
                                  2, baseline_degree=0)
    sim.add_priors(); sim.add_likelihood()
    truth = sim.sample_prior(1, np.random.default_rng(7))
    for _ in range(20):
        mu = sim.predict(truth)
        if all(np.all(np.isfinite(m)) for m in mu.values()):
            break
        truth = sim.sample_prior(1, rng)
    data = {"emission": SpecData(ve, mu["emission"][0] + 0.1 * rng.standard_normal(S_e), 0.1),
            "absorption": SpecData(va, mu["absorption"][0] + 0.01 * rng.standard_normal(S_a), 0.01)}
    model = EmissionAbsorptionModel(data, 2, baseline_degree=0)
    model.add_priors(); model.add_likelihood()
    t0 = time.perf_counter()
    post, stats = model.sample(draws=draws, tune=tune, chains=chains, seed=2, init_point=truth, jitter=0.05)
    dt = time.perf_counter() - t0
    leaps = float((2.0 ** stats["depth"]).sum())  # upper bound on the leapfrog steps of the kept draws
    ess = None
    if "velocity_norm" in post:
        # [chains, draws, clouds]; clouds are exchangeable: the sorted velocities are label-invariant
        vs = np.sort(np.asarray(post["velocity_norm"]).transpose(1, 0, 2), axis=-1)
        ess = min(_ess(vs[:, :, k]) for k in range(vs.shape[-1]))
    units_eval = chains * 2 * (S_e + S_a)
    frac_kept = draws / float(tune + draws)
    out["configs[0]: EmissionAbsorptionModel, 2 clouds, 334+166 channels, NUTS, 8 chains x (300 tune + 300 draws)"] = {
        "iterations_per_s": (tune + draws) / dt, "chain_iterations_per_s": chains * (tune + draws) / dt,
        "mean_tree_depth": float(stats["depth"].mean()), "mean_accept": float(stats["accept"].mean()),
        "divergent_fraction": float(stats["diverging"].mean()), "seconds": dt,
        "ess_per_s_min_over_sorted_velocity": (ess / (dt * frac_kept)) if ess else None,
        "reference_figure": "optimization.ipynb cell 12: 16000 chain-iterations in 28 s = 571 chain-iterations/s (n_cloud = 2, 8 processes, CPU)",
        "roofline": roof(leaps / (dt * frac_kept) * units_eval / chains, 22.0, peak64,
                         note="latency-bound by construction: one tick = the dependent chain leapfrog -> map -> channels -> contraction of 8 chains")}
    # the same model with more chains: a tick serves every chain's leapfrog step at once, so chains are
    # almost free until the machine fills -- the device sampler's operating point is hundreds of chains
    more = {}
    for c in (64, 512):
        t0 = time.perf_counter()
        _, st = model.sample(draws=100, tune=200, chains=c, seed=3, init_point=truth, jitter=0.05)
        dt_c = time.perf_counter() - t0
        more[str(c)] = {"iterations_per_s": 300 / dt_c, "chain_iterations_per_s": c * 300 / dt_c,
                        "mean_tree_depth": float(st["depth"].mean()), "divergent_fraction": float(st["diverging"].mean())}
    key0 = [k for k in out if k.startswith("configs[0]")][0]
    out[key0]["more_chains_200_tune_100_draws"] = more
    out[key0]["how"] = ("chi_nuts_sample through Model.sample(): one resident thread block per chain runs whole leapfrog "
                        "steps inside one kernel (csrc/nuts_chain.cuh)")
    model.engine.close()

    # ---- the sampler at configs[2]'s size: 64 chains x 8 clouds x 2048+2048 channels, pseudo-Voigt ----
    from caribou_hi_b200 import EmissionAbsorptionPhysicalModel

    S2, N2, chains2, tune2, draws2 = 2048, 8, 64, 200, 100
    v2 = np.linspace(-60, 60, S2)
    zero = {"emission": SpecData(v2, np.zeros(S2), 0.1), "absorption": SpecData(v2, np.zeros(S2), 0.01)}
    sim = EmissionAbsorptionPhysicalModel(zero, N2, baseline_degree=0)
    sim.add_priors(prior_fwhm_L=1.0); sim.add_likelihood()
    truth = sim.sample_prior(1, np.random.default_rng(7))
    for _ in range(20):
        mu = sim.predict(truth)
        if all(np.all(np.isfinite(m)) for m in mu.values()):
            break
        truth = sim.sample_prior(1, rng)
    sim.engine.close()
    data = {"emission": SpecData(v2, mu["emission"][0] + 0.1 * rng.standard_normal(S2), 0.1),
            "absorption": SpecData(v2, mu["absorption"][0] + 0.01 * rng.standard_normal(S2), 0.01)}
    model = EmissionAbsorptionPhysicalModel(data, N2, baseline_degree=0)
    model.add_priors(prior_fwhm_L=1.0); model.add_likelihood()
    t0 = time.perf_counter()
    _, st = model.sample(draws=draws2, tune=tune2, chains=chains2, seed=2, init_point=truth, jitter=0.05)
    dt2 = time.perf_counter() - t0
    key2 = [k for k in out if k.startswith("configs[2]")][0]
    out[key2]["nuts_64_chains_200_tune_100_draws"] = {
        "iterations_per_s": (tune2 + draws2) / dt2, "chain_iterations_per_s": chains2 * (tune2 + draws2) / dt2,
        "mean_tree_depth": float(st["depth"].mean()), "divergent_fraction": float(st["diverging"].mean()), "seconds": dt2,
        "how": "the four blocks of a chain as one thread-block cluster, whole leapfrog steps inside one kernel "
               "(csrc/nuts_chain.cuh, k_nuts_chain_cl)"}
    model.engine.close()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--draws", type=int, default=DRAWS)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the secondary configs[2] / configs[4] timings")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        args.warmup = max(args.warmup, 3)
        run_b200(args)


if __name__ == "__main__":
    main()
