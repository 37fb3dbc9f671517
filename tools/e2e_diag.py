"""Where does the end-to-end step go when N ranks share one host?  Under torchrun, every rank at the same
time: H2D alone, D2H alone, both, the compute-only step and the full host-to-host step of bench.py's e2e
(27.2 MB up, 28 MB down per rank at configs[1]); with the NUMA node of the GPU, of the pinned pages and the
CPUs the rank may run on.  usage: torchrun --nproc-per-node N tools/e2e_diag.py"""
import ctypes, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
from caribou_hi_b200 import Engine, PinnedArray
import bench

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("gloo")
bar = (lambda: dist.barrier()) if world > 1 else (lambda: None)


def page_nodes(arr):
    """NUMA node of a sample of the pages of a numpy array (move_pages with nodes = NULL)."""
    libc = ctypes.CDLL(None, use_errno=True)
    page = 4096
    addr = arr.ctypes.data & ~(page - 1)
    n = min(64, max(1, arr.nbytes // page))
    step = max(1, (arr.nbytes // page) // n)
    pages = (ctypes.c_void_p * n)(*[addr + i * step * page for i in range(n)])
    status = (ctypes.c_int * n)()
    rc = libc.syscall(279, 0, ctypes.c_ulong(n), pages, None, status, 0)
    if rc != 0:
        return f"move_pages errno {ctypes.get_errno()}"
    vals, counts = np.unique(np.array(status[:]), return_counts=True)
    return {int(v): int(c) for v, c in zip(vals, counts)}


props = torch.cuda.get_device_properties(local)
bus = f"{props.pci_domain_id:04x}:{props.pci_bus_id:02x}:{props.pci_device_id:02x}.0"
try:
    gpu_node = open(f"/sys/bus/pci/devices/{bus}/numa_node").read().strip()
except Exception as exc:
    gpu_node = f"? ({exc})"
N, draws = bench.N_CLOUDS, bench.DRAWS
eng = Engine(bench.KIND, N, device=local, prior_amplitude=1.0e21)
thetas = bench.build_inputs(eng, draws, rank, 2)
rows = eng.row_slots
pin_in = [PinnedArray((draws, rows.size, N)) for _ in range(2)]
for k in range(2):
    pin_in[k].array[...] = thetas[k][:, rows, :]
pin_out = [{"logp": PinnedArray((draws,)), "grad": PinnedArray((draws, rows.size, N)), "gcoeff_e": PinnedArray((draws, 1)),
            "gcoeff_a": PinnedArray((draws, 1))} for _ in range(2)]
outs = [{k: v.array for k, v in o.items()} for o in pin_out]
ce, ca = PinnedArray((draws, 1)), PinnedArray((draws, 1))
ce.array[...] = 0.0; ca.array[...] = 0.0
dev = torch.device("cuda", local)
d_in = torch.empty((draws, rows.size, N), dtype=torch.float64, device=dev)
d_out = torch.empty((draws, rows.size, N), dtype=torch.float64, device=dev)
h_in = torch.from_numpy(pin_in[0].array); h_out = torch.from_numpy(pin_out[0]["grad"].array)
info = (f"rank {rank}: gpu {local} bus {bus} numa_node {gpu_node}; cpus allowed {len(os.sched_getaffinity(0))} "
        f"[{min(os.sched_getaffinity(0))}..{max(os.sched_getaffinity(0))}]; pinned input pages on nodes {page_nodes(pin_in[0].array)}, "
        f"output pages {page_nodes(pin_out[0]['grad'].array)}")


def timed(fn, reps=20):
    fn(); torch.cuda.synchronize(); bar()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps
    bar()
    return dt

s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
mb = d_in.numel() * 8 / 1e6
def up():
    with torch.cuda.stream(s1): d_in.copy_(h_in, non_blocking=True)
def down():
    with torch.cuda.stream(s2): h_out.copy_(d_out, non_blocking=True)
t_up = timed(up); t_dn = timed(down); t_both = timed(lambda: (up(), down()))
eng.logp_grad_rows(pin_in[0].array, ce.array, ca.array, out=outs[0])
t_sync = timed(lambda: eng.logp_grad_rows(pin_in[0].array, ce.array, ca.array, out=outs[0]), 10)
def async_pair():
    a = eng.logp_grad_rows_async(pin_in[0].array, ce.array, ca.array, outs[0])
    b = eng.logp_grad_rows_async(pin_in[1].array, ce.array, ca.array, outs[1])
    eng.wait(a); eng.wait(b)
t_async = timed(async_pair, 10) / 2
lines = [info, f"rank {rank}: H2D alone {mb / t_up / 1e3:.1f} GB/s, D2H alone {mb / t_dn / 1e3:.1f} GB/s, both at once "
               f"{mb / t_both / 1e3:.1f} GB/s each way; step host-to-host blocking {t_sync * 1e3:.2f} ms, async (2 in flight) {t_async * 1e3:.2f} ms"]
if world > 1:
    allv = [None] * world
    dist.all_gather_object(allv, lines)
    if rank == 0:
        for l in allv:
            print("\n".join(l))
else:
    print("\n".join(lines))
eng.close()
if world > 1:
    dist.destroy_process_group()
