"""The symbolic half of the PyTensor Ops on the GPU: tests/graph_worker.py in a subprocess (it imports the
stand-in of tests/golden/graph_shim as `pytensor`)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_graphs_with_cuda_ops(cuda_device):
    """Fifteen graphs traced from the reference's unmodified classes (twelve with the fused HISpectraOp
    patch of INTEGRATION.md section 3, three with caribou_hi.physics swapped for caribou_hi_b200.physics),
    plus HILogLikeOp: logp, predicted spectra and gradients against tests/golden/ref_*.npz at 1e-10."""
    res = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "graph_worker.py")], capture_output=True, text=True,
                         timeout=1200, cwd=ROOT)
    assert res.returncode == 0 and "GRAPH_WORKER_OK 15" in res.stdout, res.stdout[-3000:] + res.stderr[-4000:]
