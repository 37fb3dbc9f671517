import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    # the oracle evaluates prior draws the reference itself turns into NaN (log10 of 0, 0/0)
    config.addinivalue_line("filterwarnings", "ignore::RuntimeWarning")
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu under gpurun)")


@pytest.fixture(scope="session")
def cuda_device():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return 0


def pytest_terminal_summary(terminalreporter):
    """How often the 60-digit arbiter was needed (helpers.assert_grad_close) and how far the two
    contestants were from it."""
    try:
        from helpers import ARBITRATIONS as a
    except Exception:
        return
    if a["count"]:
        terminalreporter.write_line(
            f"gradient components arbitrated with mpmath: {a['count']} (worst relative distance to the 60-digit value: "
            f"cuda {a['worst_cuda']:.2e}, fp64 autograd {a['worst_autograd']:.2e})")
        import numpy as np
        pr = np.array(a["pairs"])
        ratio = pr[:, 0] / np.maximum(pr[:, 1], 1e-300)
        terminalreporter.write_line(
            f"  cuda error / autograd error over those components: median {np.median(ratio):.2f}, "
            f"90th percentile {np.percentile(ratio, 90):.2f}, max {ratio.max():.3g}; cuda closer in {np.mean(ratio < 1):.0%}")
        if a.get("noise"):
            nz = np.array(a["noise"])
            for who, col in (("cuda", 0), ("autograd", 1)):
                q = np.percentile(nz[:, col], [50, 90, 99, 100])
                terminalreporter.write_line(f"  {who} error / fp64 autograd's noise on the component (max over 4 one-ulp neighbours): "
                                            f"median {q[0]:.3g}, 90% {q[1]:.3g}, 99% {q[2]:.3g}, max {q[3]:.3g}")
        if a.get("cond"):
            q = np.percentile(np.array(a["cond"]), [50, 90, 99, 100])
            terminalreporter.write_line(f"  cuda error / magnitude of the chain-rule terms of the component: median {q[0]:.3g}, "
                                        f"90% {q[1]:.3g}, 99% {q[2]:.3g}, max {q[3]:.3g} (bar: 1e-10)")
        if a.get("detail") and os.environ.get("CHI_ARB_STATS"):
            for d in sorted(a["detail"], reverse=True)[:40]:
                terminalreporter.write_line("  ratio %.3g | %s | %s[%d,%d] exact %.6g e_cuda %.3g e_auto %.3g scale %.3g" % d)
