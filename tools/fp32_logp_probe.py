"""Why logp + gradient have no fp32 form: the emission likelihood of
physics.py:294-309 + physics.py:345-365 + pm.Normal evaluated with torch-CPU autograd in float32
and in float64 on the same inputs, (a) at prior draws (model far from the data) and (b) next to the
parameters that generated the data (where a sampler spends its time).  CPU only, no GPU needed:

    python tools/fp32_logp_probe.py

Measured here: (a) logp 3e-7, gradient 6e-7 relative -- inside the north star's 1e-5 bar;
(b) logp 3e-5, gradient 1e-4 -- outside it: the residual y - mu is a difference of numbers
that agree to 3 digits (mu ~ 50 K, sigma = 0.1 K).  Doing the residual and the sums in float64
after a float32 forward model ("mixed") changes nothing: the error is already in mu."""
import numpy as np
import torch

torch.manual_seed(0)
B, N, S = 64, 4, 1000
BG = 3.77


def brightness(v, c, W2, tau, Ts, f):
    d = v[None, :, None] - c[:, None, :]
    k = 4 * np.log(2) / W2
    prof = torch.exp(-k[:, None, :] * d * d) * torch.sqrt(k / np.pi)[:, None, :]
    e = torch.exp(-tau[:, None, :] * prof)
    a = 1 - f[:, None, :] + f[:, None, :] * e
    A = torch.cumprod(a, 2)
    A = torch.cat([torch.ones_like(A[:, :, :1]), A], 2)
    return BG * A[:, :, -1] + (f[:, None, :] * Ts[:, None, :] * (1 - e) * A[:, :, :-1]).sum(2) - BG


def logp(v, params, y, sig, mixed):
    mu = brightness(v, *params)
    if mixed:
        mu, y, sig = mu.double(), y.double(), sig.double()
    r = (y[None] - mu) / sig
    return (-0.5 * r * r).sum(1)


def draw():
    f64 = dict(dtype=torch.float64)
    return [torch.randn(B, N, **f64) * 10, 500 * torch.randn(B, N, **f64) ** 2 + 1, torch.rand(B, N, **f64) * 3,
            10 ** (1.5 + torch.rand(B, N, **f64) * 2), torch.rand(B, N, **f64)]


v64 = torch.linspace(-100, 100, S, dtype=torch.float64)
truth = [p[:1] for p in draw()]
y = brightness(v64, *truth)[0] + 0.1 * torch.randn(S, dtype=torch.float64)
sig = torch.full((S,), 0.1, dtype=torch.float64)
near = [t.expand(B, N) * (1 + 1e-3 * torch.randn(B, N, dtype=torch.float64)) for t in truth]
for label, params in (("prior draws", draw()), ("next to the generating parameters", near)):
    for mixed in (False, True):
        res = {}
        for dt in (torch.float64, torch.float32):
            p = [x.clone().to(dt).requires_grad_() for x in params]
            lp = logp(v64.to(dt), p, y.to(dt), sig.to(dt), mixed and dt == torch.float32)
            lp.sum().backward()
            res[dt] = (lp.detach().double(), torch.stack([x.grad.double() for x in p], 1))
        (lp64, g64), (lp32, g32) = res[torch.float64], res[torch.float32]
        e_lp = ((lp32 - lp64).abs() / lp64.abs()).max().item()
        scale = g64.abs().amax(dim=(1, 2), keepdim=True)
        e_g = ((g32 - g64).abs() / (g64.abs() + scale)).max().item()
        print(f"{label:36s} {'fp32 model, fp64 residual' if mixed else 'all fp32':26s} worst relative error: "
              f"logp {e_lp:.1e}  gradient {e_g:.1e}  (median logp {lp64.median().item():.3g})")
