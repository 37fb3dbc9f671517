"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle on the same
seeded inputs.  Tolerance: 1e-10 relative (fp64 path, BASELINE.json north_star), with
atol = rtol * max|ref| over each output vector."""

import numpy as np
import pytest

from helpers import RTOL64, assert_close, assert_grad_close, engine_for, make_case, oracle_logp_grad
from oracle import models_ref as mr
from oracle import physics_ref as ph

pytestmark = pytest.mark.gpu

KINDS = list(mr.KINDS)


def _check_case(kind, N, S_e, S_a, B, seed, lorentzian, baseline_degree=1, weight=True, **kw):
    spec, data, free, baseline = make_case(kind, N, S_e, S_a, B, seed, lorentzian, baseline_degree, weight, **kw)
    ref_lp, ref_g = oracle_logp_grad(spec, data, free, baseline)
    eng = engine_for(spec, data, baseline_degree)
    theta = eng.pack(free)
    finite = np.isfinite(ref_lp)
    # both evaluation paths of a small batch: the map -> channel kernels -> contraction pipeline
    # (small_path 1) and the one-launch kernel of csrc/small.cuh (small_path 2)
    for path in (1, 2):
        eng.set_option("small_path", path)
        lp, grad, gce, gca = eng.logp_grad(theta, baseline.get("baseline_emission_norm"),
                                           baseline.get("baseline_absorption_norm"))
        assert_close(lp, ref_lp, what=f"{kind} logp (path {path})", axis=None)
        # per RV vector, 60-digit arbitration where a component misses 1e-10 against autograd
        got = dict(eng.unpack_grad(grad))
        if gce is not None:
            got["baseline_emission_norm"] = gce
        if gca is not None:
            got["baseline_absorption_norm"] = gca
        assert_grad_close(got, ref_g, finite, case=(spec, data, free, baseline), what=f"{kind} (path {path})")
    eng.set_option("small_path", 0)
    # predicted spectra
    ref_mu = mr.predict(spec, data, free, baseline)
    mu_e, mu_a = eng.spectra(theta, baseline.get("baseline_emission_norm"),
                             baseline.get("baseline_absorption_norm"))
    if mu_e is not None:
        assert_close(mu_e, ref_mu["emission"], what=f"{kind} mu_e")
    if mu_a is not None:
        assert_close(mu_a, ref_mu["absorption"], what=f"{kind} mu_a")
    eng.close()
    return int(finite.sum())


@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("lorentzian", [False, True])
def test_logp_grad_spectra_all_models(cuda_device, kind, lorentzian):
    n_ok = _check_case(kind, 3, 200, 100, 64, seed=11, lorentzian=lorentzian)
    assert n_ok >= 32  # most prior draws must be finite, otherwise the test is vacuous


@pytest.mark.parametrize("N", [1, 2, 4, 5, 8, 9, 12, 13, 16])
def test_cloud_counts(cuda_device, N):
    """N > 8 runs on the 12- / 16-cloud instantiations padded with null clouds."""
    _check_case("emission_absorption_physical", N, 257, 131, 16, seed=100 + N, lorentzian=(N % 2 == 0))


@pytest.mark.parametrize("kind,N,shared", [("emission_absorption", 11, True), ("emission", 10, False),
                                           ("absorption_physical", 14, False), ("emission_absorption_physical", 16, True)])
def test_many_clouds_other_paths(cuda_device, kind, N, shared):
    """Padded cloud slots on the fused kernel, the single-dataset kernels and the model-level
    (prior) epilogue."""
    from oracle import priors_ref as pr

    _check_case(kind, N, 200, 200, 12, seed=300 + N, lorentzian=(N % 2 == 1), shared_axis=shared)


@pytest.mark.parametrize("kind", ["emission_absorption", "emission_absorption_physical"])
@pytest.mark.parametrize("lorentzian", [False, True])
@pytest.mark.parametrize("N", [1, 4, 8])
def test_shared_axis_fused_kernel(cuda_device, kind, lorentzian, N):
    """Emission and absorption on the same velocity grid take the fused kernel
    (moments_ea.cuh: one exp(-k d^2) for both datasets, combined moments)."""
    _check_case(kind, N, 333, 333, 24, seed=300 + N, lorentzian=lorentzian, baseline_degree=2, shared_axis=True)


def test_shared_axis_small_batch_and_vjp(cuda_device):
    """Fused kernel with channel chunking (small batch) and in VJP mode."""
    import torch
    from caribou_hi_b200 import Engine

    _check_case("emission_absorption_physical", 5, 1500, 1500, 4, seed=77, lorentzian=True, shared_axis=True,
                half_width=100.0)
    rng = np.random.default_rng(8)
    B, N, S = 6, 3, 200
    v = np.linspace(-50, 50, S)
    inp = {"velocity": rng.normal(0, 8, (B, N)), "fwhm2": rng.uniform(5, 200, (B, N)),
           "fwhm_L": rng.uniform(0.0, 3.0, (B, N)), "tau_total": rng.uniform(0.05, 8, (B, N)),
           "tspin": rng.uniform(30, 3000, (B, N)), "filling_factor": rng.uniform(0.1, 1.0, (B, N)),
           "absorption_weight": rng.uniform(0.3, 1.2, (B, N))}
    data = {"emission": mr.OracleData(v, rng.normal(5, 3, S), 0.3), "absorption": mr.OracleData(v, rng.normal(0.2, 0.1, S), 0.02)}
    spec = mr.make_spec("emission_absorption", N)
    eng = Engine("physics", N, lorentzian=True)
    eng.set_spectra(emission=dict(spectral=v, brightness=data["emission"].brightness, noise=0.3),
                    absorption=dict(spectral=v, brightness=data["absorption"].brightness, noise=0.02))
    theta = eng.pack(inp)
    for ge, ga in ((rng.standard_normal((B, S)), rng.standard_normal((B, S))), (None, rng.standard_normal((B, S))),
                   (rng.standard_normal((B, S)), None)):
        it = {k: torch.tensor(x, requires_grad=True) for k, x in inp.items()}
        mu = mr.predict_from_inputs(spec, data, it)
        tot = 0.0
        if ge is not None:
            tot = tot + (mu["emission"] * torch.tensor(ge)).sum()
        if ga is not None:
            tot = tot + (mu["absorption"] * torch.tensor(ga)).sum()
        tot.backward()
        vg, _, _ = eng.spectra_vjp(theta, ge, ga)
        got = eng.unpack_grad(vg)
        names = sorted(got)
        ref = [it[k].grad.numpy() if it[k].grad is not None else np.zeros((B, N)) for k in names]
        assert_close(np.concatenate([got[k] for k in names], axis=1), np.concatenate(ref, axis=1), what="fused vjp")
    eng.close()


def test_emission_absorption_fixed_weight(cuda_device):
    """prior_sigma_log10_NHI=None: absorption_weight is the constant 1
    (emission_absorption_model.py:46-47)."""
    _check_case("emission_absorption", 2, 334, 166, 32, seed=5, lorentzian=False, weight=False)


def test_config1_shape(cuda_device):
    """BASELINE configs[0]: EmissionAbsorptionModel, 2 clouds, ~500 channels."""
    _check_case("emission_absorption", 2, 334, 166, 8, seed=1234, lorentzian=False, baseline_degree=0)


def test_config2_shape_sample(cuda_device):
    """BASELINE configs[1] shape (4 clouds, 1000+1000 channels) on a sample of draws."""
    _check_case("emission_absorption_physical", 4, 1000, 1000, 48, seed=2, lorentzian=False,
                baseline_degree=0, half_width=100.0)


def test_config3_shape_small_batch_is_chunked(cuda_device):
    """BASELINE configs[2]: pseudo-Voigt, 8 clouds, 2048 channels, 64 chains -- small B
    exercises the channel-chunked path (several warps per draw)."""
    _check_case("emission_absorption_physical", 8, 2048, 2048, 8, seed=3, lorentzian=True,
                baseline_degree=0, half_width=100.0)


def test_nondefault_hyperparameters(cuda_device):
    _check_case("emission_absorption_physical", 3, 150, 90, 16, seed=9, lorentzian=True,
                bg_temp=2.7, depth_nth_fwhm_power=0.4, prior_fwhm2=200.0, prior_velocity=[2.0, 5.0],
                prior_log10_n_alpha=[-5.0, 1.0], prior_ff_NHI=5.0e20, prior_fwhm_L=2.0)
    _check_case("emission_absorption", 3, 150, 90, 16, seed=10, lorentzian=True,
                bg_temp=2.7, prior_log10_nHI=[0.5, 1.0], prior_TB_fwhm=300.0)


def test_ragged_channel_counts(cuda_device):
    """Channel counts that are not multiples of the warp size, minimum 2 channels."""
    _check_case("emission_absorption", 2, 33, 2, 8, seed=21, lorentzian=True)
    _check_case("absorption", 1, 0, 31, 8, seed=22, lorentzian=False)
    _check_case("emission", 2, 2, 0, 8, seed=23, lorentzian=False)


def test_physics_level_logp_and_vjp(cuda_device):
    """MODEL_PHYSICS: the seven likelihood inputs directly (the Op boundary)."""
    import torch
    from caribou_hi_b200 import Engine

    rng = np.random.default_rng(7)
    B, N, S_e, S_a = 24, 3, 180, 90
    ve = np.linspace(-50, 50, S_e)
    va = np.linspace(-25, 25, S_a)
    inp = {
        "velocity": rng.normal(0, 8, (B, N)),
        "fwhm2": rng.uniform(5, 200, (B, N)),
        "fwhm_L": rng.uniform(0.0, 3.0, (B, N)),
        "tau_total": rng.uniform(0.05, 8, (B, N)),
        "tspin": rng.uniform(30, 3000, (B, N)),
        "filling_factor": rng.uniform(0.1, 1.0, (B, N)),
        "absorption_weight": rng.uniform(0.3, 1.2, (B, N)),
    }
    ye = rng.normal(5, 3, S_e)
    ya = rng.normal(0.2, 0.1, S_a)
    data = {"emission": mr.OracleData(ve, ye, 0.3), "absorption": mr.OracleData(va, ya, 0.02)}
    spec = mr.make_spec("emission_absorption", N, bg_temp=3.77)

    # oracle on the deterministic inputs directly
    it = {k: torch.tensor(v, requires_grad=True) for k, v in inp.items()}
    lp = mr.logp_from_inputs(spec, data, it)
    lp.sum().backward()
    eng = Engine("physics", N, lorentzian=True, bg_temp=3.77)
    eng.set_spectra(
        emission=dict(spectral=ve, brightness=ye, noise=0.3, offset=data["emission"]._brightness_offset),
        absorption=dict(spectral=va, brightness=ya, noise=0.02, offset=data["absorption"]._brightness_offset),
    )
    theta = eng.pack(inp)
    got_lp, grad, _, _ = eng.logp_grad(theta)
    assert_close(got_lp, lp.detach().numpy(), what="physics logp", axis=None)
    for name, g in eng.unpack_grad(grad).items():
        assert_close(g, it[name].grad.numpy(), what=f"physics d/d{name}")

    # spectra VJP against autograd with random cotangents
    ge = rng.standard_normal((B, S_e))
    ga = rng.standard_normal((B, S_a))
    it = {k: torch.tensor(v, requires_grad=True) for k, v in inp.items()}
    mu = mr.predict_from_inputs(spec, data, it)
    ((mu["emission"] * torch.tensor(ge)).sum() + (mu["absorption"] * torch.tensor(ga)).sum()).backward()
    vg, _, _ = eng.spectra_vjp(theta, ge, ga)
    for name, g in eng.unpack_grad(vg).items():
        assert_close(g, it[name].grad.numpy(), what=f"vjp d/d{name}")
    # emission-only cotangent
    it = {k: torch.tensor(v, requires_grad=True) for k, v in inp.items()}
    mu = mr.predict_from_inputs(spec, data, it)
    (mu["emission"] * torch.tensor(ge)).sum().backward()
    vg, _, _ = eng.spectra_vjp(theta, ge, None)
    for name, g in eng.unpack_grad(vg).items():
        r = it[name].grad.numpy() if it[name].grad is not None else np.zeros((B, N))
        assert_close(g, r, what=f"vjp(e only) d/d{name}")
    eng.close()


def test_multiple_sightlines(cuda_device):
    """BASELINE configs[3] in miniature: each draw is scored against its own sightline."""
    spec, data, free, baseline = make_case("emission_physical", 5, 300, 0, 40, seed=31)
    rng = np.random.default_rng(32)
    n_sl = 7
    ax = data["emission"].spectral
    ys = data["emission"].brightness[None, :] + 0.5 * rng.standard_normal((n_sl, ax.size))
    sig = rng.uniform(0.05, 0.2, (n_sl, ax.size))
    sl = rng.integers(0, n_sl, size=40).astype(np.int32)
    from caribou_hi_b200 import Engine

    eng = Engine("emission_physical", 5, prior_amplitude=spec["prior_ff_NHI"])
    eng.set_spectra(emission=dict(spectral=ax, brightness=ys, noise=sig), n_sightlines=n_sl)
    theta = eng.pack(free)
    lp, grad, _, _ = eng.logp_grad(theta, sightline=sl)
    got = eng.unpack_grad(grad)
    for i in range(n_sl):
        sel = np.where(sl == i)[0]
        if sel.size == 0:
            continue
        d = {"emission": mr.OracleData(ax, ys[i], sig[i])}
        d["emission"]._brightness_offset = 0.0  # no baseline in this engine
        sub = {k: v[sel] for k, v in free.items()}
        ref_lp, ref_g = oracle_logp_grad(spec, d, sub, {})
        fin = np.isfinite(ref_lp)
        assert_close(lp[sel], ref_lp, what="sightline logp", axis=None)
        for name in got:
            assert_close(got[name][sel][fin], ref_g[name][fin], what=f"sightline d/d{name}")
    with pytest.raises(Exception):
        eng.logp_grad(theta, sightline=np.full(40, n_sl, dtype=np.int32))
    eng.close()


def test_deterministics(cuda_device):
    for kind in KINDS:
        spec, data, free, baseline = make_case(kind, 3, 64, 64, 32, seed=41, lorentzian=True)
        eng = engine_for(spec, data)
        det = eng.deterministics(eng.pack(free))
        ref = mr.cloud_inputs(spec, free)
        for name, r in ref.items():
            if name in det:
                r = np.broadcast_to(r, det[name].shape)
                assert_close(det[name], r, what=f"{kind} det {name}", rtol=1e-12)
        eng.close()


def test_physics_functions(cuda_device):
    from caribou_hi_b200 import Engine

    eng = Engine("physics", 2)
    rng = np.random.default_rng(3)
    tk = np.concatenate([rng.uniform(10, 299, 500), rng.uniform(300, 1e4, 500), [300.0, 299.99999, 1000.0, 8000.0]])
    dens = 10 ** rng.uniform(-2, 3, tk.size)
    na = 10 ** rng.uniform(-9, -3, tk.size)
    ts = eng.spin_temp(tk, dens, na)
    assert_close(ts, ph.calc_spin_temp(tk, dens, na), rtol=1e-13, what="spin_temp")
    import torch

    t = [torch.tensor(x, requires_grad=True) for x in (tk, dens, na)]
    gbar = rng.standard_normal(tk.size)
    (ph.calc_spin_temp(*t) * torch.tensor(gbar)).sum().backward()
    g = eng.spin_temp_vjp(tk, dens, na, gbar)
    for got, ref, nm in zip(g, t, ("tk", "density", "n_alpha")):
        assert_close(got, ref.grad.numpy(), rtol=1e-11, what=f"spin_temp vjp {nm}", axis=None)

    v = np.linspace(-100, 100, 101)
    prof = eng.pseudo_voigt(v, np.array([-5.0, 5.0]), np.array([100.0, 200.0]), 0.0)
    assert prof.shape == (101, 2) and not np.any(np.isnan(prof))
    assert_close(prof, ph.calc_pseudo_voigt(v, np.array([-5.0, 5.0]), np.array([100.0, 200.0]), np.zeros(2)),
                 what="pseudo_voigt gaussian", axis=0)
    fl = np.array([1.5, 4.0])
    prof = eng.pseudo_voigt(v, np.array([-5.0, 5.0]), np.array([100.0, 200.0]), fl)
    assert_close(prof, ph.calc_pseudo_voigt(v, np.array([-5.0, 5.0]), np.array([100.0, 200.0]), fl),
                 what="pseudo_voigt", axis=0)
    tau = np.array([5.0, 3.0]) * prof
    tb = eng.radiative_transfer(tau, np.array([1000.0, 5000.0]), 1.0, 2.7)
    assert tb.shape == (101,) and not np.any(np.isnan(tb))
    assert_close(tb, ph.radiative_transfer(tau, np.array([1000.0, 5000.0]), 1.0, 2.7), what="radiative_transfer")
    eng.close()


def test_error_behaviour(cuda_device):
    from caribou_hi_b200 import Engine
    from caribou_hi_b200._lib import ChiError

    with pytest.raises(ChiError):
        Engine("emission", 17)  # > CHI_MAX_CLOUDS
    eng = Engine("emission", 2)
    with pytest.raises(ChiError):  # spectra not set
        eng.logp_grad(np.zeros((1, 10, 2)))
    with pytest.raises(ChiError):  # one channel: physics.py:294 needs two
        eng.set_spectra(emission=dict(spectral=np.zeros(1), brightness=np.zeros(1), noise=1.0))
    with pytest.raises(ChiError):  # missing dataset
        eng.set_spectra(absorption=dict(spectral=np.arange(4.0), brightness=np.zeros(4), noise=1.0))
    eng.set_spectra(emission=dict(spectral=np.arange(4.0), brightness=np.zeros(4), noise=1.0))
    lp, g, _, _ = eng.logp_grad(np.zeros((0, 10, 2)))  # empty batch
    assert lp.shape == (0,)
    eng.close()


def test_descending_and_nonuniform_axes(cuda_device):
    """The spectral axis only has to have >= 2 channels (physics.py:294 uses the first two for the
    channel width); descending and non-uniform axes must work like in the reference."""
    from caribou_hi_b200 import Engine

    rng = np.random.default_rng(12)
    spec, data, free, baseline = make_case("emission_absorption", 3, 120, 80, 10, seed=55, lorentzian=True)
    for variant in ("descending", "nonuniform"):
        d2 = {}
        for key, d in data.items():
            ax = d.spectral[::-1].copy() if variant == "descending" else np.sort(
                d.spectral + rng.uniform(-0.2, 0.2, d.spectral.size))
            y = d.brightness[::-1].copy() if variant == "descending" else d.brightness
            d2[key] = mr.OracleData(ax, y, d.noise)
        ref_lp, ref_g = oracle_logp_grad(spec, d2, free, baseline)
        eng = engine_for(spec, d2, 0)
        lp, grad, gce, gca = eng.logp_grad(eng.pack(free), baseline["baseline_emission_norm"],
                                           baseline["baseline_absorption_norm"])
        assert_close(lp, ref_lp, what=f"{variant} logp", axis=None)
        got = eng.unpack_grad(grad)
        names = sorted(got)
        fin = np.isfinite(ref_lp)
        assert_close(np.concatenate([got[n][fin] for n in names], axis=1),
                     np.concatenate([ref_g[n][fin] for n in names], axis=1), what=f"{variant} grad")
        eng.close()


def test_host_sub_batching_and_lanes(cuda_device):
    """A batch large enough for the host entry point to cut it into pipelined sub-batches must
    give the same numbers as the same draws evaluated in small calls."""
    spec, data, free, baseline = make_case("emission_physical", 2, 64, 0, 30000, seed=99)
    eng = engine_for(spec, data, 0)
    theta = eng.pack(free)
    ce = baseline["baseline_emission_norm"]
    lp, grad, gce, _ = eng.logp_grad(theta, ce)
    idx = np.r_[0:50, 14990:15040, 29950:30000]
    lp2, grad2, gce2, _ = eng.logp_grad(theta[idx], ce[idx])
    # a small call takes the split epilogue (another summation order): equal to rounding
    assert np.array_equal(np.isfinite(lp[idx]), np.isfinite(lp2))
    fin = np.isfinite(lp2)
    assert_close(lp[idx], lp2, rtol=1e-13, what="logp", axis=None)
    assert_close(grad[idx].reshape(idx.size, -1)[fin], grad2.reshape(idx.size, -1)[fin], rtol=1e-12, what="grad")
    assert_close(gce[idx][fin], gce2[fin], rtol=1e-12, what="gcoeff", axis=None)
    # the same draws at another position of an equally large call: bit for bit
    perm = np.random.default_rng(0).permutation(theta.shape[0])
    lp3, grad3, gce3, _ = eng.logp_grad(theta[perm], ce[perm])
    assert np.array_equal(lp3, lp[perm], equal_nan=True)
    assert np.array_equal(grad3, grad[perm], equal_nan=True)
    assert np.array_equal(gce3, gce[perm], equal_nan=True)
    mu_e, _ = eng.spectra(theta[:2000], ce[:2000])
    mu_e2, _ = eng.spectra(theta[1990:2000], ce[1990:2000])
    assert np.array_equal(mu_e[1990:2000], mu_e2, equal_nan=True)
    eng.close()


def test_fused_kernel_with_sightlines(cuda_device):
    """Shared axis (fused kernel) + several sightlines + per-draw sightline index."""
    from caribou_hi_b200 import Engine

    spec, data, free, baseline = make_case("emission_absorption_physical", 3, 150, 150, 30, seed=61, shared_axis=True)
    rng = np.random.default_rng(62)
    n_sl = 4
    ax = data["emission"].spectral
    ye = data["emission"].brightness[None, :] + 0.3 * rng.standard_normal((n_sl, ax.size))
    ya = data["absorption"].brightness[None, :] + 0.02 * rng.standard_normal((n_sl, ax.size))
    sl = rng.integers(0, n_sl, size=30).astype(np.int32)
    eng = Engine("emission_absorption_physical", 3, prior_amplitude=spec["prior_ff_NHI"])
    eng.set_spectra(emission=dict(spectral=ax, brightness=ye, noise=0.1), absorption=dict(spectral=ax, brightness=ya, noise=0.01),
                    n_sightlines=n_sl)
    lp, grad, _, _ = eng.logp_grad(eng.pack(free), sightline=sl)
    got = eng.unpack_grad(grad)
    names = sorted(got)
    for i in range(n_sl):
        sel = np.where(sl == i)[0]
        if sel.size == 0:
            continue
        d = {"emission": mr.OracleData(ax, ye[i], 0.1), "absorption": mr.OracleData(ax, ya[i], 0.01)}
        for o in d.values():
            o._brightness_offset = 0.0
        ref_lp, ref_g = oracle_logp_grad(spec, d, {k: v[sel] for k, v in free.items()}, {})
        fin = np.isfinite(ref_lp)
        assert_close(lp[sel], ref_lp, what="fused sightline logp", axis=None)
        assert_close(np.concatenate([got[n][sel][fin] for n in names], axis=1),
                     np.concatenate([ref_g[n][fin] for n in names], axis=1), what="fused sightline grad")
    eng.close()


def test_packed_rows_entry_points(cuda_device):
    """chi_logp_grad_rows[_dev]: the packed blocks (only the model's free RVs) give the same
    bits as the CHI_NSLOT-row blocks, on the host and on the device entry point."""
    import torch

    from caribou_hi_b200._lib import ChiError

    for kind, lor, shared in (("emission_absorption_physical", False, True), ("absorption", True, False),
                              ("emission_absorption", True, False)):
        spec, data, free, baseline = make_case(kind, 3, 160, 160, 40, seed=71, lorentzian=lor, baseline_degree=1,
                                               shared_axis=shared)
        eng = engine_for(spec, data, 1)
        ce, ca = baseline.get("baseline_emission_norm"), baseline.get("baseline_absorption_norm")
        theta = eng.pack(free)
        lp, grad, gce, gca = eng.logp_grad(theta, ce, ca)
        rows = eng.pack_rows(free)
        assert rows.shape[1] == len(eng.slots) < theta.shape[1]
        lp_r, grad_r, gce_r, gca_r = eng.logp_grad_rows(rows, ce, ca)
        assert np.array_equal(lp_r, lp, equal_nan=True)
        assert np.array_equal(grad_r, grad[:, eng.row_slots, :], equal_nan=True)
        for a, b in ((gce_r, gce), (gca_r, gca)):
            assert (a is None and b is None) or np.array_equal(a, b, equal_nan=True)
        got, want = eng.unpack_rows(grad_r), eng.unpack_grad(grad)
        assert all(np.array_equal(got[k], want[k], equal_nan=True) for k in want)
        # device entry point
        d = lambda x: None if x is None else torch.from_numpy(np.ascontiguousarray(x)).cuda()  # noqa: E731
        rows_d, lp_d, gr_d = d(rows), torch.empty(40, dtype=torch.float64, device="cuda"), torch.empty_like(d(rows))
        eng.logp_grad_rows_dev(40, rows_d, lp_d, gr_d, coeff_e=d(ce), coeff_a=d(ca))
        eng.synchronize()
        assert np.array_equal(lp_d.cpu().numpy(), lp, equal_nan=True)
        assert np.array_equal(gr_d.cpu().numpy(), grad_r, equal_nan=True)
        # the slot-layout entry point is unaffected afterwards
        assert np.array_equal(eng.logp_grad(theta, ce, ca)[1], grad, equal_nan=True)
        with pytest.raises(ChiError):  # duplicate slot
            eng._check(eng._lib.chi_logp_grad_rows(eng._h, 1, 2, np.array([1, 1], dtype=np.int32).ctypes.data_as(
                __import__("ctypes").POINTER(__import__("ctypes").c_int32)), None, None, None, None, None, None, None, None))
        eng.close()


@pytest.mark.parametrize("kind", ["emission", "emission_absorption", "emission_physical",
                                  "emission_absorption_physical", "absorption", "absorption_physical"])
def test_branch_edges_of_the_parameter_maps(cuda_device, kind):
    """The switch / clip branches of add_priors (emission_model.py:82-132,
    emission_physical_model.py:81-129), the Tk < 300 K switch of the spin temperature
    (physics.py:95-101), optically thick clouds with f = 1, and fwhm_L = 0 under the
    pseudo-Voigt profile: both sides of every branch must occur in the batch and match."""
    B, N = 768, 3
    spec, data, free, baseline = make_case(kind, N, 96, 64, B, seed=909, lorentzian=True)
    rng = np.random.default_rng(910)
    amp = [k for k in free if k in ("TB_fwhm_norm", "ff_NHI_norm", "tau_total_norm", "NHI_fwhm2_thermal_norm")][0]
    free[amp] = free[amp] * 10 ** rng.uniform(-3, 3, size=(B, 1))          # sweeps tkin_min/ff_min/frac_min past 1
    free["fwhm2_norm"] = free["fwhm2_norm"] * 10 ** rng.uniform(-2, 1, size=(B, 1))  # Tk on both sides of 300 K
    for k in ("filling_factor_norm", "tkin_factor_norm", "fwhm2_thermal_fraction_norm"):
        if k in free:  # exact interval ends: ff = ff_min makes the clip at 0.9999 active
            free[k][: B // 8] = 0.0
            free[k][B // 8 : B // 4] = 1.0
    free["fwhm_L_norm"][B // 2 :: 7] = 0.0                                   # eta = 0 under the PV profile
    det = mr.cloud_inputs(spec, free)
    tk = np.broadcast_to(det["tkin"], (B, N))
    assert (tk < 300).any() and (tk > 300).any()
    if "filling_factor" in det and mr.has_emission(spec):
        ff = np.broadcast_to(det["filling_factor"], (B, N))
        assert (ff == 1.0).any() and (ff < 1.0).any()                        # ff_min > 1 and ff_min <= 1
    ref_lp, ref_g = oracle_logp_grad(spec, data, free, baseline)
    eng = engine_for(spec, data, 0)
    lp, grad, _, _ = eng.logp_grad(eng.pack(free), baseline.get("baseline_emission_norm"),
                                   baseline.get("baseline_absorption_norm"))
    assert np.array_equal(np.isfinite(lp), np.isfinite(ref_lp))
    fin = np.isfinite(ref_lp)
    assert fin.sum() > B // 4
    assert_close(lp, ref_lp, what=f"{kind} logp", axis=None)
    # amplitudes swept over six decades make some draws so ill-conditioned that fp64 autograd itself
    # is 6e-10 off the 60-digit value (tests/test_oracle_arbiter.py::test_extreme_draw): those
    # components are arbitrated automatically, everything else holds 1e-10 per RV vector
    assert_grad_close(eng.unpack_grad(grad), ref_g, fin, case=(spec, data, free, baseline), what=f"{kind} grad",
                      max_arbitrations=200)
    d_dev = eng.deterministics(eng.pack(free))
    for k in ("tkin", "tspin", "filling_factor", "tau_total"):
        if k in det and k in d_dev:
            assert_close(d_dev[k][fin], np.broadcast_to(det[k], (B, N))[fin], rtol=1e-12, what=k)
    eng.close()


def test_asynchronous_host_calls(cuda_device):
    """chi_logp_grad_rows_async + chi_wait: several calls in flight on double-buffered pinned arrays
    give the same bits as the synchronous call, in issue order, including tickets that have left
    the ring."""
    from caribou_hi_b200 import PinnedArray

    spec, data, free, baseline = make_case("emission_absorption_physical", 3, 160, 160, 20000, seed=81, shared_axis=True)
    eng = engine_for(spec, data, 0)
    rows = eng.pack_rows(free)
    ce, ca = baseline["baseline_emission_norm"], baseline["baseline_absorption_norm"]
    B = rows.shape[0]
    ref = [eng.logp_grad_rows(np.roll(rows, k, axis=0), np.roll(ce, k, axis=0), np.roll(ca, k, axis=0)) for k in range(3)]
    keep, bufs = [], []
    for k in range(3):
        p = {"rows": PinnedArray(rows.shape), "ce": PinnedArray(ce.shape), "ca": PinnedArray(ca.shape),
             "logp": PinnedArray((B,)), "grad": PinnedArray(rows.shape), "gce": PinnedArray(ce.shape), "gca": PinnedArray(ca.shape)}
        p["rows"].array[...] = np.roll(rows, k, axis=0)
        p["ce"].array[...] = np.roll(ce, k, axis=0)
        p["ca"].array[...] = np.roll(ca, k, axis=0)
        keep.append(p)
        bufs.append(dict(logp=p["logp"].array, grad=p["grad"].array, gcoeff_e=p["gce"].array, gcoeff_a=p["gca"].array))
    tickets = []
    for rep in range(4):  # 12 calls: more than the ticket ring holds
        for k in range(3):
            if rep > 0:
                eng.wait(tickets[-3])  # the buffers of set k are about to be reused
            tickets.append(eng.logp_grad_rows_async(keep[k]["rows"].array, keep[k]["ce"].array, keep[k]["ca"].array, bufs[k]))
    eng.wait(tickets[0])   # long gone from the ring: completed by stream order
    eng.wait()             # everything
    for k in range(3):
        assert np.array_equal(bufs[k]["logp"], ref[k][0], equal_nan=True)
        assert np.array_equal(bufs[k]["grad"], ref[k][1], equal_nan=True)
        assert np.array_equal(bufs[k]["gcoeff_e"], ref[k][2], equal_nan=True)
        assert np.array_equal(bufs[k]["gcoeff_a"], ref[k][3], equal_nan=True)
    eng.close()


@pytest.mark.parametrize("B", [24, 20000])
def test_logp_only_calls_match_the_gradient_calls(cuda_device, B):
    """want_grad=False (no Jacobian / no gradient buffers) gives the same logp, bit for bit, as the
    call that also returns the gradient -- small batches (side-stream Jacobian + contraction) and
    large ones (fused epilogue), likelihood and total model logp."""
    spec, data, free, baseline = make_case("emission_absorption_physical", 3, 160, 160, B, seed=77, lorentzian=True,
                                           baseline_degree=1, shared_axis=True)
    eng = engine_for(spec, data, 1)
    theta = eng.pack(free)
    ce, ca = baseline["baseline_emission_norm"], baseline["baseline_absorption_norm"]
    lp_g = eng.logp_grad(theta, ce, ca)[0]
    lp_only, g, gce, gca = eng.logp_grad(theta, ce, ca, want_grad=False)
    assert g is None and gce is None and gca is None
    assert np.array_equal(lp_only, lp_g, equal_nan=True)
    if B < 1000:
        assert_close(lp_only, oracle_logp_grad(spec, data, free, baseline)[0], what="logp only", axis=None)
    eng.set_priors()
    u = eng.transform(theta, direction=-1)
    fin = np.isfinite(u).all(axis=(1, 2))
    m_g = eng.model_logp_grad(u[fin], ce[fin], ca[fin])[0]
    m_only = eng.model_logp_grad(u[fin], ce[fin], ca[fin], want_grad=False)[0]
    assert np.array_equal(m_only, m_g, equal_nan=True)
    assert np.isfinite(m_only).sum() > 0.3 * fin.sum()
    eng.close()


@pytest.mark.parametrize("shared", [False, True])
@pytest.mark.parametrize("lorentzian", [False, True])
def test_nan_inputs_propagate(cuda_device, shared, lorentzian):
    """A NaN in any likelihood input of any cloud (the reference's prior produces them: depth 0 -> spin
    temperature NaN) must make both spectra, logp and the whole gradient NaN -- also for a line much narrower
    than a channel, whose Gaussian underflows to exactly 0 in every channel (NaN * 0 is NaN), and whatever the
    sign bit of the NaN."""
    from caribou_hi_b200 import Engine

    v = np.linspace(-30, 30, 64)
    va = v if shared else np.linspace(-20, 20, 48)
    eng = Engine("physics", 3, lorentzian=lorentzian, has_emission=True, has_absorption=True)
    eng.set_spectra(emission=dict(spectral=v, brightness=np.ones(64), noise=0.1),
                    absorption=dict(spectral=va, brightness=np.full(va.size, 0.2), noise=0.01))
    base = {"velocity": [1.0, -2.4, -8.0], "fwhm2": [0.0652, 288.0, 200.0], "fwhm_L": [0.3 * lorentzian] * 3,
            "tau_total": [1.5, 2.5, 0.57], "tspin": [80.0, 432.0, 1513.0], "filling_factor": [1.0, 0.74, 0.63],
            "absorption_weight": [0.075, 0.8, 0.07]}
    neg_nan = np.float64(np.nan) * -1.0
    neg_nan = np.copysign(np.nan, -1.0)
    for name in ("tau_total", "velocity", "fwhm2", "absorption_weight"):
        for n in range(3):
            for bad in (np.nan, neg_nan):
                inp = {k: np.array([x], dtype=np.float64) for k, x in base.items()}
                inp[name][0, n] = bad
                theta = eng.pack(inp)
                mu_e, mu_a = eng.spectra(theta)
                assert np.isnan(mu_a).all(), (name, n, "mu_a")
                if name != "absorption_weight":
                    assert np.isnan(mu_e).all(), (name, n, "mu_e")
                for path in (1, 2):
                    eng.set_option("small_path", path)
                    lp, g, _, _ = eng.logp_grad(theta)
                    assert np.isnan(lp).all(), (name, n, path)
                eng.set_option("small_path", 0)
                if lorentzian or name != "fwhm2":
                    mu_e32, mu_a32 = eng.spectra(theta, dtype=np.float32)
                    assert np.isnan(mu_a32).all(), (name, n, "mu_a f32")
    eng.close()


@pytest.mark.parametrize("two_axes", [False, True])
@pytest.mark.parametrize("sightlines", [1, 3])
def test_negative_depths_in_a_large_batch(cuda_device, sightlines, two_axes):
    """Large Gaussian batches on a shared axis run the fused channel kernel as two launches (moments_ea.cuh,
    MODE): every draw without the +inf branch of 2^y, and the draws the cloud map marked -- a negative or NaN
    optical depth, which the seven physics-level inputs allow -- on a small scanning grid with it.  Marked
    draws of every flavour (mildly negative, overflowing, NaN, negative absorption weight only) in a batch
    large enough for that form must give what the one-launch form gives, bit for bit, and what the oracle
    gives: values where it is finite, the same NaN / -inf elsewhere.  two_axes: the spectra on different axes,
    i.e. the emission kernel's two-launch form next to the one-launch absorption kernel."""
    import torch
    from caribou_hi_b200 import Engine

    rng = np.random.default_rng(41)
    B, N, S = 8000, 4, 96  # past the small-batch forms (batch_is_small in csrc/chi.cu)
    v = np.linspace(-40, 40, S)
    va = np.linspace(-30, 30, S) if two_axes else v
    inp = {
        "velocity": rng.normal(0, 8, (B, N)),
        "fwhm2": rng.uniform(5, 200, (B, N)),
        "fwhm_L": np.zeros((B, N)),
        "tau_total": rng.uniform(0.05, 8, (B, N)),
        "tspin": rng.uniform(30, 3000, (B, N)),
        "filling_factor": rng.uniform(0.1, 1.0, (B, N)),
        "absorption_weight": rng.uniform(0.3, 1.2, (B, N)),
    }
    kind = rng.integers(0, 20, B)  # 0..4 are the marked flavours, one draw in four
    cloud = rng.integers(0, N, B)
    rows = np.arange(B)
    # (the last flavour: a depth so large that exp(-tau) underflows -- the plain form of the kernel leaves out
    # that select too and hands such draws to the scanning grid)
    for k, (name, value) in enumerate([("tau_total", -0.4), ("tau_total", -2.0e4), ("tau_total", np.nan),
                                       ("absorption_weight", -0.7), ("tau_total", 3.0e4)]):
        sel = kind == k
        inp[name][rows[sel], cloud[sel]] = value
    assert (kind < 5).sum() > 1000
    ye = rng.normal(5, 3, (sightlines, S))
    ya = rng.normal(0.2, 0.1, (sightlines, S))
    sl = rng.integers(0, sightlines, B).astype(np.int32)

    eng = Engine("physics", N, lorentzian=False, bg_temp=3.77)
    if sightlines == 1:
        eng.set_spectra(emission=dict(spectral=v, brightness=ye[0], noise=0.3, offset=0.0),
                        absorption=dict(spectral=va, brightness=ya[0], noise=0.02, offset=0.0))
        sl_arg = None
    else:
        eng.set_spectra(emission=dict(spectral=v, brightness=ye, noise=0.3, offset=0.0),
                        absorption=dict(spectral=va, brightness=ya, noise=0.02, offset=0.0), n_sightlines=sightlines)
        sl_arg = sl
    theta = eng.pack(inp)
    n0 = eng.launch_count()
    lp2, g2, _, _ = eng.logp_grad(theta, sightline=sl_arg)
    n_two = eng.launch_count() - n0
    eng.set_option("ea_one_pass", 1)
    n0 = eng.launch_count()
    lp1, g1, _, _ = eng.logp_grad(theta, sightline=sl_arg)
    n_one = eng.launch_count() - n0
    eng.set_option("ea_one_pass", 0)
    assert n_two > n_one, "the two-launch form did not run"
    assert np.array_equal(lp2, lp1, equal_nan=True)
    assert np.array_equal(g2, g1, equal_nan=True)

    spec = mr.make_spec("emission_absorption", N, bg_temp=3.77)
    ref_lp = np.zeros(B)
    ref_g = {k: np.zeros((B, N)) for k in inp}
    for s in range(sightlines):
        m = sl == s if sightlines > 1 else np.ones(B, bool)
        data = {"emission": mr.OracleData(v, ye[s], 0.3), "absorption": mr.OracleData(va, ya[s], 0.02)}
        for d in data.values():
            d._brightness_offset = 0.0
        it = {k: torch.tensor(x[m], requires_grad=True) for k, x in inp.items()}
        lp = mr.logp_from_inputs(spec, data, it)  # fwhm_L = 0: the pseudo-Voigt profile is the Gaussian
        lp.sum().backward()
        ref_lp[m] = lp.detach().numpy()
        for k in inp:
            ref_g[k][m] = it[k].grad.numpy()
    fin = np.isfinite(ref_lp)
    assert fin.sum() > 0.8 * B and (~fin).sum() > 300
    assert np.array_equal(np.isnan(lp2), np.isnan(ref_lp)) and np.array_equal(np.isneginf(lp2), np.isneginf(ref_lp))
    assert_close(lp2[fin], ref_lp[fin], what="logp of a batch with marked draws", axis=None)
    got = eng.unpack_grad(g2)
    mild = fin & (kind < 5)
    assert mild.sum() > 400  # negative depths that do not overflow: finite, and evaluated by the scanning grid
    for name in inp:
        if name == "fwhm_L":
            continue
        assert_close(got[name][fin], ref_g[name][fin], what=f"d/d{name} of a batch with marked draws")
    if sightlines == 1:
        # the spectra kernels choose between their two loops per block (= per draw) from the same marks
        data = {"emission": mr.OracleData(v, ye[0], 0.3), "absorption": mr.OracleData(va, ya[0], 0.02)}
        for d in data.values():
            d._brightness_offset = 0.0
        with np.errstate(all="ignore"):
            ref_mu = mr.predict_from_inputs(spec, data, {k: torch.tensor(x[:600]) for k, x in inp.items()})
        mu_e, mu_a = eng.spectra(theta[:600])
        for got_mu, key in ((mu_e, "emission"), (mu_a, "absorption")):
            ref = ref_mu[key].numpy()
            ok = np.isfinite(ref).all(axis=1)
            assert ok.sum() > 450 and (~ok).sum() > 20
            assert_close(got_mu[ok], ref[ok], what=f"mu_{key} of a batch with marked draws")
            assert np.array_equal(np.isnan(got_mu[~ok]), np.isnan(ref[~ok])), key
            assert np.array_equal(np.isinf(got_mu[~ok]), np.isinf(ref[~ok])), key
    eng.close()


@pytest.mark.parametrize("shared", [False, True])
def test_many_baseline_terms_on_the_small_paths(cuda_device, shared):
    """Baselines of degree 7 (CHI_MAX_BASELINE terms per dataset): the one-launch kernel keeps one column of
    per-thread sums per extra term in dynamic shared memory, 28 KB here on top of its 36 KB of static arrays --
    past the 48 KB a launch gets without the opt-in (csrc/small.cuh: sm_allow_dyn)."""
    kind = "emission_absorption_physical"
    if shared:
        n_ok = _check_case(kind, 3, 96, 96, 12, seed=77, lorentzian=True, baseline_degree=7, shared_axis=True)
    else:
        n_ok = _check_case(kind, 3, 96, 64, 12, seed=78, lorentzian=False, baseline_degree=7)
    assert n_ok >= 6


@pytest.mark.parametrize("kind_n", [("emission_absorption_physical", 8, True, True), ("emission_absorption", 2, False, False),
                                    ("absorption_physical", 3, True, False), ("emission_physical", 5, False, False)])
def test_cluster_form_of_the_one_launch_kernel(cuda_device, kind_n):
    """Several blocks per draw of the one-launch evaluation: as one thread-block cluster that meets in the first
    block's (distributed) shared memory (default), or through records and a counter in global memory
    (small_cluster 1).  Same sums in the same order: the results agree bit for bit, for likelihood and model
    logp, whatever the number of blocks per draw."""
    kind, N, lorentzian, shared = kind_n
    spec, data, free, baseline = make_case(kind, N, 700, 420, 10, seed=31 + N, lorentzian=lorentzian, baseline_degree=1,
                                           shared_axis=shared)
    eng = engine_for(spec, data, 1)
    theta = eng.pack(free)
    ce, ca = baseline.get("baseline_emission_norm"), baseline.get("baseline_absorption_norm")
    eng.set_option("small_path", 2)
    ref_lp, _ = oracle_logp_grad(spec, data, free, baseline)
    for chunks in (0, 2, 3, 5, 8):
        eng.set_option("small_chunks", chunks)
        out = []
        for off in (0, 1):
            eng.set_option("small_cluster", off)
            out.append(eng.logp_grad(theta, ce, ca))
        for x, y in zip(*out):
            if x is not None:
                assert np.array_equal(x, y, equal_nan=True), (chunks,)
        assert_close(out[0][0], ref_lp, what=f"{kind} logp, cluster of {chunks}", axis=None)
    eng.set_option("small_chunks", 0)
    eng.set_option("small_cluster", 0)
    # the model logp (priors and transforms on the device) through the same two forms
    eng.set_priors()
    u = np.random.default_rng(5).normal(0, 0.3, theta.shape)
    out = []
    for off in (0, 1):
        eng.set_option("small_cluster", off)
        out.append(eng.model_logp_grad(u, ce, ca))
    eng.set_option("small_cluster", 0)
    assert np.isfinite(out[0][0]).sum() >= 5
    for x, y in zip(*out):
        if x is not None:
            assert np.array_equal(x, y, equal_nan=True)
    eng.close()
