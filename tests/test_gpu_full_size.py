"""BASELINE.json's full sizes (configs[1]: 4 clouds, 1000+1000 channels, 1e5 prior draws;
configs[2]: 8 clouds, 2048+2048 channels, pseudo-Voigt, 64 chains) through size-independent
properties of the path, plus an oracle spot check of a random subset:

  * additivity of logp and of the gradient over a split of the channels,
  * per-draw independence (a permuted batch gives the permuted results bit for bit),
  * linearity of the VJP in the cotangent,
  * consistency of the two kernel families: the Gaussian log-likelihood of the spectra kernel's
    output equals the fused logp,
  * commutativity of the absorption clouds (1 - exp(-sum tau) does not depend on their order).
"""

import numpy as np
import pytest

from helpers import assert_close, assert_grad_close, make_case, oracle_logp_grad
from oracle import models_ref as mr

pytestmark = pytest.mark.gpu

B_FULL = 100_000


def _engine(kind, N, data, lorentzian=False, prior_amplitude=None):
    from caribou_hi_b200 import Engine

    kw = {"prior_amplitude": prior_amplitude} if prior_amplitude is not None else {}
    eng = Engine(kind, N, lorentzian=lorentzian, prior_fwhm_L=1.0 if lorentzian else 0.0, **kw)
    eng.set_spectra(**{k: dict(spectral=d["spectral"], brightness=d["brightness"], noise=d["noise"])
                       for k, d in data.items()})
    return eng


def _draws(eng, kind, N, B, seed, lorentzian=False):
    from caribou_hi_b200 import synthetic

    rng = np.random.default_rng(seed)
    th = lambda trial: eng.deterministics(eng.pack(trial))["fwhm2_thermal" if kind.endswith("physical") else "tkin"]  # noqa: E731
    return synthetic.draw_free(kind, N, B, rng, lorentzian=lorentzian, thermal_fn=th)


def _config2_data(seed=2024, shared=True):
    rng = np.random.default_rng(seed)
    S = 1000
    ve = np.linspace(-100.0, 100.0, S)
    va = ve if shared else np.linspace(-80.0, 80.0, S)
    g = lambda v, c, w: np.exp(-4 * np.log(2) * (v - c) ** 2 / w**2)  # noqa: E731
    ye = 30 * g(ve, -8, 12) + 12 * g(ve, 15, 30) + 0.1 * rng.standard_normal(S)
    ya = 1 - np.exp(-(0.8 * g(va, -8, 10) + 0.2 * g(va, 15, 25))) + 0.01 * rng.standard_normal(S)
    return {"emission": dict(spectral=ve, brightness=ye, noise=0.1),
            "absorption": dict(spectral=va, brightness=ya, noise=0.01)}


@pytest.mark.parametrize("shared", [True, False], ids=["fused", "separate"])
def test_config2_full_size_properties(cuda_device, shared):
    kind, N = "emission_absorption_physical", 4
    data = _config2_data(shared=shared)
    eng = _engine(kind, N, data, prior_amplitude=1e21)
    theta = eng.pack(_draws(eng, kind, N, B_FULL, 77))
    lp, grad, _, _ = eng.logp_grad(theta)
    fin = np.isfinite(lp)
    assert 0.5 < fin.mean() < 1.0  # ~30 % of the prior draws are NaN in the reference too
    assert np.array_equal(fin, np.isfinite(grad).all(axis=(1, 2)))

    # ---- per-draw independence: a permuted batch gives the permuted results bit for bit
    perm = np.random.default_rng(1).permutation(B_FULL)
    lp_p, grad_p, _, _ = eng.logp_grad(theta[perm])
    assert np.array_equal(lp_p, lp[perm], equal_nan=True)
    assert np.array_equal(grad_p, grad[perm], equal_nan=True)

    # ---- additivity over a split of the channels (same channel spacing on both halves)
    halves = []
    for sl in (slice(0, 500), slice(500, 1000)):
        part = {k: dict(spectral=d["spectral"][sl], brightness=d["brightness"][sl], noise=d["noise"]) for k, d in data.items()}
        e2 = _engine(kind, N, part, prior_amplitude=1e21)
        halves.append(e2.logp_grad(theta)[:2])
        e2.close()
    assert_close(halves[0][0] + halves[1][0], lp, rtol=1e-12, what="logp additivity", axis=None)
    g_sum = (halves[0][1] + halves[1][1]).reshape(B_FULL, -1)
    assert_close(g_sum[fin], grad.reshape(B_FULL, -1)[fin], rtol=1e-11, what="gradient additivity")

    # ---- the spectra kernel and the fused logp agree: logp = sum Normal(y | mu, sigma)
    sub = np.random.default_rng(2).choice(B_FULL, 4000, replace=False)
    mu_e, mu_a = eng.spectra(theta[sub])
    ll = np.zeros(sub.size)
    for mu, d in ((mu_e, data["emission"]), (mu_a, data["absorption"])):
        s = d["noise"]
        ll += np.sum(-0.5 * ((d["brightness"] - mu) / s) ** 2 - np.log(s) - 0.5 * np.log(2 * np.pi), axis=1)
    assert_close(ll, lp[sub], rtol=1e-11, what="logp from spectra", axis=None)

    # ---- linearity of the VJP in the cotangent
    rng = np.random.default_rng(3)
    sub = sub[:2000]
    S = 1000
    g1e, g2e, g1a, g2a = (rng.standard_normal((sub.size, S)) for _ in range(4))
    v1 = eng.spectra_vjp(theta[sub], g1e, g1a)[0]
    v2 = eng.spectra_vjp(theta[sub], g2e, g2a)[0]
    v12 = eng.spectra_vjp(theta[sub], 0.7 * g1e - 1.9 * g2e, 0.7 * g1a - 1.9 * g2a)[0]
    ok = np.isfinite(v12).all(axis=(1, 2))
    assert_close(v12.reshape(sub.size, -1)[ok], (0.7 * v1 - 1.9 * v2).reshape(sub.size, -1)[ok], rtol=1e-11, what="VJP linearity")

    # ---- oracle spot check of a random subset of the full batch
    pick = np.random.default_rng(4).choice(B_FULL, 96, replace=False)
    spec = mr.make_spec(kind, N, prior_ff_NHI=1e21)
    odata = {k: mr.OracleData(d["spectral"], d["brightness"], d["noise"]) for k, d in data.items()}
    for o in odata.values():
        o._brightness_offset = 0.0  # no baseline in this workload
    free = eng.unpack_grad(theta[pick])  # the same layout maps names -> [B, N] slices
    ref_lp, ref_g = oracle_logp_grad(spec, odata, free, {})
    assert_close(lp[pick], ref_lp, what="oracle logp", axis=None)
    got = eng.unpack_grad(grad[pick])
    f2 = np.isfinite(ref_lp)
    assert_grad_close(got, ref_g, f2, case=(spec, odata, free, {}), what="oracle gradient")
    eng.close()


def test_config3_full_size_properties(cuda_device):
    """64 chains x 8 clouds x (2048 + 2048) channels, pseudo-Voigt: the chunked small-batch path."""
    kind, N, S = "emission_absorption_physical", 8, 2048
    spec, data, free, _ = make_case(kind, N, S, S, 64, seed=31, lorentzian=True)
    d = {k: dict(spectral=o.spectral, brightness=o.brightness - o._brightness_offset, noise=o.noise) for k, o in data.items()}
    eng = _engine(kind, N, d, lorentzian=True, prior_amplitude=spec["prior_ff_NHI"])
    theta = eng.pack(free)
    lp, grad, _, _ = eng.logp_grad(theta)
    # the same 64 chains inside a large batch take the un-chunked kernels: same numbers
    big = np.concatenate([theta] * 64)
    lp_b, grad_b, _, _ = eng.logp_grad(big)
    fin = np.isfinite(lp)
    assert_close(lp_b[:64], lp, rtol=1e-12, what="chunked vs un-chunked logp", axis=None)
    assert_close(grad_b[:64].reshape(64, -1)[fin], grad.reshape(64, -1)[fin], rtol=1e-11, what="chunked vs un-chunked grad")
    assert np.array_equal(lp_b[64:128], lp_b[:64], equal_nan=True)
    eng.close()


def test_absorption_clouds_commute_at_full_size(cuda_device):
    kind, N = "absorption", 4
    data = {"absorption": _config2_data()["absorption"]}
    eng = _engine(kind, N, data)
    free = _draws(eng, kind, N, B_FULL, 5)
    theta = eng.pack(free)
    lp, grad, _, _ = eng.logp_grad(theta)
    order = [2, 0, 3, 1]
    lp_p, grad_p, _, _ = eng.logp_grad(np.ascontiguousarray(theta[:, :, order]))
    assert_close(lp_p, lp, rtol=1e-12, what="cloud order logp", axis=None)
    fin = np.isfinite(lp)
    assert_close(grad_p.reshape(B_FULL, -1)[fin], grad[:, :, order].reshape(B_FULL, -1)[fin], rtol=1e-11, what="cloud order grad")
    eng.close()


def test_config4_survey_full_size(cuda_device):
    """BASELINE configs[3]: 10^4 sightlines x 5 clouds x 1000 channels (EmissionAbsorptionPhysical,
    one spectrum pair per sightline, a few chains per sightline).  Properties: a draw's result
    depends on its own sightline only (single-sightline engines reproduce sampled rows bit for
    bit), and sharding by sightline (caribou_hi_b200.sharding) returns the same rows."""
    import time

    from caribou_hi_b200 import Engine, sharding

    kind, N, S, n_sl, per = "emission_absorption_physical", 5, 1000, 10_000, 2
    rng = np.random.default_rng(8)
    v = np.linspace(-100.0, 100.0, S)
    g = lambda c, w: np.exp(-4 * np.log(2) * (v[None, :] - c[:, None]) ** 2 / w[:, None] ** 2)  # noqa: E731
    c1, c2 = rng.normal(0, 20, n_sl), rng.normal(0, 20, n_sl)
    w1, w2 = rng.uniform(5, 25, n_sl), rng.uniform(5, 25, n_sl)
    ye = 30 * g(c1, w1) + 10 * g(c2, w2) + 0.1 * rng.standard_normal((n_sl, S))
    ya = 1 - np.exp(-(0.7 * g(c1, 0.8 * w1) + 0.2 * g(c2, 0.8 * w2))) + 0.01 * rng.standard_normal((n_sl, S))
    eng = Engine(kind, N, prior_amplitude=1e21)
    eng.set_spectra(emission=dict(spectral=v, brightness=ye, noise=0.1), absorption=dict(spectral=v, brightness=ya, noise=0.01),
                    n_sightlines=n_sl)
    B = n_sl * per
    theta = eng.pack(_draws(eng, kind, N, B, 9))
    sl = np.repeat(np.arange(n_sl, dtype=np.int32), per)
    lp, grad, _, _ = eng.logp_grad(theta, sightline=sl)
    t0 = time.perf_counter()
    lp2, grad2, _, _ = eng.logp_grad(theta, sightline=sl)
    dt = time.perf_counter() - t0
    assert np.array_equal(lp, lp2, equal_nan=True) and np.array_equal(grad, grad2, equal_nan=True)
    print(f"\\nsurvey sweep: {B} draws over {n_sl} sightlines in {dt * 1e3:.2f} ms host call "
          f"({B * N * 2 * S / dt:.3e} units/s incl. copies)")
    fin = np.isfinite(lp)
    assert 0.4 < fin.mean() < 1.0

    # a draw depends on its own sightline only
    for i in rng.choice(n_sl, 5, replace=False):
        one = Engine(kind, N, prior_amplitude=1e21)
        one.set_spectra(emission=dict(spectral=v, brightness=ye[i], noise=0.1),
                        absorption=dict(spectral=v, brightness=ya[i], noise=0.01))
        rows = np.where(sl == i)[0]
        big = np.concatenate([theta[rows]] * 4096)  # same large-call arithmetic as the survey sweep
        lp1, g1, _, _ = one.logp_grad(big)
        assert np.array_equal(lp1[: rows.size], lp[rows], equal_nan=True)
        assert np.array_equal(g1[: rows.size], grad[rows], equal_nan=True)
        one.close()

    # sharding by sightline: each shard evaluates its own sightlines, rows are the same
    for rank in range(4):
        lo, hi = sharding.shard_bounds(n_sl, 4, rank)
        sel = (sl >= lo) & (sl < hi)
        part = Engine(kind, N, prior_amplitude=1e21)
        part.set_spectra(emission=dict(spectral=v, brightness=ye[lo:hi], noise=0.1),
                         absorption=dict(spectral=v, brightness=ya[lo:hi], noise=0.01), n_sightlines=hi - lo)
        lp_s, g_s, _, _ = part.logp_grad(theta[sel], sightline=sl[sel] - lo)
        assert np.array_equal(lp_s, lp[sel], equal_nan=True)
        assert np.array_equal(g_s, grad[sel], equal_nan=True)
        part.close()
    eng.close()


def test_config3_and_config5_shapes_against_the_oracle(cuda_device):
    """The many-cloud pseudo-Voigt instantiations at the channel counts the BASELINE configs name, against the
    ORACLE (not a sibling kernel): k_moments_ea<8,true> and k_eval_small<2,true> at 2048+2048 channels
    (configs[2]) and k_spectra_ea_t<8,true> plus the fp32 spectra at 4096+4096 channels (configs[4])."""
    from helpers import engine_for

    for S, B in ((2048, 24), (4096, 16)):
        spec, data, free, baseline = make_case("emission_absorption_physical", 8, S, S, B, seed=4000 + S, lorentzian=True,
                                               baseline_degree=0, half_width=100.0, shared_axis=True)
        eng = engine_for(spec, data, 0)
        theta = eng.pack(free)
        ce, ca = baseline["baseline_emission_norm"], baseline["baseline_absorption_norm"]
        ref_mu = mr.predict(spec, data, free, baseline)
        mu_e, mu_a = eng.spectra(theta, ce, ca)
        assert_close(mu_e, ref_mu["emission"], what=f"S={S} mu_e")
        assert_close(mu_a, ref_mu["absorption"], what=f"S={S} mu_a")
        m32_e, m32_a = eng.spectra(theta, ce, ca, dtype=np.float32)
        assert_close(m32_e, ref_mu["emission"], rtol=1e-5, what=f"S={S} mu_e fp32")
        assert_close(m32_a, ref_mu["absorption"], rtol=1e-5, what=f"S={S} mu_a fp32")
        if S == 2048:
            ref_lp, ref_g = oracle_logp_grad(spec, data, free, baseline)
            fin = np.isfinite(ref_lp)
            assert fin.sum() >= 8
            for path in (1, 2):
                eng.set_option("small_path", path)
                lp, grad, gce, gca = eng.logp_grad(theta, ce, ca)
                assert_close(lp, ref_lp, what=f"S={S} logp path {path}", axis=None)
                got = dict(eng.unpack_grad(grad))
                got["baseline_emission_norm"], got["baseline_absorption_norm"] = gce, gca
                assert_grad_close(got, ref_g, fin, case=(spec, data, free, baseline), what=f"S={S} gradient path {path}")
        eng.close()
