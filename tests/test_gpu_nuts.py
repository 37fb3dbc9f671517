"""Chain-batched NUTS on the device: statistical checks.

1. With uninformative data (huge noise) the posterior is the prior, whose moments are known in
   closed form -- this exercises the sampler, the prior densities and the transforms together.
2. With informative synthetic data the posterior must concentrate on the generating parameters."""

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_nuts_recovers_the_prior(cuda_device):
    from caribou_hi_b200 import EmissionModel, SpecData

    rng = np.random.default_rng(0)
    v = np.linspace(-30, 30, 64)
    data = {"emission": SpecData(v, rng.normal(0, 1, 64), 1.0e7)}  # noise so large the likelihood is flat
    model = EmissionModel(data, 2, baseline_degree=0)
    model.add_priors(prior_tkin_factor=[2.0, 3.0])
    model.add_likelihood()
    chains, draws = 64, 400
    post, stats = model.sample(draws=draws, tune=400, chains=chains, seed=11)
    assert post["velocity_norm"].shape == (chains, draws, 2)
    assert np.all(np.isfinite(post["velocity_norm"]))
    assert stats["diverging"].mean() < 0.01
    assert 0.6 < stats["accept"].mean() < 0.95
    n_eff = chains * draws / 4.0  # conservative effective sample size

    def check(x, mean, var, name):
        x = x.reshape(-1)
        se = np.sqrt(var / n_eff)
        assert abs(x.mean() - mean) < 6 * se, f"{name}: mean {x.mean():.4f} vs {mean:.4f} (se {se:.4f})"
        assert abs(x.var() / var - 1.0) < 0.2, f"{name}: var {x.var():.4f} vs {var:.4f}"

    check(post["velocity_norm"], 0.0, 1.0, "Normal velocity_norm")
    check(post["log10_nHI_norm"], 0.0, 1.0, "Normal log10_nHI_norm")
    check(post["TB_fwhm_norm"], np.sqrt(2 / np.pi), 1 - 2 / np.pi, "HalfNormal TB_fwhm_norm")
    a, b = 2.0, 3.0
    check(post["tkin_factor_norm"], a / (a + b), a * b / ((a + b) ** 2 * (a + b + 1)), "Beta tkin_factor_norm")
    check(post["filling_factor_norm"], 0.5, 1 / 12, "Uniform filling_factor_norm")
    # chi-square(1): heavy tail, compare the median (0.4549) and the mean of log x instead
    f2 = post["fwhm2_norm"].reshape(-1)
    assert abs(np.median(f2) - 0.4549) < 0.04
    assert abs(np.mean(np.log(f2)) - (-1.2704)) < 0.12  # E[log chi2_1] = psi(1/2) + log 2
    # the baseline coefficient is scaled by the noise (bayes_spec normalisation), so the data DO
    # constrain it: N(0,1) prior x 64 channels of unit-information likelihood -> variance 1/65
    bl = post["baseline_emission_norm"].reshape(-1)
    assert abs(bl.var() * 65.0 - 1.0) < 0.2 and abs(bl.mean()) < 0.05


def test_nuts_recovers_generating_parameters(cuda_device):
    from caribou_hi_b200 import AbsorptionModel, SpecData

    rng = np.random.default_rng(5)
    v = np.linspace(-25, 25, 200)
    truth = {"fwhm2_norm": np.array([[0.05]]), "velocity_norm": np.array([[0.3]]),
             "log10_nHI_norm": np.array([[0.0]]), "log10_n_alpha_norm": np.array([[0.0]]),
             "tau_total_norm": np.array([[0.8]]), "tkin_factor": np.array([[0.5]]),
             "baseline_absorption_norm": np.array([[0.0]])}
    sim = AbsorptionModel({"absorption": SpecData(v, np.zeros(200), 0.01)}, 1, baseline_degree=0)
    sim.add_priors()
    sim.add_likelihood()
    clean = sim.predict(truth)["absorption"][0]
    data = {"absorption": SpecData(v, clean + 0.01 * rng.standard_normal(200), 0.01)}
    model = AbsorptionModel(data, 1, baseline_degree=0)
    model.add_priors()
    model.add_likelihood()
    # start near an optimiser-quality point, as the reference does with its ADVI initialisation
    init = {k: v * 1.2 for k, v in truth.items()}
    post, stats = model.sample(draws=300, tune=500, chains=32, seed=3, init_point=init)
    assert stats["diverging"].mean() < 0.05
    for name in ("velocity_norm", "fwhm2_norm", "tau_total_norm"):
        x = post[name][:, :, 0]
        chain_means = x.mean(axis=1)
        m, s = x.mean(), x.std()
        t = truth[name][0, 0]
        assert abs(m - t) < 5 * s + 1e-3, f"{name}: posterior {m:.4f} +- {s:.4f}, truth {t}"
        assert s < 0.1 * max(abs(t), 0.1) + 0.05, f"{name}: posterior not concentrated (std {s:.4f})"
        # chains agree with each other (potential-scale-reduction style check)
        assert chain_means.std() < 0.5 * s + 1e-3, f"{name}: chains disagree"
    # the likelihood at the posterior mean is as good as at the truth
    assert stats["logp"].max() > -1e4


def test_lock_step_and_asynchronous_chains_agree(cuda_device):
    """The lock-step mode (every chain waits for the deepest tree; thread per chain) and the default
    asynchronous mode (warp per chain) sample the same posterior: posterior means agree within
    their Monte-Carlo error, tree depths and acceptance rates are alike; the pooled mass adaptation
    (lock step only) runs."""
    from caribou_hi_b200 import AbsorptionModel, SpecData

    rng = np.random.default_rng(5)
    v = np.linspace(-25, 25, 200)
    truth = {"fwhm2_norm": np.array([[0.05]]), "velocity_norm": np.array([[0.3]]),
             "log10_nHI_norm": np.array([[0.0]]), "log10_n_alpha_norm": np.array([[0.0]]),
             "tau_total_norm": np.array([[0.8]]), "tkin_factor": np.array([[0.5]]),
             "baseline_absorption_norm": np.array([[0.0]])}
    sim = AbsorptionModel({"absorption": SpecData(v, np.zeros(200), 0.01)}, 1, baseline_degree=0)
    sim.add_priors(); sim.add_likelihood()
    clean = sim.predict(truth)["absorption"][0]
    model = AbsorptionModel({"absorption": SpecData(v, clean + 0.01 * rng.standard_normal(200), 0.01)}, 1, baseline_degree=0)
    model.add_priors(); model.add_likelihood()
    init = {k: x * 1.2 for k, x in truth.items()}
    runs = {}
    for mode, kw in (("async", {}), ("lock", {"lock_step": True}), ("pooled", {"pool_mass": True})):
        post, stats = model.sample(draws=300, tune=400, chains=32, seed=3, init_point=init, **kw)
        assert np.all(np.isfinite(post["velocity_norm"])) and stats["diverging"].mean() < 0.05
        runs[mode] = (post, stats)
    for name in ("velocity_norm", "fwhm2_norm", "tau_total_norm"):
        ref = runs["async"][0][name][:, :, 0]
        for mode in ("lock", "pooled"):
            x = runs[mode][0][name][:, :, 0]
            se = np.hypot(ref.std(), x.std()) / np.sqrt(32 * 300 / 10.0)  # ESS ~ draws / 10
            assert abs(x.mean() - ref.mean()) < 6 * se + 1e-4, f"{mode} {name}: {x.mean()} vs {ref.mean()}"
            assert 0.6 < x.std() / ref.std() < 1.6, f"{mode} {name}: spread {x.std()} vs {ref.std()}"
    d_async, d_lock = runs["async"][1]["depth"].mean(), runs["lock"][1]["depth"].mean()
    assert abs(d_async - d_lock) < 1.0
    assert abs(runs["async"][1]["accept"].mean() - runs["lock"][1]["accept"].mean()) < 0.1


def _sample_with(kind, N, options, chains=6, draws=40, tune=60, shared=False):
    import caribou_hi_b200 as chb
    from caribou_hi_b200 import SpecData

    cls = {"emission_absorption": chb.EmissionAbsorptionModel, "emission_absorption_physical": chb.EmissionAbsorptionPhysicalModel,
           "absorption_physical": chb.AbsorptionPhysicalModel, "emission": chb.EmissionModel}[kind]
    ve = np.linspace(-40, 40, 96)
    va = ve if shared else np.linspace(-20, 20, 48)
    data = {}
    if "emission" in kind:
        data["emission"] = SpecData(ve, np.full(96, 5.0), 0.5)
    if "absorption" in kind:
        data["absorption"] = SpecData(va, np.full(va.size, 0.2), 0.05)
    model = cls(data, N, baseline_degree=1)
    model.add_priors(prior_fwhm_L=1.0) if kind.endswith("physical") else model.add_priors()
    model.add_likelihood()
    for name, value in options.items():
        model.engine.set_option(name, value)
    n0 = model.engine.launch_count()
    post, stats = model.sample(draws=draws, tune=tune, chains=chains, seed=4)
    launches = model.engine.launch_count() - n0
    model.engine.close()
    return post, stats, launches


def _same_run(a, b):
    for k in a[0]:
        assert np.array_equal(a[0][k], b[0][k], equal_nan=True), k
    for k in a[1]:
        assert np.array_equal(a[1][k], b[1][k], equal_nan=True), k


@pytest.mark.parametrize("kind_n", [("emission_absorption", 2), ("emission_absorption_physical", 3)])
def test_tick_variants_agree(cuda_device, kind_n):
    """The sampler's common tick staged in shared memory (nuts_tick 2) does the same operations in the same
    order as the round-1 form that passes every phase through global memory (1): identical draws and
    statistics."""
    kind, N = kind_n
    _same_run(_sample_with(kind, N, {"nuts_tick": 2}), _sample_with(kind, N, {"nuts_tick": 1}))


@pytest.mark.parametrize("kind_n_shared", [("emission_absorption", 2, False), ("emission_absorption", 2, True),
                                           ("emission_absorption_physical", 3, True), ("emission_absorption_physical", 6, False),
                                           ("absorption_physical", 3, False), ("emission", 2, False)])
def test_chain_kernel_matches_the_graph_of_kernels(cuda_device, kind_n_shared):
    """The per-chain kernel (nuts_chain.cuh: one resident block runs whole leapfrog steps tick after tick; the
    default for small models) against the graph of two kernels per tick it fuses -- the staged tick and the
    one-launch evaluation with one block per draw: the same operations in the same order, so identical draws
    and statistics, from a handful of launches instead of thousands."""
    kind, N, shared = kind_n_shared
    chain = _sample_with(kind, N, {"nuts_tick": 3}, shared=shared)
    graph = _sample_with(kind, N, {"nuts_tick": 2, "small_path": 2, "small_chunks": 1}, shared=shared)
    _same_run(chain, graph)
    assert chain[2] < 60 and graph[2] > 20 * chain[2], (chain[2], graph[2])
    # the speculative form (the state machine on a ninth warp, next to the evaluation): the same draws again
    spec = _sample_with(kind, N, {"nuts_tick": 6}, shared=shared)
    _same_run(spec, chain)
    assert spec[2] < 60
    # the cluster form: the blocks of a chain (8 asked for; what the channel count allows) as one thread-block
    # cluster, against the same graph with that many blocks per draw
    cluster = _sample_with(kind, N, {"nuts_tick": 5}, shared=shared)
    graph_cl = _sample_with(kind, N, {"nuts_tick": 2, "small_path": 2, "small_chunks": 8}, shared=shared)
    _same_run(cluster, graph_cl)
    assert cluster[2] < 60
    # a model this small takes the speculative chain kernel by itself
    auto = _sample_with(kind, N, {}, shared=shared)
    _same_run(auto, spec)
    assert auto[2] == spec[2]
    # the wide form (512 threads per chain: the channels are spread over twice the warps, so the sums agree with
    # the 256-thread form to rounding and the chains part ways after a while)
    wide = _sample_with(kind, N, {"nuts_tick": 4}, shared=shared)
    assert wide[2] < 60
    for k in ("accept", "depth"):
        assert np.all(np.isfinite(wide[1][k]))
    assert abs(wide[1]["accept"].mean() - chain[1]["accept"].mean()) < 0.25, (wide[1]["accept"].mean(), chain[1]["accept"].mean())


def test_chain_kernels_with_one_sightline_per_chain(cuda_device):
    """The survey use of the sampler: every chain samples the posterior of its OWN sightline's spectra
    (chi_nuts_sample's sightline argument).  The per-chain kernel in its three forms against the graph of
    kernels they fuse: bit-identical draws for the 256-thread and the cluster form, and every chain's
    posterior mean velocity follows its own sightline's line."""
    from caribou_hi_b200 import Engine
    from caribou_hi_b200 import layout

    N, S, n_sl = 1, 128, 5
    v = np.linspace(-30, 30, S)
    centres = np.linspace(-12, 12, n_sl)
    rng = np.random.default_rng(9)
    ye = np.stack([8.0 * np.exp(-0.5 * ((v - c) / 3.0) ** 2) for c in centres]) + 0.2 * rng.standard_normal((n_sl, S))
    ya = np.stack([0.3 * np.exp(-0.5 * ((v - c) / 3.0) ** 2) for c in centres]) + 0.01 * rng.standard_normal((n_sl, S))

    def run(options):
        eng = Engine("emission_absorption", N, prior_velocity=(0.0, 15.0))
        eng.set_spectra(emission=dict(spectral=v, brightness=ye, noise=0.2, basis=np.ones((1, S)), offset=0.0),
                        absorption=dict(spectral=v, brightness=ya, noise=0.01, basis=np.ones((1, S)), offset=0.0),
                        n_sightlines=n_sl)
        eng.set_priors()
        for k, val in options.items():
            eng.set_option(k, val)
        chains = 10
        sl = (np.arange(chains) % n_sl).astype(np.int32)
        u0 = np.zeros((chains, layout.NSLOT, N))
        u0 += 0.05 * np.random.default_rng(3).standard_normal(u0.shape)
        n0 = eng.launch_count()
        out = eng.nuts_sample(u0, n_tune=150, n_draws=60, seed=11, sightline=sl)
        out["launches"] = eng.launch_count() - n0
        eng.close()
        return out, sl

    chain, sl = run({"nuts_tick": 3})
    graph, _ = run({"nuts_tick": 2, "small_path": 2, "small_chunks": 1})
    for k in ("u", "accept", "depth", "logp", "step_size", "coeff_e", "coeff_a"):
        assert np.array_equal(chain[k], graph[k], equal_nan=True), k
    assert chain["launches"] < 300 and graph["launches"] > 20 * chain["launches"], (chain["launches"], graph["launches"])
    spec, _ = run({"nuts_tick": 6})
    for k in ("u", "accept", "depth", "logp", "step_size"):
        assert np.array_equal(spec[k], chain[k], equal_nan=True), k
    cluster, _ = run({"nuts_tick": 5})
    graph_cl, _ = run({"nuts_tick": 2, "small_path": 2, "small_chunks": 8})
    for k in ("u", "accept", "depth", "logp"):
        assert np.array_equal(cluster[k], graph_cl[k], equal_nan=True), k
    wide, _ = run({"nuts_tick": 4})
    # each chain found its own sightline's line: velocity = prior mean + 15 * velocity_norm
    idx = layout.slot_names("emission_absorption", False)["velocity_norm"]  # Normal: unconstrained = constrained
    for res in (chain, cluster, wide):
        x = res["u"][:, :, idx, 0].mean(axis=0) * 15.0
        assert np.all(np.abs(x - centres[sl]) < 1.5), (x, centres[sl])
