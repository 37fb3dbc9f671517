"""Host-side mirror of the reference's six model classes for the likelihood path.

Same class names, constructor arguments, ``add_priors`` keyword names/defaults and
``add_likelihood()`` call order as ``caribou_hi/*_model.py`` (cited per class), but
instead of building a PyMC graph they configure one ``libchi`` handle: the
parameter maps, spin temperature, line profiles, radiative transfer, Normal logp and
its gradient all run in the CUDA library (``Engine``), batched over draws.

Only the likelihood path is mirrored: priors are recorded as hyper-parameters (the
maps need them) but prior densities, samplers and posterior tooling stay with
PyMC/bayes_spec (out of scope).
"""

from abc import ABC, abstractmethod

import numpy as np

from . import layout, synthetic
from .engine import Engine


class _HIBase(ABC):
    """Common machinery of HIModel (hi_model.py:21-167) and HIPhysicalModel
    (hi_physical_model.py:21-225): stores data/n_clouds/baseline_degree like
    ``bayes_spec.BaseModel.__init__`` and the priors of ``add_priors``."""

    kind = None  # set by concrete classes
    _physical = False

    def __init__(self, data, n_clouds, baseline_degree=0, seed=1234, verbose=False, device=0):
        self.data = data
        self.n_clouds = int(n_clouds)
        self.baseline_degree = int(baseline_degree)
        self.seed = seed
        self.verbose = verbose
        self.device = device
        self._cluster_features = ["velocity", "fwhm2"]  # hi_model.py:30-33
        self._priors = None
        self.engine = None

    # -- priors ---------------------------------------------------------------
    def _base_priors(self, prior_fwhm2, prior_velocity, prior_log10_n_alpha, prior_fwhm_L,
                     prior_baseline_coeffs, **extra):
        self._priors = dict(
            prior_fwhm2=float(prior_fwhm2),
            prior_velocity=tuple(map(float, prior_velocity)),
            prior_log10_n_alpha=tuple(map(float, prior_log10_n_alpha)),
            prior_fwhm_L=None if prior_fwhm_L is None else float(prior_fwhm_L),
            prior_baseline_coeffs=prior_baseline_coeffs,
        )
        self._priors.update(extra)

    @abstractmethod
    def add_likelihood(self, *args, **kwargs):  # pragma: no cover
        """Must be defined in inherited class (hi_model.py:164-167)."""

    # -- engine -----------------------------------------------------------------
    _amplitude_prior = None  # name of the prior scaling CHI_FR_AMPLITUDE_NORM
    _required_keys = ()

    def _build_engine(self, free_weight=True):
        if self._priors is None:
            raise RuntimeError("call add_priors() before add_likelihood()")
        for key in self._required_keys:
            if key not in self.data:
                raise KeyError(f'SpecData key must include "{key}"')
        p = self._priors
        self.engine = Engine(
            self.kind,
            self.n_clouds,
            device=self.device,
            lorentzian=p["prior_fwhm_L"] is not None,
            free_weight=free_weight,
            bg_temp=getattr(self, "bg_temp", 3.77),
            depth_nth_fwhm_power=getattr(self, "depth_nth_fwhm_power", 1.0 / 3.0),
            prior_fwhm2=p["prior_fwhm2"],
            prior_velocity=p["prior_velocity"],
            prior_log10_nHI=p.get("prior_log10_nHI", (0.0, 1.5)),
            prior_log10_n_alpha=p["prior_log10_n_alpha"],
            prior_fwhm_L=p["prior_fwhm_L"] or 0.0,
            prior_amplitude=p[self._amplitude_prior],
        )
        sets = {}
        for key in ("emission", "absorption"):
            if key in self._required_keys:
                sets[key] = self.data[key].as_dataset(self.baseline_degree)
        self.engine.set_spectra(emission=sets.get("emission"), absorption=sets.get("absorption"))
        # prior densities + default transforms for the unconstrained-space model logp
        bc = p.get("prior_baseline_coeffs") or {}
        self.engine.set_priors(
            beta=p.get("prior_fwhm2_thermal_fraction" if self._physical else "prior_tkin_factor", (2.0, 2.0)),
            nth_fwhm_1pc=p.get("prior_nth_fwhm_1pc", (1.75, 0.25)),
            sigma_log10_NHI=p.get("prior_sigma_log10_NHI") or 0.5,
            baseline_sigma_e=bc.get("emission"),
            baseline_sigma_a=bc.get("absorption"),
        )

    # -- unconstrained space (what PyMC's NUTS works in) -------------------------------
    @property
    def value_vars(self):
        """PyMC's names of the unconstrained value variables, in ``free_RVs`` order."""
        return [layout.value_name(k) for k in self.free_RVs]

    def _strip(self, point):
        """Accept either RV names or PyMC value-variable names as keys."""
        back = {layout.value_name(k): k for k in self.engine.slots}
        return {back.get(k, k): v for k, v in point.items()}

    def to_unconstrained(self, point):
        """Free-RV values -> unconstrained values (keys: PyMC value-variable names)."""
        theta, ce, ca, _ = self._split(point)
        u = self.engine.transform(theta, -1)
        out = {layout.value_name(k): (u[:, s, :1] if k in layout.scalar_rvs(self.kind) else u[:, s, :])
               for k, s in self.engine.slots.items()}
        out.update({k: v for k, v in point.items() if k.startswith("baseline_")})
        return out

    def to_constrained(self, point):
        point = self._strip(point)
        theta, ce, ca, _ = self._split(point)
        x = self.engine.transform(theta, +1)
        out = {k: (x[:, s, :1] if k in layout.scalar_rvs(self.kind) else x[:, s, :])
               for k, s in self.engine.slots.items()}
        out.update({k: v for k, v in point.items() if k.startswith("baseline_")})
        return out

    def model_logp_dlogp(self, point):
        """Total model logp (likelihood + priors + Jacobians) ``[B]`` and its gradient for a
        dict of UNCONSTRAINED values -- PyMC's ``model.logp_dlogp_function`` per draw."""
        point = self._strip(point)
        theta, ce, ca, _ = self._split(point)
        logp, grad, gce, gca = self.engine.model_logp_grad(theta, ce, ca)
        grads = {layout.value_name(k): v for k, v in self.engine.unpack_grad(grad).items()}
        if gce is not None:
            grads["baseline_emission_norm"] = gce
        if gca is not None:
            grads["baseline_absorption_norm"] = gca
        return logp, grads

    def sample(self, draws=1000, tune=1000, chains=8, seed=None, target_accept=0.8, max_treedepth=10,
               init_point=None, jitter=None, pool_mass=False, lock_step=False):
        """Batched NUTS over ``chains`` on the device -- the role of ``BaseModel.sample`` /
        ``pm.sample`` in the reference (optimization.ipynb cell 12), but with all chains in one
        process: small models run one resident thread block per chain that does whole leapfrog steps
        inside one kernel (``csrc/nuts_chain.cuh``), larger ones a CUDA graph of the chains' state
        machine + one batched logp+grad kernel pipeline per leapfrog step
        (``Engine.set_option("nuts_tick", ...)`` forces either).

        Starting points follow PyMC's "jitter" initialisation: the model's initial point (the
        origin of the unconstrained space, i.e. the moment of each prior -- or ``init_point``, a
        dict of constrained values such as an optimiser/ADVI result) plus Uniform(-jitter, jitter)
        noise per chain in the unconstrained space, redrawn for chains whose model logp is not
        finite.  Returns ``(posterior, stats)``: dicts of ``[chains, draws, ...]`` arrays keyed by
        free-RV name (constrained values) and by statistic."""
        seed = self.seed if seed is None else seed
        rng = np.random.default_rng(seed)
        eng = self.engine
        N, nb = self.n_clouds, self.baseline_degree + 1
        if init_point is None:
            base = {}
            for k in eng.slots:
                shape = (1, 1) if k in layout.scalar_rvs(self.kind) else (1, N)
                if k == "nth_fwhm_1pc":  # TruncatedNormal: start at its location parameter
                    base[k] = np.full(shape, np.log(self._priors.get("prior_nth_fwhm_1pc", (1.75, 0.25))[0]))
                else:
                    base[k] = np.zeros(shape)
            for key in self._required_keys:
                base[f"baseline_{key}_norm"] = np.zeros((1, nb))
            # the coupled Normal (log10_wt_ff_*) starts at its prior mean given the other values
            wt = [k for k in eng.slots if k.startswith("log10_wt_ff")]
            if wt:
                x0 = self.to_constrained({layout.value_name(k): v for k, v in base.items()})
                det = self.deterministics(x0)
                T = det["fwhm2_thermal" if self._physical else "tkin"]
                sig = self._priors.get("prior_sigma_log10_NHI") or 0.5
                base[wt[0]] = -np.log(10.0) * sig**2 / 2.0 - np.log10(T)
        else:
            base = self._strip(self.to_unconstrained(init_point))
        if jitter is None:
            jitter = 1.0 if init_point is None else 0.1

        def jittered(scale):
            out = {}
            for k, v in base.items():
                v = np.asarray(v, dtype=np.float64)
                full = np.broadcast_to(v, (chains,) + v.shape[1:]) if v.shape[0] in (1, chains) else v
                out[k] = full + rng.uniform(-scale, scale, size=full.shape)
            return out

        start, scale = jittered(jitter), jitter
        lp = self.model_logp_dlogp(start)[0]
        for _ in range(30):
            bad = ~np.isfinite(lp)
            if not bad.any():
                break
            scale *= 0.8
            cand = jittered(scale)
            for k in start:
                start[k][bad] = cand[k][bad]
            lp = self.model_logp_dlogp(start)[0]
        else:
            raise RuntimeError("could not find starting points with finite model logp")
        start = self._strip(start)
        u0, ce, ca, _ = self._split(start)
        res = eng.nuts_sample(u0, ce, ca, n_tune=tune, n_draws=draws, seed=int(rng.integers(1 << 62)),
                              target_accept=target_accept, max_treedepth=max_treedepth, pool_mass=pool_mass, lock_step=lock_step)
        x = eng.transform(res["u"].reshape(draws * chains, layout.NSLOT, self.n_clouds), +1)
        x = x.reshape(draws, chains, layout.NSLOT, self.n_clouds).transpose(1, 0, 2, 3)
        post = {k: (x[:, :, s, :1] if k in layout.scalar_rvs(self.kind) else x[:, :, s, :])
                for k, s in eng.slots.items()}
        if res["coeff_e"] is not None:
            post["baseline_emission_norm"] = res["coeff_e"].transpose(1, 0, 2)
        if res["coeff_a"] is not None:
            post["baseline_absorption_norm"] = res["coeff_a"].transpose(1, 0, 2)
        stats = {k: res[k].T for k in ("accept", "depth", "diverging", "logp", "step_size")}
        return post, stats

    # -- evaluation ---------------------------------------------------------------
    @property
    def free_RVs(self):
        """Names of the free RVs (reference naming), baseline coefficients last."""
        names = list(self.engine.slots)
        return names + [f"baseline_{k}_norm" for k in self._required_keys]

    def _split(self, point):
        cloud = {k: v for k, v in point.items() if not k.startswith("baseline_")}
        theta = self.engine.pack(cloud)
        B = theta.shape[0]
        ce = point.get("baseline_emission_norm")
        ca = point.get("baseline_absorption_norm")
        return theta, ce, ca, B

    def logp_dlogp(self, point):
        """Likelihood logp ``[B]`` and its gradient (dict keyed like ``point``) for a
        dict of free-RV arrays -- one call of the reference's compiled
        ``logp_dlogp_function`` restricted to the observed nodes, per draw."""
        theta, ce, ca, _ = self._split(point)
        logp, grad, gce, gca = self.engine.logp_grad(theta, ce, ca)
        grads = self.engine.unpack_grad(grad)
        if gce is not None:
            grads["baseline_emission_norm"] = gce
        if gca is not None:
            grads["baseline_absorption_norm"] = gca
        return logp, grads

    def logp(self, point):
        theta, ce, ca, _ = self._split(point)
        return self.engine.logp_grad(theta, ce, ca, want_grad=False)[0]

    def predict(self, point, dtype=np.float64):
        """Predicted spectra per draw: the ``mu`` of the observed Normal nodes.  ``dtype=np.float32``
        takes the single-precision kernels (1e-5 of each spectrum's largest value, half the bytes)."""
        theta, ce, ca, _ = self._split(point)
        mu_e, mu_a = self.engine.spectra(theta, ce, ca, dtype=dtype)
        out = {}
        if mu_e is not None:
            out["emission"] = mu_e
        if mu_a is not None:
            out["absorption"] = mu_a
        return out

    def sample_predictive(self, point, seed=1234, dtype=np.float64):
        """Draws of the observed nodes at the given parameter draws -- ``pm.sample_posterior_predictive``
        (posterior draws) / ``pm.sample_prior_predictive`` (prior draws) for the ``pm.Normal(mu,
        sigma=data.noise)`` likelihood nodes: ``{"emission": [B, S_e], "absorption": [B, S_a]}``."""
        theta, ce, ca, _ = self._split(point)
        y_e, y_a = self.engine.predictive(theta, ce, ca, seed=seed, dtype=dtype)
        out = {}
        if y_e is not None:
            out["emission"] = y_e
        if y_a is not None:
            out["absorption"] = y_a
        return out

    def deterministics(self, point):
        theta, _, _, _ = self._split(point)
        det = self.engine.deterministics(theta)
        return {k: v for k, v in det.items() if not np.all(np.isnan(v))}

    def sample_prior(self, draws, rng=None):
        """Draw free RVs from the default prior families (host NumPy; the thermal
        quantity the weight prior needs comes from the CUDA deterministics)."""
        rng = np.random.default_rng(self.seed) if rng is None else rng
        eng = self.engine
        key = "fwhm2_thermal" if self._physical else "tkin"

        def thermal(trial):
            return eng.deterministics(eng.pack(trial))[key]

        free = synthetic.draw_free(
            self.kind, self.n_clouds, draws, rng,
            lorentzian=eng.lorentzian, free_weight=eng.free_weight,
            sigma_log10_NHI=self._priors.get("prior_sigma_log10_NHI") or 0.5,
            prior_nth_fwhm_1pc=self._priors.get("prior_nth_fwhm_1pc", (1.75, 0.25)),
            thermal_fn=thermal,
        )
        for k in self._required_keys:
            free[f"baseline_{k}_norm"] = rng.standard_normal(size=(draws, self.baseline_degree + 1))
        return free

    def _validate(self):
        """Like ``BaseModel._validate`` in spirit: the likelihood at a prior draw is
        finite."""
        if self.engine is None:
            raise RuntimeError("call add_likelihood() first")
        point = self.sample_prior(4)
        return bool(np.all(np.isfinite(self.logp(point))))


class HIModel(_HIBase):
    """hi_model.py:21 -- observational parameterisation (log10_nHI is free)."""

    def add_priors(self, prior_fwhm2=500.0, prior_log10_nHI=(0.0, 1.5), prior_velocity=(0.0, 10.0),
                   prior_log10_n_alpha=(-6.0, 2.0), prior_fwhm_L=None, prior_baseline_coeffs=None, **extra):
        """hi_model.py:57-136."""
        self._base_priors(prior_fwhm2, prior_velocity, prior_log10_n_alpha, prior_fwhm_L,
                          prior_baseline_coeffs, prior_log10_nHI=tuple(map(float, prior_log10_nHI)), **extra)


class HIPhysicalModel(_HIBase):
    """hi_physical_model.py:21 -- physical parameterisation (depth from the
    size-linewidth relation)."""

    _physical = True

    def __init__(self, *args, depth_nth_fwhm_power=1 / 3, **kwargs):
        super().__init__(*args, **kwargs)
        self.depth_nth_fwhm_power = depth_nth_fwhm_power  # hi_physical_model.py:36

    def add_priors(self, prior_fwhm2=500.0, prior_velocity=(0.0, 10.0), prior_log10_n_alpha=(-6.0, 2.0),
                   prior_nth_fwhm_1pc=(1.75, 0.25), prior_fwhm_L=None, prior_baseline_coeffs=None, **extra):
        """hi_physical_model.py:70-151."""
        self._base_priors(prior_fwhm2, prior_velocity, prior_log10_n_alpha, prior_fwhm_L,
                          prior_baseline_coeffs, prior_nth_fwhm_1pc=tuple(map(float, prior_nth_fwhm_1pc)),
                          **extra)


class AbsorptionModel(HIModel):
    """absorption_model.py:19. SpecData key must be "absorption"."""

    kind = "absorption"
    _amplitude_prior = "prior_tau_total"
    _required_keys = ("absorption",)

    def add_priors(self, prior_tau_total=10.0, prior_tkin_factor=(2.0, 2.0), *args, **kwargs):
        """absorption_model.py:22-72."""
        super().add_priors(*args, prior_tau_total=float(prior_tau_total),
                           prior_tkin_factor=tuple(prior_tkin_factor), **kwargs)

    def add_likelihood(self):
        """absorption_model.py:77-104."""
        self._build_engine()


class EmissionModel(HIModel):
    """emission_model.py:21. SpecData key must be "emission"."""

    kind = "emission"
    _amplitude_prior = "prior_TB_fwhm"
    _required_keys = ("emission",)

    def __init__(self, *args, bg_temp=3.77, **kwargs):
        super().__init__(*args, **kwargs)
        self.bg_temp = bg_temp  # emission_model.py:36

    def add_priors(self, prior_TB_fwhm=1000.0, prior_tkin_factor=(2.0, 2.0), *args, **kwargs):
        """emission_model.py:38-132."""
        super().add_priors(*args, prior_TB_fwhm=float(prior_TB_fwhm),
                           prior_tkin_factor=tuple(prior_tkin_factor), **kwargs)

    def add_likelihood(self):
        """emission_model.py:137-169."""
        self._build_engine()


class EmissionAbsorptionModel(EmissionModel):
    """emission_absorption_model.py:20. SpecData keys "emission" and "absorption"."""

    kind = "emission_absorption"
    _required_keys = ("emission", "absorption")

    def add_priors(self, prior_sigma_log10_NHI=None, *args, **kwargs):
        """emission_absorption_model.py:23-64: ``None`` fixes absorption_weight = 1."""
        super().add_priors(*args, prior_sigma_log10_NHI=prior_sigma_log10_NHI, **kwargs)

    def add_likelihood(self):
        """emission_absorption_model.py:66-119."""
        self._build_engine(free_weight=self._priors["prior_sigma_log10_NHI"] is not None)


class AbsorptionPhysicalModel(HIPhysicalModel):
    """absorption_physical_model.py:19. SpecData key must be "absorption"."""

    kind = "absorption_physical"
    _amplitude_prior = "prior_NHI_fwhm2_thermal"
    _required_keys = ("absorption",)

    def add_priors(self, prior_NHI_fwhm2_thermal=1.0e20, prior_fwhm2_thermal_fraction=(2.0, 2.0), *args, **kwargs):
        """absorption_physical_model.py:22-80."""
        super().add_priors(*args, prior_NHI_fwhm2_thermal=float(prior_NHI_fwhm2_thermal),
                           prior_fwhm2_thermal_fraction=tuple(prior_fwhm2_thermal_fraction), **kwargs)

    def add_likelihood(self):
        """absorption_physical_model.py:82-109."""
        self._build_engine()


class EmissionPhysicalModel(HIPhysicalModel):
    """emission_physical_model.py:21. SpecData key must be "emission"."""

    kind = "emission_physical"
    _amplitude_prior = "prior_ff_NHI"
    _required_keys = ("emission",)

    def __init__(self, *args, bg_temp=3.77, **kwargs):
        super().__init__(*args, **kwargs)
        self.bg_temp = bg_temp  # emission_physical_model.py:36

    def add_priors(self, prior_ff_NHI=1.0e21, prior_fwhm2_thermal_fraction=(2.0, 2.0), *args, **kwargs):
        """emission_physical_model.py:38-132."""
        super().add_priors(*args, prior_ff_NHI=float(prior_ff_NHI),
                           prior_fwhm2_thermal_fraction=tuple(prior_fwhm2_thermal_fraction), **kwargs)

    def add_likelihood(self):
        """emission_physical_model.py:134-166."""
        self._build_engine()


class EmissionAbsorptionPhysicalModel(EmissionPhysicalModel):
    """emission_absorption_physical_model.py:19. SpecData keys "emission" and
    "absorption"."""

    kind = "emission_absorption_physical"
    _required_keys = ("emission", "absorption")

    def add_priors(self, prior_sigma_log10_NHI=0.5, *args, **kwargs):
        """emission_absorption_physical_model.py:22-60."""
        super().add_priors(*args, prior_sigma_log10_NHI=float(prior_sigma_log10_NHI), **kwargs)

    def add_likelihood(self):
        """emission_absorption_physical_model.py:62-115."""
        self._build_engine()
