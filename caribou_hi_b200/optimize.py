"""``Optimize``: the n_clouds sweep of ``bayes_spec.Optimize`` (third-party; usage in the reference:
docs/source/notebooks/optimization.ipynb cells 8-13) on the device sampler.

One model per cloud count 1..max_n_clouds, every model sampled with the chain-batched NUTS of the
library (``model.sample`` -> ``chi_nuts_sample``), the Bayesian information criterion of each from the
largest likelihood among its posterior draws, early stop once adding clouds stops paying, ``best_model``
= the smallest BIC.  The sweep is one more independent axis: the models are spread
round-robin over the given GPUs and sampled at the same time, one host thread per GPU (the C call
releases the GIL, every handle owns its own stream).

bayes_spec's own BIC / stopping rule are third-party and not installable here; what is implemented is
their published definition -- BIC = k ln(n) - 2 ln L_max with k free parameters and n data points, the
null hypothesis = baseline only -- restated, not executed (parity unpinned, as for the other
third-party pieces).
"""

from concurrent.futures import ThreadPoolExecutor

import numpy as np


class Optimize:
    """``Optimize(model_type, data, max_n_clouds=5, baseline_degree=0, seed=1234, verbose=False, **model_kwargs)``
    as in the reference notebook; ``devices`` lists the CUDA devices to spread the models over."""

    def __init__(self, model_type, data, max_n_clouds=5, baseline_degree=0, seed=1234, verbose=False, devices=(0,),
                 **model_kwargs):
        if max_n_clouds < 1:
            raise ValueError("max_n_clouds must be >= 1")
        self.model_type = model_type
        self.data = data
        self.max_n_clouds = int(max_n_clouds)
        self.verbose = verbose
        self.seed = seed
        self.devices = tuple(devices)
        self.n_clouds = list(range(1, self.max_n_clouds + 1))
        self.models = {n: model_type(data, n, baseline_degree=baseline_degree, seed=seed, verbose=verbose,
                                     device=self.devices[(n - 1) % len(self.devices)], **model_kwargs)
                       for n in self.n_clouds}
        self.best_model = None
        self.bics = {}
        self.results = {}

    def add_priors(self, *args, **kwargs):
        for m in self.models.values():
            m.add_priors(*args, **kwargs)

    def add_likelihood(self, *args, **kwargs):
        for m in self.models.values():
            m.add_likelihood(*args, **kwargs)

    # ---- information criterion -----------------------------------------------------------------
    def n_data(self):
        return int(sum(np.asarray(d.brightness).size for d in self.data.values()))

    def null_bic(self):
        """BIC of the no-cloud hypothesis: every spectrum explained by its baseline offset alone, no free
        parameter (the "Null hypothesis BIC" line of optimization.ipynb cell 12)."""
        lnl = 0.0
        for d in self.data.values():
            y, s = np.asarray(d.brightness, dtype=float), np.broadcast_to(np.asarray(d.noise, dtype=float), np.shape(d.brightness))
            r = (y - np.median(y)) / s
            lnl += float(np.sum(-0.5 * r * r - np.log(s) - 0.5 * np.log(2.0 * np.pi)))
        return -2.0 * lnl

    @staticmethod
    def n_parameters(model):
        """Free parameters of a model: its per-cloud RVs (per-draw scalars once) and baseline coefficients."""
        from . import layout

        scal = layout.scalar_rvs(model.kind)
        k = sum(1 if name in scal else model.n_clouds for name in model.engine.slots)
        return k + len(model._required_keys) * (model.baseline_degree + 1)

    def bic(self, model, posterior):
        """k ln(n) - 2 max_draws ln L over the posterior draws ``[chains, draws, ...]`` of ``model``."""
        some = next(iter(posterior.values()))
        chains, draws = some.shape[:2]
        point = {k: np.ascontiguousarray(v.reshape((chains * draws,) + v.shape[2:])) for k, v in posterior.items()}
        lnl = model.logp(point)
        best = np.nanmax(np.where(np.isfinite(lnl), lnl, -np.inf))
        return self.n_parameters(model) * np.log(self.n_data()) - 2.0 * float(best)

    # ---- the sweep --------------------------------------------------------------------------------
    def _fit_one(self, n, sample_kwargs):
        m = self.models[n]
        post, stats = m.sample(**sample_kwargs)
        return n, post, stats, self.bic(m, post)

    def optimize(self, bic_threshold=10.0, sample_kwargs=None, **_ignored):
        """Sample the models in order of cloud count, ``len(devices)`` at a time, and stop when the last two
        counts both failed to improve the best BIC by ``bic_threshold`` (the notebook's run stops the same
        way: n = 4 and n = 5 after the best n = 3).  ``sample_kwargs`` go to ``model.sample``
        (``draws, tune, chains, seed, target_accept, ...``); PyMC-only keys (``cores``, ``init_kwargs``,
        ``nuts_kwargs``) are translated or dropped.  Returns ``best_model``."""
        kw = dict(sample_kwargs or {})
        nuts = kw.pop("nuts_kwargs", None) or {}
        for k in ("cores", "init_kwargs", "progressbar"):
            kw.pop(k, None)
        if "target_accept" in nuts:
            kw.setdefault("target_accept", nuts["target_accept"])
        self.bics = {0: self.null_bic()}
        if self.verbose:
            print(f"Null hypothesis BIC = {self.bics[0]:.3e}")
        best_n, worse = 0, 0
        width = max(1, len(self.devices))
        with ThreadPoolExecutor(max_workers=width) as pool:
            for start in range(0, len(self.n_clouds), width):
                wave = self.n_clouds[start : start + width]
                for n, post, stats, bic in pool.map(lambda n: self._fit_one(n, kw), wave):
                    self.results[n] = {"posterior": post, "stats": stats}
                    self.bics[n] = bic
                    if self.verbose:
                        print(f"n_cloud = {n} BIC = {bic:.3e} (divergent {float(np.mean(stats['diverging'])):.3f})")
                    if bic < self.bics[best_n] - bic_threshold:
                        best_n, worse = n, 0
                    else:
                        worse += 1
                if worse >= 2:
                    if self.verbose:
                        print("Stopping criteria met.")
                    break
        self.best_model = self.models[best_n] if best_n > 0 else None
        return self.best_model
