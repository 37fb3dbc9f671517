"""`bayes_spec` stand-in: the attributes of BaseModel the reference's classes touch (README.md); the
baseline polynomial is the oracle's restatement (third-party, unverified) on symbolic coefficients."""
import pymc as pm


class BaseModel:
    def __init__(self, data, n_clouds, baseline_degree=0, seed=1234, verbose=False):
        self.data = data
        self.n_clouds = n_clouds
        self.baseline_degree = baseline_degree
        self.seed = seed
        self.verbose = verbose
        self._cluster_features = []
        self.var_name_map = {}
        self.model = pm.Model(coords={"cloud": range(n_clouds), "coeff": range(baseline_degree + 1)})

    def add_baseline_priors(self, prior_baseline_coeffs=None):
        with self.model:
            for key in self.data:
                pm.Normal(f"baseline_{key}_norm", mu=0.0, sigma=1.0, dims="coeff")

    def predict_baseline(self):
        out = {}
        for key, d in self.data.items():
            c = self.model[f"baseline_{key}_norm"]
            total = None
            for i in range(self.baseline_degree + 1):
                term = c[i : i + 1] / (i + 1.0) ** i * d.spectral_norm**i
                total = term if total is None else total + term
            out[key] = total * d.brightness_scale + d.brightness_offset
        return out
