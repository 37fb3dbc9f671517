"""`bayes_spec` stand-in: the attributes of BaseModel the reference's classes touch
(../README.md).  The baseline polynomial follows oracle.models_ref.OracleData.predict_baseline
(restated third-party code)."""
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
        coords = {"cloud": range(n_clouds), "coeff": range(baseline_degree + 1)}
        self.model = pm.Model(coords=coords)

    def add_baseline_priors(self, prior_baseline_coeffs=None):
        with self.model:
            for key in self.data:
                sig = [1.0] * (self.baseline_degree + 1)
                if prior_baseline_coeffs is not None:
                    sig = prior_baseline_coeffs[key]
                pm.Normal(f"baseline_{key}_norm", mu=0.0, sigma=sig, dims="coeff")

    def predict_baseline(self):
        return {key: d.predict_baseline(self.model[f"baseline_{key}_norm"]) for key, d in self.data.items()}
