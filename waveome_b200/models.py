"""Model objects of the exact-GPR path: what ``gpflow.models.GPR(data=(X, Y), kernel=k)`` is to the
reference (waveome/model_fitting.py:150-155), reduced to a description the CUDA engine consumes.

A fitted ``GPR`` exposes what the reference's downstream code reads from a model:
``.kernel``, ``.likelihood.variance``, ``.mean_function.c``, ``.kernel_name``, ``.trainable_parameters``,
``.log_marginal_likelihood_value``, ``.log_posterior_density_value``.
"""
from __future__ import annotations

from typing import List, Optional

import numpy as np

from . import kernels as K
from .program import Program, build_program


class Gaussian:
    """gpflow.likelihoods.Gaussian: variance = 1e-6 + softplus(u), default 1.0."""
    name = "gaussian"

    def __init__(self, variance=1.0, variance_lower_bound=1e-6):
        self.variance = K.Parameter(variance, transform=("softplus_shift", variance_lower_bound))

    @property
    def parameters(self):
        return [self.variance]


class Poisson:
    """gpflow.likelihoods.Poisson (exp inverse link, binsize 1): no parameters."""
    name = "poisson"
    parameters: list = []

    def __init__(self):
        self._dummy_noise = K.Parameter(1.0, transform=("softplus_shift", 1e-6), trainable=False)

    @property
    def engine_param(self):
        return 0.0


class NegativeBinomial:
    """waveome/likelihoods.py:16-66: trainable dispersion ``alpha`` with an Exp bijector (:24-28), log link.  On the
    engine alpha rides in the program's noise slot (there is no Gaussian noise on this path)."""
    name = "negative_binomial"

    def __init__(self, alpha=1.0, trainable=True):
        self.alpha = K.Parameter(alpha, transform="exp", trainable=trainable)
        self._dummy_noise = self.alpha

    @property
    def parameters(self):
        return [self.alpha]

    @property
    def engine_param(self):
        return float(self.alpha)


class Bernoulli:
    """gpflow.likelihoods.Bernoulli (inv_probit link): no parameters, y in {0, 1}."""
    name = "bernoulli"
    parameters: list = []

    def __init__(self):
        self._dummy_noise = K.Parameter(1.0, transform=("softplus_shift", 1e-6), trainable=False)

    @property
    def engine_param(self):
        return 0.0


class Gamma:
    """gpflow.likelihoods.Gamma (exp link): trainable ``shape`` with the positive() bijector; it rides in the program's
    noise slot like the negative binomial's alpha."""
    name = "gamma"

    def __init__(self, shape=1.0, trainable=True):
        self.shape = K.Parameter(shape, transform="softplus", trainable=trainable)
        self._dummy_noise = self.shape

    @property
    def parameters(self):
        return [self.shape]

    @property
    def engine_param(self):
        return float(self.shape)


class Exponential(Gamma):
    """gpflow.likelihoods.Exponential (exp link, no parameters; waveome/model_fitting.py:156-162): p(y | f) is the Gamma
    density with shape 1, so it runs as the engine's Gamma likelihood with the shape frozen at 1."""

    def __init__(self):
        super().__init__(shape=1.0, trainable=False)


class ZeroInflatedNegativeBinomial:
    """waveome/likelihoods.py:96-139: NB with dispersion ``alpha`` and structural zeros with probability
    psi = km / (km + exp(f)) (Michaelis-Menten constant ``km``), both positive() parameters.  alpha rides in the
    program's noise slot, km in its second likelihood slot."""
    name = "zinb"

    def __init__(self, alpha=1.0, km=1.0, trainable=True):
        self.alpha = K.Parameter(alpha, transform="softplus", trainable=trainable)
        self.km = K.Parameter(km, transform="softplus", trainable=trainable)
        self._dummy_noise = self.alpha
        self._aux_param = self.km

    @property
    def parameters(self):
        return [self.alpha, self.km]

    @property
    def engine_param(self):
        return (float(self.alpha), float(self.km))


def make_likelihood(name, **kw):
    """gp_likelihood_crosswalk (waveome/utilities.py:989-1009) for the likelihoods the engine covers."""
    if name == "gaussian":
        return Gaussian(**kw)
    if name == "poisson":
        return Poisson()
    if name in ("negative_binomial", "negativebinomial"):
        return NegativeBinomial(**kw)
    if name in ("bernoulli", "binomial"):
        return Bernoulli()
    if name == "gamma":
        return Gamma(**kw)
    if name == "exponential":
        return Exponential()
    if name in ("zeroinflated_negativebinomial", "zero_inflated_negative_binomial", "zinb"):
        return ZeroInflatedNegativeBinomial(**kw)
    raise NotImplementedError(f"likelihood {name!r} is not covered by the B200 engine "
                              "(gaussian, poisson, negative_binomial, bernoulli, gamma, exponential, zeroinflated_negativebinomial)")


class ConstantMean:
    """gpflow.mean_functions.Constant: trainable scalar ``c`` (identity bijector), default 0."""
    name = "constant"

    def __init__(self, c=0.0):
        self.c = K.Parameter(c, transform="identity")

    @property
    def parameters(self):
        return [self.c]


class ZeroMean:
    """gpflow.mean_functions.Zero (the GPR default in waveome/model_fitting.py:151-155)."""
    name = "zero"
    parameters: list = []


class GPR:
    def __init__(self, kernel: K.Kernel, mean_function=None, noise_variance: float = 1.0, likelihood=None):
        self.kernel = kernel
        self.mean_function = mean_function if mean_function is not None else ZeroMean()
        self.likelihood = likelihood if likelihood is not None else Gaussian(noise_variance)
        self.name = "gpr"
        self.kernel_name = ""
        self.data = None
        self.log_marginal_likelihood_value = None
        self.log_posterior_density_value = None
        self.fit_info = None

    # GPflow-like surface ---------------------------------------------------------------------
    @property
    def parameters(self) -> List[K.Parameter]:
        return list(self.kernel.parameters) + list(self.likelihood.parameters) + list(self.mean_function.parameters)

    @property
    def trainable_parameters(self) -> List[K.Parameter]:
        seen, out = set(), []
        for p in self.parameters:
            if p.trainable and id(p) not in seen:
                seen.add(id(p))
                out.append(p)
        return out

    def parameter_dict(self):
        """gpflow.utilities.parameter_dict(model)-style {path: Parameter}."""
        d = dict(self.kernel.named_parameters("kernel"))
        if isinstance(self.likelihood, Gaussian):
            d[".likelihood.variance"] = self.likelihood.variance
        elif isinstance(self.likelihood, NegativeBinomial):
            d[".likelihood.alpha"] = self.likelihood.alpha
        elif isinstance(self.likelihood, Gamma):
            d[".likelihood.shape"] = self.likelihood.shape
        elif isinstance(self.likelihood, ZeroInflatedNegativeBinomial):
            d[".likelihood.alpha"] = self.likelihood.alpha
            d[".likelihood.km"] = self.likelihood.km
        if isinstance(self.mean_function, ConstantMean):
            d[".mean_function.c"] = self.mean_function.c
        return d

    # engine interface ------------------------------------------------------------------------
    def program(self) -> Program:
        mc = self.mean_function.c if isinstance(self.mean_function, ConstantMean) else None
        # count likelihoods: the program's noise slot is a frozen placeholder the engine ignores
        noise = self.likelihood.variance if isinstance(self.likelihood, Gaussian) else self.likelihood._dummy_noise
        return build_program(self.kernel, noise, mc, likelihood_aux=getattr(self.likelihood, "_aux_param", None))

    def to_spec(self) -> dict:
        """Neutral JSON-able description (the format the test oracle consumes)."""
        noise = self.likelihood.variance if isinstance(self.likelihood, Gaussian) else self.likelihood._dummy_noise
        spec = {"kernel": self.kernel.to_spec(), "likelihood_variance": noise.to_spec()}
        if not isinstance(self.likelihood, Gaussian):
            spec["likelihood"] = {"type": self.likelihood.name}
            if isinstance(self.likelihood, NegativeBinomial):
                spec["likelihood"]["alpha"] = float(self.likelihood.alpha)
            if isinstance(self.likelihood, Gamma):
                spec["likelihood"]["shape"] = float(self.likelihood.shape)
            if isinstance(self.likelihood, ZeroInflatedNegativeBinomial):
                spec["likelihood_aux"] = self.likelihood.km.to_spec()
        if isinstance(self.mean_function, ConstantMean):
            spec["mean"] = {"type": "constant", "c": self.mean_function.c.to_spec()}
        else:
            spec["mean"] = {"type": "zero"}
        return spec

    def log_posterior_density(self):
        return self.log_posterior_density_value

    def predict_f(self, Xnew, data=None):
        """gpflow GPModel.predict_f(Xnew) (full_cov=False): ([m, 1] mean, [m, 1] variance of f).  ``data`` = (X, Y) as
        the reference's model methods take it; defaults to the data the model was fitted on when it was kept.  The
        non-Gaussian likelihoods give the latent posterior under the variational q (gpflow VGP.predict_f)."""
        from .postfit import predict_f
        data = data if data is not None else self.data
        if data is None:
            raise ValueError("predict_f needs data=(X, Y): the fitted model does not keep its training data")
        X, y = np.asarray(data[0]), np.asarray(data[1]).reshape(-1)
        mu, var = predict_f(self, X, y, Xnew)
        return mu.reshape(-1, 1), var.reshape(-1, 1)

    def predict_y(self, Xnew, data=None):
        """gpflow GPModel.predict_y = likelihood.predict_mean_and_var(predict_f): the Gaussian likelihood adds its
        variance; the others integrate their conditional moments (postfit.likelihood_predict_mean_and_var)."""
        mu, var = self.predict_f(Xnew, data=data)
        if isinstance(self.likelihood, Gaussian):
            return mu, var + float(self.likelihood.variance)
        from .postfit import likelihood_predict_mean_and_var
        return likelihood_predict_mean_and_var(self.likelihood, mu, var)

    def predict_log_density(self, data_new, data=None):
        """gpflow GPModel.predict_log_density((Xnew, Ynew)) = likelihood.predict_log_density(predict_f, Ynew), [m]."""
        Xnew, Ynew = data_new
        ynew = np.asarray(Ynew, dtype=np.float64).reshape(-1, 1)
        if isinstance(self.likelihood, Gaussian):
            mu, var = self.predict_y(Xnew, data=data)
            return (-0.5 * (np.log(2 * np.pi) + np.log(var) + (ynew - mu) ** 2 / var)).reshape(-1)
        from .postfit import likelihood_predict_log_density
        mu, var = self.predict_f(Xnew, data=data)
        return likelihood_predict_log_density(self.likelihood, mu, var, ynew).reshape(-1)

    def get_feature_importances(self, data=None, return_value="log_bf"):
        """waveome/model_classes.py:546-573"""
        from .postfit import feature_importances_batch
        data = data if data is not None else self.data
        X, y = np.asarray(data[0]), np.asarray(data[1]).reshape(1, -1)
        self.feature_importances = feature_importances_batch(X, y, [self], return_value=return_value)[0]
        return None

    def log_marginal_likelihood(self):
        return self.log_marginal_likelihood_value


# ------------------------------------------------------------------------------------------------
# penalised model: what PSVGP(penalized_options={"penalization_factor": pf}) is to the reference
# ------------------------------------------------------------------------------------------------
class PenalizedGPR(GPR):
    """Exact-GPR counterpart of ``PenalizedGP`` (waveome/model_classes.py:777-1079): horseshoe prior with
    scale 1/penalization_factor on every trainable kernel variance (:837-864), structure pruning by
    ``cut_kernel_components`` (:1029-1079)."""

    def __init__(self, kernel, mean_function=None, noise_variance=1.0, penalization_factor=1.0, likelihood=None):
        super().__init__(kernel, mean_function=mean_function, noise_variance=noise_variance, likelihood=likelihood)
        self.name = "penalized_gpr"
        self.feature_importances = None
        self.set_penalization_factor(penalization_factor)
        self.update_kernel_name()

    def set_penalization_factor(self, penalization_factor, use_prior=True):
        self.penalization_factor = float(penalization_factor)
        if use_prior:
            prior = K.Horseshoe(scale=1.0 / penalization_factor) if penalization_factor > 0 else None
            for key, val in self.parameter_dict().items():
                if "kernel" in key and "variance" in key:
                    val.prior = prior

    def penalization_search(self, data=None, penalization_factor_list=(0.0, 1.0, 10.0, 100.0), k_fold=3, fit_best=True,
                            random_seed=None, num_restart=5, selection_type="se", unit_col=None, **unused):
        """waveome/model_classes.py:866-998 for this model: k-fold cross-validation over the factors, then (fit_best) the
        model takes the best factor and the parameters of the refit on all rows.  Sets
        ``self.penalization_search_results`` ([factors x folds, 3]: factor, fold, held-out mean log density)."""
        from .penalization import penalization_search_batch
        X, y = np.asarray(data[0], dtype=np.float64), np.asarray(data[1], dtype=np.float64).reshape(1, -1)
        out = penalization_search_batch(X, y, self.kernel, mean_function=self.mean_function,
                                        penalization_factor_list=penalization_factor_list, k_fold=k_fold,
                                        unit_col=unit_col, fit_best=fit_best, random_seed=random_seed,
                                        num_restart=num_restart, selection_type=selection_type)
        res = out["results"][0]
        self.penalization_search_results = np.array([[pf, k, res[fi, k]] for fi, pf in enumerate(out["factors"])
                                                     for k in range(res.shape[1])])
        if fit_best:
            best = out["models"][0]
            self.set_penalization_factor(float(out["best_factor"][0]) if np.isfinite(out["best_factor"][0]) else 0.0)
            self.kernel, self.mean_function, self.likelihood = best.kernel, best.mean_function, best.likelihood
            self.set_penalization_factor(self.penalization_factor)
            self.log_marginal_likelihood_value = best.log_marginal_likelihood_value
            self.log_posterior_density_value = best.log_posterior_density_value
            self.fit_info = best.fit_info
        return None

    def update_kernel_name(self):
        from .utilities import kernel_name_string
        self.kernel_name = kernel_name_string(self.kernel, with_idx=True)

    def cut_kernel_components(self, X, var_cutoff: float = 0.1):
        from .utilities import find_variance_components, search_through_kernel_list_
        var_parts = find_variance_components(self.kernel, sum_reduce=False)
        var_flag = np.where(np.asarray(var_parts).reshape(-1) >= var_cutoff)[0]
        if len(var_flag) > 1:
            self.kernel = K.Sum([self.kernel.kernels[i] for i in var_flag])
        elif len(var_flag) == 1:
            if len(var_parts) > 1:
                self.kernel = self.kernel.kernels[var_flag[0]]
        else:
            self.kernel = K.Constant()
        if hasattr(self.kernel, "kernels"):
            self.kernel = search_through_kernel_list_(self.kernel.kernels, list_type=self.kernel.name, X=np.asarray(X))
        return None
