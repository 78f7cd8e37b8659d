"""Small latent / observation models with the reference's `LatentModel` / `ObservationLikelihood` interface (TEST
infrastructure: the models are producers of Q and stay in Julia). Recipes:
* AR1 -- src/latent_models/ar.jl:135-150: tridiagonal, diag tau (ends) / (1 + rho^2) tau, off-diagonal -rho tau;
* RW1 -- src/latent_models/rw.jl:170-185 (+ sum-to-zero null-space constraint :204-221, regularization 1e-5);
* IID -- src/latent_models/iid.jl:74-90 (tau I, `precision_logdet` = n log tau);
* Matern -- the SPDE recipe of gmrf_b200/spde.py wrapped in the interface, with its value basis;
* LinearGaussianLikelihood -- y ~ N(A x, sigma^2 I) (LinearlyTransformedObservationModel over Normal,
  test/workspace/test_workspace_latent_models.jl:132-165): the non-diagonal Hessian -A'A / sigma^2."""
import numpy as np
import scipy.sparse as sp


class AR1Model:
    def __init__(self, n):
        self.n = n

    def precision_matrix(self, tau, rho):
        if tau <= 0 or abs(rho) >= 1:
            raise ValueError("AR1 needs tau > 0 and |rho| < 1")
        d = np.full(self.n, (1 + rho ** 2) * tau)
        d[0] = d[-1] = tau
        e = np.full(self.n - 1, -rho * tau)
        return sp.diags([e, d, e], [-1, 0, 1]).tocsc()

    def mean(self, **_):
        return np.zeros(self.n)

    def constraints(self, **_):
        return None

    def precision_logdet(self, tau, rho):
        return self.n * np.log(tau) + np.log1p(-rho ** 2)        # det = tau^n (1 - rho^2)


class RW1Model:
    def __init__(self, n, regularization=1e-5):
        self.n, self.regularization = n, regularization

    def precision_matrix(self, tau):
        d = np.full(self.n, 2.0 * tau + self.regularization)
        d[0] = d[-1] = tau + self.regularization
        e = np.full(self.n - 1, -tau)
        return sp.diags([e, d, e], [-1, 0, 1]).tocsc()

    def mean(self, **_):
        return np.zeros(self.n)

    def constraints(self, **_):
        return np.ones((1, self.n)), np.zeros(1)


class IIDModel:
    def __init__(self, n):
        self.n = n

    def precision_matrix(self, tau):
        return (tau * sp.identity(self.n)).tocsc()

    def mean(self, **_):
        return np.zeros(self.n)

    def constraints(self, **_):
        return None

    def precision_logdet(self, tau):
        return self.n * np.log(tau)


class MaternModel:
    """MaternSPDE behind the LatentModel interface; exposes the value basis for device-side assembly."""

    def __init__(self, spde_model):
        self.m = spde_model
        self.n = spde_model.n

    def precision_matrix(self, tau, range_):
        return self.m.precision(tau, range_)

    def mean(self, **_):
        return np.zeros(self.n)

    def constraints(self, **_):
        return None

    def basis(self):
        return self.m.basis()

    def coefficients(self, tau, range_):
        return self.m.coefficients(tau, range_)


class LinearGaussianLikelihood:
    def __init__(self, A, y, sigma):
        self.A = sp.csr_matrix(A, dtype=np.float64)
        self.y = np.asarray(y, dtype=np.float64)
        self.sigma = float(sigma)
        self._H = sp.csc_matrix(-(self.A.T @ self.A) / self.sigma ** 2)
        self._H.sort_indices()

    def loglik(self, x):
        r = self.y - self.A @ x
        return float(-0.5 * (r @ r) / self.sigma ** 2 - self.y.size * np.log(self.sigma) - 0.5 * self.y.size * np.log(2 * np.pi))

    def loggrad(self, x):
        return self.A.T @ (self.y - self.A @ x) / self.sigma ** 2

    def loghessian(self, x):
        return self._H


class BernoulliLikelihood:
    """y_i ~ Bernoulli(logistic(x_i)) (canonical logit link; ExponentialFamily(Bernoulli) of the reference's tests)."""

    def __init__(self, y):
        self.y = np.asarray(y, dtype=np.float64)

    def loglik(self, x):
        return float(np.sum(self.y * x - np.logaddexp(0.0, x)))

    def loggrad(self, x):
        return self.y - 1.0 / (1.0 + np.exp(-x))

    def loghessian(self, x):
        p = 1.0 / (1.0 + np.exp(-x))
        return -p * (1.0 - p)
