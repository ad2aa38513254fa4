"""Surrogate-model containers -- API of mpc4quantum/model.py.

Only the read-only container ``DMDc`` (model.py:7-103) is on the MPC hot path.  The data-driven fitting classes
(``DiscrepDMDc`` model.py:109-213, ``OnlineDMDc`` model.py:216-313) are reached only through
``mpc(streaming=True)``; their names exist so that scripts import cleanly, fitting raises NotImplementedError.
"""
import numpy as np


class DMDc:
    """y = A_x x + A_u u with A = [A_x | A_u] (model.py:11-31, 81-103)."""

    def __init__(self, dim_y, dim_x, dim_u, A0):
        self.dim_y = dim_y
        self.dim_x = dim_x
        self.dim_u = dim_u
        self.A = A0
        self.discount = 1
        self.rcond = 1e-15

    @classmethod
    def from_data(cls, Y, X, U, **kwargs):
        raise NotImplementedError()

    @classmethod
    def from_bootstrap(cls, dim_y, dim_x, dim_u, A0, **kwargs):
        raise NotImplementedError()

    @classmethod
    def from_randn(cls, dim_y, dim_x, dim_u, **kwargs):
        raise NotImplementedError()

    def fit_iteration(self, next_y, next_x, next_u):
        raise NotImplementedError()

    def predict(self, current_x, current_u):
        A_x, A_u = self.get_discrete()
        return A_x @ np.reshape(current_x, (self.dim_x, -1)) + A_u @ np.reshape(current_u, (self.dim_u, -1))

    def get_discrete(self):
        return self.A[:self.dim_y, :self.dim_x], self.A[:self.dim_y, self.dim_x:]


class DiscrepDMDc(DMDc):
    """Placeholder for model.py:109-213 (discrepancy DMDc); fitting is outside the accelerated path."""

    def fit_iteration(self, next_y, next_x, next_u):
        raise NotImplementedError('streaming model updates are not part of the B200 hot path yet')


class OnlineDMDc(DMDc):
    """Placeholder for model.py:216-313 (recursive least squares DMDc); fitting is outside the accelerated path."""

    def fit_iteration(self, next_y, next_x, next_u):
        raise NotImplementedError('streaming model updates are not part of the B200 hot path yet')
