"""GaP -- Gamma-Poisson factor model without the zero-inflation layer (oriana/models/gap.py:14-135).

X_ij ~ Poisson(sum_k U_ik V_jk),  U_ik ~ Gamma(alpha1_k, alpha2_k),  V_jk ~ Gamma(beta1_k, beta2_k).
The Gamma rates are column sums (gap.py:98,106); the latent-count step is the dropout-free configuration of
the same CUDA kernels.
"""
import numpy as np
import torch

from ..nodes import Gamma, Poisson
from .base import FactorModel


class GaP(FactorModel):

    _dropout = False

    def __init__(self, *args, **kwargs):
        FactorModel.__init__(self, *args, **kwargs)

    def build_u_node(self):
        self._hyper[0:2] = 1.                                           # gap.py:19-20
        return Gamma(self.alpha1, self.alpha2, self.dims('n,k ~ s,d'), name='U')

    def build_v_node(self):
        self._hyper[2:4] = 1.                                           # gap.py:24-25
        return Gamma(self.beta1, self.beta2, self.dims('m,k ~ s,d'), name='V')

    def build_x_node(self, cmatrix, UV):
        X = Poisson(UV, self.dims('n,m ~ d,d'), name='X')
        X.buffer = self._X                                              # gap.py:30 (device view, no copy)
        src = cmatrix.as_tensor()                                       # the dtype the reference's buffer would have
        X.integer_samples = not src.dtype.is_floating_point if isinstance(src, torch.Tensor) \
            else bool(np.issubdtype(src.dtype, np.integer))
        return X

    def define_variational_distribution(self):
        # the reference draws Gamma(2) placeholders here that it overwrites at once (gap.py:37-45);
        # the draws are repeated only to consume the global RNG in the same order
        np.random.gamma(2., size=(self.n, self.k))
        np.random.gamma(2., size=(self.m, self.k))
        self.U_q = Gamma(self.a1, self.a2, self.dims('n,k ~ d,d'))
        self.V_q = Gamma(self.b1, self.b2, self.dims('m,k ~ d,d'))

    def initialize_variational_parameters(self):
        self._draw_factor_inits()                                       # gap.py:47-65
