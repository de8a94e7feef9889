"""Covariance operators (reference: yagremcmc/statistics/covariance.py:8-94).

On the device a covariance appears in two lowered forms:
  chol_factor()  lower-triangular L with  L L' = C   (proposal: p = s + L z)
  precision()    dense C^-1                          (noise / prior norms: x' C^-1 x)
Diagonal kinds keep the reference's storage quirk -- it stores reciprocal(var) and applies
sqrt(reciprocal(precision)) (:37-38,51-52) -- so lowered values are bit-identical.
"""
from abc import abstractmethod

import numpy as np

from .interface import CovarianceOperatorInterface


class CovarianceMatrix(CovarianceOperatorInterface):

    @property
    @abstractmethod
    def dimension(self):
        ...

    @abstractmethod
    def apply_chol_factor(self, x):
        ...

    def induced_norm_squared(self, x):
        return np.dot(x, self.apply_inverse(x))

    # lowered forms
    @abstractmethod
    def chol_factor(self):
        ...

    @abstractmethod
    def precision(self):
        ...


class DiagonalCovarianceMatrix(CovarianceMatrix):

    def __init__(self, marginalVariances):
        self._precision = np.reciprocal(np.asarray(marginalVariances, dtype=np.float64))

    @property
    def marginalVariance(self):
        return np.reciprocal(self._precision)

    @marginalVariance.setter
    def marginalVariance(self, mVar):
        self._precision = np.reciprocal(np.asarray(mVar, dtype=np.float64))

    @property
    def dimension(self):
        return self._precision.size

    def apply_chol_factor(self, x):
        return np.sqrt(np.reciprocal(self._precision)) * x

    def apply_inverse(self, x):
        return self._precision * x

    def chol_factor(self):
        return np.diag(np.sqrt(np.reciprocal(self._precision)))

    def precision(self):
        return np.diag(self._precision)          # exact zeros off the diagonal (kernels skip them)


class IIDCovarianceMatrix(DiagonalCovarianceMatrix):

    def __init__(self, dimension, variance):
        super().__init__(np.full(dimension, float(variance)))


class DenseCovarianceMatrix(CovarianceMatrix):

    def __init__(self, denseCovMat):
        C = np.asarray(denseCovMat, dtype=np.float64)
        if C.ndim != 2 or C.shape[0] != C.shape[1]:
            raise ValueError("dense covariance must be square")
        self.dim_ = C.shape[0]
        self.cholFactor_ = np.linalg.cholesky(C)          # LAPACK potrf, lower (reference :78)

    @property
    def dimension(self):
        return self.dim_

    def apply_chol_factor(self, x):
        return self.cholFactor_ @ x

    def apply_inverse(self, x):
        y = np.linalg.solve(self.cholFactor_, x)
        return np.linalg.solve(self.cholFactor_.T, y)

    def dense(self):
        return self.cholFactor_ @ self.cholFactor_.T

    def chol_factor(self):
        return self.cholFactor_

    def precision(self):
        Linv = np.linalg.inv(self.cholFactor_)
        P = Linv.T @ Linv
        return 0.5 * (P + P.T)
