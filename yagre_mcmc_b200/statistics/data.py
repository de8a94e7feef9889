"""Observations [nData, dataDim] (reference: yagremcmc/statistics/data.py:4-23)."""
import numpy as np


class Data:

    def __init__(self, dataArray):
        self.array_ = np.array(dataArray, dtype=np.float64)
        if self.array_.ndim != 2:
            raise ValueError("data must be [nData, dataDim]")
        self.size_, self.dim_ = self.array_.shape

    @property
    def size(self):
        return self.size_

    @property
    def dim(self):
        return self.dim_

    @property
    def array(self):
        return self.array_
