"""Trajectory store (reference: yagremcmc/chain/chain.py:4-21).

The reference keeps a Python list of state vectors; the ensemble keeps ONE device tensor
[chainLength, d, nChains] (chain index fastest, the layout the kernels write coalesced) and
exposes it as an array-like of shape [chainLength, nChains, d] -- or [chainLength, d] for a
single chain, which is what reference scripts index (`states[burnIn:]`, `states[-1]`)."""
import numpy as np


class Trajectory:

    def __init__(self, device_tensor, squeeze):
        self._t = device_tensor            # [length, d, n] on the device
        self._squeeze = squeeze
        self._host = None

    @property
    def device_tensor(self):
        return self._t

    def numpy(self):
        if self._host is None:
            a = self._t.permute(0, 2, 1).contiguous().cpu().numpy()     # [length, n, d]
            self._host = a[:, 0, :] if self._squeeze else a
        return self._host

    def __len__(self):
        return int(self._t.shape[0])

    def __getitem__(self, idx):
        if self._host is None and isinstance(idx, (int, np.integer)):
            row = self._t[idx].t().contiguous().cpu().numpy()           # [n, d]
            return row[0] if self._squeeze else row
        return self.numpy()[idx]

    def __iter__(self):
        return iter(self.numpy())

    def __array__(self, dtype=None, copy=None):
        a = self.numpy()
        return a.astype(dtype) if dtype is not None else a

    @property
    def shape(self):
        L, d, n = self._t.shape
        return (L, d) if self._squeeze else (L, n, d)


class Chain:

    def __init__(self):
        self._trajectory = []

    @property
    def trajectory(self):
        return self._trajectory

    @property
    def length(self):
        return len(self._trajectory)

    def set(self, trajectory):
        self._trajectory = trajectory

    def clear(self):
        self._trajectory = []
