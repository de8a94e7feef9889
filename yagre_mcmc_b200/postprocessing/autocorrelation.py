"""Integrated autocorrelation time / ESS (reference: yagremcmc/postprocessing/autocorrelation.py:5-140).

integrated_autocorrelation(seq, method) keeps the reference signature for a single chain
(seq = [N, d] array-like or list of states) and accepts an ensemble trajectory
([N, nChains, d] array-like or a chain.Trajectory), returning one IAT per chain.  The work is
done by iat_kernel (diag_kernels.cu): ACF by direct summation in shared memory up to Sokal's
window.  There is no host implementation."""
import numpy as np
import torch

from ..ensemble import iat_ess
from ..chain.chain import Trajectory


def _as_device_samples(seq):
    """-> device tensor [N, d, n], squeeze flag."""
    if isinstance(seq, Trajectory):
        return seq.device_tensor, seq.shape.__len__() == 2
    if torch.is_tensor(seq):
        t = seq
    else:
        t = torch.as_tensor(np.asarray(seq, dtype=np.float64))
    if t.dim() == 1:
        t = t.reshape(-1, 1)
    squeeze = t.dim() == 2
    if squeeze:
        t = t.unsqueeze(1)                                   # [N, 1, d]
    if not torch.cuda.is_available():
        raise RuntimeError("integrated_autocorrelation runs on the GPU (no CPU fallback)")
    return t.to('cuda', dtype=torch.float64).permute(0, 2, 1).contiguous(), squeeze


def integrated_autocorrelation(seq, method='mean', sokalConst=5.):
    if method not in ['mean', 'max']:
        raise ValueError(f"Invalid IAT - Type: {method}. Options are 'mean' and 'max'.")
    samples, squeeze = _as_device_samples(seq)
    iat, _ = iat_ess(samples, method, sokalConst)
    out = iat.cpu().numpy()
    return int(out[0]) if squeeze else out


def effective_sample_size(seq, burnIn=0, method='max', sokalConst=5.):
    """(N - burnIn) // IAT per chain, the idiom of example_inference_lotkaVolterra_twoLevel.py:117-118,132."""
    samples, squeeze = _as_device_samples(seq)
    iat, ess = iat_ess(samples[burnIn:], method, sokalConst)
    out = ess.cpu().numpy()
    return int(out[0]) if squeeze else out
