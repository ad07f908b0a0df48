"""Spike encoders: drop-in for micloc/spike_encoder.py (ZeroCrossingSpikeEncoder).

`evolve` keeps the reference signature and return type (float64 `T x C` array in
{-1, 0, +1}) and runs on the GPU through micloc_rzcc_encode_f64, which reproduces
scipy.signal.find_peaks(np.cumsum(x), distance=robust_width) exactly, including
unbounded candidate clusters (white-noise inputs).
"""
from __future__ import annotations

import numpy as np
import torch

from .engine import rzcc_encode


class SpikeEncoder:
    def evolve(self, sig_in: np.ndarray) -> np.ndarray:
        raise NotImplementedError("this methods needs to be implemented in various spike encoders!")

    def __call__(self, *args, **kwargs) -> np.ndarray:
        return self.evolve(*args, **kwargs)


class ZeroCrossingSpikeEncoder(SpikeEncoder):
    """Robust zero-crossing conjugate (RZCC) encoder (micloc/spike_encoder.py:100-137)."""

    def __init__(self, fs: float, robust_width: int = 1, bipolar: bool = False, device: int = 0):
        self.fs = fs
        self.robust_width = robust_width
        self.bipolar = bipolar
        self.device = device

    def evolve(self, sig_in: np.ndarray) -> np.ndarray:
        sig = np.asarray(sig_in, dtype=np.float64)
        if sig.ndim != 2:
            raise ValueError("sig_in should be a `T x num_chan` array")
        if self.robust_width < 1:
            raise ValueError("`distance` must be greater or equal to 1")  # scipy.signal.find_peaks
        if not torch.cuda.is_available():
            raise RuntimeError("ZeroCrossingSpikeEncoder.evolve needs a CUDA device; there is no CPU fallback")
        width = int(np.ceil(self.robust_width))
        dev = torch.device("cuda", self.device)
        x = torch.from_numpy(np.ascontiguousarray(sig)).to(dev).unsqueeze(0)
        spikes = rzcc_encode(x, width, self.bipolar)
        return spikes[0].cpu().numpy().astype(np.float64)

    def evolve_batch(self, sig: torch.Tensor) -> torch.Tensor:
        """Device batch [B,T,C] float64 -> int8 spikes (no host round trip)."""
        return rzcc_encode(sig, int(np.ceil(self.robust_width)), self.bipolar)
