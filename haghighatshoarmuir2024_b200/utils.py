"""Post-processing helpers mirrored from micloc/utils.py (host side)."""
from __future__ import annotations

import warnings

import numpy as np


def find_peak_location(sig_in: np.ndarray, win_size: int, periodic: bool = True) -> int:
    """Box-car smoothed argmax (micloc/utils.py:84-121): argmax of the 'full'
    convolution with ones(win_size), shifted back by win_size//2, modulo len."""
    sig_in = np.asarray(sig_in)
    if sig_in.ndim != 1:
        raise ValueError("input signal should be 1-dim!")
    if win_size % 2 != 1:
        raise ValueError("averaging window size should be odd to not create confusion in peak index!")
    if win_size > len(sig_in) // 2:
        raise ValueError("size of averaging window is larger than half the length of input signal!")
    full = np.convolve(np.ones(win_size), sig_in, mode="full")
    index = int(np.argmax(full)) - win_size // 2
    if periodic:
        index = index % len(sig_in)
    return index


class Envelope:
    """Asymmetric rise/fall one-pole envelope tracker (micloc/utils.py:15-81)."""

    def __init__(self, rise_time: float, fall_time: float, fs: float):
        if rise_time > fall_time:
            raise ValueError("for proper functioning, an envelope estimator should have a larger fall time!")
        self.rise_time, self.fall_time, self.fs = rise_time, fall_time, fs
        self.win_lens = np.asarray([int(fs * fall_time), int(fs * rise_time)])

    def evolve(self, sig_in: np.ndarray) -> np.ndarray:
        """`T x num_chan` in, envelopes out, on the GPU (micloc_envelope); float32 arithmetic."""
        import torch
        from .engine import envelope
        sig_in = np.asarray(sig_in)
        T, channel = sig_in.shape
        if T < channel:
            warnings.warn("number of channels in the input signal is larger than number of samples in each channel!")
        if not torch.cuda.is_available():
            raise RuntimeError("Envelope.evolve needs a CUDA device (evolve_host is the numpy restatement used by the CPU tests)")
        x = torch.from_numpy(np.ascontiguousarray(sig_in, dtype=np.float32)).cuda()
        return envelope(x, self.fs, self.rise_time, self.fall_time).cpu().numpy().astype(np.float64)

    def evolve_host(self, sig_in: np.ndarray) -> np.ndarray:
        """numpy restatement of micloc/utils.py:49-81 (test infrastructure; `evolve` is the product path)."""
        T, channel = sig_in.shape
        if T < channel:
            warnings.warn("number of channels in the input signal is larger than number of samples in each channel!")
        mag = np.abs(sig_in)
        state = mag[0].copy()
        out = np.empty_like(mag)
        for n in range(1, T):
            out[n - 1] = state
            rising = (mag[n] >= state).astype(int)
            win = self.win_lens[rising]
            state = (1 - 1 / win) * state + 1 / win * mag[n] * rising
        out[T - 1] = state
        return out
