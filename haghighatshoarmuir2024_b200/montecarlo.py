"""Batched Monte-Carlo localisation sweep: the driver loop of
paper_plots/target_snn_localization.py:431-467 (11 SNRs x num_sim random-DoA trials,
one clip per trial) re-stated as batches of independent clips per frequency band.

Clips are synthesised ON THE DEVICE (torch elementwise ops; plumbing, not the hot
path) exactly as SNNBeamformer.apply_to_template builds them
(micloc/snn_beamformer.py:243-275: per-microphone delay, linear interpolation of the
sampled source clamped at t_min, AWGN at the requested SNR), then go through the
fused CUDA chain of one `SnnEngine` per band; only DoA indices, per-DoA power and a
DoA histogram come back.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch
from scipy.signal import butter, hilbert

from . import _native as N
from .engine import ChainSpec, SnnEngine, neuron_alpha_params

SPEED_OF_SOUND = 340.0            # micloc/array_geometry.py:14


@dataclass
class BandSetup:
    band: Sequence[float]
    tau: float
    bf_mat: np.ndarray            # [2M, G]


def stht_kernel(fs: float, kernel_duration: float) -> np.ndarray:
    """micloc/snn_beamformer.py:45-53."""
    K = int(fs * kernel_duration)
    impulse = np.zeros(K)
    impulse[0] = 1
    return np.fft.fftshift(np.imag(hilbert(impulse)))


def band_chain_spec(num_mic: int, fs: float, kernel: np.ndarray, band, tau: float, T: int,
                    bipolar: bool = True) -> ChainSpec:
    """Constants SNNBeamformer derives for one band (micloc/snn_beamformer.py:58-80, 342-361)."""
    sos = butter(2, band, btype="bandpass", analog=False, output="sos", fs=fs)
    robust_width = max(int(fs / band[1]) // 2, 1)
    _, a, c, L = neuron_alpha_params(np.arange(T) / fs, [tau, tau])
    return ChainSpec(num_mic=num_mic, stht_kernel=kernel, sos=sos, robust_width=robust_width, bipolar=bipolar,
                     neuron_decay=a, neuron_scale=c, neuron_len=L)


class SnrSweep:
    """One engine per band on one GPU + on-device clip synthesis."""

    def __init__(self, bands: List[BandSetup], r_vec, theta_vec, fs: float, kernel_duration: float, T: int,
                 device: int = 0, bipolar: bool = True):
        self.fs, self.T = float(fs), int(T)
        self.device = torch.device("cuda", device)
        self.r_vec = np.asarray(r_vec, dtype=np.float64)
        self.theta_vec = np.asarray(theta_vec, dtype=np.float64)
        self.M = len(self.r_vec)
        self.kernel = stht_kernel(fs, kernel_duration)
        self.bands = bands
        self.specs = [band_chain_spec(self.M, fs, self.kernel, b.band, b.tau, T, bipolar) for b in bands]
        self.engines = [SnnEngine(s, b.bf_mat, device=device) for s, b in zip(self.specs, bands)]
        self.G = bands[0].bf_mat.shape[1]

    # ------------------------------------------------------------------
    def synthesize(self, band_idx: int, B: int, seed: int, snr_db_grid: Sequence[float],
                   dtype: torch.dtype = torch.float32, chunk: int = 256):
        """B noisy single-target clips for band `band_idx`: sine at the band's upper edge
        (freq_design, target_snn_localization.py:439-441), DoA ~ U[0, 2 pi) (:452), SNR cycling
        through `snr_db_grid` minus the bandwidth correction (:382, :449).
        Returns (audio [B,T,M] device tensor, doa_true [B] float64 numpy, snr_index [B])."""
        band = self.bands[band_idx].band
        dev = self.device
        gen = torch.Generator(device=dev)
        gen.manual_seed(int(seed))
        rng = np.random.default_rng(int(seed))
        doa = rng.uniform(0.0, 2 * np.pi, size=B)
        snr_idx = np.arange(B) % len(snr_db_grid)
        corr = 10 * np.log10((self.fs / 2) / (band[1] - band[0]))
        snr = 10 ** ((np.asarray(snr_db_grid, dtype=np.float64)[snr_idx] - corr) / 10)
        f0 = float(band[1])
        T, M = self.T, self.M
        out = torch.empty((B, T, M), dtype=dtype, device=dev)
        n = torch.arange(T, device=dev, dtype=torch.float64).view(1, T, 1)
        for b0 in range(0, B, chunk):
            b1 = min(B, b0 + chunk)
            d = -self.r_vec[None, :] * np.cos(self.theta_vec[None, :] - doa[b0:b1, None]) / SPEED_OF_SOUND
            d = d - d.min(axis=1, keepdims=True)                       # snn_beamformer.py:256-257
            dn = torch.from_numpy(d * self.fs).to(dev).view(b1 - b0, 1, M)
            pos = (n - dn).clamp_(min=0.0)                             # clamp at t_min (:262-264)
            i0 = pos.floor()
            fr = pos - i0
            w = 2 * np.pi * f0 / self.fs
            x = (1 - fr) * torch.sin(w * i0) + fr * torch.sin(w * (i0 + 1))     # np.interp of the sampled sine
            rms = x.pow(2).mean(dim=(1, 2), keepdim=True).sqrt()
            sigma = rms / torch.from_numpy(np.sqrt(snr[b0:b1])).to(dev).view(-1, 1, 1)
            noise = torch.randn((b1 - b0, T, M), generator=gen, device=dev, dtype=torch.float32)
            x = x.to(torch.float32) + sigma.to(torch.float32) * noise
            if dtype == torch.int16:
                scale = 12000.0 / x.abs().amax(dim=(1, 2), keepdim=True)
                out[b0:b1] = (x * scale).round().to(torch.int16)
            else:
                out[b0:b1] = x
        return out, doa, snr_idx

    # ------------------------------------------------------------------
    def run_band(self, band_idx: int, audio: torch.Tensor, want_spikes: bool = False, want_power: bool = True,
                 hist: Optional[torch.Tensor] = None) -> Dict[str, torch.Tensor]:
        eng = self.engines[band_idx]
        out = eng.run(audio, want_spikes=want_spikes, want_power=want_power, fused=True)
        if hist is not None:
            self.histogram(out["doa"], hist)
        return out

    def histogram(self, doa: torch.Tensor, hist: torch.Tensor) -> None:
        """hist[g] += count(doa == g) with the library's own kernel (int64 [G] on this GPU)."""
        assert hist.dtype == torch.int64 and hist.numel() == self.G and hist.device == doa.device
        st = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        N.check(N.lib().micloc_doa_histogram(C.c_void_p(doa.data_ptr()), doa.numel(), self.G,
                                             C.c_void_p(hist.data_ptr()), self.device.index or 0, st))

    def doa_error(self, doa_idx: np.ndarray, doa_true: np.ndarray, doa_list: np.ndarray) -> np.ndarray:
        """arcsin|sin(est - true)| (target_snn_localization.py:466)."""
        return np.arcsin(np.abs(np.sin(doa_list[doa_idx] - doa_true)))
