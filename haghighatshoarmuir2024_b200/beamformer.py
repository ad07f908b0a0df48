"""Beamformer: drop-in for the non-spiking Hilbert beamformer of micloc/beamformer.py.

`apply_to_signal` (micloc/beamformer.py:260-292) runs STHT + band-pass + complex
projection on the GPU (micloc_hilbert_beamform) and returns complex128 `T x G`.
"""
from __future__ import annotations

import ctypes as C
from numbers import Number
from typing import List, Tuple, Union

import numpy as np
import torch
from scipy.linalg import eigh
from scipy.signal import butter, hilbert

from . import _native as N
from .array_geometry import ArrayGeometry
from .engine import ChainSpec, SnnEngine, _ptr, _stream_ptr

Fs = 48_000


class Beamformer:
    def __init__(self, geometry: ArrayGeometry, kernel_duration: float, freq_range: List, fs: float = Fs,
                 device: int = 0):
        self.geometry = geometry
        self.kernel_duration = kernel_duration
        self.fs = fs
        ker_len = int(fs * kernel_duration)
        impulse = np.zeros(ker_len)
        impulse[0] = 1
        self.kernel = np.fft.fftshift(np.imag(hilbert(impulse)))
        self.freq_range = np.asarray(freq_range)
        try:
            f_low, f_high = freq_range
            if f_low > f_high:
                raise Exception()
        except Exception:
            raise ValueError("freq_range should be a vector consisting of two frequencies f_low < f_high!")
        self.bandpass_filter = butter(2, freq_range, btype="bandpass", analog=False, output="ba", fs=fs)
        self.bandpass_sos = butter(2, freq_range, btype="bandpass", analog=False, output="sos", fs=fs)
        self.device = device
        self.verbose = True
        self._eng = None

    def _engine(self) -> SnnEngine:
        if self._eng is None:
            M = len(self.geometry)
            # the spike/neuron constants are unused by the Hilbert path; any valid values do
            spec = ChainSpec(num_mic=M, stht_kernel=self.kernel, sos=self.bandpass_sos, robust_width=1,
                             bipolar=False, neuron_decay=0.5, neuron_scale=1.0, neuron_len=1)
            self._eng = SnnEngine(spec, np.zeros((2 * M, 1)), device=self.device)
        return self._eng

    def hilbert_signal(self, sig_in: np.ndarray) -> np.ndarray:
        """Band-passed analytic signal `T x M` complex128 (beamformer.py:281-287)."""
        eng = self._engine()
        x = torch.from_numpy(np.ascontiguousarray(sig_in, dtype=np.float32)).to(eng.device)
        z = eng.run_taps(x, want=("z",))["z"][0].cpu().numpy().astype(np.float64)
        M = len(self.geometry)
        return z[:, :M] + 1j * z[:, M:]

    def design_from_template(self, template: Tuple[np.ndarray, np.ndarray], doa_list: np.ndarray,
                             interference_removal: bool = False):
        """(bf_mat `M x G` complex, list of covariance matrices) (beamformer.py:73-192)."""
        try:
            time_temp, sig_temp = template
        except Exception:
            raise ValueError("input template should be a tuple containing (time_in, sig_in) of the template signal!")
        time_interp = np.arange(np.min(time_temp), np.max(time_temp), step=1 / self.fs)
        sig_temp = np.interp(time_interp, time_temp, sig_temp)
        time_temp = time_interp
        if self.verbose:
            print()
            print("+" * 150)
            print(" designing beamforming matrices for various DoAs ".center(150, "+"))
            print("+" * 150)
        cov_mat_list = []
        eng = self._engine()
        time_temp = np.asarray(time_temp, dtype=np.float64)
        for doa in doa_list:
            delays = self.geometry.delays(theta=doa, normalized=True)
            td = time_temp.reshape(-1, 1) - delays.reshape(1, -1)
            td[td < time_temp.min()] = time_temp.min()
            sig_vec = np.interp(td.ravel(), time_temp, sig_temp).reshape(td.shape)
            # STHT only: upstream band-passes a copy it never uses (beamformer.py:136-137)
            x = torch.from_numpy(np.ascontiguousarray(sig_vec, dtype=np.float32)).to(eng.device)
            q = eng.run_taps(x, want=("q",))["q"][0].cpu().numpy().astype(np.float64)
            sig_h = np.roll(sig_vec, len(self.kernel) // 2, axis=0) + 1j * q
            stable_part = min([len(self.kernel), sig_h.shape[0] // 2])
            stable = sig_h[stable_part:, :]
            cov_mat_list.append(1 / stable.shape[0] * (stable.conj().T @ stable))
        bf_mat = []
        if not interference_removal:
            for Cm in cov_mat_list:
                U, _, _ = np.linalg.svd(Cm)
                bf_mat.append(U[:, 0])
        else:
            C_sum = 0
            for Cm in cov_mat_list:
                C_sum = C_sum + Cm
            C_sum = C_sum + np.diag(np.mean(np.diag(C_sum)) * np.ones(C_sum.shape[0])) / 10
            for Cm in cov_mat_list:
                _, vecs = eigh(Cm, C_sum - Cm)   # ascending eigenvalues: last is the largest
                v = vecs[:, -1]
                bf_mat.append(v / np.linalg.norm(v))
        return np.asarray(bf_mat).T, cov_mat_list

    def apply_to_template(self, bf_mat: np.ndarray, template, snr_db: float) -> np.ndarray:
        """beamformer.py:194-258: delayed + noisy array signal, then apply_to_signal."""
        try:
            time_temp, sig_temp, doa_temp = template
        except Exception:
            raise ValueError(
                "input template should be a tuple containing (time_in, sig_in, doa_in) of the template signal!")
        if isinstance(doa_temp, Number):
            doa_temp = doa_temp * np.ones_like(sig_temp)
        snr = 10 ** (snr_db / 10)
        time_in = np.arange(np.min(time_temp), np.max(time_temp), step=1 / self.fs)
        sig_in = np.interp(time_in, time_temp, sig_temp)
        doa_in = np.interp(time_in, time_temp, doa_temp)
        delays = self.geometry.delays_batch(doa_in)
        delays = delays - delays.min()
        td = time_in.reshape(-1, 1) - delays
        td[td < time_in.min()] = time_in.min()
        sig_vec = np.interp(td.ravel(), time_in, sig_in).reshape(td.shape)
        sig_vec = sig_vec + np.sqrt(np.mean(sig_vec ** 2)) / np.sqrt(snr) * np.random.randn(*sig_vec.shape)
        return self.apply_to_signal(bf_mat=bf_mat, sig_in=sig_vec)

    def apply_to_signal(self, bf_mat: np.ndarray, sig_in: np.ndarray) -> np.ndarray:
        bf_mat = np.asarray(bf_mat)
        num_mic, num_grid = bf_mat.shape
        T, num_chan = np.shape(sig_in)
        if num_chan != num_mic:
            raise ValueError(
                f"number of channels in the input siganl {num_chan} should be the same as the number of microphones {num_mic}!")
        eng = self._engine()
        x = torch.from_numpy(np.ascontiguousarray(sig_in, dtype=np.float32)).to(eng.device).unsqueeze(0)
        y = torch.empty((1, T, num_grid, 2), dtype=torch.float32, device=eng.device)
        bre = np.ascontiguousarray(bf_mat.real, dtype=np.float64)
        bim = np.ascontiguousarray(bf_mat.imag, dtype=np.float64)
        N.check(N.lib().micloc_hilbert_beamform(eng._h, _ptr(x), N.F32, 1, T, bre.ctypes.data_as(N._dp),
                                                bim.ctypes.data_as(N._dp), num_grid, _ptr(y), None, None,
                                                _stream_ptr(eng.device)))
        yh = y[0].cpu().numpy().astype(np.float64)
        return yh[..., 0] + 1j * yh[..., 1]
