"""Multi-band SNN localiser: the arithmetic of micloc/localization_demo_snn.py's `Demo` on the GPU.

Same constructor as the reference class (micloc/localization_demo_snn.py:22-98).  The reference's `run()` is an endless
record -> process -> plot loop around a `sox` recorder and a matplotlib visualiser (both out of scope); `process_frame`
is the body of that loop (localization_demo_snn.py:134-193): int32 `T x 8` wav frame in, DoA in degrees (or NaN when
the activity detector finds no signal) out.  `localize` is the batched form: frames [B, T, channels] in one go.

Per frame: filterbank on the raw integers (micloc_filterbank), one fused SNN chain per band (micloc_snn_run on that
band's rows), power summed over the bands + argmax + the periodic-ML estimators (micloc_power_fuse).
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional, Sequence

import numpy as np
import torch

from . import _native as N
from .array_geometry import ArrayGeometry
from .engine import _ptr, _stream_ptr
from .filterbank import ButterworthFilterbank
from .snn_beamformer import SNNBeamformer


class Demo:
    def __init__(self, geometry: ArrayGeometry, freq_bands: np.ndarray, doa_list: np.ndarray, recording_duration: float,
                 kernel_duration: float, bipolar_spikes: bool, fs: float, device: int = 0,
                 bf_mats: Optional[Sequence[np.ndarray]] = None):
        """`bf_mats` (optional) skips the design step with matrices designed earlier (e.g. by the reference)."""
        freq_bands = np.asarray(freq_bands, dtype=np.float64)
        if freq_bands.ndim == 1:
            freq_bands = freq_bands.reshape(1, -1)
        self.beamfs, self.bf_mats = [], []
        time_temp = np.arange(0, recording_duration, step=1 / fs)
        for i, freq_range in enumerate(freq_bands):
            freq_mid = np.mean(freq_range)
            tau_mem = 1 / (2 * np.pi * freq_mid)                   # localization_demo_snn.py:62-65
            beamf = SNNBeamformer(geometry=geometry, kernel_duration=kernel_duration, freq_range=freq_range,
                                  tau_vec=[tau_mem, tau_mem], bipolar_spikes=bipolar_spikes, fs=fs, device=device)
            beamf.verbose = False
            self.beamfs.append(beamf)
            if bf_mats is not None:
                self.bf_mats.append(np.asarray(bf_mats[i], dtype=np.float64))
            else:
                sig_temp = np.sin(2 * np.pi * freq_mid * time_temp)
                self.bf_mats.append(beamf.design_from_template(template=(time_temp, sig_temp), doa_list=doa_list))
        self.filterbank = ButterworthFilterbank(freq_bands=freq_bands, order=1, fs=fs)
        self.doa_list = np.asarray(doa_list, dtype=np.float64)
        self.recording_duration = recording_duration
        self.kernel_duration = kernel_duration
        self.fs = fs
        self.device = torch.device("cuda", device)
        self.num_mic = len(geometry)
        self._doa_dev = None

    # ------------------------------------------------------------------
    def localize(self, frames: torch.Tensor, rel_threshold: float = 0.0001, full_scale: Optional[float] = None,
                 want_spikes: bool = False) -> Dict[str, torch.Tensor]:
        """frames [B, T, channels >= num_mic] (CUDA, int32 / int16 / float32; extra channels are dropped like the
        reference's `data[:, :-1]`) -> dict(power_grid [B, G], doa [B] index of the peak, doa_deg [B] (NaN = no
        activity), periodic_ml / trimmed_periodic_ml [B] radians, active [B] bool, power [F, B, G], flags [B])."""
        if frames.dim() == 2:
            frames = frames.unsqueeze(0)
        if frames.dim() != 3 or frames.shape[2] < self.num_mic:
            raise ValueError(
                f"number of channels in the input siganl {frames.shape[-1]} should be the same as the number of microphones {self.num_mic}!")
        if frames.device != self.device:
            raise ValueError(f"frames live on {frames.device}, the localiser on {self.device}")
        dt = {torch.float32: N.F32, torch.int16: N.I16, torch.int32: N.I32}.get(frames.dtype)
        if dt is None:
            raise ValueError(f"frames must be float32, int16 or int32, got {frames.dtype}")
        frames = frames.contiguous()
        B, T, ch = frames.shape
        F, M, G = len(self.beamfs), self.num_mic, len(self.doa_list)
        lib = N.lib()
        dev = self.device
        st = _stream_ptr(dev)
        sos = np.ascontiguousarray(self.filterbank.sos_list, dtype=np.float64)          # [F][nsec][6]
        banded = torch.empty((F, B, T, M), dtype=torch.float32, device=dev)
        sumsq = torch.empty(B, dtype=torch.float64, device=dev)
        N.check(lib.micloc_filterbank(_ptr(frames), dt, B, T, ch, M, F, sos.shape[1], sos.ctypes.data_as(N._dp),
                                      _ptr(banded), _ptr(sumsq), dev.index or 0, st))
        time_vec = np.arange(0, T) / self.fs
        power = torch.empty((F, B, G), dtype=torch.float32, device=dev)
        flags = torch.zeros(B, dtype=torch.int32, device=dev)
        spikes = []
        for f, (beamf, bf) in enumerate(zip(self.beamfs, self.bf_mats)):
            eng = beamf.engine(bf, time_vec)
            out = eng.run(banded[f], want_spikes=want_spikes, want_power=True, fused=True)
            spikes.append(out["spikes"])
            power[f].copy_(out["power"])
            flags |= out["flags"]
        if self._doa_dev is None:
            self._doa_dev = torch.from_numpy(self.doa_list).to(dev)
        power_grid = torch.empty((B, G), dtype=torch.float32, device=dev)
        doa = torch.empty(B, dtype=torch.int32, device=dev)
        ml = torch.empty(B, dtype=torch.float64, device=dev)
        trimmed = torch.empty(B, dtype=torch.float64, device=dev)
        fflags = torch.empty(B, dtype=torch.int32, device=dev)
        N.check(lib.micloc_power_fuse(_ptr(power), F, B, G, _ptr(self._doa_dev), _ptr(power_grid), _ptr(doa), _ptr(ml),
                                      _ptr(trimmed), _ptr(fflags), dev.index or 0, st))
        # activity detection (localization_demo_snn.py:136-163): rms over the frame against rel_threshold * full scale
        if full_scale is None:
            full_scale = float(torch.iinfo(frames.dtype).max) if not frames.dtype.is_floating_point else 1.0
        rms = torch.sqrt(sumsq / (T * M))
        active = rms >= rel_threshold * full_scale
        doa_deg = torch.where(active, self._doa_dev[doa.long()] * 180 / np.pi,
                              torch.full_like(rms, float("nan")))
        return {"power_grid": power_grid, "doa": doa, "doa_deg": doa_deg, "periodic_ml": ml, "trimmed_periodic_ml": trimmed,
                "active": active, "power": power, "flags": flags | fflags, "spikes": spikes if want_spikes else None,
                "banded": banded}

    def process_frame(self, data: np.ndarray) -> float:
        """One recording as the reference's loop handles it: `T x 8` integer wav frame -> DoA in degrees, NaN when no
        activity was detected (what the loop pushes to its visualiser)."""
        data = np.asarray(data)
        out = self.localize(torch.from_numpy(np.ascontiguousarray(data)).to(self.device))
        return float(out["doa_deg"][0])

    def run(self, frames=None):
        """The reference records from the devkit forever; here `frames` is any iterable of wav frames."""
        if frames is None:
            raise NotImplementedError("the recorder / visualiser loop is out of scope: pass an iterable of wav frames")
        for data in frames:
            yield self.process_frame(data)
