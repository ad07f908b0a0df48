"""ctypes front end of the CPU ORACLE (oracle/micloc_oracle.c).

TEST INFRASTRUCTURE ONLY.  Importable from tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs -- never from the product
package (haghighatshoarmuir2024_b200/).

Each wrapper cites the reference site it restates; see the C file header.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass
from typing import Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libmicloc_oracle.so")
_lib = None

_dp = C.POINTER(C.c_double)


def build(force: bool = False) -> str:
    """Compile the oracle with oracle/Makefile (gcc)."""
    src = os.path.join(_HERE, "micloc_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-B" if force else "-s"], check=True,
                       stdout=subprocess.DEVNULL)
    return _LIB_PATH


class _SnnCfg(C.Structure):
    _fields_ = [
        ("M", C.c_int), ("K", C.c_int), ("h", _dp),
        ("nba", C.c_int), ("b", _dp), ("a", _dp),
        ("robust_width", C.c_double), ("bipolar", C.c_int),
        ("L", C.c_int), ("nir", _dp),
        ("G", C.c_int), ("bf", _dp),
    ]


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.mo_pairwise_sum.restype = C.c_double
        _lib.mo_pairwise_sum.argtypes = [_dp, C.c_int]
        _lib.mo_neuron_kernel.restype = C.c_int
        _lib.mo_neuron_kernel.argtypes = [_dp, C.c_int, C.c_double, C.c_double, _dp]
        _lib.mo_rzcc.restype = C.c_int
        _lib.mo_rzcc.argtypes = [_dp, C.c_int, C.c_int, C.c_double, C.c_int, _dp]
        _lib.mo_snn_apply.restype = C.c_int
        _lib.mo_snn_run_batch.restype = C.c_int
        _lib.mo_xylo_encode.restype = C.c_int
        _lib.mo_xylo_lif.restype = None
        _lib.mo_xylo_rate_doa.restype = None
        _lib.mo_xylo_run_batch.restype = C.c_int
    return _lib


def _p(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(_dp)


def _f64(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float64)


def fir(x: np.ndarray, b: np.ndarray) -> np.ndarray:
    """scipy.signal.lfilter(b, [1], x, axis=0) (micloc/snn_beamformer.py:327,364)."""
    x = _f64(x); b = _f64(b)
    x2 = x.reshape(x.shape[0], -1)
    y = np.empty_like(x2)
    lib().mo_fir_df2t(_p(x2), _p(y), C.c_int(x2.shape[0]), C.c_int(x2.shape[1]), _p(b), C.c_int(len(b)))
    return y.reshape(x.shape)


def iir(x: np.ndarray, b: np.ndarray, a: np.ndarray) -> np.ndarray:
    """scipy.signal.lfilter(b, a, x, axis=0), a[0]==1 (micloc/snn_beamformer.py:330-331)."""
    x = _f64(x); b = _f64(b); a = _f64(a)
    assert len(a) == len(b) and a[0] == 1.0
    x2 = x.reshape(x.shape[0], -1)
    y = np.empty_like(x2)
    lib().mo_iir_df2t(_p(x2), _p(y), C.c_int(x2.shape[0]), C.c_int(x2.shape[1]), _p(b), _p(a), C.c_int(len(b)))
    return y.reshape(x.shape)


def stht(x: np.ndarray, h: np.ndarray):
    """(in-phase, quadrature) of micloc/snn_beamformer.py:325-327."""
    x = _f64(x); h = _f64(h)
    T, M = x.shape
    i = np.empty_like(x); q = np.empty_like(x)
    lib().mo_stht(_p(x), C.c_int(T), C.c_int(M), _p(h), C.c_int(len(h)), _p(i), _p(q))
    return i, q


def rzcc(x: np.ndarray, robust_width: float, bipolar: bool) -> np.ndarray:
    """ZeroCrossingSpikeEncoder.evolve (micloc/spike_encoder.py:115-137)."""
    x = _f64(x)
    T, Cc = x.shape
    s = np.empty_like(x)
    rc = lib().mo_rzcc(_p(x), T, Cc, float(robust_width), int(bool(bipolar)), _p(s))
    if rc != 0:
        raise ValueError("`distance` must be greater or equal to 1")
    return s


def neuron_kernel(time_vec: np.ndarray, tau_syn: float, tau_mem: float) -> np.ndarray:
    """Truncated, normalised alpha kernel (micloc/snn_beamformer.py:342-361)."""
    tv = _f64(time_vec)
    h = np.empty(len(tv))
    L = lib().mo_neuron_kernel(_p(tv), len(tv), float(tau_syn), float(tau_mem), _p(h))
    if L < 0:
        raise AssertionError("tau_syn != tau_mem is not executable in the reference")
    return h[:L].copy()


@dataclass
class SnnConfig:
    """Host-side constants of one SNNBeamformer + bf_mat (float64)."""
    h: np.ndarray            # STHT kernel [K]
    b: np.ndarray            # band-pass numerator
    a: np.ndarray            # band-pass denominator (a[0] == 1)
    robust_width: float
    bipolar: bool
    nir: np.ndarray          # truncated neuron impulse response [L]
    bf: np.ndarray           # [2M, G]

    def _c(self):
        self.h = _f64(self.h); self.b = _f64(self.b); self.a = _f64(self.a)
        self.nir = _f64(self.nir); self.bf = _f64(self.bf)
        c = _SnnCfg()
        c.M = self.bf.shape[0] // 2; c.K = len(self.h); c.h = _p(self.h)
        c.nba = len(self.b); c.b = _p(self.b); c.a = _p(self.a)
        c.robust_width = float(self.robust_width); c.bipolar = int(bool(self.bipolar))
        c.L = len(self.nir); c.nir = _p(self.nir)
        c.G = self.bf.shape[1]; c.bf = _p(self.bf)
        return c


def snn_apply(cfg: SnnConfig, x: np.ndarray, want=("q", "z", "spikes", "vmem", "y", "power")):
    """Full chain on one clip x[T, M]; returns dict with 'doa' + requested taps.

    Restates SNNBeamformer.apply_to_signal (micloc/snn_beamformer.py:283-370) and
    the callers' mean-power/argmax (paper_plots/target_snn_localization.py:462-464).
    """
    x = _f64(x)
    T, M = x.shape
    c = cfg._c()
    assert M == c.M
    out = {}
    shapes = {"q": (T, M), "z": (T, 2 * M), "spikes": (T, 2 * M), "vmem": (T, 2 * M),
              "y": (T, c.G), "power": (c.G,)}
    bufs = {k: (np.empty(shapes[k]) if k in want else None) for k in shapes}
    doa = lib().mo_snn_apply(C.byref(c), _p(x), C.c_int(T), _p(bufs["q"]), _p(bufs["z"]),
                             _p(bufs["spikes"]), _p(bufs["vmem"]), _p(bufs["y"]), _p(bufs["power"]))
    out["doa"] = int(doa)
    out.update({k: v for k, v in bufs.items() if v is not None})
    return out


def snn_run_batch(cfg: SnnConfig, audio: np.ndarray, nthreads: int = 1,
                  want_power: bool = True, want_spikes: bool = False):
    """Batched chain over audio[B, T, M] (float32 or int16) on `nthreads` host threads."""
    assert audio.dtype in (np.float32, np.int16) and audio.flags.c_contiguous
    B, T, M = audio.shape
    c = cfg._c()
    assert M == c.M
    doa = np.empty(B, dtype=np.int32)
    power = np.empty((B, c.G)) if want_power else None
    spikes = np.empty((B, T, 2 * M), dtype=np.int8) if want_spikes else None
    used = lib().mo_snn_run_batch(
        C.byref(c), audio.ctypes.data_as(C.c_void_p), int(audio.dtype == np.int16),
        C.c_int(B), C.c_int(T), doa.ctypes.data_as(C.c_void_p), _p(power),
        None if spikes is None else spikes.ctypes.data_as(C.c_void_p), C.c_int(nthreads))
    return {"doa": doa, "power": power, "spikes": spikes, "threads": int(used)}


def beamformer_apply(x, h, b, a, bf_mat) -> np.ndarray:
    """Beamformer.apply_to_signal (micloc/beamformer.py:260-292): complex [T, G]."""
    x = _f64(x); h = _f64(h); b = _f64(b); a = _f64(a)
    bf_re = _f64(np.real(bf_mat)); bf_im = _f64(np.imag(bf_mat))
    T, M = x.shape
    G = bf_re.shape[1]
    yr = np.empty((T, G)); yi = np.empty((T, G))
    lib().mo_beamformer_apply(_p(x), C.c_int(T), C.c_int(M), _p(h), C.c_int(len(h)), _p(b), _p(a),
                              C.c_int(len(b)), _p(bf_re), _p(bf_im), C.c_int(G), _p(yr), _p(yi))
    return yr + 1j * yi


# --------------------------------------------------------------------------------------
# Xylo integer chain (micloc/xylo_snn_localization.py:315-398; XyloSim restated, parity unpinned)
# --------------------------------------------------------------------------------------
class _XyloCfg(C.Structure):
    _fields_ = [
        ("M", C.c_int), ("K", C.c_int), ("h", _dp),
        ("F", C.c_int), ("nba", C.c_int), ("b", _dp), ("a", _dp),
        ("robust_width", C.c_double), ("bipolar", C.c_int),
        ("N_in", C.c_int), ("N", C.c_int), ("G", C.c_int),
        ("w_in", C.c_void_p), ("w_rec", C.c_void_p), ("threshold", C.c_void_p),
        ("dash_syn", C.c_void_p), ("dash_mem", C.c_void_p), ("bias", C.c_void_p),
        ("weight_shift_in", C.c_int), ("weight_shift_rec", C.c_int), ("max_spikes", C.c_int),
    ]


@dataclass
class XyloConfig:
    """Host constants of Demo (front end) + the quantised hidden layer."""
    h: np.ndarray                # STHT kernel [K]
    b: np.ndarray                # [F, nba] band-filter numerators
    a: np.ndarray                # [F, nba] denominators (a[:, 0] == 1)
    robust_width: float
    bipolar: bool
    num_mic: int
    num_doa: int
    w_in: np.ndarray             # int8 [N_in, N]
    threshold: np.ndarray        # int16 [N]
    dash_syn: np.ndarray         # int8 [N]
    dash_mem: np.ndarray         # int8 [N]
    w_rec: Optional[np.ndarray] = None
    bias: Optional[np.ndarray] = None
    weight_shift_in: int = 0
    weight_shift_rec: int = 0
    max_spikes: int = 31

    def _c(self):
        self.h = _f64(self.h)
        self.b = _f64(np.atleast_2d(self.b)); self.a = _f64(np.atleast_2d(self.a))
        self.w_in = np.ascontiguousarray(self.w_in, dtype=np.int8)
        self.threshold = np.ascontiguousarray(self.threshold, dtype=np.int16)
        self.dash_syn = np.ascontiguousarray(self.dash_syn, dtype=np.int8)
        self.dash_mem = np.ascontiguousarray(self.dash_mem, dtype=np.int8)
        if self.w_rec is not None:
            self.w_rec = np.ascontiguousarray(self.w_rec, dtype=np.int8)
        if self.bias is not None:
            self.bias = np.ascontiguousarray(self.bias, dtype=np.int16)
        vp = lambda arr: None if arr is None else arr.ctypes.data_as(C.c_void_p)
        c = _XyloCfg()
        c.M = self.num_mic; c.K = len(self.h); c.h = _p(self.h)
        c.F, c.nba = self.b.shape; c.b = _p(self.b); c.a = _p(self.a)
        c.robust_width = float(self.robust_width); c.bipolar = int(bool(self.bipolar))
        c.N_in, c.N = self.w_in.shape; c.G = self.num_doa
        assert c.N_in == 2 * c.M * c.F * (2 if self.bipolar else 1) and c.N == c.G * c.F
        c.w_in = vp(self.w_in); c.w_rec = vp(self.w_rec); c.threshold = vp(self.threshold)
        c.dash_syn = vp(self.dash_syn); c.dash_mem = vp(self.dash_mem); c.bias = vp(self.bias)
        c.weight_shift_in = self.weight_shift_in; c.weight_shift_rec = self.weight_shift_rec
        c.max_spikes = self.max_spikes
        return c


def xylo_encode(cfg: XyloConfig, x: np.ndarray):
    """Demo.spike_encoding (micloc/xylo_snn_localization.py:315-356) on one clip x[T, M]:
    returns (spikes_in int8 [T, N_in] in {0,1}, signed raster int8 [T, 2M*F])."""
    x = _f64(x)
    T, M = x.shape
    c = cfg._c()
    assert M == c.M
    spk = np.zeros((T, c.N_in), dtype=np.int8)
    sgn = np.zeros((T, 2 * M * c.F), dtype=np.int8)
    rc = lib().mo_xylo_encode(C.byref(c), _p(x), C.c_int(T), spk.ctypes.data_as(C.c_void_p), sgn.ctypes.data_as(C.c_void_p))
    if rc != 0:
        raise ValueError("`distance` must be greater or equal to 1")
    return spk, sgn


def xylo_lif(cfg: XyloConfig, spikes_in: np.ndarray):
    """Hidden layer of XyloSim (Demo.xylo_process, :358-377) on one clip: (raster uint8 [T, N], counts int32 [N])."""
    s = np.ascontiguousarray(spikes_in, dtype=np.int8)
    c = cfg._c()
    T = s.shape[0]
    assert s.shape[1] == c.N_in
    raster = np.zeros((T, c.N), dtype=np.uint8)
    counts = np.zeros(c.N, dtype=np.int32)
    lib().mo_xylo_lif(C.byref(c), s.ctypes.data_as(C.c_void_p), C.c_int(T), raster.ctypes.data_as(C.c_void_p),
                      counts.ctypes.data_as(C.c_void_p))
    return raster, counts


def xylo_rate_doa(counts: np.ndarray, G: int, F: int, T: int, fs: float, win: int = 0):
    """extract_rate (:379-398) + argmax + find_peak_location (micloc/utils.py:84-121): (rate [G], doa, doa_peak)."""
    cnt = np.ascontiguousarray(counts, dtype=np.int32)
    rate = np.empty(G)
    d0, d1 = C.c_int(0), C.c_int(-1)
    lib().mo_xylo_rate_doa(cnt.ctypes.data_as(C.c_void_p), C.c_int(G), C.c_int(F), C.c_int(T), C.c_double(fs),
                           C.c_int(win), _p(rate), C.byref(d0), C.byref(d1))
    return rate, d0.value, d1.value


def xylo_run_batch(cfg: XyloConfig, audio: np.ndarray, fs: float, win: int = 0, nthreads: int = 1,
                   want_spikes: bool = False):
    """Batched Xylo chain over audio[B, T, M] (float32 or int16) on `nthreads` host threads."""
    assert audio.dtype in (np.float32, np.int16) and audio.flags.c_contiguous
    B, T, M = audio.shape
    c = cfg._c()
    assert M == c.M
    counts = np.empty((B, c.N), dtype=np.int32)
    doa = np.empty(B, dtype=np.int32)
    doa_peak = np.empty(B, dtype=np.int32)
    sgn = np.empty((B, T, 2 * M * c.F), dtype=np.int8) if want_spikes else None
    used = lib().mo_xylo_run_batch(
        C.byref(c), audio.ctypes.data_as(C.c_void_p), int(audio.dtype == np.int16), C.c_int(B), C.c_int(T),
        C.c_double(fs), C.c_int(win), counts.ctypes.data_as(C.c_void_p), doa.ctypes.data_as(C.c_void_p),
        doa_peak.ctypes.data_as(C.c_void_p), None if sgn is None else sgn.ctypes.data_as(C.c_void_p), C.c_int(nthreads))
    return {"counts": counts, "doa": doa, "doa_peak": doa_peak, "spikes_signed": sgn, "threads": int(used)}
