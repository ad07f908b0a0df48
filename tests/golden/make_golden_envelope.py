"""Golden fixtures of micloc.utils.Envelope.evolve (micloc/utils.py:15-81), run from the unmodified reference.
Usage: python tests/golden/make_golden_envelope.py"""
import os
import sys

sys.path.insert(0, "/root/reference")
import numpy as np

from micloc.utils import Envelope

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    rng = np.random.default_rng(77)
    fs = 48_000.0
    out = {"fs": fs}
    cases = [(10e-3, 100e-3, 3000, 9), (1e-3, 1e-2, 2000, 64), (5e-3, 5e-3, 1500, 3)]
    for i, (rise, fall, T, C) in enumerate(cases):
        t = np.arange(T)[:, None] / fs
        # amplitude-modulated tones with gaps: rise and fall phases in every channel
        x = np.sin(2 * np.pi * (500 + 40 * np.arange(C))[None, :] * t) * (0.2 + np.abs(np.sin(2 * np.pi * 7 * t + np.arange(C)[None, :])))
        x = x * (rng.random((T, C)) > 0.05) + 0.01 * rng.standard_normal((T, C))
        out[f"x_{i}"], out[f"rise_{i}"], out[f"fall_{i}"] = x, rise, fall
        out[f"env_{i}"] = Envelope(rise_time=rise, fall_time=fall, fs=fs).evolve(x)
    out["n_cases"] = len(cases)
    np.savez_compressed(os.path.join(HERE, "envelope.npz"), **out)
    print({k: np.shape(v) for k, v in out.items()})


if __name__ == "__main__":
    main()
