"""Stateful frames (micloc_snn_stream_*) and the device Envelope tracker.

N pushed frames must give exactly what one clip of their concatenation gives; a clip whose last K/2 samples are
zero makes the stream's causal in-phase delay coincide with the reference's np.roll, so that the one-shot path
(and through it the reference goldens / oracle) is the yardstick."""
import numpy as np
import pytest
import torch

import helpers as H
from oracle import oracle as O

pytestmark = pytest.mark.gpu


def make_engine(g, T):
    from haghighatshoarmuir2024_b200.engine import SnnEngine
    return SnnEngine(H.chain_spec(g, T), g["bf_mat"], device=0)


def stream_all(stream, x, frame_lens, want_env=False):
    parts = {"spikes": [], "env": [], "doa_t": []}
    frames = []
    t = 0
    for n in frame_lens:
        out = stream.push(x[t:t + n], want_env=want_env)
        t += n
        frames.append(out)
        for k in parts:
            if out.get(k) is not None and out["n_out"]:
                parts[k].append(out[k])
    assert t == x.shape[0]
    out = stream.flush(want_env=want_env)
    frames.append(out)
    for k in parts:
        if out.get(k) is not None and out["n_out"]:
            parts[k].append(out[k])
    cat = {k: (torch.cat(v) if v else None) for k, v in parts.items()}
    return cat, frames, out["flags"]


@pytest.mark.parametrize("dtype", [torch.float32, torch.int16, torch.int32])
def test_frames_equal_one_clip_bit_for_bit(dtype):
    from haghighatshoarmuir2024_b200.engine import SnnStream
    g = H.load("snn_c1_bipolar")
    T = 9600
    x, _ = H.synth_clips(g, 1, T, seed=11, int16=dtype != torch.float32)
    x = x[0].copy()
    x[-len(g["kernel"]) // 2:] = 0                       # zero tail: np.roll's wrap-around brings in zeros
    eng = make_engine(g, T)
    xd = torch.from_numpy(x).cuda()
    if dtype == torch.int32:
        # the recorder's frames: 8 interleaved int32 channels, the last one is not a microphone
        xd = torch.cat([xd.to(torch.int32), torch.full((T, 1), 12345, dtype=torch.int32, device="cuda")], dim=1)
    one = eng.run_taps(torch.from_numpy(x).cuda(), want=("spikes", "vmem"))
    torch.cuda.synchronize()
    st = SnnStream(eng, max_frame_len=4096)
    assert st.latency > 0
    for lens in ([T // 3] * 3, [1, 31, 32, 33, 700, 4096, 2, 609, T - 5504], [4096, 4096, T - 8192]):
        st.reset()
        cat, frames, flags = stream_all(st, xd, lens)
        assert flags == 0
        assert cat["spikes"].shape == (T, 14)
        assert torch.equal(cat["spikes"], one["spikes"][0]), lens            # bit for bit
        assert sum(f["n_out"] for f in frames) == T
    # and against the float64 oracle (reference tolerance)
    cfg = H.oracle_cfg(g)
    cfg.nir = O.neuron_kernel(np.arange(T) / float(g["fs"]), float(g["tau"]), float(g["tau"]))
    ref = O.snn_apply(cfg, x.astype(np.float64), want=("spikes",))
    assert H.spike_agreement(cat["spikes"].cpu().numpy(), ref["spikes"]) >= 0.999


def test_frame_power_and_doa_match_the_oracle_window():
    from haghighatshoarmuir2024_b200.engine import SnnStream
    g = H.load("snn_c1_bipolar")
    T = 9600
    x, _ = H.synth_clips(g, 1, T, seed=5, snrs_db=(15.0,))
    x = x[0].copy(); x[-240:] = 0
    eng = make_engine(g, T)
    one = eng.run_taps(torch.from_numpy(x).cuda(), want=("vmem",))
    vm = one["vmem"][0].cpu().numpy().astype(np.float64)
    st = SnnStream(eng, max_frame_len=4800)
    f1 = st.push(torch.from_numpy(x[:4800]).cuda())
    f2 = st.push(torch.from_numpy(x[4800:]).cuda())
    f3 = st.flush()
    lo = 0
    for f in (f1, f2, f3):
        n = f["n_out"]
        y = vm[lo:lo + n] @ g["bf_mat"]
        power = np.mean(y ** 2, axis=0)
        assert H.rel_err(f["power"].cpu().numpy(), power) < 1e-4
        assert int(f["doa"][0]) == int(np.argmax(power))
        lo += n
    assert lo == T and f1["n_out"] == 4800 - st.latency and f3["n_out"] == st.latency


def test_envelope_matches_reference_golden_and_streams():
    from haghighatshoarmuir2024_b200.engine import envelope
    from haghighatshoarmuir2024_b200.utils import Envelope
    g = H.load("envelope")
    for i in range(int(g["n_cases"])):
        x = g[f"x_{i}"]
        env, idx = envelope(torch.from_numpy(x.astype(np.float32)).cuda(), float(g["fs"]), float(g[f"rise_{i}"]),
                            float(g[f"fall_{i}"]), want_argmax=True)
        ref = g[f"env_{i}"]
        assert H.rel_err(env.cpu().numpy(), ref) < 2e-5                       # float32 recurrence vs float64
        same = (idx.cpu().numpy() == np.argmax(ref, axis=1)).mean()
        assert same > 0.995                                                   # near-ties between channels may flip
        y = Envelope(float(g[f"rise_{i}"]), float(g[f"fall_{i}"]), float(g["fs"])).evolve(x)      # drop-in class
        assert y.dtype == np.float64 and H.rel_err(y, ref) < 2e-5
    with pytest.raises(ValueError):
        envelope(torch.zeros((10, 3), device="cuda"), 48000.0, 1.0, 0.1)


def test_stream_envelope_continues_across_frames():
    from haghighatshoarmuir2024_b200.engine import SnnStream, envelope
    g = H.load("snn_c1_bipolar")
    T = 6000
    x, _ = H.synth_clips(g, 1, T, seed=3, snrs_db=(20.0,))
    x = x[0].copy(); x[-240:] = 0
    eng = make_engine(g, T)
    y = eng.run_taps(torch.from_numpy(x).cuda(), want=("y",))["y"][0]
    env_ref, idx_ref = envelope(y, 48000.0, 10e-3, 100e-3, want_argmax=True)
    st = SnnStream(eng, max_frame_len=2048)
    cat, _, _ = stream_all(st, torch.from_numpy(x).cuda(), [2048, 2048, T - 4096], want_env=True)
    assert torch.equal(cat["env"], env_ref) and torch.equal(cat["doa_t"], idx_ref)


def test_stream_argument_errors():
    from haghighatshoarmuir2024_b200.engine import SnnStream
    g = H.load("snn_c1_bipolar")
    eng = make_engine(g, 4800)
    st = SnnStream(eng, max_frame_len=100)
    with pytest.raises(ValueError):
        st.push(torch.zeros((50, 6), device="cuda"))                          # fewer channels than microphones
    with pytest.raises(ValueError):
        st.push(torch.zeros((101, 7), device="cuda"))                         # longer than max_frame_len
    st.push(torch.zeros((100, 7), device="cuda"))
    st.flush()
    with pytest.raises(ValueError):
        st.push(torch.zeros((10, 7), device="cuda"))                          # flushed: reset first
    st.reset()
    assert st.push(torch.zeros((100, 7), device="cuda"))["n_out"] == 0        # shorter than the latency: nothing final yet
