"""Which mbarrier wait of the tensor-core fused kernel times out (needs tools/libmicloc_b200_rt.so: MICLOC_WAIT_DEBUG)."""
import ctypes, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers as H
from haghighatshoarmuir2024_b200 import _native as N
from haghighatshoarmuir2024_b200.engine import SnnEngine
g = H.load("snn_c1_bipolar")
B, T = int(sys.argv[1]), int(sys.argv[2])
x, _ = H.synth_clips(g, min(B, 64), T, seed=9)
x = np.tile(x, (B // x.shape[0] + 1, 1, 1))[:B]
eng = SnnEngine(H.chain_spec(g, T), g["bf_mat"], device=0)
out = eng.run(torch.from_numpy(x).cuda(), want_spikes=True, fused=True)
lib = ctypes.CDLL(N.LIB_PATH)
if not hasattr(lib, "micloc_debug_wait"):
    torch.cuda.synchronize(); print("B", B, "T", T, "ok (product build)"); sys.exit(0)
buf = (ctypes.c_uint64 * 8)()
lib.micloc_debug_wait(buf)
v = list(buf)
print("B", B, "T", T, "first timed-out wait:", "none" if v[0] == 0 else f"tag {v[0]-1} (kind {(v[0]-1)//100000}, step {(v[0]-1)%100000-8}) block {v[1]} thread {v[2]} (warp {v[2]//32}) parity {v[3]} addr {v[4]:#x} barrier word {v[5]:#018x} last part (it, ksteps) ({v[6] >> 32}, {v[6] & 0xffffffff}) commits {v[7]}")
