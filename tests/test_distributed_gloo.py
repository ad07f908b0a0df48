"""world_size-2 gloo test of the multi-GPU plumbing (clip sharding + histogram sum)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from haghighatshoarmuir2024_b200 import distributed as D


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, total, G, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(123)                       # same stream on every rank
    doa_all = torch.from_numpy(rng.integers(0, G, size=total).astype(np.int32))
    grp_all = torch.from_numpy(rng.integers(0, 3, size=total))
    lo, hi = D.shard_range(total, rank, world)
    hist = D.doa_histogram(doa_all[lo:hi], G, grp_all[lo:hi], 3)
    hist = D.reduce_histograms(hist)
    ms = D.max_over_ranks(10.0 + rank, torch.device("cpu"))
    want = D.doa_histogram(doa_all, G, grp_all, 3)
    ok = bool(torch.equal(hist, want)) and ms == 10.0 + world - 1 and int(hist.sum()) == total
    open(os.path.join(out_dir, f"rank{rank}.ok" if ok else f"rank{rank}.bad"), "w").close()
    dist.destroy_process_group()


def test_sharded_histograms_sum_to_the_global_one(tmp_path):
    world, port = 2, _free_port()
    mp.spawn(_worker, args=(world, port, 1001, 17, str(tmp_path)), nprocs=world, join=True)
    assert sorted(os.listdir(tmp_path)) == ["rank0.ok", "rank1.ok"]
