"""The N > 1 merge on CPU: world_size 2 over gloo.  Checks the reduce helper used by bench.py / the engine
against the StreamCounter::join rule (src/kmerstream/StreamCounter.hpp:95-112): counters add, every 4-bit
sketch counter becomes min(sum, 15)."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from bamqc_b200 import dist as bdist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(100 + rank)
    counters = torch.from_numpy(rng.integers(0, 2 ** 40, size=5000, dtype=np.int64))
    sketch = torch.from_numpy(rng.integers(0, 16, size=70000, dtype=np.uint8))
    mine = (counters.clone().numpy(), sketch.clone().numpy())
    bdist.reduce_tensors(counters, sketch)
    np.savez(os.path.join(out_dir, f"r{rank}.npz"), c0=mine[0], s0=mine[1], c=counters.numpy(), s=sketch.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_reduce_tensors_world2(tmp_path):
    import torch.multiprocessing as mp
    from bamqc_b200 import dist as bdist
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r = [np.load(tmp_path / f"r{i}.npz") for i in range(2)]
    want_c = r[0]["c0"] + r[1]["c0"]
    want_s = r[0]["s0"].astype(np.uint16) + r[1]["s0"].astype(np.uint16)
    for i in range(2):
        assert np.array_equal(r[i]["c"], want_c)
        assert np.array_equal(r[i]["s"].astype(np.uint16), want_s)          # no overflow before the clamp
        assert np.array_equal(bdist.clamp_sketch_numpy(r[i]["s"]), np.minimum(want_s, 15).astype(np.uint8))
