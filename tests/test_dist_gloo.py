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


# ---- one record stream cut across ranks: the coverage protocol (include/bamqc_b200.h, cov_math.h "Shards") ---------
class _MockShardEngine:
    """Plays back one piece dumped by tests/cov_selftest.cpp (--dump-shards): what an engine in shard mode answers.
    It checks that the protocol hands every piece the right predecessor and entry state."""

    def __init__(self, piece):
        from bamqc_b200 import _lib
        self.p = piece
        self.lib = _lib.load_library()   # bqc_cov_shards_combine / bqc_cov_apply are host code: no GPU needed
        self._lib = _lib

    def cov_shard_boundary(self):
        sh = self._lib.bqc_cov_shard()
        sh.n, sh.first_rid, sh.first_b, sh.last_rid, sh.last_b = (self.p[k] for k in ("n", "first_rid", "first_b", "last_rid", "last_b"))
        return sh

    def cov_shard_function(self, have_prev, prev_rid, prev_b):
        if self.p["n"]:
            assert have_prev == self.p["have_prev"]
            if have_prev:
                assert (prev_rid, prev_b) == (self.p["prev_rid"], self.p["prev_b"])
        return np.array(self.p["table"], dtype=np.uint16)

    def cov_shard_run(self, have_prev, prev_rid, prev_b, p_in):
        if self.p["n"]:
            assert p_in == self.p["p_in"], (p_in, self.p["p_in"])
        sh = self.cov_shard_boundary()
        sh.span = self.p["span"]
        for i in range(2001):
            sh.head[i] = self.p["head"][i]
            sh.tail[i] = self.p["tail"][i]
        return sh


def _shard_worker(rank, world, port, dump_path, out_dir):
    sys.path.insert(0, ROOT)
    import json
    import torch.distributed as dist
    from bamqc_b200 import dist as bdist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    d = json.load(open(dump_path))
    eng = _MockShardEngine(d["pieces"][rank])
    delta = bdist.resolve_coverage_shards(eng, rank, bdist.torch_exchange())
    np.save(os.path.join(out_dir, f"delta{rank}.npy"), delta)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_coverage_shard_protocol_over_gloo(tmp_path, world):
    """A coordinate-ordered stream cut at arbitrary records into `world` pieces (one of them empty when world is 3):
    the sum of the pieces' own histograms plus the combined correction equals the sequential reference."""
    import json
    import subprocess
    import torch.multiprocessing as mp
    exe = str(tmp_path / "cov_selftest")
    subprocess.run(["g++", "-O2", "-std=c++17", "-o", exe, os.path.join(ROOT, "tests", "cov_selftest.cpp")], check=True)
    dump = str(tmp_path / "shards.json")
    subprocess.run([exe, "--dump-shards", str(world), dump], check=True)
    port = _free_port()
    mp.spawn(_shard_worker, args=(world, port, dump, str(tmp_path)), nprocs=world, join=True)
    d = json.load(open(dump))
    total = np.sum([np.array(p["poscov"], dtype=np.int64) for p in d["pieces"]], axis=0)
    for r in range(world):
        delta = np.load(tmp_path / f"delta{r}.npy")
        assert np.array_equal(total + delta, np.array(d["expected"], dtype=np.int64))
