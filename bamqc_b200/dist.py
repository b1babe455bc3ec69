"""Multi-GPU plumbing: one process per GPU, records sharded with no data-path collective, and a single
end-of-run reduction of the per-GPU result blocks over NCCL (torch.distributed is only the transport).

Merge rule = StreamCounter::join (src/kmerstream/StreamCounter.hpp:95-112): counters, F2 tables and
sumCount add; every 4-bit sketch counter becomes min(sum, 15).  The sketch travels as one uint8 per counter
so that the sum of up to 17 ranks cannot overflow before the clamp (done by bqc_sketch_import_u8).
"""
import numpy as np


def shard_regions(lengths, rank, world, gap=5000, margin=0):
    """Genome slice of `rank`: {contig: (begin, end)}.  Neighbouring shards are separated by `gap` bp with no
    fragment starts, so that the first qualifying read of a shard always re-anchors the coverage windows
    (beginPos - shift > 2*vsize, src/OverallNumbers.hpp:91) and per-shard coverage histograms add exactly."""
    assert gap > 2000 + 1000
    out = {}
    for c, n in enumerate(lengths):
        lo = n * rank // world
        hi = n * (rank + 1) // world
        if rank > 0:
            lo += gap
        if hi - lo > margin:
            out[c] = (lo, hi)
    return out


def reduce_tensors(counters, sketch_u8, group=None):
    """All-reduce (sum) the counter block (int64 view of the uint64 counters) and the uint8 sketch in place."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(counters, op=dist.ReduceOp.SUM, group=group)
        if sketch_u8.numel():
            dist.all_reduce(sketch_u8, op=dist.ReduceOp.SUM, group=group)
    return counters, sketch_u8


def reduce_engine(engine, bufs=None, group=None):
    """Export this GPU's tables, all-reduce them over NCCL, import the merged tables back (every rank ends
    with the whole-job result).  `bufs` lets the caller reuse the two device tensors across steps."""
    import torch
    dev = torch.device("cuda", engine.device)
    if bufs is None:
        bufs = (torch.empty(engine.counters_len(), dtype=torch.int64, device=dev),
                torch.empty(max(1, engine.sketch_len()), dtype=torch.uint8, device=dev))
    c, s = bufs
    engine.export_to(c.data_ptr(), s.data_ptr())
    reduce_tensors(c, s[: engine.sketch_len()], group)
    torch.cuda.synchronize(dev)
    engine.import_from(c.data_ptr(), s.data_ptr())
    return bufs


def clamp_sketch_numpy(total_u8):
    """Host restatement of the import clamp (for CPU tests)."""
    return np.minimum(total_u8, 15).astype(np.uint8)
