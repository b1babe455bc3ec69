"""Kernel-only time of a multi-@RG batch: every lane's pass over its own index list (default) vs every lane's pass
filtering the whole batch (BQC_LANE_INDEX=0).  python profiles/probes/lanes_probe.py [n_lanes] [n_pairs]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from bamqc_b200 import Engine, synth
n_lanes = int(sys.argv[1]) if len(sys.argv) > 1 else 4
n_pairs = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
genome = synth.Genome.make(5, ["chr1", "chr2"], [40_000_000, 20_000_000])
lib_ = synth.Library(seed=9, n_pairs=n_pairs, n_lanes=n_lanes)
records, offsets = synth.generate(genome, lib_)
eng = Engine(lane_ids=synth.lane_ids(lib_), ref_names=genome.names, chroms="chr1,chr2")
for rid, (p, n) in enumerate(zip(genome.packed, genome.lengths)):
    eng.set_reference(rid, p, n)
b = eng.prepare(records[:int(offsets[-1])], offsets)
st = torch.cuda.ExternalStream(eng.stream)
ts = []
for it in range(6):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st); eng.reset(); eng.run(b); eng.finish(); e1.record(st); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
print(f"lanes {n_lanes} records {len(offsets) - 1} BQC_LANE_INDEX={os.environ.get('BQC_LANE_INDEX', '1')}: {min(ts[2:]):.2f} ms per pass, {(len(offsets) - 1) / min(ts[2:]) / 1e3:.1f} M records/s; readcounts {[eng.scalars(l)['readcount'] for l in range(n_lanes)]}")
