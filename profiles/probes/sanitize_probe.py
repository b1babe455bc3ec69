"""Small end-to-end pass over every kernel (host-framed, device-framed stream, BGZF with device inflate, resident
batches, multi k/q, three read groups) for compute-sanitizer runs:
    compute-sanitizer --tool memcheck python profiles/probes/sanitize_probe.py"""
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import bqc_testutil as util
from bamqc_b200 import Engine, synth

genome = util.small_genome()
lib_ = synth.Library(seed=99, n_pairs=1500).stress()
records, offsets = synth.generate(genome, lib_)
n_bytes = int(offsets[-1])
comp = synth.bgzf_compress(records[:n_bytes], level=6)
outs = []
with tempfile.TemporaryDirectory() as td:
    for mode in ("offsets", "stream", "bgzf"):
        eng = Engine(lane_ids=synth.lane_ids(lib_), ref_names=genome.names, chroms="chr1,chr2", klist=(15, 32), qlist=(17,),
                     staging_bytes=1 << 18, cov_ring_log2=13)
        for rid, (p, n) in enumerate(zip(genome.packed, genome.lengths)):
            eng.set_reference(rid, p, n)
        if mode == "offsets":
            for lo in range(0, len(offsets) - 1, 700):   # host-framed slices that fit the small staging buffers
                o = offsets[lo:lo + 701]
                eng.submit(records[int(o[0]):int(o[-1])], o - o[0])
        elif mode == "stream":
            step = 70001
            for p in range(0, n_bytes, step):
                eng.submit_stream(records[p:min(n_bytes, p + step)], last=p + step >= n_bytes)
        else:
            eng.submit_bgzf(comp, last=True)
        eng.finish()
        path = os.path.join(td, mode + ".bamqc")
        eng.write_bamqc("S1", path)
        outs.append(open(path).read())
        eng.close()
    # resident batches (the kernel-only path of bench.py) and three read groups (lane index lists, k_frame_lanes)
    eng = Engine(lane_ids=synth.lane_ids(lib_), ref_names=genome.names, chroms="chr1,chr2", klist=(15, 32), qlist=(17,), staging_bytes=1 << 18)
    for rid, (p, n) in enumerate(zip(genome.packed, genome.lengths)):
        eng.set_reference(rid, p, n)
    half = (len(offsets) - 1) // 2
    batches = []
    for lo, hi in ((0, half), (half, len(offsets) - 1)):
        o = offsets[lo:hi + 1]
        batches.append(eng.prepare(records[int(o[0]):int(o[-1])], o - o[0]))
    for b in batches:
        eng.run(b)
    eng.finish()
    path = os.path.join(td, "resident.bamqc")
    eng.write_bamqc("S1", path)
    outs.append(open(path).read())
    eng.close()
    lib3 = synth.Library(seed=98, n_pairs=900, n_lanes=3)
    rec3, off3 = synth.generate(genome, lib3)
    comp3 = synth.bgzf_compress(rec3[:int(off3[-1])], level=6)
    eng = Engine(lane_ids=synth.lane_ids(lib3), ref_names=genome.names, chroms="chr1,chr2", staging_bytes=1 << 18)
    for rid, (p, n) in enumerate(zip(genome.packed, genome.lengths)):
        eng.set_reference(rid, p, n)
    eng.submit_bgzf(comp3, last=True)
    eng.finish()
    eng.close()
assert outs[0] == outs[1] == outs[2] == outs[3], "paths disagree"
print("sanitize probe ok:", len(offsets) - 1, "records, four paths identical; three-lane BGZF pass done")
