import sys, time, numpy as np, torch
sys.path.insert(0,'.')
import bench
from bamqc_b200 import Engine, synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 3000000
scale = float(sys.argv[2]) if len(sys.argv) > 2 else 0.3
staging = int(sys.argv[3]) if len(sys.argv) > 3 else 256
genome, records, offsets = bench.make_workload(n, 0, 1, scale=scale, threads=16)
n_bytes=int(offsets[-1])
comp = synth.bgzf_compress(records[:n_bytes], level=6)
eng = Engine(lane_ids=["L1"], ref_names=genome.names, staging_bytes=staging<<20)
for rid,(p,nn) in enumerate(zip(genome.packed, genome.lengths)): eng.set_reference(rid,p,nn)
pinned = torch.empty(comp.size, dtype=torch.uint8, pin_memory=True); pin_np=pinned.numpy(); pin_np[:]=comp
for rep in range(3):
    eng.reset(); eng.profile_enable(True); eng.profile_read()
    t0=time.perf_counter()
    eng.submit_bgzf(pin_np, last=True)
    t1=time.perf_counter(); eng.finish(); t2=time.perf_counter(); eng.scalars(); t3=time.perf_counter()
    pr=eng.profile_read()
    print("submit %.1f finish %.1f fetch %.1f total %.1f ms | %s"%((t1-t0)*1e3,(t2-t1)*1e3,(t3-t2)*1e3,(t3-t0)*1e3, {k:round(v[0],1) for k,v in pr.items() if v[0]}), flush=True)
