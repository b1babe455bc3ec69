import sys, time, numpy as np, torch
sys.path.insert(0,'.')
import bench
from bamqc_b200 import Engine
genome, records, offsets = bench.make_workload(int(sys.argv[1]) if len(sys.argv) > 1 else 3000000, 0, 1, scale=float(sys.argv[2]) if len(sys.argv) > 2 else 0.3, threads=16)
n_bytes=int(offsets[-1])
eng = Engine(lane_ids=["L1"], ref_names=genome.names, staging_bytes=256<<20)
for rid,(p,n) in enumerate(zip(genome.packed, genome.lengths)): eng.set_reference(rid,p,n)
pinned = torch.empty(n_bytes+64, dtype=torch.uint8, pin_memory=True); pin_np=pinned.numpy(); pin_np[:]=records[:n_bytes+64]
bounds = bench.split_batches(offsets, (256<<20)-4096)
for rep in range(3):
    eng.reset()
    t0=time.perf_counter(); ts=[]
    for lo,hi in zip(bounds[:-1],bounds[1:]):
        o=offsets[lo:hi+1]; t=time.perf_counter(); eng.submit(pin_np[int(o[0]):int(o[-1])], None); ts.append(time.perf_counter()-t)
    t1=time.perf_counter(); eng.finish(); t2=time.perf_counter(); eng.scalars(); t3=time.perf_counter()
    print("submits %s ms; submit total %.1f finish %.1f fetch %.1f total %.1f ms"%([round(x*1e3,1) for x in ts],(t1-t0)*1e3,(t2-t1)*1e3,(t3-t2)*1e3,(t3-t0)*1e3), flush=True)
