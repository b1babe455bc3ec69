import sys, time, os
sys.path.insert(0, '/root/repo')
import numpy as np, torch
import bench
from bamqc_b200 import synth
spec = bench.workload_spec(sys.argv[1], 1)
genome, records, offsets = bench.make_workload(spec, 0, 1, spec["records_per_gpu"], threads=8)
eng = bench.make_engine(spec, genome, 0, 256)
cudart = torch.cuda.cudart()
print("register", cudart.cudaHostRegister(records.ctypes.data, records.nbytes, 0))
bounds = bench.split_batches(offsets, (256 << 20) - 4096)
for it in range(4):
    t0 = time.perf_counter(); eng.reset(); t1 = time.perf_counter()
    ts = []
    for lo, hi in zip(bounds[:-1], bounds[1:]):
        o = offsets[lo:hi + 1]
        a = time.perf_counter(); eng.submit(records[int(o[0]):int(o[-1])], None); ts.append(time.perf_counter() - a)
    t2 = time.perf_counter(); eng.finish(); t3 = time.perf_counter(); eng.scalars(); t4 = time.perf_counter()
    print(f"reset {1e3*(t1-t0):.2f} submits {[round(1e3*x,2) for x in ts]} finish {1e3*(t3-t2):.2f} scalars {1e3*(t4-t3):.2f} total {1e3*(t4-t0):.2f}")
