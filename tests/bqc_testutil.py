"""Shared helpers of the parity tests: build/run the CPU oracle (test infrastructure under oracle/),
make synthetic data sets, run the CUDA engine through the C ABI and diff the `.bamqc` outputs."""
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_BIN = os.path.join(ORACLE_DIR, "bamqualcheck_oracle")


def ensure_oracle():
    src = os.path.join(ORACLE_DIR, "bamqc_oracle.cpp")
    if not os.path.exists(ORACLE_BIN) or os.path.getmtime(ORACLE_BIN) < os.path.getmtime(src):
        subprocess.run(["make", "-C", ORACLE_DIR, "all"], check=True, capture_output=True)
    return ORACLE_BIN


def run_oracle(bam, fasta, out, chroms=None, isize=None, klist=None, qlist=None, e=None, seed=None, dump=None, extra=()):
    cmd = [ensure_oracle(), "-r", str(fasta), "-o", str(out)]
    if chroms is not None:
        cmd += ["-c", chroms]
    if isize is not None:
        cmd += ["-i", str(isize)]
    if klist is not None:
        cmd += ["-k", ",".join(map(str, klist))]
    if qlist is not None:
        cmd += ["-q", ",".join(map(str, qlist))]
    if e is not None:
        cmd += ["-e", repr(e)]
    if seed is not None:
        cmd += ["-s", str(seed)]
    if dump is not None:
        cmd += ["--dump", str(dump)]
    cmd += list(extra) + [str(bam)]
    return subprocess.run(cmd, capture_output=True, text=True)


def small_genome(seed=7, lengths=(400000, 300000, 100000), names=("chr1", "chr2", "chrX")):
    from bamqc_b200 import synth
    return synth.Genome.make(seed, list(names), list(lengths))


def diff_bamqc(a_path, b_path, max_report=8):
    """Return a list of human-readable differences between two .bamqc files (empty = identical)."""
    a = open(a_path).read().split("\n")
    b = open(b_path).read().split("\n")
    out = []
    if len(a) != len(b):
        out.append(f"line count {len(a)} vs {len(b)}")
    for i, (x, y) in enumerate(zip(a, b)):
        if x != y:
            xs, ys = x.split(" "), y.split(" ")
            key = xs[0]
            where = next((j for j, (p, q) in enumerate(zip(xs, ys)) if p != q), min(len(xs), len(ys)))
            out.append(f"line {i} key {key} len {len(xs)} vs {len(ys)} first diff at field {where}: "
                       f"{xs[where:where + 4]} vs {ys[where:where + 4]}")
            if len(out) >= max_report:
                break
    return out


def run_engine(genome, lib_, records, offsets, out_path, chroms="chr1,chr2", isize=1000, klist=(32,), qlist=(17,),
               e=0.01, seed=1, n_batches=1, sample_id="S1", resident=False, engine_kwargs=None, keep=False):
    """Push `records` through the CUDA engine via the C ABI and write the .bamqc text."""
    from bamqc_b200 import Engine, synth
    eng = Engine(lane_ids=synth.lane_ids(lib_), ref_names=genome.names, chroms=chroms, isize=isize, klist=klist,
                 qlist=qlist, e=e, seed=seed, **(engine_kwargs or {}))
    for rid, (p, n) in enumerate(zip(genome.packed, genome.lengths)):
        eng.set_reference(rid, p, n)
    nrec = len(offsets) - 1
    bounds = [nrec * i // n_batches for i in range(n_batches + 1)]
    batches = []
    for i in range(n_batches):
        lo, hi = bounds[i], bounds[i + 1]
        if hi == lo:
            continue
        o = offsets[lo:hi + 1]
        if resident:
            batches.append(eng.prepare(records[int(o[0]):int(o[-1])], o - o[0]))
        else:
            eng.submit(records, o)
    if resident:
        for rep in range(2):  # replay twice: results must come from the second, post-reset pass
            eng.reset()
            for b in batches:
                eng.run(b)
            eng.finish()
    else:
        eng.finish()
    eng.write_bamqc(sample_id, out_path)
    if keep:
        return eng
    for b in batches:
        b.free()
    eng.close()
    return None


def oracle_sketch(dump_path, n_sketches, size=32768, f2size=32768):
    raw = np.fromfile(str(dump_path) + ".sketch", dtype=np.uint64)
    per = 32 * size + f2size
    assert raw.size == per * n_sketches
    return [(raw[i * per:i * per + 32 * size], raw[i * per + 32 * size:(i + 1) * per]) for i in range(n_sketches)]
