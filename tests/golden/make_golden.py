#!/usr/bin/env python
"""Regenerate the committed golden fixtures.  Run in the build container (needs /root/reference):

    make -C oracle all ref && python tests/golden/make_golden.py

* kmerstream_golden.json : known-answer vectors produced by the REFERENCE's own kmerstream sources
  (oracle/_ref/libkmerstream_ref.so = src/kmerstream/{RepHash.cpp,lsb.cpp,StreamCounter.hpp,RepHash.hpp,
  mersennetwister.h} + src/ReadQualityHasher.hpp, compiled unmodified).
* genome.fa, <case>.bam, <case>.bamqc : small seeded synthetic inputs and the `.bamqc` written by
  oracle/_ref/bamqualcheck_ref, i.e. the reference's own src/bamqualcheck.cpp + statistics headers compiled
  unmodified over the SeqAn stand-in oracle/miniseqan (SeqAn 1.4.2 itself is unavailable offline).
"""
import ctypes
import json
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

CASES = {
    # name: (library kwargs, stress?, cli options)
    "standard": (dict(seed=101, n_pairs=1500), False, ["-c", "chr1,chr2"]),
    "stress": (dict(seed=102, n_pairs=1200), True, ["-c", "chr1,chr2"]),
    "two_lanes_kq": (dict(seed=103, n_pairs=1000, n_lanes=2), False, ["-c", "chr1,chr2", "-k", "15,32,63", "-q", "10,17"]),
    "long_insert": (dict(seed=104, n_pairs=1000, ins_mean=1500, ins_sd=400, ins_min=150, ins_max=6000), False,
                    ["-c", "chr1,chr2,chrX", "-i", "3000", "-s", "7"]),
}
GENOME = dict(seed=77, names=["chr1", "chr2", "chrX"], lengths=[60000, 40000, 20000])


def case_inputs(name):
    from bamqc_b200 import synth
    kw, stress, opts = CASES[name]
    genome = synth.Genome.make(GENOME["seed"], GENOME["names"], GENOME["lengths"])
    lib_ = synth.Library(**kw)
    if stress:
        lib_.stress()
    records, offsets = synth.generate(genome, lib_)
    return genome, lib_, records, offsets, opts


def main():
    from bamqc_b200 import synth
    ref_bin = os.path.join(ROOT, "oracle", "_ref", "bamqualcheck_ref")
    ref_lib = os.path.join(ROOT, "oracle", "_ref", "libkmerstream_ref.so")
    assert os.path.exists(ref_bin) and os.path.exists(ref_lib), "run `make -C oracle ref` first"
    for name in CASES:
        genome, lib_, records, offsets, opts = case_inputs(name)
        fasta = os.path.join(HERE, "genome.fa")  # every case uses the same seeded genome
        bam, out = (os.path.join(HERE, f"{name}.{e}") for e in ("bam", "bamqc"))
        genome.write_fasta(fasta)
        synth.write_bam(bam, genome, lib_, records, int(offsets[-1]), level=6)
        r = subprocess.run([ref_bin, "-r", fasta, "-o", out] + opts + [bam], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        print(name, len(offsets) - 1, "records", os.path.getsize(bam), "bytes BAM")

    # ---- kmerstream known-answer vectors ---------------------------------------------------------------
    R = ctypes.CDLL(ref_lib)
    u64p = ctypes.POINTER(ctypes.c_uint64)
    g = {}
    hv = np.zeros(64, dtype=np.uint64)
    R.ref_rephash_hvals(1, hv.ctypes.data_as(u64p))
    g["hvals_seed1"] = ["%016x" % x for x in hv]
    seq = "ACGTACGTTGCAAGCTTAGGCATCGATCGGATCCATGCAAGT"
    g["seq42"] = seq
    g["windows"] = {}
    for seed in (1, 7):
        for k in (1, 5, 15, 31, 32, 33, 63):
            if k > len(seq):
                continue
            out = np.zeros(64, dtype=np.uint64)
            n = R.ref_rephash_windows(seed, k, seq.encode(), len(seq), out.ctypes.data_as(u64p))
            g["windows"][f"{seed}:{k}"] = ["%016x" % x for x in out[:n]]
    R.ref_bitscan.restype = ctypes.c_uint64
    R.ref_bitscan.argtypes = [ctypes.c_uint64]
    g["bitscan"] = {str(v): int(R.ref_bitscan(v)) for v in (0, 1, 8, 2 ** 31, 2 ** 40, 2 ** 63, 12345678)}
    R.ref_streamcounter.argtypes = [ctypes.c_double, ctypes.c_int, u64p, ctypes.c_uint64, u64p, u64p, u64p]
    g["streamcounter"] = {}
    import random
    for n in (0, 1000, 200000):
        rng = random.Random(42 + n)
        h = np.array([rng.getrandbits(64) for _ in range(n)] + [0] * (1 if n else 0), dtype=np.uint64)
        if n:
            h[::3] = h[0]
        out = np.zeros(6, dtype=np.uint64)
        R.ref_streamcounter(0.01, 1, h.ctypes.data_as(u64p), len(h), out.ctypes.data_as(u64p), None, None)
        g["streamcounter"][str(n)] = [int(x) for x in out]
    R.ref_hasher.argtypes = [ctypes.c_double, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_char_p, ctypes.c_char_p,
                             ctypes.POINTER(ctypes.c_int), ctypes.c_int, u64p, u64p, u64p]
    g["hasher"] = {}
    for (q, k) in ((17, 32), (10, 15), (30, 63)):
        seqs, quals, lens = hasher_reads(5)
        out = np.zeros(4, dtype=np.uint64)
        la = (ctypes.c_int * len(lens))(*lens)
        R.ref_hasher(0.01, 1, q, k, seqs, quals, la, len(lens), out.ctypes.data_as(u64p), None, None)
        g["hasher"][f"{q}:{k}"] = [int(x) for x in out]
    # the fixed read of SURVEY Appendix D.4
    fixed = seq + "N" + seq
    qual = bytearray(b"I" * len(fixed))
    qual[50] = ord("1")
    out = np.zeros(4, dtype=np.uint64)
    la = (ctypes.c_int * 1)(len(fixed))
    R.ref_hasher(0.01, 1, 17, 32, fixed.encode(), bytes(qual), la, 1, out.ctypes.data_as(u64p), None, None)
    g["hasher_fixed_read_sumcount"] = int(out[0])
    with open(os.path.join(HERE, "kmerstream_golden.json"), "w") as f:
        json.dump(g, f, indent=1)
    print("kmerstream_golden.json written")


def hasher_reads(seed, n=2000):
    import random
    rng = random.Random(seed)
    seqs, quals, lens = bytearray(), bytearray(), []
    for _ in range(n):
        L = rng.randint(20, 220)
        for _ in range(L):
            seqs.append(ord("N") if rng.random() < 0.02 else rng.choice(b"ACGT"))
            quals.append(33 + (rng.randint(0, 16) if rng.random() < 0.04 else rng.randint(17, 40)))
        lens.append(L)
    return bytes(seqs), bytes(quals), lens


if __name__ == "__main__":
    main()
