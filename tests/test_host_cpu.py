"""Host-side logic that needs no GPU: the C-ABI library loads and exports everything include/*.h declares,
BGZF / BAM header / framing / FASTA helpers, the synthetic generator, the sharding plan."""
import ctypes
import os
import re

import numpy as np
import pytest

import bqc_testutil as util

ROOT = util.ROOT


def test_library_exports_every_declared_symbol(lib):
    import bamqc_b200
    # the product library exports what include/bamqc_b200.h declares, the generator library what bamqc_synth.h declares
    for hdr, path, protos, least in (("bamqc_b200.h", bamqc_b200.library_path(), bamqc_b200._lib.PROTOTYPES, 40),
                                      ("bamqc_synth.h", os.path.join(os.path.dirname(bamqc_b200.library_path()), "libbamqc_synth.so"),
                                       bamqc_b200._lib.SYNTH_PROTOTYPES, 6)):
        text = open(os.path.join(ROOT, "include", hdr)).read()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        declared = set(re.findall(r"\b(bqc_[a-z0-9_]+)\s*\(", text))
        assert len(declared) > least
        raw = ctypes.CDLL(path)
        missing = [s for s in sorted(declared) if not hasattr(raw, s)]
        assert not missing, missing
        assert declared <= set(protos), sorted(declared - set(protos))


def test_engine_fails_loudly_without_gpu(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import bamqc_b200
    with pytest.raises(bamqc_b200.BamQCError) as ei:
        bamqc_b200.Engine(ref_names=["chr1"])
    assert "no CPU fallback" in str(ei.value)


def test_bgzf_roundtrip_and_header(lib):
    from bamqc_b200 import synth, _lib
    raw = np.fromfile(os.path.join(util.GOLDEN, "standard.bam"), dtype=np.uint8)
    out = np.zeros(4 << 20, dtype=np.uint8)
    sizes = set()
    for threads in (1, 4):
        n = lib.bqc_bgzf_inflate(raw.ctypes.data, raw.size, out.ctypes.data, out.size, threads)
        assert n > 0
        sizes.add(int(n))
    n = sizes.pop()
    assert not sizes
    hdr = _lib.bqc_bam_header()
    used = lib.bqc_parse_bam_header(out.ctypes.data, n, ctypes.byref(hdr))
    assert used > 0
    assert hdr.n_ref == 3 and [hdr.ref_names[i].decode() for i in range(3)] == ["chr1", "chr2", "chrX"]
    assert [hdr.ref_lengths[i] for i in range(3)] == [60000, 40000, 20000]
    assert hdr.n_lanes == 1 and hdr.lane_ids[0] == b"L1" and hdr.sample_id == b"S1"
    lib.bqc_free_bam_header(ctypes.byref(hdr))
    # framing finds every record and ends exactly at the end of the stream
    offs = np.zeros(100000, dtype=np.uint64)
    nrec = lib.bqc_frame_records(out.ctypes.data + used, n - used, offs.ctypes.data, offs.size)
    assert nrec == 3206 and int(offs[nrec]) == n - used
    # compress -> inflate is the identity, also for empty input
    rng = np.random.default_rng(1)
    for size in (0, 1, 70000, 300001):
        data = rng.integers(0, 4, size=size, dtype=np.uint8)
        comp = synth.bgzf_compress(data, level=1)
        back = np.zeros(size + 16, dtype=np.uint8)
        m = lib.bqc_bgzf_inflate(comp.ctypes.data, comp.size, back.ctypes.data, back.size, 3)
        assert m == size and np.array_equal(back[:size], data)
    # corrupt input is rejected
    bad = raw.copy()
    bad[0] = 0
    assert lib.bqc_bgzf_inflate(bad.ctypes.data, bad.size, out.ctypes.data, out.size, 2) == 0


def test_fasta_packing(lib, tmp_path):
    genome = util.golden_genome()
    f = lib.bqc_fasta_open(os.path.join(util.GOLDEN, "genome.fa").encode())
    assert f
    for name, length, packed in zip(genome.names, genome.lengths, genome.packed):
        p = ctypes.c_void_p()
        n = lib.bqc_fasta_contig(f, name.encode(), ctypes.byref(p))
        assert n == length
        got = np.frombuffer((ctypes.c_uint8 * ((length + 3) // 4)).from_address(p.value), dtype=np.uint8)
        want = packed[: (length + 3) // 4].copy()
        if length % 4:
            want[-1] &= (1 << (2 * (length % 4))) - 1
        assert np.array_equal(got, want)
    p = ctypes.c_void_p()
    assert lib.bqc_fasta_contig(f, b"nope", ctypes.byref(p)) == -1
    lib.bqc_fasta_close(f)
    # N and lower case: N -> A (Dna5 -> Dna keeps two bits), id cut at the first blank
    fa = tmp_path / "x.fa"
    fa.write_text(">c1 description here\nACGTNacgt\nNN\n>c2\tother\nTTTT\n")
    f = lib.bqc_fasta_open(str(fa).encode())
    n = lib.bqc_fasta_contig(f, b"c1", ctypes.byref(p))
    assert n == 11
    got = np.frombuffer((ctypes.c_uint8 * 3).from_address(p.value), dtype=np.uint8)
    codes = [(int(got[i // 4]) >> (2 * (i % 4))) & 3 for i in range(11)]
    assert codes == [0, 1, 2, 3, 0, 0, 1, 2, 3, 0, 0]
    assert lib.bqc_fasta_contig(f, b"c2", ctypes.byref(p)) == 4
    lib.bqc_fasta_close(f)


def test_generator_is_deterministic_and_sorted(lib):
    from bamqc_b200 import synth
    genome = util.golden_genome()
    a = synth.generate(genome, synth.Library(seed=5, n_pairs=800))
    b = synth.generate(genome, synth.Library(seed=5, n_pairs=800))
    c = synth.generate(genome, synth.Library(seed=6, n_pairs=800))
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    assert not np.array_equal(a[0][: min(a[0].size, c[0].size)], c[0][: min(a[0].size, c[0].size)])
    rec, offs = a
    keys = []
    for i in range(len(offs) - 1):
        o = int(offs[i])
        rid, pos = np.frombuffer(rec[o + 4:o + 12].tobytes(), dtype=np.int32)
        keys.append((rid if rid >= 0 else 1 << 30, pos))
        assert int(np.frombuffer(rec[o:o + 4].tobytes(), dtype=np.int32)[0]) + 4 == int(offs[i + 1]) - o
    assert keys == sorted(keys)
    mean = (int(offs[-1])) / (len(offs) - 1)
    assert 285 < mean < 295  # ~290 B per standard record


def test_shard_regions_have_reanchoring_gaps():
    from bamqc_b200 import dist
    lengths = [1000000, 700000, 50000]
    for world in (2, 4, 8):
        prev_end = [0] * len(lengths)
        for rank in range(world):
            reg = dist.shard_regions(lengths, rank, world)
            for c, (lo, hi) in reg.items():
                assert hi > lo and hi <= lengths[c]
                if rank > 0:
                    assert lo - prev_end[c] >= 5000  # > 2*vsize + read span: the next shard always re-anchors
                prev_end[c] = hi
        assert all(prev_end[c] == lengths[c] for c in range(len(lengths)))


def test_fasta_packer_matches_a_bytewise_restatement(tmp_path):
    """bqc_fasta_open packs on all host threads (chunks of one record share boundary bytes): same result as the
    byte-at-a-time rule -- '>' only at a line start opens a record, the name ends at the first blank/tab, every other
    non-white character is a base, C/G/T (either case) are 1/2/3 and everything else (A, N, IUPAC, '>') is 0."""
    import ctypes
    import random
    from bamqc_b200 import _lib
    L = _lib.load_library()
    rng = random.Random(11)

    def seq(n, alphabet="ACGTacgtNnRY"):
        return "".join(rng.choice(alphabet) for _ in range(n))

    big = seq(2_500_003, "ACGT")  # several chunks per record
    parts = ["junk before the first header\nACGT\n", ">chr1 description here\n"]
    for i in range(0, len(big), 61):
        parts.append(big[i:i + 61] + ("\r\n" if i % 7 == 0 else "\n"))
    parts += [">chr2\tother\n", seq(70) + "\n", "AC>GT  \t" + seq(13) + "\n\n", seq(5), "\n>empty\n>last no newline\n" + seq(1_200_001, "ACGTN")]
    text = "".join(parts)
    path = tmp_path / "odd.fa"
    path.write_bytes(text.encode())
    # byte-wise restatement
    want, cur, hdr, in_hdr, line_start = [], None, "", False, True
    for c in text:
        if in_hdr:
            if c == "\n":
                in_hdr, line_start = False, True
                name = hdr.split(" ")[0].split("\t")[0].rstrip("\r\n")
                cur = [name, []]
                want.append(cur)
            else:
                hdr += c
            continue
        if c == "\n":
            line_start = True
            continue
        if line_start and c == ">":
            in_hdr, hdr, line_start = True, "", False
            continue
        line_start = False
        if cur is None or c in "\r \t":
            continue
        cur[1].append({"C": 1, "c": 1, "G": 2, "g": 2, "T": 3, "t": 3}.get(c, 0))
    if in_hdr:
        want.append([hdr.split(" ")[0].split("\t")[0].rstrip("\r\n"), []])
    fa = L.bqc_fasta_open(str(path).encode())
    assert fa
    try:
        for name, codes in want:
            ptr = ctypes.c_void_p()
            n = L.bqc_fasta_contig(fa, name.encode(), ctypes.byref(ptr))
            assert n == len(codes), (name, n, len(codes))
            if n:
                got = np.frombuffer((ctypes.c_uint8 * ((n + 3) // 4)).from_address(ptr.value), dtype=np.uint8)
                exp = np.zeros((n + 3) // 4 * 4, dtype=np.uint8)
                exp[:n] = codes
                exp = exp.reshape(-1, 4)
                packed = exp[:, 0] | (exp[:, 1] << 2) | (exp[:, 2] << 4) | (exp[:, 3] << 6)
                assert np.array_equal(got, packed), name
    finally:
        L.bqc_fasta_close(fa)
