"""CPU checks of the arithmetic inside k_inflate's match copy (bamqc_b200/csrc/kernel_inflate.cuh): the reciprocal
table that replaces `lane % dist`, and the lane schedule of the three copy branches against the byte-by-byte
semantics of an LZ77 match (RFC 1951 3.2.3: the copy may overlap its own output).  The constants are read out of the
kernel source, the schedule is restated here line by line; the kernel itself is covered by the -m gpu tests
(zlib levels 0/1/6/9, periodic fields with distances 1..63)."""
import os
import random
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = open(os.path.join(ROOT, "bamqc_b200", "csrc", "kernel_inflate.cuh")).read()


def _table(name):
    m = re.search(r"__constant__\s+\w+\s+" + name + r"\[\d+\]\s*=\s*\{([^}]*)\}", SRC)
    assert m, name
    return [int(x) for x in m.group(1).replace("\n", " ").split(",")]


def test_reciprocal_table_is_an_exact_division_for_a_warp():
    rcp = _table("c_rcp")
    assert len(rcp) == 32
    for d in range(2, 32):
        assert rcp[d] == 65536 // d + 1
        for x in range(0, 33):   # lane indices and the step 32
            assert (x * rcp[d]) >> 16 == x // d, (x, d)
    assert all(v < 65536 for v in rcp)


def test_deflate_base_tables_match_rfc1951():
    len_base, len_extra = _table("c_len_base"), _table("c_len_extra")
    dist_base, dist_extra = _table("c_dist_base"), _table("c_dist_extra")
    # RFC 1951 3.2.5: every length 3..258 and distance 1..32768 is reachable by exactly one (base, extra) pair
    covered = {}
    for b, e in zip(len_base[:-1], len_extra[:-1]):
        for x in range(1 << e):
            covered[b + x] = covered.get(b + x, 0) + 1
    covered[258] = covered.get(258, 0) + 1   # symbol 285
    assert sorted(k for k in covered if 3 <= k <= 258) == list(range(3, 259))
    assert covered[258] == 2 and all(v == 1 for k, v in covered.items() if k < 258)   # 258 is also 227 + 31 (symbol 284)
    seen = []
    for b, e in zip(dist_base, dist_extra):
        seen += [b + x for x in range(1 << e)]
    assert seen == list(range(1, 32769))
    assert _table("c_cl_order") == [16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15]


def _warp_copy(out, pos, dist, length, rcp):
    """The kernel's copy of one match, all 32 lanes, in its own order: loads of a step see the output as it was before
    the match (every source index is taken modulo the distance into bytes that existed then)."""
    o = out
    sp = pos - dist
    pend = {}
    if length <= 32 and dist >= length:
        for lane in range(32):
            if lane < length:
                pend[pos + lane] = o[sp + lane]
    else:
        rc = rcp[dist] if dist < 32 else 0
        for lane in range(32):
            m = 0 if dist == 1 else lane - dist * ((lane * rc) >> 16)
            if length <= 32:
                if lane < length:
                    pend[pos + lane] = o[sp + m]
                continue
            j = lane
            if dist >= length:
                while j + 32 < length + lane:
                    pend[pos + j] = o[sp + j]
                    j += 32
                if j < length:
                    pend[pos + j] = o[sp + j]
            elif dist < 32:
                step = 0 if dist == 1 else 32 - dist * ((32 * rc) >> 16)
                while j + 32 < length + lane:
                    pend[pos + j] = o[sp + m]
                    m += step
                    if m >= dist:
                        m -= dist
                    j += 32
                if j < length:
                    pend[pos + j] = o[sp + m]
            else:
                while j + 32 < length + lane:
                    pend[pos + j] = o[sp + j % dist]
                    j += 32
                if j < length:
                    pend[pos + j] = o[sp + j % dist]
    assert sorted(pend) == list(range(pos, pos + length))   # every byte of the match is written exactly once
    for k, v in pend.items():
        o[k] = v


def test_match_copy_schedule_equals_the_sequential_copy():
    rcp = _table("c_rcp")
    rng = random.Random(11)
    prefix = [rng.randrange(256) for _ in range(400)]
    dists = list(range(1, 72)) + [100, 257, 258, 259, 300, 400]
    for dist in dists:
        for length in list(range(3, 70)) + [96, 97, 128, 129, 255, 256, 257, 258]:
            a = prefix + [0] * length
            b = list(a)
            pos = len(prefix)
            for i in range(length):      # RFC 1951: byte by byte, may read what it just wrote
                a[pos + i] = a[pos - dist + i]
            _warp_copy(b, pos, dist, length, rcp)
            assert a == b, (dist, length)


def test_bit_reader_matches_a_plain_lsb_first_reader(tmp_path):
    """inflate_bits.h (BitWin) is host/device code: the same struct k_inflate uses, fuzzed on the CPU."""
    exe = str(tmp_path / "inflate_selftest")
    subprocess.run(["g++", "-O2", "-std=c++17", "-o", exe, os.path.join(ROOT, "tests", "inflate_selftest.cpp")], check=True)
    r = subprocess.run([exe, "3000"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert " 0 mismatching" in r.stdout
