"""CPU check of the parallel form of the coverage statistic (bamqc_b200/csrc/cov_math.h + the block decomposition that
kernel_cov.cuh runs on the device: candidates, closed-form stretches, all-states block tables, virtual coordinates as
a prefix sum, carried windows) against a sequential restatement of src/OverallNumbers.hpp:59-135 on random record
streams: dense, sparse, exact window edges, ties, unsorted input, contig changes, arbitrary batch cuts."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_parallel_coverage_anchor_matches_sequential_reference(tmp_path):
    exe = str(tmp_path / "cov_selftest")
    subprocess.run(["g++", "-O2", "-std=c++17", "-o", exe, os.path.join(ROOT, "tests", "cov_selftest.cpp")], check=True)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "cov_selftest ok" in r.stdout
