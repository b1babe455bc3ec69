"""CPU check of the word-parallel forms in bamqc_b200/csrc/swar.h (the statements k_stats runs per 8/16 bases) against
per-base restatements of src/TripletCounting.hpp:195-236 and src/QualityCheck.hpp:111-176 on random reads."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_swar_forms_match_per_base_walk(tmp_path):
    exe = str(tmp_path / "swar_selftest")
    subprocess.run(["g++", "-O2", "-o", exe, os.path.join(ROOT, "tests", "swar_selftest.cpp")], check=True)
    r = subprocess.run([exe, "200000"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "0 mismatching" in r.stdout
