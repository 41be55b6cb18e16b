"""The __host__ __device__ k-mer / necklace functions the kernels run (cbl_b200/csrc/kmer_necklace.cuh)
are compiled for the CPU and checked against the oracle: fast necklace == brute force == reference
definition (random, periodic, sparse, dense words; 64- and 128-bit), revcomp, word packing."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_host_check(tmp_path):
    exe = str(tmp_path / "host_check")
    subprocess.run(["g++", "-std=c++17", "-O2", "-o", exe, os.path.join(ROOT, "tests", "host", "host_check.cpp")], check=True)
    r = subprocess.run([exe, "60000"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "0 mismatches" in r.stdout
