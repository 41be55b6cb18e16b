"""The drop-in boundary from a compiled-language host: tests/host/cabi_driver.cpp includes include/cbl.hpp (the C++ mirror
of the reference's CBL<K, T, PREFIX_BITS>), links libcbl_gpu through the C ABI only (no torch, no Python in that
process) and checks insert_seq / contains_seq / | & - ^ and their assign forms / merge / intersect / iter / serde / error
messages on ONE GPU and on a handle sharded over several GPUs of one process (cbl_create_sharded), against expected
results written here from the CPU oracle.  The CPU half of this file only checks that the driver compiles and links."""
import os
import struct
import subprocess

import numpy as np
import pytest

import cbl_testutil as util

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "cbl_b200", "csrc")


def build_driver(out_dir) -> str:
    exe = os.path.join(str(out_dir), "cabi_driver")
    cmd = ["g++", "-std=c++17", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "host", "cabi_driver.cpp"),
           "-L", LIBDIR, "-lcbl_gpu", f"-Wl,-rpath,{LIBDIR}", "-o", exe]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    return exe


def test_cpp_facade_compiles_and_links(tmp_path):
    """include/cbl.hpp + the driver build against the shared library's exported C symbols (no GPU needed to link)."""
    exe = build_driver(tmp_path)
    assert os.path.exists(exe)


def blob(b: bytes) -> bytes:
    return struct.pack("<Q", len(b)) + b


@pytest.mark.gpu
def test_cpp_driver_single_and_sharded(tmp_path):
    from oracle.pyoracle import OracleCBL

    k, tb, pb = 25, 64, 24
    A = util.random_dna(60_000, seed=41)
    B = np.concatenate([A[20_000:45_000], util.random_dna(30_000, seed=42)])
    oa, ob = OracleCBL(k, tb, pb), OracleCBL(k, tb, pb)
    oa.insert_seq(A)
    ob.insert_seq(B)
    parts = [blob(A.tobytes()), blob(B.tobytes()), blob(oa.iter_words()[0].astype("<u8").tobytes())]
    for res in (oa | ob, oa & ob, oa - ob, oa ^ ob):
        parts.append(blob(res.iter_words()[0].astype("<u8").tobytes()))
    parts.append(blob(oa.contains_seq(B).astype(np.uint8).tobytes()))
    case = tmp_path / "case.bin"
    case.write_bytes(b"".join(parts))
    exe = build_driver(tmp_path)
    import torch

    devs = "0,1" if torch.cuda.device_count() >= 2 else "0,0"   # two shards: two GPUs when the box has them, else one device twice
    r = subprocess.run([exe, str(case), str(tmp_path), devs], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "all checks passed" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
