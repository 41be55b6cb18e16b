"""The `cbl` command line tool (cbl_b200/cli.py mirroring examples/cbl.rs): the FASTA/Q reader on the CPU, the eleven
sub-commands end to end on the GPU against the oracle (counts, listed k-mers, query statistics, set operations on files the
ORACLE reads back)."""
import gzip
import os
import subprocess
import sys

import numpy as np
import pytest

import cbl_testutil as util

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def write_fasta(path, recs, width=80, gz=False):
    lines = []
    for i, r in enumerate(recs):
        lines.append(b">r%d some description" % i)
        for j in range(0, len(r), width):
            lines.append(r[j : j + width])
    data = b"\n".join(lines) + b"\n"
    with (gzip.open if gz else open)(path, "wb") as f:
        f.write(data)


def write_fastq(path, recs):
    with open(path, "wb") as f:
        for i, r in enumerate(recs):
            f.write(b"@q%d\n" % i + r + b"\n+\n" + b"I" * len(r) + b"\n")


def test_fastx_reader(tmp_path):
    sys.path.insert(0, ROOT)
    from cbl_b200 import cli

    recs = [util.random_dna(n, seed=n).tobytes() for n in (1, 79, 80, 81, 1000, 25)]
    fa, fagz, fq = str(tmp_path / "a.fa"), str(tmp_path / "a.fa.gz"), str(tmp_path / "a.fq")
    write_fasta(fa, recs)
    write_fasta(fagz, recs, width=60, gz=True)
    write_fastq(fq, recs)
    for p in (fa, fagz, fq):
        assert list(cli.read_fastx(p)) == recs, p
    buf, off = next(cli.batches(fa))
    assert buf.tobytes() == b"".join(recs) and off.tolist() == np.concatenate([[0], np.cumsum([len(r) for r in recs])]).tolist()
    assert cli.t_bits_for(13) == 32 and cli.t_bits_for(15) == 64 and cli.t_bits_for(25) == 64 and cli.t_bits_for(29) == 64 and cli.t_bits_for(31) == 128 and cli.t_bits_for(59) == 128
    assert cli.kmer_to_nucs(0b00_01_10_11, 4) == b"ACTG"
    with pytest.raises(SystemExit, match="Failed to open"):
        list(cli.read_fastx(str(tmp_path / "missing.fa")))


def run_cli(*args, env=None):
    e = dict(os.environ, PYTHONPATH=ROOT, **(env or {}))
    r = subprocess.run([sys.executable, "-m", "cbl_b200.cli", *args], capture_output=True, text=True, env=e, timeout=600)
    return r


@pytest.mark.gpu
def test_cli_end_to_end(tmp_path):
    from oracle.pyoracle import OracleCBL

    k = 25
    A = [util.random_dna(n, seed=500 + i).tobytes() for i, n in enumerate((30000, 2500, 41000))]
    B = [A[0][5000:20000], util.random_dna(20000, seed=600).tobytes()]
    fa, fb = str(tmp_path / "a.fa"), str(tmp_path / "b.fq")
    write_fasta(fa, A)
    write_fastq(fb, B)
    ia, ib, io = str(tmp_path / "a.cbl"), str(tmp_path / "b.cbl"), str(tmp_path / "o.cbl")
    oa, ob = OracleCBL(k, 64, 24, True), OracleCBL(k, 64, 24, True)
    for r in A:
        oa.insert_seq(r)
    for r in B:
        ob.insert_seq(r)
    r = run_cli("build", fa, "-o", ia, "-c")
    assert r.returncode == 0 and f"Building the index of canonical {k}-mers contained in {fa}" in r.stderr and f"Writing the index to {ia}" in r.stderr, r.stderr
    assert run_cli("build", fb, "-o", ib, "--canonical").returncode == 0
    W = lambda o: util.to_int_list(*o.iter_words())
    assert W(oa.deserialize(open(ia, "rb").read())) == W(oa), "the oracle cannot read the file `cbl build` wrote"
    r = run_cli("count", ia)
    assert f"Reading the index stored in {ia}" in r.stderr and f"It contains {oa.count()} canonical {k}-mers" in r.stderr
    r = run_cli("list", ia)
    listed = r.stdout.split()
    assert len(listed) == oa.count() and listed == [("".join("ACTG"[(oa.recover_kmer(w) >> (2 * (k - 1 - i))) & 3] for i in range(k))) for w in W(oa)]
    r = run_cli("query", ia, fb)
    exp = np.concatenate([oa.contains_seq(x) for x in B])
    assert f"# queries: {len(exp)}" in r.stderr and f"# positive queries: {int(exp.sum())} ({int(exp.sum()) * 100 / len(exp):.2f}%)" in r.stderr, r.stderr
    for cmd, op in (("merge", "__or__"), ("inter", "__and__"), ("diff", "__sub__"), ("sym-diff", "__xor__")):
        assert run_cli(cmd, ia, ib, "-o", io).returncode == 0
        assert W(oa.deserialize(open(io, "rb").read())) == W(getattr(oa, op)(ob)), cmd
    assert run_cli("insert", ia, fb, "-o", io).returncode == 0
    assert W(oa.deserialize(open(io, "rb").read())) == W(oa | ob)
    assert run_cli("remove", ia, fb, "-o", io).returncode == 0
    assert W(oa.deserialize(open(io, "rb").read())) == W(oa - ob)
    r = run_cli("repartition", ia)
    assert "of the available prefixes are used" in r.stderr and "The biggest bucket (of size" in r.stderr and "nodes in total" in r.stderr, r.stderr
    r = run_cli("count", str(tmp_path / "nope.cbl"))
    assert r.returncode != 0 and "Failed to open" in r.stderr
    # sharded over two "GPUs" of the process (the same device twice on a one-GPU box is not possible through --gpus: skip unless 2 GPUs)
    import torch

    if torch.cuda.device_count() >= 2:
        assert run_cli("--gpus", "2", "build", fa, "-o", io, "-c").returncode == 0
        assert W(oa.deserialize(open(io, "rb").read())) == W(oa)
