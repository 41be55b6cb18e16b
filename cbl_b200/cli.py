"""`cbl` — the reference's command line tool (examples/cbl.rs:29-53, 144-367) over the B200 library: the same 11
sub-commands, arguments and stderr messages.

    python -m cbl_b200.cli build reads.fa -o index.cbl [-c]
    python -m cbl_b200.cli count | list | query | insert | remove | merge | inter | diff | sym-diff | repartition ...

K, PREFIX_BITS (and T) are compile-time constants of the reference binary, taken from the environment by its build.rs
(build.rs:9-57: K default 25, PREFIX_BITS default 24, T = the smallest of u32 / u64 / u128 that holds 2K + lg(2K) bits);
here they are read from the same environment variables (or -k / -p) at run time.  The record loops of the reference
(`while let Some(record) = reader.next() { cbl.insert_seq(&record.seq()) }`) become one library call per ~256 MB of
records.  `--gpus N` shards the set over N GPUs of this process (cbl_create_sharded).
"""
from __future__ import annotations

import argparse
import gzip
import os
import sys
from typing import Iterator, List, Tuple

import numpy as np


# ---- FASTA / FASTQ reader (needletail's parse_fastx_file: multi-line FASTA, 4-line FASTQ, optionally gzipped) --------------
def open_maybe_gz(path: str):
    try:
        f = open(path, "rb")
    except OSError:
        raise SystemExit(f"Failed to open {path}")     # examples/cbl.rs:113-116
    magic = f.read(2)
    f.seek(0)
    if magic == b"\x1f\x8b":
        return gzip.open(f, "rb")
    return f


def read_fastx(path: str) -> Iterator[bytes]:
    """the sequences of a FASTA / FASTQ file, line breaks removed (what `record.seq()` returns)"""
    with open_maybe_gz(path) as f:
        first = f.read(1)
        if not first:
            return
        if first == b">":
            parts: List[bytes] = []
            started = False
            for line in f:                     # the rest of the first header line comes first
                if not started:
                    started = True
                    continue
                if line.startswith(b">"):
                    yield b"".join(parts)
                    parts = []
                else:
                    parts.append(line.strip())
            yield b"".join(parts)
        elif first == b"@":
            f.readline()
            while True:
                seq = f.readline()
                if not seq:
                    break
                f.readline()                   # '+'
                f.readline()                   # qualities
                yield seq.strip()
                if not f.readline():           # next '@' header
                    break
        else:
            raise SystemExit("Invalid record")


def batches(path: str, limit: int = 256 << 20) -> Iterator[Tuple[np.ndarray, np.ndarray]]:
    recs, size = [], 0
    for s in read_fastx(path):
        recs.append(s)
        size += len(s)
        if size >= limit:
            yield pack(recs)
            recs, size = [], 0
    if recs:
        yield pack(recs)


def pack(recs: List[bytes]) -> Tuple[np.ndarray, np.ndarray]:
    off = np.zeros(len(recs) + 1, dtype=np.uint64)
    off[1:] = np.cumsum([len(r) for r in recs])
    return np.frombuffer(b"".join(recs), dtype=np.uint8), off


def t_bits_for(k: int) -> int:
    need = 2 * k + (2 * k - 1).bit_length()
    return 32 if need <= 32 else 64 if need <= 64 else 128


NUC = b"ACTG"    # src/kmer.rs:11


def kmer_to_nucs(x: int, k: int) -> bytes:
    return bytes(NUC[(x >> (2 * (k - 1 - i))) & 3] for i in range(k))


def main(argv=None) -> int:
    K = int(os.environ.get("K", 25))
    P = int(os.environ.get("PREFIX_BITS", 24))
    ap = argparse.ArgumentParser(prog="cbl", description=f"CBL compiled for K={K}")
    ap.add_argument("-k", type=int, default=K, help="k-mer size (the reference fixes it at compile time through the K environment variable)")
    ap.add_argument("-p", "--prefix-bits", type=int, default=P)
    ap.add_argument("--gpus", type=int, default=1, help="shard the set over this many GPUs of the process")
    sub = ap.add_subparsers(dest="command", required=True)

    def idx(p):
        p.add_argument("index", help="Index file (CBL format)")

    b = sub.add_parser("build", help="Build an index containing the k-mers of a FASTA/Q file")
    b.add_argument("input", help="Input file (FASTA/Q, possibly gzipped)")
    b.add_argument("-o", "--output", help="Output file (no serialization by default)")
    b.add_argument("-c", "--canonical", action="store_true", help="Use canonical k-mers")
    idx(sub.add_parser("count", help="Count the k-mers contained in an index"))
    ls = sub.add_parser("list", help="List the k-mers contained in an index")
    idx(ls)
    ls.add_argument("-o", "--output", help="Output file (write to stdout by default)")
    q = sub.add_parser("query", help="Query an index for every k-mer contained in a FASTA/Q file")
    idx(q)
    q.add_argument("input", help="Input file to query (FASTA/Q, possibly gzipped)")
    for name, hlp in (("insert", "Add the k-mers of a FASTA/Q file to an index"), ("remove", "Remove the k-mers of a FASTA/Q file from an index")):
        u = sub.add_parser(name, help=hlp)
        idx(u)
        u.add_argument("input", help="Input file to query (FASTA/Q, possibly gzipped)")
        u.add_argument("-o", "--output", help="Output file (no serialization by default)")
    for name, hlp in (("merge", "Compute the union of two indexes"), ("inter", "Compute the intersection of two indexes"),
                      ("diff", "Compute the difference of two indexes"), ("sym-diff", "Compute the symmetric difference of two indexes")):
        s = sub.add_parser(name, help=hlp)
        s.add_argument("first_index", help="Index file (CBL format)")
        s.add_argument("second_index", help="Index file (CBL format)")
        s.add_argument("-o", "--output", help="Output file (no serialization by default)")
    idx(sub.add_parser("repartition", help="Show the repartition of the k-mers in the data structure"))
    args = ap.parse_args(argv)
    K, P = args.k, args.prefix_bits

    from .cbl import CBL

    def new(canonical=False):
        if args.gpus > 1:
            return CBL.sharded(K, t_bits_for(K), P, canonical, list(range(args.gpus)))
        return CBL(K, t_bits_for(K), P, canonical)

    def read_index(path):
        if not os.path.exists(path):
            raise SystemExit(f"Failed to open {path}")
        print(f"Reading the index stored in {path}", file=sys.stderr)
        return new().load_from_file(path)

    def write_index(c, path):
        print(f"Writing the index to {path}", file=sys.stderr)
        c.save_to_file(path)

    canon = lambda c: "canonical " if c.is_canonical() else ""
    cmd = args.command
    if cmd == "build":
        c = new(args.canonical)
        print(f"Building the index of {canon(c)}{K}-mers contained in {args.input}", file=sys.stderr)
        for buf, off in batches(args.input):
            c.insert_seqs(buf, off)
        if args.output:
            write_index(c, args.output)
    elif cmd == "count":
        c = read_index(args.index)
        print(f"It contains {c.count()} {canon(c)}{K}-mers", file=sys.stderr)
    elif cmd == "list":
        c = read_index(args.index)
        print(f"Listing {canon(c)}{K}-mers contained in {args.index}", file=sys.stderr)
        out = open(args.output, "wb") if args.output else sys.stdout.buffer
        for x in c.iter():
            out.write(kmer_to_nucs(x, K) + b"\n")
        if args.output:
            out.close()
    elif cmd == "query":
        c = read_index(args.index)
        print(f"Querying the {canon(c)}{K}-mers contained in {args.input}", file=sys.stderr)
        total = positive = 0
        for buf, off in batches(args.input):
            ans = c.contains_seqs(buf, off)
            total += len(ans)
            positive += int(ans.sum())
        print(f"# queries: {total}", file=sys.stderr)
        print(f"# positive queries: {positive} ({positive * 100 / max(total, 1):.2f}%)", file=sys.stderr)
    elif cmd in ("insert", "remove"):
        c = read_index(args.index)
        if cmd == "insert":
            print(f"Adding the {canon(c)}{K}-mers contained in {args.input} to the index", file=sys.stderr)
        else:
            print(f"Removing the {canon(c)}{K}-mers contained in {args.input} from the index", file=sys.stderr)
        for buf, off in batches(args.input):
            (c.insert_seqs if cmd == "insert" else c.remove_seqs)(buf, off)
        if args.output:
            write_index(c, args.output)
    elif cmd in ("merge", "inter", "diff", "sym-diff"):
        c, c2 = read_index(args.first_index), read_index(args.second_index)
        if cmd == "merge":
            c |= c2
        elif cmd == "inter":
            c &= c2
        elif cmd == "diff":
            c -= c2
        else:
            c ^= c2
        if args.output:
            write_index(c, args.output)
    elif cmd == "repartition":
        c = read_index(args.index)
        print(f"{c.prefix_load() * 100.0:.1f}% of the available prefixes are used", file=sys.stderr)
        sc = c.buckets_size_count()
        total_buckets = sum(sc.values())
        total_items = sum(s * n for s, n in sc.items())
        print(f"The average bucket size is {total_items / total_buckets:.1f} items", file=sys.stderr)
        bucket_count = item_count = 0
        for size in sorted(sc):
            count = sc[size]
            bucket_count += count
            item_count += size * count
            if count > total_buckets // 100 // 2 or size * count > total_items // 100 // 2 or bucket_count == total_buckets:
                print(f"{item_count * 100 / total_items:.1f}% of items are in a bucket of size ≤ {size} ({bucket_count * 100 / total_buckets:.1f}% of buckets)",
                      file=sys.stderr)
        p, s = c.buckets_sizes()
        i = int(np.argmax(s))
        print(f"The biggest bucket (of size {int(s[i])}) corresponds to prefix {int(p[i])}", file=sys.stderr)
        nc = c.buckets_node_count()
        vec_count = sum(n for nodes, n in nc.items() if nodes <= 1024)
        vec_nodes = sum(nodes * n for nodes, n in nc.items() if nodes <= 1024)
        trie_count = sum(n for nodes, n in nc.items() if nodes > 1024)
        trie_nodes = sum(nodes * n for nodes, n in nc.items() if nodes > 1024)
        nan = float("nan")
        print(f"{vec_count} vecs, average node count = {vec_nodes / vec_count if vec_count else nan:.1f}", file=sys.stderr)
        print(f"{trie_count} tries, average node count = {trie_nodes / trie_count if trie_count else nan:.1f}", file=sys.stderr)
        print(f"{total_buckets + vec_nodes + trie_nodes} nodes in total", file=sys.stderr)
    return 0


if __name__ == "__main__":
    try:
        sys.exit(main())
    except BrokenPipeError:
        sys.exit(0)
