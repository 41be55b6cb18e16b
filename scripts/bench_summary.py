#!/usr/bin/env python
"""One-line summary of bench.py JSON lines:  python scripts/bench_summary.py gpurun_out/bench_n8.json [...]"""
import json, sys
for f in sys.argv[1:]:
    for l in open(f):
        if not l.startswith("{"):
            continue
        d = json.loads(l)
        x = d.get("extra", {})
        c, i = x.get("contains_seq") or {}, x.get("insert_seq") or {}
        e = d.get("e2e") or {}
        fl = (e.get("host_link_floor") or {}).get("ms_per_step")
        p = d.get("parity_check") or {}
        nv = (c.get("nvlink") or {}).get("egress_GBps")
        print(f"{f}: N={d.get('n_gpus')} {d['metric'][:12]} {d['value']:.4g} {d['unit']} {d['ms_per_step']:.3f} ms/step | e2e {e.get('value', 0):.4g} "
              f"({e.get('ms_per_step', 0):.1f} ms, link floor {fl if fl is None else round(fl, 1)} ms) | insert {i.get('value', 0):.4g} {i.get('ms_per_step', 0):.3f} ms "
              f"(e2e {(i.get('e2e') or {}).get('ms_per_step', 0):.1f} ms) | parity mismatches {p.get('mismatches')}/{p.get('set_mismatches')} over {p.get('set_words_checked')} words | "
              f"kernels { {k.split('<')[0]: round(v['ms'] / v['n'], 2) for k, v in (x.get('kernel_ms') or {}).items()} } | nvlink egress {nv if nv is None else round(nv)} GB/s | "
              f"roofline frac {(d.get('roofline') or {}).get('frac')}")
