#!/usr/bin/env python
"""Sliced-parity digests of the large benchmark configs (rufus_b200/workload.py BIG: BASELINE configs[2..4] scaled).

The reads are those bench.py generates on the GPU (counter-based generator, rg_synth_pairs_host writes the same bytes
on the CPU). For every sample the CPU restatement (oracle/rufus_oracle.c ro_slice_*: the k-mer extraction of
mer_iterator.hpp:59-89) counts the slice {key : slice_hash(key) % 64 == 0} over ALL reads of all ranks, keeps
count >= 2 (the -L 2 of RunJellyForRUFUS.sh:29), and the HashList of the slice is the reference's set algebra
(oracle.hashlist = scripts/CheckJellyHashList.sh:12) on those slice tables, for every min-count of the config's sweep.
Stored: md5 of the sorted (key, count) arrays per sample, md5 of each slice HashList.

Usage: make_big_digests.py CONFIG WORLD [CONFIG WORLD ...]    (CPU only; minutes per config, one process per sample)
"""
import json, os, sys, time
from concurrent.futures import ProcessPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def count_sample(args):
    name, world, si = args
    import numpy as np
    import oracle
    from rufus_b200 import api, workload
    cfg = workload.BIG[name]
    sc = oracle.SliceCounter(cfg["k"], True, workload.SLICE_MOD, 0)
    chunk = 250_000
    for rank in range(world):
        for first in range(0, cfg["pairs"][si], chunk):
            n = min(chunk, cfg["pairs"][si] - first)
            s1, _, s2, _ = api.synth_pairs_host(workload.big_spec(cfg, si, rank, world, first, n))
            sc.add(s1)
            sc.add(s2)
    keys, counts = sc.finish(2)
    return si, keys, counts


def run(name, world):
    import numpy as np
    import oracle
    from rufus_b200 import workload
    cfg = workload.BIG[name]
    t0 = time.time()
    ns = len(cfg["samples"])
    with ProcessPoolExecutor(max_workers=min(ns, os.cpu_count() or 1)) as ex:
        tabs = {si: (k, c) for si, k, c in ex.map(count_sample, [(name, world, si) for si in range(ns)])}
    res = {"config": name, "world": world, "slice_mod": workload.SLICE_MOD, "k": cfg["k"], "seed": cfg["seed"],
           "genome": cfg["genome_unit"] * world, "pairs_per_rank": list(cfg["pairs"]),
           "samples": [workload.table_digest(*tabs[si]) for si in range(ns)], "hashlist": {}}
    for mc in cfg["min_cov"]:
        hk, hc = oracle.hashlist(tabs[0], [tabs[i] for i in range(1, ns)], max(mc, 5), cfg["max_cov"])
        res["hashlist"][str(mc)] = workload.table_digest(hk, hc)
    res["cpu_seconds"] = round(time.time() - t0, 1)
    return res


if __name__ == "__main__":
    from rufus_b200 import workload
    db = json.load(open(workload.BIG_DIGESTS)) if os.path.exists(workload.BIG_DIGESTS) else {}
    a = sys.argv[1:]
    for name, world in zip(a[0::2], a[1::2]):
        r = run(name, int(world))
        db[workload.big_key(name, int(world))] = r
        print(name, world, {k: v for k, v in r.items() if k != "samples"}, [s["n"] for s in r["samples"]], flush=True)
        json.dump(db, open(workload.BIG_DIGESTS, "w"), indent=1, sort_keys=True)
