#!/usr/bin/env python
"""Multi-GPU parity check (run under torchrun, one rank per GPU):
  python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/multigpu_check.py
Every rank counts its own slice of reads through the hash-prefix-sharded pipeline
(extract+partition -> NCCL all-to-all -> tiled insert), then rank 0 checks, against
the CPU oracle on the union of all reads: the union of the shard dumps, the
all-gathered mutant k-mer list and every rank's keep mask. Bit-exact."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def main():
    from rufus_b200 import api, kmer, pipeline, synth
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    k, L, n_pairs, G = 25, 150, 40_000, 300_000
    trio, _ = synth.make_trio(17, G, n_pairs, slice_id=rank)
    if os.environ.get("RG_HEAVY", "1") == "1":
        # heavy hitters: poly-A / poly-AC reads make single groups overflow their slabs, so the owner's
        # exported spill list (p2p) and the bucket overflow paths are part of what is checked
        for s in synth.SAMPLES:
            trio[s][0][:3000] = ord("A")
            trio[s][2][:1500, 0::2] = ord("A")
            trio[s][2][:1500, 1::2] = ord("C")
    ctx = api.Context(local)
    ctx.use_stream(torch.cuda.current_stream().cuda_stream)
    streams = [torch.from_numpy(np.concatenate([trio[s][0], trio[s][2]]).reshape(-1)).to(dev) for s in synth.SAMPLES]
    subj = tuple(torch.from_numpy(a.reshape(-1)).to(dev) for a in trio["Child"])
    est = int(1.1 * G / world + 14 * n_pairs)
    exchange = os.environ.get("RG_EXCHANGE", "p2p")
    pipe = pipeline.DevicePipeline(ctx, k, streams, subj, L, est, 5, 40 * world, 15, 1, rank=rank, world=world,
                                   exchange=exchange, filter_exchange=os.environ.get("RG_FILTER_EXCHANGE", "set"))
    r = pipe.step()
    r = pipe.step()          # twice: tables are cleared and re-used
    dumps = [t.dump(min_count=1) for t in pipe.tables]
    keep = pipe.keep.cpu().numpy()
    mut = pipe.mut_keys[: 0].cpu().numpy()
    # gather everything on rank 0 through files in a shared tmp dir (same node)
    tmp = os.environ.get("RG_TMP", "/tmp/rg_multi")
    os.makedirs(tmp, exist_ok=True)
    np.savez(os.path.join(tmp, f"rank{rank}.npz"), keep=keep, n_mutant=r["n_mutant"], histo=r["histo"],
             **{f"k{i}": d[0] for i, d in enumerate(dumps)}, **{f"c{i}": d[1] for i, d in enumerate(dumps)},
             **{f"reads_{s}_{j}": trio[s][j] for s in synth.SAMPLES for j in range(4)})
    dist.barrier()
    ok = True
    if rank == 0:
        import oracle
        parts = [np.load(os.path.join(tmp, f"rank{q}.npz")) for q in range(world)]
        otabs = []
        for i, s in enumerate(synth.SAMPLES):
            reads = np.concatenate([np.concatenate([p[f"reads_{s}_0"], p[f"reads_{s}_2"]]) for p in parts])
            okk, occ = oracle.count(reads.reshape(-1), np.arange(reads.shape[0] + 1, dtype=np.uint64) * L, k, True, 1)
            gk = np.concatenate([p[f"k{i}"] for p in parts])
            gc = np.concatenate([p[f"c{i}"] for p in parts])
            o = np.argsort(gk)
            same = np.array_equal(gk[o], okk) and np.array_equal(gc[o], occ)
            sizes = [int(p[f"k{i}"].size) for p in parts]
            print(f"sample {s}: {okk.size} distinct, shard sizes {sizes}, union == oracle: {same}")
            ok &= same
            m2 = occ >= 2
            otabs.append((okk[m2], occ[m2]))
            if i == 0:
                h, _ = oracle.histo(occ[m2])
                same_h = np.array_equal(h, parts[0]["histo"])
                print("subject histogram (all-reduced) == oracle:", same_h)
                ok &= same_h
        hk, hc = oracle.hashlist(otabs[0], otabs[1:], 5, 40 * world)
        print(f"mutant k-mers: oracle {hk.size}, gpu {[int(p['n_mutant']) for p in parts]}")
        ok &= all(int(p["n_mutant"]) == hk.size for p in parts)
        rset = oracle.rufus_set_from_kmers(kmer.keys_to_text(hk, k))
        for q, p in enumerate(parts):
            s1, q1, s2, q2 = (p[f"reads_Child_{j}"] for j in range(4))
            offs = np.arange(s1.shape[0] + 1, dtype=np.uint64) * L
            okeep = oracle.filter_pairs(s1, q1, offs, s2, q2, offs, k, 15, 1, rset)
            same = np.array_equal(okeep, p["keep"])
            print(f"rank {q}: pulled {int(p['keep'].sum())} pairs, == oracle: {same}")
            ok &= same
        print("MULTI-GPU PARITY", "OK" if ok else "FAILED", f"(world={world})")
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.broadcast(flag, 0)
    dist.destroy_process_group()
    ctx.close()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
