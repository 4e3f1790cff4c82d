"""world_size-2 gloo test of the multi-GPU host logic (no GPU needed): hash-prefix
ownership (rg_key_shard, the same function the kernels use), the ragged
all-to-all of packed k-mers (pipeline.exchange_keys) and the gather of the
per-shard HashList pieces (pipeline.gather_varlen). K-mers are extracted on the CPU
by the oracle (allowed in tests/), bucketed per owner like rg_extract_partition_dev
does, exchanged, counted per shard; the union of the shards must equal the oracle
count of all reads, with every key on the rank that owns it."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT

sys.path.insert(0, os.path.join(ROOT, "oracle"))


def _kmers_cpu(seq2d, k):
    """canonical k-mer occurrences of fixed-length reads (numpy, Jellyfish packing)"""
    code = np.full(256, 255, np.uint8)
    for i, c in enumerate(b"ACGT"):
        code[c] = i
        code[c + 32] = i
    c = code[seq2d]
    n, L = c.shape
    out = []
    mask = (1 << (2 * k)) - 1
    for s in range(L - k + 1):
        w = c[:, s:s + k]
        ok = (w != 255).all(axis=1)
        f = np.zeros(n, np.uint64)
        r = np.zeros(n, np.uint64)
        for j in range(k):
            f = (f << np.uint64(2)) | w[:, j].astype(np.uint64) & np.uint64(3)
            r = r | ((np.uint64(3) - (w[:, j].astype(np.uint64) & np.uint64(3))) << np.uint64(2 * j))
        out.append(np.minimum(f, r)[ok] & np.uint64(mask))
    return np.concatenate(out)


def _worker(rank, world, port, tmp):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    from rufus_b200 import api, pipeline, synth
    L = api.lib()
    k = 21
    trio, _ = synth.make_trio(3, 20000, 400, slice_id=rank)
    s1, _, s2, _ = trio["Child"]
    mine = _kmers_cpu(np.concatenate([s1, s2]), k)
    m, groups = L.rg_minimizer_len(k, 64), 64
    owner = np.array([L.rg_key_shard(int(x), k, m, groups, world) for x in mine.tolist()], dtype=np.int64)
    order = np.argsort(owner, kind="stable")
    keys = torch.from_numpy(mine[order].view(np.int64))
    offs = np.searchsorted(owner[order], np.arange(world + 1))
    sizes = pipeline.split_sizes_from_offsets(offs)
    got = pipeline.exchange_keys(keys, sizes).numpy().view(np.uint64)
    assert all(L.rg_key_shard(int(x), k, m, groups, world) == rank for x in got[:2000].tolist())
    uk, uc = np.unique(got, return_counts=True)
    np.save(os.path.join(tmp, f"k{rank}.npy"), uk)
    np.save(os.path.join(tmp, f"c{rank}.npy"), uc)
    np.save(os.path.join(tmp, f"reads{rank}.npy"), np.concatenate([s1, s2]))
    # ragged gather of "mutant" pieces: every rank must end with the same concatenation
    piece = torch.arange(rank * 100, rank * 100 + 3 + 2 * rank, dtype=torch.int64)
    allp = pipeline.gather_varlen(piece)
    exp = torch.cat([torch.arange(r * 100, r * 100 + 3 + 2 * r, dtype=torch.int64) for r in range(world)])
    assert torch.equal(allp, exp)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_exchange_matches_oracle(tmp_path):
    import socket
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    world = 2
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    import oracle
    reads = np.concatenate([np.load(tmp_path / f"reads{r}.npy") for r in range(world)])
    ok, oc = oracle.count(reads.reshape(-1), np.arange(reads.shape[0] + 1, dtype=np.uint64) * reads.shape[1], 21, True, 1)
    ks = [np.load(tmp_path / f"k{r}.npy") for r in range(world)]
    cs = [np.load(tmp_path / f"c{r}.npy") for r in range(world)]
    assert np.intersect1d(ks[0], ks[1]).size == 0          # shards are disjoint
    allk = np.concatenate(ks)
    allc = np.concatenate(cs)
    o = np.argsort(allk)
    assert np.array_equal(allk[o], ok) and np.array_equal(allc[o].astype(np.uint32), oc)
    assert min(k.size for k in ks) > 0.4 * allk.size / world   # balanced


def test_split_sizes():
    from rufus_b200 import pipeline
    assert pipeline.split_sizes_from_offsets([0, 3, 3, 10]) == [3, 0, 7]
