"""Host orchestration of one pass of the hot path over a trio (or any subject +
controls): count every sample -> histogram -> subtract -> pull subject reads.

Mirrors the stage order of runRufus.sh:766-797 (count), :916-929 (subtract),
:946-985 (read pull). Two drivers share the same C-ABI calls:

  * DevicePipeline: inputs already resident in HBM (torch tensors are used only to
    own device memory and, for world > 1, for the NCCL all-to-all). This is what
    bench.py times as `value`.
  * host_step(): the public host-buffer API (rg_count_push / rg_subtract /
    rg_filter_pairs) — bench.py's `e2e`.

Multi-GPU: the k-mer space is partitioned by hash prefix; each rank
extracts k-mers from its own slice of reads, buckets them by owner rank
(rg_extract_partition_dev), ONE all-to-all routes the packed keys to their owners,
each owner inserts into its shard table. Subtraction is shard-local (same hash =>
same shard); the (small) mutant set is all-gathered so every rank filters its own
slice of subject reads without a second exchange of reads.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import api


def split_sizes_from_offsets(offsets) -> list[int]:
    """bucket offsets (n_parts+1) -> per-destination send counts for the all-to-all"""
    o = [int(x) for x in offsets]
    return [o[i + 1] - o[i] for i in range(len(o) - 1)]


def exchange_keys(keys, send_sizes, group=None, async_op=False):
    """One ragged all-to-all of packed k-mers (torch tensor, int64 view of u64 keys).

    keys: 1-D tensor laid out as contiguous per-destination buckets; send_sizes[r] keys
    go to rank r. Returns the tensor of keys this rank owns (async_op=True: (tensor, work);
    the caller waits on `work` before using the tensor, so that the transfer overlaps
    whatever it enqueues in between). Works on any backend (NCCL on GPUs, gloo on CPU for
    the host-logic tests)."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    send = torch.tensor(send_sizes, dtype=torch.int64, device=keys.device)
    recv = torch.empty(world, dtype=torch.int64, device=keys.device)
    dist.all_to_all_single(recv, send, group=group)
    recv_sizes = [int(x) for x in recv.tolist()]
    out = torch.empty(sum(recv_sizes), dtype=keys.dtype, device=keys.device)
    work = dist.all_to_all_single(out, keys[: sum(send_sizes)], output_split_sizes=recv_sizes,
                                  input_split_sizes=list(send_sizes), group=group, async_op=async_op)
    return (out, work) if async_op else out


def gather_varlen(local, group=None):
    """all-gather of a ragged 1-D tensor (the per-shard HashList pieces)."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    n = torch.tensor([local.numel()], dtype=torch.int64, device=local.device)
    ns = [torch.empty(1, dtype=torch.int64, device=local.device) for _ in range(world)]
    dist.all_gather(ns, n, group=group)
    sizes = [int(x.item()) for x in ns]
    mx = max(sizes + [1])
    pad = torch.zeros(mx, dtype=local.dtype, device=local.device)
    pad[: local.numel()] = local
    outs = [torch.empty(mx, dtype=local.dtype, device=local.device) for _ in range(world)]
    dist.all_gather(outs, pad, group=group)
    return torch.cat([o[:s] for o, s in zip(outs, sizes)])


class DevicePipeline:
    """count -> histo -> subtract -> filter with every input resident in HBM."""

    def __init__(self, ctx: api.Context, k: int, streams, subject_pairs, read_len: int, est_distinct: int,
                 min_cov: int = 5, max_cov: int = 150, minq: int = 15, thresh: int = 1, lower=(2, 2, 2),
                 mode: int = api.MODE_TILED, rank: int = 0, world: int = 1, max_mutant: int = 1 << 22,
                 exchange: str = "p2p", groups_per_shard: int = 0, scatter_ctas: int = 0, fast_cap: int = 1 << 16,
                 filter_exchange: str = "set"):
        import torch

        self.torch = torch
        self.ctx, self.k, self.read_len = ctx, k, read_len
        self.streams = streams                      # list of uint8 CUDA tensors [n_reads*read_len], subject first
        self.s1, self.q1, self.s2, self.q2 = subject_pairs
        self.n_pairs = self.s1.numel() // read_len
        self.min_cov, self.max_cov, self.minq, self.thresh, self.lower = min_cov, max_cov, minq, thresh, lower
        self.rank, self.world = rank, world
        dev = self.s1.device
        shard = (rank, world) if world > 1 else None
        self.tables = [api.KmerTable(ctx, k, True, est_distinct, mode, shard=shard) for _ in streams]
        self.mut_keys = torch.empty(max_mutant, dtype=torch.int64, device=dev)
        self.mut_counts = torch.empty(max_mutant, dtype=torch.int32, device=dev)
        self.hits1 = torch.empty(self.n_pairs, dtype=torch.int32, device=dev)
        self.hits2 = torch.empty(self.n_pairs, dtype=torch.int32, device=dev)
        self.keep = torch.empty(self.n_pairs, dtype=torch.uint8, device=dev)
        self.exchange = exchange if world > 1 else "none"
        import os
        self.serial = os.environ.get("RG_P2P_SERIAL", "0") == "1"
        self.schedule = os.environ.get("RG_P2P_SCHEDULE", "overlap")      # "batched" | "overlap"
        # "set" (default): the few mutant k-mers are all-gathered, every rank filters its own reads (SURVEY §8e layout 1);
        # "a2a": the set stays sharded, the reads travel and per-read hit counts are all-reduced (layout 2)
        self.filter_exchange = filter_exchange
        # Fast path: the whole pass is ENQUEUED (deferred status checks, histogram / subtraction count / kept
        # count left on the device, a set rebuilt in place, one fixed-capacity all-gather of the mutant k-mers)
        # and the host synchronises once, at the end. Falls back to the step-by-step path when a table reports
        # spilled keys or a shard finds more than fast_cap mutant k-mers.
        self.fast_cap = min(fast_cap, max_mutant)
        self.fast = k < 32 and os.environ.get("RG_PIPELINE_SLOW", "0") != "1" and filter_exchange == "set"
        if self.fast:
            L0 = api.lib()
            self.fast_set = L0.rg_set_create(ctx.h, k, self.fast_cap * world)
            if not self.fast_set:
                raise api._err("rg_set_create")
            # all-reduced in one go: histogram rows, mutant k-mer count, "some table / shard needs the slow tail"
            self.red = torch.zeros(10002 + 2, dtype=torch.int64, device=dev)
            self.count_dev = torch.zeros(1, dtype=torch.int64, device=dev)
            self.gathered = torch.empty(self.fast_cap * world, dtype=torch.int64, device=dev) if world > 1 else None
            self.res_dev = torch.zeros(2, dtype=torch.int64, device=dev)           # local mutant count, kept pairs
            self.status_dev = torch.zeros(len(streams), 9, dtype=torch.int64, device=dev)
            self.dirty_idx = torch.tensor([1, 2, 8], dtype=torch.int64, device=dev)
            self.red_host = torch.zeros(10002 + 2, dtype=torch.int64).pin_memory()
            self.res_host = torch.zeros(2, dtype=torch.int64).pin_memory()
        if self.exchange == "p2p":
            # peer-to-peer: every rank's partition pass writes straight into the owners' group slabs
            # (CUDA IPC mappings of one exported block per table); no collective in the data path
            import torch.distributed as dist
            L = api.lib()
            self.flag = torch.zeros(1, dtype=torch.int32, device=dev)
            self.flag2 = torch.zeros(1, dtype=torch.int32, device=dev)
            self.scatter_stream = torch.cuda.Stream(device=dev)
            # measured: N=2 96 CTAs (118.9 vs 113.4 Gbases/s without overlap); N=8 full grid 320 > 64: 306 >
            # 48: 280 ~ 96: 279 > 32: 203 (0 = one CTA per SM: the two streams interleave by themselves)
            self.scatter_ctas = scatter_ctas or (96 if world <= 2 else 0)
            n_keys = max(int(s.numel()) for s in streams)      # weak scaling: a shard receives ~ what it sends
            # fewer, longer runs travel better over NVLink (measured: N=2 37 groups/shard, N=8 16 > 37 > 74)
            groups_per_shard = groups_per_shard or (37 if world <= 2 else max(8, 128 // world))
            for t in self.tables:
                api._check(L.rg_table_p2p_prepare(t.h, groups_per_shard, int(n_keys * 1.15) + (1 << 20)),
                           "rg_table_p2p_prepare")
                h = np.zeros(64, np.uint8)
                api._check(L.rg_table_p2p_export(t.h, api._ptr(h)), "rg_table_p2p_export")
                mine = torch.from_numpy(h).to(dev)
                allh = [torch.empty(64, dtype=torch.uint8, device=dev) for _ in range(world)]
                dist.all_gather(allh, mine)
                for r in range(world):
                    hr = np.ascontiguousarray(allh[r].cpu().numpy())
                    api._check(L.rg_table_p2p_import(t.h, r, api._ptr(hr)), "rg_table_p2p_import")
            dist.barrier()
        elif world > 1:
            # one send buffer per sample: the all-to-all of sample i runs (NCCL's stream) while the
            # extraction of sample i+1 fills the next buffer
            self.part_keys = [torch.empty(int(s.numel()), dtype=torch.int64, device=dev) for s in streams]
            self.part_offs = torch.empty(world + 1, dtype=torch.int64, device=dev)
        self.last = {}

    def stream_barrier(self):
        """all ranks' kernels enqueued so far have completed before anything enqueued after it starts
        (a 4-byte all-reduce on the compute stream; the host does not wait)"""
        import torch.distributed as dist
        dist.all_reduce(self.flag)

    def extract_and_send(self, i: int):
        """world > 1: extract + bucket sample i by owner rank, start its all-to-all (asynchronous)."""
        L = api.lib()
        s = self.streams[i]
        rc = L.rg_extract_partition_dev(self.tables[i].h, api._ptr(s), s.numel(), self.read_len,
                                        api._ptr(self.part_keys[i]), api._ptr(self.part_offs))
        api._check(rc, "rg_extract_partition_dev")
        sizes = split_sizes_from_offsets(self.part_offs.tolist())
        self.last.setdefault("sent_keys", []).append(sum(sizes) - sizes[self.rank])
        return exchange_keys(self.part_keys[i], sizes, async_op=True)

    def step(self) -> dict:
        torch = self.torch
        L = api.lib()
        self.last = {}
        n = len(self.tables)
        for t in self.tables:
            t.clear()
        if self.world == 1:
            for t, s in zip(self.tables, self.streams):
                t.push_dev(s, s.numel(), self.read_len)   # enqueued only: no host synchronisation between the samples
        elif self.exchange == "p2p" and self.schedule == "batched":
            # every sample's scatter, ONE barrier, every sample's finish: two barriers per pass instead of four.
            # (The overlapped schedule below hid the NVLink-bound scatter of 8-byte keys behind the previous
            # sample's inserts; with 16-byte records per ~5 k-mers the scatter is no longer NVLink-bound.)
            for t in self.tables:
                api._check(L.rg_count_p2p_begin(t.h), "rg_count_p2p_begin")
            self.stream_barrier()
            for t, s in zip(self.tables, self.streams):
                api._check(L.rg_count_p2p_scatter_dev(t.h, api._ptr(s), s.numel(), self.read_len, 0),
                           "rg_count_p2p_scatter_dev")
            self.stream_barrier()
            for t in self.tables:
                api._check(L.rg_count_p2p_finish(t.h), "rg_count_p2p_finish")
        elif self.exchange == "p2p" and self.serial:
            # no overlap: scatter, barrier, finish, sample after sample on one stream (isolated kernel timings)
            import torch.distributed as dist
            for t in self.tables:
                api._check(L.rg_count_p2p_begin(t.h), "rg_count_p2p_begin")
            self.stream_barrier()
            for t, s in zip(self.tables, self.streams):
                api._check(L.rg_count_p2p_scatter_dev(t.h, api._ptr(s), s.numel(), self.read_len, 0),
                           "rg_count_p2p_scatter_dev")
                self.stream_barrier()
                api._check(L.rg_count_p2p_finish(t.h), "rg_count_p2p_finish")
        elif self.exchange == "p2p":
            import torch.distributed as dist
            main = torch.cuda.current_stream()
            side = self.scatter_stream
            for t in self.tables:
                api._check(L.rg_count_p2p_begin(t.h), "rg_count_p2p_begin")       # zero my slab cursors
            self.stream_barrier()                                                # every rank has zeroed
            side.wait_event(main.record_event())
            # The scatter of sample i+1 (NVLink-bound, a reduced grid) runs on a side stream while the
            # finish kernels of sample i (groups -> tiles -> inserts) take the remaining SMs.
            arrived = []
            self.ctx.switch_stream(side.cuda_stream)
            with torch.cuda.stream(side):
                for i, (t, s) in enumerate(zip(self.tables, self.streams)):
                    # the first scatter has nothing to overlap with: it takes every SM
                    ctas = 0 if i == 0 else self.scatter_ctas
                    api._check(L.rg_count_p2p_scatter_dev(t.h, api._ptr(s), s.numel(), self.read_len, ctas),
                               "rg_count_p2p_scatter_dev")
                    dist.all_reduce(self.flag2)                                  # every rank's scatter of this sample is done
                    arrived.append(side.record_event())
            self.ctx.switch_stream(main.cuda_stream)
            for t, ev in zip(self.tables, arrived):
                main.wait_event(ev)
                api._check(L.rg_count_p2p_finish(t.h), "rg_count_p2p_finish")
        else:
            # software pipeline: exchange of sample i overlaps extraction of sample i+1 and insertion of sample i-1
            inflight = self.extract_and_send(0)
            for i in range(n):
                nxt = self.extract_and_send(i + 1) if i + 1 < n else None
                mine, work = inflight
                work.wait()                                # the compute stream waits for the transfer, not the host
                self.tables[i].push_keys_dev(mine, mine.numel())
                inflight = nxt
        if self.fast and self._fast_tail():
            return self.last
        return self._slow_tail()

    def _fast_tail(self) -> bool:
        """histo -> subtract -> set -> read pull enqueued behind the counts, ONE host synchronisation. Returns
        False (nothing consumed) when the step-by-step tail has to run instead."""
        torch = self.torch
        L = api.lib()
        cap = self.fast_cap
        api._check(L.rg_defer_checks(self.ctx.h, 1), "rg_defer_checks")
        try:
            for i, t in enumerate(self.tables):
                api._check(L.rg_table_status_dev(t.h, api._ptr(self.status_dev[i])), "rg_table_status_dev")
            api._check(L.rg_histo_dev(self.tables[0].h, self.lower[0], 1, 10000, api._ptr(self.red)), "rg_histo_dev")
            others = self.tables[1:]
            om = np.ascontiguousarray(self.lower[1:], np.uint64)
            api._check(L.rg_subtract_enqueue(self.tables[0].h, self.lower[0], api._handles(others), api._ptr(om),
                                             len(others), 5, self.min_cov, self.max_cov, api._ptr(self.mut_keys),
                                             api._ptr(self.mut_counts), cap, api._ptr(self.count_dev)),
                       "rg_subtract_enqueue")
            self.red[10002:10003] = self.count_dev
            # foreign keys, k-mers without a slot, records spilled by peers, or more mutant k-mers than the fast buffers hold
            self.red[10003:10004] = self.status_dev.index_select(1, self.dirty_idx).sum() + (self.count_dev > cap).to(torch.int64)
            keys_all, n_max = self.mut_keys[:cap], cap
            if self.world > 1:
                import torch.distributed as dist
                dist.all_reduce(self.red)                                   # histogram + total mutant count
                dist.all_gather_into_tensor(self.gathered, self.mut_keys[:cap])
                keys_all, n_max = self.gathered, cap * self.world
            api._check(L.rg_set_rebuild_dev(self.fast_set, api._ptr(keys_all), n_max), "rg_set_rebuild_dev")
            rc = L.rg_filter_pairs_fixed_dev(self.fast_set, api._ptr(self.s1), api._ptr(self.q1), api._ptr(self.s2),
                                             api._ptr(self.q2), self.read_len, self.n_pairs, self.minq, self.thresh,
                                             api._ptr(self.hits1), api._ptr(self.hits2), api._ptr(self.keep))
            api._check(rc, "rg_filter_pairs_fixed_dev")
            self.res_dev[0:1] = self.count_dev
            self.res_dev[1:2] = self.keep.sum(dtype=torch.int64)
            self.red_host.copy_(self.red, non_blocking=True)
            self.res_host.copy_(self.res_dev, non_blocking=True)
            torch.cuda.current_stream().synchronize()                       # the pass's only host synchronisation
        finally:
            api._check(L.rg_defer_checks(self.ctx.h, 0), "rg_defer_checks")
        if int(self.red_host[10003]) != 0:      # (all-reduced: every rank takes the same path)
            return False
        self.n_local = int(self.res_host[0])
        self.last.update({"n_mutant": int(self.red_host[10002]), "n_kept": int(self.res_host[1]),
                          "histo": self.red_host[:10002].numpy().astype(np.uint64)})
        return True

    def rerun_tail(self, min_cov: int) -> dict:
        """histo -> subtract -> set -> read pull again over the counted tables with another MutantMinCov (the
        min-count sweep of BASELINE configs[4]: CheckJellyHashList.sh:12 re-cut at another cutoff)"""
        self.min_cov = min_cov
        self.last = {}
        if self.fast and self._fast_tail():
            return self.last
        return self._slow_tail()

    def _slow_tail(self) -> dict:
        torch = self.torch
        L = api.lib()
        for t in self.tables:
            t.end(trim=False)             # one read-back of the status words per table (spill / foreign keys)
        # histogram of the subject (jellyfish histo -f); all-reduced over shards
        h = self.tables[0].histo(min_count=self.lower[0])
        if self.world > 1:
            import torch.distributed as dist
            ht = torch.from_numpy(h.astype(np.int64)).to(self.s1.device)
            dist.all_reduce(ht)
            h = ht.cpu().numpy().astype(np.uint64)
        # shard-local subtraction -> mutant k-mers (device)
        others = self.tables[1:]
        om = np.ascontiguousarray(self.lower[1:], np.uint64)
        n = L.rg_subtract_dev(self.tables[0].h, self.lower[0], api._handles(others), api._ptr(om), len(others), 5,
                              self.min_cov, self.max_cov, api._ptr(self.mut_keys), api._ptr(self.mut_counts),
                              self.mut_keys.numel())
        if n < 0:
            raise api._err("rg_subtract_dev")
        self.n_local = int(n)
        keys = self.mut_keys[:n]
        if self.world > 1 and self.filter_exchange == "a2a":
            return self._filter_sharded_set(keys, h)
        if self.world > 1:
            keys = gather_varlen(keys)
        mset = api.MutantSet(self.ctx, L.rg_set_from_jf_keys_dev(self.ctx.h, api._ptr(keys), keys.numel(), self.k),
                             self.k)
        rc = L.rg_filter_pairs_fixed_dev(mset.h, api._ptr(self.s1), api._ptr(self.q1), api._ptr(self.s2),
                                         api._ptr(self.q2), self.read_len, self.n_pairs, self.minq, self.thresh,
                                         api._ptr(self.hits1), api._ptr(self.hits2), api._ptr(self.keep))
        api._check(rc, "rg_filter_pairs_fixed_dev")
        n_kept = int(self.keep.sum().item())     # device -> host read of the step's result
        mset.free()
        self.last.update({"n_mutant": int(keys.numel()), "n_kept": n_kept, "histo": h})
        return self.last

    def _filter_sharded_set(self, local_keys, h) -> dict:
        """SURVEY §8e layout (2), the north-star wording: the mutant set STAYS SHARDED and the reads travel. Every
        rank receives every rank's subject reads (one exchange of reads), counts their hits against ITS shard of the
        mutant set, and the per-read hit counts are summed over the ranks (each mutant k-mer lives in exactly one
        shard, so hits add up); a pair is pulled when either mate reaches the threshold (RUFUS.Filter.cpp:222).
        Strictly more traffic than layout (1) (all-gather of the few mutant k-mers): kept behind
        filter_exchange="a2a" for sets too large to replicate."""
        import torch.distributed as dist
        torch = self.torch
        L = api.lib()
        W, n = self.world, self.s1.numel()
        allr = [torch.empty(W * n, dtype=torch.uint8, device=self.s1.device) for _ in range(4)]
        for dst, src in zip(allr, (self.s1, self.q1, self.s2, self.q2)):
            dist.all_gather_into_tensor(dst, src)
        mset = api.MutantSet(self.ctx, L.rg_set_from_jf_keys_dev(self.ctx.h, api._ptr(local_keys), local_keys.numel(), self.k),
                             self.k)
        hits = [torch.zeros(W * self.n_pairs, dtype=torch.int32, device=self.s1.device) for _ in range(2)]
        for hh, (sq, ql) in zip(hits, ((allr[0], allr[1]), (allr[2], allr[3]))):
            api._check(L.rg_filter_hits_fixed_dev(mset.h, api._ptr(sq), api._ptr(ql), self.read_len, W * self.n_pairs,
                                                  self.minq, 1, api._ptr(hh)), "rg_filter_hits_fixed_dev")
            dist.all_reduce(hh)
        mine = slice(self.rank * self.n_pairs, (self.rank + 1) * self.n_pairs)
        self.keep.copy_(((hits[0][mine] >= self.thresh) | (hits[1][mine] >= self.thresh)).to(torch.uint8))
        self.hits1.copy_(hits[0][mine])
        self.hits2.copy_(hits[1][mine])
        nm = torch.tensor([local_keys.numel()], dtype=torch.int64, device=self.s1.device)
        dist.all_reduce(nm)
        n_kept = int(self.keep.sum().item())
        mset.free()
        self.last.update({"n_mutant": int(nm.item()), "n_kept": n_kept, "histo": h})
        return self.last

    def mutant_keys_host(self):
        n = self.last["n_mutant"] if self.world == 1 else None
        return self.mut_keys[:n].cpu().numpy().view(np.uint64) if n is not None else None

    def results(self) -> dict:
        """The last step's results on the host: subject histogram (all shards), this shard's piece of the
        HashList (keys, counts) and this rank's keep mask. What a caller reads back after a step."""
        n = self.n_local
        return {"histo": self.last["histo"], "keys": self.mut_keys[:n].cpu().numpy().view(np.uint64),
                "counts": self.mut_counts[:n].cpu().numpy().view(np.uint32), "keep": self.keep.cpu().numpy()}

    def gathered_results(self, sample_prefix: str = "C") -> dict:
        """results of the whole job on every rank (not timed by the bench): the HashList pieces of all
        shards concatenated, and the names of the pulled pairs of all ranks (workload.name_prefix naming)."""
        import torch.distributed as dist
        from . import workload
        r = self.results()
        names = workload.pulled_names(r["keep"], "Child", self.rank, self.world)
        if self.world == 1:
            return {"histo": r["histo"], "keys": r["keys"], "counts": r["counts"], "names": names}
        parts = [None] * self.world
        dist.all_gather_object(parts, (r["keys"], r["counts"], names))
        return {"histo": r["histo"], "keys": np.concatenate([p[0] for p in parts]),
                "counts": np.concatenate([p[1] for p in parts]), "names": [x for p in parts for x in p[2]]}

    def step_from_host(self, host_streams, host_pairs) -> dict:
        """One step with every input starting in (pinned) host memory and every result ending there: the
        end-to-end call of the sharded pipeline. host_streams: list of uint8 torch tensors (pinned), one per
        sample; host_pairs: the subject's (seq1, qual1, seq2, qual2) as pinned uint8 tensors."""
        for d, h in zip(self.streams, host_streams):
            d.copy_(h, non_blocking=True)
        for d, h in zip((self.s1, self.q1, self.s2, self.q2), host_pairs):
            d.copy_(h, non_blocking=True)
        self.step()
        return self.results()

    def free(self):
        if self.fast and self.fast_set:
            api.lib().rg_set_free(self.fast_set)
            self.fast_set = None
        for t in self.tables:
            t.free()


def host_step(ctx: api.Context, k: int, host_streams, subject_pairs, read_len: int, est_distinct: int,
              min_cov: int = 5, max_cov: int = 150, minq: int = 15, thresh: int = 1, lower=(2, 2, 2),
              mode: int = api.MODE_TILED, tables=None, timings: dict | None = None):
    """The same pass through the public host-buffer API: every input starts in host
    memory (pinned or not), results come back to host memory.

    host_streams: list of uint8 numpy arrays (subject first), reads separated by a
    non-ACGT byte. subject_pairs: (seq1, qual1, seq2, qual2) numpy [n_pairs, read_len]."""
    import time

    def lap(name, t0):
        if timings is not None:
            ctx.sync()
            timings[name] = timings.get(name, 0.0) + (time.perf_counter() - t0) * 1e3
        return time.perf_counter()

    own = tables is None
    if own:
        tables = [api.KmerTable(ctx, k, True, est_distinct, mode) for _ in host_streams]
    t0 = time.perf_counter()
    for t, s in zip(tables, host_streams):
        t.clear()
        t.push_stream(s)                  # asynchronous: H2D pieces overlap the kernels of the previous pieces
    for t in tables:
        t.end(trim=False)
    t0 = lap("count_ms", t0)
    h = tables[0].histo(min_count=lower[0])
    t0 = lap("histo_ms", t0)
    hk, hc = api.subtract(tables[0], tables[1:], subj_min=lower[0], other_min=list(lower[1:]), min_cov=min_cov,
                          max_cov=max_cov, sort=False)
    t0 = lap("subtract_ms", t0)
    mset = api.MutantSet.from_jf_keys(ctx, hk, k)
    s1, q1, s2, q2 = subject_pairs
    offs = np.arange(s1.shape[0] + 1, dtype=np.uint64) * np.uint64(read_len)
    t0 = lap("set_ms", t0)
    keep = mset.filter_pairs(s1, q1, offs, s2, q2, offs, minq, thresh)
    t0 = lap("filter_ms", t0)
    mset.free()
    if own:
        for t in tables:
            t.free()
    return {"histo": h, "hashlist": (hk, hc), "keep": keep}
