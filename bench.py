#!/usr/bin/env python
"""bench.py — Gbases/s through k-mer count + subtract + read pull (k=25, synthetic
30x trio of 150 bp reads), BASELINE.json's metric.

  python bench.py [--gpus N] [--steps K] [--warmup W]        our CUDA path
  python bench.py --impl reference [...]                     the reference's CPU path

One step = one pass of the hot path over one batch: count the three samples
(1 M reads x 150 bp each, per GPU) -> subject histogram -> subtract -> pull the
subject's read pairs. `value` = bases of all samples counted / device time with
inputs resident in HBM; `e2e` = the same pass with inputs in pinned host memory
(H2D of the reads and D2H of the results inside the timed region).
N > 1 (torchrun): weak scaling on a workload that stays a 30x trio: the genome is
5 Mb x N and every rank brings its own 1 M reads/sample; super-k-mer records are
routed to the GPU that owns their minimizer's hash prefix (peer-to-peer stores over
NVLink, or --exchange nccl: one all-to-all of packed keys per sample).
After the timed loop the run's RESULT (histogram, HashList, pulled read names) is
checked against digests of the reference's own output on the same workload
(tests/golden/bench_digests.json); a mismatch makes the exit code 3.
"""
import argparse
import json
import os
import shutil
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "Gbases/sec through k-mer count+filter (k=25, 30x trio)"
K, READ_LEN, MIN_COV, MAX_COV, MINQ, THRESH = 25, 150, 5, 150, 15, 1
FALLBACK_HBM_GBS = 6650.0   # /opt/skills/guides/B200_PROFILING.md fallback


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# --------------------------------------------------------------------------- clocks
class ClockSampler:
    """SM clock + throttle reasons sampled DURING the timed region. NVML in a thread (a sample every
    ~2 ms; the timed region of the default run is only tens of ms, too short for `nvidia-smi -lms`),
    with nvidia-smi as the fallback when pynvml is unavailable."""
    REASONS = {"hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40, "sw_power_cap": 0x4}

    def __init__(self, gpu_index):
        self.idx, self.sm, self.reasons, self.mx = gpu_index, [], set(), None
        self._stop, self._thr, self._nvml = threading.Event(), None, None

    def _poll(self):
        n = self._nvml
        while not self._stop.is_set():
            try:
                self.sm.append(int(n.nvmlDeviceGetClockInfo(self._h, n.NVML_CLOCK_SM)))
                try:
                    mask = int(n.nvmlDeviceGetCurrentClocksEventReasons(self._h))
                except Exception:
                    mask = int(n.nvmlDeviceGetCurrentClocksThrottleReasons(self._h))
                for name, bit in self.REASONS.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception as e:
                log("clock sample failed:", e)
                return
            time.sleep(0.002)

    def start(self):
        try:
            import pynvml as n
            n.nvmlInit()
            # NVML enumerates all GPUs of the box; CUDA_VISIBLE_DEVICES may remap: go through the PCI bus id
            import torch
            bus = torch.cuda.get_device_properties(self.idx).pci_bus_id if hasattr(
                torch.cuda.get_device_properties(self.idx), "pci_bus_id") else None
            self._h = None
            if bus is not None:
                for i in range(n.nvmlDeviceGetCount()):
                    h = n.nvmlDeviceGetHandleByIndex(i)
                    if int(n.nvmlDeviceGetPciInfo(h).bus) == int(bus):
                        self._h = h
                        break
            if self._h is None:
                self._h = n.nvmlDeviceGetHandleByIndex(self.idx)
            self.mx = int(n.nvmlDeviceGetMaxClockInfo(self._h, n.NVML_CLOCK_SM))
            self._nvml = n
            self._thr = threading.Thread(target=self._poll, daemon=True)
            self._thr.start()
        except Exception as e:
            log("NVML clock sampler unavailable:", e)

    def stop(self):
        self._stop.set()
        if self._thr:
            self._thr.join(timeout=1.0)
        if not self.sm:      # fallback: one nvidia-smi query right after the timed region
            try:
                q = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,clocks.max.sm", "--format=csv,noheader,nounits",
                                    "-i", str(self.idx)], capture_output=True, text=True, timeout=20).stdout.split(",")
                self.sm, self.mx = [int(q[0])], int(q[1])
            except Exception as e:
                log("clock sampler unavailable:", e)
        return {"sm_mhz": int(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.mx,
                "reasons": sorted(self.reasons), "samples": len(self.sm)}


def bind_to_gpu_cpus(gpu_index):
    """N > 1 (one process per GPU under torchrun, which binds nothing): run this rank on the CPUs NVML reports as local to
    its GPU, BEFORE any pinned host buffer is allocated, so that the end-to-end leg's host->device copies read memory of
    the socket the GPU hangs off instead of crossing the inter-socket link with seven other ranks. Best effort: any
    failure (no NVML, a cpuset that leaves fewer than four CPUs) leaves the affinity as it was. Returns what was done."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
        allowed = os.sched_getaffinity(0)
        words = max(max(allowed), (os.cpu_count() or 1) - 1) // 64 + 1      # (a cpuset names HOST cpu numbers)
        masks = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        local = {64 * i + b for i, m in enumerate(masks) for b in range(64) if (int(m) >> b) & 1}
        want = local & allowed
        if len(want) < 4:
            return f"kept ({len(want)} of the GPU's CPUs in this process's cpuset)"
        if want == allowed:
            return f"already local ({len(want)} CPUs)"
        os.sched_setaffinity(0, want)
        return f"{len(want)} CPUs local to GPU {gpu_index} (of {len(allowed)})"
    except Exception as e:          # noqa: BLE001 - measurement hygiene only, never fatal
        return f"kept ({type(e).__name__})"


# --------------------------------------------------------------------------- data
def make_data(args, rank, world):
    """this rank's reads of the benchmark workload (rufus_b200/workload.py: genome = unit x world, slice = rank)"""
    from rufus_b200 import workload
    t0 = time.time()
    trio = workload.make_slice(args.seed, args.genome, args.pairs, rank, world)
    log(f"synthetic trio: {args.pairs} pairs/sample/rank, genome {args.genome * world}, slice {rank}/{world}, {time.time() - t0:.1f}s")
    return trio


# --------------------------------------------------------------------------- CPU reference
def have_reference_binaries():
    """oracle/_ref holds the reference's own executables (built by oracle/Makefile where /root/reference exists)"""
    refbin = os.path.join(ROOT, "oracle", "_ref")
    return os.path.exists(os.path.join(refbin, "jellyfish")) and os.path.exists(os.path.join(refbin, "RUFUS.Filter"))


def write_reference_fastq(trio, n_pairs, fqdir):
    """the FASTQ files the reference's commands read (untimed; written once per run): first n_pairs pairs of every sample"""
    from rufus_b200 import synth
    fq = {}
    for s in synth.SAMPLES:
        s1, q1, s2, q2 = (a[:n_pairs] for a in trio[s])
        p1, p2 = os.path.join(fqdir, f"{s}.m1.fq"), os.path.join(fqdir, f"{s}.m2.fq")
        open(p1, "wb").write(synth.fastq_bytes(s1, q1, prefix=s[0]))
        open(p2, "wb").write(synth.fastq_bytes(s2, q2, prefix=s[0]))
        fq[s] = (p1, p2)
    return fq


def reference_output_digests(workdir):
    """order-free digests of what the reference's commands left in `workdir` (the same three digests as
    rufus_b200/workload.py: subject histo file, sorted HashList, sorted pulled mate-1 names)"""
    import hashlib
    from rufus_b200 import workload
    hl = sorted(open(os.path.join(workdir, "HashList")).read().splitlines())
    names = sorted(open(os.path.join(workdir, "out.Mutations.Mate1.fastq")).read().splitlines()[0::4])
    return {"histo_md5": hashlib.md5(open(os.path.join(workdir, "Child.Jhash.histo"), "rb").read()).hexdigest(),
            "hashlist_md5": workload._md5(hl), "hashlist_lines": len(hl),
            "pulled_md5": workload._md5(names), "pulled_pairs": len(names)}


def cpu_reference_pass(trio, n_pairs, workdir, threads, fq=None):
    """One pass of the reference's own commands on the first n_pairs pairs
    of every sample (FASTQ files `fq` from write_reference_fastq, written here when absent: untimed either way).
    Returns (seconds, bases, kind, info)."""
    from rufus_b200 import synth
    refbin = os.path.join(ROOT, "oracle", "_ref")
    jf = os.path.join(refbin, "jellyfish")
    have_ref = os.path.exists(jf) and os.path.exists(os.path.join(refbin, "RUFUS.Filter"))
    bases = 0
    if have_ref:
        if fq is None:
            fq = write_reference_fastq(trio, n_pairs, workdir)
        bases = len(synth.SAMPLES) * 2 * n_pairs * READ_LEN
        t0 = time.time()
        T = max(1, threads)
        for s in synth.SAMPLES:
            # RunJellyForRUFUS.sh:29 (right-sized -s: results are independent of -s) and :37
            subprocess.run(f"cat {fq[s][0]} {fq[s][1]} | {jf} count --disk -m {K} -L 2 -s 256M -t {T} "
                           f"-o {s}.Jhash -C /dev/stdin && {jf} histo -f -o {s}.Jhash.histo {s}.Jhash",
                           shell=True, check=True, cwd=workdir, executable="/bin/bash")
        # runRufus.sh:925-926 + CheckJellyHashList.sh:12
        subprocess.run(f"{jf} merge Child.Jhash Mother.Jhash Father.Jhash > merge.out && "
                       f"{jf} query -s <(awk '{{print \">\"$1\"\\n\"$1}}' merge.out) Child.Jhash | "
                       f"awk -v var={MIN_COV} ' $2 >= var ' | awk -v var={MAX_COV} ' $2 <= var ' > HashList",
                       shell=True, check=True, cwd=workdir, executable="/bin/bash")
        # runRufus.sh:981
        subprocess.run(f"{refbin}/RUFUS.Filter HashList {fq['Child'][0]} {fq['Child'][1]} out {K} {MINQ} {THRESH} {T} "
                       "> /dev/null", shell=True, check=True, cwd=workdir, executable="/bin/bash")
        dt = time.time() - t0
        dg = reference_output_digests(workdir)
        return dt, bases, "reference", {"hashlist": dg["hashlist_lines"], "pulled": dg["pulled_pairs"], "cores": T, "digests": dg}
    # fall back to the C port of the reference algorithm (single thread)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    os.environ["OMP_NUM_THREADS"] = "1"
    import oracle
    from rufus_b200 import kmer
    t0 = time.time()
    tabs = []
    for s in synth.SAMPLES:
        s1, q1, s2, q2 = (a[:n_pairs] for a in trio[s])
        both = np.concatenate([s1, s2])
        offs = np.arange(both.shape[0] + 1, dtype=np.uint64) * READ_LEN
        tabs.append(oracle.count(both.reshape(-1), offs, K, True, 2))
        oracle.histo(tabs[-1][1])
        bases += both.size
    hk, hc = oracle.hashlist(tabs[0], tabs[1:], MIN_COV, MAX_COV)
    s1, q1, s2, q2 = (a[:n_pairs] for a in trio["Child"])
    offs = np.arange(n_pairs + 1, dtype=np.uint64) * READ_LEN
    keep = oracle.filter_pairs(s1, q1, offs, s2, q2, offs, K, MINQ, THRESH,
                               oracle.rufus_set_from_kmers(kmer.keys_to_text(hk, K)))
    return time.time() - t0, bases, "port", {"hashlist": int(hk.size), "pulled": int(keep.sum()), "cores": 1}


def cli_stage_commands(bindir_jf, bindir_tools, fq, threads, size):
    """the stage commands of runRufus.sh for a trio, as (stage, shell command) pairs (RunJellyForRUFUS.sh:29,37;
    runRufus.sh:925-926 + CheckJellyHashList.sh:12; runRufus.sh:981), with the executables of the given directories"""
    jf = os.path.join(bindir_jf, "jellyfish")
    cmds = []
    for s in ("Child", "Mother", "Father"):
        cmds.append((f"count_{s}", f"cat {fq[s][0]} {fq[s][1]} | {jf} count --disk -m {K} -L 2 -s {size} -t {threads} -o {s}.Jhash -C /dev/stdin"))
        cmds.append((f"histo_{s}", f"{jf} histo -f -o {s}.Jhash.histo {s}.Jhash"))
    cmds.append(("merge", f"{jf} merge Child.Jhash Mother.Jhash Father.Jhash > merge.out"))
    cmds.append(("query", f"{jf} query -s <(awk '{{print \">\"$1\"\\n\"$1}}' merge.out) Child.Jhash | "
                          f"awk -v var={MIN_COV} ' $2 >= var ' | awk -v var={MAX_COV} ' $2 <= var ' > HashList"))
    cmds.append(("filter", f"{os.path.join(bindir_tools, 'RUFUS.Filter')} HashList {fq['Child'][0]} {fq['Child'][1]} out {K} {MINQ} {THRESH} "
                           f"{threads} > /dev/null"))
    return cmds


def cli_leg(trio):
    """What a user of the pipeline runs: the stage COMMANDS on FASTQ files, with our drop-in executables (tools/bin)
    and with the reference's (oracle/_ref), same argv, full workload. Wall seconds per stage, and the two outputs
    compared (sorted HashList, sorted pulled mate-1 names)."""
    import hashlib
    from rufus_b200 import synth
    refbin = os.path.join(ROOT, "oracle", "_ref")
    ours = os.path.join(ROOT, "tools", "bin")
    threads = os.cpu_count() or 1
    out = {}
    base = tempfile.mkdtemp(prefix="rg_cli_")
    try:
        fq = {}
        for s in synth.SAMPLES:                       # untimed: the FASTQ files both sets of commands read
            s1, q1, s2, q2 = trio[s]
            p1, p2 = os.path.join(base, f"{s}.m1.fq"), os.path.join(base, f"{s}.m2.fq")
            open(p1, "wb").write(synth.fastq_bytes(s1, q1, prefix=s[0]))
            open(p2, "wb").write(synth.fastq_bytes(s2, q2, prefix=s[0]))
            fq[s] = (p1, p2)
        for name, d_jf, d_tools, size in (("rufus_b200", ours, ours, "8G"), ("reference", refbin, refbin, "256M")):
            if not os.path.exists(os.path.join(d_jf, "jellyfish")):
                out[name] = {"unavailable": f"{d_jf}/jellyfish not built"}
                continue
            wd = os.path.join(base, name)
            os.makedirs(wd)
            stages, t_all = {}, time.time()
            for stage, cmd in cli_stage_commands(d_jf, d_tools, fq, threads, size):
                t0 = time.time()
                subprocess.run(cmd, shell=True, check=True, cwd=wd, executable="/bin/bash")
                key = stage.split("_")[0]
                stages[key] = round(stages.get(key, 0.0) + time.time() - t0, 3)
            hl = sorted(open(os.path.join(wd, "HashList")).read().splitlines())
            names = sorted(open(os.path.join(wd, "out.Mutations.Mate1.fastq")).read().splitlines()[0::4])
            out[name] = {"seconds": round(time.time() - t_all, 3), "stages_s": stages, "threads": threads, "jellyfish_s": size,
                         "hashlist_md5": hashlib.md5(("\n".join(hl) + "\n").encode()).hexdigest(), "hashlist_lines": len(hl),
                         "pulled_md5": hashlib.md5(("\n".join(names) + "\n").encode()).hexdigest(), "pulled_pairs": len(names)}
    finally:
        shutil.rmtree(base, ignore_errors=True)
    a, b = out.get("rufus_b200", {}), out.get("reference", {})
    if "seconds" in a and "seconds" in b:
        out["speedup"] = round(b["seconds"] / a["seconds"], 2)
        out["identical_outputs"] = a["hashlist_md5"] == b["hashlist_md5"] and a["pulled_md5"] == b["pulled_md5"]
    bases = sum(trio[s][0].size + trio[s][2].size for s in synth.SAMPLES)
    for side in (a, b):
        if "seconds" in side:
            side["gbases_per_s"] = round(bases / side["seconds"] / 1e9, 4)
    out["note"] = ("wall clock of the stage commands (count x3, histo x3, merge, query, RUFUS.Filter) on FASTQ files, process start-up, "
                   "FASTQ parsing and .Jhash I/O included; -s 8G for the drop-in (its table grows with the input; -s only sets the "
                   "header geometry) and a right-sized -s 256M for the reference (SURVEY 8c: results do not depend on -s)")
    return out


def reference_sample_pairs(args, world):
    """pairs/sample the reference's commands are run on. N = 1: the WHOLE configured workload (C2: 500 k pairs per
    sample, ~8 s per pass on 16 cores), so the reference arm and the CUDA arm run the same config at the same size and
    the reference's outputs can be checked against the committed digests in the same run. N > 1: the union of the rank
    slices is N times that, so the arm stays a bounded sample (--ref-pairs, default 100 k pairs of rank 0's slice)."""
    if args.ref_pairs > 0:
        return min(args.ref_pairs, args.pairs)
    return args.pairs if world == 1 and have_reference_binaries() else min(100_000, args.pairs)   # (else: the 1-thread C port)


def reference_parity(info, args, world, n_pairs):
    """the digests of the reference's OWN outputs in this run against tests/golden/bench_digests.json (only when the
    pass covered the whole N = 1 workload: the committed digests are those of the full config)"""
    from rufus_b200 import workload
    dg = (info or {}).get("digests")
    if dg is None or world != 1 or n_pairs != args.pairs:
        return None
    exp = workload.expected(1, seed=args.seed, pairs=args.pairs, genome_unit=args.genome, k=K)
    if exp is None:
        return dict(dg, ok=None, reference="no committed digests for this workload")
    bad = [f for f in ("histo_md5", "hashlist_md5", "pulled_md5") if exp.get(f) != dg[f]]
    return dict(dg, ok=not bad, mismatch=bad,
                reference=workload.digest_key(1, seed=args.seed, pairs=args.pairs, genome_unit=args.genome, k=K))


def run_reference_arm(args, rank, world):
    if rank != 0:
        return
    n_pairs = reference_sample_pairs(args, world)
    trio = make_data(args, 0, world)
    threads = os.cpu_count() or 1
    times, kind, info, bases = [], None, None, 0
    par, shrunk = None, ""
    have_ref = have_reference_binaries()
    fqdir = tempfile.mkdtemp(prefix="rg_ref_fq_")

    def one_pass(n, fq):
        wd = tempfile.mkdtemp(prefix="rg_ref_")
        try:
            return cpu_reference_pass(trio, n, wd, threads, fq=fq)
        finally:
            shutil.rmtree(wd, ignore_errors=True)

    try:
        fq = write_reference_fastq(trio, n_pairs, fqdir) if have_ref else None
        i = 0
        while i < args.warmup + args.steps:
            try:
                dt, bases, kind, info = one_pass(n_pairs, fq)
            except Exception as e:          # noqa: BLE001
                if n_pairs <= 100_000 or args.ref_pairs != 0:
                    raise
                shrunk = f"; the pass over the whole workload failed on this host ({type(e).__name__})"
                log("reference pass over the whole workload failed, falling back to the sample:", e)
                n_pairs = 100_000
                fq = write_reference_fastq(trio, n_pairs, fqdir) if have_ref else None
                continue
            log(f"reference pass {i}: {dt:.2f}s ({n_pairs} pairs/sample)")
            if par is None:
                par = reference_parity(info, args, world, n_pairs)
            # the whole run has to end within a few minutes whatever the host: when W + K passes over the whole workload
            # would take longer than --ref-budget seconds, this pass was a calibration (its outputs still checked against
            # the digests) and the W + K passes run on the bounded sample instead
            if i == 0 and n_pairs > 100_000 and args.ref_pairs == 0 and dt * (args.warmup + args.steps) > args.ref_budget:
                shrunk = (f"; one calibration pass over the whole workload took {dt:.1f}s on this host, "
                          f"x{args.warmup + args.steps} passes > {args.ref_budget:.0f}s budget")
                n_pairs = 100_000
                fq = write_reference_fastq(trio, n_pairs, fqdir) if have_ref else None
                continue
            if i >= args.warmup:
                times.append(dt)
            i += 1
    finally:
        shutil.rmtree(fqdir, ignore_errors=True)
    ms = 1e3 * sum(times) / len(times)
    val = bases / (ms / 1e3) / 1e9
    whole = world == 1 and n_pairs == args.pairs
    sample = ((f"the whole workload: {n_pairs} pairs/sample" if whole else
               f"first {n_pairs} of the {args.pairs} pairs/sample of rank 0's slice, per-base normalised")
              + f" ({bases / 1e6:.0f} Mbases per pass), jellyfish -s 256M -t {info['cores']}" + shrunk)
    out = {"impl": "reference", "metric": METRIC, "value": val, "unit": "Gbases/s", "n_gpus": world,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
           "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
           "config": dict(workload_config(args, world), reference_sample=sample),
           "cpu_baseline": {"value": val, "unit": "Gbases/s", "cores": info["cores"], "kind": kind, "sample": sample},
           "e2e": {"value": val, "unit": "Gbases/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    if par is not None:
        out["parity"] = par        # the reference's outputs ARE the committed digests (same run, same box)
    print(json.dumps(out), flush=True)


def workload_config(args, world):
    return {"workload": f"synthetic 30x trio (synth-trio v1, seed {args.seed}): genome {args.genome} bp x {world}, "
                        f"{2 * args.pairs} x {READ_LEN} bp reads per sample per GPU, k={K}, count 3 samples -> histo -> "
                        f"subtract (min {MIN_COV}, max {MAX_COV}) -> pull subject pairs (MinQ {MINQ})"
                        + (" = BASELINE configs[1]" if world == 1 and args.pairs == 500_000 and args.genome == 5_000_000 else ""),
            "reads_per_sample_per_gpu": 2 * args.pairs, "read_len": READ_LEN, "k": K, "samples": 3,
            "genome_bp": args.genome * world, "coverage": round(2.0 * args.pairs * READ_LEN / args.genome, 1),
            "partition": f"minimizer hash prefix x{world}" if world > 1 else "single GPU",
            "l2": "inputs (450 MB reads + GB-scale tables per GPU) larger than L2; no flush needed"}


# --------------------------------------------------------------------------- the large configs
def run_big(args, rank, local_rank, world, dev):
    """BASELINE configs[2..4], scaled (workload.BIG): inputs generated in HBM by the counter-based generator, the
    same count -> histo -> subtract -> pull pass (the quad sweeps the min-count cutoff over the counted tables),
    sliced parity against tests/golden/big_digests.json. One JSON line like the default config's."""
    import ctypes as C
    import torch
    from rufus_b200 import api, pipeline, workload
    if world > 1:
        import torch.distributed as dist
    cfg = workload.BIG[args.config]
    k, ns = cfg["k"], len(cfg["samples"])
    ctx = api.Context(local_rank)
    ctx.use_stream(torch.cuda.current_stream().cuda_stream)
    L = api.lib()
    t0 = time.time()
    streams, subj = [], None
    for si in range(ns):
        P = cfg["pairs"][si]
        buf = torch.empty(2 * P * READ_LEN, dtype=torch.uint8, device=dev)            # both mates, one read stream
        q = torch.empty(2 * P * READ_LEN, dtype=torch.uint8, device=dev)
        sp = workload.big_spec(cfg, si, rank, world)
        api._check(L.rg_synth_pairs_dev(ctx.h, C.byref(sp), api._ptr(buf[:P * READ_LEN]), api._ptr(q[:P * READ_LEN]),
                                        api._ptr(buf[P * READ_LEN:]), api._ptr(q[P * READ_LEN:])), "rg_synth_pairs_dev")
        streams.append(buf)
        if si == 0:
            subj = (buf[:P * READ_LEN], q[:P * READ_LEN], buf[P * READ_LEN:], q[P * READ_LEN:])
        else:
            del q
    torch.cuda.synchronize()
    log(f"rank {rank}: {args.config} inputs generated on the GPU in {time.time() - t0:.2f}s "
        f"({sum(int(s.numel()) for s in streams) / 1e9:.2f} Gbases)")
    bases_per_step = sum(int(s.numel()) for s in streams)
    reads_max = max(cfg["pairs"]) * 2
    est = int((1.05 * cfg["genome_unit"] + 0.27 * k * reads_max) * args.est_scale)
    sweep = list(cfg["min_cov"])
    pipe = pipeline.DevicePipeline(ctx, k, streams, subj, READ_LEN, est, max(sweep[0], 5), cfg["max_cov"], MINQ, THRESH,
                                   lower=(2,) * ns, rank=rank, world=world, exchange=args.exchange,
                                   groups_per_shard=args.p2p_groups, scatter_ctas=args.scatter_ctas, fast_cap=1 << 18)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def one_step():
        r = pipe.step()
        out = {sweep[0]: (r["n_mutant"], r["n_kept"])}
        for mc in sweep[1:]:                       # the quad's min-count sweep: cutoffs over the counted tables
            r = pipe.rerun_tail(max(mc, 5))
            out[mc] = (r["n_mutant"], r["n_kept"])
        pipe.min_cov = max(sweep[0], 5)
        return out

    for i in range(args.warmup):
        res = one_step()
    geo = pipe.tables[0].geometry()
    st0 = pipe.tables[0].stats()
    log(f"rank {rank}: warm; (mutant k-mers, pulled pairs) per min-count {res}; subject table {st0}, geometry {geo}")
    sampler = ClockSampler(local_rank)
    ctx.profile_reset()
    ctx.profile(True)
    launches0 = ctx.launches
    barrier()
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        res = one_step()
    e1.record()
    barrier()
    clocks = sampler.stop()
    ms_total = e0.elapsed_time(e1)
    launches = ctx.launches - launches0
    prof = ctx.profile_report()
    ctx.profile(False)
    # ---- sliced parity (untimed): every sample's slice of the table + the slice of the HashList per cutoff
    exp = workload.big_expected(args.config, world)
    got = {"samples": [], "hashlist": {}}
    cap = max(1 << 20, int(pipe.tables[0].stats()["distinct"]) // workload.SLICE_MOD * 2)
    dk = torch.empty(cap, dtype=torch.int64, device=dev)
    dc = torch.empty(cap, dtype=torch.int32, device=dev)

    def gather(keys, counts):
        if world == 1:
            return keys, counts
        parts = [None] * world
        dist.all_gather_object(parts, (keys, counts))
        return np.concatenate([p[0] for p in parts]), np.concatenate([p[1] for p in parts])

    for t in pipe.tables:
        n = L.rg_dump_slice_dev(t.h, 2, workload.SLICE_MOD, 0, api._ptr(dk), api._ptr(dc), cap, None)
        if n < 0:
            raise api._err("rg_dump_slice_dev")
        kk, cc = gather(dk[:n].cpu().numpy().view(np.uint64), dc[:n].cpu().numpy().view(np.uint32))
        got["samples"].append(workload.table_digest(kk, cc))
    for mc in sweep:
        r = pipe.rerun_tail(max(mc, 5))
        pr = pipe.results()
        kk, cc = gather(pr["keys"], pr["counts"])
        m = workload.slice_hash(kk) % workload.SLICE_MOD == 0
        got["hashlist"][str(mc)] = workload.table_digest(kk[m], cc[m])
    if exp is None:
        parity = dict(got, ok=None, note="no sliced digest committed for this config / GPU count")
    else:
        bad = [f"sample{i}" for i in range(ns) if got["samples"][i] != exp["samples"][i]] + \
              [f"hashlist{mc}" for mc in got["hashlist"] if got["hashlist"][mc] != exp["hashlist"][mc]]
        parity = {"ok": not bad, "mismatch": bad, "slice": f"slice_hash(key) % {workload.SLICE_MOD} == 0",
                  "reference": workload.big_key(args.config, world), "slice_keys": [s["n"] for s in got["samples"]],
                  "hashlist_slice": {mc: d["n"] for mc, d in got["hashlist"].items()}}
    if world > 1:
        t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
        lt = torch.tensor([launches], dtype=torch.int64, device=dev)
        dist.all_reduce(lt)
        launches = int(lt.item())
    ms_step = ms_total / args.steps
    value = bases_per_step * world / (ms_step / 1e3) / 1e9
    peak, peak_src = FALLBACK_HBM_GBS, "fallback"
    try:
        mp = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        peak, peak_src = float(mp["hbm_gbs"]), "measured"
    except Exception:
        pass
    kern = {kk: {"ms_avg": v[0] / max(v[1], 1), "launches": v[1]} for kk, v in prof.items()}
    names = [n for n in ("partition_kernel", "partition_p2p_kernel", "partition_groups_kernel", "dedup_kernel", "tile_insert_kernel") if n in kern]
    # per sample push: (sum over its kernels of total time) / pushes; algorithmic bytes B + 16 n
    pushes = args.steps * ns
    count_ms = sum(kern[n]["ms_avg"] * kern[n]["launches"] for n in names) / max(pushes, 1)
    bases_per_push = bases_per_step / ns
    alg_bytes = bases_per_push * (1.0 + 16.0 * (READ_LEN - k + 1) / READ_LEN)
    achieved = alg_bytes / (count_ms / 1e3) / 1e9 if count_ms > 0 else 0.0
    out = {"metric": METRIC.replace("k=25, 30x trio", cfg["title"]), "value": value, "unit": "Gbases/s", "n_gpus": world,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "u64", "data": "synthetic",
           "config": {"workload": f"{args.config}: {cfg['title']}; per GPU {[2 * p for p in cfg['pairs']]} x {READ_LEN} bp reads "
                                  f"per sample, genome {cfg['genome_unit']} bp x {world} (5 % repeat families), k={k}, "
                                  f"min-count {sweep}, max {cfg['max_cov']}; inputs generated on the GPU (seed {cfg['seed']})",
                      "k": k, "samples": ns, "genome_bp": cfg["genome_unit"] * world,
                      "reads_per_sample": [2 * p * world for p in cfg["pairs"]],
                      "partition": f"minimizer hash prefix x{world}" if world > 1 else "single GPU",
                      "l2": "inputs and tables (GBs per GPU) larger than L2; no flush needed"},
           "clocks": clocks, "gpu_launches": launches,
           "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                        "traffic": None, "peak_source": peak_src, "kernel": "+".join(names),
                        "algorithmic_bytes_per_launch": alg_bytes, "ms_per_launch": count_ms,
                        "note": "one 'launch' = the count of one sample on one GPU: B + 16*n bytes (SURVEY 8d)"},
           "kernels_ms": kern, "result": {str(mc): {"n_mutant_kmers": v[0], "n_pulled_pairs_rank0": v[1]} for mc, v in res.items()},
           "table": {"distinct_subject_shard": st0["distinct"], "geometry": geo}, "parity": parity}
    if rank == 0 and not args.no_cpu_baseline:
        # the reference's CPU path on 1/64 of rank 0's reads of every sample (generated on the host: same bytes)
        sub = {}
        for si in range(ns):
            n = max(1000, cfg["pairs"][si] // 64)
            sub[si] = api.synth_pairs_host(workload.big_spec(cfg, si, 0, world, 0, n))
        wd = tempfile.mkdtemp(prefix="rg_cpu_")
        try:
            dt, b = cpu_reference_big(sub, k, max(sweep[0], 5), cfg["max_cov"], wd, os.cpu_count() or 1)
        finally:
            shutil.rmtree(wd, ignore_errors=True)
        if dt:
            out["cpu_baseline"] = {"value": b / dt / 1e9, "unit": "Gbases/s", "cores": os.cpu_count() or 1, "kind": "reference",
                                   "extrapolated": True,
                                   "sample": f"first 1/64 of rank 0's pairs of every sample ({b / 1e6:.0f} Mbases), {dt:.1f}s wall, "
                                             "jellyfish -s 256M; per-base rate, extrapolated linearly to the full config"}
    if rank == 0:
        print(json.dumps(out), flush=True)
    pipe.free()
    if world > 1:
        dist.destroy_process_group()
    ctx.close()
    if rank == 0 and parity["ok"] is False:
        log("SLICED PARITY MISMATCH:", parity)
        sys.exit(3)


def cpu_reference_big(sub, k, min_cov, max_cov, workdir, threads):
    """the reference's commands on a sub-sample {sample index: (s1, q1, s2, q2)}, subject = index 0"""
    from rufus_b200 import synth
    refbin = os.path.join(ROOT, "oracle", "_ref")
    jf = os.path.join(refbin, "jellyfish")
    if not (os.path.exists(jf) and os.path.exists(os.path.join(refbin, "RUFUS.Filter"))):
        return None, 0
    bases, fq = 0, {}
    for si, (s1, q1, s2, q2) in sub.items():
        p1, p2 = os.path.join(workdir, f"s{si}.m1.fq"), os.path.join(workdir, f"s{si}.m2.fq")
        open(p1, "wb").write(synth.fastq_bytes(s1, q1, prefix=f"s{si}_"))
        open(p2, "wb").write(synth.fastq_bytes(s2, q2, prefix=f"s{si}_"))
        fq[si] = (p1, p2)
        bases += s1.size + s2.size
    T = max(1, threads)
    t0 = time.time()
    for si in fq:
        subprocess.run(f"cat {fq[si][0]} {fq[si][1]} | {jf} count --disk -m {k} -L 2 -s 256M -t {T} -o s{si}.Jhash -C /dev/stdin "
                       f"&& {jf} histo -f -o s{si}.Jhash.histo s{si}.Jhash", shell=True, check=True, cwd=workdir,
                       executable="/bin/bash")
    subprocess.run(f"{jf} merge " + " ".join(f"s{si}.Jhash" for si in fq) + " > merge.out && "
                   f"{jf} query -s <(awk '{{print \">\"$1\"\\n\"$1}}' merge.out) s0.Jhash | "
                   f"awk -v var={min_cov} ' $2 >= var ' | awk -v var={max_cov} ' $2 <= var ' > HashList",
                   shell=True, check=True, cwd=workdir, executable="/bin/bash")
    subprocess.run(f"{refbin}/RUFUS.Filter HashList {fq[0][0]} {fq[0][1]} out {k} {MINQ} {THRESH} {T} > /dev/null",
                   shell=True, check=True, cwd=workdir, executable="/bin/bash")
    return time.time() - t0, bases


# --------------------------------------------------------------------------- our arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--pairs", type=int, default=500_000, help="read pairs per sample per GPU")
    ap.add_argument("--genome", type=int, default=5_000_000, help="genome bases per GPU (the genome is N times that)")
    ap.add_argument("--seed", type=int, default=42)
    ap.add_argument("--ref-pairs", type=int, default=0,
                    help="pairs/sample the reference's commands run on (0: the whole workload at N=1, 100 k of rank 0's slice at N>1)")
    ap.add_argument("--ref-budget", type=float, default=420.0,
                    help="--impl reference: seconds the W + K passes may take before the arm falls back to the bounded sample")
    ap.add_argument("--mode", default="tiled", choices=["tiled", "direct"])
    ap.add_argument("--est-scale", type=float, default=1.0, help="scale the distinct k-mer estimate (table load factor)")
    ap.add_argument("--p2p-groups", type=int, default=0, help="hash-prefix groups per shard of the peer-to-peer partition")
    ap.add_argument("--scatter-ctas", type=int, default=0,
                    help="CTAs of the (NVLink-bound) peer-to-peer scatter when it overlaps the previous sample's inserts")
    ap.add_argument("--exchange", default="p2p", choices=["p2p", "nccl"],
                    help="N>1: peer-to-peer partition into the owners' slabs, or extract + NCCL all-to-all")
    ap.add_argument("--config", default="c2", choices=["c2", "c3", "c4", "c5"],
                    help="c2: BASELINE configs[1] (the default, the metric's config); c3/c4/c5: the large configs[2..4], "
                         "scaled (rufus_b200/workload.py BIG), inputs generated on the GPU, sliced parity")
    ap.add_argument("--filter-exchange", default="set", choices=["set", "a2a"],
                    help="N>1 read pull: all-gather the mutant set (layout 1, default) or keep it sharded and exchange the reads (layout 2)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--cli", action="store_true",
                    help="N=1: also time the stage COMMANDS (tools/bin/* vs oracle/_ref/*) on FASTQ files of the full workload")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl != "reference" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        return run_reference_arm(args, rank, world)

    cpu_binding = bind_to_gpu_cpus(local_rank) if world > 1 else "none (single process)"
    import torch
    from rufus_b200 import api, pipeline, synth
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the CUDA path has no CPU fallback "
                         "(use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")   # NCCL's banner / warnings stay off stdout: one JSON line only
        dist.init_process_group("nccl", device_id=dev)

    if args.config != "c2":
        return run_big(args, rank, local_rank, world, dev)
    trio = make_data(args, rank, world)
    ctx = api.Context(local_rank)
    ctx.use_stream(torch.cuda.current_stream().cuda_stream)
    mode = api.MODE_TILED if args.mode == "tiled" else api.MODE_DIRECT

    # resident inputs: per sample both mates as one [n_reads*150] stream; subject mates + quals for the pull
    streams = []
    for s in synth.SAMPLES:
        s1, _, s2, _ = trio[s]
        streams.append(torch.from_numpy(np.concatenate([s1, s2]).reshape(-1)).to(dev))
    subj = tuple(torch.from_numpy(a.reshape(-1)).to(dev) for a in trio["Child"])
    bases_per_step = sum(int(s.numel()) for s in streams)
    # distinct k-mers per shard: genomic (two haplotypes, ~3 % private; the genome grows with the GPU count, so a
    # shard holds one unit's worth) + ~6.2 error k-mers per read
    est = int((1.03 * args.genome + 6.2 * 2 * args.pairs) * args.est_scale)
    pipe = pipeline.DevicePipeline(ctx, K, streams, subj, READ_LEN, est, MIN_COV, MAX_COV, MINQ, THRESH,
                                   mode=mode, rank=rank, world=world, exchange=args.exchange,
                                   groups_per_shard=args.p2p_groups, scatter_ctas=args.scatter_ctas,
                                   filter_exchange=args.filter_exchange)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(args.warmup):
        r = pipe.step()
    st0 = pipe.tables[0].stats()
    log(f"rank {rank}: warm; mutant k-mers {r['n_mutant']}, pulled {r['n_kept']}, subject table {st0}, "
        f"geometry {pipe.tables[0].geometry()}")
    n_occ = [t.stats()["occurrences"] for t in pipe.tables]

    sampler = ClockSampler(local_rank)
    ctx.profile_reset()
    ctx.profile(True)
    launches0 = ctx.launches
    barrier()
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        r = pipe.step()
    e1.record()
    barrier()
    clocks = sampler.stop()
    # ---- parity of the benchmarked run (untimed): the job's results against the digests of the REFERENCE's
    # output on the same workload (tests/golden/bench_digests.json, made by tests/golden/make_bench_golden.py)
    from rufus_b200 import workload
    g = pipe.gathered_results()
    parity = workload.check(world, g["histo"], g["keys"], g["counts"], g["names"], seed=args.seed, pairs=args.pairs,
                            genome_unit=args.genome, k=K) if rank == 0 else None
    ms_total = e0.elapsed_time(e1)
    launches = ctx.launches - launches0
    prof = ctx.profile_report()
    ctx.profile(False)
    if world > 1:
        t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
        lt = torch.tensor([launches], dtype=torch.int64, device=dev)
        dist.all_reduce(lt)
        launches = int(lt.item())
    ms_step = ms_total / args.steps
    value = bases_per_step * world / (ms_step / 1e3) / 1e9

    # ---- roofline of the count stage (dominant): partition_kernel + tile_insert_kernel per sample push
    peak, peak_src = FALLBACK_HBM_GBS, "fallback"
    try:
        mp = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        peak, peak_src = float(mp["hbm_gbs"]), "measured"
    except Exception:
        pass
    kern = {k: {"ms_avg": v[0] / max(v[1], 1), "launches": v[1]} for k, v in prof.items()}
    if world == 1:
        names = ["partition_kernel", "partition_groups_kernel", "dedup_kernel", "tile_insert_kernel"] if mode == api.MODE_TILED \
            else ["count_direct_kernel"]          # (partition_groups_kernel only runs for tables beyond one partition pass)
    elif args.exchange == "p2p":
        names = ["partition_p2p_kernel", "partition_groups_kernel", "dedup_kernel", "tile_insert_kernel"]
    else:
        names = ["extract_parts_kernel<count>", "extract_parts_kernel<scatter>", "partition_keys_kernel",
                 "tile_insert_kernel"] if mode == api.MODE_TILED else ["extract_parts_kernel<count>",
                                                                      "extract_parts_kernel<scatter>",
                                                                      "insert_keys_kernel"]
    count_ms = sum(kern[n]["ms_avg"] for n in names if n in kern)
    n_occ_avg = float(np.mean(n_occ)) * (1 if world == 1 else 1.0)
    bases_per_push = bases_per_step / 3
    alg_bytes = bases_per_push + 16.0 * (n_occ_avg if world == 1 else bases_per_push * (READ_LEN - K + 1) / READ_LEN)
    achieved = alg_bytes / (count_ms / 1e3) / 1e9 if count_ms > 0 else 0.0
    traffic = None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        traffic = sum(tj[n] for n in names)
    except Exception:
        pass
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "peak_source": peak_src, "kernel": "+".join(names),
                "algorithmic_bytes_per_launch": alg_bytes, "ms_per_launch": count_ms,
                "note": "one 'launch' = the count of one sample (150 Mbases): B + 16*n bytes (SURVEY 8d)"}

    out = {"metric": METRIC, "value": value, "unit": "Gbases/s", "n_gpus": world, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "u64", "data": "synthetic",
           "config": dict(workload_config(args, world), cpu_binding=cpu_binding),
           "clocks": clocks, "gpu_launches": launches, "roofline": roofline, "kernels_ms": kern,
           "result": {"n_mutant_kmers": r["n_mutant"], "n_pulled_pairs_rank0": r["n_kept"]}, "parity": parity}

    # ---- e2e: the same pass with every input starting in pinned HOST memory and every result ending there
    # (H2D + D2H inside the timed region). N = 1: the host-buffer C ABI (rg_count_push / rg_histo / rg_subtract /
    # rg_filter_pairs, H2D pipelined against the kernels). N > 1: the sharded pipeline fed from host buffers.
    if not args.no_e2e:
        e2e_steps = max(3, args.steps // 2)
        e2e_binding, saved_cpus = cpu_binding, None
        if world == 1:
            # the pinned buffers of this leg are allocated (and first touched) on the CPUs next to the GPU; the previous
            # affinity is restored after the leg, before the CPU baseline runs the reference's binaries on every core
            try:
                saved_cpus = os.sched_getaffinity(0)
            except Exception:           # noqa: BLE001
                saved_cpus = None
            e2e_binding = bind_to_gpu_cpus(local_rank) if saved_cpus else "kept"
            pipe.free()
            del pipe
            torch.cuda.empty_cache()
            nl = np.full((2 * args.pairs, 1), 10, np.uint8)
            host_streams = []
            for s in synth.SAMPLES:
                s1, _, s2, _ = trio[s]
                a = torch.from_numpy(np.concatenate([np.concatenate([s1, s2]), nl], axis=1).reshape(-1)).pin_memory()
                host_streams.append(a.numpy())
            pairs_host = tuple(torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy() for a in trio["Child"])
            tables = [api.KmerTable(ctx, K, True, est, mode) for _ in host_streams]
            h2d = sum(a.size for a in host_streams) + sum(a.size for a in pairs_host) + 2 * 8 * (args.pairs + 1)
            res = None
            for i in range(2):
                tm = {}
                res = pipeline.host_step(ctx, K, host_streams, pairs_host, READ_LEN, est, MIN_COV, MAX_COV, MINQ, THRESH,
                                         mode=mode, tables=tables, timings=tm)
                log("e2e warm-up pass, host-side phase times (ms):", {k: round(v, 2) for k, v in tm.items()})
            barrier()
            e0.record()
            for i in range(e2e_steps):
                res = pipeline.host_step(ctx, K, host_streams, pairs_host, READ_LEN, est, MIN_COV, MAX_COV, MINQ, THRESH,
                                         mode=mode, tables=tables)
            e1.record()
            barrier()
            d2h = res["histo"].nbytes + res["hashlist"][0].nbytes + res["hashlist"][1].nbytes + res["keep"].nbytes
            e2e_ok = workload.check(1, res["histo"], res["hashlist"][0], res["hashlist"][1],
                                    workload.pulled_names(res["keep"], "Child", 0, 1), seed=args.seed, pairs=args.pairs,
                                    genome_unit=args.genome, k=K)["ok"]
            note = "rg_count_push/rg_histo/rg_subtract/rg_filter_pairs on pinned host buffers"
            for t in tables:
                t.free()
        else:
            host_streams = [s.cpu().pin_memory() for s in streams]
            pairs_host = tuple(a.cpu().pin_memory() for a in subj)
            h2d = sum(a.numel() for a in host_streams) + sum(a.numel() for a in pairs_host)
            for i in range(2):
                res = pipe.step_from_host(host_streams, pairs_host)
            barrier()
            e0.record()
            for i in range(e2e_steps):
                res = pipe.step_from_host(host_streams, pairs_host)
            e1.record()
            barrier()
            d2h = res["histo"].nbytes + res["keys"].nbytes + res["counts"].nbytes + res["keep"].nbytes
            g2 = pipe.gathered_results()
            e2e_ok = workload.check(world, g2["histo"], g2["keys"], g2["counts"], g2["names"], seed=args.seed,
                                    pairs=args.pairs, genome_unit=args.genome, k=K)["ok"] if rank == 0 else None
            note = "sharded pipeline (DevicePipeline.step_from_host): H2D of every rank's reads, peer-to-peer count, D2H of results"
        ms_e2e = e0.elapsed_time(e1) / e2e_steps
        if world > 1:
            t = torch.tensor([ms_e2e], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms_e2e = float(t.item())
        if saved_cpus:
            try:
                os.sched_setaffinity(0, saved_cpus)
            except Exception:           # noqa: BLE001
                pass
        out["e2e"] = {"value": bases_per_step * world / (ms_e2e / 1e3) / 1e9, "unit": "Gbases/s",
                      "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "ms_per_step": ms_e2e,
                      "steps": e2e_steps, "parity_ok": e2e_ok, "note": note, "cpu_binding": e2e_binding}

    # ---- CPU baseline beside it (rank 0, N = 1 only): ONE pass of the reference's commands over the whole workload
    # (C2: ~8 s wall on 16 cores), its outputs checked against the same committed digests as the CUDA path's
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        n_ref = reference_sample_pairs(args, world)
        for attempt in (n_ref, min(100_000, n_ref)):      # (the measured line is never lost to the baseline leg: a failed
            wd = tempfile.mkdtemp(prefix="rg_cpu_")       # whole-workload pass — disk, memory — falls back to the sample)
            try:
                dt, b, kind, info = cpu_reference_pass(trio, attempt, wd, os.cpu_count() or 1)
                what = "the whole workload" if attempt == args.pairs else f"first {attempt} pairs/sample of the same workload"
                out["cpu_baseline"] = {"value": b / dt / 1e9, "unit": "Gbases/s", "cores": info["cores"], "kind": kind,
                                       "sample": f"{what} ({b / 1e6:.0f} Mbases), one pass, {dt:.1f}s wall, jellyfish -s 256M; "
                                                 f"hashlist {info['hashlist']}, pulled {info['pulled']}"}
                par = reference_parity(info, args, world, attempt)
                if par is not None:
                    out["cpu_baseline"]["parity"] = par
                break
            except Exception as e:          # noqa: BLE001
                log(f"cpu_baseline pass on {attempt} pairs/sample failed: {type(e).__name__}: {e}")
                out["cpu_baseline"] = {"value": None, "unit": "Gbases/s", "cores": os.cpu_count() or 1, "kind": "reference",
                                       "sample": f"failed: {type(e).__name__}"}
            finally:
                shutil.rmtree(wd, ignore_errors=True)
            if attempt == min(100_000, n_ref):
                break
    if rank == 0 and world == 1 and args.cli:
        ctx.close()
        ctx = None
        try:
            out["cli"] = cli_leg(trio)
        except Exception as e:              # noqa: BLE001
            out["cli"] = {"error": f"{type(e).__name__}: {e}"}
    if rank == 0:
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()
    if ctx is not None:
        ctx.close()
    # a run whose RESULT differs from the reference's is not a measurement: fail it
    if rank == 0 and (parity["ok"] is False or out.get("e2e", {}).get("parity_ok") is False):
        log("PARITY MISMATCH against the reference digests:", parity)
        sys.exit(3)


if __name__ == "__main__":
    main()
