#!/usr/bin/env python3
"""bench.py — reads/s and DP GCUPS of the CDR3 extraction hot path on ONE sample of synthetic hs-TRB reads (BASELINE.json configs[2]:
100 M x 150 bp), on 1 GPU or split over the N GPUs of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--reads 100000000] [--length 150] [--impl gpu|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

A step = one pass of the whole hot path (seeding, SW, selection, ordering/dedup, assign/finder, k-mer table, extension) over THE
sample (SURVEY 8d config-3 generator, read i drawn from splitmix64(seed+i)).  With N > 1 the same sample is split into N contiguous
blocks, one per rank (strong scaling): the per-read stages run on each block's GPU, the global steps of the reference (SAM order /
QNAME dedup / V-J join, last-writer-wins k-mer table, result lists) are merged over NCCL by the library's own communicator
(catt_ctx_join; torch.distributed only carries the 128-byte id, the barrier and the timing max).
  value : reads/s with the sample already resident in HBM (catt_shard_upload once, catt_shard_run per step)
  e2e   : reads/s through the public call with HOST buffers (catt_run_reads / catt_run_reads_shard: staging, H2D, kernels, collectives, D2H)
  one_sample : result digest (identical on every N), Stage A spread over the ranks, per-collective time and bytes, limiting collective
  roofline : the SW kernel against the measured integer-ALU (packed DPX add-max) issue peak of the GPU(s)
  replicas : round 1's weak-scaling number (every GPU its own 2 M-read sample, no exchange)
  cpu_baseline : the CPU oracle (OpenMP restatement of the reference path, not the Julia program) on a bounded sample
`--impl reference` times the CPU restatement on all host cores on the same workload (rank 0 only).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tarfile
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SEED = 20251019


def fixture_dir():
    d = tempfile.mkdtemp(prefix="catt_bench_")
    with tarfile.open(os.path.join(ROOT, "tests", "golden", "catt_fixture.tar.gz")) as tar:
        tar.extractall(d)
    return d


class ClockSampler(threading.Thread):
    """SM clocks / throttle reasons during the timed region, sampled in-process through NVML (nvidia_ml_py): forking
    nvidia-smi every 200 ms from a multi-GB process stalls the CUDA driver calls of the step being timed."""

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu, self.rows, self.stop_flag = gpu_index, [], False

    def run(self):
        try:
            import pynvml as N
            N.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = int(vis.split(",")[self.gpu]) if vis and vis.split(",")[self.gpu].isdigit() else self.gpu
            h = N.nvmlDeviceGetHandleByIndex(idx)
            mx = N.nvmlDeviceGetMaxClockInfo(h, N.NVML_CLOCK_SM)
        except Exception:
            return
        bits = {"hw_slowdown": getattr(N, "nvmlClocksEventReasonHwSlowdown", 0x8),
                "hw_thermal_slowdown": getattr(N, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                "sw_thermal_slowdown": getattr(N, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                "sw_power_cap": getattr(N, "nvmlClocksEventReasonSwPowerCap", 0x4)}
        while not self.stop_flag:
            try:
                sm = N.nvmlDeviceGetClockInfo(h, N.NVML_CLOCK_SM)
                try:
                    rs = N.nvmlDeviceGetCurrentClocksEventReasons(h)
                except Exception:
                    rs = N.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                self.rows.append((sm, mx, [k for k, b in bits.items() if rs & b]))
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        sm = sorted(r[0] for r in self.rows)
        reasons = sorted({x for r in self.rows for x in r[2]})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.rows[0][1] if self.rows else None, "reasons": reasons,
                "samples": len(self.rows), "how": "NVML in-process, 200 ms period, during both timed regions"}


def ncu_traffic():
    """dram bytes read + written per launch of the dominant kernel, from the committed ncu --set full capture
    (profiles/traffic.json, written by profiles/summarize_ncu.py); None when no capture is committed."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as fh:
            return json.load(fh)["sw_bulk2_kernel"]
    except Exception:
        return None


def ncu_pipe_alu():
    """sm__inst_executed_pipe_alu (% of peak) of the dominant kernel from the committed ncu capture, None when absent."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as fh:
            return json.load(fh).get("sw_bulk2_kernel_pipe_alu_pct")
    except Exception:
        return None


def ncu_facts():
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as fh:
            return json.load(fh)
    except Exception:
        return {}


def roofline_other(infos, length, world, n_total):
    """The two HBM-side kernels of SURVEY 8d against the measured copy bandwidth (MEASURED_PEAKS.json, else the recipe's fallback).
    Algorithmic bytes: seeding = packed read in (ceil(L/4) + ceil(L/8) B) + the two candidate bitmaps and the task count out per
    read and reference set; k-mer table = 16 B per insert, 11 inserts per Vpart/Jpart read.  Both kernels are latency / random-
    access bound by design (tables are L2-resident), so the fractions are small; they are reported, not optimised for."""
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            peak, src = float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        peak, src = 6650.0, "fallback (B200_PROFILING.md)"
    i = infos[-1]
    n = n_total / world                          # per GPU (the stage times are the slowest GPU's)
    ms_seed = sum(x["ms_seed_kernel"] for x in infos) / len(infos)
    ms_pool = sum(x["ms_pool_kernel"] for x in infos) / len(infos)
    seed_bytes = n * ((-(-length // 4) + -(-length // 8)) * 2 + 2 * 4 * (11 + 1) + 2 * 4)      # CW = 11 words (V) + 1 (J)
    pool_bytes = 16 * 11 * (i["vpart"] + i["jpart"]) / world
    out = {}
    if ms_seed > 0:
        # instruction-issue bound (one warp per read, tables L2-resident): the HBM yardstick SURVEY gave it reads 0.003 and says
        # nothing, so the line carries what bounds it - warp instructions per read and issue-slot use from the committed ncu capture
        a = world * seed_bytes / (ms_seed * 1e-3) / 1e9
        f = ncu_facts()
        step_ms = sum(x["ms_seed"] + x["ms_sw"] + x["ms_select"] + x["ms_order"] + x["ms_extract"] + x["ms_pool"] for x in infos) / len(infos)
        out["seed_kernel"] = {"bound": "instruction issue", "ms_per_step": ms_seed, "share_of_step": ms_seed / step_ms if step_ms else None,
                              "ns_per_read_per_gpu": 1e6 * ms_seed / n,
                              "warp_instr_per_read": f.get("seed_kernel_warp_instr_per_read"), "issue_active_pct": f.get("seed_kernel_issue_active_pct"),
                              "lanes_per_instr": f.get("seed_kernel_lanes_per_instr"), "hbm_gbs_algorithmic": a,
                              "note": "V reference set, both seeding launches; ncu figures from profiles/traffic.json (" + str(f.get("_source")) + ")"}
    if ms_pool > 0:
        a = world * pool_bytes / (ms_pool * 1e-3) / 1e9
        out["pool_kernel"] = {"bound": "hbm", "achieved": a, "peak": peak * world, "unit": "GB/s", "frac": a / (peak * world), "peak_source": src}
    return out


def nbsorb_scans(ctx, np, peak_ops):
    """the low-count neighbour scan of nbsorb (Jtool.jl:553-558) through catt_near_any, 10^6 x 10^4 random 45-nt
    sequences from host buffers (no early exit: every pair is compared), and linkRLG (Jtool.jl:396-421) through catt_link_leaves."""
    rng = np.random.default_rng(1)
    lut = np.frombuffer(b"ACGT", dtype=np.uint8)
    lgt, nq, nt = 45, 1_000_000, 10_000
    q, t = lut[rng.integers(0, 4, size=(nq, lgt))], lut[rng.integers(0, 4, size=(nt, lgt))]
    ctx.near_any(lgt, q, t)
    ts = []
    for _ in range(3):
        t0 = time.perf_counter(); ctx.near_any(lgt, q, t); ts.append(time.perf_counter() - t0)
    dt = sorted(ts)[1]
    nl, nr = 200_000, 2_000
    up = np.tile(np.array([40.0, 12.0, 4.0, 1.5, 1.0, 1.0, 1.0]), (nr, 1))
    cnt = rng.integers(1, 50, size=nl)
    ctx.link_leaves(lgt, t[:nr], up, q[:nl], cnt)
    t0 = time.perf_counter(); ctx.link_leaves(lgt, t[:nr], up, q[:nl], cnt); dl = time.perf_counter() - t0
    # a CLUSTERED repertoire, as real ones are (uniform-random pairs all die in the word-0 reject test): 4,000 roots, every query a
    # root with 1-5 substitutions, so every query has near targets and the full three-plane test and the early exit both matter
    nroot = 4000
    roots = rng.integers(0, 4, size=(nroot, lgt), dtype=np.uint8)
    parent = rng.integers(0, nroot, size=nq)
    cq = roots[parent].copy()
    for _ in range(5):
        hit = rng.random(nq) < 0.6
        pos = rng.integers(0, lgt, size=nq)
        cq[hit, pos[hit]] = (cq[hit, pos[hit]] + rng.integers(1, 4, size=int(hit.sum()))) % 4
    ct = np.concatenate([roots, cq[:nt - nroot]])            # targets: the roots plus 6,000 of the mutated sequences
    cqa, cta = lut[cq], lut[ct]
    flags = ctx.near_any(lgt, cqa, cta)
    tc = []
    for _ in range(3):
        t0 = time.perf_counter(); ctx.near_any(lgt, cqa, cta); tc.append(time.perf_counter() - t0)
    dc = sorted(tc)[1]
    cl_cnt = rng.integers(1, 50, size=nl)
    ctx.link_leaves(lgt, lut[roots[:nr]], up, cqa[:nl], cl_cnt)
    t0 = time.perf_counter(); st, _ = ctx.link_leaves(lgt, lut[roots[:nr]], up, cqa[:nl], cl_cnt); dlc = time.perf_counter() - t0
    return {"near_any": {"pairs": nq * nt, "ms_per_call": dt * 1e3, "pairs_per_s": nq * nt / dt, "base_compares_per_s": nq * nt * (lgt - 6) / dt},
            "link_leaves": {"pairs": nl * nr, "ms_per_call": dl * 1e3, "pairs_per_s": nl * nr / dl},
            "near_any_clustered": {"pairs": nq * nt, "ms_per_call": dc * 1e3, "pairs_per_s": nq * nt / dc, "flagged": int(flags.sum()),
                                   "note": "4,000 roots; queries = roots with up to 5 substitutions; targets = roots + 6,000 mutated sequences: every query "
                                           "has a target within 3 mismatches unless its own mutations exceed that, so blocks leave early"},
            "link_leaves_clustered": {"pairs": nl * nr, "ms_per_call": dlc * 1e3, "pairs_per_s": nl * nr / dlc, "linked_once": int((st == 1).sum()),
                                      "linked_twice_or_more": int((st >= 2).sum())},
            "note": "wall clock of the C-ABI call, pageable host buffers in / flags out (upload + bit-plane packing + scan + download); "
                    "kernel-only times in profiles/r01i_nbsorb_scans.txt; integer-ALU bound (2 LOP3 + 1 POPC per rejected pair), "
                    f"measured ALU issue peak {peak_ops / 1e12:.2f} T thread-instr/s"}


def file_e2e(ctx, np, R, length, seq_buf, qual_buf, first):
    """The same batch as a FASTQ file on local disk through catt_run_file - the call the patched Julia driver makes
    (INTEGRATION.md): parallel pread into pinned memory, text to HBM, indexing / validation / 2-bit packing on the device."""
    d = tempfile.mkdtemp(prefix="catt_bench_fq_")
    path = os.path.join(d, "synth.fq")
    w = 10
    rec = np.empty((R, 1 + 1 + w + 1 + length + 3 + length + 1), dtype=np.uint8)
    rec[:, 0] = ord("@"); rec[:, 1] = ord("s")
    rec[:, 2:2 + w] = np.frombuffer("".join(np.char.zfill((np.arange(R) + first).astype(str), w).tolist()).encode(), dtype=np.uint8).reshape(R, w)
    o = 2 + w
    rec[:, o] = 10
    rec[:, o + 1:o + 1 + length] = seq_buf.reshape(R, length)
    o += 1 + length
    rec[:, o] = 10; rec[:, o + 1] = ord("+"); rec[:, o + 2] = 10
    rec[:, o + 3:o + 3 + length] = qual_buf.reshape(R, length)
    rec[:, o + 3 + length] = 10
    rec.tofile(path)
    size = os.path.getsize(path)
    ctx.mainflow(path).close()
    ts = []
    for _ in range(3):
        t0 = time.perf_counter()
        res = ctx.mainflow(path)
        ts.append(time.perf_counter() - t0)
        res.close()
    os.remove(path)
    dt = sorted(ts)[1]
    return {"value": R / dt, "unit": "reads/s", "ms_per_call": dt * 1e3, "file_bytes": size,
            "note": "catt_run_file on a FASTQ in the page cache, median of 3 calls after 1 warm-up; QNAMEs are zero-padded read numbers"}


def dilute(np, seq_buf, qual_buf, lo, hi, length, tcr_fraction, seed):
    """RNA-seq-like mix (BASELINE configs[1] substitute): keep a fraction of the TCR-derived reads, replace the
    rest by uniform-random ACGT reads.  Decided per GLOBAL read index (1 Mi-read tiles, each with its own generator), so the sample
    is the same however it is split over GPUs.  seq_buf / qual_buf hold reads [lo, hi)."""
    if tcr_fraction >= 1.0 or hi <= lo:
        return
    tile = 1 << 20
    lut = np.frombuffer(b"ACGT", dtype=np.uint8)
    seqs, quals = seq_buf.reshape(hi - lo, length), qual_buf.reshape(hi - lo, length)
    for t0 in range(lo - lo % tile, hi, tile):
        rng = np.random.default_rng(seed + t0)
        noise = rng.random(tile) >= tcr_fraction
        bases = rng.integers(0, 4, size=(tile, length), dtype=np.uint8)
        a, b = max(t0, lo), min(t0 + tile, hi)
        sel = np.nonzero(noise[a - t0:b - t0])[0]
        seqs[(a - lo) + sel] = lut[bases[(a - t0) + sel]]
        quals[(a - lo) + sel] = ord("I")


def cpu_reference_run(fx, length, n_sample, first_index=0, species="hs", chain="TRB", tcr_fraction=1.0):
    """The CPU restatement (oracle) on n_sample reads with all host threads; returns (reads/s, seconds, info)."""
    import numpy as np
    from catt_b200.refg import chain_config
    from oracle import oracle as O
    cc = chain_config(species, chain, "CDR3", os.path.join(fx, "resource"))
    V, J = O.RefSet(cc["vregion"]), O.RefSet(cc["jregion"])
    seqs, quals = O.synth_reads(V, J, n_sample, length, first_index=first_index, seed=SEED)
    dilute(np, seqs, quals, first_index, first_index + n_sample, length, tcr_fraction, SEED)
    cfg = O.make_config(cc)
    t0 = time.perf_counter()
    R = O.run_flat(V, J, cfg, seqs, quals, first_index=first_index, threads=len(os.sched_getaffinity(0)))   # explicit: torchrun exports OMP_NUM_THREADS=1
    dt = time.perf_counter() - t0
    return n_sample / dt, dt, R.info


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--reads", "--reads-per-step", dest="reads", type=int, default=100_000_000,
                    help="reads of THE sample (BASELINE configs[2]: 100 M); every step runs the whole sample, split over the GPUs")
    ap.add_argument("--replica-reads", type=int, default=2_000_000, help="reads per GPU of the small side measurements (0 = skip)")
    ap.add_argument("--e2e-steps", type=int, default=5, help="timed host-buffer calls (at most --steps)")
    ap.add_argument("--length", type=int, default=150)
    ap.add_argument("--species", default="hs", help="other BASELINE configs: ms / pig (configs[4])")
    ap.add_argument("--chain", default="TRB", help="TRA / IGH / ... (configs[3] with --length 50)")
    ap.add_argument("--tcr-fraction", type=float, default=1.0, help="< 1: RNA-seq-like mix, the rest are random reads (configs[1] substitute)")
    ap.add_argument("--impl", default="gpu", choices=["gpu", "reference"])
    ap.add_argument("--cpu-sample", type=int, default=0, help="reads in the CPU-baseline sample (0 = auto, ~15 s)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-other-configs", action="store_true", help="skip the 1 M-read runs of the other BASELINE configs")
    ap.add_argument("--no-file-e2e", action="store_true", help="skip the catt_run_file measurement (FASTQ written to a temp file)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    cores = len(os.sched_getaffinity(0)) or 1
    workload = (f"one sample of {args.reads} synthetic {args.species} {args.chain} {args.length}bp reads from IMGT V/J with random CDR3 insertions "
                "(BASELINE configs[2], SURVEY 8d config 3)")
    if (args.species, args.chain, args.length, args.tcr_fraction) != ("hs", "TRB", 150, 1.0):
        workload = (f"one sample of {args.reads} synthetic {args.species} {args.chain} {args.length}bp reads, TCR-derived fraction {args.tcr_fraction} "
                    "(one of the other BASELINE configs; same generator, SURVEY 8d)")

    # ------------------------------------------------------------------------------------------------- reference arm
    if args.impl == "reference":
        if rank != 0:
            return 0
        fx = fixture_dir()
        import __graft_entry__ as g
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle")])
        n = args.cpu_sample or max(8000, 5000 * cores)         # ~5 s per step on `cores` threads
        for i in range(args.warmup):
            cpu_reference_run(fx, args.length, max(1000, n // 8), first_index=i * n, species=args.species, chain=args.chain, tcr_fraction=args.tcr_fraction)
        t_tot, cells = 0.0, 0
        for i in range(args.steps):
            rps, dt, info = cpu_reference_run(fx, args.length, n, first_index=i * n, species=args.species, chain=args.chain, tcr_fraction=args.tcr_fraction)
            t_tot += dt
            cells += info["cells_v"] + info["cells_j"]
        value = n * args.steps / t_tot
        line = {"impl": "reference", "metric": "reads_per_sec", "value": value, "unit": "reads/s", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_tot / args.steps, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "int16", "data": "synthetic",
                "gcups": cells / t_tot / 1e9,
                "config": {"workload": workload, "sample_reads": args.reads, "reads_per_step": n, "read_length": args.length, "chain": f"{args.species} {args.chain}",
                           "step": f"a bounded sample of the workload: {n} reads of the same generator per step (the CPU path at full size would take hours)",
                           "note": "CPU restatement of the reference path (oracle, OpenMP) - the Julia+bwa+samtools program cannot run here"},
                "cpu_baseline": {"value": value, "unit": "reads/s", "cores": cores, "kind": "port",
                                 "sample": f"{args.steps} x {n} reads of the same generator"},
                "e2e": {"value": value, "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return 0

    # ------------------------------------------------------------------------------------------------------- GPU arm
    import numpy as np
    import torch
    import torch.distributed as dist

    if not torch.cuda.is_available():
        print(json.dumps({"error": "no CUDA device: libcatt_b200 has no CPU fallback"}))
        return 2
    torch.cuda.set_device(local_rank)
    if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
        os.environ["NCCL_DEBUG"] = "WARN"          # keep NCCL's banner off stdout: rank 0 prints exactly one JSON line
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    from catt_b200 import _lib
    from catt_b200.host import Catt, comm_id, set_host_threads
    from catt_b200.refg import chain_config
    from catt_b200.shard import shard_bounds
    fx = fixture_dir()
    cc = chain_config(args.species, args.chain, "CDR3", os.path.join(fx, "resource"))
    # one process per GPU: every rank gets its share of the host cores (torchrun exports OMP_NUM_THREADS=1)
    local_world = int(os.environ.get("LOCAL_WORLD_SIZE", str(world)))
    host_threads = max(1, len(os.sched_getaffinity(0)) // max(1, local_world))
    set_host_threads(host_threads)
    # the production setting of INTEGRATION.md: the lists as catt() uses them (after catt.jl:581-582); e2e_full_lists below measures
    # the same call returning every Vpart/Jpart read
    ctx = Catt(chain_cfg=cc, device=local_rank, only_cdr3=True)
    if world > 1:                                  # ONE sample over the ranks: the library's own NCCL communicator (catt_ctx_join)
        idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idt.copy_(torch.frombuffer(bytearray(comm_id()), dtype=torch.uint8))
        dist.broadcast(idt, 0)
        ctx.join(world, rank, bytes(idt.cpu().numpy().tobytes()))
    stream = torch.cuda.ExternalStream(ctx.stream(), device=torch.device("cuda", local_rank))
    peak_ops, peak_ms = ctx.alu_peak()            # packed DPX add-max thread-instructions per second

    # ---- the sample: args.reads synthetic reads, read i drawn from splitmix64(seed + i); this rank generates ITS block
    n_total = args.reads
    lo, hi = shard_bounds(n_total, world, rank)
    n_loc = hi - lo
    t_gen = time.perf_counter()
    name_buf, name_off, seq_buf, seq_off, qual_buf = ctx.synth_reads(n_loc, args.length, first_index=lo, seed=SEED)
    dilute(np, seq_buf, qual_buf, lo, hi, args.length, args.tcr_fraction, SEED)
    t_gen = time.perf_counter() - t_gen

    def run_e2e():
        if world > 1:
            return ctx.mainflow_shard(n_total, lo, n_loc, name_buf, name_off, seq_buf, seq_off, qual_buf)
        return ctx.mainflow_buffers(n_total, name_buf, name_off, seq_buf, seq_off, qual_buf)

    # ---- value: the sample resident in HBM (packed reads, QNAMEs, quality strings); a step = the whole hot path over the whole sample
    ctx.shard_upload(n_total, lo, n_loc, name_buf, name_off, seq_buf, seq_off, qual_buf)

    def step_resident(keep=False):
        res = ctx.shard_run()
        info = res.info
        dig = res.digest() if keep else None
        res.close()
        return info, dig

    for _ in range(args.warmup):
        step_resident()
    sampler = ClockSampler(local_rank)
    sampler.start()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    ev0.record(stream)
    infos, digest = [], None
    for k in range(args.steps):
        inf, dg = step_resident(keep=(k == args.steps - 1))
        infos.append(inf)
        digest = dg or digest
    ev1.record(stream)
    barrier()
    wall_ms = (time.perf_counter() - t0) * 1e3
    ev_ms = ev0.elapsed_time(ev1)
    dev_ms = max(wall_ms, ev_ms)

    # ---- e2e: the same sample from HOST buffers through the public call (catt_run_reads / catt_run_reads_shard): staging copies,
    #      H2D, every kernel, the collectives, D2H of the lists - all inside the timed region
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    for _ in range(2):
        run_e2e().close()
    barrier()
    t0 = time.perf_counter()
    e2e_infos, e2e_digest = [], None
    for k in range(e2e_steps):
        res = run_e2e()
        e2e_infos.append(res.info)
        if k == e2e_steps - 1:
            e2e_digest = res.digest()
        res.close()
    barrier()
    e2e_ms = (time.perf_counter() - t0) * 1e3
    sampler.stop_flag = True
    sampler.join(timeout=2)

    t = torch.tensor([dev_ms, e2e_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms = float(t[0]), float(t[1])

    info = infos[-1]                                            # counters are sums over the ranks, stage times the slowest rank's
    ei = e2e_infos[-1]
    cells = info["cells_v"]                                     # the V reference set: its DP kernel is the one the roofline is about
    sw_ms = sum(i["ms_sw"] for i in infos) / len(infos)         # (the J set, ~3 % of the cells, runs concurrently on its own stream)
    achieved = 12.0 * cells / (sw_ms * 1e-3)                    # algorithmic int ops/s (12 per cell), all GPUs
    peak = peak_ops * 4.0 * world                               # a packed add-max retires 2 adds + 2 max per thread-instruction
    executed = 12.0 * info["cells_computed_v"] / (sw_ms * 1e-3)
    packed_bytes = n_loc * (4 * ((args.length + 15) // 16 + (args.length + 31) // 32) + 8)
    h2d, d2h = ei["bytes_h2d"], ei["bytes_d2h"]
    stage_keys = ("ms_seed", "ms_sw", "ms_select", "ms_order", "ms_extract", "ms_pool", "ms_strings", "ms_d2h", "ms_host", "ms_pack", "ms_j_stream",
                  "ms_comm", "ms_stage_a_max", "ms_stage_a_min")

    def coll_block(inf_list):
        ms = [sum(i["ms_coll"][k] for i in inf_list) / len(inf_list) for k in range(len(_lib.COLL_NAMES))]
        by = [inf_list[-1]["bytes_coll"][k] for k in range(len(_lib.COLL_NAMES))]
        rows = {nm: {"ms": ms[k], "bytes_received_all_ranks": by[k]} for k, nm in enumerate(_lib.COLL_NAMES)}
        top = max(range(len(ms)), key=lambda k: ms[k])
        return {"per_collective": rows, "limiting": _lib.COLL_NAMES[top] if world > 1 else None,
                "ms_total": sum(ms), "bytes_total": sum(by)}

    line = {
        "metric": "reads_per_sec", "value": n_total * args.steps / (dev_ms * 1e-3), "unit": "reads/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "int16", "data": "synthetic",
        "config": {"workload": workload, "sample_reads": n_total, "reads_per_gpu": n_loc, "read_length": args.length, "chain": f"{args.species} {args.chain}",
                   "step": "ONE sample of sample_reads reads through the whole hot path; with N GPUs the SAME sample is split into N contiguous blocks "
                           "(strong scaling) and the global steps - SAM order / QNAME dedup / V-J join, the last-writer-wins k-mer table, the result lists - "
                           "are merged over NCCL (catt_ctx_join + catt_shard_run / catt_run_reads_shard)",
                   "l2": f"packed block {packed_bytes >> 20} MiB per GPU and step vs 126 MB L2" + (" (larger than L2)" if packed_bytes > 126e6 else " (compute-bound kernels; see DESIGN.md)"),
                   "host_threads_per_rank": host_threads, "generate_s": t_gen,
                   "result": "only_cdr3 = 1: Vpart/Jpart as they stand after catt.jl:581-582 (the reads catt() goes on to use); counters of the full lists in info"},
        "gcups": cells / (sw_ms * 1e-3) / 1e9,
        "gcups_note": "DP cell updates of the V reference set, all GPUs / the slowest GPU's SW-kernel time (CUDA events on its stream, in the library)",
        "stage_ms": {k: sum(i[k] for i in infos) / len(infos) for k in stage_keys},
        "timing": {"wall_ms_per_step": wall_ms / args.steps, "cuda_event_ms_per_step": ev_ms / args.steps},
        "e2e_stage_ms": {k: sum(i[k] for i in e2e_infos) / len(e2e_infos) for k in stage_keys},
        "e2e": {"value": n_total * e2e_steps / (e2e_ms * 1e-3), "unit": "reads/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "ms_per_step": e2e_ms / e2e_steps, "steps": e2e_steps, "warmup": 2,
                "note": "catt_run_reads (1 GPU) / catt_run_reads_shard (N GPUs) on pageable host buffers; bytes summed over the GPUs"},
        "gpu_launches": int(sum(i["kernel_launches"] for i in infos)),
        "clocks": sampler.summary(),
        "one_sample": {"reads": n_total, "n_ranks": max(1, info["n_ranks"]), "ms_resident": dev_ms / args.steps, "ms_e2e": e2e_ms / e2e_steps,
                       "result_digest": {"vpart": f"{digest[0]:016x}", "jpart": f"{digest[1]:016x}", "e2e_vpart": f"{e2e_digest[0]:016x}", "e2e_jpart": f"{e2e_digest[1]:016x}",
                                         "note": "order-sensitive digest of every returned read (catt_result_digest): the same on 1, 2, 4 and 8 GPUs"},
                       "stage_a_ms_slowest_rank": info["ms_stage_a_max"], "stage_a_ms_fastest_rank": info["ms_stage_a_min"],
                       "collectives": coll_block(infos), "collectives_e2e": coll_block(e2e_infos),
                       "transport": "NCCL (library-owned communicator, ncclCommInitRank) over NVLink" if world > 1 else "none (one GPU)"},
        "roofline": {"bound": "alu", "kernel": "sw_bulk2_kernel<4,25> (V reference set; the J set goes through sw_score_kernel<4,8,TRACK> once per candidate)", "achieved": achieved / 1e12, "peak": peak / 1e12,
                     "unit": "Tiop/s", "frac": achieved / peak, "frac_executed": executed / peak, "traffic": ncu_traffic(),
                     "cells_algorithmic_v": int(cells), "cells_computed_v": int(info["cells_computed_v"]),
                     "alu_instr_per_cell_pair": 4.5, "ncu_pipe_alu_pct": ncu_pipe_alu(), "ncu_issue_active_pct": ncu_facts().get("sw_bulk2_kernel_issue_active_pct"),
                     "traffic_note": "dram bytes read + written by ONE launch of the ncu capture (" + str(ncu_facts().get("_source")) + "); a bench launch covers a 1 M-read chunk",
                     "note": "achieved = 12 int ops x ALGORITHMIC DP cells of the V reference set (every scored (read, record, strand) pair, len x len) / its SW kernel time; "
                             "frac_executed counts only the cells the kernel computes (byte-identical records are scored once); the kernel issues 4.5 ALU-pipe "
                             "instructions per cell pair where the 12-op model implies 6, so neither is pipe utilisation - ncu_pipe_alu_pct is (profiles/); "
                             "peak = measured packed-DPX (VIADDMNMX.S16x2) issue rate x 4 "
                             f"int16 ops per thread-instruction ({peak_ops / 1e12:.2f} T thread-instr/s per GPU, in-repo microbenchmark) x GPUs; "
                             "MEASURED_PEAKS.json has no integer entry"},
        "roofline_other": roofline_other(infos, args.length, world, n_total),
        "counters": {k: info[k] for k in ("v_hits", "j_hits", "v_final", "j_final", "share", "vpart", "jpart", "direct", "left", "rescued", "cand_v", "cand_j", "gapped_v")},
    }

    # ---- side measurements on a small batch (2 M reads per GPU): independent samples per GPU (the round-1 weak-scaling number),
    #      full lists, file input, the nbsorb scans, the CPU baseline
    R = min(args.replica_reads, n_loc)
    if R > 0:
        rb = ctx.synth_reads(R, args.length, first_index=rank * R, seed=SEED + 7)
        for _ in range(2):
            ctx.mainflow_buffers(R, *rb).close()
        barrier()
        t0 = time.perf_counter()
        for _ in range(3):
            ctx.mainflow_buffers(R, *rb).close()
        barrier()
        tt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        line["replicas"] = {"value": world * R * 3 / float(tt[0]), "unit": "reads/s", "reads_per_gpu_per_step": R, "steps": 3,
                            "note": "weak scaling: every GPU runs its OWN sample of reads_per_gpu_per_step reads (catt_run_reads from host buffers), no exchange"}
    if rank == 0 and world == 1 and R > 0:
        ctx_full = Catt(chain_cfg=cc, device=local_rank)
        for _ in range(2):
            ctx_full.mainflow_buffers(R, *rb).close()
        t0 = time.perf_counter()
        fi = None
        for _ in range(3):
            rr = ctx_full.mainflow_buffers(R, *rb)
            fi = rr.info
            rr.close()
        dtf = time.perf_counter() - t0
        line["e2e_full_lists"] = {"value": R * 3 / dtf, "unit": "reads/s", "ms_per_step": 1e3 * dtf / 3, "reads_per_step": R,
                                  "h2d_bytes_per_step": fi["bytes_h2d"], "d2h_bytes_per_step": fi["bytes_d2h"],
                                  "note": "only_cdr3 = 0: every Vpart/Jpart read comes back, cdr3 == None included (the lists before catt.jl:581)"}
        ctx_full.close()
    if rank == 0 and world == 1 and not args.no_file_e2e and R > 0:
        line["e2e_file"] = file_e2e(ctx, np, R, args.length, rb[2], rb[4], rank * R)
        line["nbsorb_scans"] = nbsorb_scans(ctx, np, peak_ops)
    if rank == 0 and world == 1 and not args.no_other_configs:
        # the other BASELINE configs (configs[3]: 50-bp reads of three chains; configs[4]: mouse / pig TRB) at 1 M reads each, one GPU,
        # host buffers in / lists out; parity of each against the CPU checker at this size: profiles/r02_parity_10M.txt
        oc = {}
        for sp, ch, ln in (("hs", "TRA", 50), ("hs", "TRB", 50), ("hs", "IGH", 50), ("hs", "IGH", 150), ("ms", "TRB", 150), ("pig", "TRB", 150)):
            cx = Catt(chain_cfg=chain_config(sp, ch, "CDR3", os.path.join(fx, "resource")), device=local_rank, only_cdr3=True)
            nn = 1_000_000
            b = cx.synth_reads(nn, ln, seed=SEED)
            for _ in range(2):
                cx.mainflow_buffers(nn, *b).close()
            t0 = time.perf_counter()
            ii = []
            for _ in range(3):
                r_ = cx.mainflow_buffers(nn, *b)
                ii.append(r_.info)
                r_.close()
            dt = (time.perf_counter() - t0) / 3
            msw = sum(i["ms_sw"] for i in ii) / 3
            oc[f"{sp} {ch} {ln}bp"] = {"reads_per_s_e2e": nn / dt, "ms_per_call": dt * 1e3, "gcups": ii[-1]["cells_v"] / (msw * 1e-3) / 1e9,
                                      "dp_frac_algorithmic": 12.0 * ii[-1]["cells_v"] / (msw * 1e-3) / (peak_ops * 4.0),
                                      "dp_frac_executed": 12.0 * ii[-1]["cells_computed_v"] / (msw * 1e-3) / (peak_ops * 4.0),
                                      "stage_ms": {k: sum(i[k] for i in ii) / 3 for k in ("ms_seed", "ms_sw", "ms_select", "ms_order", "ms_extract", "ms_pool")},
                                      "cand_per_read": ii[-1]["cand_v"] / nn, "kmer": ii[-1]["kmer"], "left": ii[-1]["left"], "rescued": ii[-1]["rescued"]}
            cx.close()
        line["other_configs"] = {"reads_per_call": 1_000_000, "what": "catt_run_reads from host buffers, 3 calls after 2 warm-ups, one GPU", "configs": oc}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle")])
        n = args.cpu_sample or max(16000, 15000 * cores)       # ~15 s of CPU work on `cores` threads
        rps, dt, cinfo = cpu_reference_run(fx, args.length, n, species=args.species, chain=args.chain, tcr_fraction=args.tcr_fraction)
        line["cpu_baseline"] = {"value": rps, "unit": "reads/s", "cores": cores, "kind": "port",
                                "sample": f"first {n} reads of the same generator, {dt:.1f} s; CPU restatement (oracle, OpenMP, scalar full-matrix SW at -O2, "
                                          "~0.27 GCUPS per core), not the Julia + bwa + samtools program",
                                "gcups": (cinfo["cells_v"] + cinfo["cells_j"]) / dt / 1e9}
    if rank == 0:
        print(json.dumps(line))
    ctx.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
