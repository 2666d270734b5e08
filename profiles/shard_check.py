#!/usr/bin/env python3
"""One sample over N GPUs (torchrun): Stage A sharded, records all-gathered over NCCL between device buffers, Stage B on rank 0;
checks the result against the unsharded call on rank 0 and prints timings.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P profiles/shard_check.py [reads]
"""
import os, sys, tarfile, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist
from catt_b200.host import Catt, set_host_threads
from catt_b200.refg import chain_config
from catt_b200.shard import ShardedCatt

n = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 2000000
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
set_host_threads(max(1, len(os.sched_getaffinity(0)) // world))
d = tempfile.mkdtemp(prefix="catt_shard_")
with tarfile.open(os.path.join(ROOT, "tests", "golden", "catt_fixture.tar.gz")) as tar:
    tar.extractall(d)
cc = chain_config("hs", "TRB", "CDR3", os.path.join(d, "resource"))
ctx_full = Catt(chain_cfg=cc, device=local)
ctx_cdr3 = Catt(chain_cfg=cc, device=local, only_cdr3=True)      # what catt.jl's default output needs (no full read lists)
bufs = ctx_full.synth_reads(n, 150)                  # every rank generates the same whole sample
modes = ("host", "device", "device+only_cdr3") if "--all" in sys.argv else ("device", "device+only_cdr3")
for mode in modes:
    ctx = ctx_cdr3 if mode.endswith("only_cdr3") else ctx_full
    sc = ShardedCatt(ctx, device=dev, on_device=mode.startswith("device"))
    for rep in range(3):
        dist.barrier(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        res = sc.mainflow_buffers(n, *bufs)
        dist.barrier(); torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if rank == 0 and rep == 2:
            if sc.laps_ms:
                print(f"[{mode}] rank 0 laps: " + ", ".join(f"{k} {v:.1f} ms" for k, v in sc.laps_ms.items()), flush=True)
            whole_t = []
            for _ in range(2):
                t1 = time.perf_counter(); w2 = ctx.mainflow_buffers(n, *bufs); whole_t.append(time.perf_counter() - t1); w2.close()
            print(f"[1 GPU, {mode}] the same sample in one unsharded call: {min(whole_t) * 1e3:.1f} ms", flush=True)
            print(f"[{world} GPUs, records gathered through {mode}] one sample of {n} reads: {dt * 1e3:.1f} ms = {n / dt / 1e6:.2f} M reads/s", flush=True)
        if rank == 0 and rep == 0:
            whole = ctx.mainflow_buffers(n, *bufs)
            assert res.info["vpart"] == whole.info["vpart"] and res.info["jpart"] == whole.info["jpart"], (res.info["vpart"], whole.info["vpart"])
            import ctypes as C
            from catt_b200 import _lib
            # full comparison of every string through a sampled set of entries + all counters
            for which in (0, 1):
                pa, ka = C.POINTER(_lib.Read)(), C.c_int64()
                pb, kb = C.POINTER(_lib.Read)(), C.c_int64()
                _lib.check(_lib.load().catt_result_reads(res._h, which, C.byref(pa), C.byref(ka)))
                _lib.check(_lib.load().catt_result_reads(whole._h, which, C.byref(pb), C.byref(kb)))
                assert ka.value == kb.value
                step = max(1, ka.value // 50000)
                for i in range(0, ka.value, step):
                    x, y = pa[i], pb[i]
                    assert (C.string_at(x.seq, x.seq_len), C.string_at(x.qual, x.qual_len), x.v_id, x.j_id, x.cdr3_off, x.cdr3_len, x.ordinal) == \
                           (C.string_at(y.seq, y.seq_len), C.string_at(y.qual, y.qual_len), y.v_id, y.j_id, y.cdr3_off, y.cdr3_len, y.ordinal), i
            for k in ("kmer", "the_err", "v_hits", "j_hits", "v_final", "j_final", "share", "vpart", "jpart", "direct", "left", "rescued"):
                assert res.info[k] == whole.info[k], (k, res.info[k], whole.info[k])
            print(f"[{mode}] sharded == whole: Vpart {res.info['vpart']}, Jpart {res.info['jpart']}, rescued {res.info['rescued']}", flush=True)
            whole.close()
        if res is not None:
            res.close()
dist.destroy_process_group()
