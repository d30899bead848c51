#!/usr/bin/env python
"""Host ↔ device copy ceiling of the box, all GPUs at once (one process per GPU, torchrun) — what bounds the e2e line at N GPUs.

    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 profiles/host_copy_ceiling.py [--mb 512]

Every rank pins `mb` MiB, then (all ranks together, barrier before each phase) times H2D alone, D2H alone and both at once on two
streams, with CUDA events, 5 repetitions after a warm-up; rank 0 prints per-GPU and aggregate GB/s per direction, plus the NUMA
node / CPU list every rank ran on (hostmem.bind_host_to_gpu, which is what bench.py does before it pins)."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as td  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mb", type=int, default=512)
    ap.add_argument("--no-numa-bind", action="store_true")
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        td.init_process_group("nccl", device_id=dev)
    from ctgcn_b200 import hostmem
    numa = None if args.no_numa_bind else hostmem.bind_host_to_gpu(local, cpus=world > 1, memory=True)
    n = args.mb << 20
    h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
    h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
    h_in.fill_(1)
    d_a = torch.empty(n, dtype=torch.uint8, device=dev)
    d_b = torch.ones(n, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)

    def phase(up, down, reps=5):
        def once():
            if up:
                with torch.cuda.stream(s1):
                    d_a.copy_(h_in, non_blocking=True)
            if down:
                with torch.cuda.stream(s2):
                    h_out.copy_(d_b, non_blocking=True)
        once()
        torch.cuda.synchronize()
        if world > 1:
            td.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        s1.wait_event(e0)
        s2.wait_event(e0)
        for _ in range(reps):
            once()
        cur = torch.cuda.current_stream()
        cur.wait_stream(s1)
        cur.wait_stream(s2)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            g = [torch.zeros_like(t) for _ in range(world)]
            td.all_gather(g, t)
            return [float(x.item()) for x in g]
        return [ms]

    res = {"h2d_only": phase(True, False), "d2h_only": phase(False, True), "both": phase(True, True)}
    info = [None] * world
    if world > 1:
        td.all_gather_object(info, numa)
    else:
        info = [numa]
    if rank == 0:
        gb = n / 1e9
        out = {"mb_per_direction_per_gpu": args.mb, "gpus": world, "numa": info}
        for k, v in res.items():
            worst = max(v)
            out[k] = {"per_gpu_GBps": [round(gb / (x * 1e-3), 1) for x in v], "aggregate_GBps_per_direction": round(world * gb / (worst * 1e-3), 1)}
        print(json.dumps(out))
    if world > 1:
        td.destroy_process_group()


if __name__ == "__main__":
    main()
