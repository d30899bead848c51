"""Launched under torch.distributed.run by tests/test_dist_gpu.py (and usable by hand):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/dist_check.py
Every rank computes the snapshot-parallel CTGCN forward; rank 0 also computes the single-GPU forward of the same model
and inputs and checks agreement (same kernels, same per-row arithmetic → identical results)."""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as td

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def log(rank, msg):
    print(f"[rank {rank} +{time.time() - T0:6.1f}s] {msg}", flush=True)


T0 = time.time()


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    td.init_process_group("nccl", device_id=dev)
    log(rank, "process group up")
    t = torch.ones(4, device=dev)
    td.all_reduce(t)
    torch.cuda.synchronize()
    log(rank, f"all_reduce ok {t[0].item()}")
    import ctgcn_b200 as pkg
    from ctgcn_b200 import dist, synth
    from oracle import cases

    n, d, T, K = 5003, 128, 5, 4            # n and T not divisible by the world size
    snaps = [synth.make_snapshot("er", n, 30000, K, seed=t) for t in range(T)]
    sd = cases.ctgcn_params(np.random.default_rng(0), d, d, d, 1, 1, T, "S")
    ok = True
    for exchange in ("all_to_all", "all_gather", "p2p"):
        model = pkg.CTGCN(d, d, d, 1, 1, T, model_type="S").to(dev)
        model.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
        model.exchange = exchange
        model.snapshot_parallel = True
        xs = [synth.features(n, d, 1000 + t).to(dev) for t in range(T)]
        plans = [s.plan(dev) for s in snaps]
        with torch.no_grad():
            out, trans = model(xs, plans)                      # sharded (world > 1), gathered output
            model.gather_output = False
            out_slice, _ = model(xs, plans)
        torch.cuda.synchronize()
        log(rank, f"{exchange}: sharded forward done")
        assert tuple(out.shape) == (T, n, d)
        s, e = dist.node_slices(n, world)[rank]
        ok &= torch.equal(out[:, s:e], out_slice)
        owned = dist.owned_snapshots(T, world, rank)
        ok &= all((trans[t] is not None) == (t in owned) for t in range(T))
        if rank == 0:
            model.snapshot_parallel = False                   # single-process reference path
            with torch.no_grad():
                ref, ref_trans = model(xs, plans)
            maxdiff = (ref - out).abs().max().item()
            print(f"[{exchange}] sharded vs single-GPU: equal={torch.equal(ref, out)} max|diff|={maxdiff:.3e}", flush=True)
            ok &= maxdiff <= 1e-6
            ok &= all(torch.equal(ref_trans[t], trans[t]) for t in owned)
    flag = torch.tensor([1 if ok else 0], device=dev)
    td.all_reduce(flag, op=td.ReduceOp.MIN)
    torch.cuda.synchronize()
    td.destroy_process_group()
    if flag.item() != 1:
        raise SystemExit("dist_check FAILED")
    if rank == 0:
        print("dist_check OK", flush=True)


if __name__ == "__main__":
    main()
