"""world_size-2 gloo tests (CPU) of the snapshot-parallel exchange logic in ctgcn_b200/dist.py."""
import os
import sys

import pytest
import torch
import torch.distributed as td
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, T, n, d, mode, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    td.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from ctgcn_b200 import dist
        g = torch.Generator().manual_seed(0)
        full = torch.randn(n, T, d, generator=g)                 # [N, T, D]: what a single process would stack
        owned = dist.owned_snapshots(T, world, rank)
        tl = (T + world - 1) // world
        local = torch.zeros(n, tl, d)
        for j, t in enumerate(owned):
            local[:, j] = full[:, t]
        seq = dist.exchange_to_node_slices(local, T, mode=mode)
        s, e = dist.node_slices(n, world)[rank]
        ok1 = torch.equal(seq, full[s:e])
        back = dist.gather_node_slices(seq * 2.0, n)
        ok2 = torch.equal(back, full * 2.0)
        q.put((rank, bool(ok1), bool(ok2), dist.world_size()))
    except Exception as exc:  # surface the failure instead of letting the parent time out
        q.put((rank, False, False, repr(exc)))
    finally:
        td.destroy_process_group()


@pytest.mark.parametrize("mode", ["all_to_all", "all_gather"])
@pytest.mark.parametrize("T,n", [(8, 10), (5, 7), (1, 4)])
def test_exchange_world2(mode, T, n):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() + T * 7 + n + (0 if mode == "all_to_all" else 50)) % 400
    procs = [ctx.Process(target=_worker, args=(r, 2, port, T, n, 6, mode, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, ok1, ok2, ws in res:
        assert ws == 2, ws
        assert ok1, f"rank {rank}: exchange result differs from the single-process stack ({mode})"
        assert ok2, f"rank {rank}: gathered output differs"
