"""Snapshot-parallel execution of CTGCN.forward over torch.distributed (one process per GPU, NCCL).

Sharding: snapshot t is owned by rank t mod G — MLP_t / CDN_t weights, x_t and the graph plan
of adj_list[t] are disjoint per snapshot (reference models.py:225-231, 243-247), so there is no data-path
collective until the stack at models.py:248.  There, ONE exchange moves the per-snapshot embeddings so that
every rank holds all T snapshots of ITS node slice ([N/G, T, D]); the temporal GRU + LayerNorm
(models.py:249-250) is node-wise independent and runs on that slice.  ``exchange='all_to_all'``
sends each rank only its slice; ``exchange='all_gather'`` is the literal all-gather of whole snapshots;
``exchange='p2p'`` fuses the exchange into the producing kernel's epilogue over NVLink peer memory (PeerExchange);
``'auto'`` (default) = p2p when symmetric memory is available, else all_to_all.
A second collective gathers the output only if ``model.gather_output`` is set.

The tensor-shuffling helpers below are device-agnostic so that the host logic is testable with gloo on CPU.
"""
from __future__ import annotations

import torch
import torch.distributed as td


def world_size() -> int:
    return td.get_world_size() if td.is_available() and td.is_initialized() else 1


def rank() -> int:
    return td.get_rank() if td.is_available() and td.is_initialized() else 0


def owned_snapshots(T: int, G: int, r: int):
    return list(range(r, T, G))


def node_slices(n: int, G: int):
    """Balanced contiguous node ranges [(start, stop)] for the G ranks."""
    base, rem = divmod(n, G)
    out, s = [], 0
    for g in range(G):
        e = s + base + (1 if g < rem else 0)
        out.append((s, e))
        s = e
    return out


def exchange_to_node_slices(hx_local: torch.Tensor, T: int, group=None, mode: str = "all_to_all") -> torch.Tensor:
    """hx_local [N, Tl, D]: this rank's snapshots (slot j ↔ snapshot rank + G·j, Tl = ceil(T/G), unused slots
    arbitrary) → seq [rows_of_this_rank, T, D] holding every snapshot of this rank's node slice."""
    G, r = td.get_world_size(group), td.get_rank(group)
    n, tl, d = hx_local.shape
    assert tl == (T + G - 1) // G, (tl, T, G)
    slices = node_slices(n, G)
    my0, my1 = slices[r]
    rows = my1 - my0
    if mode == "all_to_all":
        recv = hx_local.new_empty((rows * G, tl, d))
        td.all_to_all_single(recv, hx_local.contiguous(), output_split_sizes=[rows] * G,
                             input_split_sizes=[e - s for s, e in slices], group=group)
        parts = recv.view(G, rows, tl, d)
    elif mode == "all_gather":
        full = hx_local.new_empty((G * n, tl, d))
        td.all_gather_into_tensor(full, hx_local.contiguous(), group=group)
        parts = full.view(G, n, tl, d)[:, my0:my1]
    else:
        raise ValueError("exchange must be 'all_to_all' or 'all_gather'")
    # parts[g, n, j] is snapshot t = g + G·j  →  seq[n, t]
    seq = parts.permute(1, 2, 0, 3).reshape(rows, tl * G, d)
    return seq[:, :T].contiguous()


def gather_node_slices(out_slice: torch.Tensor, n: int, group=None) -> torch.Tensor:
    """out_slice [rows_r, T, D] on every rank → [N, T, D] on every rank."""
    G = td.get_world_size(group)
    slices = node_slices(n, G)
    max_rows = max(e - s for s, e in slices)
    _, T, d = out_slice.shape
    pad = out_slice.new_zeros((max_rows, T, d))
    pad[: out_slice.shape[0]] = out_slice
    full = out_slice.new_empty((G * max_rows, T, d))
    td.all_gather_into_tensor(full, pad, group=group)
    full = full.view(G, max_rows, T, d)
    return torch.cat([full[g, : e - s] for g, (s, e) in enumerate(slices)], dim=0)


class PeerExchange:
    """Snapshot exchange fused into the CoreDiffusion epilogue over NVLink peer memory.

    Every rank owns a symmetric buffer seq[rows_max, T, D] (torch symmetric memory: the same allocation is mapped into
    every peer's address space).  The last CoreDiffusion layer of snapshot t stores row r directly into the buffer of the
    rank that owns r's node slice (ctgcn_core_diffusion_fwd_scatter) — no staging copy, no NCCL call, no permute; the
    transfer overlaps the GRU math tile by tile.  Two device-side barriers on the stream bracket the writes."""

    _cache = {}

    def __init__(self, n, T, d, dev, group):
        import torch.distributed._symmetric_memory as symm
        self.n, self.T, self.d = n, T, d
        G = td.get_world_size(group)
        self.slices = node_slices(n, G)
        rows_max = max(e - s for s, e in self.slices)
        self.buf = symm.empty((rows_max, T, d), dtype=torch.float32, device=dev)
        self.handle = symm.rendezvous(self.buf, group if group is not None else td.group.WORLD)
        ptrs = [int(p) for p in self.handle.buffer_ptrs]
        self.ptrs = torch.tensor(ptrs, dtype=torch.int64, device=dev)
        s, e = self.slices[td.get_rank(group)]
        self.rows = e - s
        self.group = group
        self._flag = torch.zeros(1, dtype=torch.float32, device=dev)

    @classmethod
    def get(cls, n, T, d, dev, group=None):
        key = (n, T, d, str(dev), id(group))
        if key not in cls._cache:
            cls._cache[key] = cls(n, T, d, dev, group)
        return cls._cache[key]

    def scatter_args(self, t):
        return (self.ptrs, self.T * self.d, t * self.d)

    def barrier(self):
        # A 4-byte NCCL all-reduce on the compute stream: it starts after this rank's producing kernel has completed
        # (peer stores are visible system-wide at kernel completion) and completes only when every rank got there.
        td.all_reduce(self._flag, group=self.group)

    def local_seq(self):
        return self.buf[: self.rows]


_p2p_state = {}


def _p2p_available(dev) -> bool:
    """Symmetric (peer-mapped) memory works for this process group?  Probed once, collectively (every rank takes the same
    branch: the outcome is all-reduced), with a tiny allocation; any failure selects the NCCL all_to_all exchange."""
    key = str(dev)
    if key not in _p2p_state:
        ok = 1
        try:
            import torch.distributed._symmetric_memory as symm
            t = symm.empty((16,), dtype=torch.float32, device=dev)
            symm.rendezvous(t, td.group.WORLD).barrier()
        except Exception as exc:  # noqa: BLE001 - any failure means "use NCCL"
            import sys
            print(f"ctgcn_b200: peer-memory exchange unavailable ({exc!r}); using NCCL all_to_all", file=sys.stderr)
            ok = 0
        flag = torch.tensor([ok], device=dev)
        td.all_reduce(flag, op=td.ReduceOp.MIN)
        _p2p_state[key] = bool(flag.item())
    return _p2p_state[key]


def _forward_p2p(model, x_list, adj_list, owned, T, dev):
    from .models import _HostFeatureStager
    stager = _HostFeatureStager(x_list, owned, dev)
    trans_list = [None] * T
    ex, n = None, None
    for t in owned:
        trans = model.mlp_list[t](stager.get(t))
        trans_list[t] = trans
        if ex is None:
            n = trans.shape[0]
            ex = PeerExchange.get(n, T, model.output_dim, dev)
            ex.barrier()                      # every peer has finished reading the previous call's sequence buffer
        model.duffision_list[t].forward_into(trans, adj_list[t], scatter=ex.scatter_args(t))
    if ex is None:                            # this rank owns no snapshot but still takes part in the barriers
        n = int(getattr(model, "node_num", 0)) or _infer_rows(x_list, adj_list)
        ex = PeerExchange.get(n, T, model.output_dim, dev)
        ex.barrier()
    ex.barrier()                              # all peers' rows have landed
    return ex.local_seq(), n, trans_list


def ctgcn_forward_sharded(model, x_list, adj_list):
    """CTGCN.forward under snapshot parallelism.  x_list[t] / adj_list[t] are only touched for owned t
    (entries for other snapshots may be None).  Returns the reference's [T, N, D] view when
    ``model.gather_output`` else this rank's slice [T, rows, D]; model_type 'S' adds the list of MLP outputs
    (None for snapshots owned by other ranks)."""
    from .layers import _guard
    from .models import _HostFeatureStager

    G, r = world_size(), rank()
    T = len(x_list)
    owned = owned_snapshots(T, G, r)
    tl = (T + G - 1) // G
    dev = model.norm.weight.device
    mode = getattr(model, "exchange", "auto")
    if mode == "auto":
        mode = "p2p" if _p2p_available(dev) else "all_to_all"
    # The sharded forward is inference-only (the exchange has no backward yet): compute without autograd and mark the result
    # so that .backward() fails loudly instead of producing zero gradients.
    with torch.no_grad():
        if mode == "p2p":
            seq, n, trans_list = _forward_p2p(model, x_list, adj_list, owned, T, dev)
        else:
            trans_list = [None] * T
            hx_local, n = None, None
            stager = _HostFeatureStager(x_list, owned, dev)
            for j, t in enumerate(owned):
                trans = model.mlp_list[t](stager.get(t))
                trans_list[t] = trans
                if hx_local is None:
                    n = trans.shape[0]
                    hx_local = torch.zeros(n, tl, model.output_dim, dtype=torch.float32, device=dev)
                model.duffision_list[t].forward_into(trans, adj_list[t], out=hx_local[:, j, :])
            if hx_local is None:  # more ranks than snapshots: this rank owns nothing but still takes part
                n = int(getattr(model, "node_num", 0)) or _infer_rows(x_list, adj_list)
                hx_local = torch.zeros(n, tl, model.output_dim, dtype=torch.float32, device=dev)
            seq = exchange_to_node_slices(hx_local, T, mode=mode)
        if getattr(model, "keep_exchanged", False):     # bench / tests: the [rows, T, D] sequence this rank received (a view)
            model._exchanged = seq
        out = model._temporal(seq)
        if getattr(model, "gather_output", True):
            out = gather_node_slices(out, n)
    out = _guard(out, model).transpose(0, 1)
    return out if model.model_type == 'C' else (out, trans_list)


def _infer_rows(x_list, adj_list):
    for x in x_list:
        if x is not None:
            return x.shape[0]
    raise ValueError("cannot infer the node count: this rank owns no snapshot and x_list holds no tensor")
