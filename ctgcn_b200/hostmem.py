"""Host-side placement for the end-to-end path: keep a rank's pinned buffers on the NUMA node its GPU hangs off.

The e2e step uploads this rank's snapshot features and reads its node slice back over PCIe every step.  With one process per
GPU on a two-socket box, a process that runs (and first-touches its pinned pages) on the other socket pushes every byte
through the inter-socket link as well: 8 ranks measured 63 GB/s aggregate per direction where the four PCIe switches offer
about 200 (profiles/r01_bench.md, 8-GPU e2e).  `bind_host_to_gpu` restricts the calling process to the CPUs local to the GPU
and prefers that node for new pages — what `numactl --cpunodebind --preferred` would do from outside.  Everything is
best effort: a container without sysfs PCI nodes, a cpuset that excludes the node, or a non-Linux host leave the process
untouched and say so in the returned record.  Call it BEFORE allocating pinned memory.
"""
from __future__ import annotations

import ctypes
import os

_MPOL_PREFERRED = 1
_SYS_SET_MEMPOLICY = {"x86_64": 238, "aarch64": 237}


def parse_cpulist(text: str):
    """'0-3,8,10-11' → [0, 1, 2, 3, 8, 10, 11] (the kernel's cpulist format; empty string → [])."""
    cpus = []
    for part in text.strip().split(","):
        part = part.strip()
        if not part:
            continue
        if "-" in part:
            lo, hi = part.split("-", 1)
            cpus.extend(range(int(lo), int(hi) + 1))
        else:
            cpus.append(int(part))
    return sorted(set(cpus))


def gpu_sysfs_dir(device_index: int):
    """/sys/bus/pci/devices/<domain:bus:device.0> of a CUDA device, or None."""
    import torch
    try:
        p = torch.cuda.get_device_properties(device_index)
        name = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
    except (AttributeError, RuntimeError, AssertionError):
        return None
    path = os.path.join("/sys/bus/pci/devices", name)
    return path if os.path.isdir(path) else None


def _read(path):
    try:
        with open(path) as f:
            return f.read().strip()
    except OSError:
        return None


def _prefer_node(node: int) -> bool:
    machine = os.uname().machine
    nr = _SYS_SET_MEMPOLICY.get(machine)
    if nr is None or node < 0 or node >= 1024:
        return False
    try:
        libc = ctypes.CDLL(None, use_errno=True)
        words = node // 64 + 1
        mask = (ctypes.c_ulong * words)()
        mask[node // 64] = 1 << (node % 64)
        rc = libc.syscall(ctypes.c_long(nr), ctypes.c_int(_MPOL_PREFERRED), mask, ctypes.c_ulong(64 * words + 1))
        return rc == 0
    except (OSError, AttributeError, ValueError):
        return False


def reset_memory_policy() -> bool:
    """Back to the default page placement (MPOL_DEFAULT) for the calling thread."""
    nr = _SYS_SET_MEMPOLICY.get(os.uname().machine)
    if nr is None:
        return False
    try:
        libc = ctypes.CDLL(None, use_errno=True)
        return libc.syscall(ctypes.c_long(nr), ctypes.c_int(0), None, ctypes.c_ulong(0)) == 0
    except (OSError, AttributeError, ValueError):
        return False


def bind_host_to_gpu(device_index: int, cpus: bool = True, memory: bool = True) -> dict:
    """Pin the calling process to the CPUs of the GPU's NUMA node and prefer that node for new pages.  Never raises;
    returns {"node": int|None, "cpus": n_bound|None, "mem_preferred": bool, "note": str}."""
    rec = {"node": None, "cpus": None, "mem_preferred": False, "note": ""}
    sysdir = gpu_sysfs_dir(device_index)
    if sysdir is None:
        rec["note"] = "no sysfs PCI node for the device"
        return rec
    node = _read(os.path.join(sysdir, "numa_node"))
    try:
        rec["node"] = int(node) if node is not None else None
    except ValueError:
        rec["node"] = None
    if cpus and hasattr(os, "sched_setaffinity"):
        local = parse_cpulist(_read(os.path.join(sysdir, "local_cpulist")) or "")
        try:
            allowed = os.sched_getaffinity(0)
            keep = allowed.intersection(local)
            if keep and keep != allowed:
                os.sched_setaffinity(0, keep)
                rec["cpus"] = len(keep)
            elif keep:
                rec["cpus"] = len(keep)
                rec["note"] = "already on the local CPUs"
            else:
                rec["note"] = "the GPU's local CPUs are outside this process's cpuset"
        except OSError as exc:
            rec["note"] = f"sched_setaffinity: {exc}"
    if memory and rec["node"] is not None and rec["node"] >= 0:
        rec["mem_preferred"] = _prefer_node(rec["node"])
    return rec
