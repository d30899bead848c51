import os, torch, torch.distributed as td
import torch.distributed._symmetric_memory as sm
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
td.init_process_group("nccl", device_id=dev)
t = sm.empty(1024, dtype=torch.float32, device=dev)
t.fill_(rank)
h = sm.rendezvous(t, td.group.WORLD)
print(rank, "rendezvous ok", h.world_size, [hex(p) for p in h.buffer_ptrs], "multicast", h.has_multicast_support(dev.type, local) if hasattr(h, "has_multicast_support") else None, flush=True)
h.barrier()
peer = h.get_buffer((rank + 1) % world, (1024,), torch.float32)
peer[:4] = 100 + rank           # P2P store into the next rank's buffer
torch.cuda.synchronize(); h.barrier(); torch.cuda.synchronize()
print(rank, "my buffer head after peer write:", t[:6].tolist(), flush=True)
td.destroy_process_group()
