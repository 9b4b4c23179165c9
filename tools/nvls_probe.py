"""Probe (run under torchrun): can this box hand out NVLink-multicast (NVLS) addresses through torch's symmetric memory?
Prints, per rank, whether the rendezvous works, the peer pointers and the multicast pointer."""
import os
import sys

import torch
import torch.distributed as dist

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device(f"cuda:{local}")
dist.init_process_group("nccl", device_id=dev)
import torch.distributed._symmetric_memory as symm_mem  # noqa: E402

try:
    print(rank, "backend", symm_mem.get_backend(dev), flush=True)
except Exception as e:
    print(rank, "get_backend failed", repr(e), flush=True)
try:
    t = symm_mem.empty(1 << 20, dtype=torch.float32, device=dev)
    t.fill_(rank + 1)
    hdl = symm_mem.rendezvous(t, dist.group.WORLD)
    print(rank, "rendezvous ok: world", hdl.world_size, "multicast support", hdl.has_multicast_support, "multicast_ptr",
          hex(hdl.multicast_ptr), "buffer_ptrs", [hex(p) for p in hdl.buffer_ptrs], "local ptr", hex(t.data_ptr()), flush=True)
    hdl.barrier()
    peer = hdl.get_buffer((rank + 1) % world, (16,), torch.float32)
    print(rank, "peer value", float(peer[0]), flush=True)
    hdl.barrier()
except Exception as e:
    print(rank, "symmetric memory failed:", repr(e)[:500], flush=True)
dist.barrier()
dist.destroy_process_group()
