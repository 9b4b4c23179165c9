"""2..8-rank check of the peer-memory allreduce against NCCL (run under torchrun)."""
import os, sys, time
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from virus_nerf_b200 import _lib

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device(f"cuda:{local}")
dist.init_process_group("nccl", device_id=dev)
n = 11429472
g = torch.Generator(device=dev).manual_seed(rank)
grad = torch.randn(n, device=dev, generator=g)
ref = grad.clone()
flags = torch.zeros(world, dtype=torch.int32, device=dev)
err = torch.zeros(1, dtype=torch.int32, device=dev)
keep = _lib.p2p_setup(grad, flags, err, rank, world)
dist.all_reduce(ref)
_lib.call("vn_p2p_allreduce", n)
torch.cuda.synchronize()
assert int(err) == 0, "barrier timeout"
diff = (grad - ref).abs().max().item()
# bit-identical across ranks?
lo, hi = grad.clone(), grad.clone()
dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
same = bool(torch.equal(lo, hi))
def timeit(fn, reps=20):
    for _ in range(3): fn()
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
t_p2p = timeit(lambda: _lib.call("vn_p2p_allreduce", n))
t_nccl = timeit(lambda: dist.all_reduce(ref))
assert int(err) == 0
if rank == 0:
    print(f"world {world}: max |p2p - nccl| = {diff:.3e}, replicas identical = {same}, p2p {t_p2p:.4f} ms, nccl {t_nccl:.4f} ms", flush=True)
dist.destroy_process_group()
