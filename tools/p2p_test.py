"""2..8-rank check of the peer-memory exchange kernels (run under torchrun):
  * vn_p2p_allreduce against NCCL;
  * vn_p2p_allreduce_small (mailbox sum of the loss normalisers) against NCCL;
  * vn_p2p_step (the same as ONE kernel, inf flag as input, step count on the device) against the same reference;
  * vn_p2p_reduce_adam (reduce-scatter + inf check + sharded Adam + parameter push + scaler
    update) against vn_p2p_allreduce + vn_grad_check + vn_adam_step + vn_scaler_update: the
    parameters must be BIT-identical on every rank, incl. a step with an injected inf.
Prints one line per check on rank 0; exits non-zero on a mismatch."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from virus_nerf_b200 import _lib  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device(f"cuda:{local}")
dist.init_process_group("nccl", device_id=dev)
n = 11429472
gen = torch.Generator(device=dev).manual_seed(rank)
grad = torch.randn(n, device=dev, generator=gen)
params = torch.randn(n, device=dev, generator=torch.Generator(device=dev).manual_seed(99))   # same on every rank
flags = torch.zeros(world, dtype=torch.int32, device=dev)
err = torch.zeros(1, dtype=torch.int32, device=dev)
mbox = torch.zeros(2, world, 8, device=dev)
keep = _lib.p2p_setup(grad, flags, err, rank, world, params=params, mbox=mbox)


def say(msg):
    if rank == 0:
        print(msg, flush=True)


def identical_everywhere(t):
    lo, hi = t.clone(), t.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    return bool(torch.equal(lo, hi))


def timeit(fn, reps=20):
    # the in-place allreduce grows the buffer by `world` per call: start every series small enough that 23 calls stay
    # finite (an overflow would set found_inf and make the optimiser kernels skip their work)
    grad.normal_().mul_(1e-30)
    for _ in range(3):
        fn()
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


ok = True
# ---- 1. big allreduce ------------------------------------------------------------------------
ref = grad.clone()
dist.all_reduce(ref)
_lib.call("vn_p2p_allreduce", n)
torch.cuda.synchronize()
diff = (grad - ref).abs().max().item()
same = identical_everywhere(grad)
ok &= same and diff < 1e-3 and int(err) == 0
say(f"world {world}: allreduce max |p2p - nccl| = {diff:.3e}, replicas identical = {same}")

# ---- 2. small allreduce ----------------------------------------------------------------------
for it in range(5):
    small = torch.arange(4, device=dev, dtype=torch.float32) * (rank + 1) + it
    ref_s = small.clone(); dist.all_reduce(ref_s)
    _lib.call("vn_p2p_allreduce_small", small, 4, 0)
    torch.cuda.synchronize()
    ok &= bool(torch.equal(small, ref_s)) and int(err) == 0
say(f"world {world}: small allreduce matches nccl = {ok}")

# ---- 3. fused sharded optimiser vs allreduce + dense Adam -------------------------------------
lr, b1, b2, eps = 1e-2, 0.9, 0.999, 1e-15
m_f, v_f = torch.zeros(n, device=dev), torch.zeros(n, device=dev)
m_r, v_r = torch.zeros(n, device=dev), torch.zeros(n, device=dev)
p_ref = params.clone()
found_f, found_r = torch.zeros(1, device=dev), torch.zeros(1, device=dev)
scale_f, scale_r = torch.tensor([2.0 ** 19], device=dev), torch.tensor([2.0 ** 19], device=dev)
tr_f, tr_r = torch.zeros(1, dtype=torch.int32, device=dev), torch.zeros(1, dtype=torch.int32, device=dev)
chunk = ((n // 4 + world - 1) // world) * 4
lo, hi = rank * chunk, min((rank + 1) * chunk, n)
for step in range(1, 5):
    g0 = torch.randn(n, device=dev, generator=gen) * 2.0 ** 19
    if step == 3 and rank == world - 1:
        g0[12345] = float("inf")                      # GradScaler must skip this step on EVERY rank
    # reference: allreduce (same summation order as the fused reduce-scatter) + dense optimiser
    grad.copy_(g0)
    _lib.call("vn_p2p_allreduce", n)
    _lib.call("vn_grad_check", grad, n, found_r)
    _lib.call("vn_adam_step", p_ref, grad, m_r, v_r, n, 1.0, lr, b1, b2, eps, step, found_r, scale_r)
    _lib.call("vn_scaler_update", scale_r, tr_r, found_r, 2.0, 0.5, 2000)
    # fused
    grad.copy_(g0)
    torch.cuda.synchronize(); dist.barrier()
    _lib.call("vn_p2p_reduce_adam", n, m_f, v_f, lr, b1, b2, eps, step, found_f, scale_f, tr_f)
    torch.cuda.synchronize()
    eq_p = bool(torch.equal(params, p_ref))
    eq_m = bool(torch.equal(m_f[lo:hi], m_r[lo:hi]) and torch.equal(v_f[lo:hi], v_r[lo:hi]))
    eq_s = float(scale_f) == float(scale_r) and int(tr_f) == int(tr_r)
    same = identical_everywhere(params)
    ok &= eq_p and eq_m and eq_s and same and int(err) == 0
    say(f"world {world}: fused optimiser step {step}: params == reference {eq_p}, own m/v slice {eq_m}, scaler {eq_s} "
        f"(scale {float(scale_f):.0f}), replicas identical {same}")

# ---- 3b. the single-kernel step (vn_p2p_step) vs allreduce + vn_grad_check + vn_adam_step_dev + vn_scaler_update_dev ----
m_s, v_s = torch.zeros(n, device=dev), torch.zeros(n, device=dev)
m_q, v_q = torch.zeros(n, device=dev), torch.zeros(n, device=dev)
p_ref = params.clone()
found_s, found_q = torch.zeros(1, device=dev), torch.zeros(1, device=dev)
scale_s, scale_q = torch.tensor([2.0 ** 19], device=dev), torch.tensor([2.0 ** 19], device=dev)
tr_s, tr_q = torch.zeros(1, dtype=torch.int32, device=dev), torch.zeros(1, dtype=torch.int32, device=dev)
st_s, st_q = torch.zeros(4, device=dev), torch.zeros(4, device=dev)
_lib.call("vn_opt_state_init", st_s, 0, lr, b1, b2); _lib.call("vn_opt_state_init", st_q, 0, lr, b1, b2)
for step in range(1, 6):
    g0 = torch.randn(n, device=dev, generator=gen) * 2.0 ** 19
    if step == 2 and rank == world - 1:
        g0[777] = float("nan")                        # one rank overflows: every rank must skip, Adam's count must not advance
    grad.copy_(g0)
    _lib.call("vn_p2p_allreduce", n)
    _lib.call("vn_grad_check", grad, n, found_q)
    _lib.call("vn_adam_step_dev", p_ref, grad, m_q, v_q, n, lr, b1, b2, eps, st_q, found_q, scale_q)
    _lib.call("vn_scaler_update_dev", scale_q, tr_q, found_q, 2.0, 0.5, 2000, st_q, lr, b1, b2)
    grad.copy_(g0)
    torch.cuda.synchronize(); dist.barrier()
    _lib.call("vn_grad_check", grad, n, found_s)       # the rank's OWN check is the input of vn_p2p_step
    _lib.call("vn_p2p_step", n, m_s, v_s, lr, b1, b2, eps, st_s, found_s, scale_s, tr_s)
    torch.cuda.synchronize()
    eq_p = bool(torch.equal(params, p_ref))
    eq_m = bool(torch.equal(m_s[lo:hi], m_q[lo:hi]) and torch.equal(v_s[lo:hi], v_q[lo:hi]))
    eq_s = float(scale_s) == float(scale_q) and int(tr_s) == int(tr_q) and bool(torch.equal(st_s[:3], st_q[:3])) \
        and float(found_s) == 0.0
    same = identical_everywhere(params)
    ok &= eq_p and eq_m and eq_s and same and int(err) == 0
    say(f"world {world}: one-kernel step {step}: params == reference {eq_p}, own m/v slice {eq_m}, scaler + step count {eq_s} "
        f"(scale {float(scale_s):.0f}, applied {int(st_s[2:3].view(torch.int32))}), replicas identical {same}")

# ---- 4. timings --------------------------------------------------------------------------------
grad.normal_()
t_p2p = timeit(lambda: _lib.call("vn_p2p_allreduce", n))
t_nccl = timeit(lambda: dist.all_reduce(ref))


def unfused():
    _lib.call("vn_p2p_allreduce", n)
    _lib.call("vn_grad_check", grad, n, found_r)
    _lib.call("vn_adam_step", p_ref, grad, m_r, v_r, n, 1.0, lr, b1, b2, eps, 5, found_r, scale_r)
    _lib.call("vn_scaler_update", scale_r, tr_r, found_r, 2.0, 0.5, 2000)


def nccl_unfused():
    dist.all_reduce(ref)
    _lib.call("vn_grad_check", ref, n, found_r)
    _lib.call("vn_adam_step", p_ref, ref, m_r, v_r, n, 1.0, lr, b1, b2, eps, 5, found_r, scale_r)
    _lib.call("vn_scaler_update", scale_r, tr_r, found_r, 2.0, 0.5, 2000)


t_unf = timeit(unfused)
t_nccl_unf = timeit(nccl_unfused)
t_fused = timeit(lambda: _lib.call("vn_p2p_reduce_adam", n, m_f, v_f, lr, b1, b2, eps, 5, found_f, scale_f, tr_f))
t_step = timeit(lambda: _lib.call("vn_p2p_step", n, m_s, v_s, lr, b1, b2, eps, st_s, found_s, scale_s, tr_s))
small = torch.ones(4, device=dev)
t_small = timeit(lambda: _lib.call("vn_p2p_allreduce_small", small, 4, 0))
small2 = torch.ones(4, device=dev)
t_small_nccl = timeit(lambda: dist.all_reduce(small2))
ok &= int(err) == 0
say(f"world {world}: allreduce p2p {t_p2p:.4f} ms, nccl {t_nccl:.4f} ms | exchange + optimiser: nccl+dense {t_nccl_unf:.4f} ms, "
    f"p2p+dense {t_unf:.4f} ms, FUSED sharded {t_fused:.4f} ms, ONE-KERNEL sharded {t_step:.4f} ms | 4-float allreduce: mailbox {t_small:.4f} ms, nccl {t_small_nccl:.4f} ms")
# ---- 5. the same step through NVLink multicast (NVLS): symmetric-memory buffers, multimem.ld_reduce / multimem.st ----------
try:
    _lib.p2p_shutdown()
    g2, g_ptrs, mc_g, hg = _lib.symmetric_empty(n, dev)
    p2, p_ptrs, mc_p, hp = _lib.symmetric_empty(n, dev)
    have_nvls = bool(mc_g and mc_p)
except Exception as e:      # no symmetric memory here
    have_nvls = False
    say(f"world {world}: symmetric memory unavailable ({e!r}); NVLS step not tested")
if have_nvls:
    flags2 = torch.zeros(world, dtype=torch.int32, device=dev); err2 = torch.zeros(1, dtype=torch.int32, device=dev)
    mbox2 = torch.zeros(2, world, 8, device=dev)
    keep2 = _lib.p2p_setup_symmetric(g_ptrs, p_ptrs, mc_g, mc_p, flags2, err2, mbox2, rank, world)
    p2.copy_(params)
    p_ref = params.clone()
    m_s, v_s, m_q, v_q = (torch.zeros(n, device=dev) for _ in range(4))
    found_s.zero_(); found_q.zero_(); scale_s.fill_(2.0 ** 19); scale_q.fill_(2.0 ** 19); tr_s.zero_(); tr_q.zero_()
    _lib.call("vn_opt_state_init", st_s, 0, lr, b1, b2); _lib.call("vn_opt_state_init", st_q, 0, lr, b1, b2)
    for step in range(1, 5):
        g0 = torch.randn(n, device=dev, generator=gen) * 2.0 ** 19
        if step == 2 and rank == 0:
            g0[4321] = float("inf")
        ref_g = g0.clone(); dist.all_reduce(ref_g)               # NCCL sum: the reference up to the order of the additions
        _lib.call("vn_grad_check", ref_g, n, found_q)
        _lib.call("vn_adam_step_dev", p_ref, ref_g, m_q, v_q, n, lr, b1, b2, eps, st_q, found_q, scale_q)
        _lib.call("vn_scaler_update_dev", scale_q, tr_q, found_q, 2.0, 0.5, 2000, st_q, lr, b1, b2)
        g2.copy_(g0)
        torch.cuda.synchronize(); dist.barrier()
        _lib.call("vn_grad_check", g2, n, found_s)
        _lib.call("vn_p2p_step", n, m_s, v_s, lr, b1, b2, eps, st_s, found_s, scale_s, tr_s)
        torch.cuda.synchronize()
        # Adam's update is lr * m / (sqrt(v) + eps): a last-bit difference of the gradient sum moves a parameter by
        # far less than lr; compare at 1e-3 * lr absolute
        dmax = float((p2 - p_ref).abs().max())
        eq_s = float(scale_s) == float(scale_q) and bool(torch.equal(st_s[:3], st_q[:3]))
        same = identical_everywhere(p2)
        good = dmax <= 1e-3 * lr and eq_s and same and int(err2) == 0
        ok &= good
        say(f"world {world}: NVLS step {step}: max |params - (nccl allreduce + dense Adam)| = {dmax:.3e}, scaler + step count {eq_s} "
            f"(scale {float(scale_s):.0f}, applied {int(st_s[2:3].view(torch.int32))}), replicas identical {same}")

    def nvls_step():
        _lib.call("vn_p2p_step", n, m_s, v_s, lr, b1, b2, eps, st_s, found_s, scale_s, tr_s)

    g2.normal_().mul_(1e-30)
    for _ in range(3):
        nvls_step()
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        nvls_step()
    e1.record(); torch.cuda.synchronize()
    ok &= int(err2) == 0
    say(f"world {world}: ONE-KERNEL sharded step through NVLS {e0.elapsed_time(e1) / 20:.4f} ms (peer loads / stores: {t_step:.4f} ms)")
flag = torch.tensor([1.0 if ok else 0.0], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
torch.cuda.synchronize()
_lib.p2p_shutdown()
dist.destroy_process_group()
if float(flag) != 1.0:
    raise SystemExit("p2p_test: MISMATCH")
