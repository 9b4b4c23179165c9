"""bench.py -- headline benchmark of the hot path (see DESIGN.md "Measurement").

  python bench.py --gpus N --steps K --warmup W            (this framework, N ranks via torchrun)
  python bench.py --impl reference --gpus N --steps K --warmup W   (CPU restatement of the reference)

One "step" = one full training step of the ETHZ-shaped workload (BASELINE.json configs[1]):
occupancy-grid update every 8 steps, ray/AABB, ray march over the occupancy bitfield, hash
encode, MLPs, composite, RGB+USS+ToF losses, backward, gradient allreduce (N>1), GradScaler
unscale + Adam.  4096 rays per GPU per step (weak scaling).  Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

RAYS_PER_GPU = 4096
HASH_BYTES_PER_POINT = 1164          # fwd or bwd, fp32 L16 F2 (SURVEY 8(d) / BASELINE.md section 3)
# the two candidates for "dominant kernel of the step" (DESIGN.md section 4): both are bracketed with CUDA events
# INSIDE the timed region and the one with the larger total there is the roofline kernel
ROOFLINE_CANDIDATES = ("hash_encode_bwd", "mlp_bwd", "mlp_bwd_hash_scatter")
# the fused MLP-backward + hash-scatter kernel (vn_mlp_bwd_scatter): the table-gradient read-modify-write of the hash
# backward (16 levels x 8 corners x 8 B = 1024) + xyz (12) + the MLP backward's inputs (fp16 encoding 64, SH planes 32,
# dsigma 4, drgb 12); the 128 B/point d(enc) round trip of the unfused pair no longer exists
FUSED_BWD_BYTES_PER_POINT = 1024 + 12 + 64 + 32 + 4 + 12
FUSED_BWD_BYTES_PER_POINT_HALF = FUSED_BWD_BYTES_PER_POINT      # the half encoder's backward scatters fp32 too
WORKLOAD = ("ETHZ-shaped synthetic scene, hash grid L=16 F=2 T=2^19 fp32 tables, 4096 rays/batch/GPU, RGB+USS+ToF "
            "losses, VIRUS-NeRF occupancy update every 8 steps, training from the initial (all-occupied) grid")


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=8)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--rays", type=int, default=RAYS_PER_GPU)
    ap.add_argument("--cpu-rays", type=int, default=RAYS_PER_GPU,
                    help="rays per step of the CPU arm (default: the full 4096-ray batch of the GPU arm's config)")
    ap.add_argument("--windows", type=int, default=7, help="repetitions of the K-step timed window (median reported)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: --rays per GPU; strong: --rays is ONE global batch, rank r takes its contiguous shard")
    ap.add_argument("--no-extra", action="store_true", help="skip extra_configs (configs 3 and 5)")
    ap.add_argument("--no-config3", action="store_true", help="skip the T=2^22 / 2^18-ray / half-encoder configuration")
    ap.add_argument("--no-fused-scatter", action="store_true", help="MLP backward and hash backward as two kernels (A/B)")
    ap.add_argument("--no-early-expand", action="store_true", help="sample expansion inside the step instead of ahead of it (A/B)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-autocast", action="store_true")
    ap.add_argument("--autograd-step", action="store_true", help="use the torch-autograd step instead of the fused step")
    ap.add_argument("--comm", default="auto", choices=["auto", "nccl", "p2p", "p2p_fused"],
                    help="gradient exchange: nccl allreduce, own peer-memory allreduce, or the sharded optimiser fused "
                         "with the peer-memory exchange (auto = p2p_fused, nccl if peer mapping fails)")
    ap.add_argument("--enc-layout", default="chunks", choices=["chunks", "planar", "rows"],
                    help="layout of the encoding inside the fused step (rows = the reference's [S,32])")
    ap.add_argument("--two-pass-march", action="store_true", help="re-march in the write pass (reference structure)")
    ap.add_argument("--no-e2e", action="store_true", help="skip the end-to-end phase (profiling runs only)")
    return ap.parse_args()


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def peaks_tensor():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["bf16_tflops_sustained"])
    except Exception:
        return 1400.0


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)"""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.proc, self.index = None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill(); out = ""
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ----------------------------------------------------------------------------------------
def run_reference(a):
    """CPU restatement of the reference's train step on the host cores (Taichi is not
    installable here: see oracle/oracle.cpp header).  Rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import numpy as np
    import torch
    import oracle
    from oracle import pipeline
    from virus_nerf_b200 import synthetic
    oracle.build()
    cores = len(os.sched_getaffinity(0))
    torch.set_num_threads(cores)
    args = synthetic.make_args(device="cpu")
    ds = synthetic.SyntheticDataset(pool_size=1 << 16, device="cpu")
    model = pipeline.OracleNGP(threads=cores)
    tr = pipeline.OracleTrainer(model, lr=args.training.lr)
    occ = pipeline.OracleOccupancyGrid(model, ds, args)
    rng = np.random.default_rng(0)
    n = a.cpu_rays
    times, samples = [], []
    for it in range(a.warmup + a.steps):
        t0 = time.perf_counter()
        if it % args.occ_grid.update_interval == 0:
            occ.update()
        data = ds(n, args.training.sampling_strategy)
        noise = rng.random(n, dtype=np.float32)
        _, res = tr.step(data, occ.bitfield, noise)
        dt = time.perf_counter() - t0
        if it >= a.warmup:
            times.append(dt); samples.append(res["rm_samples"])
    total = sum(times)
    value = n * a.steps / total
    sample = (f"{n} of the 4096 rays of each step ({a.steps} steps, grid update every 8; dense Adam over the full "
              f"11.4M-parameter table runs every step as in the reference)")
    line = {"impl": "reference", "metric": "train_rays_per_sec", "value": value, "unit": "rays/s", "n_gpus": a.gpus,
            "steps": a.steps, "warmup": a.warmup, "ms_per_step": 1e3 * total / a.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "rays_per_step": n, "samples_per_step": float(np.mean(samples)),
                       "reference_kind": "restated reference (Taichi ti.cpu unavailable): oracle.cpp OpenMP + torch-CPU fp32"},
            "cpu_baseline": {"value": value, "unit": "rays/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------
def _snapshot(eng):
    """everything a train step changes: parameters, Adam moments, scaler / optimiser state, occupancy grid, RNG streams"""
    import torch
    og = eng.model.occupancy_grid
    snap = {"t": [x.clone() for x in (eng.flat_p, eng.flat_m, eng.flat_v, eng.scale, eng.growth_tracker, eng.found_inf,
                                      eng.opt_state)],
            "grid": og.occ_3d_grid.clone(), "bf": og.getBitfield().clone(), "update_step": og.update_step,
            "ints": (eng.step_idx, eng._prep_step, eng._grid_updates, eng.adam_step),
            "rng": torch.cuda.get_rng_state(eng.device),
            "gen": og.dataset.gen.get_state() if getattr(og, "dataset", None) is not None else None}
    return snap


def _restore(eng, snap):
    import torch
    og = eng.model.occupancy_grid
    for dst, src in zip((eng.flat_p, eng.flat_m, eng.flat_v, eng.scale, eng.growth_tracker, eng.found_inf, eng.opt_state),
                        snap["t"]):
        dst.copy_(src)
    og.occ_3d_grid.copy_(snap["grid"]); og.bitfield = snap["bf"].clone(); og.update_step = snap["update_step"]
    eng.step_idx, eng._prep_step, eng._grid_updates, eng.adam_step = snap["ints"]
    eng._ticket = None                                   # it was prepared with the bitfield of the previous window
    torch.cuda.set_rng_state(snap["rng"], eng.device)
    if snap["gen"] is not None:
        og.dataset.gen.set_state(snap["gen"])


def run_ours(a):
    import torch
    import torch.distributed as dist
    from virus_nerf_b200 import _lib, synthetic
    from virus_nerf_b200.engine import TrainEngine

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- virus-nerf_b200 has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _lib.lib()
    strong = a.scaling == "strong"
    n_global = a.rays if strong else a.rays * world         # strong: --rays is the GLOBAL batch, split over the ranks
    lo, hi = (n_global * rank) // world, (n_global * (rank + 1)) // world
    n = hi - lo if strong else a.rays                       # rays of this rank per step
    args = synthetic.make_args(device=str(dev), batch_size=n)
    scene = synthetic.RoomScene()
    K, W, R = a.steps, a.warmup, max(1, a.windows)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t)
        return ms

    def global_loss(loss):
        """a rank's loss_out is ITS share sum_k w_k sums_k(local) / counts_k(global): the shares add up to the loss"""
        t = loss.detach().clone().reshape(1).float()
        if world > 1:
            dist.all_reduce(t)
        return float(t)

    def check_p2p(eng):
        if world > 1 and eng._p2p is not None and int(eng._p2p_err) != 0:
            raise SystemExit("bench.py: peer-memory exchange barrier timed out")

    def make_batches(ds, count, n_rank, strategy, shard=None):
        """weak: every rank draws its own batches; strong (shard=(lo, hi, n_global)): every rank draws the SAME global
        batch (identically seeded sampler) and keeps its contiguous shard"""
        out = []
        for _ in range(count):
            if shard is None:
                out.append(ds(n_rank, strategy))
            else:
                b = ds(shard[2], strategy)
                sl = slice(shard[0], shard[1])
                out.append({"rays_o": b["rays_o"][sl].contiguous(), "rays_d": b["rays_d"][sl].contiguous(),
                            "rgb": b["rgb"][sl].contiguous(), "depth": {k: v[sl].contiguous() for k, v in b["depth"].items()}})
        return out

    def run_phase(pinned):
        """W warm-up steps, then R windows of K timed steps, every window from the SAME restored state and over the same
        K batches (so the windows are repetitions of one measurement).  pinned=False: the batches are resident in HBM;
        pinned=True: every step's batch is copied from pinned host memory inside the timed region and the loss is read
        back to the host (end-to-end)."""
        ds = synthetic.SyntheticDataset(scene, pool_size=1 << 18, device=str(dev), pinned=False, seed=21)
        ds.gen.manual_seed(1000 + (0 if strong else rank))    # same pool on every rank; weak: different training batches
        eng = TrainEngine(args, ds, dev, world_size=world, rank=rank, autocast=not a.no_autocast, comm=a.comm,
                          enc_layout=a.enc_layout, single_pass_march=not a.two_pass_march,
                          fused_scatter=False if a.no_fused_scatter else "auto", early_expand=not a.no_early_expand)
        batches = make_batches(ds, W + K + 1, n, args.training.sampling_strategy, (lo, hi, n_global) if strong else None)
        noises = None
        if strong:          # the jitter of the GLOBAL batch, so that N ranks march exactly the samples one rank would
            gen = torch.Generator(device=dev); gen.manual_seed(9)
            noises = [torch.rand(n_global, device=dev, generator=gen)[lo:hi].contiguous() for _ in batches]
        host_batches = None
        if pinned:
            # ONE flat pinned buffer per batch: rays_o | rays_d | rgb | USS | ToF  (11 n floats, one H2D copy per step)
            host_batches = []
            for b in batches:
                flat = [b["rays_o"], b["rays_d"], b["rgb"], b["depth"]["USS"], b["depth"]["ToF"]]
                host_batches.append(torch.cat([t.reshape(-1).float() for t in flat]).cpu().pin_memory())
            loss_host = torch.zeros(1).pin_memory()
        h2d = d2h = 0

        copy_stream = torch.cuda.Stream(device=dev) if pinned else None

        def get_batch(it):
            if pinned:
                # the H2D copy of a batch goes through its own stream, so that it runs under the tail of the step that is
                # executing instead of in front of the next one; the compute stream waits for it (still inside the timed
                # region) before anything that reads the batch is enqueued
                with torch.cuda.stream(copy_stream):
                    buf = host_batches[it].to(dev, non_blocking=True)
                    landed = torch.cuda.Event(); landed.record()
                torch.cuda.current_stream().wait_event(landed)
                buf.record_stream(torch.cuda.current_stream())
                ro, rd, rgb_t = buf[0:3 * n].view(n, 3), buf[3 * n:6 * n].view(n, 3), buf[6 * n:9 * n].view(n, 3)
                return {"rays_o": ro, "rays_d": rd, "rgb": rgb_t, "depth": {"USS": buf[9 * n:10 * n], "ToF": buf[10 * n:11 * n]}}
            return batches[it]

        def steps(first, count):
            nonlocal h2d, d2h
            smp, loss = [], None
            data = get_batch(first)
            for it in range(first, first + count):
                # the next batch is fetched (H2D copy in the e2e phase) before this step is enqueued so that the
                # engine can pipeline its front half; every batch is copied exactly once per window
                nxt = get_batch(it + 1)
                if pinned:
                    h2d = host_batches[it].numel() * host_batches[it].element_size()
                if a.autograd_step:
                    loss = eng.step(data)
                elif noises is not None:
                    loss = eng.step_fast(data, noise=noises[it], next_data=nxt, next_noise=noises[it + 1])
                else:
                    loss = eng.step_fast(data, next_data=nxt)
                data = nxt
                if pinned:
                    loss_host.copy_(loss.reshape(1), non_blocking=True)   # D2H read of the step's result (pinned, in stream)
                    d2h = 4
                smp.append(eng.last_samples)
            return smp, loss

        steps(0, W)
        torch.cuda.synchronize()
        snap = _snapshot(eng)
        window_ms, samples, launches, loss = [], [], 0, None
        live_prof = {}
        for w in range(R):
            _restore(eng, snap)
            barrier()
            launches0 = _lib.launch_count()
            if not pinned:
                # only the candidate dominant kernels are bracketed with events inside the timed region (an event between
                # two kernels serialises them); the full per-kernel table comes from the untimed steps below
                _lib.profile_start(list(ROOFLINE_CANDIDATES))
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record()
            samples, loss = steps(W, K)
            ev1.record()
            barrier()
            window_ms.append(max_over_ranks(ev0.elapsed_time(ev1)))
            launches = _lib.launch_count() - launches0
            if not pinned:
                for k, v in _lib.profile_stop().items():
                    live_prof.setdefault(k, []).extend(v)
        check_p2p(eng)
        prof = {}
        if not pinned:
            # untimed breakdown: the same K steps again with every major kernel timed
            _restore(eng, snap)
            _lib.profile_start()
            steps(W, K)
            torch.cuda.synchronize()
            prof = _lib.profile_stop()
            prof["__timed__"] = {k: live_prof.get(k, []) for k in ROOFLINE_CANDIDATES}
        same = True
        if world > 1:
            c = eng.replica_checksum()
            clo, chi = c.clone(), c.clone()
            dist.all_reduce(clo, op=dist.ReduceOp.MIN); dist.all_reduce(chi, op=dist.ReduceOp.MAX)
            same = bool((clo == chi).all())
        return {"window_ms": window_ms, "launches": launches, "samples": [int(x) for x in samples], "prof": prof, "h2d": h2d,
                "d2h": d2h, "loss": global_loss(loss), "same": same, "eng": eng,
                "comm": eng.comm + ("+nvls" if getattr(eng, "nvls", False) else "")}

    clocks = ClockSampler(local)
    clocks.start()
    ph = run_phase(pinned=False)
    clk = clocks.stop()
    eng = ph["eng"]
    extra = {} if a.no_extra else extra_configs(a, eng, scene, dev, world, rank, barrier, max_over_ranks, make_batches, check_p2p, global_loss)
    eng.close()
    del eng
    ph["eng"] = None
    torch.cuda.empty_cache()
    ph_e2e = None if a.no_e2e else run_phase(pinned=True)
    if ph_e2e is not None:
        ph_e2e["eng"] = None

    if rank == 0:
        peak, peak_src = peaks()
        ms = statistics.median(ph["window_ms"])                 # K steps
        ms_e2e = statistics.median(ph_e2e["window_ms"]) if ph_e2e else ms
        total_rays = n_global * K
        value = total_rays / (ms * 1e-3)
        prof, samples = ph["prof"], ph["samples"]
        traffic = ncu_traffic()
        # roofline of the dominant kernel: algorithmic bytes / measured kernel time (CUDA events around the launches,
        # on the launching stream)
        kern = {}
        timed = prof.pop("__timed__", {})
        for name, calls in prof.items():
            t_ms = sum(c[0] for c in calls); pts = sum(c[1] for c in calls)
            if t_ms > 0:
                kern[name] = {"ms_total": round(t_ms, 4), "launches": len(calls), "units": pts,
                              "share_of_step": round(t_ms / ms, 4)}
                if name.startswith("hash_encode"):
                    kern[name]["achieved_gbs"] = pts * HASH_BYTES_PER_POINT / (t_ms * 1e-3) / 1e9
                    kern[name]["frac_of_hbm_peak"] = kern[name]["achieved_gbs"] / peak
                if name.startswith("mlp"):
                    flop = 18816 if name == "mlp_fwd" else 56448
                    kern[name]["achieved_tflops"] = pts * flop / (t_ms * 1e-3) / 1e12
                if name == "mlp_bwd_hash_scatter":
                    kern[name]["achieved_gbs"] = pts * FUSED_BWD_BYTES_PER_POINT / (t_ms * 1e-3) / 1e9
                    kern[name]["frac_of_hbm_peak"] = kern[name]["achieved_gbs"] / peak
        tflops_peak = peaks_tensor()
        # kernels timed live inside the timed windows: totals there decide which one dominates the step
        live = {}
        for name, calls in timed.items():
            t_ms = sum(c[0] for c in calls); pts = sum(c[1] for c in calls)
            if t_ms > 0:
                live[name] = (t_ms, pts, len(calls))
                kern.setdefault(name, {}).update({"ms_total_timed_region": round(t_ms, 4), "launches_timed_region": len(calls),
                                                  "units_timed_region": pts, "ms_per_launch_timed_region": round(t_ms / len(calls), 5),
                                                  "share_of_timed_step": round(t_ms / (ms * R), 4)})
        dom = max(live, key=lambda k: live[k][0]) if live else None
        roof = None
        if dom and (dom.startswith("hash_encode") or dom == "mlp_bwd_hash_scatter"):
            t_ms, pts, nl = live[dom]
            bpp = FUSED_BWD_BYTES_PER_POINT if dom == "mlp_bwd_hash_scatter" else HASH_BYTES_PER_POINT
            gbs = pts * bpp / (t_ms * 1e-3) / 1e9
            tr = traffic.get(dom)
            roof = {"kernel": dom, "bound": "hbm", "achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak,
                    "measured": f"CUDA events around every launch of this kernel inside the {R} timed windows ({nl} launches)",
                    "traffic": (tr["dram_bytes_per_point"] * pts / nl) if tr else None,
                    "traffic_note": (f"DRAM bytes per launch = dram__bytes_read.sum + dram__bytes_write.sum per point of this round's "
                                     f"ncu --set full capture ({tr['source']}) x mean points per launch; far below the algorithmic "
                                     f"bytes because the table and its gradient are L2 resident") if tr else "no ncu capture",
                    "peak_source": peak_src, "algorithmic_bytes_per_launch": bpp * pts / nl,
                    "algorithmic_bytes_per_point": bpp,
                    "what": ("fused MLP backward + hash-gradient scatter (one kernel): table-gradient RMW 1024 B + xyz 12 B + "
                             "MLP-backward inputs 112 B per point; also 56 448 FLOP per point on the tensor cores.  Round 1's "
                             "roofline kernel was the hash backward ALONE; it now runs inside this kernel together with the "
                             "whole MLP backward (extra_configs.backward_as_two_kernels has the two-kernel figures of the same "
                             "build: hash backward alone >= 0.62 of the HBM peak)"
                             if dom == "mlp_bwd_hash_scatter" else "hash-grid backward scatter"),
                    "runner_up": {k: round(v[0] / v[2], 5) for k, v in live.items() if k != dom}}
        elif dom and dom.startswith("mlp"):
            t_ms, pts, nl = live[dom]
            tf = pts * 56448 / (t_ms * 1e-3) / 1e12
            tr = traffic.get(dom)
            roof = {"kernel": dom, "bound": "tensor", "achieved": tf, "peak": tflops_peak, "unit": "TFLOP/s",
                    "frac": tf / tflops_peak, "traffic": (tr["dram_bytes_per_point"] * pts / nl) if tr else None,
                    "measured": f"CUDA events around every launch of this kernel inside the {R} timed windows ({nl} launches)",
                    "peak_source": "bf16_tflops_sustained, " + peak_src,
                    "runner_up": {k: round(v[0] / v[2], 5) for k, v in live.items() if k != dom}}
        spread = lambda xs: (max(xs) - min(xs)) / statistics.median(xs)
        line = {"metric": "train_rays_per_sec", "value": value, "unit": "rays/s", "n_gpus": world, "steps": K,
                "warmup": W, "ms_per_step": ms / K, "higher_is_better": True, "scaling": "strong" if strong else "weak",
                "vs_baseline": None,
                "dtype": "f32 tables/encoder/composite, fp16-autocast MLP" if not a.no_autocast else "f32",
                "data": "synthetic",
                "config": {"workload": WORKLOAD, "rays_per_step_per_gpu": n, "global_rays_per_step": n_global,
                           "samples_per_step_mean": statistics.mean(samples) if samples else 0,
                           "parallelism": f"dp{world}", "grad_exchange": ph["comm"],
                           "l2_policy": "inputs larger than L2: table+grad+Adam state 183 MB and per-step sample "
                                        "buffers are streamed; a fresh ray batch every step"},
                "windows": {"count": R, "what": f"the K = {K} timed steps are repeated {R} times, every window from the same "
                                               "restored state (parameters, Adam moments, scaler, occupancy grid, RNG) over the same "
                                               "batches; value / ms_per_step = median window, max over ranks per window",
                            "ms_per_step": [round(x / K, 5) for x in ph["window_ms"]], "spread": round(spread(ph["window_ms"]), 4),
                            "e2e_ms_per_step": [round(x / K, 5) for x in ph_e2e["window_ms"]] if ph_e2e else None},
                "roofline": roof, "kernels": kern,
                "kernels_note": "per-kernel table: K untimed steps after the timed windows with CUDA events around every major "
                                "launch (share_of_step relative to the timed ms/step); the roofline candidates are also timed "
                                "inside the timed windows",
                "e2e": {"value": total_rays / (ms_e2e * 1e-3), "unit": "rays/s", "h2d_bytes_per_step": ph_e2e["h2d"] if ph_e2e else 0,
                        "d2h_bytes_per_step": ph_e2e["d2h"] if ph_e2e else 0, "ms_per_step": ms_e2e / K},
                "gpu_launches": ph["launches"], "clocks": clk, "final_loss": ph["loss"],
                "replicas_bit_identical": ph["same"] and (ph_e2e["same"] if ph_e2e else True),
                "extra_configs": extra}
        if "render_1080p" in extra:
            line["render_1080p"] = extra["render_1080p"]
        if world == 1 and not a.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(a)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def ncu_traffic():
    """DRAM bytes per point of this round's `ncu --set full` captures (profiles/r2_ncu_traffic.json, written from the
    .ncu-rep files by tools/ncu_traffic.py)"""
    try:
        with open(os.path.join(ROOT, "profiles", "r2_ncu_traffic.json")) as f:
            return json.load(f)
    except Exception:
        return {}


def extra_configs(a, eng, scene, dev, world, rank, barrier, max_over_ranks, make_batches, check_p2p, global_loss):
    """BASELINE.json configs[2] and [4], untimed by the headline (after its timed windows, before the end-to-end phase):
    config 5 = 1080p test-time frame (rays sharded, no collective) + occupancy-update sweep; config 3 = Robot@Home2-shaped
    scene, T = 2^22, 2^18 rays per step split over the ranks (strong scaling), half-precision encoder."""
    import torch
    from virus_nerf_b200 import _lib, synthetic
    from virus_nerf_b200.engine import TrainEngine
    from virus_nerf_b200.modules.occupancy_grid import OccupancyGrid
    out = {}

    def timed(fn, reps=3):
        fn(); barrier()
        ts = []
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); barrier()
            ts.append(max_over_ranks(e0.elapsed_time(e1)))
        return statistics.median(ts)

    # ---- config 5a: one 1920x1080 frame, (i) with the grid as trained so far, (ii) with the scene's carved grid
    Wp, Hp = 1920, 1080
    u, v = torch.meshgrid(torch.arange(Wp, device=dev, dtype=torch.float32),
                          torch.arange(Hp, device=dev, dtype=torch.float32), indexing="xy")
    d = torch.stack([(u + 0.5 - Wp / 2) / 960, torch.ones_like(u), -(v + 0.5 - Hp / 2) / 960], -1).reshape(-1, 3)
    d = (d / d.norm(dim=1, keepdim=True)).contiguous()
    o = torch.tensor([0.0, -0.2, -0.05], device=dev).expand_as(d).contiguous()
    og = eng.model.occupancy_grid
    ms_trained = timed(lambda: eng.render_frame(o, d))
    bf_keep = og.getBitfield().clone()
    og.bitfield = torch.from_numpy(synthetic.morton_pack(scene.occupancy_bitfield(128))).to(dev)
    ms_carved = timed(lambda: eng.render_frame(o, d))
    og.bitfield = bf_keep
    out["render_1080p"] = {"ms_per_frame": ms_carved, "rays_per_s": 1920 * 1080 / (ms_carved * 1e-3),
                           "ms_per_frame_grid_as_trained": ms_trained, "n_gpus": world,
                           "note": "test-time render of one 1920x1080 frame (raymarching_test + fused MLP + composite_test), rays "
                                   "sharded over the ranks in bands, no collective; carved = the synthetic scene's occupancy "
                                   "(room shell + boxes), grid_as_trained = the grid after the benchmark's few dozen steps"}

    # ---- config 5b: occupancy-grid update (OccupancyGrid.update: ray update + NeRF update + decay + bitfield), rank 0's clock
    sweep = []
    for G in (128, 256):
        for B in (1024, 8192, 65536):
            args_g = synthetic.make_args(device=str(dev), occ_batch_size=B)
            ds_g = synthetic.SyntheticDataset(scene, pool_size=1 << 18, device=str(dev), seed=5)
            grid = OccupancyGrid(args=args_g, grid_size=G, dataset=ds_g, fct_density=eng.model.density)
            n0 = _lib.launch_count()
            grid.update(elapse_time=0.0)
            per_update = _lib.launch_count() - n0
            torch.cuda.synchronize()
            ts = []
            for _ in range(5):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); grid.update(elapse_time=0.0); e1.record(); torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            # the same back to back without a host sync in between (what the training loop does: the host runs ahead)
            reps = 20
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(); e0.record()
            for _ in range(reps):
                grid.update(elapse_time=0.0)
            e1.record(); torch.cuda.synchronize()
            sweep.append({"grid": G, "rays": B, "ms_per_update": round(statistics.median(ts), 4),
                          "ms_per_update_pipelined": round(e0.elapsed_time(e1) / reps, 4), "kernels_per_update": per_update,
                          "note": "ms_per_update = one update with a host sync on both sides (host latency of two dataset "
                                  "samplings + one vn_occ_update call); pipelined = 20 updates enqueued back to back"})
            del grid
    out["occupancy_update_sweep"] = sweep

    # ---- the backward as TWO kernels (MLP backward, hash backward), for comparison with round 1's roofline kernel: the
    # headline runs them fused (mlp_bwd_hash_scatter) when the table is L2 resident.  Rank 0 only, single engine, untimed.
    if world == 1 and eng.fused_scatter:
        args_u = synthetic.make_args(device=str(dev), batch_size=a.rays)
        ds_u = synthetic.SyntheticDataset(scene, pool_size=1 << 18, device=str(dev), pinned=False, seed=21)
        ds_u.gen.manual_seed(1000)
        eng_u = TrainEngine(args_u, ds_u, dev, fused_scatter=False)
        bu = [ds_u(a.rays, args_u.training.sampling_strategy) for _ in range(a.warmup + a.steps + 1)]
        for it in range(a.warmup):
            eng_u.step_fast(bu[it], next_data=bu[it + 1])
        torch.cuda.synchronize()
        _lib.profile_start(["hash_encode_bwd", "mlp_bwd"])
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for it in range(a.warmup, a.warmup + a.steps):
            eng_u.step_fast(bu[it], next_data=bu[it + 1])
        e1.record(); torch.cuda.synchronize()
        pu = _lib.profile_stop()
        peak, _ = peaks()
        hb, mb = pu.get("hash_encode_bwd", []), pu.get("mlp_bwd", [])
        if hb and mb:
            t_h, p_h = sum(c[0] for c in hb), sum(c[1] for c in hb)
            t_m, p_m = sum(c[0] for c in mb), sum(c[1] for c in mb)
            out["backward_as_two_kernels"] = {
                "ms_per_step": e0.elapsed_time(e1) / a.steps,
                "hash_encode_bwd": {"ms_per_launch": round(t_h / len(hb), 5), "points_per_launch": p_h // len(hb),
                                    "achieved_gbs": round(p_h * HASH_BYTES_PER_POINT / (t_h * 1e-3) / 1e9, 1),
                                    "frac_of_hbm_peak": round(p_h * HASH_BYTES_PER_POINT / (t_h * 1e-3) / 1e9 / peak, 4),
                                    "algorithmic_bytes_per_point": HASH_BYTES_PER_POINT},
                "mlp_bwd": {"ms_per_launch": round(t_m / len(mb), 5), "achieved_tflops": round(p_m * 56448 / (t_m * 1e-3) / 1e12, 1)},
                "note": "the same steps with fused_scatter=False (round 1's structure, this round's kernels): the hash backward "
                        "alone is the kernel round 1 quoted its roofline fraction on"}
        del eng_u
        torch.cuda.empty_cache()

    # ---- config 3: RH2-shaped (RGBD + USS + ToF), T = 2^22, 2^18 rays per step over all ranks, half-precision encoder
    if not a.no_config3:
        n3 = 1 << 18
        lo3, hi3 = (n3 * rank) // world, (n3 * (rank + 1)) // world
        args3 = synthetic.make_args(device=str(dev), batch_size=hi3 - lo3, sensors=("RGBD", "USS", "ToF"))
        args3.training.sampling_strategy = {"imgs": "all", "pixs": "random"}
        ds3 = synthetic.SyntheticDataset(scene, kind="rh2", pool_size=1 << 19, device=str(dev), seed=33)
        ds3.gen.manual_seed(77)                                              # the same global batches on every rank
        eng.close()                                                          # one owner of the peer-memory exchange at a time
        eng3 = TrainEngine(args3, ds3, dev, world_size=world, rank=rank, log2_T=22, half_opt=True, comm=a.comm,
                           fused_scatter=False if a.no_fused_scatter else "auto", early_expand=not a.no_early_expand)
        W3, K3 = 2, 4
        b3 = make_batches(ds3, W3 + K3 + 1, hi3 - lo3, args3.training.sampling_strategy, (lo3, hi3, n3))
        gen = torch.Generator(device=dev); gen.manual_seed(5)
        noise3 = [torch.rand(n3, device=dev, generator=gen)[lo3:hi3].contiguous() for _ in b3]   # same jitter as the 1-GPU run
        losses = []
        for it in range(W3):
            losses.append(eng3.step_fast(b3[it], noise=noise3[it]).clone())
        barrier()
        _lib.profile_start(["hash_encode_fwd", "hash_encode_bwd", "mlp_fwd", "mlp_bwd", "mlp_bwd_hash_scatter"])
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        smp = []
        for it in range(W3, W3 + K3):
            losses.append(eng3.step_fast(b3[it], noise=noise3[it]).clone())
            smp.append(eng3.last_samples)
        e1.record()
        barrier()
        ms3 = max_over_ranks(e0.elapsed_time(e1))
        prof3 = _lib.profile_stop()
        check_p2p(eng3)
        peak, _ = peaks()
        k3 = {}
        for name, calls in prof3.items():
            t_ms = sum(c[0] for c in calls); pts = sum(c[1] for c in calls)
            if t_ms > 0:
                k3[name] = {"ms_per_launch": round(t_ms / len(calls), 4), "points": pts // len(calls)}
                if name == "mlp_bwd_hash_scatter":
                    k3[name]["achieved_gbs"] = round(pts * FUSED_BWD_BYTES_PER_POINT_HALF / (t_ms * 1e-3) / 1e9, 1)
                    k3[name]["frac_of_hbm_peak"] = round(k3[name]["achieved_gbs"] / peak, 4)
                    k3[name]["algorithmic_bytes_per_point"] = FUSED_BWD_BYTES_PER_POINT_HALF
                if name.startswith("hash_encode"):
                    # algorithmic bytes per point, half encoder: fwd 588 (fp16 table reads + fp16 out),
                    # bwd 1100 (fp32 scatter); reported next to the fp32 figure of the headline
                    bpp = 588 if name == "hash_encode_fwd" else 1100
                    k3[name]["achieved_gbs"] = round(pts * bpp / (t_ms * 1e-3) / 1e9, 1)
                    k3[name]["frac_of_hbm_peak"] = round(k3[name]["achieved_gbs"] / peak, 4)
                    k3[name]["algorithmic_bytes_per_point"] = bpp
        out["config3_rh2_T22_half"] = {
            "metric": "train_rays_per_sec", "value": n3 * K3 / (ms3 * 1e-3), "unit": "rays/s", "ms_per_step": ms3 / K3,
            "n_gpus": world, "scaling": "strong", "global_rays_per_step": n3, "rays_per_step_per_gpu": hi3 - lo3,
            "samples_per_step_per_gpu": int(statistics.mean(smp)), "steps": K3, "warmup": W3,
            "encoder": "half (fp16 table copy per step, fp16 encoding and encoding gradient, fp32 scatter)", "log2_T": 22,
            "table_mb_fp32": round(eng3.model.pos_encoder.hash_table.numel() * 4 / 2 ** 20, 1), "grad_exchange": eng3.comm + ("+nvls" if getattr(eng3, "nvls", False) else ""),
            "losses": [round(global_loss(x), 6) for x in losses], "kernels": k3,
            "note": "one globally seeded batch per step, rank r trains on rays [r N/n, (r+1) N/n) with the jitter of the global "
                    "batch: `losses` must agree between N = 1 and N > 1 (global loss normalisers, summed gradients)"}
        eng3.close()
        del eng3
        torch.cuda.empty_cache()
    return out


def cpu_baseline(a):
    """bounded sample of the same workload on the host cores through the oracle pipeline"""
    try:
        out = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", "12",
                              "--warmup", "2", "--cpu-rays", str(a.cpu_rays)], capture_output=True, text=True,
                             timeout=600, env={**os.environ, "RANK": "0", "WORLD_SIZE": "1"})
        line = json.loads(out.stdout.strip().splitlines()[-1])
        return line["cpu_baseline"]
    except Exception as e:   # the baseline is a reported number, never a reason to lose the bench line
        return {"value": None, "unit": "rays/s", "cores": len(os.sched_getaffinity(0)), "kind": "port",
                "sample": f"failed: {e!r}"}


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
