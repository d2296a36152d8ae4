#!/usr/bin/env python
"""bench.py -- DirectVoxGO hot-path benchmark on B200 (contract: see the task's "Measurement").

Workload (BASELINE.json configs[1]): fine stage, 160^3 density + 12-channel k0 + rgbnet (width 128),
8192 incoherent synthetic Blender-geometry rays per step, one step = forward + loss + backward +
total-variation + MaskedAdam.  Random-init N(0,1) grids, procedural targets (`data: synthetic`).

  python bench.py [--gpus N --steps K --warmup W]          our arm (N>1 under torchrun)
  python bench.py --impl reference [...]                    the reference's algorithm on host cores

One JSON line on stdout (rank 0).  `value` = rays/s with inputs resident in HBM; `e2e` = rays/s
through the public step() call with pinned-host rays copied H2D and the loss read back D2H every
step; `roofline` = achieved algorithmic GB/s of the dominant kernel vs MEASURED_PEAKS.json;
`cpu_baseline` = the CPU oracle model (oracle/model_ref.py) timed on a bounded sample.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "train rays/s (fwd+bwd+TV+MaskedAdam) @160^3 fine stage"
WORKLOAD = "DVGO fine stage %d^3 density + 12ch k0 + rgbnet(128), 8192 rays/iter/GPU, fwd+bwd+TV(dense)+MaskedAdam"
N_RAYS = 8192
N_BATCHES = 16  # distinct ray batches cycled through (each step sees different incoherent rays)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--exchange", default="auto", choices=["auto", "peer", "peer_p2p", "nccl"],
                    help="multi-GPU gradient exchange of the fused trainer (see DESIGN.md section 6)")
    ap.add_argument("--ramp-s", dest="ramp_s", type=float, default=1.5,
                    help="seconds of untimed steps before the W warm-up steps (GPU clock ramp)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--path", default=os.environ.get("DVGO_BENCH_PATH", "auto"),
                    choices=["auto", "fused", "module"], help="fused B200 trainer or op-by-op module path")
    ap.add_argument("--grid", type=int, default=160)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-render", action="store_true", help="skip the secondary 800x800 render measurement")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("hbm_gbs", 6650.0), d.get("bf16_tflops_sustained", 1400.0), "measured"
    return 6650.0, 1590.0, "fallback"


class ClockSampler(threading.Thread):
    """Samples SM clocks and throttle reasons with nvidia-smi while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self._stop_evt = index, [], threading.Event()

    def run(self):
        while not self._stop_evt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                f = [x.strip() for x in out.strip().split(",")]
                if len(f) >= 6:
                    self.samples.append(f)
            except Exception:
                pass
            self._stop_evt.wait(0.2)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=6)
        sm = sorted(int(s[0]) for s in self.samples if s[0].isdigit())
        mx = [int(s[1]) for s in self.samples if s[1].isdigit()]
        reasons = set()
        for s in self.samples:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), s[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(self.samples)}


def build_problem(grid, device, seed=777):
    """Model (fine stage at grid^3) + render kwargs + train cfg, per SURVEY.md 8d cfg 2."""
    import torch
    from directvoxgo_b200 import synthetic as syn
    from directvoxgo_b200.dvgo import DirectVoxGO
    lo, hi = syn.fine_bbox()
    kw = dict(syn.FINE_MODEL, num_voxels=grid ** 3, num_voxels_base=grid ** 3)
    torch.manual_seed(seed)
    model = DirectVoxGO(lo, hi, **kw)
    syn.randomize_grids_(model, seed)
    return model.to(device), dict(syn.RENDER_KWARGS), dict(syn.FINE_TRAIN)


def make_batches(n_batches, n_rays, device, rank):
    import torch
    from directvoxgo_b200 import synthetic as syn
    out = []
    for b in range(n_batches):
        ro, rd, vd, tgt = syn.random_training_rays(n_rays, n_views=100, seed=1000 * (rank + 1) + b, device="cpu")
        out.append(tuple(t.pin_memory() if torch.cuda.is_available() else t for t in (ro, rd, vd, tgt)))
    dev = [tuple(t.to(device, non_blocking=True) for t in b) for b in out]
    return out, dev


def algorithmic_bytes(model, batch, rk):
    """B_alg of SURVEY.md 8d for one step: U*(1+C)*4*3 (gather once + scatter RMW) + E*4 (read every
    grad) + U'*(1+C)*28 (Adam on touched cells; U' = G under dense TV).  U = unique voxels touched as
    trilinear corners by the step's in-bbox samples (computed here with torch.unique)."""
    import torch
    ro, rd, vd, tgt = batch
    with torch.no_grad():
        pts, ray_id, step_id = model.sample_ray(ro, rd, **rk)
        X, Y, Z = (int(s) for s in model.world_size)
        f = (pts - model.xyz_min) / (model.xyz_max - model.xyz_min) * torch.tensor([X - 1, Y - 1, Z - 1], device=pts.device)
        i0 = f.floor().long()
        idx = []
        for dx in (0, 1):
            for dy in (0, 1):
                for dz in (0, 1):
                    c = i0 + torch.tensor([dx, dy, dz], device=pts.device)
                    ok = ((c >= 0) & (c < torch.tensor([X, Y, Z], device=pts.device))).all(-1)
                    idx.append(((c[:, 0] * Y + c[:, 1]) * Z + c[:, 2])[ok])
        U = int(torch.unique(torch.cat(idx)).numel())
    C = model.k0.shape[1]
    G = X * Y * Z
    E = (1 + C) * G
    return {"U": U, "G": G, "M0": int(pts.shape[0]),
            "gather_scatter": U * (1 + C) * 4 * 3, "grad_read": E * 4, "adam_dense": G * (1 + C) * 28,
            "total": U * (1 + C) * 4 * 3 + E * 4 + G * (1 + C) * 28}


def sphere_scene(model, device):
    """The 'procedural occupancy' scene of SURVEY.md 8d: density +5 inside a ball of radius 0.6 half-extents,
    -5 outside, occupancy mask derived from it by the reference rule (lib/dvgo.py:254-259).  Labelled an extra."""
    import copy
    import torch
    from directvoxgo_b200.dvgo import MaskCache
    m = copy.deepcopy(model)
    with torch.no_grad():
        X, Y, Z = m.density.shape[2:]
        ax = [torch.linspace(-1, 1, n, device=device) for n in (X, Y, Z)]
        r = torch.stack(torch.meshgrid(*ax, indexing="ij"), -1).norm(dim=-1)
        m.density.copy_(torch.where(r < 0.6, 5.0, -5.0)[None, None])
        alpha = torch.nn.functional.max_pool3d(m.activate_density(m.density), 3, 1, 1)[0, 0]
        m.mask_cache = MaskCache(mask=(alpha > m.fast_color_thres), xyz_min=m.xyz_min, xyz_max=m.xyz_max).to(device)
    return m


def train_sphere_metric(model, rk, cfg, dev_batches, device, steps=100):
    """Extra (not the headline): the same fused training step on the sphere-occupancy scene, where the four-mask
    cascade and the early stop cull most samples as on a real scene (the random-init grids cull nothing)."""
    import torch
    from directvoxgo_b200.fused import FusedTrainer
    tr = FusedTrainer(sphere_scene(model, device), cfg, rk)
    for i in range(50):
        tr.step(*dev_batches[i % len(dev_batches)])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        tr.step(*dev_batches[i % len(dev_batches)])
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    n = dev_batches[0][0].shape[0]
    return {"ms_per_step": ms, "rays_per_s": n / ms * 1e3, "survivors_last_step": int(tr._workspace(n, True).counters[0].item()),
            "grid": "sphere-occupancy (extra)", "steps": steps}


def render_metric(model, rk, device, n_frames=2, chunk=65536, sphere=False, rank=0, world=1):
    """Secondary metric of BASELINE.json: ms per rendered 800x800 frame (run.py:57-110: rays of a view ->
    chunks -> forward, render_depth=True), device-timed with CUDA events, rays generated on the device.
    sphere=True: the 'procedural occupancy' variant of SURVEY.md 8d (density +5 inside a ball of radius 0.6
    half-extents, -5 outside, occupancy mask derived from it) -- labelled as an extra."""
    import copy
    import torch
    from directvoxgo_b200 import synthetic as syn
    from directvoxgo_b200.dvgo import MaskCache
    from directvoxgo_b200.fused import FusedRenderer
    m = sphere_scene(model, device) if sphere else model
    renderer = FusedRenderer(m, rk)
    H = W = syn.BLENDER["H"]
    K = syn.intrinsics(H, W)
    # BASELINE config 3: whole views are sharded over the ranks (view i -> rank i mod world), no communication;
    # every rank renders n_frames views, the aggregate is (n_frames * world) frames in the max-over-ranks time.
    from directvoxgo_b200.parallel import shard_views
    poses = syn.random_poses(n_frames * world + 1, seed=4242)
    mine = [poses[1 + i] for i in shard_views(n_frames * world, rank, world)]
    renderer.render_view(H, W, K, poses[0], chunk=chunk)  # warm-up (allocates the workspace)
    torch.cuda.synchronize()
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for c2w in mine:   # rays generated on the device per 65 536-pixel chunk (dvgo_rays_of_view), then rendered
        img = renderer.render_view(H, W, K, c2w, chunk=chunk)["rgb_marched"]
    ev1.record()
    torch.cuda.synchronize()
    t = torch.tensor([ev0.elapsed_time(ev1)], device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    ws = renderer._workspace(min(chunk, H * W), False)
    return {"ms_per_frame": ms / (n_frames * world), "ms_per_frame_per_gpu": ms / n_frames, "frames": n_frames * world,
            "frames_per_s": n_frames * world / ms * 1e3, "sharding": "view i -> rank i mod %d, no communication" % world,
            "rays_per_call": chunk, "resolution": "%dx%d" % (H, W),
            "grid": "sphere-occupancy (extra)" if sphere else "random-init N(0,1), nothing culled",
            "survivors_last_call": int(ws.counters[0].item()), "mean_rgb": float(img.mean())}


def cpu_reference_run(grid, n_rays, steps, warmup, threads):
    """The reference's algorithm on the host cores: oracle/model_ref.py (torch-CPU + C oracle)."""
    import torch
    from directvoxgo_b200 import synthetic as syn
    from oracle.model_ref import RefDVGO
    torch.set_num_threads(threads)
    model, rk, cfg = build_problem(grid, "cpu")
    ref = RefDVGO.from_module(model)
    del model
    ro, rd, vd, tgt = syn.random_training_rays(n_rays, n_views=100, seed=1000, device="cpu")
    for _ in range(warmup):
        ref.train_step(ro, rd, vd, tgt, rk, cfg)
    t0 = time.perf_counter()
    for _ in range(steps):
        ref.train_step(ro, rd, vd, tgt, rk, cfg)
    dt = (time.perf_counter() - t0) / max(steps, 1)
    return n_rays / dt, dt


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    n_rays = 1024  # bounded sample of the 8192-ray step (same grid, same ops)
    steps = max(1, min(args.steps, 5))   # bounded: ~2 s per step on 16 host cores
    rays_s, dt = cpu_reference_run(args.grid, n_rays, steps, min(args.warmup, 1), threads)
    line = {"metric": METRIC, "value": rays_s, "unit": "rays/s", "n_gpus": args.gpus, "steps": steps,
            "warmup": min(args.warmup, 1), "ms_per_step": dt * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "impl": "reference",
            "config": {"workload": WORKLOAD % args.grid, "rays_per_step_per_gpu": N_RAYS,
                       "sample": "each timed step = a %d-ray slice of the 8192-ray step" % n_rays},
            "cpu_baseline": {"value": rays_s, "unit": "rays/s", "cores": threads, "kind": "port",
                             "sample": "%d-ray slice of the 8192-ray step on the full %d^3 grid, %d step(s); "
                                       "reference CUDA ops have no CPU path, so this is oracle/model_ref.py" % (n_rays, args.grid, steps)},
            "e2e": {"value": rays_s, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def run_ours(args):
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product has no CPU path")
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        import datetime
        # a desynchronised collective must abort the run quickly instead of hanging the box
        dist.init_process_group("nccl", device_id=device, timeout=datetime.timedelta(seconds=120))
    import directvoxgo_b200 as pkg
    from directvoxgo_b200.trainer import ModuleTrainer

    model, rk, cfg = build_problem(args.grid, device)
    path = args.path
    trainer = None
    if path in ("auto", "fused"):
        try:
            from directvoxgo_b200.fused import FusedTrainer
            trainer = FusedTrainer(model, cfg, rk, world_size=world, exchange=args.exchange)
            path = "fused"
        except ImportError:
            if path == "fused":
                raise
    if trainer is None:
        trainer = ModuleTrainer(model, cfg, rk, world_size=world)
        path = "module"

    host_batches, dev_batches = make_batches(N_BATCHES, N_RAYS, device, rank)
    balg = algorithmic_bytes(model, dev_batches[0], rk) if rank == 0 else None
    torch.cuda.synchronize()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing -----------------------------------------------------------------
    # Untimed clock ramp: a B200 idling at its floor clock needs ~1 s of load to reach its boost clock; W warm-up
    # steps (a few ms) are not enough and the first timed steps would otherwise run at a lower clock.
    # The steps contain collectives when world > 1, so every rank must run the SAME number of rounds: the ranks
    # agree on "done" with a MAX all-reduce after each round (a per-rank wall-clock test would desynchronise them).
    t_ramp = time.time()
    while args.ramp_s > 0:
        for i in range(50):
            trainer.step(*dev_batches[i % N_BATCHES])
        torch.cuda.synchronize()
        done = torch.tensor([1.0 if time.time() - t_ramp >= args.ramp_s else 0.0], device=device)
        if world > 1:
            dist.all_reduce(done, op=dist.ReduceOp.MAX)
        if float(done.item()) > 0:
            break
    for i in range(args.warmup):
        trainer.step(*dev_batches[i % N_BATCHES])
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    l0 = pkg._C.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for i in range(args.steps):
        loss = trainer.step(*dev_batches[(args.warmup + i) % N_BATCHES])
    ev1.record()
    barrier()
    launches = pkg._C.launch_count() - l0
    ms = ev0.elapsed_time(ev1)
    t = torch.tensor([ms], device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_per_step = float(t.item()) / args.steps
    value = N_RAYS * world / (ms_per_step * 1e-3)

    # ---- end to end: pinned host rays in, loss out, every step --------------------------------------
    stage = [torch.empty_like(x, device=device) for x in host_batches[0]]
    for i in range(min(3, args.warmup)):
        for d, h in zip(stage, host_batches[i % N_BATCHES]):
            d.copy_(h, non_blocking=True)
        float(trainer.step(*stage).item())
    barrier()
    ev0.record()
    for i in range(args.steps):
        for d, h in zip(stage, host_batches[i % N_BATCHES]):
            d.copy_(h, non_blocking=True)
        loss_host = float(trainer.step(*stage).item())
    ev1.record()
    barrier()
    t = torch.tensor([ev0.elapsed_time(ev1)], device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms = float(t.item()) / args.steps
    clocks = sampler.stop()  # sampled across both timed regions (device-resident and end-to-end)
    h2d = sum(x.numel() * x.element_size() for x in host_batches[0])

    # Per-stage device times: CUDA events recorded on the launching stream around each stage of extra steps
    # (outside the timed regions).  EVERY rank runs them -- the steps contain collectives when world > 1.
    stages = None
    if path == "fused" and hasattr(trainer, "stage_events"):
        barrier()
        for i in range(3):              # re-align the ranks after the end-to-end loop before recording events
            trainer.step(*dev_batches[i % N_BATCHES])
        barrier()
        trainer.stage_events = {}
        for i in range(10):
            trainer.step(*dev_batches[i % N_BATCHES])
        torch.cuda.synchronize()
        stages = trainer.stage_times_ms()
        if "sweep_grids" in stages:     # multi-GPU: grid sweeps and the rgbnet Adam are marked separately
            stages["sweep"] = stages["sweep"] + stages.pop("sweep_grids")
        trainer.stage_events = None
    barrier()
    render = None
    if not args.no_render:      # every rank renders its share of the views
        try:
            if hasattr(trainer, "sync_to_model"):
                trainer.sync_to_model()
            render = {"dense": render_metric(model, rk, device, 2, 65536, False, rank, world),
                      "sphere": render_metric(model, rk, device, 2, 65536, True, rank, world)}
            if world == 1 and path == "fused":
                render["train_sphere_extra"] = train_sphere_metric(model, rk, cfg, dev_batches, device)
        except Exception as e:  # secondary metric: never let it break the headline line
            if world > 1:
                raise
            render = {"error": repr(e)[:200]}
    barrier()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    hbm_peak, tensor_peak, peak_kind = peaks()
    # Roofline of the dominant kernel (the dominant stage is a single kernel).
    roof = None
    if stages is not None:
        M = balg["M0"]
        C = model.k0.shape[1]
        G = balg["G"]
        # algorithmic work per launch: bytes for the HBM-bound kernels, FLOPs for the MLP
        alg = {
            "march_fwd": ("hbm", balg["U"] * (1 + C) * 4 + M * (16 + C * 4 + 12)),   # touched cells once + sample stream out
            "mlp_fwd": ("tensor", M * 43520.0),
            "mlp_bwd": ("tensor", M * 43520.0 * 2),
            "march_bwd": ("hbm", balg["U"] * (1 + C) * 8 + M * (16 + C * 4 + 4)),    # grad cells RMW + sample stream in
            "sweep": ("hbm", G * (1 + C) * 32 / (world if getattr(trainer, "_slab", lambda: None)() else 1)),  # p,g,m,v in; p,m,v,g=0 out
        }
        dom = max((k for k in stages if k in alg), key=lambda k: stages[k])
        kind, work = alg[dom]
        t = stages[dom] * 1e-3
        if kind == "hbm":
            ach, peak, unit = work / t / 1e9, hbm_peak, "GB/s"
        else:
            ach, peak, unit = work / t / 1e12, tensor_peak, "TFLOP/s"
        kernel_names = {"march_fwd": "march_fwd_kernel<12> + k0_gather_kernel<12>", "mlp_fwd": "mlp_fwd_kernel", "mlp_bwd": "mlp_bwd_kernel",
                        "march_bwd": "k0_scatter_kernel<12> + march_bwd_kernel<12>", "sweep": "sweep_kernel<4,true> (+density sweep, rgbnet Adam)"}
        # dram__bytes_read.sum + dram__bytes_write.sum per launch, from the committed ncu --set full capture of this
        # command (profiles/r01_ncu_final_kernels.md); valid for the default workload only.
        ncu_traffic = {"march_fwd": 35.9e6 + 457.6e6, "mlp_fwd": 142.6e6, "mlp_bwd": 268.7e6, "march_bwd": 65.5e6 + 598.1e6, "sweep": 1473.3e6 + 75.6e6}
        traffic = ncu_traffic.get(dom) if (args.grid == 160 and world == 1) else None
        comm = {k: stages[k] for k in ("grad_exchange", "param_gather") if k in stages}
        roof = {"bound": kind, "kernel": kernel_names[dom], "achieved": ach, "peak": peak, "unit": unit,
                "frac": ach / peak, "traffic": traffic, "peak_kind": peak_kind + (" (bf16 sustained; fp16 runs at the same rate)" if kind == "tensor" else ""),
                "kernel_ms": stages[dom], "algorithmic_work_per_launch": work,
                "all_stages": {k: {"ms": stages[k], "bound": alg[k][0],
                                   "frac": (alg[k][1] / (stages[k] * 1e-3) / (1e9 * hbm_peak if alg[k][0] == "hbm" else 1e12 * tensor_peak))}
                               for k in stages if k in alg}}
        if comm:
            roof["collectives_ms"] = comm
    if roof is None:
        # module path: the grid-optimiser sweep (masked Adam over density+k0) is the one pure-HBM kernel
        from directvoxgo_b200 import adam_upd_cuda
        p = model.k0.detach()
        g = torch.randn_like(p)
        m_, v_ = torch.zeros_like(p), torch.zeros_like(p)
        for _ in range(3):
            adam_upd_cuda.masked_adam_upd(p, g, m_, v_, 1, 0.9, 0.99, 0.0, 1e-8)
        torch.cuda.synchronize()
        ev0.record()
        reps = 10
        for _ in range(reps):
            adam_upd_cuda.masked_adam_upd(p, g, m_, v_, 1, 0.9, 0.99, 0.0, 1e-8)
        ev1.record()
        torch.cuda.synchronize()
        k_ms = ev0.elapsed_time(ev1) / reps
        bytes_alg = p.numel() * 28
        roof = {"bound": "hbm", "kernel": "adam_kernel<masked> over k0", "achieved": bytes_alg / (k_ms * 1e-3) / 1e9,
                "peak": hbm_peak, "unit": "GB/s", "frac": bytes_alg / (k_ms * 1e-3) / 1e9 / hbm_peak,
                "traffic": None, "peak_kind": peak_kind, "kernel_ms": k_ms, "algorithmic_bytes": bytes_alg}

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        n_cpu = 1024
        rays_s, dt = cpu_reference_run(args.grid, n_cpu, 4, 1, threads)
        cpu = {"value": rays_s, "unit": "rays/s", "cores": threads, "kind": "port",
               "sample": "%d-ray slice of the 8192-ray step on the full %d^3 grid, 4 steps (oracle/model_ref.py: "
                         "torch-CPU + C oracle; the reference's CUDA ops have no CPU path)" % (n_cpu, args.grid)}

    line = {
        "metric": METRIC, "value": value, "unit": "rays/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32 (grids, sampling, compositing, Adam); rgbnet GEMM operands fp16, fp32 accumulate",
        "data": "synthetic",
        "config": {"workload": WORKLOAD % args.grid,
                   "rays_per_step_per_gpu": N_RAYS, "path": path, "rgbnet": getattr(trainer, "mlp_mode", "torch"), "parallelism": "ray-sharded dp%d" % world,
                   "grad_exchange": getattr(trainer, "exchange", "nccl all-reduce" if world > 1 else "none") +
                                    (" + NVLS multicast" if getattr(trainer, "multicast", False) else ""),
                   "samples_per_step": balg["M0"], "unique_voxels_touched": balg["U"],
                   "l2_policy": "working set (params+grads+Adam state = %.2f GB) larger than the 126 MB L2; "
                                "%d distinct ray batches cycled" % (balg["G"] * 13 * 16 / 1e9, N_BATCHES),
                   "algorithmic_bytes_per_step": balg["total"], "clock_ramp_s": args.ramp_s},
        "e2e": {"value": N_RAYS * world / (e2e_ms * 1e-3), "unit": "rays/s", "ms_per_step": e2e_ms,
                "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4, "last_loss": loss_host},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": roof,
        "step_hbm_frac": balg["total"] / (ms_per_step * 1e-3) / 1e9 / hbm_peak,
        "cpu_baseline": cpu,
        "train_sphere_extra": (render or {}).pop("train_sphere_extra", None),
        "render_800x800": render,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference_arm(a)
    else:
        run_ours(a)
