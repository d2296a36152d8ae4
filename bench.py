#!/usr/bin/env python
"""bench.py -- DirectVoxGO hot-path benchmark on B200 (contract: see the task's "Measurement").

Workload (BASELINE.json configs[1]): fine stage, 160^3 density + 12-channel k0 + rgbnet (width 128),
8192 incoherent synthetic Blender-geometry rays per step and GPU, one step = forward + loss + backward +
total-variation + MaskedAdam.  Random-init N(0,1) grids, procedural targets (`data: synthetic`).

  python bench.py [--gpus N --steps K --warmup W]          our arm (N>1 under torchrun)
  python bench.py --workload cfg5 [...]                     BASELINE configs[4]: 320^3, 65 536 rays split over N (strong)
  python bench.py --impl reference [...]                    the reference's algorithm on the host cores

One JSON line on stdout (rank 0).  What is measured, and from which state:

  * every measured pass starts from the SAME random-init state: the clock ramp and the W warm-up steps run on the
    trainer, then its parameters / Adam state are restored from a snapshot taken at construction
    (FusedTrainer.snapshot / restore), so neither trains the model that is timed;
  * `value`    K steps, inputs resident in HBM, CUDA events around the loop, max over ranks;
  * `e2e`      the same K steps through the public host-fed loop (trainer.HostFedLoop): pinned-host rays copied H2D and
    the loss read back for every step; `serialised_ms_per_step` = the same with copy -> step -> read strictly in turn;
  * `roofline` a third pass of the same K steps with CUDA events between the stages (per-stage MEDIAN over the steps:
    a host-side launch hiccup otherwise lands in whichever stage was waiting); the survivor count M4 of exactly
    those steps is accumulated on the device (dvgo_fused_step_begin) and printed (`survivors_per_step`), and the MLP
    FLOPs / gather-scatter bytes are computed from it -- every `frac` can be recomputed from the line;
  * `cpu_baseline` / `--impl reference`: oracle/model_ref.py (the reference's algorithm restated on torch-CPU + the C
    oracle; its CUDA ops have no CPU path) on one full 8192-ray step;
  * `ref_gpu_baseline` (N=1): the reference's OWN CUDA kernels (oracle/_ref) + ATen on this GPU, same step.
"""
import argparse
import glob
import importlib.util
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (grid, rays per step, split over ranks?, metric, workload text, scaling)
    "cfg2": (160, 8192, False, "train rays/s (fwd+bwd+TV+MaskedAdam) @160^3 fine stage",
             "DVGO fine stage %d^3 density + 12ch k0 + rgbnet(128), 8192 rays/iter/GPU, fwd+bwd+TV(dense)+MaskedAdam", "weak"),
    "cfg5": (320, 65536, True, "train rays/s (fwd+bwd+TV+MaskedAdam) @320^3, 65536 rays/iter ray-sharded",
             "DVGO large grid %d^3 density + 12ch k0 + rgbnet(128), 65536 rays/iter split over the GPUs, "
             "fwd+bwd+TV(dense)+MaskedAdam", "strong"),
}
N_BATCHES = 16  # distinct ray batches cycled through (each step sees different incoherent rays)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--exchange", default="auto", choices=["auto", "peer", "peer_p2p", "nccl"],
                    help="multi-GPU gradient exchange of the fused trainer (see DESIGN.md section 6)")
    ap.add_argument("--ramp-s", dest="ramp_s", type=float, default=1.5,
                    help="seconds of untimed steps before the W warm-up steps (GPU clock ramp)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--path", default=os.environ.get("DVGO_BENCH_PATH", "auto"),
                    choices=["auto", "fused", "module"], help="fused B200 trainer or op-by-op module path")
    ap.add_argument("--grid", type=int, default=0, help="override the workload's grid size")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-ref-gpu", action="store_true", help="skip the reference-CUDA-kernels baseline (N=1)")
    ap.add_argument("--no-render", action="store_true", help="skip the secondary 800x800 render measurement")
    ap.add_argument("--no-extras", action="store_true", help="skip the sphere-scene and cfg5 extras")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("hbm_gbs", 6650.0), d.get("bf16_tflops_sustained", 1400.0), "measured"
    return 6650.0, 1590.0, "fallback"


def load_synthetic():
    """directvoxgo_b200/synthetic.py loaded BY PATH: it is pure torch / numpy, and importing it this way does not
    import the product package (no native .so is mapped) -- the reference arm must not run on, or load, our code."""
    spec = importlib.util.spec_from_file_location("_dvgo_synthetic", os.path.join(ROOT, "directvoxgo_b200", "synthetic.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


class ClockSampler(threading.Thread):
    """Samples SM clocks and throttle reasons with nvidia-smi while the timed regions run."""
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
            self._stop_evt.wait(0.1)

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


def make_batches(n_batches, n_rays, device, rank, syn=None):
    import torch
    if syn is None:
        from directvoxgo_b200 import synthetic as syn
    out = []
    for b in range(n_batches):
        ro, rd, vd, tgt = syn.random_training_rays(n_rays, n_views=100, seed=1000 * (rank + 1) + b, device="cpu")
        out.append(tuple(t.pin_memory() if torch.cuda.is_available() else t for t in (ro, rd, vd, tgt)))
    dev = [tuple(t.to(device, non_blocking=True) for t in b) for b in out]
    return out, dev


def unique_voxels(model, batch, rk):
    """U of SURVEY.md 8d: unique voxels touched as trilinear corners by the step's in-bbox samples (torch.unique)."""
    import torch
    ro, rd, vd, tgt = batch
    with torch.no_grad():
        pts, ray_id, step_id = model.sample_ray(ro, rd, **rk)
        X, Y, Z = (int(s) for s in model.world_size)
        f = (pts - model.xyz_min) / (model.xyz_max - model.xyz_min) * torch.tensor([X - 1, Y - 1, Z - 1], device=pts.device)
        i0 = f.floor().long()
        seen = torch.zeros(X * Y * Z, dtype=torch.bool, device=pts.device)
        for dx in (0, 1):
            for dy in (0, 1):
                for dz in (0, 1):
                    c = i0 + torch.tensor([dx, dy, dz], device=pts.device)
                    ok = ((c >= 0) & (c < torch.tensor([X, Y, Z], device=pts.device))).all(-1)
                    seen[((c[:, 0] * Y + c[:, 1]) * Z + c[:, 2])[ok]] = True
        return int(seen.sum().item()), int(pts.shape[0])


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


def time_trainer(trainer, batches, steps, warm, snap=None):
    """warm untimed steps, restore, `steps` timed steps (CUDA events); returns (ms/step, survivors/step, last loss)."""
    import torch
    for i in range(warm):
        trainer.step(*batches[i % len(batches)])
    if snap is not None:
        trainer.restore(snap)
    torch.cuda.synchronize()
    if hasattr(trainer, "stats_snapshot"):
        trainer.stats_snapshot(reset=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        loss = trainer.step(*batches[i % len(batches)])
    e1.record()
    torch.cuda.synchronize()
    surv = None
    if hasattr(trainer, "stats_snapshot"):
        st = trainer.stats_snapshot(reset=True)
        surv = st[0] / max(st[1], 1)
    return e0.elapsed_time(e1) / steps, surv, float(loss)


def train_sphere_metric(model, rk, cfg, dev_batches, device, steps=100):
    """Extras (not the headline): the fused training step on the sphere-occupancy scene, where the four-mask cascade
    and the early stop cull most samples as on a real scene -- (a) with the bench's dense TV, (b) with the reference's
    actual Blender fine-stage settings: TV off, masked Adam on density + k0 (configs/default.py:53-54,67)."""
    from directvoxgo_b200.fused import FusedTrainer
    n = dev_batches[0][0].shape[0]
    out = {}
    for name, c in (("dense_tv", cfg),
                    ("blender_default_no_tv_masked_adam", dict(cfg, weight_tv_density=0.0, weight_tv_k0=0.0))):
        tr = FusedTrainer(sphere_scene(model, device), c, rk)
        snap = tr.snapshot()
        ms, surv, _ = time_trainer(tr, dev_batches, steps, 50, snap)
        out[name] = {"ms_per_step": ms, "rays_per_s": n / ms * 1e3, "survivors_per_step": surv, "steps": steps}
        del tr
    out["grid"] = "sphere-occupancy (extra)"
    return out


def exact_transmittance_metric(grid, device, dev_batches, steps, warm):
    """Extra (not the headline): the same cfg-2 step from the same random-init state with
    FusedTrainer(exact_transmittance=True), i.e. the forward march replaying the reference's per-sample float T_cum
    recurrence (survivor set / weights bit-identical to the reference kernels) instead of the double product scan."""
    from directvoxgo_b200.fused import FusedTrainer
    fresh, rk2, cfg2 = build_problem(grid, device)
    tr = FusedTrainer(fresh, cfg2, rk2, exact_transmittance=True)
    snap = tr.snapshot()
    k = warm % len(dev_batches)            # the headline pass times its step i on batch (warm + i) mod nb: same order here
    ms, surv, loss = time_trainer(tr, dev_batches[k:] + dev_batches[:k], steps, max(warm, 3), snap)
    n = dev_batches[0][0].shape[0]
    return {"ms_per_step": ms, "rays_per_s": n / ms * 1e3, "survivors_per_step": surv, "steps": steps, "last_loss": loss,
            "what": "cfg2 step with the bit-exact transmittance instantiation of march_fwd (opt-in; default = double scan)"}


def render_metric(model, rk, device, n_frames=2, chunk=65536, label="", rank=0, world=1):
    """Secondary metric of BASELINE.json: ms per rendered 800x800 frame (run.py:57-110: rays of a view ->
    chunks -> forward, render_depth=True), device-timed with CUDA events, rays generated on the device."""
    import torch
    from directvoxgo_b200 import synthetic as syn
    from directvoxgo_b200.fused import FusedRenderer
    from directvoxgo_b200.parallel import shard_views
    renderer = FusedRenderer(model, rk)
    H = W = syn.BLENDER["H"]
    K = syn.intrinsics(H, W)
    # BASELINE config 3: whole views are sharded over the ranks (view i -> rank i mod world), no communication;
    # every rank renders n_frames views, the aggregate is (n_frames * world) frames in the max-over-ranks time.
    poses = syn.random_poses(n_frames * world + 1, seed=4242)
    mine = [poses[1 + i] for i in shard_views(n_frames * world, rank, world)]
    renderer.render_view(H, W, K, poses[0], chunk=chunk)  # warm-up (allocates the workspace)
    torch.cuda.synchronize()
    renderer.stats_snapshot(reset=True)
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
    st = renderer.stats_snapshot(reset=True)
    surv_frame = st[0] / n_frames
    _, tensor_peak, _ = peaks()
    return {"ms_per_frame": ms / (n_frames * world), "ms_per_frame_per_gpu": ms / n_frames, "frames": n_frames * world,
            "frames_per_s": n_frames * world / ms * 1e3, "sharding": "view i -> rank i mod %d, no communication" % world,
            "rays_per_call": chunk, "resolution": "%dx%d" % (H, W), "grid": label,
            "survivors_per_frame": surv_frame, "rgbnet_tflops_per_frame": surv_frame * 43520.0 / 1e12,
            "tensor_frac_of_frame_time": surv_frame * 43520.0 / (ms / n_frames * 1e-3) / 1e12 / tensor_peak,
            "mean_rgb": float(img.mean())}


# ---- the reference's algorithm on the host cores (oracle/model_ref.py) -- never imports the product package ------
def cpu_reference_problem(grid, syn, seed=777):
    """RefDVGO at the bench's fine-stage configuration, built WITHOUT the product package: same bbox, grid size, N(0,1)
    grids from the same generator, default-initialised rgbnet 39 -> 128 -> 128 -> 3 with zero last bias
    (lib/dvgo.py:123-131), act_shift from alpha_init (lib/dvgo.py:57), all-true mask."""
    import math
    import torch
    from oracle.model_ref import RefDVGO
    lo, hi = syn.fine_bbox()
    kw = syn.FINE_MODEL
    torch.manual_seed(seed)
    dim0 = kw["rgbnet_dim"] + 3 + 3 * kw["viewbase_pe"] * 2
    net = [torch.nn.Linear(dim0, kw["rgbnet_width"]), torch.nn.Linear(kw["rgbnet_width"], kw["rgbnet_width"]),
           torch.nn.Linear(kw["rgbnet_width"], 3)]
    torch.nn.init.constant_(net[-1].bias, 0)
    g = torch.Generator(device="cpu").manual_seed(seed)
    density = torch.randn(1, 1, grid, grid, grid, generator=g)
    k0 = torch.randn(1, kw["rgbnet_dim"], grid, grid, grid, generator=g)
    ref = RefDVGO(lo, hi, density, k0, [(l.weight, l.bias) for l in net], None,
                  act_shift=math.log(1 / (1 - kw["alpha_init"]) - 1), voxel_size_ratio=1.0, voxel_size=None,
                  fast_color_thres=kw["fast_color_thres"], viewbase_pe=kw["viewbase_pe"], rgbnet_direct=True)
    return ref, dict(syn.RENDER_KWARGS), dict(syn.FINE_TRAIN)


def cpu_reference_run(grid, n_rays, steps, warmup, threads):
    import torch
    syn = load_synthetic()
    torch.set_num_threads(threads)
    ref, rk, cfg = cpu_reference_problem(grid, syn)
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
    grid, n_rays, split, metric, wl, scaling = WORKLOADS["cfg2"]   # the arm the driver compares is cfg 2
    grid = args.grid or grid
    threads = os.cpu_count() or 1
    steps = max(1, min(args.steps, 2))     # one FULL 8192-ray step is ~10-20 s on the box's host cores
    warm = min(args.warmup, 1)
    rays_s, dt = cpu_reference_run(grid, n_rays, steps, warm, threads)
    loaded = [l.split()[-1] for l in open("/proc/self/maps") if "libdvgo_b200" in l or "directvoxgo_b200/_C" in l]
    line = {"metric": metric, "value": rays_s, "unit": "rays/s", "n_gpus": args.gpus, "steps": steps,
            "warmup": warm, "ms_per_step": dt * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "impl": "reference",
            "config": {"workload": wl % grid, "rays_per_step_per_gpu": n_rays,
                       "sample": "each timed step = one full %d-ray step on the full %d^3 grid (same config as our arm); "
                                 "%d timed step(s) after %d warm-up" % (n_rays, grid, steps, warm),
                       "product_native_code_loaded": sorted(set(loaded))},
            "cpu_baseline": {"value": rays_s, "unit": "rays/s", "cores": threads, "kind": "port",
                             "sample": "%d full %d-ray step(s) on the %d^3 grid; the reference's CUDA ops have no CPU path "
                                       "(CHECK_CUDA), so this is oracle/model_ref.py: its op sequence on torch-CPU + the C oracle"
                                       % (steps, n_rays, grid)},
            "e2e": {"value": rays_s, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def ref_gpu_baseline(grid, batches, device, steps=6):
    """The reference's OWN CUDA kernels (oracle/_ref, compiled from the unmodified lib/cuda sources) + ATen
    grid_sample / index_add / fp32 nn.Linear on this GPU: the op sequence of lib/dvgo.py:450-577 + run.py:377-397 with
    its host syncs (oracle/model_ref.py with oracle/_ref as the op provider).  'Beat THAT kernel on the same box.'"""
    import types
    import torch
    ref_dir = os.path.join(ROOT, "oracle", "_ref")
    if not glob.glob(os.path.join(ref_dir, "ref_render_utils_cuda*.so")):
        return {"unavailable": "oracle/_ref not built"}
    sys.path.insert(0, ref_dir)
    import ref_adam_upd_cuda
    import ref_render_utils_cuda
    import ref_total_variation_cuda
    from oracle import model_ref
    ns = types.SimpleNamespace()
    for mod in (ref_render_utils_cuda, ref_total_variation_cuda, ref_adam_upd_cuda):
        for k in dir(mod):
            if not k.startswith("_"):
                setattr(ns, k, getattr(mod, k))
    model, rk, cfg = build_problem(grid, device)
    ref = model_ref.RefDVGO.from_module(model).to(device)
    del model
    prev = model_ref.set_ops(ns)
    try:
        for i in range(2):
            ref.train_step(*batches[i % len(batches)], rk, cfg)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            loss, _ = ref.train_step(*batches[i % len(batches)], rk, cfg)
        e1.record()
        torch.cuda.synchronize()
    finally:
        model_ref.set_ops(prev)
    ms = e0.elapsed_time(e1) / steps
    n = batches[0][0].shape[0]
    return {"value": n / ms * 1e3, "unit": "rays/s", "ms_per_step": ms, "steps": steps, "kind": "reference CUDA kernels",
            "what": "oracle/_ref (unmodified lib/cuda/*.cu built for sm_100a) + ATen grid_sample/index_add/nn.Linear fp32, "
                    "op sequence and host syncs of lib/dvgo.py + run.py, same GPU, same rays", "last_loss": loss}


def ncu_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from the newest committed ncu --set full summary
    (profiles/r*_ncu_traffic.json, written by tools/ncu_traffic.py from the .ncu-rep of this command)."""
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_ncu_traffic.json")))
    if not files:
        return {}, None
    d = json.load(open(files[-1]))
    return d.get("kernels", {}), {"file": os.path.relpath(files[-1], ROOT), "commit": d.get("commit"),
                                  "workload": d.get("workload")}


def sampling_hbm(hbm_peak):
    """BASELINE metric (iii): achieved HBM GB/s of the sampling (gather) kernels, taken from ncu as SURVEY.md section
    8(d) defines it -- (dram__bytes_read.sum + dram__bytes_write.sum) / gpu__time_duration per launch, from the newest
    committed capture of this command (never from a run under a profiler of THIS process: the table is read, not
    measured here)."""
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_ncu_traffic.json")))
    if not files:
        return None
    d = json.load(open(files[-1]))
    out = {"unit": "GB/s", "source": {"file": os.path.relpath(files[-1], ROOT), "commit": d.get("commit")},
           "definition": "ncu dram bytes (read + write) per launch / ncu gpu__time_duration, cold-cache serialised launches"}
    for name, k in d.get("per_kernel", {}).items():
        if name.startswith(("k0_gather", "march_fwd")) and k.get("time_us"):
            gbs = (k["dram_rd"] + k["dram_wr"]) / (k["time_us"] * 1e-6) / 1e9
            out[name] = {"achieved": gbs, "frac_of_measured_hbm_peak": gbs / hbm_peak, "dram_bytes_per_launch": k["dram_rd"] + k["dram_wr"],
                         "time_us": k["time_us"], "l1_lsu_wavefront_pct": k.get("l1_lsu_wavefront_pct"),
                         "l2_hit_pct": k.get("l2_hit_pct")}
    return out


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
        dist.init_process_group("nccl", device_id=device, timeout=datetime.timedelta(seconds=180))
    import directvoxgo_b200 as pkg
    from directvoxgo_b200.trainer import ModuleTrainer

    grid, rays_total, split, metric, wl, scaling = WORKLOADS[args.workload]
    grid = args.grid or grid
    n_rays = rays_total // world if split else rays_total      # rays per step on THIS rank
    n_global = n_rays * world

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([x], device=device)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def measure(grid, n_rays, steps, warm, ramp_s, with_e2e, with_stages, exchange):
        """Build the problem and run the passes described in the module docstring.  Every rank executes the same
        sequence (the steps contain collectives when world > 1)."""
        model, rk, cfg = build_problem(grid, device)
        path = args.path
        trainer = None
        if path in ("auto", "fused"):
            try:
                from directvoxgo_b200.fused import FusedTrainer
                trainer = FusedTrainer(model, cfg, rk, world_size=world, exchange=exchange)
                path = "fused"
            except ImportError:
                if path == "fused":
                    raise
        if trainer is None:
            trainer = ModuleTrainer(model, cfg, rk, world_size=world)
            path = "module"
        nb = N_BATCHES if n_rays <= 16384 else 4
        host_batches, dev_batches = make_batches(nb, n_rays, device, rank)
        res = {"path": path, "trainer": trainer, "model": model, "rk": rk, "cfg": cfg, "dev_batches": dev_batches}
        if rank == 0 and grid <= 200:
            res["U"], res["M0"] = unique_voxels(model, dev_batches[0], rk)
        torch.cuda.synchronize()
        snap = trainer.snapshot() if hasattr(trainer, "snapshot") else None

        def restore():
            if snap is not None:
                trainer.restore(snap)
            barrier()
            if hasattr(trainer, "stats_snapshot"):
                trainer.stats_snapshot(reset=True)

        # Untimed clock ramp: a B200 idling at its floor clock needs ~1 s of load to reach its boost clock.  Every rank
        # runs the SAME number of rounds: "done" is agreed with a MAX all-reduce after each round.
        t_ramp = time.time()
        while ramp_s > 0:
            for i in range(50 if n_rays <= 16384 else 5):
                trainer.step(*dev_batches[i % nb])
            torch.cuda.synchronize()
            if max_over_ranks(1.0 if time.time() - t_ramp >= ramp_s else 0.0) > 0:
                break
        # ---- pass 1: device-resident ---------------------------------------------------------------------------
        for i in range(warm):
            trainer.step(*dev_batches[i % nb])
        restore()
        l0 = pkg._C.launch_count()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for i in range(steps):
            loss = trainer.step(*dev_batches[(warm + i) % nb])
        ev1.record()
        barrier()
        res["launches"] = pkg._C.launch_count() - l0
        res["ms_per_step"] = max_over_ranks(ev0.elapsed_time(ev1)) / steps
        if hasattr(trainer, "stats_snapshot"):
            st = trainer.stats_snapshot(reset=True)
            res["survivors_value_pass"] = st[0] / max(st[1], 1)
        res["last_loss_value_pass"] = float(loss)
        # ---- pass 2: end to end: pinned host rays in, loss out, every step ----------------------------------------
        if with_e2e:
            stage = [torch.empty_like(x, device=device) for x in host_batches[0]]
            for i in range(min(3, warm)):
                for d, h in zip(stage, host_batches[i % nb]):
                    d.copy_(h, non_blocking=True)
                float(trainer.step(*stage).item())
            restore()
            ev0.record()
            for i in range(steps):
                for d, h in zip(stage, host_batches[(warm + i) % nb]):
                    d.copy_(h, non_blocking=True)
                loss_host = float(trainer.step(*stage).item())
            ev1.record()
            barrier()
            res["e2e_sync_ms"] = max_over_ranks(ev0.elapsed_time(ev1)) / steps     # copy -> step -> read, serialised
            # the public host-fed loop (directvoxgo_b200.trainer.HostFedLoop): the same K copies, K steps and K loss
            # reads, but step i's copy runs under step i-1 and its loss is read after step i+1 was enqueued
            from directvoxgo_b200.trainer import HostFedLoop
            restore()
            loop = HostFedLoop(trainer, host_batches[0], device)
            losses = []
            torch.cuda.synchronize()
            ev0.record()
            for i in range(steps):
                prev = loop.step(host_batches[(warm + i) % nb])
                if prev is not None:
                    losses.append(prev)
            losses.append(loop.drain())
            ev1.record()
            barrier()
            # every step's loss was read, and the pipelined loop reproduces the serialised loop's final loss (same state,
            # same batches; grid gradients are accumulated with unordered atomics, hence a tolerance) -- reported, not fatal
            res["e2e_check"] = {"losses_read": len(losses), "last_loss_serialised": loss_host,
                                "agrees": bool(len(losses) == steps and
                                               abs(losses[-1] - loss_host) <= 1e-3 * max(1.0, abs(loss_host)))}
            res["e2e_ms"] = max_over_ranks(ev0.elapsed_time(ev1)) / steps
            res["e2e_loss"] = losses[-1]
            res["h2d"] = sum(x.numel() * x.element_size() for x in host_batches[0])
        # ---- pass 3: the same K steps with CUDA events between the stages, survivors counted on the device ----------
        if with_stages and path == "fused":
            for i in range(min(3, warm)):
                trainer.step(*dev_batches[i % nb])
            restore()
            trainer.stage_events = {}
            for i in range(steps):
                trainer.step(*dev_batches[(warm + i) % nb])
            torch.cuda.synchronize()
            stages = trainer.stage_times_ms()
            if "sweep_grids" in stages:     # multi-GPU: grid sweeps and the rgbnet Adam are marked separately
                stages["sweep"] = stages["sweep"] + stages.pop("sweep_grids")
            trainer.stage_events = None
            st = trainer.stats_snapshot(reset=True)
            res["stages"] = stages
            res["survivors_per_step"] = st[0] / max(st[1], 1)
            res["survivors_max"] = st[3]
            barrier()
        return res

    sampler = ClockSampler(local)
    sampler.start()
    main = measure(grid, n_rays, args.steps, args.warmup, args.ramp_s, True, True, args.exchange)
    clocks = sampler.stop()  # sampled across the three measured passes
    trainer, model, rk, cfg, path = main["trainer"], main["model"], main["rk"], main["cfg"], main["path"]
    dev_batches = main["dev_batches"]
    ms_per_step = main["ms_per_step"]
    value = n_global / (ms_per_step * 1e-3)

    # ---- secondary measurements (never inside the timed regions) -------------------------------------------------
    render, sphere_extra, cfg5_extra, exact_extra = None, None, None, None
    if not args.no_render and args.workload == "cfg2":      # every rank renders its share of the views
        try:
            fresh, rk2, _ = build_problem(grid, device)      # random-init N(0,1): what BASELINE configs[2] names
            render = {"dense": render_metric(fresh, rk2, device, 2, 65536, "random-init N(0,1) (untrained), nothing culled", rank, world),
                      "sphere": render_metric(sphere_scene(fresh, device), rk2, device, 2, 65536, "sphere-occupancy (extra)", rank, world)}
            del fresh
            if hasattr(trainer, "sync_to_model"):
                trainer.sync_to_model()
                render["after_bench_training"] = render_metric(
                    model, rk, device, 1, 65536, "the bench's model after its %d training steps (extra)" % args.steps, rank, world)
        except Exception as e:  # secondary metric: never let it break the headline line
            if world > 1:
                raise
            render = {"error": repr(e)[:300]}
    barrier()
    if not args.no_extras and world == 1 and path == "fused" and args.workload == "cfg2":
        try:
            fresh, rk2, cfg2 = build_problem(grid, device)
            sphere_extra = train_sphere_metric(fresh, rk2, cfg2, dev_batches, device)
            del fresh
        except Exception as e:
            sphere_extra = {"error": repr(e)[:300]}
        try:
            exact_extra = exact_transmittance_metric(grid, device, dev_batches, args.steps, args.warmup)
        except Exception as e:
            exact_extra = {"error": repr(e)[:300]}
    stages_main, U, M0 = main.get("stages"), main.get("U"), main.get("M0")
    surv = main.get("survivors_per_step")
    launches = main["launches"]
    main_e2e_ms, main_h2d, main_e2e_loss = main["e2e_ms"], main["h2d"], main["e2e_loss"]
    main_e2e_sync_ms = main.get("e2e_sync_ms")
    main_e2e_check = main.get("e2e_check")
    exchange_used = getattr(trainer, "exchange", "nccl all-reduce" if world > 1 else "none") + \
        (" + NVLS multicast" if getattr(trainer, "multicast", False) else "")
    surv_value_pass = main.get("survivors_value_pass")
    # free the main problem before the large-grid extra
    del main, trainer, model, dev_batches
    torch.cuda.empty_cache()
    if not args.no_extras and path == "fused" and args.workload == "cfg2":
        # BASELINE configs[4] as an extra in every line: 320^3, 65 536 rays per step split over the ranks (strong scaling)
        g5, rays5 = WORKLOADS["cfg5"][0], WORKLOADS["cfg5"][1]
        r5 = measure(g5, rays5 // world, 10, 3, 0.0, False, True, args.exchange)
        cfg5_extra = {"workload": WORKLOADS["cfg5"][4] % g5, "scaling": "strong", "n_gpus": world,
                      "rays_per_step_total": rays5, "rays_per_step_per_gpu": rays5 // world, "steps": 10, "warmup": 3,
                      "ms_per_step": r5["ms_per_step"], "rays_per_s": rays5 / (r5["ms_per_step"] * 1e-3),
                      "survivors_per_step_rank0": r5.get("survivors_per_step"), "stages_ms_rank0": r5.get("stages"),
                      "grad_exchange": getattr(r5["trainer"], "exchange", "none")}
        del r5
        torch.cuda.empty_cache()
    barrier()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    hbm_peak, tensor_peak, peak_kind = peaks()
    C, G = 12, grid ** 3
    roof, balg = None, None
    if stages_main is not None and surv is not None:
        M4 = surv                    # survivors per step of exactly the steps whose stages were timed
        Ueff = U if U is not None else G
        # algorithmic work per launch: bytes for the HBM-bound stages, FLOPs for the rgbnet
        alg = {
            "march_fwd": ("hbm", Ueff * (1 + C) * 4 + M4 * (16 + C * 4 + 12)),   # touched cells once + sample stream out
            "mlp_fwd": ("tensor", M4 * 43520.0),
            "mlp_bwd": ("tensor", M4 * 43520.0 * 2),
            "march_bwd": ("hbm", Ueff * (1 + C) * 8 + M4 * (16 + C * 4 + 4)),    # grad cells RMW + sample stream in
            "sweep": ("hbm", G * (1 + C) * 32 / (world if world > 1 else 1)),       # p,g,m,v in; p,m,v,g=0 out (x-slab per rank)
        }
        dom = max((k for k in stages_main if k in alg), key=lambda k: stages_main[k])
        kind, work = alg[dom]
        t = stages_main[dom] * 1e-3
        if kind == "hbm":
            ach, peak, unit = work / t / 1e9, hbm_peak, "GB/s"
        else:
            ach, peak, unit = work / t / 1e12, tensor_peak, "TFLOP/s"
        kernel_names = {"march_fwd": "march_fwd_kernel<12> + k0_gather_kernel<12>", "mlp_fwd": "mlp_fwd_kernel", "mlp_bwd": "mlp_bwd_kernel",
                        "march_bwd": "k0_scatter_kernel<12> + march_bwd_kernel<12>", "sweep": "sweep_kernel (k0) + sweep_kernel (density) + rgbnet Adam"}
        traffic_tab, traffic_src = ncu_traffic()
        traffic = traffic_tab.get(dom) if (grid == 160 and world == 1 and args.workload == "cfg2") else None
        comm = {k: stages_main[k] for k in ("grad_exchange", "param_gather") if k in stages_main}
        roof = {"bound": kind, "kernel": kernel_names[dom], "achieved": ach, "peak": peak, "unit": unit,
                "frac": ach / peak, "traffic": traffic, "traffic_source": traffic_src,
                "peak_kind": peak_kind + (" (bf16 sustained; fp16 runs at the same rate)" if kind == "tensor" else ""),
                "kernel_ms": stages_main[dom], "algorithmic_work_per_launch": work,
                "survivors_per_step": M4, "work_term": "M4 = survivors_per_step counted on the device over the same %d steps "
                                                       "whose stage events give kernel_ms (median over the steps)" % args.steps,
                "limiter_observed": ("ncu, profiles/r02c_ncu_kernels.md: mlp_bwd_kernel moves 0.85 shared-memory wavefronts per "
                                     "SM per cycle (0.49 tensor-core operand fetch + 0.36 LSU) -- shared-memory bandwidth, not "
                                     "the tensor pipe (39 % busy), bounds it") if dom == "mlp_bwd" else None,
                "all_stages": {k: {"ms": stages_main[k], "bound": alg[k][0], "work": alg[k][1],
                                   "frac": (alg[k][1] / (stages_main[k] * 1e-3) / (1e9 * hbm_peak if alg[k][0] == "hbm" else 1e12 * tensor_peak))}
                               for k in stages_main if k in alg},
                "other_stages_ms": {k: v for k, v in stages_main.items() if k not in alg}}
        if comm:
            roof["collectives_ms"] = comm
        balg = Ueff * (1 + C) * 4 * 3 + G * (1 + C) * 4 + G * (1 + C) * 28

    cpu = None
    if world == 1 and not args.no_cpu_baseline and args.workload == "cfg2":
        threads = os.cpu_count() or 1
        rays_s, dt = cpu_reference_run(grid, 8192, 1, 0, threads)
        cpu = {"value": rays_s, "unit": "rays/s", "cores": threads, "kind": "port", "seconds": dt,
               "sample": "one full 8192-ray step on the %d^3 grid, no warm-up (oracle/model_ref.py: torch-CPU + C oracle; "
                         "the reference's CUDA ops have no CPU path)" % grid}
    refgpu = None
    if world == 1 and not args.no_ref_gpu and args.workload == "cfg2":
        try:
            _, dev_b = make_batches(4, 8192, device, rank)
            refgpu = ref_gpu_baseline(grid, dev_b, device)
        except Exception as e:
            refgpu = {"unavailable": repr(e)[:300]}

    line = {
        "metric": metric, "value": value, "unit": "rays/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": scaling,
        "vs_baseline": None, "dtype": "f32 (grids, sampling, compositing, Adam); rgbnet GEMM operands fp16 (saturating), fp32 accumulate",
        "data": "synthetic",
        "config": {"workload": wl % grid, "rays_per_step_per_gpu": n_rays, "rays_per_step_total": n_global, "path": path,
                   "rgbnet": "tc" if path == "fused" else "torch", "parallelism": "ray-sharded dp%d" % world,
                   "samples_emitted_batch0": M0, "unique_voxels_touched_batch0": U,
                   "survivors_per_step": surv, "survivors_per_step_value_pass": surv_value_pass,
                   "grad_exchange": exchange_used,
                   "state": "every measured pass starts from the random-init state (parameters + Adam state restored after "
                            "the clock ramp and the warm-up steps)",
                   "l2_policy": "working set (params+grads+Adam state = %.2f GB) larger than the 126 MB L2; "
                                "%d distinct ray batches cycled" % (G * 13 * 16 / 1e9, N_BATCHES),
                   "algorithmic_bytes_per_step": balg, "clock_ramp_s": args.ramp_s},
        "e2e": {"value": n_global / (main_e2e_ms * 1e-3), "unit": "rays/s", "ms_per_step": main_e2e_ms,
                "h2d_bytes_per_step": main_h2d, "d2h_bytes_per_step": 4, "last_loss": main_e2e_loss,
                "api": "directvoxgo_b200.trainer.HostFedLoop(trainer).step(pinned host batch): every step copies its own "
                       "rays H2D (copy stream, under the previous step) and its own loss D2H (read one call later)",
                "serialised_ms_per_step": main_e2e_sync_ms, "check": main_e2e_check},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": roof,
        "step_hbm_frac": (balg / (ms_per_step * 1e-3) / 1e9 / hbm_peak) if balg else None,
        "sampling_hbm": sampling_hbm(hbm_peak) if (grid == 160 and world == 1 and args.workload == "cfg2") else None,
        "cpu_baseline": cpu,
        "ref_gpu_baseline": refgpu,
        "train_sphere_extra": sphere_extra,
        "cfg5_extra": cfg5_extra,
        "exact_transmittance_extra": exact_extra,
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
