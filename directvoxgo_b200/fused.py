"""The fused B200 path: FusedTrainer (one training iteration of run.py:372-397) and FusedRenderer
(one chunk of run.py:91-98), built on include/dvgo_b200_fused.h.

Per step the trainer launches ~12 kernels and never synchronises with the host (the reference's
forward alone does ~120 launches and 17 syncs):

    ray_setup(2) -> march_fwd -> rgb (rgbnet) -> composite -> ray_finish -> sample_grad
      -> rgbnet backward -> march_bwd -> [NCCL all-reduce of grid grads when ray-sharded]
      -> sweep(density) -> sweep(k0) -> Adam(rgbnet)

State lives in the trainer's own buffers: density [X,Y,Z], k0 channel-last [X,Y,Z,C] (two copies,
ping-ponged by the TV sweep), persistent zero-initialised gradient accumulators (re-zeroed inside
the sweep), Adam moments in the same layouts.  `sync_to_model()` converts back to the reference's
[1,C,X,Y,Z] parameters so checkpoints / state_dicts stay interchangeable (run.py:420-437).

rgbnet modes
    'torch'  the MLP runs as plain cuBLAS fp32 GEMMs through torch autograd (exact-fp32 parity mode;
             needs one host read of the survivor count per step)
    'tc'     hand-written tcgen05 (TF32 tensor-core, fp32 accumulate) forward+backward kernels with
             in-TMEM weight-gradient accumulation; no host sync   [fused_mlp.cu]
"""

import os

import torch

from . import adam_upd_cuda, ext


def _scene_of(model, rk, ndc=False, ndc_samples=0, exact_transmittance=False):
    X, Y, Z = (int(s) for s in model.density.shape[2:])
    C = int(model.k0.shape[1])
    mc = model.mask_cache
    stepdist = 0.0 if ndc else float(rk["stepsize"] * model.voxel_size)
    return ext.Scene(X, Y, Z, C, model.xyz_min.contiguous(), model.xyz_max.contiguous(),
                     mc.mask if mc is not None else None,
                     mc.xyz2ijk_scale if mc is not None else None,
                     mc.xyz2ijk_shift if mc is not None else None,
                     float(rk["near"]), float(rk["far"]), stepdist, float(model.act_shift),
                     float(rk["stepsize"] * model.voxel_size_ratio), float(model.fast_color_thres),
                     bool(ndc), int(ndc_samples), bool(exact_transmittance))


class _Workspace:
    """Preallocated per-call buffers for N rays (capacities from the scene's max steps per ray)."""

    def __init__(self, scene, n_rays, C, device, train):
        f32 = dict(dtype=torch.float32, device=device)
        i32 = dict(dtype=torch.int32, device=device)
        cap = n_rays * scene.max_steps()
        if cap >= 2 ** 31:      # slot offsets (ray_off) and survivor indices are int32 on the device
            raise ValueError("fused path: %d rays x %d steps per ray exceeds the int32 slot index range; render / train "
                             "in smaller ray chunks, or check near / far / stepsize" % (n_rays, scene.max_steps()))
        self.n_rays, self.cap = n_rays, cap
        self.t_min = torch.empty(n_rays, **f32)
        self.n_steps = torch.empty(n_rays, **i32)
        self.ray_off = torch.empty(n_rays + 1, **i32)
        # per-slot record of the forward march, read by march_bwd only: a rendering workspace passes empty tensors and
        # march_fwd skips those 16 B / sample of stores (and 16 B x cap of memory)
        slots = cap if train else 0
        self.slot_alpha = torch.empty(slots, **f32)
        self.slot_T = torch.empty(slots, **f32)
        self.slot_expd = torch.empty(slots, **f32)
        self.slot_code = torch.empty(slots, **i32)
        self.feat = torch.empty(cap, C, **f32)
        self.s_ray = torch.empty(cap, **i32)
        self.s_slot = torch.empty(cap, **i32)
        self.s_weight = torch.empty(cap, **f32)
        self.s_pos = torch.empty(cap, 4, **f32)   # per survivor: continuous voxel coordinates + ray index (int bits)
        self.rgb = torch.empty(cap, 3, **f32)
        self.alphainv_last = torch.empty(n_rays, **f32)
        # one zero-able block: counters(2) | loss(2) | rgb_acc(3N) | depth_acc(N)
        self.zblock = torch.zeros(4 + 4 * n_rays, **f32)
        self.counters = self.zblock[0:2].view(torch.int32)
        self.loss_acc = self.zblock[2:4]
        self.rgb_acc = self.zblock[4:4 + 3 * n_rays].view(n_rays, 3)
        self.depth_acc = self.zblock[4 + 3 * n_rays:4 + 4 * n_rays]
        # device-side running statistics of the calls made on this workspace, folded in by step_begin at the START of
        # the next call (no host sync): [sum survivors, calls, status bits (1 = a capacity was too small, 2 = non-finite
        # rgbnet value), max survivors of a call]; `pending`: the last call's counters are not folded in yet
        self.stats = torch.zeros(4, dtype=torch.int64, device=device)
        self.pending = False
        if train:
            self.G = torch.empty(n_rays, 3, **f32)
            self.g_last = torch.empty(n_rays, **f32)
            self.d_rgb = torch.empty(cap, 3, **f32)
            self.d_w = torch.empty(cap, **f32)
            self.d_feat = torch.empty(cap, C, **f32)

    def tiles(self, C, pe_stride, train):
        """Survivor tiles of the tensor-core rgbnet (fp16, operand layout): X~ written by k0_gather_tiles, dZ3 by
        sample_grad; zero-initialised once (rows past the count are only ever rewritten with zeros)."""
        if getattr(self, "_xt_key", None) != (C, pe_stride):
            self.xt = torch.zeros(ext.mlp_xtile_bytes(self.cap, C, pe_stride), dtype=torch.uint8, device=self.rgb.device)
            self._xt_key = (C, pe_stride)
        if train and getattr(self, "dzt", None) is None:
            self.dzt = torch.zeros(ext.mlp_dztile_bytes(self.cap), dtype=torch.uint8, device=self.rgb.device)
        return self.xt


def view_embedding(viewdirs, viewfreq):
    """[N, 3 + 6*len(viewfreq)] = cat(viewdirs, sin(v*f), cos(v*f)) (lib/dvgo.py:524-525)."""
    emb = (viewdirs.unsqueeze(-1) * viewfreq).flatten(-2)
    return torch.cat([viewdirs, emb.sin(), emb.cos()], -1)


SUPPORTED_C = (3, 4, 6, 8, 9, 12, 16)   # k0 channel counts the fused kernels are instantiated for (DVGO_DISPATCH_C)


class _FusedBase:
    def __init__(self, model, render_kwargs, mlp="auto", exact_transmittance=False):
        """exact_transmittance: the forward march replays the reference's per-sample `float T_cum` recurrence
        (render_utils_kernel.cu:447-451) instead of its double product scan -- T, weights, alphainv_last, the early-stop
        index and the survivor set are then bit-exact with the reference kernels (default: scan, T / weights rel 5e-6)."""
        self.model = model
        self.rk = dict(render_kwargs)
        self.device = model.density.device
        self.ndc = hasattr(model, "mpi_depth")
        ndc_samples = int((model.mpi_depth - 1) / self.rk["stepsize"]) + 1 if self.ndc else 0
        self.exact_transmittance = bool(exact_transmittance)
        self.scene = _scene_of(model, self.rk, self.ndc, ndc_samples, self.exact_transmittance)
        self.X, self.Y, self.Z = (int(s) for s in model.density.shape[2:])
        self.C = int(model.k0.shape[1])
        if self.C not in SUPPORTED_C:
            raise NotImplementedError(
                "fused path: k0 with %d channels (rgbnet_dim) is not instantiated; supported: %s -- use the op-by-op "
                "module path (directvoxgo_b200.dvgo.DirectVoxGO) for other widths" % (self.C, SUPPORTED_C))
        if mlp == "auto":
            mlp = "tc" if (model.rgbnet is not None and hasattr(ext, "mlp_fwd") and self._tc_supported(model)) else "torch"
        if mlp == "tc" and not hasattr(ext, "mlp_fwd"):
            raise ImportError("tensor-core rgbnet kernels are not built")
        self.mlp_mode = mlp
        self.split_k0 = True
        # tensor-core rgbnet: False (default) = k0_gather_tiles -> X~ tiles in HBM -> mlp_fwd (bulk copies); True = the k0
        # gather inside the forward kernel (mlp_fwd_gather: four producer warps per CTA fill the shared-memory tile).
        # Measured on B200 the fused form is SLOWER (dense 800x800 render 41.5 vs 28.6 ms; training forward 0.65 vs
        # 0.33 ms): the gather is L1-wavefront bound and needs ~24 resident warps per SM to keep enough loads in flight,
        # the forward kernel's register budget leaves room for 8.  Kept as an opt-in experiment.
        self.fuse_gather = os.environ.get("DVGO_FUSE_GATHER", "0") == "1"
        if model.rgbnet is not None and not getattr(model, "rgbnet_direct", True):
            if mlp == "tc":
                raise NotImplementedError("tc rgbnet implements rgbnet_direct=True (the configs' default)")
        self.density = model.density.detach().reshape(self.X, self.Y, self.Z).contiguous().clone()
        self.k0 = ext.ncdhw_to_cl(model.k0.detach().contiguous())
        self._ws = {}
        self._density_swept = False

    def _viewfreq(self):
        vf = getattr(self.model, "viewfreq", None)     # DirectMPIGO with viewbase_pe=0 has an empty table
        return vf if vf is not None else torch.zeros(0, device=self.device)

    @staticmethod
    def _tc_supported(model):
        lin = [m for m in model.rgbnet.modules() if isinstance(m, torch.nn.Linear)]
        return (len(lin) == 3 and lin[0].out_features <= 128 and lin[1].out_features == lin[0].out_features and
                getattr(model, "rgbnet_direct", True) and getattr(model, "posbase_pe", 0) == 0)

    def _workspace(self, n_rays, train, slot=0):
        """slot: independent workspaces of the same shape (render_view pipelines its chunks over two streams)."""
        key = (n_rays, train, slot)
        if key not in self._ws:
            self._ws[key] = _Workspace(self.scene, n_rays, self.C, self.device, train)
        return self._ws[key]

    def stats_snapshot(self, reset=False):
        """(survivors summed over the completed calls, number of calls, status bits, max survivors) over all workspaces
        -- one host read per workspace.  Includes calls whose counters have not been folded in yet.  Raises on a
        non-zero status."""
        st = [0, 0, 0, 0]
        for ws in self._ws.values():
            w = [int(v) for v in ws.stats.tolist()]
            if ws.pending:
                c = [int(v) for v in ws.counters.tolist()]
                w = [w[0] + c[0], w[1] + 1, w[2] | c[1], max(w[3], c[0])]
            st = [st[0] + w[0], st[1] + w[1], st[2] | w[2], max(st[3], w[3])]
            if reset:
                ws.counters.zero_()
                ws.stats.zero_()
                ws.pending = False
        self._raise_on_status(st[2])
        return tuple(st)

    @staticmethod
    def _raise_on_status(bits):
        if bits & 1:
            raise RuntimeError("fused path: a sample / survivor capacity was too small and samples were dropped "
                               "(near / far / stepsize changed after the workspace was sized?)")
        if bits & 2:
            raise FloatingPointError("fused path: non-finite value in the tensor-core rgbnet (NaN input, or weights / "
                                     "features far outside FP16 range); use mlp='torch' for the exact fp32 rgbnet")

    def check_status(self):
        """Host check of the device status bits; call where the host synchronises anyway."""
        self.stats_snapshot()

    def _march(self, ws, rays_o, rays_d, pe=None):
        """pe (tensor-core rgbnet): (padded view-embedding table, its fp16 row form) as TensorCoreMLP.embed(.., C) returns
        them; the k0 gather then writes the rgbnet's X~ tiles (ws.xt) instead of the fp32 feature stream ws.feat."""
        ext.step_begin(ws.zblock, ws.stats if ws.pending else None)   # folds the previous call's counters, then zeroes
        ws.pending = True
        ext.ray_setup(self.scene, rays_o, rays_d, ws.t_min, ws.n_steps, ws.ray_off)
        # density march (scan-bound, per ray) then the k0 gather over the compacted stream (fully parallel)
        ext.march_fwd(self.scene, rays_o, rays_d, self.density, None if (self.split_k0 or pe is not None) else self.k0, ws.t_min,
                      ws.n_steps, ws.ray_off, ws.slot_alpha, ws.slot_T, ws.slot_expd, ws.slot_code, ws.feat, ws.s_ray,
                      ws.s_slot, ws.s_weight, ws.alphainv_last, ws.counters, ws.s_pos)
        if pe is not None and self.fuse_gather:
            pass      # the gather happens inside _rgb_tc (mlp_fwd_gather)
        elif pe is not None:
            ext.k0_gather_tiles(self.scene, rays_o, rays_d, self.k0, ws.t_min, ws.ray_off, ws.s_ray, ws.s_slot,
                                ws.counters, ws.s_pos, pe[1], pe[0].shape[1],
                                ws.tiles(self.C, pe[0].shape[1], hasattr(ws, "d_feat")))
        elif self.split_k0:
            ext.k0_gather(self.scene, rays_o, rays_d, self.k0, ws.t_min, ws.ray_off, ws.s_ray, ws.s_slot,
                          ws.counters, ws.feat)

    def _rgb_tc(self, ws, pe, train, wp=None):
        """Tensor-core rgbnet forward over the survivor stream -> ws.rgb (and, when training, the X~ tiles in ws.xt).
        wp: weight tiles already packed for this (C, pe_stride) (render_view packs once per view)."""
        pe_stride = pe[0].shape[1]
        if self.fuse_gather:
            self._tc.forward_gather(self.scene, self.k0, ws.s_pos, pe[1], self.C, pe_stride, ws.counters, ws.cap, ws.rgb,
                                    ws.tiles(self.C, pe_stride, True) if train else None, wp)
        else:
            self._tc.forward_tiles(ws.xt, self.C, pe_stride, ws.counters, ws.cap, ws.rgb, wp)

    def _rgb_torch(self, ws, viewdirs, m4, grad):
        """rgbnet through torch/cuBLAS fp32 on the first m4 survivors; returns (rgb, feat leaf)."""
        model = self.model
        feat = ws.feat[:m4].detach()
        if grad:
            feat.requires_grad_(True)
        pe = view_embedding(viewdirs, model.viewfreq)[ws.s_ray[:m4].long()]
        if getattr(model, "rgbnet_direct", True):
            rgb = torch.sigmoid(model.rgbnet(torch.cat([feat, pe], -1)))
        else:
            rgb = torch.sigmoid(model.rgbnet(torch.cat([feat[:, 3:], pe], -1)) + feat[:, :3])
        return rgb, feat

    @torch.no_grad()
    def sync_to_model(self):
        """Write the trainer-owned parameters back into the reference-layout module ([1,C,X,Y,Z]).  OWNERSHIP: between
        construction and this call the trainer's buffers are the truth and `model.density / model.k0` are stale;
        model-side edits made in between are NOT seen by the trainer (use `sync_from_model()` after them)."""
        self.check_status()
        self.model.density.data.copy_(self.density.reshape(self.model.density.shape))
        self.model.k0.data.copy_(ext.cl_to_ncdhw(self.k0))
        return self.model

    @torch.no_grad()
    def sync_from_model(self):
        """Re-read density / k0 (and the tensor-core rgbnet copy) from the module after a model-side edit
        (maskout_near_cam_vox, density.data.sub_(1), load_state_dict ...).  Grid shapes must be unchanged."""
        self.density.copy_(self.model.density.detach().reshape(self.X, self.Y, self.Z))
        self.k0.copy_(ext.ncdhw_to_cl(self.model.k0.detach().contiguous()))
        if getattr(self, "_tc", None) is not None:
            self._tc.refresh_from_module()

    @torch.no_grad()
    def update_occupancy_cache(self):
        """The periodic occupancy refresh of the training loop (run.py:330-332) on the TRAINER's live density
        (the module's copy is stale during fused training): mask &= maxpool3(alpha(density)) > fast_color_thres,
        in place in the mask tensor the fused kernels already read."""
        m = self.model.mask_cache.mask
        assert tuple(m.shape) == (self.X, self.Y, self.Z), "the refresh needs the mask at the density resolution"
        m.copy_(ext.alpha_maxpool_mask(self.density.reshape(1, 1, self.X, self.Y, self.Z), float(self.model.act_shift),
                                       float(self.model.voxel_size_ratio), float(self.model.fast_color_thres), m))


class FusedRenderer(_FusedBase):
    """Forward-only rendering of a chunk of rays: rgb_marched [N,3], depth [N], alphainv_last [N]
    (the keys run.py:89,98-99 keeps)."""

    @torch.no_grad()
    def render(self, rays_o, rays_d, viewdirs, render_depth=True, slot=0, out=None):
        """out: optional (rgb [n,3], alphainv_last [n], depth [n] or None) views to write the results into (render_view
        hands in slices of the frame buffers); default: fresh tensors."""
        n = rays_o.shape[0]
        ws = self._workspace(n, False, slot)
        tc = self.model.rgbnet is not None and self.mlp_mode == "tc"
        pe = self._tc_embed(viewdirs, slot) if tc else None
        self._march(ws, rays_o.contiguous(), rays_d.contiguous(), pe)
        if self.model.rgbnet is None:
            ext.rgb_direct(ws.feat, ws.counters, ws.rgb)
        elif tc:
            self._rgb_tc(ws, pe, False, getattr(self, "_view_wp", None))
        else:
            m4 = int(ws.counters[0].item())
            if m4:
                rgb, _ = self._rgb_torch(ws, viewdirs, m4, False)
                ws.rgb[:m4].copy_(rgb)
        ext.composite(ws.rgb, ws.s_weight, ws.s_ray, ws.s_slot, ws.ray_off, ws.counters, ws.rgb_acc,
                      ws.depth_acc if render_depth else None)
        ext.ray_finish(ws.rgb_acc, ws.alphainv_last, None, float(self.rk["bg"]), n, n, 1.0, 0.0, None, None, None)
        if getattr(self, "check_every_call", False):
            self.check_status()
        if out is not None:
            out[0].copy_(ws.rgb_acc)
            out[1].copy_(ws.alphainv_last)
            if render_depth:
                out[2].copy_(ws.depth_acc)
            return None
        res = {"rgb_marched": ws.rgb_acc.clone(), "alphainv_last": ws.alphainv_last.clone()}
        if render_depth:
            res["depth"] = ws.depth_acc.clone()
        return res

    @torch.no_grad()
    def render_view(self, H, W, K, c2w, ndc=False, inverse_y=False, flip_x=False, flip_y=False, chunk=65536,
                    render_depth=True, streams=2):
        """One full view (run.py:82-99): the rays of each chunk of pixels are generated on the device straight
        into a reusable [chunk,3] workspace (`dvgo_rays_of_view`, lib/ray_utils.py:80-85) instead of three
        [H*W,3] tensors per view, then rendered.  Returns [H,W,3] rgb, [H,W] depth / alphainv_last.

        streams = 2: consecutive chunks run on two CUDA streams with a workspace each, so the kernels of chunk i + 1
        (march: L1 / latency bound; k0 gather: L1-wavefront bound) fill the SM resources the tensor-core rgbnet kernel
        of chunk i leaves idle (its two 256-thread CTAs hold 41 K of the 64 K registers and little L1 bandwidth)
        instead of queueing behind it; streams = 1 is the serial order."""
        from .ray_utils import make_view
        view = make_view(H, W, K, c2w, ndc, inverse_y, flip_x, flip_y)
        n = H * W
        chunk = min(chunk, n)
        n_slots = max(1, min(int(streams), (n + chunk - 1) // chunk))
        if getattr(self, "_view_rays", None) is None or self._view_rays[0][0].shape[0] != chunk or \
                len(self._view_rays) < n_slots:
            self._view_rays = [[torch.empty(chunk, 3, device=self.device) for _ in range(3)] for _ in range(n_slots)]
        rgb = torch.empty(n, 3, device=self.device)
        last = torch.empty(n, device=self.device)
        depth = torch.empty(n, device=self.device) if render_depth else None
        cur = torch.cuda.current_stream(self.device)
        if self.model.rgbnet is not None and self.mlp_mode == "tc":
            # the weights do not change during a view: pack the fp16 tiles once, before the streams fork
            self._tc_embed(torch.zeros(1, 3, device=self.device))           # creates self._tc on first use
            P = 3 + 6 * int(self._viewfreq().numel())
            self._view_wp = self._tc.pack(self.C, (P + 1 + 3) // 4 * 4)
        if n_slots > 1:
            if len(getattr(self, "_side_streams", [])) < n_slots:
                self._side_streams = [torch.cuda.Stream(self.device) for _ in range(n_slots)]
            for st in self._side_streams[:n_slots]:
                st.wait_stream(cur)
        for ci, p0 in enumerate(range(0, n, chunk)):
            m = min(chunk, n - p0)
            slot = ci % n_slots
            ro, rd, vd = self._view_rays[slot]
            with torch.cuda.stream(self._side_streams[slot] if n_slots > 1 else cur):
                ext.rays_of_view(view, p0, m, ro, rd, vd)
                self.render(ro[:m], rd[:m], vd[:m], render_depth, slot,
                            (rgb[p0:p0 + m], last[p0:p0 + m], depth[p0:p0 + m] if render_depth else None))
        if n_slots > 1:
            for st in self._side_streams[:n_slots]:
                cur.wait_stream(st)
        self._view_wp = None
        res = {"rgb_marched": rgb.reshape(H, W, 3), "alphainv_last": last.reshape(H, W)}
        if render_depth:
            res["depth"] = depth.reshape(H, W)
        return res

    def _tc_embed(self, viewdirs, slot=0):
        from .fused_mlp import TensorCoreMLP
        if not hasattr(self, "_tc"):
            self._tc = TensorCoreMLP(self.model.rgbnet, self.device)
        return self._tc.embed(viewdirs, self._viewfreq(), self.C)


class FusedTrainer(_FusedBase):
    def __init__(self, model, cfg_train, render_kwargs, world_size=1, dist_group=None, mlp="auto",
                 betas=(0.9, 0.99), eps=1e-8, rank=None, shard_sweep=True, global_step=0, exchange="auto",
                 exact_transmittance=False):
        super().__init__(model, render_kwargs, mlp, exact_transmittance)
        self.cfg = dict(cfg_train)
        self.world_size = world_size
        self.dist_group = dist_group
        self.shard_sweep = shard_sweep
        if rank is None and world_size > 1:
            import torch.distributed as dist
            rank = dist.get_rank(dist_group)
        self.rank = rank or 0
        self.betas, self.eps = betas, eps
        self.global_step = global_step        # resume point: lr pre-decayed as lib/utils.py:21-22 does
        self.opt_step = 0
        z = torch.zeros_like
        self.density_next = torch.empty_like(self.density)
        self.k0_next = torch.empty_like(self.k0)
        self.g_density, self.m_density, self.v_density = z(self.density), z(self.density), z(self.density)
        self.g_k0, self.m_k0, self.v_k0 = z(self.k0), z(self.k0), z(self.k0)
        self.per_lr = None
        self.lr = {k: float(self.cfg.get("lrate_" + k, 0.0)) for k in ("density", "k0", "rgbnet")}
        decay_steps = self.cfg.get("lrate_decay", 20) * 1000
        self.decay = 0.1 ** (1.0 / decay_steps)
        for k in self.lr:
            self.lr[k] *= 0.1 ** (global_step / decay_steps)
        skip = self.cfg.get("skip_zero_grad_fields", []) or []
        self.masked = {k: (k in skip) for k in ("density", "k0")}
        self.rgbnet_state = {}
        self.stage_events = None  # set to {} to record per-stage CUDA events (bench.py roofline pass)
        if model.rgbnet is not None and self.mlp_mode == "tc":
            from .fused_mlp import TensorCoreMLP
            self._tc = TensorCoreMLP(model.rgbnet, self.device, train=True)
        # gradient exchange of ray-sharded data parallel training:
        #   "peer": ONE kernel per grid = reduce-scatter + TV/Adam sweep + all-gather over NVLink peer memory
        #           (torch symmetric memory gives every rank the device addresses of every rank's buffers);
        #   "nccl": reduce_scatter -> slab sweep -> all_gather (or plain all-reduce when shard_sweep=False).
        self._pp = None
        self._flags = None
        self._side, self._zero_pending = None, False
        self._side2, self.overlap_density_bwd = None, True
        self.exchange = "none"
        if world_size > 1:
            self.exchange = "nccl"
            if exchange in ("auto", "peer", "peer_p2p") and shard_sweep and world_size <= 8:
                # "peer_p2p": per-peer loads and stores; "peer": NVLS multicast (multimem.ld_reduce / multimem.st through
                # the switch); "auto": multicast from 4 ranks up.  Measured on B200 (round 2, profiles/r02_scaling.md):
                # 2 ranks p2p is ahead (same bytes per link, plain accesses run faster); 4 ranks 1.90 vs 1.95 ms/step
                # (320^3: 9.20 vs 9.41); 8 ranks 1.90 vs 2.03 ms/step (320^3: 6.62 vs 7.66) in favour of multicast.
                use_mc = exchange == "peer" or (exchange == "auto" and world_size >= 4)
                self._setup_peer(required=(exchange != "auto"), multicast=use_mc)

    def _setup_peer(self, required=False, multicast=True):
        """Re-home the six exchanged buffers (both parameter ping-pong buffers and the gradient accumulators of
        density and k0) in symmetric memory and collect every rank's device addresses of them."""
        import torch.distributed as dist
        ok, pp, mc, keep, err = True, {}, {}, [], None
        try:
            import torch.distributed._symmetric_memory as symm
            group = self.dist_group if self.dist_group is not None else dist.group.WORLD
            new = {}
            for name in ("density", "density_next", "k0", "k0_next", "g_density", "g_k0"):
                old = getattr(self, name)
                t = symm.empty(tuple(old.shape), dtype=old.dtype, device=self.device)
                t.copy_(old)
                hdl = symm.rendezvous(t, group)
                pp[name] = [int(p) for p in hdl.buffer_ptrs]
                assert len(pp[name]) == self.world_size and pp[name][self.rank] == t.data_ptr()
                mc[name] = int(hdl.multicast_ptr or 0) if multicast else 0
                keep.append(hdl)
                new[name] = t
        except Exception as e:  # symmetric memory unavailable on this system: keep the NCCL exchange
            ok, err = False, e
        flag = torch.tensor([1.0 if ok else 0.0], device=self.device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.dist_group)
        if float(flag.item()) < 1.0:
            if required:
                raise RuntimeError("peer-memory gradient exchange unavailable: %r" % (err,))
            return
        for name, t in new.items():
            setattr(self, name, t)
        # multicast is used only if every buffer on every rank got a multicast mapping
        have = torch.tensor([1.0 if all(mc.values()) else 0.0], device=self.device)
        dist.all_reduce(have, op=dist.ReduceOp.MIN, group=self.dist_group)
        if float(have.item()) < 1.0:
            mc = {k: 0 for k in mc}
        self._pp, self._mc, self._symm_handles = pp, mc, keep
        self.multicast = bool(mc["k0"])
        self._bar = torch.zeros(1, device=self.device)
        self.exchange = "peer"
        # Ordering without a collective library: per-rank flag arrays in symmetric memory (ext.peer_barrier), and --
        # with the tensor-core rgbnet, whose gradient is one flat buffer -- that buffer too, so that its sum over ranks
        # is read inside the Adam kernel (ext.adam_upd_peer).  Round 1 used two NCCL all-reduces per step for this
        # (~55 + ~70 us at 8 ranks).  mlp='torch' keeps its autograd gradients and the NCCL all-reduce.
        self._flags = None
        if self.model.rgbnet is None or self.mlp_mode == "tc":
            ok2 = True
            try:
                flags = symm.empty((8,), dtype=torch.int32, device=self.device)
                flags.zero_()
                fh = symm.rendezvous(flags, group)
                fptrs = [int(p) for p in fh.buffer_ptrs]
                gptrs, gmc, gflat = None, 0, None
                if self.model.rgbnet is not None:
                    gflat = symm.empty(tuple(self._tc.grad_flat.shape), dtype=torch.float32, device=self.device)
                    gflat.zero_()
                    gh = symm.rendezvous(gflat, group)
                    gptrs = [int(p) for p in gh.buffer_ptrs]
                    gmc = int(gh.multicast_ptr or 0) if self.multicast else 0
                    keep.append(gh)
                keep.append(fh)
            except Exception:
                ok2 = False
            f2 = torch.tensor([1.0 if ok2 else 0.0], device=self.device)
            dist.all_reduce(f2, op=dist.ReduceOp.MIN, group=self.dist_group)
            if float(f2.item()) >= 1.0:
                torch.cuda.synchronize()
                dist.barrier(group=self.dist_group)          # every rank's flags are zeroed before anyone signals
                self._flags, self._flag_ptrs, self._epoch = flags, fptrs, 0
                if gflat is not None:
                    self._tc.grad_flat = gflat
                    self._gflat_ptrs, self._gflat_mc = gptrs, gmc

    def set_pervoxel_lr(self, count):
        """View-count learning-rate table for the density grid (lib/masked_adam.py:35-37)."""
        self.per_lr = (count.float() / count.max()).reshape(self.X, self.Y, self.Z).contiguous()

    def _mark(self, name):
        if self.stage_events is not None:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            self.stage_events.setdefault(name, []).append(ev)

    def stage_times_ms(self):
        """Median duration of each stage between consecutive marks over the recorded steps (after a synchronize).
        The median, not the mean: with an event recorded between every two kernels the host occasionally falls behind
        the GPU, and the idle gap lands in whichever stage was waiting for its launch (seen as +0.08 ms on the first
        stage of one run in three); a kernel's own duration does not vary by more than ~1 % from step to step."""
        names = list(self.stage_events.keys())
        out = {}
        for a, b in zip(names[:-1], names[1:]):
            ts = sorted(x.elapsed_time(y) for x, y in zip(self.stage_events[a], self.stage_events[b]))
            n = len(ts)
            out[b] = 0.0 if n == 0 else (ts[n // 2] if n % 2 else 0.5 * (ts[n // 2 - 1] + ts[n // 2]))
        return out

    def _join_side(self):
        if self._zero_pending:
            torch.cuda.current_stream().wait_event(self._zero_done)
            self._zero_pending = False

    # -- state snapshot / restore (bench.py: the clock ramp and warm-up must not train the model that is timed) -----
    @torch.no_grad()
    def snapshot(self):
        snap = {"t": {n: getattr(self, n).clone() for n in ("density", "k0", "m_density", "v_density", "m_k0", "v_k0")},
                "opt_step": self.opt_step, "global_step": self.global_step, "lr": dict(self.lr)}
        if getattr(self, "_tc", None) is not None:
            snap["tc"] = [self._tc.params.clone(), self._tc.exp_avg.clone(), self._tc.exp_avg_sq.clone()]
        elif self.model.rgbnet is not None:
            snap["rgbnet"] = [p.detach().clone() for p in self.model.rgbnet.parameters()]
        return snap

    @torch.no_grad()
    def restore(self, snap):
        """In place (the buffers may live in symmetric memory that peers hold addresses of)."""
        self._join_side()
        for n, t in snap["t"].items():
            getattr(self, n).copy_(t)
        for g in (self.g_density, self.g_k0):
            g.zero_()
        self.opt_step, self.global_step, self.lr = snap["opt_step"], snap["global_step"], dict(snap["lr"])
        if "tc" in snap:
            for dst, src in zip((self._tc.params, self._tc.exp_avg, self._tc.exp_avg_sq), snap["tc"]):
                dst.copy_(src)
        elif "rgbnet" in snap:
            for p, src in zip(self.model.rgbnet.parameters(), snap["rgbnet"]):
                p.data.copy_(src)
            self.rgbnet_state = {}

    # -- the step -----------------------------------------------------------------------------------
    def step(self, rays_o, rays_d, viewdirs, target):
        cfg, model = self.cfg, self.model
        n = rays_o.shape[0]
        n_global = n * self.world_size
        self.global_step += 1
        ws = self._workspace(n, True)
        rays_o, rays_d, target = rays_o.contiguous(), rays_d.contiguous(), target.contiguous()
        tc = model.rgbnet is not None and self.mlp_mode == "tc"
        self._mark("start")
        pe = self._tc.embed(viewdirs, self._viewfreq(), self.C) if tc else None
        self._mark("embed")
        self._march(ws, rays_o, rays_d, pe)
        self._mark("march_fwd")

        dens_async = False
        w_main = float(cfg.get("weight_main", 1.0))
        w_ent = float(cfg.get("weight_entropy_last", 0.0))
        w_per = float(cfg.get("weight_rgbper", 0.0))
        bg = float(self.rk["bg"])


        def density_backward_async():
            """alpha2weight / raw2alpha backward + density scatter (march_bwd) depend on d_w from the loss kernels only,
            not on the rgbnet backward: launch them on a side stream so they run under it (the rgbnet kernel holds one
            CTA per SM for its shared memory but leaves registers and issue slots; march_bwd uses no shared memory)."""
            if not self.split_k0 or not self.overlap_density_bwd or self.stage_events is not None:
                return False    # (per-stage event timing wants the stages one after the other)
            if self._side2 is None:
                self._side2 = torch.cuda.Stream(device=self.device)
                self._dens_done = torch.cuda.Event()
            ev = torch.cuda.Event()
            ev.record()
            self._side2.wait_event(ev)
            if self._zero_pending:
                self._side2.wait_event(self._zero_done)
            with torch.cuda.stream(self._side2):
                ext.march_bwd(self.scene, rays_o, rays_d, ws.t_min, ws.n_steps, ws.ray_off, ws.slot_alpha, ws.slot_T,
                              ws.slot_expd, ws.slot_code, ws.d_feat, ws.d_w, ws.alphainv_last, ws.g_last, self.g_density,
                              None)
                if self.world_size == 1:
                    # single GPU: the density gradient is final here, so its TV + Adam sweep (27 us at 160^3) follows on
                    # the side stream, under the rgbnet backward / k0 scatter, instead of after them.  (Ray-sharded runs
                    # exchange gradients first: their sweeps stay behind the barrier in _optimise.)
                    tv_on, tv_dense = self._tv_now()
                    self._density_swept = self._sweep_grid("density", 1, n_global, self.opt_step + 1, tv_on, tv_dense,
                                                           False, 0, self.X) or self.lr["density"] <= 0
                self._dens_done.record()
            return True

        def after_rgb():
            ext.composite(ws.rgb, ws.s_weight, ws.s_ray, ws.s_slot, ws.ray_off, ws.counters, ws.rgb_acc, None)
            ext.ray_finish(ws.rgb_acc, ws.alphainv_last, target, bg, n, n_global, w_main, w_ent, ws.G, ws.g_last,
                           ws.loss_acc)
            # tensor-core rgbnet: the dZ3 tiles are the only consumer of dL/d(rgb): the fp32 stream is not written
            ext.sample_grad(ws.rgb, ws.s_weight, ws.s_ray, ws.G, target, ws.counters, n_global, w_per,
                            None if tc else ws.d_rgb, ws.d_w, ws.loss_acc, ws.dzt if tc else None,
                            self._tc.grad_scale(n_global) if tc else 1.0)

        if model.rgbnet is None:
            ext.rgb_direct(ws.feat, ws.counters, ws.rgb)
            after_rgb()
            ext.rgb_direct_bwd(ws.rgb, ws.d_rgb, ws.counters, ws.d_feat)
        elif tc:
            self._rgb_tc(ws, pe, True)
            self._mark("mlp_fwd")
            after_rgb()
            self._mark("loss")
            dens_async = density_backward_async()
            self._tc.backward_tiles(ws.xt, ws.dzt, self.C, pe[0].shape[1], ws.counters, ws.cap, ws.d_feat, n_global)
        else:
            m4 = int(ws.counters[0].item())  # parity mode: one host read of the survivor count
            for p in model.rgbnet.parameters():
                p.grad = None
            if m4:
                rgb, feat = self._rgb_torch(ws, viewdirs, m4, True)
                ws.rgb[:m4].copy_(rgb.detach())
            after_rgb()
            if m4:
                rgb.backward(ws.d_rgb[:m4])
                ws.d_feat[:m4].copy_(feat.grad)

        self._mark("mlp_bwd")
        self._join_side()     # the side-stream re-zeroing of the gradient slabs must have finished before the scatter
        if not dens_async:
            ext.march_bwd(self.scene, rays_o, rays_d, ws.t_min, ws.n_steps, ws.ray_off, ws.slot_alpha, ws.slot_T,
                          ws.slot_expd, ws.slot_code, ws.d_feat, ws.d_w, ws.alphainv_last, ws.g_last, self.g_density,
                          None if self.split_k0 else self.g_k0)
        if self.split_k0:
            ext.k0_scatter(self.scene, rays_o, rays_d, ws.t_min, ws.ray_off, ws.s_ray, ws.s_slot, ws.counters,
                           ws.d_feat, self.g_k0, ws.s_pos)
        if dens_async:
            torch.cuda.current_stream().wait_event(self._dens_done)
        self._mark("march_bwd")
        self._optimise(n_global)
        self._mark("sweep")
        return ws.loss_acc[0].clone()

    # -- gradient exchange (ray-sharded data parallel) -----------------------------------------------
    def _slab(self):
        """This rank's x-slab when the sweep is sharded: needs X divisible by the world size."""
        if self.world_size > 1 and self.shard_sweep and self.X % self.world_size == 0:
            n = self.X // self.world_size
            return self.rank * n, (self.rank + 1) * n
        return None

    def _rgbnet_grads(self):
        """Autograd-mode rgbnet gradients, materialised as zeros where a rank had no surviving sample: every rank
        must enter the same collectives."""
        out = []
        for p in self.model.rgbnet.parameters():
            if p.grad is None:
                p.grad = torch.zeros_like(p)
            out.append(p.grad)
        return out

    def _reduce_grads(self, slab):
        import torch.distributed as dist
        from .parallel import allreduce_sum_
        small = []
        if self.model.rgbnet is not None:
            small = [self._tc.grad_flat] if self.mlp_mode == "tc" else self._rgbnet_grads()
        if slab is None:
            allreduce_sum_([self.g_density, self.g_k0] + small, self.dist_group)
            return
        x0, x1 = slab
        for g in (self.g_density, self.g_k0):
            # in-place reduce-scatter: rank r receives the sum of slab r into its own slab of the buffer
            dist.reduce_scatter_tensor(g[x0:x1], g, op=dist.ReduceOp.SUM, group=self.dist_group)
        allreduce_sum_(small, self.dist_group)

    def _gather_params(self, slab, names):
        import torch.distributed as dist
        x0, x1 = slab
        for name in names:
            p = getattr(self, name)
            dist.all_gather_into_tensor(p, p[x0:x1], group=self.dist_group)   # in place
        # the other ranks' slabs of the gradient accumulators still hold this rank's partial sums
        for g in (self.g_density, self.g_k0):
            if x0 > 0:
                ext.zero_(g[:x0])
            if x1 < self.X:
                ext.zero_(g[x1:])

    def _tv_now(self):
        from .trainer import tv_schedule
        return tv_schedule(self.cfg, self.global_step)

    def _peer_barrier(self, with_small_grads=False):
        """Cross-rank ordering point on the current stream: a (tiny) NCCL all-reduce completes on a rank only after
        every rank has reached it in ITS stream order.  Before the sweep it also sums the rgbnet gradients."""
        import torch.distributed as dist
        from .parallel import allreduce_sum_
        small = []
        if with_small_grads and self.model.rgbnet is not None:
            small = [self._tc.grad_flat] if self.mlp_mode == "tc" else self._rgbnet_grads()
        if small:
            allreduce_sum_(small, self.dist_group)
        else:
            dist.all_reduce(self._bar, group=self.dist_group)

    def _sweep_grid(self, name, C, n_global, opt_step, tv_on, tv_dense, peer, x0, x1):
        """TV + (masked) Adam + gradient re-zero of one grid's x-slab [x0, x1) in one kernel; swaps the ping-pong
        buffers when TV is on.  Returns False when the grid is frozen (lr <= 0)."""
        cfg = self.cfg
        lr = self.lr[name]
        if lr <= 0:
            return False
        b1, b2 = self.betas
        wt = float(cfg.get("weight_tv_" + name, 0.0))
        tv = tv_on and wt > 0
        if self.ndc:  # lib/dmpigo.py:147-157 anisotropic weights
            wx = wy = wt / n_global * float(max(self.X, self.Y)) / 128
            wz = wt / n_global * float(self.Z) / 128
        else:
            wx = wy = wz = wt / n_global * float(max(self.X, self.Y, self.Z)) / 128  # lib/dvgo.py:297-305, run.py:392
        cur = getattr(self, name)
        nxt = getattr(self, name + "_next") if tv else cur
        per_lr = self.per_lr if (name == "density" and self.per_lr is not None) else None
        masked = self.masked[name] and per_lr is None  # dispatch of lib/masked_adam.py:60-71
        if peer:
            out = name + "_next" if tv else name
            ext.sweep_peer(cur, self._pp[out], self._pp["g_" + name], self._mc[out], self._mc["g_" + name], self.rank,
                           getattr(self, "g_" + name), getattr(self, "m_" + name), getattr(self, "v_" + name),
                           per_lr, self.X, self.Y, self.Z, C, tv, tv_dense, wx, wy, wz, masked, opt_step,
                           b1, b2, lr, self.eps, x0, x1)
        else:
            ext.sweep(cur, nxt, getattr(self, "g_" + name), getattr(self, "m_" + name), getattr(self, "v_" + name),
                      per_lr, self.X, self.Y, self.Z, C, tv, tv_dense, wx, wy, wz, masked, opt_step,
                      b1, b2, lr, self.eps, x0, x1)
        if tv:
            setattr(self, name, nxt)
            setattr(self, name + "_next", cur)
            if peer:
                for tab in (self._pp, self._mc):
                    tab[name], tab[name + "_next"] = tab[name + "_next"], tab[name]
        return True

    def _optimise(self, n_global):
        cfg = self.cfg
        peer = self.exchange == "peer"
        slab = None if peer else self._slab()
        flags = peer and self._flags is not None
        if flags:
            self._epoch += 1
            ext.peer_barrier(self._flag_ptrs, self.rank, self._epoch)     # every rank's backward has finished
            self._mark("grad_exchange")
            x0, x1 = (self.X * self.rank) // self.world_size, (self.X * (self.rank + 1)) // self.world_size
        elif peer:
            self._peer_barrier(with_small_grads=True)      # every rank's backward has finished
            self._mark("grad_exchange")
            x0, x1 = (self.X * self.rank) // self.world_size, (self.X * (self.rank + 1)) // self.world_size
        else:
            if self.world_size > 1:
                self._reduce_grads(slab)
                self._mark("grad_exchange")
            x0, x1 = slab if slab is not None else (0, self.X)
        updated = []
        self.opt_step += 1
        b1, b2 = self.betas
        tv_on, tv_dense = self._tv_now()
        for name, C in (("density", 1), ("k0", self.C)):
            if name == "density" and self._density_swept:
                self._density_swept = False      # done on the side stream right behind march_bwd (step())
                updated.append(name)
                continue
            if self._sweep_grid(name, C, n_global, self.opt_step, tv_on, tv_dense, peer, x0, x1):
                updated.append(name)
        rgbnet_done = False
        if flags and self.model.rgbnet is not None and self.lr["rgbnet"] > 0:
            # rgbnet Adam on the gradient summed over ranks, read from the peers' flat buffers inside the kernel
            ext.adam_upd_peer(self._tc.params, self._gflat_ptrs, self._gflat_mc, self._tc.exp_avg, self._tc.exp_avg_sq,
                              self.opt_step, b1, b2, self.lr["rgbnet"], self.eps)
            rgbnet_done = True
        if peer:
            self._mark("sweep_grids")
            if flags:
                self._epoch += 1
                ext.peer_barrier(self._flag_ptrs, self.rank, self._epoch)   # every rank has read my gradients and written
            else:                                                           # my parameters
                self._peer_barrier()
            # The slabs other ranks own still hold my partial sums: (n-1)/n of both gradient accumulators to re-zero
            # (187 MB at 8 ranks).  Nothing reads them before the next backward's scatter, so the zeroing runs on a side
            # stream under the next forward pass instead of on the critical path.
            if self._side is None:
                self._side = torch.cuda.Stream(device=self.device)
                self._zero_done = torch.cuda.Event()
            ev = torch.cuda.Event()
            ev.record()
            self._side.wait_event(ev)
            with torch.cuda.stream(self._side):
                for g in (self.g_density, self.g_k0):
                    if x0 > 0:
                        ext.zero_(g[:x0])
                    if x1 < self.X:
                        ext.zero_(g[x1:])
                self._zero_done.record()
            self._zero_pending = True
            self._mark("param_gather")
        elif slab is not None:
            self._mark("sweep_grids")     # slab-sharded grid sweeps end here; "sweep" then only holds the rgbnet Adam
            self._gather_params(slab, updated)
            self._mark("param_gather")
        if self.model.rgbnet is not None and self.lr["rgbnet"] > 0 and not rgbnet_done:
            if self.mlp_mode == "tc":
                self._tc.adam_step(self.opt_step, b1, b2, self.lr["rgbnet"], self.eps)
            else:
                for p in self.model.rgbnet.parameters():
                    if p.grad is None:
                        continue
                    st = self.rgbnet_state.setdefault(p, {"exp_avg": torch.zeros_like(p), "exp_avg_sq": torch.zeros_like(p)})
                    adam_upd_cuda.adam_upd(p.data, p.grad, st["exp_avg"], st["exp_avg_sq"], self.opt_step, b1, b2,
                                           self.lr["rgbnet"], self.eps)
        for k in self.lr:  # run.py:401-406
            self.lr[k] *= self.decay

    # -- optimiser state in the reference's MaskedAdam.state_dict() format (checkpoint boundary, run.py:420-437) --
    def _opt_entries(self):
        """(group name, [parameter names]) in the reference's group order (lrate_* keys of cfg_train that exist
        on the model and have lr > 0, lib/utils.py:24-44)."""
        out = []
        for key in self.cfg:
            if not key.startswith("lrate_") or key == "lrate_decay":
                continue
            name = key[len("lrate_"):]
            if name in ("density", "k0") and self.lr.get(name, 0) > 0:
                out.append((name, [name]))
            elif name == "rgbnet" and self.model.rgbnet is not None and self.lr.get(name, 0) > 0:
                out.append((name, ["rgbnet." + n for n, _ in self.model.rgbnet.named_parameters()]))
        return out

    @torch.no_grad()
    def optimizer_state_dict(self):
        self.check_status()
        state, groups, idx = {}, [], 0
        for name, pnames in self._opt_entries():
            ids = []
            if name == "density":
                tensors = [(self.m_density.reshape(1, 1, self.X, self.Y, self.Z).clone(),
                            self.v_density.reshape(1, 1, self.X, self.Y, self.Z).clone())]
            elif name == "k0":
                tensors = [(ext.cl_to_ncdhw(self.m_k0), ext.cl_to_ncdhw(self.v_k0))]
            elif self.mlp_mode == "tc":
                tensors = [(m.clone(), v.clone()) for m, v in zip(self._tc.unflatten(self._tc.exp_avg),
                                                                  self._tc.unflatten(self._tc.exp_avg_sq))]
            else:
                tensors = []
                for p in self.model.rgbnet.parameters():
                    st = self.rgbnet_state.get(p)
                    tensors.append((st["exp_avg"].clone(), st["exp_avg_sq"].clone()) if st else
                                   (torch.zeros_like(p), torch.zeros_like(p)))
            for m, v in tensors:
                if self.opt_step > 0:
                    state[idx] = {"step": self.opt_step, "exp_avg": m, "exp_avg_sq": v}
                ids.append(idx)
                idx += 1
            groups.append({"lr": self.lr[name], "skip_zero_grad": bool(self.masked.get(name, False)),
                           "betas": tuple(self.betas), "eps": self.eps, "params": ids})
        return {"state": state, "param_groups": groups}

    @torch.no_grad()
    def load_optimizer_state_dict(self, sd):
        """Accepts the `optimizer_state_dict` of a reference checkpoint (same model, same cfg_train)."""
        entries = self._opt_entries()
        assert len(entries) == len(sd["param_groups"]), "optimizer groups do not match cfg_train"
        steps = set()
        for (name, _), grp in zip(entries, sd["param_groups"]):
            self.lr[name] = float(grp["lr"])
            sts = [sd["state"].get(i) for i in grp["params"]]
            if any(s is None for s in sts):
                continue
            steps.update(int(s["step"]) for s in sts)
            dev = self.device
            if name == "density":
                self.m_density.copy_(sts[0]["exp_avg"].to(dev).reshape(self.X, self.Y, self.Z))
                self.v_density.copy_(sts[0]["exp_avg_sq"].to(dev).reshape(self.X, self.Y, self.Z))
            elif name == "k0":
                self.m_k0.copy_(ext.ncdhw_to_cl(sts[0]["exp_avg"].to(dev).float().contiguous()))
                self.v_k0.copy_(ext.ncdhw_to_cl(sts[0]["exp_avg_sq"].to(dev).float().contiguous()))
            elif self.mlp_mode == "tc":
                for key, flat in (("exp_avg", self._tc.exp_avg), ("exp_avg_sq", self._tc.exp_avg_sq)):
                    for dst, s in zip(self._tc.unflatten(flat), sts):
                        dst.copy_(s[key].to(dev))
            else:
                for p, s in zip(self.model.rgbnet.parameters(), sts):
                    self.rgbnet_state[p] = {"exp_avg": s["exp_avg"].to(dev).clone(),
                                            "exp_avg_sq": s["exp_avg_sq"].to(dev).clone()}
        assert len(steps) <= 1, "parameters with different step counts are not supported"
        self.opt_step = steps.pop() if steps else 0

    @torch.no_grad()
    def sync_to_model(self):
        if self.model.rgbnet is not None and self.mlp_mode == "tc":
            self._tc.sync_to_module(self.model.rgbnet)
        return super().sync_to_model()
