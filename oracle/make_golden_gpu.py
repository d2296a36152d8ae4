"""TEST INFRASTRUCTURE ONLY -- record outputs of the reference's OWN CUDA kernels (oracle/_ref, built
from the unmodified /root/reference/lib/cuda sources) on a B200, as golden vectors that pin the C
oracle and the product.  Run on the GPU box:

    python -m oracle.make_golden_gpu gpurun_out/golden/ref_gpu_ops.npz
    python -m oracle.make_golden_gpu gpurun_out/golden/ref_gpu_ops_f64.npz f64

then copy the files to tests/golden/.  The second form records the reference's DOUBLE instantiation
(AT_DISPATCH_FLOATING_TYPES with float64 tensors) on the same seeded inputs, widened to float64 and scaled
by (1 + 2^-30) so that they are not float32-representable; it pins oracle/dvgo_oracle_f64.c.
Inputs are seeded and stored in the file, so the fixture is self-contained.  Sizes are small (a few hundred rays) to keep the fixture < 1 MB.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "_ref"))
sys.path.insert(0, os.path.join(HERE, ".."))


def main(out_path, f64=False):
    import ref_adam_upd_cuda as ad
    import ref_render_utils_cuda as ru
    import ref_total_variation_cuda as tv
    from tests.util import make_rays, sorted_ray_ids

    dev = "cuda"
    c = lambda t: t.detach().cpu().numpy()
    save = {}
    if f64:
        # every floating input of the run below goes through Tensor.to(dev): widen it there
        _to = torch.Tensor.to

        def widen(t, *a, **k):
            t = _to(t, *a, **k)
            return _to(t, torch.float64) * (1.0 + 2.0 ** -30) if t.dtype == torch.float32 else t
        torch.Tensor.to = widen

    lo = torch.tensor([-1.0, -0.9, -0.8]).to(dev)
    hi = torch.tensor([1.0, 0.9, 0.8]).to(dev)
    ro, rd, vd, _ = make_rays(192, 42)
    ro, rd = ro.to(dev), rd.to(dev)
    near, far, stepdist = 0.2, 6.0, 0.037
    pts, mask, ray_id, step_id, N_steps, t_min, t_max = ru.sample_pts_on_rays(ro, rd, lo, hi, near, far, stepdist)
    save.update(rays_o=c(ro), rays_d=c(rd), xyz_min=c(lo), xyz_max=c(hi), near=np.float32(near),
                far=np.float32(far), stepdist=np.float32(stepdist), rays_pts=c(pts), mask_outbbox=c(mask),
                ray_id=c(ray_id), step_id=c(step_id), N_steps=c(N_steps), t_min=c(t_min), t_max=c(t_max))
    start, dirs = ru.infer_ray_start_dir(ro, rd, t_min)
    save.update(rays_start=c(start), rays_dir=c(dirs))

    g = torch.Generator(device="cpu").manual_seed(7)
    world = (torch.rand(23, 19, 17, generator=g) > 0.4).to(dev)
    shape = torch.tensor([23.0, 19.0, 17.0]).to(dev)
    scale = (shape - 1) / (hi - lo)
    shift = -lo * scale
    occ = ru.maskcache_lookup(world, pts, scale, shift)
    save.update(world=c(world), scale=c(scale), shift=c(shift), maskcache=c(occ))

    ndc_o = torch.cat([(torch.rand(40, 2, generator=g) - 0.5) * 2.4, -torch.ones(40, 1)], -1).to(dev).contiguous()
    ndc_d = torch.cat([(torch.rand(40, 2, generator=g) - 0.5) * 0.9, 2 * torch.ones(40, 1)], -1).to(dev).contiguous()
    ndc_pts, ndc_mask = ru.sample_ndc_pts_on_rays(ndc_o, ndc_d, lo, hi, 33)
    save.update(ndc_o=c(ndc_o), ndc_d=c(ndc_d), ndc_n=np.int32(33), ndc_pts=c(ndc_pts), ndc_mask=c(ndc_mask))

    density = torch.cat([torch.randn(4000, generator=g) * 4, torch.tensor([100.0, -100.0, 0.0, 88.0, 30.0])]).to(dev)
    shift_a, interval = -4.595, 0.5
    exp_d, alpha = ru.raw2alpha(density, shift_a, interval)
    grad_back = torch.randn(density.shape, generator=g).to(dev)
    r2a_grad = ru.raw2alpha_backward(exp_d, grad_back, interval)
    save.update(density=c(density), shift_a=np.float32(shift_a), interval=np.float32(interval), exp_d=c(exp_d),
                alpha=c(alpha), grad_back=c(grad_back), raw2alpha_grad=c(r2a_grad))

    n_rays, n_pts = 120, 9000
    rid = sorted_ray_ids(n_rays, n_pts, 11).to(dev)
    a2w_alpha = (torch.rand(n_pts, generator=g) ** 3 * 0.6).to(dev)
    w, T, last, i_s, i_e = ru.alpha2weight(a2w_alpha, rid, n_rays)
    gw = torch.randn(n_pts, generator=g).to(dev)
    gl = torch.randn(n_rays, generator=g).to(dev)
    a2w_grad = ru.alpha2weight_backward(a2w_alpha, w, T, last, i_s, i_e, n_rays, gw, gl)
    save.update(a2w_alpha=c(a2w_alpha), a2w_ray_id=c(rid), a2w_n_rays=np.int32(n_rays), weight=c(w), T=c(T),
                alphainv_last=c(last), i_start=c(i_s), i_end=c(i_e), a2w_gw=c(gw), a2w_gl=c(gl), a2w_grad=c(a2w_grad))

    tv_param = (torch.randn(1, 3, 9, 8, 7, generator=g) * 1.5).to(dev)
    tv_grad_in = torch.randn(1, 3, 9, 8, 7, generator=g)
    tv_grad_in[torch.rand(tv_grad_in.shape, generator=g) < 0.5] = 0
    tv_grad_in = tv_grad_in.to(dev)
    wy, wz = 0.7, 1.3
    for dense in (0, 1):
        gcopy = tv_grad_in.clone()
        tv.total_variation_add_grad(tv_param, gcopy, 0.3, wy, wz, bool(dense))
        save["tv_out_dense%d" % dense] = c(gcopy)
    save.update(tv_param=c(tv_param), tv_grad_in=c(tv_grad_in), tv_wy=np.float32(wy), tv_wz=np.float32(wz))

    N = 5003
    p0, m0, v0 = torch.randn(N, generator=g), torch.randn(N, generator=g) * 0.01, torch.rand(N, generator=g) * 1e-3
    gr = torch.randn(N, generator=g)
    gr[torch.rand(N, generator=g) < 0.4] = 0
    perlr = torch.rand(N, generator=g)
    p0, m0, v0, gr, perlr = (t.to(dev) for t in (p0, m0, v0, gr, perlr))   # (widened here in the f64 run)
    save.update(adam_p=c(p0), adam_m=c(m0), adam_v=c(v0), adam_g=c(gr), adam_perlr=c(perlr))
    for name in ("adam_upd", "masked_adam_upd", "adam_upd_with_perlr"):
        p, m, v = p0.clone(), m0.clone(), v0.clone()
        for step in (1, 2, 3):
            args = (p, gr, m, v) + ((perlr,) if name == "adam_upd_with_perlr" else ())
            getattr(ad, name)(*args, step, 0.9, 0.99, 0.1, 1e-8)
        save["adam_out_p_" + name], save["adam_out_m_" + name], save["adam_out_v_" + name] = c(p), c(m), c(v)

    if f64:
        torch.Tensor.to = _to
        assert save["rays_pts"].dtype == np.float64 and save["adam_out_p_adam_upd"].dtype == np.float64
    torch.cuda.synchronize()
    os.makedirs(os.path.dirname(os.path.abspath(out_path)), exist_ok=True)
    np.savez_compressed(out_path, **save)
    print("wrote", out_path, "M0 =", len(save["ray_id"]))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/golden/ref_gpu_ops.npz", f64="f64" in sys.argv[2:])
