"""Synthetic Blender-geometry inputs (there is no dataset in the build container).

Cameras follow the reference's NeRF-synthetic loader: `pose_spherical(theta, phi, 4.0)`
(lib/load_blender.py:25-42), H = W = 800, focal = 0.5*W / tan(0.5 * 0.6911112) = 1111.11,
near = 2, far = 6 (lib/load_data.py:55), white background; rays are pixel-centre rays of
`get_rays_of_a_view` (lib/ray_utils.py:9-47, 80-85; mode='center', OpenGL convention).
Images are procedural (targets only enter the loss).  Everything is generated with torch on the
device it is asked for, from an explicit seed.
"""
import math

import numpy as np
import torch

BLENDER = dict(H=800, W=800, camera_angle_x=0.6911112070083618, radius=4.0, near=2.0, far=6.0, bg=1.0)


def blender_focal(W=BLENDER["W"], camera_angle_x=BLENDER["camera_angle_x"]):
    return 0.5 * W / math.tan(0.5 * camera_angle_x)


def pose_spherical(theta_deg, phi_deg, radius):
    """c2w of lib/load_blender.py:37-42 (translate along z, rotate by phi about x, theta about y,
    then swap to the Blender world frame)."""
    th, ph = math.radians(theta_deg), math.radians(phi_deg)
    trans = np.eye(4, dtype=np.float64); trans[2, 3] = radius
    rot_phi = np.array([[1, 0, 0, 0], [0, math.cos(ph), -math.sin(ph), 0],
                        [0, math.sin(ph), math.cos(ph), 0], [0, 0, 0, 1]], dtype=np.float64)
    rot_th = np.array([[math.cos(th), 0, -math.sin(th), 0], [0, 1, 0, 0],
                       [math.sin(th), 0, math.cos(th), 0], [0, 0, 0, 1]], dtype=np.float64)
    swap = np.array([[-1, 0, 0, 0], [0, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, 1]], dtype=np.float64)
    return (swap @ rot_th @ rot_phi @ trans).astype(np.float32)


def random_poses(n, seed=777, radius=BLENDER["radius"]):
    rng = np.random.RandomState(seed)
    thetas = rng.uniform(-180, 180, size=n)
    phis = rng.uniform(-90, 0, size=n)
    return np.stack([pose_spherical(t, p, radius) for t, p in zip(thetas, phis)])


def intrinsics(H=BLENDER["H"], W=BLENDER["W"], focal=None):
    f = blender_focal(W) if focal is None else focal
    return np.array([[f, 0, 0.5 * W], [0, f, 0.5 * H], [0, 0, 1]], dtype=np.float32)


def rays_of_view(H, W, K, c2w, device="cpu"):
    """rays_o, rays_d, viewdirs, each [H,W,3] (lib/ray_utils.py:9-47 mode='center', :80-85)."""
    c2w = torch.as_tensor(c2w, dtype=torch.float32, device=device)
    i, j = torch.meshgrid(torch.linspace(0, W - 1, W, device=device),
                          torch.linspace(0, H - 1, H, device=device), indexing="ij")
    i, j = i.t() + 0.5, j.t() + 0.5
    dirs = torch.stack([(i - K[0][2]) / K[0][0], -(j - K[1][2]) / K[1][1], -torch.ones_like(i)], -1)
    rays_d = torch.sum(dirs[..., None, :] * c2w[:3, :3], -1)
    rays_o = c2w[:3, 3].expand(rays_d.shape)
    viewdirs = rays_d / rays_d.norm(dim=-1, keepdim=True)
    return rays_o.contiguous(), rays_d.contiguous(), viewdirs.contiguous()


def random_training_rays(n_rays, n_views=100, seed=777, H=BLENDER["H"], W=BLENDER["W"], device="cpu"):
    """`n_rays` incoherent training rays: random (view, pixel) picks from `n_views` random cameras --
    the access pattern of the reference's random-permutation batches (lib/ray_utils.py:283-290)."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    poses = torch.as_tensor(random_poses(n_views, seed), device=device)
    K = intrinsics(H, W)
    v = torch.randint(n_views, (n_rays,), generator=g).to(device)
    px = torch.randint(W, (n_rays,), generator=g).to(device).float() + 0.5
    py = torch.randint(H, (n_rays,), generator=g).to(device).float() + 0.5
    dirs = torch.stack([(px - K[0][2]) / K[0][0], -(py - K[1][2]) / K[1][1], -torch.ones_like(px)], -1)
    R = poses[v, :3, :3]
    rays_d = torch.sum(dirs[:, None, :] * R, -1)
    rays_o = poses[v, :3, 3]
    viewdirs = rays_d / rays_d.norm(dim=-1, keepdim=True)
    target = procedural_rgb(px / W, py / H, v.float() / max(n_views, 1))
    return rays_o.contiguous(), rays_d.contiguous(), viewdirs.contiguous(), target.contiguous()


def procedural_rgb(u, v, w):
    """Any smooth RGB in [0,1]; targets only enter the loss."""
    two_pi = 2 * math.pi
    return torch.stack([0.5 + 0.5 * torch.sin(two_pi * (u + w)),
                        0.5 + 0.5 * torch.sin(two_pi * (v + 2 * w) + 1.0),
                        0.5 + 0.5 * torch.sin(two_pi * (u + v) + 2.0)], -1)


# The BASELINE.json configurations made concrete.
def fine_bbox():
    """Cube +-1.5*1.05 (world_bound_scale 1.05, configs/default.py:104) -> exactly 160^3 voxels."""
    e = 1.5 * 1.05
    return np.array([-e, -e, -e], np.float32), np.array([e, e, e], np.float32)


def coarse_bbox():
    """lego's coarse bbox: +-(3.3, 3.3, 2.7)-ish half extents -> 107 x 107 x 88 (lib/dvgo.py:499)."""
    return np.array([-3.3, -3.3, -2.7], np.float32), np.array([3.3, 3.3, 2.7], np.float32)


FINE_MODEL = dict(num_voxels=160 ** 3, num_voxels_base=160 ** 3, alpha_init=1e-2, fast_color_thres=1e-4,
                  rgbnet_dim=12, rgbnet_direct=True, rgbnet_depth=3, rgbnet_width=128, viewbase_pe=4)
COARSE_MODEL = dict(num_voxels=1024000, num_voxels_base=1024000, alpha_init=1e-6, fast_color_thres=1e-7,
                    rgbnet_dim=0)
FINE_TRAIN = dict(N_rand=8192, lrate_density=1e-1, lrate_k0=1e-1, lrate_rgbnet=1e-3, lrate_decay=20,
                  weight_main=1.0, weight_entropy_last=1e-3, weight_rgbper=1e-2,
                  weight_tv_density=1e-5, weight_tv_k0=1e-5, tv_dense=True,
                  skip_zero_grad_fields=["density", "k0"])
RENDER_KWARGS = dict(near=BLENDER["near"], far=BLENDER["far"], bg=BLENDER["bg"], stepsize=0.5,
                     inverse_y=False, flip_x=False, flip_y=False)


def randomize_grids_(model, seed=777):
    """density ~ N(0,1), k0 ~ N(0,1) (random-init grids of BASELINE.json); rgbnet keeps its default init."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    with torch.no_grad():
        model.density.copy_(torch.randn(model.density.shape, generator=g))
        model.k0.copy_(torch.randn(model.k0.shape, generator=g))
    return model
