"""Synthetic workload generators for the benchmarks and profiling tools (no dataset exists offline): camera rays on the
radius-3 sphere looking at the object, look-at poses, seeded parameter perturbation, an icosphere test scene.

These are INPUT generators only -- no part of the rendering algorithm.  They reproduce, value for value, the generators the
test oracle uses for its fixtures (`tests/test_synthetic_inputs.py` asserts equality), so that the benchmarks run the very
workloads the parity tests pin, without the measured arm importing anything from `oracle/`.
"""
import numpy as np
import torch
import torch.nn.functional as F


def near_far_from_sphere(rays_o, rays_d):
    a = torch.sum(rays_d ** 2, dim=-1, keepdim=True)
    b = 2.0 * torch.sum(rays_o * rays_d, dim=-1, keepdim=True)
    mid = 0.5 * (-b) / a
    return torch.clamp(mid - 1.0, min=1e-3), mid + 1.0


def human_coordinate_poses(poses, fixed_camera=False):
    """Capturer frame of every camera (same construction as NeROShapeRenderer.get_human_coordinate_poses)."""
    pn = poses.shape[0]
    cam_cen = (-poses[:, :, :3].permute(0, 2, 1) @ poses[:, :, 3:])[..., 0]
    if not fixed_camera:
        cam_cen = torch.cat([cam_cen[:, :2], torch.zeros_like(cam_cen[:, 2:])], -1)
    Y = torch.zeros(pn, 3, dtype=poses.dtype)
    Y[:, 2] = -1.0
    Z = torch.cat([poses[:, 2, :2], torch.zeros_like(poses[:, 2, 2:3])], -1)
    Z = F.normalize(Z, dim=-1)
    X = torch.cross(Y, Z, dim=-1)
    R = torch.stack([X, Y, Z], 1)
    t = -R @ cam_cen[:, :, None]
    return torch.cat([R, t], -1)


def synthetic_rays(R, seed=6033, dtype=torch.float32):
    """o = 3*normalize(N(0,I)), d = normalize(-o + 0.2*N(0,I)), target rgb ~ U[0,1], look-at poses."""
    g = torch.Generator().manual_seed(seed)
    o = 3.0 * F.normalize(torch.randn(R, 3, generator=g), dim=-1)
    d = F.normalize(-o + 0.2 * torch.randn(R, 3, generator=g), dim=-1)
    rgb = torch.rand(R, 3, generator=g)
    zc = F.normalize(-o, dim=-1)
    up = torch.tensor([0.0, 0.0, 1.0]).expand(R, 3)
    xc = F.normalize(torch.cross(up, zc, dim=-1) + 1e-3, dim=-1)
    yc = torch.cross(zc, xc, dim=-1)
    Rm = torch.stack([xc, yc, zc], 1)
    poses = torch.cat([Rm, -(Rm @ o[:, :, None])], -1)
    near, far = near_far_from_sphere(o, d)
    hp = human_coordinate_poses(poses)
    cv = lambda x: x.to(dtype)
    return {'rays_o': cv(o), 'rays_d': cv(d), 'near': cv(near), 'far': cv(far), 'rgb': cv(rgb), 'human_poses': cv(hp), 'poses': cv(poses)}


def perturb_params(p, seed=7, rel=0.05, sdf_shift=0.45):
    """Deterministic perturbation of a freshly initialised state_dict away from init statistics (zero-initialised PE columns,
    unit weight-norm gains and constant biases would otherwise hide errors); shifts the SDF bias back so the synthetic rays
    still cross a zero level set inside the unit sphere."""
    g = torch.Generator().manual_seed(seed)
    out = {}
    for k in sorted(p.keys()):
        v = p[k]
        if k.endswith('FG_LUT') or k.endswith('light_pts') or not torch.is_floating_point(v) or k.endswith('variance'):
            out[k] = v.clone()
            continue
        noise = torch.randn(v.shape, generator=g).to(v.dtype)
        if k.endswith('weight_g'):
            out[k] = v * (1.0 + rel * noise)
        elif k.endswith('bias'):
            out[k] = v + 0.02 * noise
        else:
            scale = v.std() if v.numel() > 1 and float(v.std()) > 0 else torch.tensor(0.05)
            out[k] = v + rel * scale * noise
    k = 'sdf_network.lin8.bias'
    if k in out:
        out[k] = out[k].clone()
        out[k][0] -= sdf_shift
    return out


def icosphere(subdiv=2, radius=0.5, seed=None, bump=0.0):
    """Subdivided icosahedron (20*4^subdiv faces) pushed to `radius`, optional seeded radial bumps."""
    t = (1.0 + 5.0 ** 0.5) / 2.0
    v = [(-1, t, 0), (1, t, 0), (-1, -t, 0), (1, -t, 0), (0, -1, t), (0, 1, t), (0, -1, -t), (0, 1, -t),
         (t, 0, -1), (t, 0, 1), (-t, 0, -1), (-t, 0, 1)]
    f = [(0, 11, 5), (0, 5, 1), (0, 1, 7), (0, 7, 10), (0, 10, 11), (1, 5, 9), (5, 11, 4), (11, 10, 2), (10, 7, 6),
         (7, 1, 8), (3, 9, 4), (3, 4, 2), (3, 2, 6), (3, 6, 8), (3, 8, 9), (4, 9, 5), (2, 4, 11), (6, 2, 10), (8, 6, 7),
         (9, 8, 1)]
    v = [np.asarray(x, np.float64) / np.linalg.norm(x) for x in v]
    for _ in range(subdiv):
        cache, nf = {}, []

        def mid(a, b):
            key = (min(a, b), max(a, b))
            if key not in cache:
                m = v[a] + v[b]
                v.append(m / np.linalg.norm(m))
                cache[key] = len(v) - 1
            return cache[key]
        for a, b, c in f:
            ab, bc, ca = mid(a, b), mid(b, c), mid(c, a)
            nf += [(a, ab, ca), (b, bc, ab), (c, ca, bc), (ab, bc, ca)]
        f = nf
    v = np.stack(v, 0)
    if bump > 0:
        rng = np.random.RandomState(seed if seed is not None else 0)
        k = rng.randn(6, 3)
        v = v * (1.0 + bump * sum(np.sin(3.0 * v @ k[i] + i) for i in range(6)) / 6.0)[:, None]
    return (v * radius).astype(np.float32), np.asarray(f, np.int32)


def test_scene(subdiv=2, seed=3):
    """A bumpy icosphere (radius 0.5) plus a small satellite sphere; faces wound so that cross(v1-v0, v2-v0) points INTO the
    solid (the convention of NeuS-extracted meshes, flipped back by NeROMaterialRenderer.trace)."""
    v0, f0 = icosphere(subdiv, 0.5, seed=seed, bump=0.15)
    v1, f1 = icosphere(max(subdiv - 1, 0), 0.2)
    v1 = v1 + np.asarray([0.62, 0.1, 0.15], np.float32)
    f = np.concatenate([f0, f1 + v0.shape[0]], 0)
    return np.concatenate([v0, v1], 0), np.ascontiguousarray(f[:, [0, 2, 1]])


test_scene.__test__ = False   # not a pytest test
