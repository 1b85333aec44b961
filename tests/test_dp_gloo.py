"""World-size-2 gloo test of the data-parallel host logic (nero_b200/dp.py) on CPU: two ranks run the ORACLE on their
ray shards, rescale the eikonal mean with dp.global_mean_weight, average flat gradients with dp.sync_gradients, and
must reproduce the single-process gradient of the whole batch."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import nero_oracle as O
from helpers import build_params
from nero_b200 import dp

CFG = {'n_samples': 16, 'n_importance': 16, 'n_bg_samples': 8, 'up_sample_steps': 2}
STEP = 10000


def _loss_and_grads(sd, rays, sl, eik_scale_fn):
    p = {k: v.clone().requires_grad_(torch.is_floating_point(v) and not k.endswith('FG_LUT')) for k, v in sd.items()}
    c = O.merged_cfg(CFG)
    lut = sd['color_network.FG_LUT'][0]
    r = {k: v[sl] for k, v in rays.items()}
    out = O.render(p, CFG, lut, r['rays_o'], r['rays_d'], r['near'], r['far'], r['human_poses'], O.get_anneal_val(c, STEP), STEP)
    n_in = out['gradient_error'].numel()
    loss = torch.mean(O.compute_rgb_loss(out['ray_rgb'], r['rgb'])) + eik_scale_fn(n_in) * torch.mean(out['gradient_error'] * 0.1)
    loss.backward()
    names = sorted(k for k in p if p[k].grad is not None)
    return torch.cat([p[k].grad.reshape(-1) for k in names]), n_in


def _worker(rank, world, port, q):
    os.environ['MASTER_ADDR'], os.environ['MASTER_PORT'] = '127.0.0.1', str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    torch.set_num_threads(2)
    sd = build_params(CFG)
    rays = O.synthetic_rays(16, seed=3)
    # rank 1 passes the count as the engine keeps it in graph mode: a device-side int32 tensor (never read back to the host)
    flat, n_in = _loss_and_grads(sd, rays, dp.shard_slice(16, rank, world),
                                 lambda n: dp.global_mean_weight(torch.tensor([n], dtype=torch.int32) if rank else n, world))
    dp.sync_gradients(flat, world)
    if rank == 0:
        q.put(flat.numpy())
    dist.destroy_process_group()


def test_ray_sharded_dp_matches_single_process():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p_ in procs:
        p_.start()
    got = q.get(timeout=300)
    for p_ in procs:
        p_.join(timeout=60)
        assert p_.exitcode == 0
    torch.set_num_threads(4)
    sd = build_params(CFG)
    rays = O.synthetic_rays(16, seed=3)
    want, _ = _loss_and_grads(sd, rays, slice(0, 16), lambda n: 1.0)
    want = want.numpy()
    assert np.abs(got - want).max() <= 1e-5 * np.abs(want).max() + 1e-8


def test_shard_slices_cover_batch():
    idx = np.concatenate([np.arange(64)[dp.shard_slice(64, r, 4)] for r in range(4)])
    assert (idx == np.arange(64)).all()


# ------------------------------------------------------------------------------------------------ stage II (surface points)
MCFG = {'shader_cfg': {'human_lights': True, 'diffuse_sample_num': 8, 'specular_sample_num': 8}}
MSTEP = 5000      # >= 2000: every stage-II loss term is a plain mean over points (field.py:1079: the min/max regulariser of
                  # earlier steps is a SUM over the batch and would need a factor `world` per rank)


def _material_grads(sd, batch, rands, sl, verts, tris):
    import nero_oracle_mat as OM
    p = {k: v.clone().requires_grad_(torch.is_floating_point(v) and not k.endswith('light_pts')) for k, v in sd.items()}
    tabs = (OM.direction_samples(8), OM.direction_samples(8))
    b = {k: v[sl] for k, v in batch.items()}
    r = {k: v[sl] for k, v in rands.items()}
    out = OM.material_train_outputs(p, MCFG, tabs, lambda o, d: OM.renderer_trace(verts, tris, o, d), b, MSTEP, r)
    OM.material_training_loss(out).backward()
    names = sorted(k for k in p if p[k].grad is not None)
    return torch.cat([p[k].grad.reshape(-1) for k in names])


def _material_setup():
    import nero_oracle_mat as OM
    from helpers import build_material_params
    verts, tris = OM.test_scene(1)
    sd = build_material_params(MCFG['shader_cfg'])
    batch = OM.synthetic_surface_batch(verts, tris, 8, seed=5)
    return sd, batch, OM.draw_rands(8), verts, tris


def _material_worker(rank, world, port, q):
    os.environ['MASTER_ADDR'], os.environ['MASTER_PORT'] = '127.0.0.1', str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    torch.set_num_threads(2)
    sd, batch, rands, verts, tris = _material_setup()
    flat = _material_grads(sd, batch, rands, dp.shard_slice(8, rank, world), verts, tris)
    dp.sync_gradients(flat, world)
    if rank == 0:
        q.put(flat.numpy())
    dist.destroy_process_group()


def test_point_sharded_material_dp_matches_single_process():
    """Stage II shards over surface points: equal shards + mean losses => averaged shard gradients == full-batch gradient."""
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_material_worker, args=(r, 2, port, q)) for r in range(2)]
    for p_ in procs:
        p_.start()
    got = q.get(timeout=300)
    for p_ in procs:
        p_.join(timeout=60)
    sd, batch, rands, verts, tris = _material_setup()
    want = _material_grads(sd, batch, rands, slice(0, 8), verts, tris).numpy()
    assert np.abs(got - want).max() <= 1e-5 * np.abs(want).max() + 1e-9
