"""Size-independent properties of the CUDA path at the bench workload's FULL size (1024 rays x (64+64)+32 samples, the
configuration BASELINE.json quotes) and the edge cases of the domain: rays that never enter the unit sphere (no inner
samples), single / odd ray counts, stage II with no geometry hit and with a single surface point.
The oracle is too slow for these sizes; what is checked here does not need it: invariants of the algorithm, determinism,
batch-split invariance (rays are independent units), linearity of the backward pass in the loss scale.
"""
import numpy as np
import pytest
import torch

import nero_oracle as O
import nero_oracle_mat as OM
from helpers import build_params, build_material_params

pytestmark = pytest.mark.gpu
DEV = 'cuda'
STEP = 30000


def make_shape_net(cfg=None):
    from nero_b200.renderer import NeROShapeRenderer
    cfg = cfg or {}
    net = NeROShapeRenderer(cfg, training=False)
    net.load_state_dict(build_params(cfg))
    return net.cuda()


def render(net, r, step=STEP, sl=slice(None)):
    torch.manual_seed(0)     # the occlusion loss draws a random subset of its candidates (renderer.py:535-541)
    z = net.sample_ray(r['rays_o'][sl], r['rays_d'][sl], r['near'][sl], r['far'][sl], 0)
    out = net.render_core(r['rays_o'][sl], r['rays_d'][sl], z, r['human_poses'][sl], net.get_anneal_val(step), step)
    return z, out


def loss_of(net, out, rgb):
    return torch.mean(net.compute_rgb_loss(out['ray_rgb'], rgb)) + torch.mean(out['gradient_error'] * 0.1) + torch.mean(out['loss_occ'])


def grads_of(net):
    return torch.cat([p.grad.reshape(-1) for p in net.parameters()]).clone()


def test_full_size_invariants_determinism_split_and_linearity():
    net = make_shape_net()
    R = 1024
    r = {k: v.to(DEV) for k, v in O.synthetic_rays(R, seed=6033).items()}
    z, out = render(net, r)
    S = z.shape[1]
    assert S == 160 and out['ray_rgb'].shape == (R, 3)
    e = net.engine
    n_in, n_out = e.state['N_in'], int(e.w['n_out'].item())
    assert n_in + n_out == R * S and out['gradient_error'].shape == (n_in,)
    zi = z[:, :128]
    assert bool((zi[:, 1:] >= zi[:, :-1]).all()), 'inner samples sorted'
    assert bool((zi >= r['near'] - 1e-5).all()) and bool((zi <= r['far'] + 1e-5).all())
    assert bool((z[:, 128:] > r['far']).all()), 'background samples lie beyond far'
    rgb = out['ray_rgb']
    assert bool(torch.isfinite(rgb).all()) and float(rgb.min()) >= 0.0 and float(rgb.max()) <= 1.0 + 1e-5, 'sum w <= 1, colours in [0,1]'
    assert float(out['gradient_error'].min()) >= 0.0
    # determinism: same inputs -> bit-identical outputs, gradients reproducible to rounding
    net.zero_grad()
    loss_of(net, out, r['rgb']).backward()
    g1 = grads_of(net)
    net.zero_grad()
    z2, out2 = render(net, r)
    assert torch.equal(z, z2) and torch.equal(out['ray_rgb'], out2['ray_rgb']) and torch.equal(out['gradient_error'], out2['gradient_error'])
    loss_of(net, out2, r['rgb']).backward()
    g1b = grads_of(net)
    # (bias / scalar gradients are accumulated with floating-point atomics: reproducible to rounding, not bit for bit)
    assert float((g1 - g1b).abs().max()) <= 2e-6 * float(g1.abs().max()), 'backward is reproducible'
    # linearity of the backward pass in the loss scale
    net.zero_grad()
    _, out3 = render(net, r)
    (2.0 * loss_of(net, out3, r['rgb'])).backward()
    g2 = grads_of(net)
    assert float((g2 - 2.0 * g1).abs().max()) <= 1e-5 * float(g1.abs().max())
    # rays are independent units: rendering the two halves separately reproduces the full batch
    za, oa = render(net, r, sl=slice(0, 512))
    zb, ob = render(net, r, sl=slice(512, 1024))
    assert torch.equal(torch.cat([za, zb]), z)
    assert float((torch.cat([oa['ray_rgb'], ob['ray_rgb']]) - rgb).abs().max()) <= 1e-6


def test_rays_that_miss_the_unit_sphere():
    """No inner samples at all: renderer.py:574-575, 586-587, 600-601 (zeros(1) placeholders), outer NeRF only."""
    net = make_shape_net({'n_samples': 32, 'n_importance': 32})
    R = 37
    g = torch.Generator().manual_seed(1)
    o = torch.tensor([0.0, 0.0, 3.0]).repeat(R, 1) + 0.1 * torch.randn(R, 3, generator=g)
    d = torch.nn.functional.normalize(torch.tensor([1.0, 0.2, 0.0]).repeat(R, 1) + 0.1 * torch.randn(R, 3, generator=g), dim=-1)
    near, far = O.near_far_from_sphere(o, d)
    hp = O.synthetic_rays(R, seed=2)['human_poses']
    r = {'rays_o': o.to(DEV), 'rays_d': d.to(DEV), 'near': near.to(DEV), 'far': far.to(DEV), 'human_poses': hp.to(DEV)}
    z, out = render(net, r)
    assert net.engine.state['N_in'] == 0
    assert out['gradient_error'].shape == (1,) and float(out['gradient_error']) == 0.0
    assert float(torch.as_tensor(out['loss_occ']).sum()) == 0.0 and float(torch.as_tensor(out['std']).sum()) == 0.0
    assert bool(torch.isfinite(out['ray_rgb']).all())
    sd = build_params({'n_samples': 32, 'n_importance': 32})
    c = O.merged_cfg({'n_samples': 32, 'n_importance': 32})
    with torch.no_grad():
        want = O.render_core(sd, c, sd['color_network.FG_LUT'][0], o, d, z.cpu(), hp, O.get_anneal_val(c, STEP), STEP)
    assert float((out['ray_rgb'].cpu() - want['ray_rgb']).abs().max()) <= 1e-4 * float(want['ray_rgb'].abs().max()) + 2e-5
    net.zero_grad()
    torch.mean(out['ray_rgb']).backward()          # only the outer NeRF receives gradients
    P = dict(net.named_parameters())
    assert float(P['outer_nerf.rgb_linear.weight'].grad.abs().max()) > 0
    assert float(P['sdf_network.lin0.weight_v'].grad.abs().max()) == 0.0


@pytest.mark.parametrize('R', [1, 3, 129])
def test_odd_ray_counts(R):
    net = make_shape_net({'n_samples': 32, 'n_importance': 32})
    rays = O.synthetic_rays(max(R, 4), seed=6033)
    r = {k: v[:R].to(DEV).contiguous() for k, v in rays.items()}
    z, out = render(net, r)
    sd = build_params({'n_samples': 32, 'n_importance': 32})
    c = O.merged_cfg({'n_samples': 32, 'n_importance': 32})
    with torch.no_grad():
        want = O.render_core(sd, c, sd['color_network.FG_LUT'][0], rays['rays_o'][:R], rays['rays_d'][:R], z.cpu(), rays['human_poses'][:R],
                             O.get_anneal_val(c, STEP), STEP)
    assert out['ray_rgb'].shape == (R, 3)
    assert float((out['ray_rgb'].cpu() - want['ray_rgb']).abs().max()) <= 1e-4 * float(want['ray_rgb'].abs().max()) + 2e-5
    loss_of(net, out, r['rgb']).backward()
    assert all(bool(torch.isfinite(p.grad).all()) for p in net.parameters())


def test_material_all_rays_escape_and_single_point():
    """Stage II edge cases: a mesh far from the shaded points (every secondary ray escapes: no inner-light rows) and P = 1."""
    from nero_b200.material import NeROMaterialRenderer
    cfg = {'shader_cfg': {'human_lights': True, 'diffuse_sample_num': 16, 'specular_sample_num': 16}}
    verts, tris = OM.icosphere(1, 0.05)
    verts = verts + np.asarray([5.0, 5.0, 5.0], np.float32)
    net = NeROMaterialRenderer(cfg, is_train=False, mesh=(verts, tris))
    sd = build_material_params(cfg['shader_cfg'])
    net.load_state_dict(sd)
    net = net.cuda()
    rays = O.synthetic_rays(8, seed=4)
    pts = 0.5 * torch.nn.functional.normalize(rays['rays_o'], dim=-1)
    normals = torch.nn.functional.normalize(pts, dim=-1)
    tabs = (OM.direction_samples(16), OM.direction_samples(16))
    trace_fn = lambda o, d: OM.renderer_trace(verts, tris, o, d)
    for P in (8, 1):
        batch = {'pts': pts[:P], 'rays_d': rays['rays_d'][:P], 'normals': normals[:P], 'rgb': rays['rgb'][:P], 'human_poses': rays['human_poses'][:P]}
        rands = OM.draw_rands(P)
        net.zero_grad()
        out = net.shade_batch({k: v.to(DEV) for k, v in batch.items()}, 5000, {k: v.to(DEV) for k, v in rands.items()})
        assert net.engine.state['n_hit'] == 0 and net.engine.state['n_miss'] == P * 32
        p = {k: v.clone().requires_grad_(torch.is_floating_point(v) and not k.endswith('light_pts')) for k, v in sd.items()}
        want = OM.material_train_outputs(p, cfg, tabs, trace_fn, batch, 5000, rands)
        for k in ('rgb_pr', 'diffuse_light', 'specular_light', 'human_lights', 'loss_rgb'):
            assert float((out[k].detach().cpu() - want[k].detach()).abs().max()) <= 2e-4 * float(want[k].abs().max()) + 3e-5, k
        sum(torch.mean(v) for k, v in out.items() if k.startswith('loss')).backward()
        OM.material_training_loss(want).backward()
        G = dict(net.named_parameters())
        for k in ('shader_network.outer_light.6.bias', 'shader_network.human_light.6.bias', 'shader_network.roughness_predictor.6.bias'):
            a, b = G[k].grad.cpu(), p[k].grad
            assert float((a - b).abs().max()) <= 5e-3 * float(b.abs().max()) + 1e-9, k
        assert float(G['shader_network.inner_light.6.bias'].grad.abs().max()) == 0.0
