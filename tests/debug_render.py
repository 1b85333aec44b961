"""Stage-by-stage comparison of the CUDA path against the oracle on a golden fixture (diagnostic script, GPU)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'oracle'), os.path.join(ROOT, 'tests')):
    sys.path.insert(0, p)
import numpy as np
import torch
import nero_oracle as O
from helpers import load_golden, build_params, t, rays_from_golden, FIXTURE_CFGS

name = sys.argv[1] if len(sys.argv) > 1 else 'shape_bell_r32'
step = int(sys.argv[2]) if len(sys.argv) > 2 else 30000
from nero_b200.renderer import NeROShapeRenderer
from nero_b200 import ops
from nero_b200.engine import *

g = load_golden(name)
cfg = FIXTURE_CFGS[name]
sd = build_params(cfg, int(g['seed']), int(g['pseed']))
net = NeROShapeRenderer(cfg, training=False)
net.load_state_dict(sd)
DRY = bool(os.environ.get('NERO_DRY_RUN'))
if not DRY:
    net = net.cuda()
rays = rays_from_golden(g)
c = O.merged_cfg(cfg)
dev = torch.device('cpu' if os.environ.get('NERO_DRY_RUN') else 'cuda')
cu = {k: v.to(dev).contiguous() for k, v in rays.items()}
R = cu['rays_o'].shape[0]


def rep(nm, got, want, mask=None):
    got = got.detach().float().cpu(); want = want.detach().float().cpu()
    if got.shape != want.shape:
        print(nm, 'SHAPE MISMATCH', tuple(got.shape), tuple(want.shape)); return
    e = (got - want).abs()
    rel = e / (want.abs() + 1e-3)
    print(f'{nm:28s} shape {tuple(got.shape)} max_abs {float(e.max()):.3e} max_rel(+1e-3) {float(rel.max()):.3e} mean_abs {float(e.mean()):.3e} |want|max {float(want.abs().max()):.3e}')


# ---- sampling
z_gold = t(g['z_vals'])
z = net.sample_ray(cu['rays_o'], cu['rays_d'], cu['near'], cu['far'], 0)
(None if os.environ.get('NERO_DRY_RUN') else torch.cuda.synchronize())
dz = (z.cpu() - z_gold).abs()
print('z_vals: max', float(dz.max()), 'frac>1e-4', float((dz > 1e-4).float().mean()), 'frac>1e-3', float((dz > 1e-3).float().mean()))
zp = net.sample_ray(cu['rays_o'], cu['rays_d'], cu['near'], cu['far'], 1.0, t(g['rand_inner']).to(dev), t(g['rand_bg']).to(dev))
dzp = (zp.cpu() - t(g['z_vals_perturbed'])).abs()
print('z_vals perturbed: max', float(dzp.max()), 'frac>1e-4', float((dzp > 1e-4).float().mean()))
# sdf on coarse samples vs oracle
with torch.no_grad():
    zb, _ = O.base_z_vals(c, rays['near'], rays['far'])
    pts = rays['rays_o'][:, None] + rays['rays_d'][:, None] * zb[..., None]
    sdf_o = O.sdf_forward(sd, pts)[..., 0]

# ---- render core with golden z
car = O.get_anneal_val(c, step)
p = {k: v.clone().requires_grad_(torch.is_floating_point(v) and not k.endswith('FG_LUT')) for k, v in sd.items()}
lut = sd['color_network.FG_LUT'][0]
out_o = O.render_core(p, c, lut, rays['rays_o'], rays['rays_d'], z_gold, rays['human_poses'], car, step, return_intermediates=True)
it = out_o['_inter']
loss_o = O.training_loss(out_o, rays['rgb'], c, step)
loss_o.backward()

net.zero_grad()
out = net.render_core(cu['rays_o'], cu['rays_d'], z_gold.to(dev), cu['human_poses'], car, step)
(None if os.environ.get('NERO_DRY_RUN') else torch.cuda.synchronize())
w = net.engine.w
N = net.engine.state['N_in']
print('N_in', N, 'oracle', int(it['inner_mask'].sum()), 'P', net.engine.state['P'])
rep('sdf', w['Y8'][:N, 260], it['sdf'])
rep('feats', w['Y8'][:N, :256], it['feats'])
rep('gradients', w['G'][:N, :3], it['gradients'])
rep('metallic', w['OUTS'][:N, 0], it['sh_metallic'][:, 0])
rep('roughness', w['OUTS'][:N, 4], it['sh_roughness'][:, 0])
rep('albedo', w['OUTS'][:N, 8:11], it['sh_albedo'])
rep('ide_refl', w['E'][:N, 92:164], it['sh_ide_refl'])
rep('ide_normal', w['E'][:N, 164:236], it['sh_ide_normal'])
rep('diffuse_light', w['OUTS'][:N, 12:15], it['sh_diffuse_light'])
rep('direct_light', w['OUTS'][:N, 16:19], it['sh_direct_light'])
rep('indirect_light', w['OUTS'][:N, 20:23], it['sh_indirect_light'])
rep('occ_prob', w['OCCP'][:N], it['occ_prob'][:, 0])
rep('NoV', w['GEO'][:N, 3], it['sh_NoV'][:, 0])
rep('reflective', w['REFL'][:N, :3], it['reflective'])
rep('inner_color', w['COLOR_IN'][:N, :3], it['inner_color'])
rep('inner_alpha', w['ALPHA_IN'][:N], it['inner_alpha'])
No = int(w['n_out'].item())
rep('outer_alpha', w['ALPHA_OUT'][:No], it['outer_alpha'])
rep('outer_color', w['COLOR_OUT'][:No, :3], it['outer_color'])
rep('ray_rgb', out['ray_rgb'], out_o['ray_rgb'])
rep('gradient_error', out['gradient_error'], out_o['gradient_error'])
rep('loss_occ', out['loss_occ'].reshape(-1), out_o['loss_occ'].reshape(-1))
if 'sh_human_light' in it and torch.is_tensor(it['sh_human_light']):
    rep('human hl', w['OUTS'][:N, 28:31] * w['GEO'][:N, 7:8], it['sh_human_light'])

loss = O.training_loss(out, cu['rgb'], c, step)
print('loss', float(loss), 'oracle', float(loss_o), 'golden', float(g[f's{step}_loss']))
loss.backward()
(None if os.environ.get('NERO_DRY_RUN') else torch.cuda.synchronize())
worst = []
for n_, q in net.named_parameters():
    go = p[n_].grad
    if go is None:
        go = torch.zeros_like(p[n_])
    gm = q.grad.detach().cpu()
    e = float((gm - go).abs().max())
    sc = float(go.abs().max()) + 1e-12
    worst.append((e / sc, n_, e, sc, float(gm.norm()), float(go.norm())))
worst.sort(reverse=True)
for r_ in worst[:25]:
    print('grad %-52s relmax %.3e abs %.3e scale %.3e |g| %.4e oracle %.4e' % (r_[1], r_[0], r_[2], r_[3], r_[4], r_[5]))
print('median rel', np.median([r_[0] for r_ in worst]), 'launches', ops.launch_count)
