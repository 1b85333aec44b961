import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'oracle'), os.path.join(ROOT, 'tests')):
    sys.path.insert(0, p)
import torch, numpy as np
import nero_oracle as O
from helpers import load_golden, build_params, t, rays_from_golden, FIXTURE_CFGS
from nero_b200.renderer import NeROShapeRenderer
from nero_b200.ops import K, Mat
name = 'shape_bell_r32'
g = load_golden(name); cfg = FIXTURE_CFGS[name]
sd = build_params(cfg, int(g['seed']), int(g['pseed']))
net = NeROShapeRenderer(cfg, training=False); net.load_state_dict(sd); net = net.cuda()
rays = rays_from_golden(g); c = O.merged_cfg(cfg)
dev = torch.device('cuda')
cu = {k: v.to(dev).contiguous() for k, v in rays.items()}
e = net.engine; e.prepare_weights()
R = 32; n, nb, nimp, steps = c['n_samples'], c['n_bg_samples'], c['n_importance'], c['up_sample_steps']
S = n + nimp + nb; nn_ = nimp // steps
e._alloc(R, S); w = e.w
z_vals = torch.zeros(R, S, device=dev)
var = net.deviation_network.variance.detach()
K('nero_sample_init', cu['rays_o'], cu['rays_d'], cu['near'], cu['far'], R, n, nb, e.t_lin, e.t_bg, e.t_bg_lo, e.t_bg_hi, None, None,
  w['ZA'], 128, Mat(z_vals, n + nimp), S, w['SX0'], 64, w['SC'], 256)
zb, zo = O.base_z_vals(c, rays['near'], rays['far'])
print('base z diff', float((w['ZA'][:, :n].cpu() - zb).abs().max()), 'bg diff', float((z_vals[:, n + nimp:].cpu() - zo).abs().max()))
e.sdf.sdf_only(w['SX0'], w['SA'], w['SB'], w['SC'], w['SSDF'], None, R * n)
pts = rays['rays_o'][:, None] + rays['rays_d'][:, None] * zb[..., None]
with torch.no_grad():
    sdf_o = O.sdf_forward(sd, pts)[..., 0]
sdf_g = w['SSDF'][:R * n, 0].reshape(R, n).cpu()
print('coarse sdf diff', float((sdf_g - sdf_o).abs().max()))
cur_z, cur_sdf = w['ZA'][:, :n].cpu().clone(), sdf_g.clone()
cz_dev, cs_dev, lds = w['ZA'], w['SSDF'], n
nxt_z, nxt_sdf = w['ZB'], w['SDFB']
cur_n = n
for i in range(steps):
    last = i + 1 == steps
    K('nero_upsample', cu['rays_o'], cu['rays_d'], R, cz_dev, 128, cs_dev, lds, cur_n, nn_, var, float(64 * 2 ** i), 1, 0, w['NEWZ'], 32,
      None if last else w['SX0'], 64, None if last else w['SC'], 256, None)
    inv_s = torch.clamp(torch.exp(sd['deviation_network.variance'] * 10), max=64 * 2 ** i).expand(R, cur_n - 1)
    with torch.no_grad():
        nz_o = O.upsample(rays['rays_o'], rays['rays_d'], cur_z, cur_sdf, nn_, inv_s)
    nz_g = w['NEWZ'][:, :nn_].cpu()
    d = (nz_g - nz_o).abs()
    print(f'iter {i}: upsample new_z diff max {float(d.max()):.3e} frac>1e-5 {float((d>1e-5).float().mean()):.3f}  (inv_s {float(inv_s[0,0]):.3f})')
    if d.max() > 1e-3:
        r = int(d.max(1)[0].argmax()); print('  worst ray', r, 'gpu', nz_g[r].numpy(), 'oracle', nz_o[r].numpy())
    if last:
        K('nero_merge_samples', cz_dev, 128, None, 0, cur_n, w['NEWZ'], 32, None, 0, nn_, z_vals, S, None, 0, R)
        zc, _ = torch.sort(torch.cat([cur_z, nz_g], -1), -1)
        print('final merge diff', float((z_vals[:, :n + nimp].cpu() - zc).abs().max()))
    else:
        e.sdf.sdf_only(w['SX0'], w['SA'], w['SB'], w['SC'], w['SSDF'], None, R * nn_)
        npts = rays['rays_o'][:, None] + rays['rays_d'][:, None] * nz_g[..., None]
        with torch.no_grad():
            ns_o = O.sdf_forward(sd, npts.reshape(-1, 3))[..., 0].reshape(R, nn_)
        ns_g = w['SSDF'][:R * nn_, 0].reshape(R, nn_).cpu()
        print('   new sdf diff', float((ns_g - ns_o).abs().max()))
        K('nero_merge_samples', cz_dev, 128, cs_dev, lds, cur_n, w['NEWZ'], 32, w['SSDF'], nn_, nn_, nxt_z, 128, nxt_sdf, 128, R)
        zc, idx = torch.sort(torch.cat([cur_z, nz_g], -1), -1)
        sc = torch.gather(torch.cat([cur_sdf, ns_g], -1), 1, idx)
        print('   merge z diff', float((nxt_z[:, :cur_n + nn_].cpu() - zc).abs().max()), 'sdf diff', float((nxt_sdf[:, :cur_n + nn_].cpu() - sc).abs().max()))
        cur_z, cur_sdf = nxt_z[:, :cur_n + nn_].cpu().clone(), nxt_sdf[:, :cur_n + nn_].cpu().clone()
        cz_dev, nxt_z = nxt_z, (w['ZA'] if nxt_z is w['ZB'] else w['ZB'])
        cs_dev, nxt_sdf = nxt_sdf, (w['SDFA'] if nxt_sdf is w['SDFB'] else w['SDFB'])
        lds = 128
    cur_n += nn_
