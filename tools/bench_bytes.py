"""HBM-bound kernels of the stage-I step, timed one by one (run on the GPU box).

One training step of the bench workload runs with a CUDA-event pair around every C-ABI launch (a wrapper around
engine.K installed here, not in the product); the byte-moving kernels are reported with their ALGORITHMIC bytes per launch
(unique inputs + outputs, from the sample counts of the step) and the resulting GB/s against the measured HBM peak
(MEASURED_PEAKS.json, else the 7.7 TB/s spec figure).  The stand-alone encodings nero_pe / nero_ide are timed on the same
number of rows.  Prints one JSON object."""
import os, sys, json, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'oracle')):
    sys.path.insert(0, p)
import torch
import bench
from nero_b200 import synthetic as O, ops, engine as E

dev = torch.device('cuda')
bear = os.environ.get('WORKLOAD', 'bell') == 'bear'
net, _ = bench.build_net({'shader_config': {'human_light': True}} if bear else {}, dev)
R = int(os.environ.get('RAYS', 2048 if bear else 1024))
rays = O.synthetic_rays(R, seed=6033)
r = {k: v.to(dev).contiguous() for k, v in rays.items()}
car = net.get_anneal_val(bench.STEP)
records = None
orig_K = E.K


def timed_K(name, *args):
    if records is None:
        return orig_K(name, *args)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    orig_K(name, *args)
    e1.record()
    records.append((name, e0, e1))


E.K = timed_K


def step():
    net.zero_grad()
    z = net.sample_ray(r['rays_o'], r['rays_d'], r['near'], r['far'], 0)
    out = net.render_core(r['rays_o'], r['rays_d'], z, r['human_poses'], car, bench.STEP)
    bench.training_loss(net, out, r['rgb']).backward()
    torch.cuda.synchronize()


for _ in range(3):
    step()
records = []
step()
st = net.engine.state
S = st['S']
n_in, n_out = st['N_in'], int(net.engine.w['n_out'].item())
P = st['P']
agg = collections.OrderedDict()
for name, e0, e1 in records:
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += e0.elapsed_time(e1) * 1e3
human = 24 if bear else 0
F = 4
# algorithmic bytes per launch (unique inputs + outputs), floats -> bytes
alg = {
    'nero_ray_fill': F * (R * S + 6 * R + n_in * (39 + 39 + 3 + 4 + 1) + n_out * (84 + 84 + 27 + 2) + R * S),
    'nero_shade_prep_fwd': F * n_in * (4 + 4 + 1 + 1 + 39 + 51 + 72 + 72 + 8 + human),
    'nero_shade_prep_bwd': F * n_in * (4 + 4 + 1 + 1 + 8 + 72 + 72 + 72 + human + 1 + 4 + 2),
    'nero_shade_combine_fwd': F * n_in * (16 + 8 + 4 + 1 + 4),
    'nero_shade_combine_bwd': F * n_in * (16 + 8 + 4 + 1 + 16 + 1),
    'nero_composite_fwd': F * (R * S + 5 * (n_in + n_out) + 3 * R),
    'nero_composite_bwd': F * (R * S + 5 * (n_in + n_out) + 3 * R + 5 * (n_in + n_out)),
    'nero_sdf_alpha_fwd': F * n_in * (1 + 4 + 1 + 1 + 2),
    'nero_sdf_alpha_bwd': F * n_in * (1 + 4 + 1 + 1 + 2 + 1 + 4),
    'nero_dact_times_row': F * n_in * 512,
    'nero_row_axpy': F * n_in * (1 + 256 * 3),
    'nero_pe_grad': F * n_in * (39 * 3 + 4),
    'nero_pe_tangent': F * n_in * (39 * 3 + 4),
    'nero_nerf_post_fwd': F * n_out * (1 + 3 + 1 + 1 + 4),
    'nero_nerf_post_bwd': F * n_out * (1 + 3 + 1 + 1 + 4 + 1 + 3),
}
peak = 7700.0
try:
    peak = float(json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))['hbm_gbs'])
    src = 'MEASURED_PEAKS.json'
except Exception:
    src = 'spec fallback'
out = {'workload': 'bear' if bear else 'bell', 'rays': R, 'n_in': n_in, 'n_out': n_out, 'p_occ': P, 'hbm_peak_gbs': peak, 'peak_source': src,
       'kernels': {}}
for name, (n, us) in agg.items():
    ent = {'launches': n, 'us_per_launch': round(us / n, 2)}
    if name in alg:
        gbs = alg[name] / (us / n * 1e-6) / 1e9
        ent.update(algorithmic_mb=round(alg[name] / 1e6, 2), gbs=round(gbs, 1), frac_of_hbm_peak=round(gbs / peak, 3))
    out['kernels'][name] = ent


def time_fn(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps


M = n_in
x = torch.randn(M, 3, device=dev) * 0.5
d = torch.nn.functional.normalize(torch.randn(M, 3, device=dev), dim=-1)
k = torch.rand(M, 1, device=dev)
for name, fn, nbytes in [('nero_pe (L=6, d=3)', lambda: ops.positional_encoding(x, 6), F * M * (3 + 39)),
                         ('nero_pe (L=8, d=3)', lambda: ops.positional_encoding(x, 8), F * M * (3 + 51)),
                         ('nero_ide', lambda: ops.integrated_dir_enc(d, k), F * M * (4 + 72))]:
    us = time_fn(fn)
    out['kernels'][name] = {'rows': M, 'us_per_launch': round(us, 2), 'algorithmic_mb': round(nbytes / 1e6, 2),
                            'gbs': round(nbytes / (us * 1e-6) / 1e9, 1), 'frac_of_hbm_peak': round(nbytes / (us * 1e-6) / 1e9 / peak, 3),
                            'note': 'includes the torch.empty of the output'}
print(json.dumps(out))
