"""Stress check (GPU box): the same training step N times; forward outputs must be bit-identical, parameter gradients equal to
rounding (fp32 atomics).  Prints which parameters deviate, if any -- a deviation beyond rounding means a data race."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'oracle')):
    sys.path.insert(0, p)
import torch
import bench
from nero_b200 import synthetic as O
dev = torch.device('cuda')
bear = os.environ.get('WORKLOAD', 'bell') == 'bear'
net, _ = bench.build_net({'shader_config': {'human_light': True}} if bear else {}, dev)
R = int(os.environ.get('RAYS', 1024))
N = int(sys.argv[1]) if len(sys.argv) > 1 else 30
rays = O.synthetic_rays(R, seed=6033)
r = {k: v.to(dev).contiguous() for k, v in rays.items()}
car = net.get_anneal_val(bench.STEP)
names = [n for n, _ in net.named_parameters()]
ref = None
bad = 0
for it in range(N):
    torch.manual_seed(0)
    net.zero_grad()
    z = net.sample_ray(r['rays_o'], r['rays_d'], r['near'], r['far'], 0)
    out = net.render_core(r['rays_o'], r['rays_d'], z, r['human_poses'], car, bench.STEP)
    scale = 1.0 + (it % 3)                       # the backward pass is linear in the loss scale
    (scale * bench.training_loss(net, out, r['rgb'])).backward()
    torch.cuda.synchronize()
    cur = dict(z=z.clone(), rgb=out['ray_rgb'].detach().clone(), gerr=out['gradient_error'].detach().clone(),
               locc=out['loss_occ'].detach().clone(), g=[p.grad.detach().clone() / scale for p in net.parameters()])
    if ref is None:
        ref = cur
        continue
    msgs = []
    for k in ('z', 'rgb', 'gerr'):
        if not torch.equal(cur[k], ref[k]):
            msgs.append(f'{k}: max diff {float((cur[k] - ref[k]).abs().max()):.3e}')
    if float((cur['locc'] - ref['locc']).abs().max()) > 1e-6 * float(ref['locc'].abs().max()):      # an atomic fp32 sum
        msgs.append(f"locc: max diff {float((cur['locc'] - ref['locc']).abs().max()):.3e}")
    for n, a, b in zip(names, cur['g'], ref['g']):
        d = float((a - b).abs().max())
        if d > 2e-5 * float(b.abs().max()) + 1e-9:
            msgs.append(f'grad {n}: max diff {d:.3e} of {float(b.abs().max()):.3e}')
    if msgs:
        bad += 1
        print(f'iteration {it} (scale {scale}):', '; '.join(msgs[:12]))
print(f'{bad} deviating iterations of {N - 1}')
