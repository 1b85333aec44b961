"""Benchmark of the NeRO stage-I training hot path on B200 (BASELINE.json metric: train rays/sec at 128 samples/ray).

    python bench.py --gpus 1 --steps 20 --warmup 5                 # this repo's CUDA path
    torchrun ... bench.py --gpus N ...                             # ray-sharded data parallel, one rank per GPU
    python bench.py --impl reference ...                           # the reference algorithm (oracle port) on the host CPU

A "step" = one training step on the workload of BASELINE.json configs[1] ("bell shape stage, 1024 rays x 128 samples"):
sample_ray (64 coarse + 4x16 up-sampled + 32 background samples) + render_core forward + the YAML loss set
(charbonnier rgb + 0.1*eikonal + occlusion loss, step 30000 so the occlusion march is active) + backward + Adam step.
Prints ONE JSON line (see the driver contract in the task statement).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, 'oracle')):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np
import torch

A_SDF, A_SHADE, A_SHADE_H, A_NERF = 524544, 1211648, 1349888, 604160   # MAC / sample
STEP = 30000
RAYS_PER_GPU = 1024
WORKLOAD = 'bell_shape_stage_1024rays_x_(64+64)samples_+32bg_step30000_occ_on'


def algorithmic_flops(R, n_in, n_out, p_occ, human=False):
    """F = 2*[R*112*A_sdf + N_in*(6*A_sdf + 3*A_shade) + N_out*3*A_nerf + P_occ*80*A_sdf]   """
    a_sh = A_SHADE_H if human else A_SHADE
    return 2.0 * (R * 112 * A_SDF + n_in * (6 * A_SDF + 3 * a_sh) + n_out * 3 * A_NERF + p_occ * 80 * A_SDF)


def measured_peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        d = json.load(open(path))
        return d.get('bf16_tflops_sustained', 1386.8), d.get('hbm_gbs', 6569.6), 'measured'
    return 1400.0, 6650.0, 'fallback'


class ClockSampler(threading.Thread):
    def __init__(self, index=0):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.samples, self.reasons = index, False, [], set()
        self.max_mhz = None

    def run(self):
        q = 'clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,' \
            'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        while not self.stop_flag:
            try:
                o = subprocess.run(['nvidia-smi', f'--id={self.index}', f'--query-gpu={q}', '--format=csv,noheader,nounits'],
                                   capture_output=True, text=True, timeout=5).stdout.strip().split(',')
                self.samples.append(float(o[0]))
                self.max_mhz = float(o[1])
                for n, v in zip(names, o[2:]):
                    if 'Active' in v and 'Not' not in v:
                        self.reasons.add(n)
            except Exception:
                pass
            time.sleep(0.1)

    def result(self):
        return {'sm_mhz': float(np.median(self.samples)) if self.samples else None, 'sm_max_mhz': self.max_mhz,
                'reasons': sorted(self.reasons)}


def build_net(cfg, device):
    from nero_b200 import params as P, synthetic
    from nero_b200.renderer import NeROShapeRenderer
    sd = synthetic.perturb_params(P.build_shape_state_dict(cfg, seed=6033))
    net = NeROShapeRenderer(cfg, training=False)
    net.load_state_dict(sd)
    return net.to(device), sd


def synthetic_dataset(net, n_images, h, w, seed, device):
    """Fills the renderer's train ray table exactly like _init_dataset does from a database, from synthetic look-at
    cameras on the radius-3 sphere (no dataset exists offline): CPU pinned dirs/rgbs/idxs + poses."""
    g = torch.Generator().manual_seed(seed)
    o = 3.0 * torch.nn.functional.normalize(torch.randn(n_images, 3, generator=g), dim=-1)
    zc = torch.nn.functional.normalize(-o, dim=-1)
    up = torch.tensor([0.0, 0.0, 1.0]).expand(n_images, 3)
    xc = torch.nn.functional.normalize(torch.cross(up, zc, dim=-1) + 1e-3, dim=-1)
    yc = torch.cross(zc, xc, dim=-1)
    Rm = torch.stack([xc, yc, zc], 1)
    poses = torch.cat([Rm, -(Rm @ o[:, :, None])], -1)
    f = 1.2 * w
    K = torch.tensor([[f, 0, w / 2], [0, f, h / 2], [0, 0, 1.0]]).expand(n_images, 3, 3)
    info = {'imgs': torch.rand(n_images, 3, h, w, generator=g), 'Ks': K.contiguous(), 'poses': poses}
    net.train_batch, net.train_poses, net.tbn, _, _ = net._construct_ray_batch(info)
    net.train_poses = net.train_poses.float()
    gs = torch.Generator().manual_seed(seed + 1)
    idx = torch.randperm(net.tbn, generator=gs)
    net.train_batch = {k: v[idx].pin_memory() for k, v in net.train_batch.items()}
    net.train_batch_i = 0
    net._shuffle_train_batch = lambda: setattr(net, 'train_batch_i', 0)


def training_loss(net, out, rgb_gt):
    loss = torch.mean(net.compute_rgb_loss(out['ray_rgb'], rgb_gt)) + torch.mean(out['gradient_error'] * 0.1)
    if 'loss_occ' in out:
        loss = loss + torch.mean(out['loss_occ'])
    return loss


def run_ours(args):
    import torch.distributed as dist
    from nero_b200 import ops
    assert not ops.DEBUG_GEMM and not ops.DRY_RUN, 'bench.py measures the real kernels only'
    rank, world = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1))
    local = int(os.environ.get('LOCAL_RANK', 0))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    from nero_b200 import synthetic
    bear = args.workload == 'bear'        # BASELINE.json configs[2] / SURVEY 8d config 3: human light on, 2048 rays on one GPU
    cfg = {'shader_config': {'human_light': True}} if bear else {}
    net, sd = build_net(cfg, dev)
    # weak scaling: fixed rays per GPU, global batch = R * world.  bear: 2048 rays on one GPU (configs[2]); under torchrun
    # 1024 per GPU, i.e. 8192 rays on 8 GPUs (configs[4])
    R = (2048 if world == 1 else 1024) if bear else RAYS_PER_GPU
    rays = synthetic.synthetic_rays(R * world, seed=6033)
    r = {k: v[rank * R:(rank + 1) * R].to(dev).contiguous() for k, v in rays.items()}
    from nero_b200.optim import FlatAdam
    opt = FlatAdam(net, lr=5e-4 * 0.05)        # one nero_adam_flat launch over the flat parameter / gradient buffers
    car = net.get_anneal_val(STEP)
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)

    def sync_grads():
        if world > 1:
            flat = net.engine.grads.flat
            dist.all_reduce(flat)          # ONE NCCL all-reduce over the flat fp32 gradient buffer per step
            flat.div_(world)

    def resident_step():
        opt.zero_grad(set_to_none=True)
        ri = torch.rand([R, 1], device=dev, generator=gen)
        rb = torch.rand([R, 32], device=dev, generator=gen)
        z = net.sample_ray(r['rays_o'], r['rays_d'], r['near'], r['far'], 1.0, ri, rb)
        out = net.render_core(r['rays_o'], r['rays_d'], z, r['human_poses'], car, STEP)
        loss = training_loss(net, out, r['rgb'])
        loss.backward()
        sync_grads()
        opt.step()
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        resident_step()
    barrier()
    sampler = ClockSampler(local)      # clocks / throttle reasons of rank 0's GPU only (one nvidia-smi poller per job, not per rank)
    if rank == 0:
        sampler.start()
    l0 = ops.launch_count
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        resident_step()
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1) / args.steps
    launches = (ops.launch_count - l0) // args.steps
    st = net.engine.state
    n_in, n_out, p_occ = st['N_in'], int(net.engine.w['n_out'].item()), st['P']

    # ---- end to end through the public API with host buffers (H2D of the ray batch + D2H of the loss every step)
    net.cfg['train_ray_num'] = R
    synthetic_dataset(net, 8, 128, 128, seed=99 + rank, device=dev)
    h2d = R * (12 + 12 + 8)

    def e2e_step(step):
        opt.zero_grad(set_to_none=True)
        out = net({'step': step})
        loss = out['loss_rgb'].mean() + torch.mean(out['gradient_error'] * 0.1) + torch.mean(out['loss_occ'])
        loss.backward()
        sync_grads()
        opt.step()
        return float(loss.detach().cpu())          # the trainer's per-step host read (train/trainer.py:168)

    for _ in range(max(3, args.warmup // 2)):
        e2e_step(STEP)
    barrier()
    ev0.record()
    for _ in range(args.steps):
        e2e_step(STEP)
    ev1.record()
    barrier()
    ms_e2e = ev0.elapsed_time(ev1) / args.steps
    sampler.stop_flag = True
    if rank == 0:
        sampler.join(timeout=2)

    # ---- roofline of the dominant kernel (tcgen05 linear, N=256 tiles): per-launch CUDA events on the launch stream
    prof = None
    if rank == 0:
        prof = profile_linear(net, r, car)
    t = torch.tensor([ms, ms_e2e], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e = float(t[0]), float(t[1])
    if rank == 0:
        peak_tf, peak_bw, which = measured_peaks()
        F = algorithmic_flops(R, n_in, n_out, p_occ, human=bear)
        line = {
            'metric': 'train rays/sec (128 samples/ray)', 'value': R * world / (ms * 1e-3), 'unit': 'rays/s', 'n_gpus': world,
            'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'f32 (split-bf16 x3 tensor-core MMAs, fp32 accumulate)', 'data': 'synthetic',
            'config': {'workload': WORKLOAD.replace('bell', 'bear_human_light').replace('1024rays', f'{R}rays') if bear else WORKLOAD, 'rays_per_gpu': R, 'global_rays': R * world, 'n_in': n_in, 'n_out': n_out, 'p_occ': p_occ,
                       'parallelism': f'ray-sharded dp{world}, one NCCL all-reduce of the flat grad buffer' if world > 1 else 'single gpu',
                       'l2': 'per-step working set ~6 GB of activations >> 126 MB L2 (inputs larger than L2)',
                       'optimizer': 'Adam (nero_adam_flat over the flat parameter buffer) inside the timed region'},
            'e2e': {'value': R * world / (ms_e2e * 1e-3), 'unit': 'rays/s', 'ms_per_step': ms_e2e, 'h2d_bytes_per_step': h2d,
                    'd2h_bytes_per_step': 4},
            'gpu_launches': int(launches),
            'clocks': sampler.result(),
            'step_tensor_tflops': F / (ms * 1e-3) / 1e12,
            'step_tensor_frac_of_bf16_peak': F / (ms * 1e-3) / 1e12 / peak_tf,
            'roofline': None, 'cpu_baseline': None,
        }
        if prof is not None and prof['reverse_sweep'] is not None:
            one = prof['reverse_sweep']
            ach = one['flops'] / one['seconds'] / 1e12
            traffic = None
            tp = os.path.join(ROOT, 'profiles', 'r01_chain_traffic.json')
            if os.path.exists(tp):
                traffic = json.load(open(tp)).get('dram_bytes_per_launch')
            line['roofline'] = {'bound': 'tensor', 'kernel': 'umma_chain_kernel: SDF reverse-sweep chain (8 fused 256-wide layers, one launch)',
                                'achieved': ach, 'peak': peak_tf, 'unit': 'TFLOP/s', 'frac': ach / peak_tf,
                                'frac_of_split3_ceiling': ach / (peak_tf / 3.0), 'peak_source': which + ' bf16 sustained',
                                'launch_us': one['seconds'] * 1e6, 'rows': one['rows'], 'traffic': traffic,
                                'all_chain_launches': {'launches': prof['launches'], 'total_ms': prof['seconds'] * 1e3,
                                                       'achieved': prof['flops'] / prof['seconds'] / 1e12,
                                                       'frac': prof['flops'] / prof['seconds'] / 1e12 / peak_tf},
                                'note': 'achieved counts ALGORITHMIC fp32 GEMM flops (2*M*K*N of the un-padded layers); the '
                                        'split-bf16 scheme issues 3 bf16 MMAs per product, so 1/3 of peak is its ceiling'}
        if not args.no_cpu and world == 1:      # the CPU leg runs on rank 0 at N=1 only
            line['cpu_baseline'] = cpu_baseline(rays_n=128, steps=1)
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def profile_linear(net, r, car):
    """One extra training step with a CUDA-event pair (on the launch stream) around every fused MLP-chain launch
    (ops.PROFILE hook).  Returns the aggregate over all chain launches and the SDF reverse-sweep chain alone (the
    launch whose ncu capture is committed under profiles/)."""
    from nero_b200 import ops
    ops.PROFILE = []
    try:
        net.zero_grad()
        z = net.sample_ray(r['rays_o'], r['rays_d'], r['near'], r['far'], 0)
        out = net.render_core(r['rays_o'], r['rays_d'], z, r['human_poses'], car, STEP)
        training_loss(net, out, r['rgb']).backward()
        torch.cuda.synchronize()
        recs = ops.PROFILE
    finally:
        ops.PROFILE = None
    w = net.engine.w
    counts = {w['n_in'].data_ptr(): net.engine.state['N_in'], w['n_out'].data_ptr(): int(w['n_out'].item())}
    tot_f = tot_s = 0.0
    one = None
    for tag, e0, e1, fl_row, mp, m_cap in recs:
        M = m_cap if mp is None else counts.get(mp, m_cap)
        sec = e0.elapsed_time(e1) * 1e-3
        tot_f += fl_row * M
        tot_s += sec
        if tag == 'sdf_reverse_sweep':
            one = {'flops': fl_row * M, 'seconds': sec, 'rows': M}
    return {'flops': tot_f, 'seconds': tot_s, 'launches': len(recs), 'reverse_sweep': one}


def cpu_baseline(rays_n=128, steps=1, threads=None):
    """The reference algorithm (oracle port, bit-exact to the reference on CPU) timed on the host cores: one training
    step (sample_ray + render_core + loss + backward) on a bounded sample of the same workload."""
    import nero_oracle as O
    from nero_b200 import params as P
    cfg = {}
    sd = O.perturb_params(P.build_shape_state_dict(cfg, seed=6033))
    rays = O.synthetic_rays(rays_n, seed=6033)
    c = O.merged_cfg(cfg)
    lut = sd['color_network.FG_LUT'][0]
    car = O.get_anneal_val(c, STEP)

    def one():
        p = {k: v.clone().requires_grad_(torch.is_floating_point(v) and not k.endswith('FG_LUT')) for k, v in sd.items()}
        g = torch.Generator().manual_seed(0)
        ri, rb = torch.rand([rays_n, 1], generator=g), torch.rand([rays_n, 32], generator=g)
        out = O.render(p, cfg, lut, rays['rays_o'], rays['rays_d'], rays['near'], rays['far'], rays['human_poses'], car, STEP,
                       rand_inner=ri, rand_bg=rb)
        O.training_loss(out, rays['rgb'], c, STEP).backward()
    # "all the host threads it can use": torch's intra-op pool is slower when oversubscribed on these small ops, so
    # pick the best of a few pool sizes on a short calibration slice and report the count actually used
    ncpu = os.cpu_count() or 1
    cands = [threads] if threads else sorted({c for c in (8, 16, 32, 64, ncpu) if c <= ncpu})
    if len(cands) > 1:
        full_n, rays_cal = rays_n, max(16, rays_n // 8)
        best = None
        for c_ in cands:
            torch.set_num_threads(c_)
            rays_n = rays_cal
            rays_save = rays
            rays = {k: v[:rays_cal] for k, v in rays_save.items()}
            one()
            t0 = time.time()
            one()
            dt_c = time.time() - t0
            rays, rays_n = rays_save, full_n
            if best is None or dt_c < best[0]:
                best = (dt_c, c_)
        cores = best[1]
    else:
        cores = cands[0]
    torch.set_num_threads(cores)
    one()   # warm-up
    t0 = time.time()
    for _ in range(steps):
        one()
    dt = (time.time() - t0) / steps
    model = ''
    try:
        model = [l.split(':')[1].strip() for l in open('/proc/cpuinfo') if l.startswith('model name')][0]
    except Exception:
        pass
    return {'value': rays_n / dt, 'unit': 'rays/s', 'cores': cores, 'kind': 'port', 'seconds_per_step': dt, 'cpu': model,
            'sample': f'{rays_n} rays x (64+64)+32bg samples, step {STEP} (occlusion march on), fwd+bwd, {steps} timed step(s) after 1 warm-up'}


def run_reference(args):
    rank = int(os.environ.get('RANK', 0))
    if rank != 0:
        return
    rays_n = 128
    cb = cpu_baseline(rays_n=rays_n, steps=max(1, min(args.steps, 3)))
    line = {'metric': 'train rays/sec (128 samples/ray)', 'value': cb['value'], 'unit': 'rays/s', 'n_gpus': args.gpus,
            'steps': max(1, min(args.steps, 3)), 'warmup': 1, 'ms_per_step': cb['seconds_per_step'] * 1e3, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic', 'impl': 'reference',
            'config': {'workload': WORKLOAD, 'sample': cb['sample'],
                       'note': 'the reference is pure PyTorch and is not importable on the GPU box; this is its bit-exact '
                               'CPU port (oracle/nero_oracle.py, pinned by tests/golden) on all host cores'},
            'cpu_baseline': cb,
            'e2e': {'value': cb['value'], 'unit': 'rays/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours')
    ap.add_argument('--workload', default='bell', choices=['bell', 'bear'],
                    help="bell = BASELINE.json configs[1] (the headline, default); bear = configs[2]: human light, 2048 rays")
    ap.add_argument('--no-cpu', action='store_true', help='skip the cpu_baseline leg (profiling runs)')
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()
