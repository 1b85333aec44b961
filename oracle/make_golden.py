"""Generate tests/golden/*.npz by running the UNMODIFIED reference (/root/reference) on CPU via ref_shim.

Run in the build container only (the reference does not exist on the GPU box):
    python oracle/make_golden.py
The fixtures are small (inputs, outputs, per-parameter gradient norms, parameter checksums); the parameters
themselves are regenerated from the seed by nero_b200.params (bit-identical to the reference's initialisers,
asserted here) plus oracle.perturb_params.
TEST INFRASTRUCTURE ONLY.
"""
import os
import sys
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)
warnings.filterwarnings('ignore')

import ref_shim  # noqa: E402
import nero_oracle as O  # noqa: E402
import nero_oracle_mat as OM  # noqa: E402

GOLD = os.path.join(ROOT, 'tests', 'golden')


def npy(d):
    return {k: (v.detach().cpu().numpy() if torch.is_tensor(v) else np.asarray(v)) for k, v in d.items()}


def param_checksums(sd):
    keys = sorted(k for k in sd if not k.endswith('FG_LUT'))
    return np.array([[float(sd[k].double().sum()), float(sd[k].double().abs().sum())] for k in keys])


def make_encoding_kats():
    ref_shim.install()
    from network.field import get_embedder, IPE, sample_pdf
    from utils.ref_utils import generate_ide_fn
    from utils.raw_utils import linear_to_srgb
    g = torch.Generator().manual_seed(11)
    x = torch.rand(64, 3, generator=g) * 2 - 1
    x[0] = torch.tensor([0.1, 0.2, 0.3])
    out = {'x': x}
    for L in (4, 6, 8):
        out[f'pe{L}'] = get_embedder(L, 3)[0](x)
    x4 = torch.rand(32, 4, generator=g) * 2 - 1
    out['x4'] = x4
    out['pe10_4'] = get_embedder(10, 4)[0](x4)
    ide = generate_ide_fn(5)
    dirs = torch.nn.functional.normalize(torch.randn(96, 3, generator=g), dim=-1)
    dirs[0] = torch.nn.functional.normalize(torch.tensor([1e-6, 1e-6, 1.0]), dim=-1)
    dirs[1] = torch.tensor([1.0, 0.0, 0.0])
    dirs[2] = torch.tensor([0.0, -1.0, 0.0])
    kap = torch.rand(96, 1, generator=g)
    kap[:8] = 0.0
    kap[8:16] = 1.0
    out.update(ide_dirs=dirs, ide_kappa=kap, ide=ide(dirs, kap))
    mean = torch.randn(40, 2, generator=g)
    var = torch.rand(40, 2, generator=g)
    out.update(ipe_mean=mean, ipe_var=var, ipe=IPE(mean, var, 0, 6))
    lin = torch.cat([torch.linspace(-0.1, 1.5, 200), torch.tensor([0.0031308, 0.0031307, 0.0031309])])
    out.update(srgb_in=lin, srgb=linear_to_srgb(lin))
    bins = torch.sort(torch.rand(24, 33, generator=g), -1)[0]
    w = torch.rand(24, 32, generator=g) ** 4
    w[0] = 0.0
    out.update(pdf_bins=bins, pdf_w=w, pdf_out=sample_pdf(bins, w, 16, det=True))
    np.savez_compressed(os.path.join(GOLD, 'kat_encodings.npz'), **npy(out))
    print('kat_encodings', {k: tuple(v.shape) for k, v in out.items()})


def make_shape_fixture(name, cfg, R, steps, seed=6033, pseed=7):
    from nero_b200 import params as P  # product-side containers: must equal the reference's init
    net = ref_shim.build_reference_shape_renderer(cfg, seed=seed)
    sd_ref = {k: v.detach().clone() for k, v in net.state_dict().items()}
    mine = P.build_shape_state_dict(cfg, seed=seed)
    assert list(mine.keys()) == list(sd_ref.keys())
    assert all(torch.equal(mine[k], sd_ref[k]) for k in sd_ref), 'nero_b200.params init != reference init'
    sd = O.perturb_params(sd_ref, seed=pseed)
    net.load_state_dict(sd)
    rays = O.synthetic_rays(R, seed=seed)
    c = O.merged_cfg(cfg)
    out = {'param_checksums': param_checksums(sd), 'R': R, 'seed': seed, 'pseed': pseed}
    out.update({'in_' + k: v for k, v in rays.items()})
    with torch.no_grad():
        z = net.sample_ray(rays['rays_o'], rays['rays_d'], rays['near'], rays['far'], 0)
    out['z_vals'] = z
    # a perturbed-sampling variant with recorded draws (renderer.py:416,422 order)
    torch.manual_seed(seed + 1)
    with torch.no_grad():
        zp = net.sample_ray(rays['rays_o'], rays['rays_d'], rays['near'], rays['far'], 1.0)
    torch.manual_seed(seed + 1)
    out['rand_inner'] = torch.rand([R, 1])
    out['rand_bg'] = torch.rand([R, c['n_bg_samples']])
    out['z_vals_perturbed'] = zp
    names = [n for n, _ in net.named_parameters()]
    for step in steps:
        car = O.get_anneal_val(c, step)
        net.zero_grad()
        o = net.render_core(rays['rays_o'], rays['rays_d'], z, rays['human_poses'], cos_anneal_ratio=car, step=step,
                            is_train=True)
        loss = O.training_loss(o, rays['rgb'], c, step)
        loss.backward()
        pre = f's{step}_'
        for k, v in o.items():
            out[pre + k] = v
        out[pre + 'loss'] = loss
        out[pre + 'grad_norms'] = np.array([float(p.grad.double().norm()) if p.grad is not None else 0.0
                                            for _, p in net.named_parameters()])
        # a few full gradients (small tensors) for element-wise checks
        gd = dict(net.named_parameters())
        for k in ['deviation_network.variance', 'sdf_network.lin8.bias', 'sdf_network.lin0.weight_g',
                  'color_network.albedo_predictor.6.weight_v', 'color_network.inner_weight.6.bias',
                  'outer_nerf.rgb_linear.weight', 'sdf_network.lin4.weight_g']:
            if gd[k].grad is not None:
                out[pre + 'grad::' + k] = gd[k].grad.clone()
    out['param_names'] = np.array(names)
    np.savez_compressed(os.path.join(GOLD, name + '.npz'), **npy(out))
    print(name, 'saved; loss', {s: float(out[f's{s}_loss']) for s in steps},
          'occ', {s: float(out[f's{s}_loss_occ'].mean()) for s in steps})


def make_validation_fixture(name, cfg, R, step, seed=6033, pseed=7):
    """is_train=False render (network/renderer.py:465-482, 602-604): depth / normal / shading intermediates /
    occ_prob_gt on the expected surface points."""
    net = ref_shim.build_reference_shape_renderer(cfg, seed=seed)
    sd = O.perturb_params({k: v.detach().clone() for k, v in net.state_dict().items()}, seed=pseed)
    net.load_state_dict(sd)
    rays = O.synthetic_rays(R, seed=seed)
    c = O.merged_cfg(cfg)
    with torch.no_grad():
        z = net.sample_ray(rays['rays_o'], rays['rays_d'], rays['near'], rays['far'], 0)
    o = net.render_core(rays['rays_o'], rays['rays_d'], z, rays['human_poses'], cos_anneal_ratio=O.get_anneal_val(c, step),
                        step=step, is_train=False)
    out = {'param_checksums': param_checksums(sd), 'R': R, 'seed': seed, 'pseed': pseed, 'step': step, 'z_vals': z}
    out.update({'val_' + k: v for k, v in o.items()})
    np.savez_compressed(os.path.join(GOLD, name + '.npz'), **npy(out))
    print(name, 'saved;', sorted(o.keys()), 'inner', float((o['normal'].abs().sum(-1) > 0).float().mean()),
          'occ_gt mean', float(o['occ_prob_gt'].mean()))


MATERIAL_FIXTURES = {
    'material_bell_p24': ({'shader_cfg': {'human_lights': False, 'diffuse_sample_num': 32, 'specular_sample_num': 16}}, 24, [500, 5000]),
    'material_bear_p16': ({'shader_cfg': {'human_lights': True, 'diffuse_sample_num': 32, 'specular_sample_num': 16}}, 16, [5000]),
    'material_ggx_p16': ({'shader_cfg': {'human_lights': False, 'diffuse_sample_num': 16, 'specular_sample_num': 16,
                                         'geometry_type': 'ggx_smith', 'outer_light_version': 'sphere_direction'}}, 16, [5000]),
}


def make_material_fixture(name, cfg, P, steps, seed=6033, pseed=7):
    """Stage II: the reference's MCShadingNetwork + NeROMaterialRenderer.train_step glue (network/renderer.py:825-848,
    the renderer object is created without its open3d / dataset constructor) on a synthetic mesh traced by the oracle's
    brute-force tracer.  The in-function random draws (field.py:782,805,1070,1074) are recorded by re-seeding."""
    from nero_b200 import params as PR
    ref_shim.install()
    from network.renderer import NeROMaterialRenderer
    scfg = cfg['shader_cfg']
    verts, tris = OM.test_scene(2)
    trace_fn = lambda o, d: OM.renderer_trace(verts, tris, o, d)
    shader = ref_shim.build_reference_mc_shader(dict(scfg), trace_fn, seed=seed)
    sd_ref = {'shader_network.' + k: v.detach().clone() for k, v in shader.state_dict().items()}
    mine = PR.build_material_state_dict(scfg, seed=seed)
    assert list(mine.keys()) == list(sd_ref.keys()) and all(torch.equal(mine[k], sd_ref[k]) for k in sd_ref)
    sd = O.perturb_params(sd_ref, seed=pseed)
    shader.load_state_dict({k[len('shader_network.'):]: v for k, v in sd.items()})
    net = NeROMaterialRenderer.__new__(NeROMaterialRenderer)
    torch.nn.Module.__init__(net)
    net.cfg = {**NeROMaterialRenderer.default_cfg, **cfg}
    net.shader_network = shader
    batch = OM.synthetic_surface_batch(verts, tris, P, seed=seed)
    net.train_batch = {'inters': batch['pts'], 'rays_d': batch['rays_d'], 'normals': batch['normals'], 'rgb': batch['rgb'],
                       'human_poses': batch['human_poses']}
    net.tbn = 10 ** 9
    net.cfg['train_ray_num'] = P
    out = {'param_checksums': param_checksums(sd), 'P': P, 'seed': seed, 'pseed': pseed}
    out.update({'in_' + k: v for k, v in batch.items()})
    names = [n for n, _ in net.named_parameters()]
    c_eps = shader.cfg['change_eps']
    for step in steps:
        net.train_batch_i = 0
        net.zero_grad()
        torch.manual_seed(seed + step)
        o = net.train_step(step)
        loss = sum(torch.mean(v) for k, v in o.items() if k.startswith('loss'))
        loss.backward()
        torch.manual_seed(seed + step)
        pre = f's{step}_'
        out[pre + 'rand_d'], out[pre + 'rand_s'] = torch.rand(P, 1, 1), torch.rand(P, 1, 1)
        out[pre + 'rand_ang'] = torch.rand(P, 1)
        out[pre + 'rand_eps'] = torch.normal(mean=0.0, std=c_eps, size=[P, 1])
        for k, v in o.items():
            out[pre + k] = v
        out[pre + 'loss'] = loss
        gd = dict(net.named_parameters())
        out[pre + 'grad_norms'] = np.array([float(gd[n].grad.double().norm()) if gd[n].grad is not None else 0.0 for n in names])
        for k in ['shader_network.roughness_predictor.6.bias', 'shader_network.feats_network.module0.0.weight_g',
                  'shader_network.albedo_predictor.6.weight_v', 'shader_network.outer_light.6.bias',
                  'shader_network.inner_light.0.weight_g', 'shader_network.human_light.6.bias']:
            if k in gd and gd[k].grad is not None:
                out[pre + 'grad::' + k] = gd[k].grad.clone()
    out['param_names'] = np.array(names)
    np.savez_compressed(os.path.join(GOLD, name + '.npz'), **npy(out))
    print(name, 'saved; loss', {s: float(out[f's{s}_loss']) for s in steps}, 'keys', sorted(o.keys()))


def make_round2_fixtures(seed=6033, pseed=7):
    """Round 2: stage-I sphere_direction lighting (field.py:560-563, 583-586), NeROShapeRenderer.predict_materials
    (renderer.py:629-647 -> field.py:653-657), the SDF field of extract_fields (field.py:1090-1104, val_geometry /
    extract_mesh.py) and the IDE of directions within 1e-3 of the poles (ref_utils.py:85-115)."""
    make_shape_fixture('shape_sphere_r16', {'n_samples': 32, 'n_importance': 32, 'shader_config': {'sphere_direction': True}}, 16,
                       [500, 30000])
    # ---- predict_materials + extract_fields on the perturbed bell parameters
    cfg = {'n_samples': 32, 'n_importance': 32}
    net = ref_shim.build_reference_shape_renderer(cfg, seed=seed)
    sd = O.perturb_params({k: v.detach().clone() for k, v in net.state_dict().items()}, seed=pseed)
    net.load_state_dict(sd)
    g = torch.Generator().manual_seed(21)
    xyz = torch.nn.functional.normalize(torch.randn(300, 3, generator=g), dim=-1) * (0.3 + 0.4 * torch.rand(300, 1, generator=g))
    with torch.no_grad():
        feats = net.sdf_network(xyz)[:, 1:]
        m, r, a = net.color_network.predict_materials(xyz, feats)
        from network.field import extract_fields
        u = extract_fields(-torch.ones(3), torch.ones(3), 24, lambda x: net.sdf_network.sdf(x), batch_size=16)
    out = {'param_checksums': param_checksums(sd), 'seed': seed, 'pseed': pseed, 'xyz': xyz, 'metallic': m, 'roughness': r, 'albedo': a,
           'field24': u}
    np.savez_compressed(os.path.join(GOLD, 'shape_materials_field.npz'), **npy(out))
    print('shape_materials_field saved: metallic', float(m.mean()), 'roughness', float(r.mean()), 'field min', float(u.min()))
    # ---- IDE near the poles: directions within 1e-3 (and 1e-2, 5e-2) of +-z, roughness 0 / 0.3 / 1
    from utils.ref_utils import generate_ide_fn
    ide = generate_ide_fn(5)
    eps = torch.tensor([1e-4, 3e-4, 1e-3, 1e-2, 5e-2])
    az = torch.linspace(0.1, 6.0, 8)
    dirs = []
    for sgn in (1.0, -1.0):
        for e in eps:
            for a_ in az:
                dirs.append(torch.tensor([torch.sin(e) * torch.cos(a_), torch.sin(e) * torch.sin(a_), sgn * torch.cos(e)]))
    dirs = torch.stack(dirs).float()
    kap = torch.tensor([0.0, 0.3, 1.0]).repeat(dirs.shape[0] // 3 + 1)[:dirs.shape[0], None]
    np.savez_compressed(os.path.join(GOLD, 'kat_ide_poles.npz'), **npy({'dirs': dirs, 'kappa': kap, 'ide': ide(dirs, kap)}))
    print('kat_ide_poles saved', dirs.shape)


def main():
    os.makedirs(GOLD, exist_ok=True)
    make_encoding_kats()
    make_shape_fixture('shape_bell_r32', {'n_samples': 32, 'n_importance': 32}, 32, [500, 10000, 30000])
    make_shape_fixture('shape_bear_r24', {'n_samples': 32, 'n_importance': 32, 'shader_config': {'human_light': True}},
                       24, [500, 30000])
    make_shape_fixture('shape_bell_full_r16', {}, 16, [30000])
    make_validation_fixture('shape_val_bell_r32', {'n_samples': 32, 'n_importance': 32}, 32, 30000)
    make_validation_fixture('shape_val_bear_r24', {'n_samples': 32, 'n_importance': 32, 'shader_config': {'human_light': True}}, 24, 30000)
    for k, (cfg, P, steps) in MATERIAL_FIXTURES.items():
        make_material_fixture(k, cfg, P, steps)
    make_round2_fixtures()


if __name__ == '__main__':
    if '--material-only' in sys.argv:
        for k, (cfg, P, steps) in MATERIAL_FIXTURES.items():
            make_material_fixture(k, cfg, P, steps)
    elif '--round2-only' in sys.argv:
        ref_shim.install()
        make_round2_fixtures()
    elif '--val-only' in sys.argv:
        ref_shim.install()
        make_validation_fixture('shape_val_bell_r32', {'n_samples': 32, 'n_importance': 32}, 32, 30000)
        make_validation_fixture('shape_val_bear_r24', {'n_samples': 32, 'n_importance': 32, 'shader_config': {'human_light': True}}, 24, 30000)
    else:
        main()
