"""Parameter containers with the reference's state_dict layout and initialisers.

These modules only OWN parameters (names, shapes, order and initial values identical to the reference so
checkpoints and optimizer state are interchangeable -- SURVEY.md Appendix B); all arithmetic happens in the
CUDA kernels driven from nero_b200/shape_renderer.py.

Reference: network/field.py:60-128 (SDFNetwork ctor, geometric init), :184-188 (SingleVarianceNetwork),
:205-256 (NeRFNetwork ctor), :310-346 (make_predictor), :496-533 (AppShadingNetwork ctor).
The RNG draw ORDER of the reference constructors is reproduced (one nn.Linear at a time, then the
overriding initialisers) so that torch.manual_seed(s) yields bit-identical parameters.
"""
import os

import numpy as np
import torch
import torch.nn as nn

ASSET_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'assets')


class WNLinear(nn.Module):
    """A weight-normalised linear layer's parameters: bias, weight_g [out,1], weight_v [out,in]
    (registration order = torch.nn.utils.weight_norm applied to nn.Linear: bias first)."""

    def __init__(self, lin: nn.Linear):
        super().__init__()
        w = lin.weight.detach()
        self.bias = nn.Parameter(lin.bias.detach().clone())
        self.weight_g = nn.Parameter(torch.linalg.norm(w, dim=1, keepdim=True).clone())
        self.weight_v = nn.Parameter(w.clone())

    @property
    def in_features(self):
        return self.weight_v.shape[1]

    @property
    def out_features(self):
        return self.weight_v.shape[0]


class PlainLinear(nn.Module):
    def __init__(self, lin: nn.Linear):
        super().__init__()
        self.weight = nn.Parameter(lin.weight.detach().clone())
        self.bias = nn.Parameter(lin.bias.detach().clone())


class SDFParams(nn.Module):
    """lin0..lin8 of the 8-hidden-layer softplus SDF MLP (d_hidden 256, PE6, skip at 4, d_out 257)."""

    def __init__(self, d_out=257, d_in=3, d_hidden=256, n_layers=8, skip_in=(4,), multires=6, bias=0.5,
                 geometric_init=True):
        super().__init__()
        dims = [d_in] + [d_hidden] * n_layers + [d_out]
        input_ch = d_in * (1 + 2 * multires) if multires > 0 else d_in
        dims[0] = input_ch
        self.dims, self.skip_in, self.multires = dims, tuple(skip_in), multires
        self.num_layers = len(dims)
        for l in range(self.num_layers - 1):
            out_dim = dims[l + 1] - dims[0] if (l + 1) in self.skip_in else dims[l + 1]
            lin = nn.Linear(dims[l], out_dim)
            if geometric_init:
                with torch.no_grad():
                    if l == self.num_layers - 2:
                        nn.init.normal_(lin.weight, mean=np.sqrt(np.pi) / np.sqrt(dims[l]), std=0.0001)
                        nn.init.constant_(lin.bias, -bias)
                    elif multires > 0 and l == 0:
                        nn.init.constant_(lin.bias, 0.0)
                        nn.init.constant_(lin.weight[:, 3:], 0.0)
                        nn.init.normal_(lin.weight[:, :3], 0.0, np.sqrt(2) / np.sqrt(out_dim))
                    elif multires > 0 and l in self.skip_in:
                        nn.init.constant_(lin.bias, 0.0)
                        nn.init.normal_(lin.weight, 0.0, np.sqrt(2) / np.sqrt(out_dim))
                        nn.init.constant_(lin.weight[:, -(dims[0] - 3):], 0.0)
                    else:
                        nn.init.constant_(lin.bias, 0.0)
                        nn.init.normal_(lin.weight, 0.0, np.sqrt(2) / np.sqrt(out_dim))
            setattr(self, f'lin{l}', WNLinear(lin))

    def layers(self):
        return [getattr(self, f'lin{l}') for l in range(self.num_layers - 1)]


class VarianceParams(nn.Module):
    def __init__(self, init_val=0.3):
        super().__init__()
        self.variance = nn.Parameter(torch.tensor(init_val))


class NeRFParams(nn.Module):
    """Outer NeRF++ (D=8, W=256, PE10 on 4-d, PE4 on view, skip after layer 4)."""

    def __init__(self, D=8, W=256, d_in=4, d_in_view=3, multires=10, multires_view=4, skips=(4,)):
        super().__init__()
        self.input_ch = d_in * (1 + 2 * multires)
        self.input_ch_view = d_in_view * (1 + 2 * multires_view)
        self.skips = tuple(skips)
        lins = [nn.Linear(self.input_ch, W)] + \
               [nn.Linear(W, W) if i not in self.skips else nn.Linear(W + self.input_ch, W) for i in range(D - 1)]
        self.pts_linears = nn.ModuleList([PlainLinear(l) for l in lins])
        self.views_linears = nn.ModuleList([PlainLinear(nn.Linear(self.input_ch_view + W, W // 2))])
        self.feature_linear = PlainLinear(nn.Linear(W, W))
        self.alpha_linear = PlainLinear(nn.Linear(W, 1))
        self.rgb_linear = PlainLinear(nn.Linear(W // 2, 3))


class PredictorParams(nn.Module):
    """make_predictor: Sequential(WN Linear, ReLU, WN Linear, ReLU, WN Linear, ReLU, WN Linear, act) --
    parameters live at indices 0, 2, 4, 6 like the reference's nn.Sequential."""

    def __init__(self, feats_dim, output_dim, run_dim=256):
        super().__init__()
        dims = [(feats_dim, run_dim), (run_dim, run_dim), (run_dim, run_dim), (run_dim, output_dim)]
        for i, (a, b) in zip((0, 2, 4, 6), dims):
            self.add_module(str(i), WNLinear(nn.Linear(a, b)))

    def layers(self):
        return [getattr(self, str(i)) for i in (0, 2, 4, 6)]

    def set_last_bias(self, value):
        with torch.no_grad():
            getattr(self, '6').bias.fill_(float(value))


class ShadingParams(nn.Module):
    """AppShadingNetwork parameters + FG_LUT buffer (network/field.py:486-533)."""
    default_cfg = {
        'human_light': False, 'sphere_direction': False, 'light_pos_freq': 8, 'inner_init': -0.95,
        'roughness_init': 0.0, 'metallic_init': 0.0, 'light_exp_max': 0.0,
    }

    def __init__(self, cfg):
        super().__init__()
        self.cfg = {**self.default_cfg, **cfg}
        feats_dim = 256
        self.metallic_predictor = PredictorParams(feats_dim + 3, 1)
        if self.cfg['metallic_init'] != 0:
            self.metallic_predictor.set_last_bias(self.cfg['metallic_init'])
        self.roughness_predictor = PredictorParams(feats_dim + 3, 1)
        if self.cfg['roughness_init'] != 0:
            self.roughness_predictor.set_last_bias(self.cfg['roughness_init'])
        self.albedo_predictor = PredictorParams(feats_dim + 3, 3)
        lut_path = 'assets/bsdf_256_256.bin' if os.path.exists('assets/bsdf_256_256.bin') \
            else os.path.join(ASSET_DIR, 'bsdf_256_256.bin')
        lut = torch.from_numpy(np.fromfile(lut_path, dtype=np.float32).reshape(1, 256, 256, 2).copy())
        self.register_buffer('FG_LUT', lut)
        pos_dim = 3 * (1 + 2 * self.cfg['light_pos_freq'])
        dir_dim = 3 * (1 + 2 * 6)
        self.outer_light = PredictorParams(72 * 2 if self.cfg['sphere_direction'] else 72, 3)
        self.outer_light.set_last_bias(np.log(0.5))
        self.inner_light = PredictorParams(pos_dim + 72, 3)
        self.inner_light.set_last_bias(np.log(0.5))
        self.inner_weight = PredictorParams(pos_dim + dir_dim, 1)
        self.inner_weight.set_last_bias(self.cfg['inner_init'])
        if self.cfg['human_light']:
            self.human_light_predictor = PredictorParams(2 * 2 * 6, 4)
            self.human_light_predictor.set_last_bias(np.log(0.01))


class MaterialFeatsParams(nn.Module):
    """MaterialFeatsNetwork (network/field.py:660-683): module0 = 4x(WN Linear, ReLU) on PE8(x)=51; module1 = 4 WN linears on
    cat(256, 51), ReLU between, none after the last.  Parameters at Sequential indices 0, 2, 4, 6."""

    def __init__(self, run_dim=256, input_dim=51):
        super().__init__()
        self.module0, self.module1 = nn.Module(), nn.Module()
        for i, a in zip((0, 2, 4, 6), (input_dim, run_dim, run_dim, run_dim)):
            self.module0.add_module(str(i), WNLinear(nn.Linear(a, run_dim)))
        for i, a in zip((0, 2, 4, 6), (input_dim + run_dim, run_dim, run_dim, run_dim)):
            self.module1.add_module(str(i), WNLinear(nn.Linear(a, run_dim)))

    def layers(self):
        return [getattr(self.module0, str(i)) for i in (0, 2, 4, 6)] + [getattr(self.module1, str(i)) for i in (0, 2, 4, 6)]


def _sample_sphere(n):
    """utils/base_utils.py:800-813 with begin_elevation=0 (upper-hemisphere golden-ratio spiral)."""
    total = int(n // 0.5)
    idx = np.arange(total - n, total)
    return 2 * np.pi * idx * ((np.sqrt(5) - 1.0) / 2.) % (2 * np.pi), np.arcsin(2. * idx / total - 1.)


class MCShadingParams(nn.Module):
    """MCShadingNetwork parameters, `light_pts` buffer and the fixed (az, el) sample tables (network/field.py:694-751)."""
    default_cfg = {
        'diffuse_sample_num': 512, 'specular_sample_num': 256, 'human_lights': True, 'light_exp_max': 5.0,
        'inner_light_exp_max': 5.0, 'outer_light_version': 'direction', 'geometry_type': 'schlick',
        'reg_change': True, 'change_eps': 0.05, 'change_type': 'gaussian', 'reg_lambda1': 0.005, 'reg_min_max': True,
        'random_azimuth': True, 'is_real': False,
    }

    def __init__(self, cfg):
        super().__init__()
        self.cfg = {**self.default_cfg, **cfg}
        self.feats_network = MaterialFeatsParams()
        self.metallic_predictor = PredictorParams(256 + 3, 1)
        self.roughness_predictor = PredictorParams(256 + 3, 1)
        self.albedo_predictor = PredictorParams(256 + 3, 3)
        ver = self.cfg['outer_light_version']
        if ver not in ('direction', 'sphere_direction'):
            raise NotImplementedError(ver)
        self.outer_light = PredictorParams(72 if ver == 'direction' else 144, 3)
        self.outer_light.set_last_bias(np.log(0.5))
        if self.cfg['human_lights']:
            self.human_light = PredictorParams(2 * 2 * 6, 4)
            self.human_light.set_last_bias(np.log(0.02))
        self.inner_light = PredictorParams(51 + 72, 3)
        self.inner_light.set_last_bias(np.log(0.5))

        def table(n):
            az, el = _sample_sphere(n)
            return torch.from_numpy(np.stack([az * 0.5 / np.pi, 1 - 2 * el / np.pi], -1).astype(np.float32))
        # plain attributes in the reference (not in the state_dict), field.py:737-745
        self.diffuse_direction_samples = table(self.cfg['diffuse_sample_num'])
        self.specular_direction_samples = table(self.cfg['specular_sample_num'])
        az, el = _sample_sphere(8192)
        pts = np.stack([np.cos(az) * np.cos(el), np.sin(az) * np.cos(el), np.sin(el)], -1)
        self.register_buffer('light_pts', torch.from_numpy(pts.astype(np.float32)))


def build_material_state_dict(shader_cfg, seed=6033):
    """state_dict of a freshly initialised stage-II renderer (keys 'shader_network.*') under torch.manual_seed(seed)."""
    torch.manual_seed(seed)
    m = MCShadingParams(shader_cfg)
    return {'shader_network.' + k: v.detach().clone() for k, v in m.state_dict().items()}


class ShapeParams(nn.Module):
    """All stage-I parameters in the reference's module order (network/renderer.py:117-130)."""

    def __init__(self, cfg):
        super().__init__()
        self.sdf_network = SDFParams(d_out=cfg.get('sdf_d_out', 257), d_in=3, d_hidden=256,
                                     n_layers=cfg.get('sdf_n_layers', 8), skip_in=(cfg.get('sdf_n_layers', 8) // 2,),
                                     multires=cfg.get('sdf_freq', 6), bias=cfg.get('sdf_bias', 0.5),
                                     geometric_init=cfg.get('geometry_init', True))
        self.deviation_network = VarianceParams(cfg.get('inv_s_init', 0.3))
        self.outer_nerf = NeRFParams()
        with torch.no_grad():
            self.outer_nerf.rgb_linear.bias.fill_(float(np.log(0.5)))
        self.color_network = ShadingParams(cfg.get('shader_config', {}))


def build_shape_state_dict(cfg, seed=6033):
    """state_dict of freshly initialised stage-I parameters under torch.manual_seed(seed) (CPU tensors)."""
    torch.manual_seed(seed)
    m = ShapeParams(cfg)
    return {k: v.detach().clone() for k, v in m.state_dict().items()}
