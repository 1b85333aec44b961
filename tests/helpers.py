"""Shared test helpers: seeded parameters (product-side initialisers + oracle.perturb_params), fixtures."""
import os
import warnings

import numpy as np
import torch

import nero_oracle as O
import nero_oracle_mat as OM
from nero_b200 import params as P

warnings.filterwarnings('ignore')
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def load_golden(name):
    return dict(np.load(os.path.join(GOLD, name + '.npz'), allow_pickle=False))


def build_params(cfg, seed=6033, pseed=7, perturb=True):
    sd = P.build_shape_state_dict(cfg, seed=seed)
    return O.perturb_params(sd, seed=pseed) if perturb else sd


def param_checksums(sd):
    keys = sorted(k for k in sd if not k.endswith('FG_LUT'))
    return np.array([[float(sd[k].double().sum()), float(sd[k].double().abs().sum())] for k in keys])


def t(x, dtype=torch.float32):
    return torch.from_numpy(np.asarray(x)).to(dtype)


def rays_from_golden(g, dtype=torch.float32):
    return {k[3:]: t(v, dtype) for k, v in g.items() if k.startswith('in_')}


FIXTURE_CFGS = {
    'shape_bell_r32': {'n_samples': 32, 'n_importance': 32},
    'shape_bear_r24': {'n_samples': 32, 'n_importance': 32, 'shader_config': {'human_light': True}},
    'shape_bell_full_r16': {},
    'shape_sphere_r16': {'n_samples': 32, 'n_importance': 32, 'shader_config': {'sphere_direction': True}},
}
FIXTURE_STEPS = {'shape_bell_r32': [500, 10000, 30000], 'shape_bear_r24': [500, 30000], 'shape_bell_full_r16': [30000],
                 'shape_sphere_r16': [500, 30000]}
VAL_FIXTURES = {'shape_val_bell_r32': FIXTURE_CFGS['shape_bell_r32'], 'shape_val_bear_r24': FIXTURE_CFGS['shape_bear_r24']}
MATERIAL_FIXTURES = {
    'material_bell_p24': ({'shader_cfg': {'human_lights': False, 'diffuse_sample_num': 32, 'specular_sample_num': 16}}, [500, 5000]),
    'material_bear_p16': ({'shader_cfg': {'human_lights': True, 'diffuse_sample_num': 32, 'specular_sample_num': 16}}, [5000]),
    'material_ggx_p16': ({'shader_cfg': {'human_lights': False, 'diffuse_sample_num': 16, 'specular_sample_num': 16,
                                         'geometry_type': 'ggx_smith', 'outer_light_version': 'sphere_direction'}}, [5000]),
}


def build_material_params(shader_cfg, seed=6033, pseed=7):
    return O.perturb_params(P.build_material_state_dict(shader_cfg, seed=seed), seed=pseed)


def material_batch_from_golden(g, dtype=torch.float32):
    return {k[3:]: t(v, dtype) for k, v in g.items() if k.startswith('in_')}


def material_rands(g, step, dtype=torch.float32):
    return {k: t(g[f's{step}_{k}'], dtype) for k in ('rand_d', 'rand_s', 'rand_ang', 'rand_eps')}


def act_ref(x, act, p=0.0):
    """fp64/fp32 torch reference of the GEMM epilogue activations (codes of include/nero_b200.h)."""
    if act == 1:
        return torch.nn.functional.softplus(x, beta=100)
    if act == 2:
        return torch.relu(x)
    if act == 3:
        return torch.sigmoid(x)
    if act == 4:
        return torch.exp(torch.clamp(x, max=p))
    return x


def dact_ref(h, dact):
    """derivative factor recovered from the stored post-activation value"""
    if dact == 1:
        return torch.where(100 * h > 20, torch.ones_like(h), -torch.expm1(-100 * h))
    if dact == 2:
        return (h > 0).to(h.dtype)
    return torch.ones_like(h)
