"""CPU: the product-side workload generators (nero_b200/synthetic.py, used by bench.py and tools/) produce exactly the inputs
the oracle's generators produce for the fixtures -- so the benchmarks measure the workloads the parity tests pin while the
measured arm imports nothing from oracle/."""
import numpy as np
import torch

import nero_oracle as O
import nero_oracle_mat as OM
from nero_b200 import params as P, synthetic as S


def test_generators_match_the_oracle_generators():
    a, b = S.synthetic_rays(257, seed=6033), O.synthetic_rays(257, seed=6033)
    assert set(a) == set(b) and all(torch.equal(a[k], b[k]) for k in a)
    for build, cfg in ((P.build_shape_state_dict, {'shader_config': {'human_light': True}}),
                       (P.build_material_state_dict, {'human_lights': True})):
        sd = build(cfg, seed=6033)
        x, y = S.perturb_params(sd), O.perturb_params(sd)
        assert list(x) == list(y) and all(torch.equal(x[k], y[k]) for k in x)
    for sub in (1, 3):
        (v0, f0), (v1, f1) = S.test_scene(sub), OM.test_scene(sub)
        assert np.array_equal(v0, v1) and np.array_equal(f0, f1)


def test_measured_arms_do_not_import_the_oracle():
    import ast
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = open(os.path.join(root, 'bench.py')).read()
    tree = ast.parse(src)
    fns = {n.name: n for n in ast.walk(tree) if isinstance(n, (ast.FunctionDef, ast.ClassDef))}
    for name in ('run_ours', 'build_net', 'synthetic_dataset', 'profile_chains', 'training_loss', 'Workload', 'timed'):
        body = ast.get_source_segment(src, fns[name])
        assert 'nero_oracle' not in body, f'bench.py:{name} must not use the oracle'
    assert 'nero_oracle' in ast.get_source_segment(src, fns['cpu_baseline'])     # the one leg that times it
