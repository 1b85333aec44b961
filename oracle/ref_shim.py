"""Import shim that lets the UNMODIFIED reference (/root/reference) run on CPU in the build container.

TEST INFRASTRUCTURE ONLY.  Used by oracle/make_golden.py to generate the committed fixtures under
tests/golden/ and by tests that pin oracle/nero_oracle.py against the real reference when
/root/reference is present.  Nothing on the product path (nero_b200/) imports this file, and nothing
run on the GPU box needs it (the reference does not travel there).

What it does:
  * numpy.math / numpy.bool aliases removed in numpy 2 (utils/ref_utils.py:10, dataset/database.py:225)
  * identity .cuda() on tensors/modules, device-less torch.randperm (renderer.py:537)
  * stub modules for the un-vendored third-party imports (nvdiffrast, raytracing, open3d, mcubes, ...)
    with nvdiffrast.torch.texture restated as grid_sample(bilinear, border, align_corners=False),
    the documented semantics of dr.texture(filter_mode='linear', boundary_mode='clamp') (field.py:612).
"""
import math
import os
import sys
import types

import numpy as np
import torch
import torch.nn.functional as F

REFERENCE_ROOT = os.environ.get('NERO_REFERENCE_ROOT', '/root/reference')


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, 'network'))


def _texture(tex, uv, filter_mode='linear', boundary_mode='clamp'):
    # tex [1,H,W,C], uv [1,h,w,2] in [0,1]; u -> width axis, v -> height axis
    assert filter_mode == 'linear' and boundary_mode == 'clamp'
    out = F.grid_sample(tex.permute(0, 3, 1, 2), uv * 2.0 - 1.0, mode='bilinear', padding_mode='border',
                        align_corners=False)
    return out.permute(0, 2, 3, 1)


_installed = False


def install():
    """Patch the process so `import network.renderer` from the reference works on CPU."""
    global _installed
    if _installed:
        return
    if not reference_available():
        raise RuntimeError(f'reference not found at {REFERENCE_ROOT}')
    np.math = math
    if not hasattr(np, 'bool'):
        np.bool = bool
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.nn.Module.cuda = lambda self, *a, **k: self
    _randperm = torch.randperm

    def randperm(n, *a, **k):
        k.pop('device', None)
        return _randperm(n, *a, **k)
    torch.randperm = randperm

    def stub(name, **attrs):
        m = types.ModuleType(name)
        for k, v in attrs.items():
            setattr(m, k, v)
        sys.modules[name] = m
        return m

    dr = stub('nvdiffrast.torch', texture=_texture)
    stub('nvdiffrast', torch=dr)
    for name in ['mcubes', 'h5py', 'open3d', 'trimesh', 'raytracing', 'xatlas', 'plyfile', 'transforms3d',
                 'transforms3d.axangles', 'transforms3d.euler', 'transforms3d.quaternions', 'skimage',
                 'skimage.io', 'skimage.metrics', 'tensorboardX', 'matplotlib', 'matplotlib.pyplot']:
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                stub(name)
    dummy = lambda *a, **k: None
    for mod, names in {
        'plyfile': ['PlyData', 'PlyElement'],
        'tensorboardX': ['SummaryWriter'],
        'transforms3d.axangles': ['mat2axangle', 'axangle2mat'],
        'transforms3d.euler': ['mat2euler', 'euler2mat'],
        'transforms3d.quaternions': ['mat2quat', 'quat2mat', 'qmult', 'qinverse'],
        'skimage.io': ['imread', 'imsave'],
        'skimage.metrics': ['structural_similarity'],
    }.items():
        m = sys.modules[mod]
        for n in names:
            if not hasattr(m, n):
                setattr(m, n, dummy)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    _installed = True


class _cwd:
    def __init__(self, path):
        self.path = path

    def __enter__(self):
        self.old = os.getcwd()
        os.chdir(self.path)

    def __exit__(self, *a):
        os.chdir(self.old)


def build_reference_shape_renderer(cfg=None, seed=6033):
    """NeROShapeRenderer(cfg, training=False) from the reference with its own initialisers under `seed`
    (network/renderer.py:113-134; Trainer seed train/trainer.py:36)."""
    install()
    with _cwd(REFERENCE_ROOT):  # assets/bsdf_256_256.bin is CWD-relative (field.py:510)
        from network.renderer import NeROShapeRenderer
        torch.manual_seed(seed)
        net = NeROShapeRenderer(cfg or {}, training=False)
    return net


def build_reference_mc_shader(shader_cfg, trace_fn, seed=6033):
    """The reference's MCShadingNetwork(cfg, ray_trace_fun) (network/field.py:713-751) under `seed`; `trace_fn` stands in
    for NeROMaterialRenderer.trace (the third-party ray tracer is not in the container)."""
    install()
    with _cwd(REFERENCE_ROOT):
        from network.field import MCShadingNetwork
        torch.manual_seed(seed)
        net = MCShadingNetwork(shader_cfg, trace_fn)
    return net
