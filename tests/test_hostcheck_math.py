"""CPU check of the hand-derived per-sample math (nero_b200/csrc/math_*.cuh compiled for the HOST by nvcc)
against the oracle (values) and the oracle's autograd in float64 (gradients)."""
import ctypes
import os
import subprocess

import numpy as np
import pytest
import torch

import nero_oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, 'hostcheck', 'hostcheck.cu')
SO = os.path.join(HERE, 'hostcheck', 'libhostcheck.so')


@pytest.fixture(scope='module')
def hc():
    deps = [SRC] + [os.path.join(HERE, '..', 'nero_b200', 'csrc', f) for f in ('math_enc.cuh', 'math_shade.cuh', 'common.cuh')]
    if not os.path.exists(SO) or any(os.path.getmtime(d) > os.path.getmtime(SO) for d in deps):
        subprocess.check_call(['/usr/local/cuda/bin/nvcc', '-O2', '-std=c++17', '-Xcompiler', '-fPIC', '-shared', '--fmad=false',
                               '-Wno-deprecated-gpu-targets', '-o', SO, SRC])
    lib = ctypes.CDLL(SO)
    _, mat = O.ide_tables(5)
    lib.hc_set_ide(np.ascontiguousarray(mat, dtype=np.float32).ctypes.data_as(ctypes.c_void_p))
    return lib


def P(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def f32(x):
    return np.ascontiguousarray(x.detach().numpy() if torch.is_tensor(x) else x, dtype=np.float32)


def close(got, want, rtol, atol, name=''):
    want = want.detach().numpy() if torch.is_tensor(want) else want
    err = np.abs(got - want) - (atol + rtol * np.abs(want))
    assert err.max() <= 0, f'{name}: max violation {err.max():.3e} (abs err {np.abs(got - want).max():.3e})'


def test_ide(hc):
    g = torch.Generator().manual_seed(3)
    n = 400
    d = torch.nn.functional.normalize(torch.randn(n, 3, generator=g), dim=-1)
    kap = torch.rand(n, 1, generator=g)
    kap[:50] = 0
    dout = torch.randn(n, 72, generator=g)
    dd = d.double().requires_grad_(True)
    kk = kap.double().requires_grad_(True)
    want = O.ide(dd, kk)
    (want * dout.double()).sum().backward()
    out, gd, gk = np.zeros((n, 72), np.float32), np.zeros((n, 3), np.float32), np.zeros(n, np.float32)
    hc.hc_ide(n, P(f32(d)), P(f32(kap[:, 0])), P(f32(dout)), P(out), P(gd), P(gk))
    close(out, want, 2e-5, 2e-5, 'ide value vs fp64 oracle')
    close(gd, dd.grad, 1e-4, 5e-4, 'ide d/ddir')
    close(gk, kk.grad[:, 0], 1e-4, 5e-4, 'ide d/dkappa')
    # against the fp32 oracle (the reference's own arithmetic) the high bands carry the reference's cancellation error
    w32 = O.ide(d, kap)
    assert np.abs(out - w32.numpy()).max() < 2e-2


def test_pe(hc):
    g = torch.Generator().manual_seed(4)
    n, L = 100, 6
    x = torch.rand(n, 3, generator=g) * 2 - 1
    u = torch.randn(n, 39, generator=g)
    dv = torch.randn(n, 3, generator=g)
    xd = x.double().requires_grad_(True)
    pe = O.embed(xd, L)
    (pe * u.double()).sum().backward()
    jt = torch.autograd.functional.jvp(lambda a: O.embed(a, L), x.double(), dv.double())[1]
    o_pe, o_g, o_t = np.zeros((n, 39), np.float32), np.zeros((n, 3), np.float32), np.zeros((n, 39), np.float32)
    hc.hc_pe(n, L, P(f32(x)), P(f32(u)), P(f32(dv)), P(o_pe), P(o_g), P(o_t))
    close(o_pe, pe, 1e-5, 2e-6, 'pe')
    close(o_g, xd.grad, 1e-4, 1e-4, 'pe backward')
    close(o_t, jt, 1e-4, 1e-4, 'pe tangent')
    x4 = torch.rand(50, 4, generator=g) * 2 - 1
    o4 = np.zeros((50, 84), np.float32)
    hc.hc_pe4(50, 10, P(f32(x4)), P(o4))
    close(o4, O.embed(x4.double(), 10), 1e-4, 1e-4, 'pe10')


def test_sdf_alpha(hc):
    g = torch.Generator().manual_seed(5)
    n = 500
    sdf = torch.randn(n, generator=g) * 0.05
    gr = torch.nn.functional.normalize(torch.randn(n, 3, generator=g), dim=-1) * (1 + 0.2 * torch.randn(n, 1, generator=g))
    dr = torch.nn.functional.normalize(torch.randn(n, 3, generator=g), dim=-1)
    dist = torch.rand(n, generator=g) * 0.05 + 1e-3
    da, dge = torch.randn(n, generator=g), torch.randn(n, generator=g)
    for inv_s, car in [(20.0, 0.0), (80.0, 0.3), (300.0, 1.0)]:
        s, G = sdf.double().requires_grad_(True), gr.double().requires_grad_(True)
        iv = torch.tensor(inv_s, dtype=torch.float64, requires_grad=True)
        tc = (dr.double() * G).sum(-1)
        ic = -(torch.relu(-tc * 0.5 + 0.5) * (1 - car) + torch.relu(-tc) * car)
        pc = torch.sigmoid((s - ic * dist.double() * 0.5) * iv)
        nc = torch.sigmoid((s + ic * dist.double() * 0.5) * iv)
        alpha = ((pc - nc + 1e-5) / (pc + 1e-5)).clip(0, 1)
        ge = (torch.linalg.norm(G, dim=-1) - 1) ** 2
        ((alpha * da.double()).sum() + (ge * dge.double()).sum()).backward()
        oa, oge, ods, odg, odi = [np.zeros(s_, np.float32) for s_ in (n, n, n, (n, 3), n)]
        hc.hc_sdf_alpha(n, P(f32(sdf)), P(f32(gr)), P(f32(dr)), P(f32(dist)), ctypes.c_float(inv_s), ctypes.c_float(car), P(f32(da)),
                        P(f32(dge)), P(oa), P(oge), P(ods), P(odg), P(odi))
        close(oa, alpha, 1e-4, 1e-5, 'alpha')
        close(oge, ge, 1e-4, 1e-6, 'gerr')
        close(ods, s.grad, 2e-3, 2e-3, 'dsdf')
        close(odg, G.grad, 2e-3, 2e-3, 'dg')
        assert abs(odi.sum() - float(iv.grad)) < 2e-3 * (1 + abs(float(iv.grad)))


def test_geometry(hc):
    g = torch.Generator().manual_seed(6)
    n = 300
    G = torch.randn(n, 3, generator=g)
    V = torch.nn.functional.normalize(torch.randn(n, 3, generator=g), dim=-1)
    dn, dr, dnov = torch.randn(n, 3, generator=g), torch.randn(n, 3, generator=g), torch.randn(n, generator=g)
    Gd = G.double().requires_grad_(True)
    nn_ = torch.nn.functional.normalize(Gd, dim=-1)
    vv = V.double()
    nov = (nn_ * vv).sum(-1)
    r = nov[:, None] * nn_ * 2 - vv
    ((nn_ * dn.double()).sum() + (r * dr.double()).sum() + (nov * dnov.double()).sum()).backward()
    on, orr, onov, odg = np.zeros((n, 3), np.float32), np.zeros((n, 3), np.float32), np.zeros(n, np.float32), np.zeros((n, 3), np.float32)
    hc.hc_geometry(n, P(f32(G)), P(f32(V)), P(f32(dn)), P(f32(dr)), P(f32(dnov)), P(on), P(orr), P(onov), P(odg))
    close(on, nn_, 1e-5, 1e-6, 'n')
    close(orr, r, 1e-5, 2e-6, 'r')
    close(odg, Gd.grad, 1e-3, 1e-3, 'dg')


def test_combine(hc):
    g = torch.Generator().manual_seed(7)
    n = 600
    lut = torch.from_numpy(np.fromfile('assets/bsdf_256_256.bin', dtype=np.float32).reshape(256, 256, 2).copy())
    x = torch.rand(n, 20, generator=g)
    x[:, 14] = torch.randn(n, generator=g) * 1.5          # inner weight (occ clamp both sides)
    x[:, 19] = torch.rand(n, generator=g) * 1.4 - 0.2     # NoV beyond [0,1]
    x[:, 5:14] *= 1.5
    dc = torch.randn(n, 3, generator=g)
    xd = x.double().requires_grad_(True)
    m, rough, alb = xd[:, 0:1], xd[:, 1:2], xd[:, 2:5]
    Ld, Ldir, Li, iw, Lh, wh, NoV = xd[:, 5:8], xd[:, 8:11], xd[:, 11:14], xd[:, 14:15], xd[:, 15:18], xd[:, 18:19], xd[:, 19:20]
    occ = torch.clamp(iw * 0.5 + 0.5, 0, 1)
    Ls = Li * occ + (Lh * torch.clamp(wh, 0, 1) + Ldir * (1 - torch.clamp(wh, 0, 1))) * (1 - occ)
    fg = O.fg_lookup(lut.double(), torch.cat([torch.clamp(NoV, 0, 1), torch.clamp(rough, 0, 1)], -1))
    color = torch.clamp(O.linear_to_srgb((1 - m) * alb * Ld + ((0.04 * (1 - m) + m * alb) * fg[:, 0:1] + fg[:, 1:2]) * Ls), 0, 1)
    (color * dc.double()).sum().backward()
    oc, od = np.zeros((n, 3), np.float32), np.zeros((n, 20), np.float32)
    hc.hc_combine(n, P(f32(x)), P(f32(lut)), P(f32(dc)), P(oc), P(od))
    close(oc, color, 1e-5, 2e-6, 'color')
    close(od, xd.grad, 2e-3, 2e-3, 'combine grads')


def test_nerf_post_and_srgb(hc):
    g = torch.Generator().manual_seed(8)
    n = 300
    dens = torch.randn(n, generator=g) * 5
    dens[:5] = 25.0
    dist = torch.rand(n, generator=g) * 0.5
    rgb = torch.randn(n, 3, generator=g) * 3
    da, dc = torch.randn(n, generator=g), torch.randn(n, 3, generator=g)
    D, Rg = dens.double().requires_grad_(True), rgb.double().requires_grad_(True)
    alpha = 1 - torch.exp(-torch.nn.functional.softplus(D) * dist.double())
    col = O.linear_to_srgb(torch.exp(torch.clamp(Rg, max=5.0)))
    ((alpha * da.double()).sum() + (col * dc.double()).sum()).backward()
    oa, oc, odd, odr = np.zeros(n, np.float32), np.zeros((n, 3), np.float32), np.zeros(n, np.float32), np.zeros((n, 3), np.float32)
    hc.hc_nerf_post(n, P(f32(dens)), P(f32(dist)), P(f32(rgb)), P(f32(da)), P(f32(dc)), P(oa), P(oc), P(odd), P(odr))
    close(oa, alpha, 1e-5, 1e-6, 'nerf alpha')
    close(oc, col, 2e-5, 1e-5, 'nerf color')
    close(odd, D.grad, 1e-3, 1e-5, 'ddens')
    close(odr, Rg.grad, 1e-3, 1e-4, 'drgb')


def test_human(hc):
    g = torch.Generator().manual_seed(9)
    n = 400
    rays = O.synthetic_rays(n, seed=12)
    p = torch.randn(n, 3, generator=g) * 0.4
    r = torch.nn.functional.normalize(torch.randn(n, 3, generator=g), dim=-1)
    rough = torch.rand(n, 1, generator=g)
    pose = rays['human_poses']
    dipe = torch.randn(n, 24, generator=g)
    rd, rg = r.double().requires_grad_(True), rough.double().requires_grad_(True)
    inter, dists, hits = O.get_camera_plane_intersection(p.double(), rd, pose.double())
    mean = inter[..., :2] * 0.3
    var = rg * (dists[:, None] * 0.3) ** 2
    hits = hits & (torch.norm(mean, dim=-1) < 1.5) & (dists > 0)
    hf = hits.double().unsqueeze(-1)
    enc = O.ipe(mean * hf, (var * hf).expand(n, 2), 0, 6)
    (enc * dipe.double()).sum().backward()
    oe, oh, odr, odg = np.zeros((n, 24), np.float32), np.zeros(n, np.float32), np.zeros((n, 3), np.float32), np.zeros(n, np.float32)
    hc.hc_human(n, P(f32(p)), P(f32(r)), P(f32(pose.reshape(n, 12))), P(f32(rough[:, 0])), P(f32(dipe)), P(oe), P(oh), P(odr), P(odg))
    assert (oh > 0).sum() > 20, 'too few human-light hits in the test inputs'
    np.testing.assert_array_equal(oh > 0, hits.numpy())
    close(oe, enc, 1e-3, 1e-4, 'ipe')
    close(odr, rd.grad, 5e-3, 5e-3, 'human dr')
    close(odg, rg.grad[:, 0], 5e-3, 5e-3, 'human drough')
