// CPU harness (TEST INFRASTRUCTURE): runs the host+device math functions of nero_b200/csrc/math_*.cuh on the HOST
// so the hand-derived backward formulas are checked against the oracle's autograd without a GPU.
// Built by tests/test_hostcheck_math.py with nvcc; never part of libnero_b200.so.
#include "../../nero_b200/csrc/math_shade.cuh"
#include "../../nero_b200/csrc/math_mc.cuh"
using namespace nero;

static IdeTable g_tab;
extern "C" {
void hc_set_ide(const float* mat) {
  int i = 0;
  for (int e = 0; e < 5; ++e) { int l = 1 << e; for (int m = 0; m <= l; ++m) { g_tab.m[i] = m; g_tab.l[i] = l; ++i; } }
  for (int k = 0; k < 17; ++k) for (int j = 0; j < 36; ++j) { g_tab.mat[k][j] = double(mat[k * 36 + j]); g_tab.matf[k][j] = mat[k * 36 + j]; }
}
void hc_ide(int n, const float* d, const float* kap, const float* dout, float* out, float* dd, float* dk) {
  for (int i = 0; i < n; ++i) {
    ide_forward(g_tab, d + 3 * i, kap[i], out + 72 * i);
    dd[3 * i] = dd[3 * i + 1] = dd[3 * i + 2] = 0.f;
    dk[i] = ide_backward(g_tab, d + 3 * i, kap[i], dout + 72 * i, dd + 3 * i);
  }
}
void hc_pe(int n, int L, const float* x, const float* u, const float* dvec, float* pe, float* g, float* t) {
  int w = 3 + 6 * L;
  for (int i = 0; i < n; ++i) {
    pe_encode<3>(x + 3 * i, L, pe + w * i);
    pe_backward<3>(pe + w * i, L, u + w * i, g + 3 * i);
    pe_tangent<3>(pe + w * i, L, dvec + 3 * i, t + w * i);
  }
}
void hc_pe4(int n, int L, const float* x, float* pe) { for (int i = 0; i < n; ++i) pe_encode<4>(x + 4 * i, L, pe + (4 + 8 * L) * i); }
void hc_sdf_alpha(int n, const float* sdf, const float* g, const float* dir, const float* dist, float inv_s, float car,
                  const float* dalpha, const float* dgerr, float* alpha, float* gerr, float* dsdf, float* dg, float* dinvs) {
  for (int i = 0; i < n; ++i) {
    SdfAlphaOut o = sdf_alpha_fwd(sdf[i], g + 3 * i, dir + 3 * i, dist[i], inv_s, car);
    alpha[i] = o.alpha; gerr[i] = o.grad_err;
    dg[3 * i] = dg[3 * i + 1] = dg[3 * i + 2] = 0.f;
    dinvs[i] = sdf_alpha_bwd(sdf[i], g + 3 * i, dir + 3 * i, dist[i], inv_s, car, dalpha[i], dgerr[i], dsdf + i, dg + 3 * i);
  }
}
void hc_geometry(int n, const float* g, const float* view, const float* dn, const float* dr, const float* dNoV, float* nrm, float* r,
                 float* NoV, float* dg) {
  for (int i = 0; i < n; ++i) {
    float v[3];
    shade_geometry_fwd(g + 3 * i, view + 3 * i, nrm + 3 * i, v, r + 3 * i, NoV + i);
    dg[3 * i] = dg[3 * i + 1] = dg[3 * i + 2] = 0.f;
    shade_geometry_bwd(g + 3 * i, nrm + 3 * i, v, NoV[i], dn + 3 * i, dr + 3 * i, dNoV[i], dg + 3 * i);
  }
}
// in: [n,20] = metallic, roughness, albedo3, Ld3, Ldir3, Li3, iw, Lh3, wh, NoV ; grads same layout
void hc_combine(int n, const float* in, const float* lut, const float* dcolor, float* color, float* din) {
  for (int i = 0; i < n; ++i) {
    const float* a = in + 20 * i;
    ShadeIn s; s.metallic = a[0]; s.roughness = a[1];
    for (int c = 0; c < 3; ++c) { s.albedo[c] = a[2 + c]; s.Ld[c] = a[5 + c]; s.Ldir[c] = a[8 + c]; s.Li[c] = a[11 + c]; s.Lh[c] = a[15 + c]; }
    s.iw = a[14]; s.wh = a[18]; s.NoV = a[19];
    shade_combine_fwd(s, lut, color + 3 * i);
    ShadeGrad d; shade_combine_bwd(s, lut, dcolor + 3 * i, d);
    float* o = din + 20 * i;
    o[0] = d.metallic; o[1] = d.roughness;
    for (int c = 0; c < 3; ++c) { o[2 + c] = d.albedo[c]; o[5 + c] = d.Ld[c]; o[8 + c] = d.Ldir[c]; o[11 + c] = d.Li[c]; o[15 + c] = d.Lh[c]; }
    o[14] = d.iw; o[18] = d.wh; o[19] = d.NoV;
  }
}
void hc_nerf_post(int n, const float* dens, const float* dist, const float* rgb, const float* dalpha, const float* dcolor, float* alpha,
                  float* color, float* ddens, float* drgb) {
  for (int i = 0; i < n; ++i) {
    nerf_post_fwd(dens[i], dist[i], rgb + 3 * i, alpha + i, color + 3 * i);
    nerf_post_bwd(dens[i], dist[i], rgb + 3 * i, dalpha[i], dcolor + 3 * i, ddens + i, drgb + 3 * i);
  }
}
void hc_human(int n, const float* p, const float* r, const float* pose, const float* rough, const float* dipe, float* ipe_out, float* hit,
              float* dr, float* drough) {
  for (int i = 0; i < n; ++i) {
    HumanGeo h = human_geo_fwd(p + 3 * i, r + 3 * i, pose + 12 * i, rough[i]);
    float var2[2] = {h.var, h.var};
    ipe_forward(h.mean, var2, ipe_out + 24 * i);
    hit[i] = h.hit;
    dr[3 * i] = dr[3 * i + 1] = dr[3 * i + 2] = 0.f; drough[i] = 0.f;
    if (h.hit > 0.f) {
      float dmean[2], dvar[2];
      ipe_backward(h.mean, var2, dipe + 24 * i, dmean, dvar);
      drough[i] = human_geo_bwd(p + 3 * i, r + 3 * i, pose + 12 * i, rough[i], dmean, dvar[0] + dvar[1], dr + 3 * i);
    }
  }
}
// sphere_direction: s = unit-sphere exit direction of the ray (p, d); gd = (ds/dd)^T gs
void hc_sphere_dir(int n, const float* p, const float* d, const float* gs, float* s, float* gd) {
  for (int i = 0; i < n; ++i) {
    sphere_dir_fwd(p + 3 * i, d + 3 * i, s + 3 * i);
    gd[3 * i] = gd[3 * i + 1] = gd[3 * i + 2] = 0.f;
    sphere_dir_bwd(p + 3 * i, d + 3 * i, gs + 3 * i, gd + 3 * i);
  }
}
void hc_srgb(int n, const float* x, float* y, float* dy) { for (int i = 0; i < n; ++i) { y[i] = linear_to_srgb(x[i]); dy[i] = dlinear_to_srgb(x[i]); } }
// stage II: sampled direction, specular weight and Schlick factor of one (point, sample) with forward-mode d/d(roughness)
void hc_mc(int n, const float* normal, const float* view, const float* a, const float* az01, const float* el, const int* spec, int ggx,
           float frac_d, float frac_s, float* dir, float* w, float* f5, float* ddir, float* dw, float* df5) {
  for (int i = 0; i < n; ++i) {
    const McPoint q = mc_point(normal + 3 * i, view + 3 * i);
    const Dual A = mk(a[i], 1.f);
    Dual d[3];
    const float ang = mc_azimuth(az01[i], 0.f, false);
    if (spec[i]) mc_specular_dir<Dual>(q, ang, el[i], A, d);
    else { float df[3]; mc_diffuse_dir(q, ang, el[i], df); for (int k = 0; k < 3; ++k) d[k] = mk(df[k]); }
    Dual W, F;
    mc_weights<Dual>(q, d, A, spec[i] != 0, frac_d, frac_s, ggx, W, F);
    for (int k = 0; k < 3; ++k) { dir[3 * i + k] = d[k].v; ddir[3 * i + k] = d[k].d; }
    w[i] = W.v; dw[i] = W.d; f5[i] = F.v; df5[i] = F.d;
  }
}
}
