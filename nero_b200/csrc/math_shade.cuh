// Per-sample shading / opacity math (forward + hand-derived backward), host+device (see math_enc.cuh).
#pragma once
#include "math_enc.cuh"

namespace nero {

NERO_HD float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }
NERO_HD float clamp01(float x) { return fminf(fmaxf(x, 0.0f), 1.0f); }
NERO_HD float in01(float x) { return (x >= 0.0f && x <= 1.0f) ? 1.0f : 0.0f; }

// ------------------------------------------------------------------ NeuS SDF -> alpha   (network/renderer.py:497-511, :574)
struct SdfAlphaOut { float alpha, grad_err; };
NERO_HD SdfAlphaOut sdf_alpha_fwd(float sdf, const float* g, const float* dir, float dist, float inv_s, float car) {
  const float tc = dir[0] * g[0] + dir[1] * g[1] + dir[2] * g[2];
  const float ic = -(fmaxf(-tc * 0.5f + 0.5f, 0.f) * (1.0f - car) + fmaxf(-tc, 0.f) * car);
  const float e_next = sdf + ic * dist * 0.5f, e_prev = sdf - ic * dist * 0.5f;
  const float pc = sigmoidf_(e_prev * inv_s), nc = sigmoidf_(e_next * inv_s);
  const float a = (pc - nc + 1e-5f) / (pc + 1e-5f);
  const float gn = sqrtf(g[0] * g[0] + g[1] * g[1] + g[2] * g[2]);
  SdfAlphaOut o;
  o.alpha = clamp01(a);
  o.grad_err = (gn - 1.0f) * (gn - 1.0f);
  return o;
}
// returns d inv_s; writes dsdf, accumulates into dg[3]
NERO_HD float sdf_alpha_bwd(float sdf, const float* g, const float* dir, float dist, float inv_s, float car, float dalpha,
                            float dgrad_err, float* dsdf, float* dg) {
  const float tc = dir[0] * g[0] + dir[1] * g[1] + dir[2] * g[2];
  const float ic = -(fmaxf(-tc * 0.5f + 0.5f, 0.f) * (1.0f - car) + fmaxf(-tc, 0.f) * car);
  const float e_next = sdf + ic * dist * 0.5f, e_prev = sdf - ic * dist * 0.5f;
  const float pc = sigmoidf_(e_prev * inv_s), nc = sigmoidf_(e_next * inv_s);
  const float num = pc - nc + 1e-5f, den = pc + 1e-5f;
  const float a = num / den;
  const float da = dalpha * in01(a);
  const float dnum = da / den, dden = -da * num / (den * den);
  const float tp = (dnum + dden) * pc * (1.0f - pc), tn = -dnum * nc * (1.0f - nc);
  const float de_prev = tp * inv_s, de_next = tn * inv_s;
  *dsdf = de_prev + de_next;
  const float dic = (de_next - de_prev) * dist * 0.5f;
  const float dtc = dic * (0.5f * (1.0f - car) * ((-tc * 0.5f + 0.5f) > 0.f ? 1.f : 0.f) + car * (-tc > 0.f ? 1.f : 0.f));
  const float gn = sqrtf(g[0] * g[0] + g[1] * g[1] + g[2] * g[2]);
  const float ge = gn > 0.f ? dgrad_err * 2.0f * (gn - 1.0f) / gn : 0.f;
  for (int c = 0; c < 3; ++c) dg[c] += dtc * dir[c] + ge * g[c];
  return tp * e_prev + tn * e_next;
}

// ------------------------------------------------------------------ shading geometry   (network/field.py:592-595)
// n = normalize(g), v = normalize(view), r = 2 (v.n) n - v, NoV = n.v
NERO_HD void shade_geometry_fwd(const float* g, const float* view, float* n, float* v, float* r, float* NoV) {
  const float gn = fmaxf(sqrtf(g[0] * g[0] + g[1] * g[1] + g[2] * g[2]), 1e-12f);
  const float vn = fmaxf(sqrtf(view[0] * view[0] + view[1] * view[1] + view[2] * view[2]), 1e-12f);
  for (int c = 0; c < 3; ++c) { n[c] = g[c] / gn; v[c] = view[c] / vn; }
  const float d = n[0] * v[0] + n[1] * v[1] + n[2] * v[2];
  for (int c = 0; c < 3; ++c) r[c] = 2.0f * d * n[c] - v[c];
  *NoV = d;
}
// given dn (direct), dr, dNoV -> accumulates dg
NERO_HD void shade_geometry_bwd(const float* g, const float* n, const float* v, float NoV, const float* dn_in, const float* dr,
                                float dNoV, float* dg) {
  // r = 2 (v.n) n - v  ->  dn += 2 [(dr.n) v + (v.n) dr];  NoV = n.v -> dn += dNoV v
  const float drn = dr[0] * n[0] + dr[1] * n[1] + dr[2] * n[2];
  float dn[3];
  for (int c = 0; c < 3; ++c) dn[c] = dn_in[c] + 2.0f * (drn * v[c] + NoV * dr[c]) + dNoV * v[c];
  const float gn = sqrtf(g[0] * g[0] + g[1] * g[1] + g[2] * g[2]);
  if (gn < 1e-12f) { for (int c = 0; c < 3; ++c) dg[c] += dn[c] / 1e-12f; return; }
  const float dnn = dn[0] * n[0] + dn[1] * n[1] + dn[2] * n[2];
  for (int c = 0; c < 3; ++c) dg[c] += (dn[c] - dnn * n[c]) / gn;
}

// ------------------------------------------------------------------ sphere_direction light encoding (network/field.py:380-396, 560-563, 583-586)
// s = normalize(p' + d t): the unit-sphere exit point of the ray (p', d), p' = p pulled back to radius 0.999 when it lies
// outside (offset_points_to_sphere), t = -p'.d + sqrt((p'.d)^2 - |p'|^2 + 1 + 1e-6) (get_sphere_intersection).
NERO_HD void sphere_dir_fwd(const float* p_raw, const float* d, float* s) {
  float p[3] = {p_raw[0], p_raw[1], p_raw[2]};
  const float pn = sqrtf(p[0] * p[0] + p[1] * p[1] + p[2] * p[2]);
  if (pn > 0.999f) { for (int c = 0; c < 3; ++c) { p[c] /= pn; p[c] *= 0.999f; } }
  const float dtx = p[0] * d[0] + p[1] * d[1] + p[2] * d[2], xtx = p[0] * p[0] + p[1] * p[1] + p[2] * p[2];
  const float t = -dtx + sqrtf(dtx * dtx - xtx + 1.0f + 1e-6f);
  float q[3];
  for (int c = 0; c < 3; ++c) q[c] = p[c] + d[c] * t;
  const float qn = fmaxf(sqrtf(q[0] * q[0] + q[1] * q[1] + q[2] * q[2]), 1e-12f);
  for (int c = 0; c < 3; ++c) s[c] = q[c] / qn;
}
// (ds/dd)^T gs accumulated into gd (the point is a constant of the render)
NERO_HD void sphere_dir_bwd(const float* p_raw, const float* d, const float* gs, float* gd) {
  float p[3] = {p_raw[0], p_raw[1], p_raw[2]};
  const float pn = sqrtf(p[0] * p[0] + p[1] * p[1] + p[2] * p[2]);
  if (pn > 0.999f) { for (int c = 0; c < 3; ++c) { p[c] /= pn; p[c] *= 0.999f; } }
  const float dtx = p[0] * d[0] + p[1] * d[1] + p[2] * d[2], xtx = p[0] * p[0] + p[1] * p[1] + p[2] * p[2];
  const float root = sqrtf(dtx * dtx - xtx + 1.0f + 1e-6f);
  const float t = -dtx + root;
  float q[3];
  for (int c = 0; c < 3; ++c) q[c] = p[c] + d[c] * t;
  const float qn = fmaxf(sqrtf(q[0] * q[0] + q[1] * q[1] + q[2] * q[2]), 1e-12f);
  float s[3], gq[3];
  for (int c = 0; c < 3; ++c) s[c] = q[c] / qn;
  const float gss = gs[0] * s[0] + gs[1] * s[1] + gs[2] * s[2];
  for (int c = 0; c < 3; ++c) gq[c] = (gs[c] - gss * s[c]) / qn;
  const float gt = gq[0] * d[0] + gq[1] * d[1] + gq[2] * d[2];
  const float dt_ddtx = -1.0f + dtx / root;
  for (int c = 0; c < 3; ++c) gd[c] += gq[c] * t + gt * dt_ddtx * p[c];
}

// ------------------------------------------------------------------ split-sum combine   (network/field.py:601-623, :571-576)
struct ShadeIn {
  float metallic, roughness, albedo[3];
  float Ld[3];        // diffuse light  = outer_light(IDE(n,1))
  float Ldir[3];      // direct light   = outer_light(IDE(r,rough))
  float Li[3];        // indirect light = inner_light(...)
  float iw;           // inner_weight output (occ_prob = 0.5 iw + 0.5)
  float Lh[3], wh;    // human light rgb and weight (already multiplied by the hit mask); 0 when disabled
  float NoV;
};
struct ShadeGrad {
  float metallic, roughness, albedo[3], Ld[3], Ldir[3], Li[3], iw, Lh[3], wh, NoV;
};

NERO_HD void shade_combine_fwd(const ShadeIn& s, const float* lut, float* color) {
  const float occ = clamp01(s.iw * 0.5f + 0.5f);
  const float whc = clamp01(s.wh);
  float fg[2], du[2], dv[2];
  fg_lookup(lut, clamp01(s.NoV), clamp01(s.roughness), fg, du, dv);
  for (int c = 0; c < 3; ++c) {
    const float mix = s.Lh[c] * whc + s.Ldir[c] * (1.0f - whc);
    const float Ls = s.Li[c] * occ + mix * (1.0f - occ);
    const float kd = (1.0f - s.metallic) * s.albedo[c];
    const float F0 = 0.04f * (1.0f - s.metallic) + s.metallic * s.albedo[c];
    const float lin = kd * s.Ld[c] + (F0 * fg[0] + fg[1]) * Ls;
    color[c] = clamp01(linear_to_srgb(lin));
  }
}

NERO_HD void shade_combine_bwd(const ShadeIn& s, const float* lut, const float* dcolor, ShadeGrad& d) {
  const float occ_raw = s.iw * 0.5f + 0.5f;
  const float occ = clamp01(occ_raw);
  const float whc = clamp01(s.wh);
  float fg[2], du[2], dv[2];
  fg_lookup(lut, clamp01(s.NoV), clamp01(s.roughness), fg, du, dv);
  float dm = 0.f, docc = 0.f, dwh = 0.f, dfg0 = 0.f, dfg1 = 0.f;
  for (int c = 0; c < 3; ++c) {
    const float mix = s.Lh[c] * whc + s.Ldir[c] * (1.0f - whc);
    const float Ls = s.Li[c] * occ + mix * (1.0f - occ);
    const float kd = (1.0f - s.metallic) * s.albedo[c];
    const float F0 = 0.04f * (1.0f - s.metallic) + s.metallic * s.albedo[c];
    const float sref = F0 * fg[0] + fg[1];
    const float lin = kd * s.Ld[c] + sref * Ls;
    const float dlin = dcolor[c] * in01(linear_to_srgb(lin)) * dlinear_to_srgb(lin);
    const float dkd = dlin * s.Ld[c];
    d.Ld[c] = dlin * kd;
    const float dsref = dlin * Ls, dLs = dlin * sref;
    const float dF0 = dsref * fg[0];
    dfg0 += dsref * F0;
    dfg1 += dsref;
    dm += dkd * (-s.albedo[c]) + dF0 * (s.albedo[c] - 0.04f);
    d.albedo[c] = dkd * (1.0f - s.metallic) + dF0 * s.metallic;
    d.Li[c] = dLs * occ;
    docc += dLs * (s.Li[c] - mix);
    const float dmix = dLs * (1.0f - occ);
    d.Ldir[c] = dmix * (1.0f - whc);
    d.Lh[c] = dmix * whc;
    dwh += dmix * (s.Lh[c] - s.Ldir[c]);
  }
  d.metallic = dm;
  d.iw = 0.5f * docc * in01(occ_raw);
  d.wh = dwh * in01(s.wh);
  d.NoV = (dfg0 * du[0] + dfg1 * du[1]) * in01(s.NoV);
  d.roughness = (dfg0 * dv[0] + dfg1 * dv[1]) * in01(s.roughness);
}

// ------------------------------------------------------------------ outer NeRF post-processing (renderer.py:346-347, 514-520)
NERO_HD void nerf_post_fwd(float density_raw, float dist, const float* rgb_raw, float* alpha, float* color) {
  const float sig = density_raw > 20.0f ? density_raw : log1pf(expf(density_raw));
  *alpha = 1.0f - expf(-sig * dist);
  for (int c = 0; c < 3; ++c) color[c] = linear_to_srgb(expf(fminf(rgb_raw[c], 5.0f)));
}
NERO_HD void nerf_post_bwd(float density_raw, float dist, const float* rgb_raw, float dalpha, const float* dcolor,
                           float* ddensity, float* drgb) {
  const float sig = density_raw > 20.0f ? density_raw : log1pf(expf(density_raw));
  const float dsig = dalpha * dist * expf(-sig * dist);
  *ddensity = dsig * (density_raw > 20.0f ? 1.0f : sigmoidf_(density_raw));
  for (int c = 0; c < 3; ++c) {
    const float e = expf(fminf(rgb_raw[c], 5.0f));
    drgb[c] = rgb_raw[c] <= 5.0f ? dcolor[c] * dlinear_to_srgb(e) * e : 0.0f;
  }
}

// ------------------------------------------------------------------ human-light plane intersection (field.py:348-367, 536-546)
// pose [3][4] = [R | t].  Outputs mean[2], var (scalar, same for both axes), hit (0/1) -- before IPE.
struct HumanGeo { float mean[2], var, hit, dist, pz, rz, q[2]; };
NERO_HD HumanGeo human_geo_fwd(const float* p, const float* r, const float* pose, float roughness) {
  float pp[3], rr[3];
  for (int i = 0; i < 3; ++i) {
    pp[i] = pose[i * 4 + 0] * p[0] + pose[i * 4 + 1] * p[1] + pose[i * 4 + 2] * p[2] + pose[i * 4 + 3];
    rr[i] = pose[i * 4 + 0] * r[0] + pose[i * 4 + 1] * r[1] + pose[i * 4 + 2] * r[2];
  }
  HumanGeo h;
  bool hit = fabsf(rr[2]) > 1e-4f;
  const float rz = hit ? rr[2] : 1e-4f;
  const float dist = -pp[2] / rz;
  const float qx = pp[0] + dist * rr[0], qy = pp[1] + dist * rr[1];
  h.mean[0] = qx * 0.3f; h.mean[1] = qy * 0.3f;
  h.var = roughness * (dist * 0.3f) * (dist * 0.3f);
  hit = hit && (sqrtf(h.mean[0] * h.mean[0] + h.mean[1] * h.mean[1]) < 1.5f) && (dist > 0.f);
  h.hit = hit ? 1.0f : 0.0f;
  h.mean[0] *= h.hit; h.mean[1] *= h.hit; h.var *= h.hit;
  h.dist = dist; h.pz = pp[2]; h.rz = rz; h.q[0] = rr[0]; h.q[1] = rr[1];
  return h;
}
// given dmean[2], dvar (w.r.t. the masked outputs) -> accumulates dr[3] (world) and returns droughness
NERO_HD float human_geo_bwd(const float* p, const float* r, const float* pose, float roughness, const float* dmean, float dvar,
                            float* dr) {
  float pp[3], rr[3];
  for (int i = 0; i < 3; ++i) {
    pp[i] = pose[i * 4 + 0] * p[0] + pose[i * 4 + 1] * p[1] + pose[i * 4 + 2] * p[2] + pose[i * 4 + 3];
    rr[i] = pose[i * 4 + 0] * r[0] + pose[i * 4 + 1] * r[1] + pose[i * 4 + 2] * r[2];
  }
  const bool hit0 = fabsf(rr[2]) > 1e-4f;
  const float rz = hit0 ? rr[2] : 1e-4f;
  const float dist = -pp[2] / rz;
  const float mx = (pp[0] + dist * rr[0]) * 0.3f, my = (pp[1] + dist * rr[1]) * 0.3f;
  const bool hit = hit0 && (sqrtf(mx * mx + my * my) < 1.5f) && (dist > 0.f);
  if (!hit) return 0.0f;
  // mean = 0.3 (pp_xy + dist rr_xy); var = rough * 0.09 dist^2
  float drr[3] = {0.3f * dmean[0] * dist, 0.3f * dmean[1] * dist, 0.f};
  const float ddist = 0.3f * (dmean[0] * rr[0] + dmean[1] * rr[1]) + dvar * roughness * 0.18f * dist;
  drr[2] = ddist * (pp[2] / (rz * rz));   // d(-pz/rz)/drz (rz == rr[2] since hit0)
  for (int j = 0; j < 3; ++j) dr[j] += pose[0 * 4 + j] * drr[0] + pose[1 * 4 + j] * drr[1] + pose[2 * 4 + j] * drr[2];
  return dvar * 0.09f * dist * dist;
}

}  // namespace nero
