// Stage II (material estimation) per-ray math: importance-sampled directions and the microfacet BRDF weights of
// MCShadingNetwork.shade_mixed (network/field.py:768-812, 886-1003).  Everything is templated on the scalar type so
// that the SAME expression tree evaluates either plain floats (forward) or forward-mode duals d/d(roughness) (backward):
// the roughness is the only learnable quantity these terms depend on (normals, view directions and points are data).
#pragma once
#include "common.cuh"

namespace nero {

struct Dual { float v, d; };
NERO_HD Dual mk(float v, float d = 0.f) { Dual r; r.v = v; r.d = d; return r; }
NERO_HD Dual operator+(Dual a, Dual b) { return mk(a.v + b.v, a.d + b.d); }
NERO_HD Dual operator-(Dual a, Dual b) { return mk(a.v - b.v, a.d - b.d); }
NERO_HD Dual operator*(Dual a, Dual b) { return mk(a.v * b.v, a.d * b.v + a.v * b.d); }
NERO_HD Dual operator/(Dual a, Dual b) { const float q = a.v / b.v; return mk(q, (a.d - q * b.d) / b.v); }
NERO_HD Dual operator+(Dual a, float b) { return mk(a.v + b, a.d); }
NERO_HD Dual operator+(float a, Dual b) { return mk(a + b.v, b.d); }
NERO_HD Dual operator-(Dual a, float b) { return mk(a.v - b, a.d); }
NERO_HD Dual operator-(float a, Dual b) { return mk(a - b.v, -b.d); }
NERO_HD Dual operator*(Dual a, float b) { return mk(a.v * b, a.d * b); }
NERO_HD Dual operator*(float a, Dual b) { return mk(a * b.v, a * b.d); }
NERO_HD Dual operator/(Dual a, float b) { return mk(a.v / b, a.d / b); }
NERO_HD Dual operator/(float a, Dual b) { const float q = a / b.v; return mk(q, -q * b.d / b.v); }

NERO_HD float val(float x) { return x; }
NERO_HD float val(Dual x) { return x.v; }
NERO_HD float der(float) { return 0.f; }
NERO_HD float der(Dual x) { return x.d; }
NERO_HD float t_sqrt(float x) { return sqrtf(x); }
NERO_HD Dual t_sqrt(Dual x) { const float s = sqrtf(x.v); return mk(s, 0.5f * x.d / s); }
// torch.clamp passes the gradient on the closed interval [lo, hi]
NERO_HD float t_clamp01(float x) { return fminf(fmaxf(x, 0.f), 1.f); }
NERO_HD Dual t_clamp01(Dual x) { return mk(fminf(fmaxf(x.v, 0.f), 1.f), (x.v >= 0.f && x.v <= 1.f) ? x.d : 0.f); }
NERO_HD float t_pow5(float x) { const float x2 = x * x; return x2 * x2 * x; }
NERO_HD Dual t_pow5(Dual x) { const float x2 = x.v * x.v; return mk(x2 * x2 * x.v, 5.f * x2 * x2 * x.d); }
template <class T> NERO_HD T lift(float x);
template <> NERO_HD float lift<float>(float x) { return x; }
template <> NERO_HD Dual lift<Dual>(float x) { return mk(x, 0.f); }

constexpr float kPi = 3.14159265358979323846f;

// get_orthogonal_directions (field.py:755-766) + y = z cross x (field.py:771-772): tangent frame around a unit vector
NERO_HD void mc_frame(const float* z, float* x, float* y) {
  const float o0[3] = {z[1], -z[0], 0.f}, o1[3] = {-z[2], 0.f, z[0]};
  const float n0 = sqrtf(o0[0] * o0[0] + o0[1] * o0[1]), n1 = sqrtf(o1[0] * o1[0] + o1[2] * o1[2]);
  const float* o = n0 > n1 ? o0 : o1;
  const float inv = 1.0f / fmaxf(n0 > n1 ? n0 : n1, 1e-12f);
  x[0] = o[0] * inv; x[1] = o[1] * inv; x[2] = o[2] * inv;
  y[0] = z[1] * x[2] - z[2] * x[1]; y[1] = z[2] * x[0] - z[0] * x[2]; y[2] = z[0] * x[1] - z[1] * x[0];
}

NERO_HD void normalize3(const float* a, float* out) {
  const float inv = 1.0f / fmaxf(sqrtf(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]), 1e-12f);
  out[0] = a[0] * inv; out[1] = a[1] * inv; out[2] = a[2] * inv;
}

// per-point constants of shade_mixed
struct McPoint {
  float n[3], v[3], refl[3];      // unit normal, unit view direction, reflection of v about n
  float xd[3], yd[3];             // diffuse frame (z = n)
  float xs[3], ys[3];             // specular frame (z = refl)
  float NoV;
};
NERO_HD McPoint mc_point(const float* normal_raw, const float* view_raw) {
  McPoint q;
  normalize3(normal_raw, q.n);
  normalize3(view_raw, q.v);
  const float vn = q.v[0] * q.n[0] + q.v[1] * q.n[1] + q.v[2] * q.n[2];
  for (int i = 0; i < 3; ++i) q.refl[i] = vn * q.n[i] * 2.0f - q.v[i];
  mc_frame(q.n, q.xd, q.yd);
  mc_frame(q.refl, q.xs, q.ys);
  q.NoV = fminf(fmaxf(vn, 0.f), 1.f);
  return q;
}

// azimuth handling shared by both samplers: (az*2pi + rand*2pi) mod 2pi   (field.py:781-783, 803-806)
NERO_HD float mc_azimuth(float az01, float rnd, bool use_rnd) {
  float az = az01 * kPi * 2.0f;
  if (use_rnd) az = fmodf(az + rnd * kPi * 2.0f, 2.0f * kPi);
  return az;
}

// cosine-weighted direction around the normal (sample_diffuse_directions, field.py:768-788); no roughness dependence
NERO_HD void mc_diffuse_dir(const McPoint& q, float az, float el, float* d) {
  const float es = sqrtf(el + 1e-7f), cz = sqrtf(1.0f - el + 1e-7f);
  const float cx = es * cosf(az), cy = es * sinf(az);
  for (int i = 0; i < 3; ++i) d[i] = cx * q.xd[i] + cy * q.yd[i] + cz * q.n[i];
}
// GGX half-angle distribution around the reflection (sample_specular_directions, field.py:790-812)
template <class T>
NERO_HD void mc_specular_dir(const McPoint& q, float phi, float el, T a, T* d) {
  const T cos_t = t_sqrt((1.0f - el + 1e-6f) / (1.0f + (a * a - 1.0f) * el + 1e-6f) + 1e-6f);
  const T sin_t = t_sqrt(1.0f - cos_t * cos_t + 1e-6f);
  const float cp = cosf(phi), sp = sinf(phi);
  for (int i = 0; i < 3; ++i) d[i] = (cp * q.xs[i] + sp * q.ys[i]) * sin_t + q.refl[i] * cos_t;
}

template <class T>
NERO_HD T mc_ggx_d(T NoH, T a) {   // distribution_ggx, field.py:916-921
  const T a2 = a * a;
  const T den = NoH * NoH * (a2 - 1.0f) + 1.0f;
  return a2 / (kPi * den * den + 1e-4f);
}
template <class T>
NERO_HD T mc_geometry(float NoV, T NoL, T a, int ggx_smith) {   // field.py:869-894, 923-930
  if (!ggx_smith) {
    const T k = a / 2.0f;
    const T g2 = NoV / (NoV * (1.0f - k) + k + 1e-5f);
    const T g1 = NoL / (NoL * (1.0f - k) + k + 1e-5f);
    return g2 * g1;
  }
  const T a2 = a * a;
  const float cv2 = NoV * NoV;
  const T lv = 0.5f * t_sqrt(1.0f + a2 * ((1.0f - cv2) / (cv2 + 1e-7f))) - 0.5f;
  const T cl2 = NoL * NoL;
  const T ll = 0.5f * t_sqrt(1.0f + a2 * ((1.0f - cl2) / (cl2 + 1e-7f))) - 0.5f;
  return 1.0f / (1.0f + lv + ll);
}

// specular weight D*G/(4 NoV p + 1e-5) and Schlick factor (1-HoV)^5 of one sampled direction (field.py:941-975)
template <class T>
NERO_HD void mc_weights(const McPoint& q, const T* dir, T a, bool specular_sample, float frac_d, float frac_s, int ggx_smith, T& w,
                        T& f5) {
  T h[3];
  for (int i = 0; i < 3; ++i) h[i] = q.v[i] + dir[i];
  const T hn = t_sqrt(h[0] * h[0] + h[1] * h[1] + h[2] * h[2]);
  const T inv = 1.0f / (val(hn) > 1e-12f ? hn : lift<T>(1e-12f));
  for (int i = 0; i < 3; ++i) h[i] = h[i] * inv;
  const T HoV = t_clamp01(h[0] * q.v[0] + h[1] * q.v[1] + h[2] * q.v[2]);
  const T NoH = t_clamp01(h[0] * q.n[0] + h[1] * q.n[1] + h[2] * q.n[2]);
  const T NoL = t_clamp01(dir[0] * q.n[0] + dir[1] * q.n[1] + dir[2] * q.n[2]);
  const T D = mc_ggx_d(NoH, a);
  T prob;
  if (specular_sample) prob = D * NoH / (4.0f * HoV + 1e-5f) * frac_s;
  else prob = NoL / kPi * frac_d;
  const T G = mc_geometry(q.NoV, NoL, a, ggx_smith);
  w = D * G / (4.0f * q.NoV * prob + 1e-5f);
  f5 = t_pow5(t_clamp01(1.0f - HoV));
}

// get_sphere_intersection (field.py:390-396) for the 'sphere_direction' outer light (field.py:843-853)
NERO_HD void mc_sphere_point(const float* p_raw, const float* d, float* sp, float* p_used) {
  float p[3] = {p_raw[0], p_raw[1], p_raw[2]};
  if (sqrtf(p[0] * p[0] + p[1] * p[1] + p[2] * p[2]) > 0.999f) { p[0] *= 0.999f; p[1] *= 0.999f; p[2] *= 0.999f; }
  const float dtx = p[0] * d[0] + p[1] * d[1] + p[2] * d[2], xtx = p[0] * p[0] + p[1] * p[1] + p[2] * p[2];
  const float dist = -dtx + sqrtf(dtx * dtx - xtx + 1.0f + 1e-6f);
  for (int i = 0; i < 3; ++i) { sp[i] = p[i] + d[i] * dist; p_used[i] = p[i]; }
}
// d(sphere point)/d(direction) applied to an incoming gradient gs -> accumulates into gd
NERO_HD void mc_sphere_point_bwd(const float* p, const float* d, const float* gs, float* gd) {
  const float dtx = p[0] * d[0] + p[1] * d[1] + p[2] * d[2], xtx = p[0] * p[0] + p[1] * p[1] + p[2] * p[2];
  const float root = sqrtf(dtx * dtx - xtx + 1.0f + 1e-6f);
  const float dist = -dtx + root;
  const float gdist = gs[0] * d[0] + gs[1] * d[1] + gs[2] * d[2];
  const float ddist_ddtx = -1.0f + dtx / root;
  for (int i = 0; i < 3; ++i) gd[i] += gs[i] * dist + gdist * ddist_ddtx * p[i];
}

}  // namespace nero
