// Per-sample encoding / colour math, forward and hand-derived backward, as host+device inline functions.
// The CUDA kernels (k_render.cu, k_shade.cu) call these per thread; tests/hostcheck compiles the very same
// functions with g++ so the derivations are checked on CPU against the oracle before touching a GPU.
#pragma once
#include <cmath>
#include "common.cuh"

namespace nero {

// ---------------------------------------------------------------- positional encoding
// out = [x(d), sin(2^0 x)(d), cos(2^0 x)(d), ..., sin(2^(L-1) x)(d), cos(2^(L-1) x)(d)]   network/field.py:14-58
template <int D>
NERO_HD void pe_encode(const float* x, int L, float* out, float scale = 1.0f) {
  for (int c = 0; c < D; ++c) out[c] = x[c] * scale;
  float f = 1.0f;
  for (int k = 0; k < L; ++k) {
    for (int c = 0; c < D; ++c) {
      float sn, cs;
      sincosf(x[c] * f, &sn, &cs);             // one shared range reduction (the encodings are ~half of a fill kernel's instructions)
      out[D + (2 * k) * D + c] = sn * scale;
      out[D + (2 * k + 1) * D + c] = cs * scale;
    }
    f *= 2.0f;
  }
}
// g[c] = sum_j dPE_j/dx_c * u[j]  given the PE VALUES (d sin(fx) = f cos(fx), d cos(fx) = -f sin(fx))
template <int D>
NERO_HD void pe_backward(const float* pe, int L, const float* u, float* g) {
  for (int c = 0; c < D; ++c) {
    float acc = u[c];
    float f = 1.0f;
    for (int k = 0; k < L; ++k) {
      const float s = pe[D + (2 * k) * D + c], co = pe[D + (2 * k + 1) * D + c];
      acc += f * (co * u[D + (2 * k) * D + c] - s * u[D + (2 * k + 1) * D + c]);
      f *= 2.0f;
    }
    g[c] = acc;
  }
}
// tangent: t[j] = sum_c dPE_j/dx_c * d[c]
template <int D>
NERO_HD void pe_tangent(const float* pe, int L, const float* d, float* t, float scale = 1.0f) {
  for (int c = 0; c < D; ++c) t[c] = d[c] * scale;
  float f = 1.0f;
  for (int k = 0; k < L; ++k) {
    for (int c = 0; c < D; ++c) {
      const float s = pe[D + (2 * k) * D + c], co = pe[D + (2 * k + 1) * D + c];
      t[D + (2 * k) * D + c] = f * co * d[c] * scale;
      t[D + (2 * k + 1) * D + c] = -f * s * d[c] * scale;
    }
    f *= 2.0f;
  }
}

// ---------------------------------------------------------------- integrated directional encoding
// utils/ref_utils.py:53-117.  36 (m,l) pairs: l in {1,2,4,8,16}, m = 0..l.  out = [Re(36) | Im(36)].
// The z-polynomials P_i(z) = sum_k mat[k][i] z^k are ill-conditioned near |z| = 1 in fp32 (the l = 16 band cancels ~1e5 :
// the reference's own result is up to 2.4e-3 away from exact arithmetic there), so the VALUE is evaluated exactly the way
// the reference evaluates it -- verified bit for bit against torch on 7e5 entries (oracle/make_golden.py, round 2):
//   * vmz[k] = z**k is the correctly rounded fp32 power (torch.pow): here the fp64 power rounded to fp32;
//   * torch.matmul(vmz, mat) on the CPU accumulates k = 0..16 in ascending order with one fp32 FMA per term.
// The backward pass (no bit-level reference exists for it) differentiates the exact polynomial: Horner in fp64.
struct IdeTable {
  double mat[17][36];
  float matf[17][36];
  int m[36];
  int l[36];
};

NERO_HD void ide_poly(const IdeTable& T, int i, double z, double& P, double& dP) {
  const int deg = T.l[i] - T.m[i];
  double p = T.mat[deg][i], dp = 0.0;
  for (int k = deg - 1; k >= 0; --k) {
    dp = dp * z + p;
    p = p * z + T.mat[k][i];
  }
  P = p; dP = dp;
}
// vmz[k] = fp32(z^k), k = 0..16
NERO_HD void ide_powers(float z, float* pw) {
  double p = 1.0;
  const double zd = double(z);
  for (int k = 0; k <= 16; ++k) { pw[k] = float(p); p *= zd; }
}
// the reference's fp32 value of P_i(z): ascending-k FMA chain over the Vandermonde row
NERO_HD float ide_poly_value(const IdeTable& T, int i, const float* pw) {
  const int deg = T.l[i] - T.m[i];
  float acc = 0.0f;
  for (int k = 0; k <= deg; ++k) acc = fmaf(pw[k], T.matf[k][i], acc);
  return acc;
}

NERO_HD void ide_forward(const IdeTable& T, const float* d, float kappa_inv, float* out) {
  float pr[17], pi[17], pw[17];
  pr[0] = 1.0f; pi[0] = 0.0f;
  for (int m = 1; m <= 16; ++m) { pr[m] = pr[m - 1] * d[0] - pi[m - 1] * d[1]; pi[m] = pr[m - 1] * d[1] + pi[m - 1] * d[0]; }
  ide_powers(d[2], pw);
  for (int i = 0; i < 36; ++i) {
    const float P = ide_poly_value(T, i, pw);
    const float sigma = 0.5f * float(T.l[i] * (T.l[i] + 1));
    const float att = expf(-sigma * kappa_inv);
    out[i] = (pr[T.m[i]] * P) * att;
    out[36 + i] = (pi[T.m[i]] * P) * att;
  }
}
// dout[72] -> dd[3] (accumulated into), dkappa (returned)
NERO_HD float ide_backward(const IdeTable& T, const float* d, float kappa_inv, const float* dout, float* dd) {
  float pr[17], pi[17];
  pr[0] = 1.0f; pi[0] = 0.0f;
  for (int m = 1; m <= 16; ++m) { pr[m] = pr[m - 1] * d[0] - pi[m - 1] * d[1]; pi[m] = pr[m - 1] * d[1] + pi[m - 1] * d[0]; }
  float dx = 0.f, dy = 0.f, dz = 0.f, dk = 0.f;
  for (int i = 0; i < 36; ++i) {
    double P, dP;
    ide_poly(T, i, double(d[2]), P, dP);
    const int m = T.m[i];
    const float sigma = 0.5f * float(T.l[i] * (T.l[i] + 1));
    const float att = expf(-sigma * kappa_inv);
    const float A = float(P) * att;
    const float gr = dout[i], gi = dout[36 + i];
    if (m > 0) {
      const float wr = float(m) * pr[m - 1], wi = float(m) * pi[m - 1];
      dx += A * (gr * wr + gi * wi);
      dy += A * (-gr * wi + gi * wr);
    }
    const float zc = gr * pr[m] + gi * pi[m];
    dz += zc * att * float(dP);
    dk -= sigma * zc * A;
  }
  dd[0] += dx; dd[1] += dy; dd[2] += dz;
  return dk;
}

// ---------------------------------------------------------------- integrated positional encoding (human light)
// network/field.py:369-378 with min_deg=0,max_deg=6 on a 2-d mean: out[24] = [sin part (12) | cos part (12)],
// index = k*2 + c for scale 2^k; value exp(-0.5*var*4^k) * sin(mean*2^k (+pi/2)).
NERO_HD void ipe_forward(const float* mean, const float* var, float* out) {
  float f = 1.0f;
  for (int k = 0; k < 6; ++k) {
    for (int c = 0; c < 2; ++c) {
      const float e = expf(-0.5f * var[c] * f * f);
      const float a = mean[c] * f;
      out[k * 2 + c] = e * sinf(a);
      out[12 + k * 2 + c] = e * sinf(a + 0.5f * 3.14159265358979323846f);
    }
    f *= 2.0f;
  }
}
NERO_HD void ipe_backward(const float* mean, const float* var, const float* dout, float* dmean, float* dvar) {
  dmean[0] = dmean[1] = dvar[0] = dvar[1] = 0.f;
  float f = 1.0f;
  for (int k = 0; k < 6; ++k) {
    for (int c = 0; c < 2; ++c) {
      const float e = expf(-0.5f * var[c] * f * f);
      const float a = mean[c] * f;
      const float a2 = a + 0.5f * 3.14159265358979323846f;
      const float g1 = dout[k * 2 + c], g2 = dout[12 + k * 2 + c];
      dmean[c] += f * e * (g1 * cosf(a) + g2 * cosf(a2));
      dvar[c] += -0.5f * f * f * e * (g1 * sinf(a) + g2 * sinf(a2));
    }
    f *= 2.0f;
  }
}

// ---------------------------------------------------------------- sRGB  (utils/raw_utils.py:4-10)
NERO_HD float linear_to_srgb(float x) {
  const float eps = 1.1920928955078125e-07f;
  return x <= 0.0031308f ? (323.0f / 25.0f) * x : (211.0f * powf(fmaxf(x, eps), 5.0f / 12.0f) - 11.0f) / 200.0f;
}
NERO_HD float dlinear_to_srgb(float x) {
  const float eps = 1.1920928955078125e-07f;
  if (x <= 0.0031308f) return 323.0f / 25.0f;
  if (x < eps) return 0.0f;
  return (211.0f / 200.0f) * (5.0f / 12.0f) * powf(x, -7.0f / 12.0f);
}

// ---------------------------------------------------------------- split-sum LUT (bilinear, clamp)
// network/field.py:610-612 via nvdiffrast.texture(filter 'linear', boundary 'clamp'); lut[row=v][col=u][2], 256x256.
// Returns fg[2] and d fg / d(u,v).
NERO_HD void fg_lookup(const float* lut, float u, float v, float* fg, float* dfg_du, float* dfg_dv) {
  const int N = 256;
  const float fx = u * N - 0.5f, fy = v * N - 0.5f;
  const float x0f = floorf(fx), y0f = floorf(fy);
  const float tx = fx - x0f, ty = fy - y0f;
  int x0 = int(x0f), y0 = int(y0f);
  int x1 = x0 + 1, y1 = y0 + 1;
  x0 = x0 < 0 ? 0 : (x0 > N - 1 ? N - 1 : x0); x1 = x1 < 0 ? 0 : (x1 > N - 1 ? N - 1 : x1);
  y0 = y0 < 0 ? 0 : (y0 > N - 1 ? N - 1 : y0); y1 = y1 < 0 ? 0 : (y1 > N - 1 ? N - 1 : y1);
  for (int c = 0; c < 2; ++c) {
    const float t00 = lut[(y0 * N + x0) * 2 + c], t01 = lut[(y0 * N + x1) * 2 + c];
    const float t10 = lut[(y1 * N + x0) * 2 + c], t11 = lut[(y1 * N + x1) * 2 + c];
    const float a = t00 * (1.f - tx) + t01 * tx, b = t10 * (1.f - tx) + t11 * tx;
    fg[c] = a * (1.f - ty) + b * ty;
    dfg_du[c] = float(N) * ((t01 - t00) * (1.f - ty) + (t11 - t10) * ty);
    dfg_dv[c] = float(N) * (b - a);
  }
}

}  // namespace nero
