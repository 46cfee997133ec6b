// Shared device/host definitions of libdpn_b200 (sm_100a).
// Math references: SURVEY.md App. A; DESIGN.md section 3 for the executed algorithm.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/dpn_b200.h"

namespace dpn {

constexpr int H = DPN_H;   // hidden width (variable_net.py: hidden_channels)
constexpr int C = DPN_C;   // encoded coordinate width (in_channels)
constexpr int NF = 32;     // coordinate PE frequencies (interface_physics.py:44: SineCosPE(3), N_freqs=32)
constexpr int NF6 = 16;    // data PE frequencies (variable_net.py:45: SineCosPE(6, N_freqs=192/2/6))

// Device-side copy of DpnConsts plus derived scales and the fp32 frequency buffers of the reference
// (position_encoding.py:27 builds them in fp32: 2**linspace(0,4,F)).
struct DevConsts {
  double sx, sy, st;            // 1/(dx (W-1)), 1/(dy (H-1)), 1/t_span : d z_c / d (x,y,t)
  float dxf, dyf;               // z is formed as the reference does, in fp32: (x / dx) / (W-1)
  float wm1, hm1, t_span;
  int with_clip;
  double mean[6], std[6], lo[6], hi[6], factor[6];
  double c_p, L, R_v, R_d;
  float band[NF];
  float band6[NF6];
};

// torch.linspace(0, 4, n) in fp32 (start + i*step for the first half, end - (n-1-i)*step for the second)
// followed by 2**v.  Fallback only: the Python binding passes the torch-computed buffers (bit-identical to
// position_encoding.py:27); this host version may differ by 1 ulp (moves outputs by ~1e-7 relative).
inline void fill_bands(float* dst, int n) {
  float step = 4.0f / (float)(n - 1);
  for (int i = 0; i < n; ++i) {
    float v = (i < n / 2) ? (0.0f + step * (float)i) : (4.0f - step * (float)(n - 1 - i));
    dst[i] = exp2f(v);
  }
}

inline DevConsts make_dev_consts(const DpnConsts& c) {
  DevConsts d;
  memset(&d, 0, sizeof(d));
  d.sx = 1.0 / (c.dx * (double)(c.lon_size - 1));
  d.sy = 1.0 / (c.dy * (double)(c.lat_size - 1));
  d.st = 1.0 / c.t_span;
  d.dxf = (float)c.dx;
  d.dyf = (float)c.dy;
  d.wm1 = (float)(c.lon_size - 1);
  d.hm1 = (float)(c.lat_size - 1);
  d.t_span = (float)c.t_span;
  d.with_clip = c.with_clip;
  for (int i = 0; i < 6; ++i) {
    d.mean[i] = c.mean[i]; d.std[i] = c.std[i]; d.lo[i] = c.lo[i]; d.hi[i] = c.hi[i]; d.factor[i] = c.factor[i];
  }
  d.c_p = c.c_p; d.L = c.L; d.R_v = c.R_v; d.R_d = c.R_d;
  if (c.band_coord[0] != 0.0f) {
    memcpy(d.band, c.band_coord, sizeof(d.band));
    memcpy(d.band6, c.band_data, sizeof(d.band6));
  } else {
    fill_bands(d.band, NF);
    fill_bands(d.band6, NF6);
  }
  return d;
}

// ------------------------------------------------------------------------------------------------
// Error plumbing (thread-local text for dpn_last_error)
// ------------------------------------------------------------------------------------------------
void set_error(const char* fmt, ...);
extern thread_local int g_launches;

#define DPN_CUDA_OK(expr)                                                                         \
  do {                                                                                            \
    cudaError_t _e = (expr);                                                                      \
    if (_e != cudaSuccess) {                                                                      \
      dpn::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e));       \
      return (int)_e;                                                                             \
    }                                                                                             \
  } while (0)

#define DPN_LAUNCH_OK()                                                                           \
  do {                                                                                            \
    ++dpn::g_launches;                                                                            \
    cudaError_t _e = cudaGetLastError();                                                          \
    if (_e != cudaSuccess) {                                                                      \
      dpn::set_error("%s:%d: kernel launch -> %s", __FILE__, __LINE__, cudaGetErrorString(_e));   \
      return (int)_e;                                                                             \
    }                                                                                             \
  } while (0)

// ------------------------------------------------------------------------------------------------
// Per-point physics: inverse_norm (+clip), the six residuals of interface_physics.py:97-179 and the
// seeds dL/do, dL/d(do/dz_c).  Evaluated in fp64 (a few hundred flops per point).
//   o[k]      normalised net outputs (u,v,p,T,q,rho)
//   od[k][c]  do_k/dz_c, z = normalised (x,y,t)
//   inv_n     1 / n_norm ; seed_scale multiplies the seeds only
// Outputs: r2w[e] = factor_e * r_e^2 * inv_n ; dov[k], dod[k][c] ; vals[k], jac[k][c] (physical).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void residual_point(const DevConsts& K, const float* o, const float* od, double f,
                                               double inv_n, double seed_scale, double* r2w, double* dov,
                                               double* dod, double* vals, double* jac) {
  double kap[6];
  const double sc[3] = {K.sx, K.sy, K.st};
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    double raw = (double)o[k] * K.std[k] + K.mean[k];
    kap[k] = 1.0;
    if (K.with_clip && k >= 2) {                        // interface_physics.py:256-261: u, v never clipped
      kap[k] = (raw >= K.lo[k] && raw <= K.hi[k]) ? 1.0 : 0.0;   // clamp backward: 1 on the closed interval
      raw = fmin(fmax(raw, K.lo[k]), K.hi[k]);
    }
    vals[k] = raw;
#pragma unroll
    for (int c = 0; c < 3; ++c) jac[k * 3 + c] = (double)od[k * 3 + c] * K.std[k] * kap[k] * sc[c];
  }
  const double u = vals[0], v = vals[1], p = vals[2], T = vals[3], q = vals[4], r = vals[5];
#define GX(i) jac[(i)*3 + 0]
#define GY(i) jac[(i)*3 + 1]
#define GT(i) jac[(i)*3 + 2]
#define DD(i) (GT(i) + u * GX(i) + v * GY(i))
  const double eps = 1e-6;
  const double Dp = DD(2), Dq = DD(4);
  double res[6];
  res[0] = DD(0) + GX(2) / r - f * v;                                    // :97-104
  res[1] = DD(1) + GY(2) / r + f * u;                                    // :106-114
  res[2] = DD(5) + r * (GX(0) + GY(1));                                  // :116-124
  res[3] = K.c_p * DD(3) - Dp / (r + eps) + K.L * Dq;                    // :126-144
  const double tc = T - 273.15;                                          // :181-185
  const double e_s = 6.112 * exp(17.67 * tc / (tc + 243.5)) * 100.0;
  double q_s = 0.622 * e_s / (p - 0.378 * e_s);
  q_s = fmax(q_s, 1e-6);                                                 // :166 (NaN-propagation differs from torch.maximum only for NaN inputs)
  const double delta = (Dp < 0.0 && q >= q_s) ? 1.0 : 0.0;               // :147-149
  const double Rm = (1.0 + 0.608 * q) * K.R_d;
  const double Fv = (K.L * Rm - K.c_p * K.R_v * T) / (K.c_p * K.R_v + T * T + K.L * K.L * q_s) * q_s * T;  // :151-155
  const double Kf = delta * Fv / (p + eps);
  res[4] = -Dp * Kf + Dq;                                                // :171-173
  res[5] = p - r * (1.0 + 0.608 * q) * K.R_d * T;                        // :177-179
  double a[6];
#pragma unroll
  for (int e = 0; e < 6; ++e) {
    r2w[e] = K.factor[e] * res[e] * res[e] * inv_n;
    a[e] = 2.0 * K.factor[e] * res[e] * inv_n * seed_scale;
  }
  double dv[6] = {0, 0, 0, 0, 0, 0};
  double dj[18];
#pragma unroll
  for (int i = 0; i < 18; ++i) dj[i] = 0.0;
#define ADD_D(i, coef)                                  \
  do {                                                  \
    double _c = (coef);                                 \
    dj[(i)*3 + 2] += _c;                                \
    dj[(i)*3 + 0] += _c * u;                            \
    dj[(i)*3 + 1] += _c * v;                            \
    dv[0] += _c * GX(i);                                \
    dv[1] += _c * GY(i);                                \
  } while (0)
  ADD_D(0, a[0]); dj[2 * 3 + 0] += a[0] / r; dv[5] += -a[0] * GX(2) / (r * r); dv[1] += -a[0] * f;
  ADD_D(1, a[1]); dj[2 * 3 + 1] += a[1] / r; dv[5] += -a[1] * GY(2) / (r * r); dv[0] += a[1] * f;
  ADD_D(5, a[2]); dv[5] += a[2] * (GX(0) + GY(1)); dj[0 * 3 + 0] += a[2] * r; dj[1 * 3 + 1] += a[2] * r;
  ADD_D(3, a[3] * K.c_p); ADD_D(2, -a[3] / (r + eps)); ADD_D(4, a[3] * K.L);
  dv[5] += a[3] * Dp / ((r + eps) * (r + eps));
  ADD_D(2, -a[4] * Kf); ADD_D(4, a[4]); dv[2] += a[4] * Dp * delta * Fv / ((p + eps) * (p + eps));
  dv[2] += a[5];
  dv[5] += -a[5] * (1.0 + 0.608 * q) * K.R_d * T;
  dv[4] += -a[5] * r * 0.608 * K.R_d * T;
  dv[3] += -a[5] * r * (1.0 + 0.608 * q) * K.R_d;
#undef ADD_D
#undef DD
#undef GX
#undef GY
#undef GT
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    dov[k] = dv[k] * K.std[k] * kap[k];
#pragma unroll
    for (int c = 0; c < 3; ++c) dod[k * 3 + c] = dj[k * 3 + c] * K.std[k] * kap[k] * sc[c];
  }
}

}  // namespace dpn
