// Row N2 of SURVEY 8(f): the query-point producer on the GPU.
// Replaces the CPU `xarray.DataArray.interp(x=..., y=..., t=...)` of dataset/physics_dataset.py:477-486 (interior points),
// :406-415 (margin points) and :567-576 (dense grid): point-wise trilinear interpolation of the normalised coarse
// (1 degree, 6-hourly) field stack at continuous (lon, lat, hour) queries, plus the Coriolis parameter of :521-526.
// Values only - the reference does not differentiate through this interpolation and neither do we.
#include "dpn_common.cuh"

namespace dpn {

__global__ void sample_field_kernel(const DpnSampler S, const float* __restrict__ coarse, const float* __restrict__ x,
                                    const float* __restrict__ y, const float* __restrict__ t,
                                    float* __restrict__ coord_data, float* __restrict__ f) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)S.B * S.N;
  if (i >= total) return;
  const int b = (int)(i / S.N);
  const double fx = (double)x[i] / S.dx, fy = (double)y[i] / S.dy;          // fine-grid cell coordinates
  double gx = fx / S.cells_per_coarse, gy = fy / S.cells_per_coarse, gt = (double)t[i] / S.t_step;
  // interval search of a regular grid; the last node belongs to the last interval (as scipy's interpn does)
  int ix = min(max((int)floor(gx), 0), S.Wc - 2), iy = min(max((int)floor(gy), 0), S.Hc - 2);
  int it = min(max((int)floor(gt), 0), S.Tt - 2);
  const double wx = gx - ix, wy = gy - iy, wt = gt - it;
  const float* base = coarse + (size_t)b * S.Tt * S.Hc * S.Wc * 6;
  double acc[6] = {0, 0, 0, 0, 0, 0};
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const int dt_ = c >> 2, dy_ = (c >> 1) & 1, dx_ = c & 1;
    const double wgt = (dt_ ? wt : 1.0 - wt) * (dy_ ? wy : 1.0 - wy) * (dx_ ? wx : 1.0 - wx);
    const float2* p = reinterpret_cast<const float2*>(base + (((size_t)(it + dt_) * S.Hc + (iy + dy_)) * S.Wc + (ix + dx_)) * 6);
    const float2 a = __ldg(p), bq = __ldg(p + 1), cq = __ldg(p + 2);     // one 24-byte texel: 6 variables
    acc[0] += wgt * a.x; acc[1] += wgt * a.y; acc[2] += wgt * bq.x;
    acc[3] += wgt * bq.y; acc[4] += wgt * cq.x; acc[5] += wgt * cq.y;
  }
  float2* o = reinterpret_cast<float2*>(coord_data + (size_t)i * 6);
  o[0] = make_float2((float)acc[0], (float)acc[1]);
  o[1] = make_float2((float)acc[2], (float)acc[3]);
  o[2] = make_float2((float)acc[4], (float)acc[5]);
  if (f) {
    const double lat = S.begin_lat + fy * S.deg_per_cell;
    f[i] = (float)(2.0 * S.omega * sin(lat / 180.0 * 3.14159265358979323846));
  }
}

int run_sampler(const DpnSampler& S, const float* coarse, const float* x, const float* y, const float* t,
                float* coord_data, float* f, cudaStream_t st) {
  const long long total = (long long)S.B * S.N;
  sample_field_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(S, coarse, x, y, t, coord_data, f);
  DPN_LAUNCH_OK();
  return 0;
}

}  // namespace dpn
