// Row N2 of SURVEY 8(f): the query-point producer on the GPU.
// Replaces the CPU `xarray.DataArray.interp(x=..., y=..., t=...)` of dataset/physics_dataset.py:477-486 (interior points),
// :406-415 (margin points) and :567-576 (dense grid): point-wise trilinear interpolation of the normalised coarse
// (1 degree, 6-hourly) field stack at continuous (lon, lat, hour) queries, plus the Coriolis parameter of :521-526.
// Values only - the reference does not differentiate through this interpolation and neither do we.
#include "dpn_common.cuh"

namespace dpn {

// Coordinates -> cell index and weights in fp64 (the fractional part of a cell coordinate ~64 needs more than fp32 to stay within
// 2e-6 of the fp64 reference); the 8 x 6 multiply-adds in fp32; the [N,6] output leaves through a shared-memory transpose so that
// every warp store covers 128 contiguous bytes.  The texels (6 floats = 24 bytes) are fetched with three 8-byte read-only loads
// from the L2-resident coarse stack.
__global__ void __launch_bounds__(256) sample_field_kernel(const DpnSampler S, const float* __restrict__ coarse, const float* __restrict__ x,
                                                           const float* __restrict__ y, const float* __restrict__ t,
                                                           float* __restrict__ coord_data, float* __restrict__ f) {
  __shared__ float stage[8][32 * 6];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long total = (long long)S.B * S.N;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const bool valid = i < total;
  const double rx = 1.0 / (S.dx * S.cells_per_coarse), ry = 1.0 / (S.dy * S.cells_per_coarse), rt = 1.0 / S.t_step;
  float acc[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  float yv = 0.f;
  if (valid) {
    const int b = (int)(i / S.N);
    yv = y[i];
    const double gx = (double)x[i] * rx, gy = (double)yv * ry, gt = (double)t[i] * rt;
    // interval search of a regular grid; the last node belongs to the last interval (as scipy's interpn does)
    const int ix = min(max((int)floor(gx), 0), S.Wc - 2), iy = min(max((int)floor(gy), 0), S.Hc - 2);
    const int it = min(max((int)floor(gt), 0), S.Tt - 2);
    const float wx = (float)(gx - ix), wy = (float)(gy - iy), wt = (float)(gt - it);
    const float* base = coarse + (size_t)b * S.Tt * S.Hc * S.Wc * 6;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const int dt_ = c >> 2, dy_ = (c >> 1) & 1, dx_ = c & 1;
      const float wgt = (dt_ ? wt : 1.f - wt) * (dy_ ? wy : 1.f - wy) * (dx_ ? wx : 1.f - wx);
      const float2* p = reinterpret_cast<const float2*>(base + (((size_t)(it + dt_) * S.Hc + (iy + dy_)) * S.Wc + (ix + dx_)) * 6);
      const float2 a = __ldg(p), bq = __ldg(p + 1), cq = __ldg(p + 2);     // one 24-byte texel: 6 variables
      acc[0] = fmaf(wgt, a.x, acc[0]); acc[1] = fmaf(wgt, a.y, acc[1]); acc[2] = fmaf(wgt, bq.x, acc[2]);
      acc[3] = fmaf(wgt, bq.y, acc[3]); acc[4] = fmaf(wgt, cq.x, acc[4]); acc[5] = fmaf(wgt, cq.y, acc[5]);
    }
  }
#pragma unroll
  for (int c = 0; c < 6; ++c) stage[warp][lane * 6 + c] = acc[c];
  __syncwarp();
  const long long w0 = ((long long)blockIdx.x * blockDim.x + warp * 32) * 6;       // first output word of this warp
#pragma unroll
  for (int c = 0; c < 6; ++c) {
    const long long word = w0 + c * 32 + lane;
    if (word < total * 6) coord_data[word] = stage[warp][c * 32 + lane];
  }
  if (f && valid) {
    const double lat = S.begin_lat + (double)yv / S.dy * S.deg_per_cell;
    f[i] = (float)(2.0 * S.omega) * sinf((float)(lat * (3.14159265358979323846 / 180.0)));
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
