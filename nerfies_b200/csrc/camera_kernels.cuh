// Camera -> rays on the GPU: what the reference does on the
// host in numpy for every evaluated frame,
//   datasets/core.py:50-75  camera_to_rays
//   camera.py:317-321       get_pixel_centers
//   camera.py:225-269       pixel_to_local_rays / pixels_to_rays
//   camera.py:26-105        radial + tangential undistortion (10 Newton steps)
// HBM-bound elementwise work: 0 or 8 B read and 24-32 B written per pixel; one
// thread per pixel, fully coalesced stores of consecutive pixels.  fp32, written
// operation for operation as the reference's float32 numpy (the build uses
// -fmad=false and IEEE division / sqrt).
#pragma once
#include "../../include/nerfies_b200.h"

namespace nfb {

struct CameraArgs {
  nfb_camera cam;
  const float* pixels_in;   // (n,2) or nullptr: pixel centres of the linear range
  long long first, count;
  float* origins;           // (n,3) nullable
  float* directions;        // (n,3)
  float* pixels_out;        // (n,2) nullable
  int has_distortion;
};

__device__ __forceinline__ void undistort_point(float xd, float yd, float k1, float k2, float k3,
                                                float p1, float p2, float& xo, float& yo) {
  float x = xd, y = yd;
  const float eps = 1e-9f;
  const float two_p1 = 2.f * p1, two_p2 = 2.f * p2, six_p1 = 6.f * p1, six_p2 = 6.f * p2;
  const float two_k2 = 2.f * k2, three_k3 = 3.f * k3;
#pragma unroll 1
  for (int it = 0; it < 10; ++it) {                       // camera.py:90 (no early exit)
    const float r = x * x + y * y;                        // camera.py:42
    const float d = 1.f + r * (k1 + r * (k2 + k3 * r));   // camera.py:43
    const float fx = d * x + two_p1 * x * y + p2 * (r + 2.f * x * x) - xd;   // camera.py:55
    const float fy = d * y + two_p2 * x * y + p1 * (r + 2.f * y * y) - yd;   // camera.py:56
    const float d_r = k1 + r * (two_k2 + three_k3 * r);   // camera.py:59
    const float d_x = 2.f * x * d_r;
    const float d_y = 2.f * y * d_r;
    const float fx_x = d + d_x * x + two_p1 * y + six_p2 * x;   // camera.py:64
    const float fx_y = d_y * x + two_p1 * x + two_p2 * y;
    const float fy_x = d_x * y + two_p2 * y + two_p1 * x;       // camera.py:68
    const float fy_y = d + d_y * y + two_p2 * x + six_p1 * y;
    const float den = fy_x * fx_y - fx_x * fy_y;                // camera.py:93-95
    const float xn = fx * fy_y - fy * fx_y;
    const float yn = fy * fx_x - fx * fy_x;
    const bool ok = fabsf(den) > eps;
    x = x + (ok ? xn / den : 0.f);
    y = y + (ok ? yn / den : 0.f);
  }
  xo = x; yo = y;
}

// One pixel: unit world-space direction and the pixel position used.
__device__ __forceinline__ void camera_ray_of_pixel(const CameraArgs& a, long long i, float* out_d, float* out_p) {
  const nfb_camera& c = a.cam;
  float px, py;
  if (a.pixels_in) {
    px = __ldg(a.pixels_in + 2 * i);
    py = __ldg(a.pixels_in + 2 * i + 1);
  } else {
    const long long p = a.first + i;                       // row-major pixel index
    const long long row = p / c.image_size[0];
    px = (float)(p - row * c.image_size[0]) + 0.5f;        // camera.py:319-321
    py = (float)row + 0.5f;
  }
  const float sy = c.focal_length * c.pixel_aspect_ratio;  // camera.py:186-187
  float y = (py - c.principal_point[1]) / sy;              // camera.py:227
  float x = (px - c.principal_point[0] - y * c.skew) / c.focal_length;   // camera.py:228-229
  if (a.has_distortion)
    undistort_point(x, y, c.radial_distortion[0], c.radial_distortion[1], c.radial_distortion[2],
                    c.tangential_distortion[0], c.tangential_distortion[1], x, y);
  // camera.py:241-242: dirs / ||dirs||, dirs = (x, y, 1)
  const float n = sqrtf(x * x + y * y + 1.f);
  const float l0 = x / n, l1 = y / n, l2 = 1.f / n;
  // camera.py:262: orientation^T @ local (orientation is world-to-camera, row-major)
  float d[3];
#pragma unroll
  for (int j = 0; j < 3; ++j)
    d[j] = c.orientation[j] * l0 + c.orientation[3 + j] * l1 + c.orientation[6 + j] * l2;
  const float n2 = sqrtf(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);        // camera.py:266
#pragma unroll
  for (int j = 0; j < 3; ++j) d[j] = d[j] / n2;
  out_d[0] = d[0]; out_d[1] = d[1]; out_d[2] = d[2];
  out_p[0] = px; out_p[1] = py;
}

// 12-byte-per-pixel outputs are staged through shared memory so that a block
// writes its 256 x 12 B as 192 aligned 16-byte vectors (a per-thread stride-12
// scalar store pattern leaves HBM sectors partially written: 208 GB/s measured on
// an 8K frame before this, see profiles/).
__global__ void __launch_bounds__(256) camera_rays_kernel(const CameraArgs a) {
  __shared__ __align__(16) float s_dir[256 * 3];
  const long long block0 = (long long)blockIdx.x * 256;
  const long long i = block0 + threadIdx.x;
  const bool valid = i < a.count;
  float d[3] = {0.f, 0.f, 0.f}, p[2] = {0.f, 0.f};
  if (valid) camera_ray_of_pixel(a, i, d, p);
  s_dir[threadIdx.x * 3 + 0] = d[0];
  s_dir[threadIdx.x * 3 + 1] = d[1];
  s_dir[threadIdx.x * 3 + 2] = d[2];
  if (valid && a.pixels_out)
    reinterpret_cast<float2*>(a.pixels_out)[i] = make_float2(p[0], p[1]);   // 8 B/thread: coalesced as is
  __syncthreads();
  const long long n_here = min((long long)256, a.count - block0);          // pixels of this block
  if (n_here == 256) {
    // block0 * 12 B is a multiple of 3072 B: 16-byte aligned if the buffer is
    if (threadIdx.x < 192) {
      reinterpret_cast<float4*>(a.directions + block0 * 3)[threadIdx.x] =
          reinterpret_cast<const float4*>(s_dir)[threadIdx.x];
      if (a.origins) {
        // origins repeat (x,y,z): element e of the block's 768 floats is position[e % 3]
        const int e = threadIdx.x * 4;
        const float* q = a.cam.position;
        reinterpret_cast<float4*>(a.origins + block0 * 3)[threadIdx.x] =
            make_float4(q[e % 3], q[(e + 1) % 3], q[(e + 2) % 3], q[(e + 3) % 3]);   // datasets/core.py:66-67
      }
    }
  } else {
    for (int e = threadIdx.x; e < n_here * 3; e += 256) {
      a.directions[block0 * 3 + e] = s_dir[e];
      if (a.origins) a.origins[block0 * 3 + e] = a.cam.position[e % 3];
    }
  }
}

}  // namespace nfb
