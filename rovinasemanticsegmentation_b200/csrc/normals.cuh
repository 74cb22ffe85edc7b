// PCL IntegralImageNormalEstimation<PointXYZ, Normal> (AVERAGE_3D_GRADIENT, MaxDepthChangeFactor 0.02,
// NormalSmoothingSize 10, BORDER_POLICY_IGNORE) evaluated at one pixel from the prepared distance map and
// double-precision integral images.  Call site in the reference: include/feature_extractor.h:256-261.
// PCL itself is not part of the reference tree; the algorithm follows
// pcl/features/impl/integral_image_normal.hpp (computeFeature / computePointNormal) - parity unpinned.
#pragma once
#include <cuda_runtime.h>

namespace rss {

__host__ __device__ __forceinline__ int skew_pitch(int H) { return (H + 31) & ~31; }
__host__ __device__ __forceinline__ size_t skew_index(int r, int c, int HP) { return (size_t)(r + c) * HP + r; }
__host__ __device__ __forceinline__ size_t skew_elems(int W, int H) { return (size_t)(W + H + 16) * skew_pitch(H); }
// value of an integral image at integral coordinates (rr in [0,H], cc in [0,W]); row 0 / column 0 are zero
template <class T>
__device__ __forceinline__ T integral_at(const T* __restrict__ plane, int rr, int cc, int HP) {
    return (rr == 0 || cc == 0) ? T(0) : plane[skew_index(rr - 1, cc - 1, HP)];
}

// integ: 6 skewed planes [(img*3+ch)] of double, img 0 = d/dx gradients, 1 = d/dy; cnt: 2 skewed planes of int
__device__ __forceinline__ float3 pcl_normal_at(const float4* __restrict__ xyz, const float* __restrict__ dist,
                                                const double* __restrict__ integ, const int* __restrict__ cnt,
                                                int W, int H, int x, int y) {
    const float qnan = __int_as_float(0x7fc00000);
    const float3 bad = make_float3(qnan, qnan, qnan);
    const int border = 10;  // int(normal_smoothing_size_)
    if (x < border || x >= W - border || y < border || y >= H - border) return bad;
    const size_t idx = (size_t)y * W + x;
    if (!isfinite(xyz[idx].z)) return bad;
    const float sm = fminf(dist[idx], 10.0f);
    if (!(sm > 2.0f)) return bad;
    const int w = (int)sm, sx = x - w / 2, sy = y - w / 2;
    const int HP = skew_pitch(H);
    const size_t plane = skew_elems(W, H);
    const int r0 = sy, r1 = sy + w, c0 = sx, c1 = sx + w;  // corners: ul (r0,c0), ur (r0,c1), ll (r1,c0), lr (r1,c1)
    const unsigned cx = (unsigned)integral_at(cnt, r1, c1, HP) + (unsigned)integral_at(cnt, r0, c0, HP) -
                        (unsigned)integral_at(cnt, r0, c1, HP) - (unsigned)integral_at(cnt, r1, c0, HP);
    const unsigned cy = (unsigned)integral_at(cnt + plane, r1, c1, HP) + (unsigned)integral_at(cnt + plane, r0, c0, HP) -
                        (unsigned)integral_at(cnt + plane, r0, c1, HP) - (unsigned)integral_at(cnt + plane, r1, c0, HP);
    if (cx == 0 || cy == 0) return bad;
    double gx[3], gy[3];
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const double* Ix = integ + (size_t)k * plane;
        const double* Iy = integ + (size_t)(3 + k) * plane;
        gx[k] = __dsub_rn(__dsub_rn(__dadd_rn(integral_at(Ix, r1, c1, HP), integral_at(Ix, r0, c0, HP)),
                                    integral_at(Ix, r0, c1, HP)), integral_at(Ix, r1, c0, HP));
        gy[k] = __dsub_rn(__dsub_rn(__dadd_rn(integral_at(Iy, r1, c1, HP), integral_at(Iy, r0, c0, HP)),
                                    integral_at(Iy, r0, c1, HP)), integral_at(Iy, r1, c0, HP));
    }
    // normal = gradient_y.cross(gradient_x)
    const double n0 = __dsub_rn(__dmul_rn(gy[1], gx[2]), __dmul_rn(gy[2], gx[1]));
    const double n1 = __dsub_rn(__dmul_rn(gy[2], gx[0]), __dmul_rn(gy[0], gx[2]));
    const double n2 = __dsub_rn(__dmul_rn(gy[0], gx[1]), __dmul_rn(gy[1], gx[0]));
    const double len = __dadd_rn(__dadd_rn(__dmul_rn(n0, n0), __dmul_rn(n1, n1)), __dmul_rn(n2, n2));
    if (len == 0.0) return bad;
    const double s = sqrt(len);
    return make_float3((float)(n0 / s), (float)(n1 / s), (float)(n2 / s));
}

}  // namespace rss
