// Preparation kernels for PCL-style integral-image normals (see normals.cuh), sm_100a.
//
// Three data products per frame:
//   (1) gradient planes g[6][H][W] (float): central differences of the organised cloud, 0 where PCL skips
//       the element (non-finite), plus finite flags fin[2][H][W];
//   (2) two 3-channel DOUBLE integral images + finite counts, built with PCL's serial recurrence
//       I[r][c+1] = (I[r-1][c+1] + I[r][c]) - I[r-1][c] + e.  Floating-point rounding makes the result depend
//       on that exact evaluation order, so the kernel keeps it and parallelises along anti-diagonals:
//       one CTA per (image, channel) plane, one thread per row, thread r handles column s-r at step s
//       and hands its value to thread r+1 through double-buffered shared memory;
//   (3) PCL's two-pass 1.0/1.4 chamfer distance map over the depth-change mask.  Only min(dist, 10) is
//       ever consumed, and a value < 10 is reached through at most 9 steps, so each pass is computed
//       exactly in independent row bands with a 10-row run-in; within a row the serial dependency
//       cur[c] = min(m[c], cur[c-1] + 1.0f) is resolved by a log-step min-plus scan (x -> x + 1.0f is
//       monotone, so min commutes with it and the float additions happen in the reference order).
#include "kernels.hpp"
#include "normals.cuh"

namespace rss {

// Skewed ("anti-diagonal major") storage: element (r, c) of an H x W plane lives at [(r + c) * HP + r], HP = H
// rounded up to 32.  At wavefront step s = r + c the H threads of a CTA touch one contiguous run of memory.
size_t integral_elems(int W, int H) { return skew_elems(W, H); }

// ------------------------------------------------------------------------------------------------
// (1) gradients (initAverage3DGradientMethod) + depth-change mask -> initial distance map
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool depth_change(float z, float zn) {
    const float thr = __fmul_rn(__fmul_rn(0.02f, __fadd_rn(fabsf(z), 1.0f)), 2.0f);
    return fabsf(__fsub_rn(z, zn)) > thr || !isfinite(z) || !isfinite(zn);
}

// One CTA per 32 x 32 pixel tile.  The cloud tile (+ a one-pixel halo) is staged in shared memory with coalesced loads; the
// gradient planes and finite flags live in the SKEWED layout the wavefront kernel reads, where consecutive rows of one
// anti-diagonal are contiguous - so a warp walks an anti-diagonal of the tile (lane = row) and stores 128 contiguous bytes
// per plane (a thread-per-pixel, row-major kernel scatters every 4-byte store into its own 32-byte sector: 4.5e6 L2
// sectors for 8 MB of output).  Shared-memory reads at lane stride 33 float4 are conflict-free.  The row-major initial
// distance map is written in a second, row-major sweep over the same tile.
constexpr int GM_T = 32;
__global__ void __launch_bounds__(256) gradient_mask_kernel(const float4* __restrict__ xyz, int W, int H,
                                                            float* __restrict__ grad, uint8_t* __restrict__ fin,
                                                            float* __restrict__ dist_init) {
    __shared__ float4 tile[(GM_T + 2) * (GM_T + 2)];
    constexpr int TPITCH = GM_T + 2;
    const int c0 = blockIdx.x * GM_T, r0 = blockIdx.y * GM_T;
    for (int k = threadIdx.x; k < TPITCH * TPITCH; k += 256) {
        const int rr = k / TPITCH, cc = k - rr * TPITCH;
        const int r = r0 - 1 + rr, c = c0 - 1 + cc;
        tile[k] = (r >= 0 && r < H && c >= 0 && c < W) ? xyz[(size_t)r * W + c] : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const size_t SP = skew_elems(W, H);
    const int HP = skew_pitch(H);
    for (int dl = warp; dl < 2 * GM_T - 1; dl += 8) {  // anti-diagonal dl of the tile, lane = local row
        const int rl = lane, cl = dl - rl;
        const int r = r0 + rl, c = c0 + cl;
        if (cl < 0 || cl >= GM_T || r >= H || c >= W) continue;
        const float4* t = tile + (rl + 1) * TPITCH + cl + 1;
        float gx[3] = {0.f, 0.f, 0.f}, gy[3] = {0.f, 0.f, 0.f};
        if (r >= 1 && r < H - 1 && c >= 1 && c < W - 1) {
            const float4 L = t[-1], Rr = t[1], U = t[-TPITCH], Dn = t[TPITCH];
            gx[0] = __fsub_rn(Rr.x, L.x); gx[1] = __fsub_rn(Rr.y, L.y); gx[2] = __fsub_rn(Rr.z, L.z);
            gy[0] = __fsub_rn(Dn.x, U.x); gy[1] = __fsub_rn(Dn.y, U.y); gy[2] = __fsub_rn(Dn.z, U.z);
        }
        const bool fx = isfinite(__fadd_rn(__fadd_rn(gx[0], gx[1]), gx[2]));
        const bool fy = isfinite(__fadd_rn(__fadd_rn(gy[0], gy[1]), gy[2]));
        const size_t si = skew_index(r, c, HP);
#pragma unroll
        for (int k = 0; k < 3; k++) {
            grad[(size_t)k * SP + si] = fx ? gx[k] : 0.f;
            grad[(size_t)(3 + k) * SP + si] = fy ? gy[k] : 0.f;
        }
        fin[si] = fx ? 1 : 0;
        fin[SP + si] = fy ? 1 : 0;
    }
    // depth-change map, gathered: a pixel is cleared by its own tests and by its left / upper neighbour's
    for (int rl = warp; rl < GM_T; rl += 8) {
        const int cl = lane, r = r0 + rl, c = c0 + cl;
        if (r >= H || c >= W) continue;
        const float4* t = tile + (rl + 1) * TPITCH + cl + 1;
        const float z = t[0].z;
        bool cleared = false;
        if (r < H - 1 && c < W - 1) cleared = depth_change(z, t[1].z) || depth_change(z, t[TPITCH].z);
        if (!cleared && c >= 1 && r < H - 1) cleared = depth_change(t[-1].z, z);
        if (!cleared && r >= 1 && c < W - 1) cleared = depth_change(t[-TPITCH].z, z);
        dist_init[(size_t)r * W + c] = cleared ? 0.0f : (float)(W + H);
    }
}

// ------------------------------------------------------------------------------------------------
// (3) chamfer passes.  One CTA per band of rows; shared arrays hold the previous row and the scan buffers.
// ------------------------------------------------------------------------------------------------
constexpr int DIST_RUNIN = 10;
constexpr int DIST_PAD = 16;
constexpr float DIST_BIG = 1.0e30f;

// min-plus scan along a row: v[c] = min_{0<=j<=15} (m[c -/+ j] (+1.0f) j times).  dir = +1 looks left.
// a/b: padded buffers (DIST_PAD entries of DIST_BIG on both sides); result ends in the returned buffer.
template <int DIR>
__device__ __forceinline__ float* minplus_scan(float* a, float* b, int W) {
    float* src = a;
    float* dst = b;
#pragma unroll
    for (int step = 1; step <= 8; step <<= 1) {
        for (int c = threadIdx.x; c < W; c += blockDim.x) {
            float o = src[DIST_PAD + c - DIR * step];
            for (int k = 0; k < step; k++) o = __fadd_rn(o, 1.0f);
            dst[DIST_PAD + c] = fminf(src[DIST_PAD + c], o);
        }
        __syncthreads();
        float* t = src; src = dst; dst = t;
    }
    return src;
}

__global__ void __launch_bounds__(1024) dist_forward_kernel(const float* __restrict__ init, int W, int H, int band,
                                                            float* __restrict__ out) {
    extern __shared__ float sm[];
    float* prev = sm;                        // W
    float* va = prev + W;                    // W + 2*PAD
    float* vb = va + W + 2 * DIST_PAD;       // W + 2*PAD
    const int y0 = blockIdx.x * band, y1 = min(y0 + band, H);
    const int ys = max(y0 - DIST_RUNIN, 0);
    for (int c = threadIdx.x; c < W + 2 * DIST_PAD; c += blockDim.x) va[c] = vb[c] = DIST_BIG;
    for (int c = threadIdx.x; c < W; c += blockDim.x) {
        const float v = init[(size_t)ys * W + c];
        prev[c] = v;
        if (ys >= y0) out[(size_t)ys * W + c] = v;  // row 0 is never updated by the forward pass
    }
    __syncthreads();
    for (int r = ys + 1; r < y1; r++) {
        const float* irow = init + (size_t)r * W;
        for (int c = threadIdx.x; c < W; c += blockDim.x) {
            const float ce = irow[c];
            float m = ce;  // column 0 is never updated
            if (c >= 1) {
                const float ul = __fadd_rn(prev[c - 1], 1.4f), up = __fadd_rn(prev[c], 1.0f);
                // previous_row[ci+1] at ci = W-1 is the first element of the current row (PCL quirk)
                const float ur = __fadd_rn(c + 1 < W ? prev[c + 1] : irow[0], 1.4f);
                m = fminf(ce, fminf(fminf(ul, up), ur));
            }
            va[DIST_PAD + c] = m;
        }
        __syncthreads();
        float* res = minplus_scan<+1>(va, vb, W);
        for (int c = threadIdx.x; c < W; c += blockDim.x) {
            const float v = res[DIST_PAD + c];
            prev[c] = v;
            if (r >= y0) out[(size_t)r * W + c] = v;
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(1024) dist_backward_kernel(const float* __restrict__ fwd, int W, int H, int band,
                                                             float* __restrict__ out) {
    extern __shared__ float sm[];
    float* next = sm;
    float* va = next + W;
    float* vb = va + W + 2 * DIST_PAD;
    const int y0 = blockIdx.x * band, y1 = min(y0 + band, H);
    const int ye = min(y1 - 1 + DIST_RUNIN, H - 1);
    for (int c = threadIdx.x; c < W + 2 * DIST_PAD; c += blockDim.x) va[c] = vb[c] = DIST_BIG;
    for (int c = threadIdx.x; c < W; c += blockDim.x) {
        const float v = fwd[(size_t)ye * W + c];
        next[c] = v;
        if (ye < y1) out[(size_t)ye * W + c] = v;  // row H-1 is never updated by the backward pass
    }
    __syncthreads();
    for (int r = ye - 1; r >= y0; r--) {
        const float* frow = fwd + (size_t)r * W;
        for (int c = threadIdx.x; c < W; c += blockDim.x) {
            const float ce = frow[c];
            float m = ce;  // column W-1 is never updated
            if (c < W - 1) {
                // next_row[ci-1] at ci = 0 is the last element of the current row (PCL quirk)
                const float ll = __fadd_rn(c >= 1 ? next[c - 1] : frow[W - 1], 1.4f);
                const float lo = __fadd_rn(next[c], 1.0f), lr = __fadd_rn(next[c + 1], 1.4f);
                m = fminf(ce, fminf(fminf(ll, lo), lr));
            }
            va[DIST_PAD + c] = m;
        }
        __syncthreads();
        float* res = minplus_scan<-1>(va, vb, W);
        for (int c = threadIdx.x; c < W; c += blockDim.x) {
            const float v = res[DIST_PAD + c];
            next[c] = v;
            if (r < y1) out[(size_t)r * W + c] = v;
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------
// (2) integral images along anti-diagonals.  grid = 6 CTAs (image x channel), block = H threads (<= 1024).
// Thread r owns image row r.  At step s it computes column c = s - r:
//     v = ((I[r][c+1] + I[r+1][c]) - I[r][c]) + g      (g = 0 where PCL skips the element)
// "up" comes from thread r-1 (shared memory, written one step earlier), "left"/"upleft" stay in registers.
// Inputs and outputs use the skewed layout, so each step is one coalesced load and one coalesced store per warp;
// the gradient values are prefetched WF_PF steps ahead.  The recurrence itself is three dependent DADDs per step.
// ------------------------------------------------------------------------------------------------
constexpr int WF_PF = 8;
// grid = 8 CTAs: planes 0..5 = the six double integral images, 6..7 = the two finite-count images (int, exact).
template <int MAXT, typename T, typename TIn>
__device__ __forceinline__ void wavefront_plane(const TIn* __restrict__ g, T* __restrict__ I, T* sval, int W, int H) {
    const int r = threadIdx.x;
    const int HP = skew_pitch(H);
    for (int k = threadIdx.x; k < 2 * (H + 1); k += blockDim.x) sval[k] = T(0);
    __syncthreads();
    T left = T(0), upleft = T(0);
    const int steps = W + H - 1;
    const bool live = r < H;
    TIn cur[WF_PF], nxt[WF_PF];
    const TIn* gp = g + r;          // advances by HP per step
    T* ip = I + r;
#pragma unroll
    for (int k = 0; k < WF_PF; k++) cur[k] = live ? __ldg(gp + (size_t)k * HP) : TIn(0);
    gp += (size_t)WF_PF * HP;
    T* rd0 = sval + r;              // buffer 0, slot of thread r-1 (index r)
    T* rd1 = sval + (H + 1) + r;    // buffer 1
    for (int s0 = 0; s0 < steps; s0 += WF_PF) {
#pragma unroll
        for (int k = 0; k < WF_PF; k++) nxt[k] = live ? __ldg(gp + (size_t)k * HP) : TIn(0);  // padded by 16 steps
        gp += (size_t)WF_PF * HP;
#pragma unroll
        for (int k = 0; k < WF_PF; k++) {
            const int s = s0 + k;
            const int c = s - r;
            // step s reads the buffer written at step s-1 (parity (s+1)&1) and writes parity s&1; WF_PF is even, so the
            // parity of s equals the parity of k and the buffer choice is a compile-time constant.
            T* rd = (k & 1) ? rd0 : rd1;
            T* wr = (k & 1) ? rd1 : rd0;
            if (live && c >= 0 && c < W) {
                const T up = *rd;
                T v;
                if constexpr (sizeof(T) == 8) v = __dadd_rn(__dsub_rn(__dadd_rn(up, left), upleft), (double)cur[k]);
                else v = up + left - upleft + (T)cur[k];
                *ip = v;
                wr[1] = v;
                upleft = up;
                left = v;
            }
            ip += HP;
            __syncthreads();
        }
#pragma unroll
        for (int k = 0; k < WF_PF; k++) cur[k] = nxt[k];
    }
}
template <int MAXT>
__global__ void __launch_bounds__(MAXT) integral_wavefront_kernel(const float* __restrict__ grad,
                                                                  const uint8_t* __restrict__ fin, int W, int H,
                                                                  double* __restrict__ integ, int* __restrict__ cnt) {
    extern __shared__ double smd[];  // [2][H+1] doubles (or ints)
    const size_t SP = skew_elems(W, H);
    const int plane_id = blockIdx.x;
    if (plane_id < 6)
        wavefront_plane<MAXT, double, float>(grad + (size_t)plane_id * SP, integ + (size_t)plane_id * SP, smd, W, H);
    else
        wavefront_plane<MAXT, int, uint8_t>(fin + (size_t)(plane_id - 6) * SP, cnt + (size_t)(plane_id - 6) * SP,
                                            reinterpret_cast<int*>(smd), W, H);
}

void launch_normals_prepare(rss_ctx* c, cudaStream_t st, const float4* xyz, int W, int H, float* dist_a,
                            float* dist_b, double* integ, int* integ_cnt, float* grad, uint8_t* fin) {
    dim3 grid(rss_div_up(W, GM_T), rss_div_up(H, GM_T));
    RSS_LAUNCH(c, gradient_mask_kernel, grid, 256, 0, st, xyz, W, H, grad, fin, dist_b);
    // distance map: init (dist_b) -> forward (dist_a) -> backward (dist_b)
    const int band = 8;
    const int threads = min(1024, rss_div_up(W, 32) * 32);
    const size_t smem = (size_t)(3 * W + 4 * DIST_PAD) * sizeof(float);
    RSS_LAUNCH(c, dist_forward_kernel, rss_div_up(H, band), threads, smem, st, dist_b, W, H, band, dist_a);
    RSS_LAUNCH(c, dist_backward_kernel, rss_div_up(H, band), threads, smem, st, dist_a, W, H, band, dist_b);
    const int wt = min(1024, rss_div_up(H, 32) * 32);
    const size_t wsmem = (size_t)2 * (H + 1) * sizeof(double);
    if (wt <= 512) RSS_LAUNCH(c, integral_wavefront_kernel<512>, 8, wt, wsmem, st, grad, fin, W, H, integ, integ_cnt);
    else RSS_LAUNCH(c, integral_wavefront_kernel<1024>, 8, wt, wsmem, st, grad, fin, W, H, integ, integ_cnt);
}

__global__ void __launch_bounds__(256) normals_full_kernel(const float4* __restrict__ xyz,
                                                           const float* __restrict__ dist,
                                                           const double* __restrict__ integ,
                                                           const int* __restrict__ cnt, int W, int H,
                                                           float* __restrict__ normals) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= W) return;
    const float3 n = pcl_normal_at(xyz, dist, integ, cnt, W, H, x, y);
    float* o = normals + ((size_t)y * W + x) * 3;
    o[0] = n.x; o[1] = n.y; o[2] = n.z;
}
void launch_normals_full(rss_ctx* c, cudaStream_t st, const float4* xyz, const float* dist, const double* integ,
                         const int* integ_cnt, int W, int H, float* normals) {
    dim3 grid(rss_div_up(W, 256), H);
    RSS_LAUNCH(c, normals_full_kernel, grid, 256, 0, st, xyz, dist, integ, integ_cnt, W, H, normals);
}

}  // namespace rss
