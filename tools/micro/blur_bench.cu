// Micro-benchmark for the design of the cooperative multi-axis blur (csrc/meanfield.cu, blur_multi_coop_kernel):
// random-neighbour 3-tap gather over float4 items between grid barriers, 6 phases (4 with 273 k items, 2 with 62 k)
// like the keyframe's two lattices.  Variants: U items per thread in flight, neighbour prefetch before the barrier,
// arrive with RED.release instead of fence + ATOM.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o blur_bench blur_bench.cu && ./blur_bench
#include <cstdio>
#include <cuda_runtime.h>
template <bool RED>
__device__ __forceinline__ void bar(unsigned* counter, unsigned target) {
    __syncthreads();
    if (threadIdx.x == 0) {
        if (RED) {
            asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(counter) : "memory");
        } else {
            __threadfence();
            atomicAdd(counter, 1u);
        }
        unsigned v;
        do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory"); } while ((int)(v - target) < 0);
    }
    __syncthreads();
}
__device__ __forceinline__ float4 blur3(float4 o, float4 x, float4 y) {
    o.x += 0.5f * (x.x + y.x); o.y += 0.5f * (x.y + y.y); o.z += 0.5f * (x.z + y.z); o.w += 0.5f * (x.w + y.w);
    return o;
}
// U items per thread, all loads of a phase issued before any use; PF: neighbour pairs of the next phase are loaded before the barrier
template <int U, bool PF, bool RED>
__global__ void __launch_bounds__(1024) k_blur(unsigned* counter, unsigned base, int nphase, float4* a, float4* b, const int2* nbr, int items_hi, int items_lo, int stride) {
    const unsigned tid = blockIdx.x * blockDim.x + threadIdx.x, nthr = gridDim.x * blockDim.x;
    float4* src = a; float4* dst = b;
    int2 nb[U];
    auto load_nb = [&](int j) {
        const int items = j < 4 ? items_hi : items_lo;
#pragma unroll
        for (int u = 0; u < U; u++) { const unsigned it = tid + u * nthr; if (it < items) nb[u] = __ldg(nbr + (size_t)j * stride + it); }
    };
    if (PF) load_nb(0);
    for (int j = 0; j < nphase; j++) {
        const int items = j < 4 ? items_hi : items_lo;
        for (unsigned b0 = 0; b0 < (unsigned)items; b0 += U * nthr) {   // one trip when U * nthr >= items
            if (!PF || b0) {
#pragma unroll
                for (int u = 0; u < U; u++) { const unsigned it = b0 + tid + u * nthr; if (it < items) nb[u] = __ldg(nbr + (size_t)j * stride + it); }
            }
            float4 o[U], x[U], y[U];
#pragma unroll
            for (int u = 0; u < U; u++) { const unsigned it = b0 + tid + u * nthr; if (it < items) { o[u] = __ldcg(src + it); x[u] = __ldcg(src + nb[u].x); y[u] = __ldcg(src + nb[u].y); } }
#pragma unroll
            for (int u = 0; u < U; u++) { const unsigned it = b0 + tid + u * nthr; if (it < items) __stcg(dst + it, blur3(o[u], x[u], y[u])); }
        }
        if (j + 1 < nphase) {
            if (PF) load_nb(j + 1);
            bar<RED>(counter, base + (j + 1) * gridDim.x);
        }
        float4* t = src; src = dst; dst = t;
    }
}
template <int U, bool PF, bool RED>
void run(const char* name, int sms, int block, unsigned* counter, float4* a, float4* b, int2* nbr, int hi, int lo, int stride, int nphase = 6) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int reps = 200; unsigned base = 0;
    cudaMemset(counter, 0, 4); cudaDeviceSynchronize();
    for (int pass = 0; pass < 2; pass++) {
        cudaEventRecord(e0);
        for (int r = 0; r < reps; r++) {
            void* args[] = {&counter, &base, &nphase, &a, &b, &nbr, &hi, &lo, &stride};
            cudaLaunchCooperativeKernel((void*)k_blur<U, PF, RED>, dim3(sms), dim3(block), args, 0, 0);
            base += (nphase - 1) * sms;
        }
        cudaEventRecord(e1); cudaEventSynchronize(e1);
    }
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("%-28s block %4d phases %d: %6.2f us/launch (%s)\n", name, block, nphase, 1000 * ms / reps, cudaGetErrorString(cudaGetLastError()));
}
int main() {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    unsigned* counter; cudaMalloc(&counter, 256);
    const int hi = 273000, lo = 61600, stride = 280000;
    float4 *a, *b; int2* nbr; cudaMalloc(&a, (size_t)stride * 16); cudaMalloc(&b, (size_t)stride * 16); cudaMalloc(&nbr, (size_t)6 * stride * 8);
    cudaMemset(a, 0, (size_t)stride * 16); cudaMemset(b, 0, (size_t)stride * 16);
    int2* h = new int2[(size_t)6 * stride];
    for (int j = 0; j < 6; j++) for (int i = 0; i < stride; i++) { const int n = j < 4 ? hi : lo; h[(size_t)j * stride + i].x = (int)(((long long)i * 7919 + j * 31) % n); h[(size_t)j * stride + i].y = (int)(((long long)i * 104729 + 13 + j) % n); }
    cudaMemcpy(nbr, h, (size_t)6 * stride * 8, cudaMemcpyHostToDevice);
    // fixed cost vs per-phase cost: the same kernel with 0..6 phases (0 = an empty cooperative launch)
    for (int np = 0; np <= 6; np++) run<2, false, false>("U2 atom", sms, 512, counter, a, b, nbr, hi, lo, stride, np);
    for (int block : {128, 256, 512, 1024}) {
        run<1, false, false>("U1 atom", sms, block, counter, a, b, nbr, hi, lo, stride);
        run<2, false, false>("U2 atom", sms, block, counter, a, b, nbr, hi, lo, stride);
        run<4, false, false>("U4 atom", sms, block, counter, a, b, nbr, hi, lo, stride);
        run<2, true, false>("U2 prefetch atom", sms, block, counter, a, b, nbr, hi, lo, stride);
        run<4, true, false>("U4 prefetch atom", sms, block, counter, a, b, nbr, hi, lo, stride);
        run<2, true, true>("U2 prefetch red", sms, block, counter, a, b, nbr, hi, lo, stride);
        run<4, true, true>("U4 prefetch red", sms, block, counter, a, b, nbr, hi, lo, stride);
    }
    return 0;
}
