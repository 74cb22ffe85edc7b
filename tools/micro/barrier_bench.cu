// Micro-benchmark: cost of a grid barrier built on an L2 counter (cooperative launch, one CTA per SM).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o barrier_bench barrier_bench.cu && ./barrier_bench
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void bar_poll(unsigned* counter, unsigned target) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(counter, 1u);
        unsigned v;
        do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory"); } while ((int)(v - target) < 0);
    }
    __syncthreads();
}
__device__ __forceinline__ void bar_flag(unsigned* counter, unsigned* flag, unsigned target, unsigned gen) {
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned prev;
        asm volatile("atom.add.release.gpu.global.u32 %0, [%1], 1;" : "=r"(prev) : "l"(counter) : "memory");
        if (prev == target - 1) {
            asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(flag), "r"(gen) : "memory");
        } else {
            unsigned v;
            do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory"); } while ((int)(v - gen) < 0);
        }
    }
    __syncthreads();
}
__global__ void k_poll(unsigned* counter, unsigned base, int nbar, float* sink) {
    for (int j = 0; j < nbar; j++) bar_poll(counter, base + (j + 1) * gridDim.x);
    if (sink && threadIdx.x == 9999) sink[0] = 1;
}
__global__ void k_flag(unsigned* counter, unsigned* flag, unsigned base, unsigned gen0, int nbar) {
    for (int j = 0; j < nbar; j++) bar_flag(counter, flag, base + (j + 1) * gridDim.x, gen0 + j + 1);
}
// barrier + a dependent L2 load/store phase like the blur
__global__ void k_work(unsigned* counter, unsigned* flag, unsigned base, unsigned gen0, int nbar, float4* a, float4* b, const int2* nbr, int items) {
    const unsigned tid = blockIdx.x * blockDim.x + threadIdx.x, nthr = gridDim.x * blockDim.x;
    float4* src = a; float4* dst = b;
    for (int j = 0; j < nbar + 1; j++) {
        for (unsigned it = tid; it < items; it += nthr) {
            int2 nb = __ldg(nbr + it);
            float4 o = __ldcg(src + it), x = __ldcg(src + nb.x), y = __ldcg(src + nb.y);
            o.x += 0.5f * (x.x + y.x); o.y += 0.5f * (x.y + y.y); o.z += 0.5f * (x.z + y.z); o.w += 0.5f * (x.w + y.w);
            __stcg(dst + it, o);
        }
        if (j < nbar) bar_flag(counter, flag, base + (j + 1) * gridDim.x, gen0 + j + 1);
        float4* t = src; src = dst; dst = t;
    }
}
int main() {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    unsigned *counter, *flag; cudaMalloc(&counter, 256); cudaMalloc(&flag, 256); cudaMemset(counter, 0, 256); cudaMemset(flag, 0, 256);
    const int items = 257000; float4 *a, *b; int2* nbr; cudaMalloc(&a, items * 16); cudaMalloc(&b, items * 16); cudaMalloc(&nbr, items * 8);
    cudaMemset(a, 0, items * 16); cudaMemset(b, 0, items * 16);
    int2* h = new int2[items]; for (int i = 0; i < items; i++) { h[i].x = (i * 7919) % items; h[i].y = (i * 104729 + 13) % items; }
    cudaMemcpy(nbr, h, items * 8, cudaMemcpyHostToDevice);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int reps = 200;
    for (int block : {128, 256, 512}) for (int nbar : {0, 5, 20}) {
        unsigned base = 0; float* sink = nullptr;
        // reset
        cudaMemset(counter, 0, 4); cudaDeviceSynchronize();
        cudaEventRecord(e0);
        for (int r = 0; r < reps; r++) { void* args[] = {&counter, &base, (void*)&nbar, &sink}; cudaLaunchCooperativeKernel((void*)k_poll, dim3(sms), dim3(block), args, 0, 0); base += nbar * sms; }
        cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("poll  block %3d nbar %2d: %.2f us/launch\n", block, nbar, 1000 * ms / reps);
        cudaMemset(counter, 0, 4); cudaMemset(flag, 0, 4); cudaDeviceSynchronize(); base = 0; unsigned gen = 0;
        cudaEventRecord(e0);
        for (int r = 0; r < reps; r++) { void* args[] = {&counter, &flag, &base, &gen, (void*)&nbar}; cudaLaunchCooperativeKernel((void*)k_flag, dim3(sms), dim3(block), args, 0, 0); base += nbar * sms; gen += nbar; }
        cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
        printf("flag  block %3d nbar %2d: %.2f us/launch\n", block, nbar, 1000 * ms / reps);
        cudaMemset(counter, 0, 4); cudaMemset(flag, 0, 4); cudaDeviceSynchronize(); base = 0; gen = 0;
        cudaEventRecord(e0);
        for (int r = 0; r < reps; r++) { int it = items; void* args[] = {&counter, &flag, &base, &gen, (void*)&nbar, &a, &b, &nbr, &it}; cudaLaunchCooperativeKernel((void*)k_work, dim3(sms), dim3(block), args, 0, 0); base += nbar * sms; gen += nbar; }
        cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
        printf("work  block %3d nbar %2d: %.2f us/launch (%s)\n", block, nbar, 1000 * ms / reps, cudaGetErrorString(cudaGetLastError()));
    }
    // plain (non-cooperative) empty kernel launch cost for reference
    cudaEventRecord(e0);
    for (int r = 0; r < reps; r++) { unsigned base = 0; k_poll<<<sms, 256>>>(counter, base, 0, nullptr); }
    cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("plain empty launch: %.2f us/launch\n", 1000 * ms / reps);
    return 0;
}
