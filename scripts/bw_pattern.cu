// Micro-benchmark: HBM read bandwidth of the time-tiled access pattern of the temporal pass.  A thread owns VB bytes
// of every frame (frames HW bytes apart) and walks through T frames with K loads in flight; a CTA of NT threads thus
// reads NT*VB contiguous bytes per frame.  Reports GB/s for several (NT, VB, K) and for a plain linear copy-style read.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

template <int VW, int K>
__global__ void walk(const uint32_t *__restrict__ src, size_t hw_words, int T, int groups, uint32_t *out) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= groups) return;
    const uint32_t *p = src + (size_t)g * VW;
    uint32_t buf[K][VW];
    uint32_t acc = 0;
#pragma unroll
    for (int k = 0; k < K; k++)
#pragma unroll
        for (int w = 0; w < VW; w++) buf[k][w] = 0;
#pragma unroll
    for (int k = 0; k < K; k++) {
        if (k < T) {
            if (VW == 4) { uint4 v = __ldg((const uint4 *)(p + (size_t)k * hw_words)); buf[k][0] = v.x; buf[k][1 % VW] = v.y; buf[k][2 % VW] = v.z; buf[k][3 % VW] = v.w; }
            else if (VW == 2) { uint2 v = __ldg((const uint2 *)(p + (size_t)k * hw_words)); buf[k][0] = v.x; buf[k][1 % VW] = v.y; }
            else buf[k][0] = __ldg(p + (size_t)k * hw_words);
        }
    }
    for (int t = 0; t < T; t += K) {
#pragma unroll
        for (int k = 0; k < K; k++) {
#pragma unroll
            for (int w = 0; w < VW; w++) acc = __vmaxu4(acc, buf[k][w]) + 1;
            const int tn = t + k + K;
            if (tn < T) {
                if (VW == 4) { uint4 v = __ldg((const uint4 *)(p + (size_t)tn * hw_words)); buf[k][0] = v.x; buf[k][1 % VW] = v.y; buf[k][2 % VW] = v.z; buf[k][3 % VW] = v.w; }
                else if (VW == 2) { uint2 v = __ldg((const uint2 *)(p + (size_t)tn * hw_words)); buf[k][0] = v.x; buf[k][1 % VW] = v.y; }
                else buf[k][0] = __ldg(p + (size_t)tn * hw_words);
            }
        }
    }
    if (acc == 0x12345678u) out[g] = acc;
}

__global__ void linear(const uint4 *__restrict__ src, size_t n, uint32_t *out) {
    uint32_t acc = 0;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        uint4 v = __ldg(src + i);
        acc += v.x ^ v.y ^ v.z ^ v.w;
    }
    if (acc == 0x12345678u) out[0] = acc;
}

template <int VW, int K>
static int run(const uint32_t *d, size_t hw_bytes, int T, int nt, uint32_t *out, const char *name) {
    const int groups = (int)(hw_bytes / (4 * VW));
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    float best = 1e9f;
    for (int r = 0; r < 4; r++) {
        cudaEventRecord(a);
        walk<VW, K><<<(groups + nt - 1) / nt, nt>>>(d, hw_bytes / 4, T, groups, out);
        cudaEventRecord(b); CK(cudaEventSynchronize(b));
        float ms; cudaEventElapsedTime(&ms, a, b); if (r && ms < best) best = ms;
    }
    printf("%-34s NT=%3d VB=%2d K=%2d : %7.3f ms  %7.0f GB/s\n", name, nt, 4 * VW, K, best, (double)hw_bytes * T / best / 1e6);
    return 0;
}

int main() {
    const size_t HW = 3840ull * 2160; const int T = 541;
    uint32_t *d, *out; CK(cudaMalloc(&d, HW * T)); CK(cudaMalloc(&out, 64 << 20)); CK(cudaMemset(d, 1, HW * T));
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    for (int r = 0; r < 3; r++) {
        cudaEventRecord(a); linear<<<148 * 16, 512>>>((const uint4 *)d, HW * T / 16, out); cudaEventRecord(b); CK(cudaEventSynchronize(b));
        float ms; cudaEventElapsedTime(&ms, a, b); printf("linear read %.3f ms %.0f GB/s\n", ms, (double)HW * T / ms / 1e6);
    }
    run<2, 5>(d, HW, T, 128, out, "walk");
    run<2, 10>(d, HW, T, 128, out, "walk");
    run<2, 16>(d, HW, T, 128, out, "walk");
    run<2, 16>(d, HW, T, 256, out, "walk");
    run<2, 16>(d, HW, T, 512, out, "walk");
    run<2, 32>(d, HW, T, 128, out, "walk");
    run<4, 8>(d, HW, T, 128, out, "walk");
    run<4, 16>(d, HW, T, 128, out, "walk");
    run<4, 16>(d, HW, T, 256, out, "walk");
    run<1, 16>(d, HW, T, 128, out, "walk");
    run<1, 32>(d, HW, T, 256, out, "walk");
    return 0;
}
