// Microbenchmarks of the primitive latencies the ordered-accumulation kernels depend on (sm_100a).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o latency latency.cu && ./latency
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__global__ void k_dadd(double *out, double seed, int n, long long *cyc) {
    double a = seed, b = seed * 0.5;
    long long t0 = clock64();
    for (int i = 0; i < n; i++) { a = __dadd_rn(a, 1.0000001); }
    long long t1 = clock64();
    for (int i = 0; i < n; i++) { a = __dadd_rn(a, 1.0000001); b = __dadd_rn(b, 1.5); }
    long long t2 = clock64();
    out[threadIdx.x] = a + b;
    if (threadIdx.x == 0) { cyc[0] = t1 - t0; cyc[1] = t2 - t1; }
}
__global__ void k_fadd(float *out, float seed, int n, long long *cyc) {
    float a = seed;
    long long t0 = clock64();
    for (int i = 0; i < n; i++) { a = __fadd_rn(a, 1.0000001f); }
    long long t1 = clock64();
    out[threadIdx.x] = a;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_lds_dadd(double *out, int n, long long *cyc) {
    __shared__ double sm[32 * 33];
    for (int i = threadIdx.x; i < 32 * 33; i += blockDim.x) sm[i] = 1.0 + i * 1e-9;
    __syncthreads();
    double a = 0, b = 0;
    const double *row = sm + (threadIdx.x & 31) * 33;
    long long t0 = clock64();
    for (int r = 0; r < n; r++) {
#pragma unroll
        for (int k = 0; k < 32; k++) { a = __dadd_rn(a, row[k]); b = __dadd_rn(b, row[(k + 1) & 31]); }
    }
    long long t1 = clock64();
    out[threadIdx.x] = a + b;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_fns(unsigned *out, unsigned mask, int n, long long *cyc) {
    unsigned acc = 0;
    long long t0 = clock64();
    for (int i = 0; i < n; i++) acc += __fns(mask, 0, (i & 7) + 1);
    long long t1 = clock64();
    out[threadIdx.x] = acc;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_chase(const int *idx, int n, int *out, long long *cyc) {
    int p = 0;
    long long t0 = clock64();
    for (int i = 0; i < n; i++) p = idx[p];
    long long t1 = clock64();
    out[0] = p; cyc[0] = t1 - t0;
}
__global__ void k_null() {}
__global__ void k_hostwrite(volatile float *h) { h[threadIdx.x] = 1.0f; }

int main() {
    double *d; float *f; long long *c, hc[4]; unsigned *u;
    cudaMalloc(&d, 4096); cudaMalloc(&f, 4096); cudaMalloc(&c, 64); cudaMalloc(&u, 4096);
    const int n = 4096;
    for (int rep = 0; rep < 2; rep++) {
        k_dadd<<<1, 32>>>(d, 1.0, n, c); cudaMemcpy(hc, c, 32, cudaMemcpyDeviceToHost);
        if (rep) printf("DADD dependent: %.1f cycles/op; two interleaved chains: %.1f cycles/pair\n", (double) hc[0] / n, (double) hc[1] / n);
        k_fadd<<<1, 32>>>(f, 1.0f, n, c); cudaMemcpy(hc, c, 32, cudaMemcpyDeviceToHost);
        if (rep) printf("FADD dependent: %.1f cycles/op\n", (double) hc[0] / n);
        k_lds_dadd<<<1, 32>>>(d, 64, c); cudaMemcpy(hc, c, 32, cudaMemcpyDeviceToHost);
        if (rep) printf("LDS-fed ordered row sum (2 chains, 32 terms): %.1f cycles/row  (%.1f per term)\n", (double) hc[0] / 64, (double) hc[0] / 64 / 32);
        k_lds_dadd<<<1, 128>>>(d, 64, c); cudaMemcpy(hc, c, 32, cudaMemcpyDeviceToHost);
        if (rep) printf("  same with 4 warps/CTA: %.1f cycles/row\n", (double) hc[0] / 64);
        k_fns<<<1, 32>>>(u, 0x0F0F3355u, n, c); cudaMemcpy(hc, c, 32, cudaMemcpyDeviceToHost);
        if (rep) printf("__fns: %.1f cycles/call\n", (double) hc[0] / n);
    }
    // pointer chase through L2 (16 MB array, stride pattern)
    {
        const int N = 1 << 22; int *h = new int[N]; int *di, *o;
        for (int i = 0; i < N; i++) h[i] = (int) (((long long) i + 40009 * 32) % N);
        cudaMalloc(&di, N * 4); cudaMalloc(&o, 4); cudaMemcpy(di, h, N * 4, cudaMemcpyHostToDevice);
        for (int rep = 0; rep < 2; rep++) { k_chase<<<1, 1>>>(di, 2000, o, c); cudaMemcpy(hc, c, 8, cudaMemcpyDeviceToHost); }
        printf("global load dependent chain (L2-resident after warm-up): %.0f cycles/load\n", (double) hc[0] / 2000);
    }
    // launch + sync round trips
    cudaStream_t st; cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
    float *hm; cudaHostAlloc(&hm, 4096, cudaHostAllocMapped);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int which = 0; which < 2; which++) {
        for (int i = 0; i < 200; i++) { if (which) k_hostwrite<<<1, 32, 0, st>>>(hm); else k_null<<<1, 32, 0, st>>>(); cudaStreamSynchronize(st); }
        cudaEventRecord(e0, st);
        const int R = 2000;
        for (int i = 0; i < R; i++) { if (which) k_hostwrite<<<1, 32, 0, st>>>(hm); else k_null<<<1, 32, 0, st>>>(); cudaStreamSynchronize(st); }
        cudaEventRecord(e1, st); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("launch + cudaStreamSynchronize round trip (%s): %.2f us\n", which ? "kernel writes mapped host memory" : "null kernel", ms * 1e3 / R);
    }
    {   // back-to-back launches without sync: issue rate
        cudaEventRecord(e0, st);
        const int R = 5000;
        for (int i = 0; i < R; i++) k_null<<<1, 32, 0, st>>>();
        cudaEventRecord(e1, st); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("back-to-back null launches: %.2f us each\n", ms * 1e3 / R);
    }
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0); printf("SM clock attr: %d kHz\n", clk);
    return 0;
}
