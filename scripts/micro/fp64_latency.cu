// dependent-chain latency of the FP64 ops on the CUSUM critical path (one warp): DADD, DADD+DSETP+FSEL (the clamp), full step
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_dadd(double *out, double x, int iters) {
    double a = out[0];
    long long t0 = clock64();
    for (int i = 0; i < iters; i++) { a = __dadd_rn(a, x); a = __dadd_rn(a, x); a = __dadd_rn(a, x); a = __dadd_rn(a, x); }
    long long t1 = clock64();
    out[0] = a; if (threadIdx.x == 0) out[1] = (double)(t1 - t0) / (4.0 * iters);
}
__global__ void k_clamp(double *out, double x, int iters) {
    double a = out[0];
    long long t0 = clock64();
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int q = 0; q < 4; q++) { double b = __dadd_rn(a, x); a = b > 0.0 ? b : 0.0; }
    }
    long long t1 = clock64();
    out[0] = a; if (threadIdx.x == 0) out[2] = (double)(t1 - t0) / (4.0 * iters);
}
__global__ void k_step(double *out, double x, double lam, int iters) {
    double sp = out[0], sn = -out[0];
    unsigned w = 0;
    long long t0 = clock64();
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const double a = __dadd_rn(sp, x), b = __dadd_rn(sn, x);
            const double p = a > 0.0 ? a : 0.0, n = b < 0.0 ? b : 0.0;
            const bool hp = p >= lam, hn = !hp && (n <= -lam);
            sp = hp ? 0.0 : p; sn = hn ? 0.0 : n;
            w += hp | hn;
        }
    }
    long long t1 = clock64();
    out[0] = sp + sn + w; if (threadIdx.x == 0) out[3] = (double)(t1 - t0) / (4.0 * iters);
}
__global__ void k_fadd(double *out, float x, int iters) {
    float a = (float)out[0];
    long long t0 = clock64();
    for (int i = 0; i < iters; i++) { a = __fadd_rn(a, x); a = __fadd_rn(a, x); a = __fadd_rn(a, x); a = __fadd_rn(a, x); }
    long long t1 = clock64();
    out[0] = a; if (threadIdx.x == 0) out[4] = (double)(t1 - t0) / (4.0 * iters);
}
int main() {
    double *d; cudaMalloc(&d, 64); cudaMemset(d, 0, 64);
    for (int threads = 32; threads <= 1024; threads *= 4) {
        k_dadd<<<1, threads>>>(d, 1e-9, 100000); k_clamp<<<1, threads>>>(d, 1e-9, 100000); k_step<<<1, threads>>>(d, 1e-9, 0.5, 100000);
        k_fadd<<<1, threads>>>(d, 1e-9f, 100000);
        double h[8]; cudaMemcpy(h, d, 64, cudaMemcpyDeviceToHost);
        printf("threads %4d: cycles per dependent DADD %.1f | DADD+clamp %.1f | full cusum step %.1f | FADD %.1f\n", threads, h[1], h[2], h[3], h[4]);
    }
    return 0;
}
