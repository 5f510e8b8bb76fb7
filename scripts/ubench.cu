// Micro-benchmarks of the FP64 pipe on B200: dependent-issue latency and per-SMSP throughput vs warps.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void lat_dfma(double* out, long long* cyc, int n) {
    double a = threadIdx.x * 1e-9 + 1.0, b = 0.999999, c = 1e-7;
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) { a = fma(a, b, c); a = fma(a, b, c); a = fma(a, b, c); a = fma(a, b, c); }
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    out[blockIdx.x * blockDim.x + threadIdx.x] = a;
}
__global__ void lat_dadd(double* out, long long* cyc, int n) {
    double a = threadIdx.x * 1e-9 + 1.0, c = 1e-7;
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) { a = a + c; a = a + c; a = a + c; a = a + c; }
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    out[blockIdx.x * blockDim.x + threadIdx.x] = a;
}
__global__ void lat_rsq(double* out, long long* cyc, int n) {
    double a = threadIdx.x * 1e-3 + 2.0;
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) { double y; asm volatile("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a)); a = y + 2.0; asm volatile("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a)); a = y + 2.0; }
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    out[blockIdx.x * blockDim.x + threadIdx.x] = a;
}
__global__ void lat_lds(double* out, long long* cyc, int n) {
    __shared__ double s[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) s[i] = (double)((i * 7 + 3) & 1023);
    __syncthreads();
    int idx = threadIdx.x;
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) { idx = (int)s[idx]; idx = (int)s[idx]; }
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    out[blockIdx.x * blockDim.x + threadIdx.x] = idx;
}
// throughput: ILP independent chains per thread, W warps per block, 1 block per SM
template <int ILP>
__global__ void thr_dfma(double* out, long long* cyc, int n) {
    double a[ILP];
    for (int k = 0; k < ILP; ++k) a[k] = threadIdx.x * 1e-9 + k;
    double b = 0.999999, c = 1e-7;
    __syncthreads();
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) {
#pragma unroll
        for (int k = 0; k < ILP; ++k) a[k] = fma(a[k], b, c);
    }
    long long t1 = clock64();
    double s = 0; for (int k = 0; k < ILP; ++k) s += a[k];
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// conversion throughput relative to DFMA: ops per thread per iteration = 8
__global__ void thr_cvt_f2f(double* out, long long* cyc, int n) {
    double a[8]; for (int k = 0; k < 8; ++k) a[k] = threadIdx.x * 1e-3 + k + 0.5;
    float f[8];
    __syncthreads();
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) {
#pragma unroll
        for (int k = 0; k < 8; ++k) { asm volatile("cvt.rn.f32.f64 %0, %1;" : "=f"(f[k]) : "d"(a[k])); }
#pragma unroll
        for (int k = 0; k < 8; ++k) a[k] += 1.0;
    }
    long long t1 = clock64();
    double s = 0; for (int k = 0; k < 8; ++k) s += a[k] + f[k];
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void thr_rint_f64(double* out, long long* cyc, int n) {
    double a[8]; for (int k = 0; k < 8; ++k) a[k] = threadIdx.x * 1e-3 + k + 0.5;
    double f[8];
    __syncthreads();
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) {
#pragma unroll
        for (int k = 0; k < 8; ++k) { asm volatile("cvt.rni.f64.f64 %0, %1;" : "=d"(f[k]) : "d"(a[k])); }
#pragma unroll
        for (int k = 0; k < 8; ++k) a[k] += 1.0;
    }
    long long t1 = clock64();
    double s = 0; for (int k = 0; k < 8; ++k) s += a[k] + f[k];
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void thr_dadd_only(double* out, long long* cyc, int n) {
    double a[8]; for (int k = 0; k < 8; ++k) a[k] = threadIdx.x * 1e-3 + k + 0.5;
    __syncthreads();
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) {
#pragma unroll
        for (int k = 0; k < 8; ++k) a[k] += 1.0;
    }
    long long t1 = clock64();
    double s = 0; for (int k = 0; k < 8; ++k) s += a[k];
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
    double* out; long long* cyc; cudaMalloc(&out, 1 << 24); cudaMalloc(&cyc, 4096);
    long long h[8]; int n = 4096;
    auto rep = [&](const char* name, double per) { cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost); printf("%-28s %8.2f cycles per op\n", name, (double)h[0] / per); };
    lat_dfma<<<1, 32>>>(out, cyc, n); rep("DFMA dependent latency", 4.0 * n);
    lat_dadd<<<1, 32>>>(out, cyc, n); rep("DADD dependent latency", 4.0 * n);
    lat_rsq<<<1, 32>>>(out, cyc, n); rep("RSQ64H+DADD dependent", 2.0 * n);
    lat_lds<<<1, 32>>>(out, cyc, n); rep("LDS.64+F2I dependent", 2.0 * n);
    for (int w : {1, 2, 4, 8, 16, 32}) {
        thr_dfma<1><<<1, 32 * w>>>(out, cyc, n); cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost);
        double c1 = (double)h[0] / n;
        thr_dfma<2><<<1, 32 * w>>>(out, cyc, n); cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost);
        double c2 = (double)h[0] / n;
        thr_dfma<4><<<1, 32 * w>>>(out, cyc, n); cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost);
        double c4 = (double)h[0] / n;
        printf("warps/SM %2d: cycles per loop iter ILP1 %.2f  ILP2 %.2f  ILP4 %.2f   => DFMA warp-instr/clk/SM %.3f %.3f %.3f\n", w, c1, c2, c4, w / c1, 2 * w / c2, 4 * w / c4);
    }
    for (int w : {4, 16}) {
        thr_dadd_only<<<1, 32 * w>>>(out, cyc, n); cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost); double c0 = (double)h[0] / n;
        thr_cvt_f2f<<<1, 32 * w>>>(out, cyc, n); cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost); double c1 = (double)h[0] / n;
        thr_rint_f64<<<1, 32 * w>>>(out, cyc, n); cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost); double c2 = (double)h[0] / n;
        printf("warps/SM %2d: cycles per iter: 8 DADD %.1f | 8 DADD + 8 cvt.f32.f64 %.1f | 8 DADD + 8 cvt.rni.f64.f64 %.1f\n", w, c0, c1, c2);
    }
    cudaError_t e = cudaDeviceSynchronize(); printf("%s\n", cudaGetErrorString(e));
    return 0;
}
