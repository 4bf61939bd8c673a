// microbench.cu — latencies that bound the transport kernel's dependent chains on this GPU (one warp, clock64()).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench tools/microbench.cu && tools/microbench
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define N 4096
__global__ void k_dfma(double* out, long long* cyc, double a, double b) {
    double x = a;
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) x = fma(x, b, a);
    long long t1 = clock64();
    out[threadIdx.x] = x; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_dfma2(double* out, long long* cyc, double a, double b) {  // two independent chains
    double x = a, y = b;
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) { x = fma(x, b, a); y = fma(y, a, b); }
    long long t1 = clock64();
    out[threadIdx.x] = x + y; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_dfma4(double* out, long long* cyc, double a, double b) {
    double x = a, y = b, z = a + 1, w = b + 1;
    long long t0 = clock64();
#pragma unroll 8
    for (int i = 0; i < N; i++) { x = fma(x, b, a); y = fma(y, a, b); z = fma(z, b, a); w = fma(w, a, b); }
    long long t1 = clock64();
    out[threadIdx.x] = x + y + z + w; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_imad(uint32_t* out, long long* cyc, uint32_t a) {
    uint32_t x = a + threadIdx.x;
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) { x = __umulhi(x, 0xD2511F53u) ^ (x * 0xCD9E8D57u); }
    long long t1 = clock64();
    out[threadIdx.x] = x; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_lop(uint32_t* out, long long* cyc, uint32_t a) {
    uint32_t x = a + threadIdx.x;
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) { x = (x ^ 0x9E3779B9u) + (x >> 3); }
    long long t1 = clock64();
    out[threadIdx.x] = x; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_rsq(double* out, long long* cyc, double a) {
    double x = a + threadIdx.x;
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) { double s; asm volatile("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(s) : "d"(x)); x = s + 1.5; }
    long long t1 = clock64();
    out[threadIdx.x] = x; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_i2f(double* out, long long* cyc, unsigned long long a) {
    unsigned long long x = a + threadIdx.x;
    double acc = 0;
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) { double d = __ull2double_rn(x >> 11); x = (unsigned long long)__double_as_longlong(d) * 3 + 1; acc += d; }
    long long t1 = clock64();
    out[threadIdx.x] = acc; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_lds(double* out, long long* cyc) {
    __shared__ int s[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) s[i] = (i * 17 + 5) & 1023;
    __syncthreads();
    int x = threadIdx.x;
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) x = s[x];
    long long t1 = clock64();
    out[threadIdx.x] = x; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_ldl(double* out, long long* cyc, int sel) {
    int loc[64];
    for (int i = 0; i < 64; i++) loc[i] = (i * 7 + sel) & 63;
    int x = threadIdx.x & 63;
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) x = loc[x];
    long long t1 = clock64();
    out[threadIdx.x] = x; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
// throughput: many warps of independent DFMA / mixed chains per SM -> achieved DFMA per clock per SM
__global__ void k_tp(double* out, int iters, double a, double b) {
    double x = a + threadIdx.x, y = b, z = a + 1, w = b + 1;
    for (int i = 0; i < iters; i++) { x = fma(x, b, a); y = fma(y, a, b); z = fma(z, b, a); w = fma(w, a, b); }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x + y + z + w;
}
int main() {
    double* d; long long* c; cudaMalloc(&d, 1 << 20); cudaMalloc(&c, 64);
    long long h;
#define RUN(name, call, per) call; cudaDeviceSynchronize(); call; cudaDeviceSynchronize(); cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost); \
    printf("%-28s %7.2f cycles per %s\n", name, (double)h / N, per);
    RUN("DFMA dependent", (k_dfma<<<1, 32>>>(d, c, 1.0000001, 0.9999999)), "op")
    RUN("DFMA 2 chains", (k_dfma2<<<1, 32>>>(d, c, 1.0000001, 0.9999999)), "pair")
    RUN("DFMA 4 chains", (k_dfma4<<<1, 32>>>(d, c, 1.0000001, 0.9999999)), "quad")
    RUN("IMAD.HI+IMAD+LOP dependent", (k_imad<<<1, 32>>>((uint32_t*)d, c, 12345u)), "round")
    RUN("LOP3+SHF+IADD dependent", (k_lop<<<1, 32>>>((uint32_t*)d, c, 12345u)), "step")
    RUN("MUFU.RSQ64H + DADD", (k_rsq<<<1, 32>>>(d, c, 2.0)), "pair")
    RUN("I2F.F64.U64 + IMAD + shift", (k_i2f<<<1, 32>>>(d, c, 0x123456789abcull)), "step")
    RUN("LDS dependent", (k_lds<<<1, 32>>>(d, c)), "load")
    RUN("LDL dependent", (k_ldl<<<1, 32>>>(d, c, 3)), "load")
    // 4 and 8 warps per SMSP of DFMA-only work
    for (int wps : {1, 2, 4, 8}) {
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        int blocks = 148, threads = wps * 4 * 32, iters = 1 << 16;
        k_tp<<<blocks, threads>>>(d, 1000, 1.0000001, 0.9999999);
        cudaEventRecord(e0); k_tp<<<blocks, threads>>>(d, iters, 1.0000001, 0.9999999); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("DFMA x4 chains, %d warps/SMSP: %.2f TFLOP/s\n", wps, 2.0 * 4 * iters * blocks * threads / (ms * 1e-3) / 1e12);
    }
    return 0;
}
