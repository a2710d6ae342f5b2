// Issue-rate probe (sm_100a): cycles per warp instruction of the ops the attention softmax is made of, one and two warps per
// scheduler.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 tools/mufu_probe.cu -o tools/build/mufu_probe && tools/build/mufu_probe
#include <cstdio>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
template <int OP>
__global__ void probe(float* out, unsigned long long* cyc, int iters) {
    float a[16];
    unsigned u[16];
    unsigned long long p[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) { a[i] = -0.001f * (threadIdx.x + i); u[i] = 0xb800b400u + i + threadIdx.x; p[i] = ((unsigned long long)__float_as_uint(a[i]) << 32) | __float_as_uint(a[i] * 0.5f); }
    __syncthreads();
    const unsigned long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            if (OP == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
            if (OP == 1) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(u[i]));
            if (OP == 2) asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(u[i]));
            if (OP == 3) asm volatile("fma.rn.f32x2 %0, %0, %0, %0;" : "+l"(p[i]));
            if (OP == 4) asm volatile("add.rn.f32x2 %0, %0, %0;" : "+l"(p[i]));
            if (OP == 5) asm volatile("fma.rn.f32 %0, %0, %0, %0;" : "+f"(a[i]));
            if (OP == 6) asm volatile("add.rn.f16x2 %0, %0, %0;" : "+r"(u[i]));
            if (OP == 7) asm volatile("{ .reg .b16 lo, hi; mov.b32 {lo, hi}, %1; cvt.f32.f16 %0, lo; }" : "=f"(a[i]) : "r"(u[i]));
            if (OP == 8) asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(u[i]) : "f"(a[i]), "f"(a[(i + 1) & 15]));
            if (OP == 9) asm volatile("tanh.approx.f32 %0, %0;" : "+f"(a[i]));
        }
    }
    const unsigned long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += a[i] + __uint_as_float(u[i]) + (float)(p[i] & 0xff);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
int main() {
    float* out; unsigned long long* cyc;
    cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 8);
    const char* names[] = {"ex2.approx.ftz.f32", "ex2.approx.ftz.f16x2", "ex2.approx.ftz.bf16x2", "fma.rn.f32x2", "add.rn.f32x2", "fma.rn.f32", "add.rn.f16x2", "cvt.f32.f16 (lo half)", "cvt.rn.f16x2.f32", "tanh.approx.f32"};
    const int iters = 2000;
    for (int warps_per_smsp = 1; warps_per_smsp <= 2; ++warps_per_smsp)
        for (int op = 0; op < 10; ++op) {
            const int threads = 128 * warps_per_smsp;
            unsigned long long h = 0;
#define RUN(O) case O: probe<O><<<148, threads>>>(out, cyc, iters); break;
            for (int rep = 0; rep < 2; ++rep) {
                switch (op) { RUN(0) RUN(1) RUN(2) RUN(3) RUN(4) RUN(5) RUN(6) RUN(7) RUN(8) RUN(9) }
                cudaDeviceSynchronize();
            }
            cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
            printf("%d warp(s)/scheduler  %-24s %6.2f cycles per warp instruction (%s)\n", warps_per_smsp, names[op],
                   (double)h / (iters * 16.0), cudaGetErrorString(cudaGetLastError()));
        }
    return 0;
}
