// Microbenchmark: issue/pipe throughput of packed FFMA2/FADD2 vs scalar FFMA/FADD on sm_100a.
// Reports thread-level FMA lanes per clock per SM and warp-instructions per clock per SMSP.
#include <cstdio>
#include <cuda_runtime.h>
#define NCH 8
template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, int iters, float c0, float c1) {
    float2 a[NCH];
    int ia[4] = {(int)threadIdx.x, 3, 5, 7};
    for (int i = 0; i < NCH; ++i) a[i] = make_float2(threadIdx.x * 1e-3f + i, i * 0.5f);
    const float2 m = make_float2(c0, c0), ad = make_float2(c1, c1);
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
#pragma unroll
            for (int i = 0; i < NCH; ++i) {
                if (MODE == 0) {  // scalar FFMA x2
                    a[i].x = fmaf(a[i].x, c0, c1);
                    a[i].y = fmaf(a[i].y, c0, c1);
                } else if (MODE == 1) {  // FFMA2
                    a[i] = __ffma2_rn(a[i], m, ad);
                } else if (MODE == 2) {  // FADD2
                    a[i] = __fadd2_rn(a[i], ad);
                } else if (MODE == 3) {  // scalar FADD x2
                    a[i].x += c1;
                    a[i].y += c1;
                } else if (MODE == 4) {  // FFMA2 + 2 integer ALU ops per packed op
                    a[i] = __ffma2_rn(a[i], m, ad);
                    ia[i & 3] = (ia[i & 3] ^ (ia[(i + 1) & 3] >> 1)) + i;
                } else if (MODE == 5) {  // 2 scalar FFMA + same ALU ops
                    a[i].x = fmaf(a[i].x, c0, c1);
                    a[i].y = fmaf(a[i].y, c0, c1);
                    ia[i & 3] = (ia[i & 3] ^ (ia[(i + 1) & 3] >> 1)) + i;
                } else if (MODE == 6) {  // FMUL2
                    a[i] = __fmul2_rn(a[i], m);
                }
            }
        }
    }
    float s = 0;
    for (int i = 0; i < NCH; ++i) s += a[i].x + a[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + ia[0] + ia[1] + ia[2] + ia[3];
}
template <int MODE>
void run(const char* name, int blocks_per_sm) {
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    int clk = 0;
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    float* out;
    cudaMalloc(&out, sizeof(float) * sms * blocks_per_sm * 256);
    const int iters = 20000;
    k<MODE><<<sms * blocks_per_sm, 256>>>(out, 1000, 1.0001f, 0.5f);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<MODE><<<sms * blocks_per_sm, 256>>>(out, iters, 1.0001f, 0.5f);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    // per thread per iteration: 4*NCH "pair-ops" (each = 2 lane flops-ops)
    double pairops = (double)iters * 4 * NCH * 256.0 * blocks_per_sm;  // per SM, thread-level
    double cycles = ms * 1e-3 * clk * 1e3;
    printf("%-34s occ %d CTA/SM: %.3f ms  pair-ops/clk/SM %.1f (lane-ops %.1f)  [clk %d kHz nominal]\n", name,
           blocks_per_sm, ms, pairops / cycles, 2 * pairops / cycles, clk);
    cudaFree(out);
}
int main() {
    for (int occ = 2; occ <= 4; occ += 2) {
        run<0>("scalar FFMA x2", occ);
        run<1>("FFMA2", occ);
        run<3>("scalar FADD x2", occ);
        run<2>("FADD2", occ);
        run<6>("FMUL2", occ);
        run<5>("2 FFMA + 3 int ALU", occ);
        run<4>("FFMA2 + 3 int ALU", occ);
    }
    return 0;
}
