// Issue-port probe: interleave independent DFMA chains with independent integer (LOP3/IADD) or FP32 FFMA
// chains.  If DFMA only occupied the FP64 pipe, DFMA throughput would stay ~2 warp-instr/cycle/SM
// while the extra instructions ride along for free.
#include <cstdio>
#include <cuda_runtime.h>
template <int NI, int NF>  // per 8 DFMA: NI integer ops and NF FFMA ops
__global__ void __launch_bounds__(128, 4) body(double* out, int iters) {
    double a[8]; unsigned b[8]; float f[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { a[i] = threadIdx.x * 1e-3 + i; b[i] = threadIdx.x * 7 + i; f[i] = threadIdx.x * 0.5f + i; }
    const double m = 0.999999, c = 1e-7;
#pragma unroll 1
    for (int k = 0; k < iters; k++) {
#pragma unroll
        for (int r = 0; r < 16; r++) {
#pragma unroll
            for (int i = 0; i < 8; i++) {
                a[i] = fma(a[i], m, c);
                if (i < NI) b[i] = (b[i] ^ (b[i] >> 3)) + 0x9e3779b9u;   // 2-3 integer instrs
                if (i < NF) f[i] = fmaf(f[i], 0.999f, 0.001f);
            }
        }
    }
    double s = 0; unsigned t = 0; float g = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) { s += a[i]; t += b[i]; g += f[i]; }
    if (s == 123.456 || t == 12345u || g == 1.5f) out[0] = s + t + g;
}
template <int NI, int NF>
void run(double* d, int sms) {
    int iters = 1 << 15;
    body<NI, NF><<<sms * 4, 128>>>(d, 16);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    body<NI, NF><<<sms * 4, 128>>>(d, iters);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double dfma = (double)sms * 16 * (double)iters * 128;
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    double cycles = ms * 1e-3 * clk * 1e3;
    printf("per 8 DFMA: %d int-chains, %d FFMA: %.3f warp-DFMA/cycle/SM (%.2f ms)\n", NI, NF, dfma / cycles / sms, ms);
}
int main() {
    double* d; cudaMalloc(&d, 8);
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount;
    run<0, 0>(d, sms); run<2, 0>(d, sms); run<4, 0>(d, sms); run<8, 0>(d, sms);
    run<0, 4>(d, sms); run<0, 8>(d, sms); run<4, 4>(d, sms); run<8, 8>(d, sms);
    return 0;
}
