// I-cache probe: loop bodies of N independent-chain DFMA instructions, 4 warps/SMSP.
// Reports warp-instructions per cycle per SM as a function of the loop-body footprint.
#include <cstdio>
#include <cuda_runtime.h>
template <int N>
__global__ void __launch_bounds__(128, 4) body(double* out, int iters) {
    double a[8];
#pragma unroll
    for (int i = 0; i < 8; i++) a[i] = threadIdx.x * 1e-3 + i;
    const double m = 0.999999, c = 1e-7;
#pragma unroll 1
    for (int k = 0; k < iters; k++) {
#pragma unroll
        for (int r = 0; r < N / 8; r++) {
#pragma unroll
            for (int i = 0; i < 8; i++) a[i] = fma(a[i], m, c);
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += a[i];
    if (s == 123.456) out[0] = s;
}
template <int N>
void run(double* d, int sms) {
    int iters = (1 << 22) / N;
    body<N><<<sms * 4, 128>>>(d, 16);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    body<N><<<sms * 4, 128>>>(d, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double winst = (double)sms * 16 * (double)iters * N;   // warp instructions (DFMA only)
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    double cycles = ms * 1e-3 * clk * 1e3;
    printf("body %6d instrs (%7.1f KB): %.3f warp-DFMA/cycle/SM  (%.2f ms)\n", N, N * 16 / 1024.0, winst / cycles / sms, ms);
}
int main() {
    double* d; cudaMalloc(&d, 8);
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount;
    run<64>(d, sms); run<128>(d, sms); run<256>(d, sms); run<384>(d, sms); run<512>(d, sms); run<768>(d, sms);
    run<1024>(d, sms); run<1536>(d, sms); run<2048>(d, sms); run<3072>(d, sms); run<4096>(d, sms); run<8192>(d, sms);
    return 0;
}
