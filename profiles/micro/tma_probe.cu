// Standalone probe: TMA 2-D tile load of an f64 plane and an i32 plane with OOB zero fill.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>
#include <cstdlib>
#define BW 68
#define BH 20
struct Maps { CUtensorMap d; CUtensorMap c; };
struct Tile { double d[BH*BW]; int32_t c[BH*BW]; unsigned long long mbar; };
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void k(const __grid_constant__ Maps maps, int c0, int c1, double* outd, int* outc, int mode) {
    extern __shared__ __align__(128) unsigned char sm[];
    Tile& T = *reinterpret_cast<Tile*>(sm);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&T.mbar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int bytes = ((mode & 1) ? BH*BW*8 : 0) + ((mode & 2) ? BH*BW*4 : 0);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&T.mbar)), "r"(bytes) : "memory");
        if (mode & 1) asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                     :: "r"(s32(T.d)), "l"((unsigned long long)&maps.d), "r"(s32(&T.mbar)), "r"(c0), "r"(c1) : "memory");
        if (mode & 2) asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                     :: "r"(s32(T.c)), "l"((unsigned long long)&maps.c), "r"(s32(&T.mbar)), "r"(c0), "r"(c1) : "memory");
    }
    uint32_t done = 0;
    while (!done) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(s32(&T.mbar)) : "memory");
    }
    for (int q = threadIdx.x; q < BH*BW; q += blockDim.x) { outd[q] = T.d[q]; outc[q] = T.c[q]; }
}
typedef CUresult (*enc_t)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main(int argc, char** argv) {
    int mode = argc > 1 ? atoi(argv[1]) : 3; int cc = argc > 2 ? atoi(argv[2]) : -2; int f64as = argc > 3 ? atoi(argv[3]) : 0;
    int Nx = 100, rows = 50, rp = 100;
    std::vector<double> hd(rp*rows); std::vector<int> hc(rp*rows);
    for (int j = 0; j < rows; j++) for (int i = 0; i < rp; i++) { hd[j*rp+i] = j*1000.0+i; hc[j*rp+i] = j*1000+i+1; }
    double* dd; int* dc; cudaMalloc(&dd, hd.size()*8); cudaMalloc(&dc, hc.size()*4);
    cudaMemcpy(dd, hd.data(), hd.size()*8, cudaMemcpyHostToDevice); cudaMemcpy(dc, hc.data(), hc.size()*4, cudaMemcpyHostToDevice);
    void* fn = nullptr; cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    printf("entry point: %s q=%d fn=%p\n", cudaGetErrorString(e), (int)q, fn);
    enc_t enc = (enc_t)fn;
    Maps m;
    cuuint64_t dims[2] = {(cuuint64_t)Nx, (cuuint64_t)rows}; cuuint32_t box[2] = {BW, BH}, es[2] = {1, 1};
    cuuint64_t sd[1] = {(cuuint64_t)rp*8}, sc[1] = {(cuuint64_t)rp*4};
    CUtensorMapDataType dt64 = f64as == 1 ? CU_TENSOR_MAP_DATA_TYPE_UINT64 : f64as == 2 ? CU_TENSOR_MAP_DATA_TYPE_INT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT64;
    CUresult r1 = enc(&m.d, dt64, 2, dd, dims, sd, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    CUresult r2 = enc(&m.c, CU_TENSOR_MAP_DATA_TYPE_INT32, 2, dc, dims, sc, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode: %d %d\n", (int)r1, (int)r2);
    double* od; int* oc; cudaMalloc(&od, BH*BW*8); cudaMalloc(&oc, BH*BW*4);
    int smem = sizeof(Tile) + 128;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    int c0 = cc, c1 = cc;
    k<<<1, 256, smem>>>(m, c0, c1, od, oc, mode);
    e = cudaDeviceSynchronize();
    printf("kernel: %s\n", cudaGetErrorString(e));
    std::vector<double> rd(BH*BW); std::vector<int> rc(BH*BW);
    cudaMemcpy(rd.data(), od, BH*BW*8, cudaMemcpyDeviceToHost); cudaMemcpy(rc.data(), oc, BH*BW*4, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int j = 0; j < BH; j++) for (int i = 0; i < BW; i++) {
        int gi = c0 + i, gj = c1 + j; bool in = gi >= 0 && gi < Nx && gj >= 0 && gj < rows;
        double ed = in ? gj*1000.0+gi : 0.0; int ec = in ? gj*1000+gi+1 : 0;
        if (rd[j*BW+i] != ed || rc[j*BW+i] != ec) bad++;
    }
    printf("mismatches: %d\n", bad);
    return 0;
}
