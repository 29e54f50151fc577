// variants of the TMA box load: argv: <dtype 4|8> <rank 2|3> <boxw> <loc 0=param 1=global> <interleave-less l2promo 0..3>
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>
#include <stdlib.h>
#include <vector>
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void k(const __grid_constant__ CUtensorMap tm, const CUtensorMap *gtm, int loc, int rank, uint32_t bytes, int x0, int y0, int *flag, unsigned char *out)
{
    extern __shared__ __align__(128) unsigned char raw[];
    __shared__ __align__(8) unsigned long long bar;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned long long desc = loc ? reinterpret_cast<unsigned long long>(gtm) : reinterpret_cast<unsigned long long>(&tm);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(bytes) : "memory");
        if (rank == 3)
            asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                         ::"r"(smem_u32(raw)), "l"(desc), "r"(x0), "r"(y0), "r"(0), "r"(smem_u32(&bar)) : "memory");
        else
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                         ::"r"(smem_u32(raw)), "l"(desc), "r"(x0), "r"(y0), "r"(smem_u32(&bar)) : "memory");
    }
    uint32_t done = 0;
    for (uint32_t spin = 0; !done && spin < (1u << 22); ++spin)
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}" : "=r"(done) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
    if (threadIdx.x == 0) *flag = done;
    for (uint32_t t = threadIdx.x; t < bytes; t += blockDim.x) out[t] = raw[t];
}
int main(int argc, char **argv)
{
    const int es = atoi(argv[1]), rank = atoi(argv[2]), boxw = atoi(argv[3]), loc = atoi(argv[4]), promo = atoi(argv[5]);
    const int boxh = 26, nr = 100, rows = 60, pitch = 100; const long plane = (long)pitch * rows;
    std::vector<unsigned char> h(4 * plane * es);
    for (long t = 0; t < 4 * plane; ++t) { if (es == 8) ((double *)h.data())[t] = (double)t; else ((float *)h.data())[t] = (float)t; }
    unsigned char *d, *o; int *flag; CUtensorMap *gtm;
    cudaMalloc(&d, h.size()); cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice);
    const uint32_t bytes = (rank == 3 ? 4u : 1u) * boxh * boxw * es;
    cudaMalloc(&o, bytes); cudaMalloc(&flag, 4); cudaMalloc(&gtm, sizeof(CUtensorMap));
    typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void *fn = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    alignas(64) CUtensorMap tm;
    cuuint64_t gdim[3] = {(cuuint64_t)nr, (cuuint64_t)rows, 4}; cuuint64_t gs[2] = {(cuuint64_t)pitch * es, (cuuint64_t)plane * es};
    cuuint32_t box[3] = {(cuuint32_t)boxw, (cuuint32_t)boxh, 4}, est[3] = {1, 1, 1};
    CUresult r = ((EncodeFn)fn)(&tm, es == 8 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, rank, d, gdim, gs, box, est,
                                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, (CUtensorMapL2promotion)promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    cudaMemcpy(gtm, &tm, sizeof tm, cudaMemcpyHostToDevice);
    const int x0 = argc > 6 ? atoi(argv[6]) : -5; k<<<1, 256, bytes>>>(tm, gtm, loc, rank, bytes, x0, -5, flag, o); printf("x0=%d ", x0);
    cudaError_t e = cudaDeviceSynchronize();
    int hf = -1; cudaMemcpy(&hf, flag, 4, cudaMemcpyDeviceToHost);
    std::vector<unsigned char> ho(bytes); cudaMemcpy(ho.data(), o, bytes, cudaMemcpyDeviceToHost);
    double v = es == 8 ? ((double *)ho.data())[6 * boxw + 7] : ((float *)ho.data())[6 * boxw + 7];
    printf("es=%d rank=%d boxw=%d loc=%d promo=%d encode=%d sync=%d(%s) done=%d val[r6,x7]=%g (want 102)\n", es, rank, boxw, loc, promo, (int)r, (int)e, cudaGetErrorString(e), hf, v);
    return 0;
}
