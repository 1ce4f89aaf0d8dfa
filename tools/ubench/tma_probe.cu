// tma_probe.cu -- which 2-D TMA forms the strip kernel relies on are legal on sm_100a (run one case per process).
//   nvcc -gencode arch=compute_100a,code=sm_100a -o tma_probe tma_probe.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>
#include <stdlib.h>
__device__ __forceinline__ uint32_t s32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void probe(const __grid_constant__ CUtensorMap map, int c0, int c1, int bytes, int do_store, int sc0, int sc1, uint8_t *out)
{
    extern __shared__ __align__(128) uint8_t buf[];
    __shared__ __align__(8) unsigned long long bar;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(s32(&bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(s32(&bar)), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];\n" ::"r"(s32(buf)),
                     "l"(&map), "r"(c0), "r"(c1), "r"(s32(&bar)) : "memory");
    }
    uint32_t ok = 0;
    for (int i = 0; i < (1 << 22) && !ok; ++i)
        asm volatile("{\n.reg .pred P1;\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\nselp.u32 %0, 1, 0, P1;\n}\n" : "=r"(ok) : "r"(s32(&bar)), "r"(0) : "memory");
    __syncthreads();
    if (threadIdx.x == 0 && !ok) printf("TIMEOUT\n");
    for (int i = threadIdx.x; i < bytes; i += blockDim.x) out[i] = buf[i];
    if (do_store) {
        for (int i = threadIdx.x; i < bytes; i += blockDim.x) buf[i] = (uint8_t)(buf[i] + 1);
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
        __syncthreads();
        if (threadIdx.x == 0) {
            asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%1, %2}], [%3];\n" ::"l"(&map), "r"(sc0), "r"(sc1), "r"(s32(buf)) : "memory");
            asm volatile("cp.async.bulk.commit_group;\ncp.async.bulk.wait_group 0;\n" ::: "memory");
        }
    }
}
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main(int argc, char **argv)
{
    // args: esz dim0_elems rows pitch box0 box1 c0 c1 do_store
    int esz = atoi(argv[1]); long d0 = atol(argv[2]), d1 = atol(argv[3]); size_t pitch = atol(argv[4]);
    int b0 = atoi(argv[5]), b1 = atoi(argv[6]), c0 = atoi(argv[7]), c1 = atoi(argv[8]), do_store = atoi(argv[9]);
    uint8_t *img, *out;
    cudaMalloc(&img, pitch * d1); cudaMalloc(&out, 1 << 20);
    uint8_t *h = (uint8_t *)malloc(pitch * d1);
    for (size_t i = 0; i < pitch * (size_t)d1; ++i) h[i] = (uint8_t)(1 + (i % 251));
    cudaMemcpy(img, h, pitch * d1, cudaMemcpyHostToDevice);
    void *fp = nullptr; cudaDriverEntryPointQueryResult qr;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &qr);
    alignas(64) CUtensorMap map;
    cuuint64_t dims[2] = { (cuuint64_t)d0, (cuuint64_t)d1 }, strides[1] = { pitch };
    cuuint32_t box[2] = { (cuuint32_t)b0, (cuuint32_t)b1 }, es[2] = { 1, 1 };
    CUresult r = ((EncodeTiledFn)fp)(&map, esz == 4 ? CU_TENSOR_MAP_DATA_TYPE_UINT32 : CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, img, dims, strides, box, es,
                                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode rc=%d ", (int)r);
    int bytes = b0 * esz * b1;
    probe<<<1, 128, bytes + 128>>>(map, c0, c1, bytes, do_store, c0, c1, out);
    cudaError_t e = cudaDeviceSynchronize();
    printf("esz=%d dims=%ldx%ld pitch=%zu box=%dx%d at (%d,%d) store=%d -> %s", esz, d0, d1, pitch, b0, b1, c0, c1, do_store, cudaGetErrorString(e));
    if (e == cudaSuccess) {
        uint8_t *o = (uint8_t *)malloc(bytes); cudaMemcpy(o, out, bytes, cudaMemcpyDeviceToHost);
        int bad = 0;
        for (int y = 0; y < b1; ++y) for (int x = 0; x < b0 * esz; ++x) {
            long gx = (long)c0 * esz + x, gy = c1 + y;
            uint8_t want = (gx >= 0 && gx < d0 * esz && gy >= 0 && gy < d1) ? h[gy * pitch + gx] : 0;
            if (o[y * b0 * esz + x] != want) ++bad;
        }
        printf(" load mismatches=%d", bad);
    }
    printf("\n");
    return 0;
}
