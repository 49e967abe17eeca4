// Minimal TMA 3D u8 box-load test (debugging aid for block_match_tma.cu).
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

constexpr int BW = 160, BH = 48;
__global__ void k(const __grid_constant__ CUtensorMap map, uint8_t* out, int x, int y, int z, int step)
{
    extern __shared__ __align__(128) uint8_t sm[];
    uint64_t* bar = reinterpret_cast<uint64_t*>(sm + 4 * 7936);
    uint32_t b32 = (uint32_t)__cvta_generic_to_shared(bar);
    uint32_t dst = (uint32_t)__cvta_generic_to_shared(sm);
    if (threadIdx.x == 0) {
        printf("smem base %u bar %u\n", dst, b32);
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b32));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b32), "r"(4 * BW * BH) : "memory");
        for (int s = 0; s < 4; s++)
            asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                         ::"r"(dst + s * 7936), "l"(&map), "r"(x + s * step), "r"(y), "r"(z), "r"(b32) : "memory");
    }
    __syncthreads();
    uint32_t done = 0;
    while (!done)
        asm volatile("{ .reg .pred q; mbarrier.try_wait.parity.shared::cta.b64 q, [%1], %2; selp.u32 %0, 1, 0, q; }" : "=r"(done) : "r"(b32), "r"(0u) : "memory");
    for (int i = threadIdx.x; i < 4 * 7936; i += blockDim.x) out[i] = sm[i];
}

int main(int argc, char** argv)
{
    const int step = argc > 1 ? atoi(argv[1]) : 1;
    const int W = 1920, H = 1080, P = 2;
    std::vector<uint8_t> h((size_t)W * H * P);
    for (size_t i = 0; i < h.size(); i++) h[i] = (uint8_t)((i * 2654435761u) >> 24);
    uint8_t *d, *o;
    CK(cudaMalloc(&d, h.size())); CK(cudaMalloc(&o, 4 * 7936));
    CK(cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice));
    void* fn = nullptr; cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    auto enc = (PFN_cuTensorMapEncodeTiled)fn;
    CUtensorMap map;
    cuuint64_t dims[3] = {W, H, P}; cuuint64_t strides[2] = {W, (cuuint64_t)W * H};
    cuuint32_t box[3] = {BW, BH, 1}; cuuint32_t es[3] = {1, 1, 1};
    CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode -> %d\n", (int)r);
    CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * 7936 + 64));
    for (int x : {0, -16, 128, 1900}) {
        k<<<1, 256, 4 * 7936 + 64>>>(map, o, x, -16, 1, step);
        CK(cudaDeviceSynchronize());
        std::vector<uint8_t> res(4 * 7936);
        CK(cudaMemcpy(res.data(), o, res.size(), cudaMemcpyDeviceToHost));
        int bad = 0;
        for (int s = 0; s < 4; s++) for (int row = 0; row < BH; row++) for (int c = 0; c < BW; c++) {
            int gx = x + s * step + c, gy = -16 + row;
            uint8_t want = (gx >= 0 && gx < W && gy >= 0 && gy < H) ? h[(size_t)W * H + (size_t)gy * W + gx] : 0;
            if (res[s * 7936 + row * BW + c] != want) bad++;
        }
        printf("x=%d mismatches %d\n", x, bad);
    }
    return 0;
}
