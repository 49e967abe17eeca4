// Does cuTensorMapEncodeTiled accept a "transposing" u8 tensor — dims {16 bytes, rows, 16-byte column blocks} with byte
// strides {row stride, 16} (not ascending) — and does the box land as [colblk][row][16]?   nvcc -arch=sm_100a -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <vector>

__global__ void k(const __grid_constant__ CUtensorMap map, uint8_t* out, int x16, int y)
{
    __shared__ __align__(128) uint8_t buf[8 * 64 * 16];
    __shared__ __align__(8) uint64_t bar;
    if (threadIdx.x == 0) {
        uint32_t b32 = (uint32_t)__cvta_generic_to_shared(&bar);
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b32));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b32), "r"(8 * 64 * 16) : "memory");
        asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
                     ::"r"((uint32_t)__cvta_generic_to_shared(buf)), "l"(&map), "r"(0), "r"(y), "r"(x16), "r"(0), "r"(b32) : "memory");
        uint32_t done = 0;
        while (!done)
            asm volatile("{ .reg .pred q; mbarrier.try_wait.parity.shared::cta.b64 q, [%1], %2; selp.u32 %0, 1, 0, q; }" : "=r"(done) : "r"(b32), "r"(0) : "memory");
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 8 * 64 * 16; i += blockDim.x) out[i] = buf[i];
}

int main()
{
    const int W = 1920, H = 200;
    std::vector<uint8_t> h((size_t)W * H);
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++) h[(size_t)y * W + x] = (uint8_t)((x * 7 + y * 13) & 255);
    uint8_t *d, *o;
    cudaMalloc(&d, h.size());
    cudaMalloc(&o, 8 * 64 * 16);
    cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice);
    cuInit(0);
    CUtensorMap map;
    cuuint64_t dims[4] = {16, (cuuint64_t)H, (cuuint64_t)(W / 16), 1};
    cuuint64_t strides[3] = {(cuuint64_t)W, 16, (cuuint64_t)W * H};
    cuuint32_t box[4] = {16, 64, 8, 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = cuTensorMapEncodeTiled(&map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 4, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                        CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode: %d\n", (int)r);
    if (r != CUDA_SUCCESS) return 1;
    const int x16 = 3, y0 = 150;   // rows 150..213: 200.. are out of the tensor -> zeros
    k<<<1, 128>>>(map, o, x16, y0);
    cudaError_t e = cudaDeviceSynchronize();
    printf("kernel: %s\n", cudaGetErrorString(e));
    std::vector<uint8_t> got(8 * 64 * 16);
    cudaMemcpy(got.data(), o, got.size(), cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int c = 0; c < 8; c++)
        for (int rr = 0; rr < 64; rr++)
            for (int b = 0; b < 16; b++) {
                const int y = y0 + rr, x = (x16 + c) * 16 + b;
                const uint8_t want = y < H ? h[(size_t)y * W + x] : 0;
                bad += got[(c * 64 + rr) * 16 + b] != want;
            }
    printf("layout [colblk][row][16]: %d mismatches\n", bad);
    return bad != 0;
}
