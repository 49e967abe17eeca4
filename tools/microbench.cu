// Instruction-throughput microbenchmark for the integer pipes K1 (block matching) lives on.
// Measures lane-ops / clk / SM with clock64() inside one resident CTA per SM.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench tools/microbench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

constexpr int ITER = 4096;
constexpr int ILP = 8;

enum Op { OP_VABSDIFF4_ACC, OP_IDP4A, OP_IADD3, OP_LOP3, OP_SHF, OP_PRMT, OP_IMAD, OP_MIX_VAD_SHF,
          OP_MIX_VAD_IDP, OP_MIX_VAD_IMAD, OP_MIX_VAD_LDS, OP_LDS32, OP_LDS128, OP_VABSDIFF4_NOACC,
          OP_MIX_VAD_PRMT, OP_VIMNMX, OP_COUNT };
static const char* names[] = { "VABSDIFF4.U8.ACC", "IDP.4A.U8.U8", "IADD3", "LOP3", "SHF.R.W (funnel)", "PRMT", "IMAD",
                               "mix VABSDIFF4.ACC + SHF (1:1)", "mix VABSDIFF4.ACC + IDP.4A (1:1)",
                               "mix VABSDIFF4.ACC + IMAD (1:1)", "mix VABSDIFF4.ACC x4 + LDS.32 x1", "LDS.32",
                               "LDS.128", "VABSDIFF4.U8 (no acc)", "mix VABSDIFF4.ACC + PRMT (1:1)", "VIMNMX.U32" };

template <int OP>
__global__ void __launch_bounds__(1024, 1) bench(uint32_t* out, long long* cycles, uint32_t seed)
{
    __shared__ uint32_t sm[4096];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = i * 2654435761u + seed;
    __syncthreads();
    uint32_t acc[ILP], b[ILP];
#pragma unroll
    for (int i = 0; i < ILP; i++) { acc[i] = threadIdx.x * 7 + i + seed; b[i] = threadIdx.x * 13 + i * 5 + seed; }
    uint32_t a = threadIdx.x ^ seed;
    uint32_t lidx = (threadIdx.x * 4) & 4095;
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITER; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) {
            if (OP == OP_VABSDIFF4_ACC) asm volatile("vabsdiff4.u32.u32.u32.add %0, %1, %2, %0;" : "+r"(acc[i]) : "r"(a), "r"(b[i]));
            if (OP == OP_VABSDIFF4_NOACC) asm volatile("vabsdiff4.u32.u32.u32 %0, %0, %1, %2;" : "+r"(acc[i]) : "r"(b[i]), "r"(0));
            if (OP == OP_IDP4A) asm volatile("dp4a.u32.u32 %0, %1, %2, %0;" : "+r"(acc[i]) : "r"(a), "r"(b[i]));
            if (OP == OP_IADD3) asm volatile("add.u32 %0, %0, %1;" : "+r"(acc[i]) : "r"(b[i]));
            if (OP == OP_LOP3) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(acc[i]) : "r"(a), "r"(b[i]));
            if (OP == OP_SHF) asm volatile("shf.r.wrap.b32 %0, %0, %1, %2;" : "+r"(acc[i]) : "r"(b[i]), "r"(8));
            if (OP == OP_PRMT) asm volatile("prmt.b32 %0, %0, %1, %2;" : "+r"(acc[i]) : "r"(b[i]), "r"(0x4321));
            if (OP == OP_IMAD) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(acc[i]) : "r"(a), "r"(b[i]));
            if (OP == OP_VIMNMX) asm volatile("min.u32 %0, %0, %1;" : "+r"(acc[i]) : "r"(b[i]));
            if (OP == OP_MIX_VAD_SHF) {
                asm volatile("vabsdiff4.u32.u32.u32.add %0, %1, %2, %0;" : "+r"(acc[i]) : "r"(a), "r"(b[i]));
                asm volatile("shf.r.wrap.b32 %0, %0, %1, %2;" : "+r"(b[i]) : "r"(a), "r"(8));
            }
            if (OP == OP_MIX_VAD_PRMT) {
                asm volatile("vabsdiff4.u32.u32.u32.add %0, %1, %2, %0;" : "+r"(acc[i]) : "r"(a), "r"(b[i]));
                asm volatile("prmt.b32 %0, %0, %1, %2;" : "+r"(b[i]) : "r"(a), "r"(0x4321));
            }
            if (OP == OP_MIX_VAD_IDP) {
                asm volatile("vabsdiff4.u32.u32.u32.add %0, %1, %2, %0;" : "+r"(acc[i]) : "r"(a), "r"(b[i]));
                asm volatile("dp4a.u32.u32 %0, %1, %2, %0;" : "+r"(b[i]) : "r"(a), "r"(a));
            }
            if (OP == OP_MIX_VAD_IMAD) {
                asm volatile("vabsdiff4.u32.u32.u32.add %0, %1, %2, %0;" : "+r"(acc[i]) : "r"(a), "r"(b[i]));
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(b[i]) : "r"(a), "r"(a));
            }
            if (OP == OP_MIX_VAD_LDS) {
                uint32_t v;
                if ((i & 3) == 0) { asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"((uint32_t)__cvta_generic_to_shared(sm) + ((lidx + i * 128) & 16383))); b[i] ^= v; }
                asm volatile("vabsdiff4.u32.u32.u32.add %0, %1, %2, %0;" : "+r"(acc[i]) : "r"(a), "r"(b[i]));
            }
            if (OP == OP_LDS32) {
                uint32_t v;
                asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"((uint32_t)__cvta_generic_to_shared(sm) + ((lidx + i * 128) & 16383)));
                acc[i] ^= v;
            }
            if (OP == OP_LDS128) {
                uint32_t v0, v1, v2, v3;
                asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v0), "=r"(v1), "=r"(v2), "=r"(v3)
                             : "r"((uint32_t)__cvta_generic_to_shared(sm) + ((lidx * 4 + i * 512) & 16383)));
                acc[i] ^= v0 ^ v1 ^ v2 ^ v3;
            }
        }
    }
    long long t1 = clock64();
    uint32_t r = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) r ^= acc[i] ^ b[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int OP>
int run(int sms, int threads, uint32_t* d_out, long long* d_cyc)
{
    bench<OP><<<sms, threads>>>(d_out, d_cyc, 1);
    CK(cudaDeviceSynchronize());
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    bench<OP><<<sms, threads>>>(d_out, d_cyc, 2);
    cudaEventRecord(e1);
    CK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long cyc[256];
    CK(cudaMemcpy(cyc, d_cyc, sizeof(long long) * sms, cudaMemcpyDeviceToHost));
    double avg = 0; for (int i = 0; i < sms; i++) avg += (double)cyc[i]; avg /= sms;
    int per_iter = (OP == OP_MIX_VAD_SHF || OP == OP_MIX_VAD_IDP || OP == OP_MIX_VAD_IMAD || OP == OP_MIX_VAD_PRMT) ? 2 : 1;
    double instr = (double)ITER * ILP * per_iter * threads;   // lane-instructions per SM
    if (OP == OP_MIX_VAD_LDS) instr = (double)ITER * ILP * threads; // count VABSDIFF4 only
    printf("%-36s threads/SM %4d : %7.2f lane-ops/clk/SM  (%.3f ms, %.0f cyc, eff clk %.0f MHz)\n",
           names[OP], threads, instr / avg, ms, avg, avg / (ms * 1e3));
    return 0;
}

int main()
{
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    int sms = p.multiProcessorCount;
    printf("device %s, %d SMs, cc %d.%d, clock %d kHz, smem/block optin %zu, L2 %d\n", p.name, sms, p.major, p.minor,
           p.clockRate, p.sharedMemPerBlockOptin, p.l2CacheSize);
    uint32_t* d_out; long long* d_cyc;
    CK(cudaMalloc(&d_out, sizeof(uint32_t) * 1024 * sms));
    CK(cudaMalloc(&d_cyc, sizeof(long long) * sms));
    for (int threads : { 128, 256, 512, 1024 }) {
        run<OP_VABSDIFF4_ACC>(sms, threads, d_out, d_cyc);
    }
    int threads = 1024;
    run<OP_VABSDIFF4_NOACC>(sms, threads, d_out, d_cyc);
    run<OP_IDP4A>(sms, threads, d_out, d_cyc);
    run<OP_IADD3>(sms, threads, d_out, d_cyc);
    run<OP_LOP3>(sms, threads, d_out, d_cyc);
    run<OP_SHF>(sms, threads, d_out, d_cyc);
    run<OP_PRMT>(sms, threads, d_out, d_cyc);
    run<OP_IMAD>(sms, threads, d_out, d_cyc);
    run<OP_VIMNMX>(sms, threads, d_out, d_cyc);
    run<OP_MIX_VAD_SHF>(sms, threads, d_out, d_cyc);
    run<OP_MIX_VAD_PRMT>(sms, threads, d_out, d_cyc);
    run<OP_MIX_VAD_IDP>(sms, threads, d_out, d_cyc);
    run<OP_MIX_VAD_IMAD>(sms, threads, d_out, d_cyc);
    run<OP_MIX_VAD_LDS>(sms, threads, d_out, d_cyc);
    run<OP_LDS32>(sms, threads, d_out, d_cyc);
    run<OP_LDS128>(sms, threads, d_out, d_cyc);
    return 0;
}
