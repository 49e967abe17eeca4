// TEST INFRASTRUCTURE — never part of libofps_b200.so.
//
// A minimal CPU stand-in for the CUDA execution model, just large enough to run the kernels of
// ofps_b200/csrc/cv_front.cu unmodified on the build container (which has no GPU): one OS thread per CUDA
// thread of a CTA, CTAs one after another, `__shared__` as function-local statics, __syncthreads and the warp
// collectives as barriers.  It exists so that the index arithmetic, barriers and border handling of a kernel
// can be checked against the oracle BEFORE GPU time is spent (tests/test_emu_cv_front.py).  It says nothing
// about performance and is not a fallback: the product library refuses to run without an sm_100 device.
#pragma once

#include <algorithm>
#include <atomic>
#include <barrier>
#include <cmath>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <thread>
#include <vector>

// ---- types ---------------------------------------------------------------------------------------
struct uint3 { unsigned x, y, z; };
struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct float2 { float x, y; };
struct alignas(16) float4 { float x, y, z, w; };
struct uint2 { unsigned x, y; };
struct alignas(16) uint4 { unsigned x, y, z, w; };
inline float2 make_float2(float x, float y) { return {x, y}; }
inline float4 make_float4(float x, float y, float z, float w) { return {x, y, z, w}; }
inline uint2 make_uint2(unsigned x, unsigned y) { return {x, y}; }
inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return {x, y, z, w}; }

typedef int cudaError_t;
typedef void* cudaStream_t;
typedef void* cudaEvent_t;
constexpr cudaError_t cudaSuccess = 0;
enum cudaMemcpyKind { cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice };
inline const char* cudaGetErrorString(cudaError_t) { return "emu"; }
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t) { memset(d, v, n); return cudaSuccess; }
inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) { memcpy(d, s, n); return cudaSuccess; }

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define __shared__ static
#define __align__(n) __attribute__((aligned(n)))
#define __grid_constant__

// ---- per-thread context -----------------------------------------------------------------------------
namespace emu {
struct Warp {
    std::barrier<> bar{32};
    uint32_t vals[32];
};
struct Block {
    explicit Block(int n) : bar(n), warps((n + 31) / 32) {
        for (auto& w : warps) w = std::make_unique<Warp>();
    }
    std::barrier<> bar;
    std::vector<std::unique_ptr<Warp>> warps;
};
inline thread_local Block* t_block = nullptr;
inline thread_local int t_linear = 0;
}  // namespace emu
inline thread_local uint3 threadIdx, blockIdx;
inline thread_local dim3 blockDim, gridDim;

inline void __syncthreads() { emu::t_block->bar.arrive_and_wait(); }

// full-warp collectives only (every use in the kernels passes 0xffffffff with all lanes converged)
inline void emu_exchange(uint32_t v, uint32_t out[32])
{
    emu::Warp& w = *emu::t_block->warps[emu::t_linear >> 5];
    w.vals[emu::t_linear & 31] = v;
    w.bar.arrive_and_wait();
    memcpy(out, w.vals, sizeof(w.vals));
    w.bar.arrive_and_wait();
}
inline void __syncwarp(unsigned = 0xffffffffu) { uint32_t t[32]; emu_exchange(0, t); }
inline unsigned __ballot_sync(unsigned, int pred)
{
    uint32_t t[32];
    emu_exchange(pred ? 1u : 0u, t);
    unsigned m = 0;
    for (int i = 0; i < 32; i++) m |= (t[i] & 1u) << i;
    return m;
}
inline uint32_t __shfl_up_sync(unsigned, uint32_t v, unsigned d)
{
    uint32_t t[32];
    emu_exchange(v, t);
    const unsigned lane = emu::t_linear & 31;
    return lane >= d ? t[lane - d] : v;
}
inline int __shfl_sync(unsigned, int v, int src)
{
    uint32_t t[32];
    emu_exchange((uint32_t)v, t);
    return (int)t[src & 31];
}
inline uint32_t __shfl_xor_sync(unsigned, uint32_t v, unsigned m)
{
    uint32_t t[32];
    emu_exchange(v, t);
    return t[(emu::t_linear & 31) ^ m];
}
inline float __uint_as_float(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
inline uint32_t __reduce_add_sync(unsigned, uint32_t v)
{
    uint32_t t[32], s = 0;
    emu_exchange(v, t);
    for (int i = 0; i < 32; i++) s += t[i];
    return s;
}
inline uint32_t __reduce_or_sync(unsigned mask, uint32_t v)   // subgroup masks: every lane passes the mask of its own group
{
    uint32_t t[32], s = 0;
    emu_exchange(v, t);
    for (int i = 0; i < 32; i++)
        if ((mask >> i) & 1u) s |= t[i];
    return s;
}
inline uint32_t __reduce_max_sync(unsigned, uint32_t v)
{
    uint32_t t[32], s = 0;
    emu_exchange(v, t);
    for (int i = 0; i < 32; i++) s = t[i] > s ? t[i] : s;
    return s;
}
inline void __trap() { abort(); }
inline uint32_t __reduce_min_sync(unsigned, uint32_t v)
{
    uint32_t t[32], s = 0xffffffffu;
    emu_exchange(v, t);
    for (int i = 0; i < 32; i++) s = std::min(s, t[i]);
    return s;
}

template <typename T> inline T __ldg(const T* p) { return *p; }
inline float __fmul_rn(float a, float b) { volatile float r = a * b; return r; }
inline float __fadd_rn(float a, float b) { volatile float r = a + b; return r; }
inline float __fdiv_rn(float a, float b) { volatile float r = a / b; return r; }
inline float __fsub_rn(float a, float b) { volatile float r = a - b; return r; }
inline double __dadd_rn(double a, double b) { volatile double r = a + b; return r; }
inline double __dmul_rn(double a, double b) { volatile double r = a * b; return r; }
inline int __float2int_rn(float a) { return (int)lrintf(a); }
inline unsigned __byte_perm(unsigned a, unsigned b, unsigned sel)
{
    const unsigned long long v = ((unsigned long long)b << 32) | a;
    unsigned r = 0;
    for (int i = 0; i < 4; i++) r |= (unsigned)((v >> (8 * ((sel >> (4 * i)) & 7))) & 255u) << (8 * i);
    return r;
}
inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline int __popcll(unsigned long long v) { return __builtin_popcountll(v); }
inline int __ffsll(long long v) { return __builtin_ffsll(v); }
inline unsigned __funnelshift_r(unsigned lo, unsigned hi, unsigned sh) { return (unsigned)((((unsigned long long)hi << 32) | lo) >> (sh & 31)); }
inline uint32_t atomicOr(uint32_t* p, uint32_t v) { return __atomic_fetch_or(p, v, __ATOMIC_RELAXED); }
inline uint32_t atomicAdd(uint32_t* p, uint32_t v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
inline uint32_t __dp4a(uint32_t a, uint32_t b, uint32_t c)
{
    for (int i = 0; i < 4; i++) c += ((a >> (8 * i)) & 255u) * ((b >> (8 * i)) & 255u);
    return c;
}
inline uint32_t __usad(uint32_t a, uint32_t b, uint32_t c) { return c + (a > b ? a - b : b - a); }
inline uint32_t __vabsdiffu4(uint32_t a, uint32_t b)
{
    uint32_t r = 0;
    for (int i = 0; i < 4; i++) {
        const int x = (int)((a >> (8 * i)) & 255u), y = (int)((b >> (8 * i)) & 255u);
        r |= (uint32_t)(x > y ? x - y : y - x) << (8 * i);
    }
    return r;
}
inline int __ffs(unsigned v) { return __builtin_ffs((int)v); }
using std::max;
using std::min;

namespace emu {
// CTAs run one after another; the threads of one CTA are real threads.
template <typename F> void launch(dim3 grid, dim3 block, F body)
{
    const int n = (int)(block.x * block.y * block.z);
    for (unsigned bz = 0; bz < grid.z; bz++)
        for (unsigned by = 0; by < grid.y; by++)
            for (unsigned bx = 0; bx < grid.x; bx++) {
                Block blk(n);
                std::vector<std::thread> th;
                th.reserve(n);
                for (int t = 0; t < n; t++)
                    th.emplace_back([&, t] {
                        t_block = &blk;
                        t_linear = t;
                        threadIdx = {(unsigned)t % block.x, ((unsigned)t / block.x) % block.y, (unsigned)t / (block.x * block.y)};
                        blockIdx = {bx, by, bz};
                        blockDim = block;
                        gridDim = grid;
                        body();
                    });
                for (auto& x : th) x.join();
            }
}
}  // namespace emu

#define OFPSB_LAUNCH_SMEM(kernel, grid, block, smem, stream, ...) ::emu::launch(dim3(grid), dim3(block), [=] { kernel(__VA_ARGS__); })
#define OFPSB_DYN_SMEM(name) static __attribute__((aligned(16))) uint8_t name[232448]
#define OFPSB_KEEP_LOADED(a, b) ((void)0)
#define OFPSB_LAUNCH(kernel, grid, block, stream, ...) ::emu::launch(dim3(grid), dim3(block), [=] { kernel(__VA_ARGS__); })
