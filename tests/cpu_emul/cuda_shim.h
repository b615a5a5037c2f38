// TEST INFRASTRUCTURE ONLY.  A minimal SIMT-on-CPU shim: compiles a plain CUDA C kernel (no inline PTX) as host code and runs it
// with ONE OS THREAD PER CUDA THREAD, so that bar.sync, warp shuffles, shared memory and the cooperative grid barrier keep their
// semantics (a missing barrier is a real data race here as well, and ThreadSanitizer finds it).
//   __syncthreads()        -> pthread barrier over the block's threads
//   __shfl_xor_sync()      -> exchange through a per-warp buffer, one pthread barrier per shuffle (double buffered)
//   __syncwarp()           -> pthread barrier over the warp's threads
//   extern __shared__      -> B200_DYN_SMEM(name): one allocation per block; static __shared__ -> a function-local static (one block at a time)
//   b200trk::grid_barrier  -> bar.sync + pthread barrier over the blocks' leader threads + bar.sync
// The device helpers of csrc/common.cuh used by the emulated kernels (warp_sum, block_sum, grid_barrier) are restated below with
// the same summation order.  Grids are kept small (a few blocks of 64-128 threads): the kernels under test take their
// decomposition from gridDim / blockDim.
#pragma once
#include <pthread.h>

#include <algorithm>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

#define B200_CPU_EMUL 1

struct float2 { float x, y; };
struct alignas(16) float4 { float x, y, z, w; };
static inline float2 make_float2(float x, float y) { return float2{x, y}; }
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
struct emul_dim3 { unsigned x, y, z; };
typedef void* cudaStream_t;                               // only named by host declarations the kernel headers pull in
using std::max;
using std::min;

namespace cpu_emul {

struct Warp {
    pthread_barrier_t bar;
    float buf[2][32];
};
struct Block {
    pthread_barrier_t bar;
    std::vector<unsigned char> smem;
    std::vector<Warp> warps;
};
struct Grid {
    pthread_barrier_t leaders;
    std::vector<Block> blocks;
};
struct ThreadCtx {
    emul_dim3 tid, bid, bdim, gdim;
    Block* blk;
    Warp* warp;
    Grid* grid;
    int parity;
};
inline thread_local ThreadCtx* tctx = nullptr;

inline unsigned char* dyn_smem() { return tctx->blk->smem.data(); }

template <class Kernel, class... Args>
void launch(Kernel kernel, unsigned grid_dim, unsigned block_dim, size_t smem_bytes, Args... params) {
    Grid g;
    g.blocks = std::vector<Block>(grid_dim);
    pthread_barrier_init(&g.leaders, nullptr, grid_dim);
    const unsigned nwarps = (block_dim + 31) / 32;
    for (auto& b : g.blocks) {
        pthread_barrier_init(&b.bar, nullptr, block_dim);
        b.smem.assign(smem_bytes + 16, 0xCD);                 // poisoned: uninitialised shared memory shows up as garbage
        b.warps = std::vector<Warp>(nwarps);
        for (unsigned w = 0; w < nwarps; ++w) {
            const unsigned lanes = std::min(32u, block_dim - w * 32);
            pthread_barrier_init(&b.warps[w].bar, nullptr, lanes);
        }
    }
    std::vector<std::thread> threads;
    threads.reserve((size_t)grid_dim * block_dim);
    for (unsigned b = 0; b < grid_dim; ++b)
        for (unsigned t = 0; t < block_dim; ++t)
            threads.emplace_back([&, b, t]() {
                ThreadCtx c{};
                c.tid = {t, 0, 0};
                c.bid = {b, 0, 0};
                c.bdim = {block_dim, 1, 1};
                c.gdim = {grid_dim, 1, 1};
                c.blk = &g.blocks[b];
                c.warp = &g.blocks[b].warps[t / 32];
                c.grid = &g;
                c.parity = 0;
                tctx = &c;
                kernel(params...);
                tctx = nullptr;
            });
    for (auto& th : threads) th.join();
    for (auto& b : g.blocks) {
        pthread_barrier_destroy(&b.bar);
        for (auto& w : b.warps) pthread_barrier_destroy(&w.bar);
    }
    pthread_barrier_destroy(&g.leaders);
}

// Kernels without a grid barrier: the blocks of a (3-D) grid one after the other, each with its own OS threads.  Function-local
// `__shared__` arrays (static here) are safe in this mode because only one block is alive at a time.
template <class Kernel, class... Args>
void launch_blocks2(Kernel kernel, unsigned gx, unsigned gy, unsigned gz, unsigned block_x, unsigned block_y, size_t smem_bytes, Args... params) {
    const unsigned block_dim = block_x * block_y;            // threads are linearised x-fastest, as on the device (warps = 32 consecutive)
    for (unsigned bz = 0; bz < gz; ++bz)
        for (unsigned by = 0; by < gy; ++by)
            for (unsigned bx = 0; bx < gx; ++bx) {
                Grid g;
                g.blocks = std::vector<Block>(1);
                pthread_barrier_init(&g.leaders, nullptr, 1);
                Block& b = g.blocks[0];
                pthread_barrier_init(&b.bar, nullptr, block_dim);
                b.smem.assign(smem_bytes + 16, 0xCD);
                const unsigned nwarps = (block_dim + 31) / 32;
                b.warps = std::vector<Warp>(nwarps);
                for (unsigned w = 0; w < nwarps; ++w) pthread_barrier_init(&b.warps[w].bar, nullptr, std::min(32u, block_dim - w * 32));
                std::vector<std::thread> threads;
                threads.reserve(block_dim);
                for (unsigned t = 0; t < block_dim; ++t)
                    threads.emplace_back([&, t]() {
                        ThreadCtx c{};
                        c.tid = {t % block_x, t / block_x, 0};
                        c.bid = {bx, by, bz};
                        c.bdim = {block_x, block_y, 1};
                        c.gdim = {gx, gy, gz};
                        c.blk = &b;
                        c.warp = &b.warps[t / 32];
                        c.grid = &g;
                        tctx = &c;
                        kernel(params...);
                        tctx = nullptr;
                    });
                for (auto& th : threads) th.join();
                pthread_barrier_destroy(&b.bar);
                for (auto& w : b.warps) pthread_barrier_destroy(&w.bar);
                pthread_barrier_destroy(&g.leaders);
            }
}

template <class Kernel, class... Args>
void launch_blocks(Kernel kernel, unsigned gx, unsigned gy, unsigned gz, unsigned block_dim, size_t smem_bytes, Args... params) {
    launch_blocks2(kernel, gx, gy, gz, block_dim, 1u, smem_bytes, params...);
}

// Kernels without any barrier or shuffle (one thread = one output element): every thread of a 2-D grid in turn, on the calling thread.
template <class Kernel, class... Args>
void launch_serial(Kernel kernel, unsigned grid_x, unsigned grid_y, unsigned block_dim, Args... params) {
    ThreadCtx c{};
    c.bdim = {block_dim, 1, 1};
    c.gdim = {grid_x, grid_y, 1};
    tctx = &c;
    for (unsigned by = 0; by < grid_y; ++by)
        for (unsigned bx = 0; bx < grid_x; ++bx)
            for (unsigned t = 0; t < block_dim; ++t) {
                c.tid = {t, 0, 0};
                c.bid = {bx, by, 0};
                kernel(params...);
            }
    tctx = nullptr;
}

}  // namespace cpu_emul

#define threadIdx (::cpu_emul::tctx->tid)
#define blockIdx (::cpu_emul::tctx->bid)
#define blockDim (::cpu_emul::tctx->bdim)
#define gridDim (::cpu_emul::tctx->gdim)
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __align__(n) __attribute__((aligned(n)))

static inline void __syncthreads() { pthread_barrier_wait(&::cpu_emul::tctx->blk->bar); }

static inline float __shfl_xor_sync(unsigned, float v, int lane_mask) {
    auto* c = ::cpu_emul::tctx;
    const int lane = c->tid.x & 31;
    float* buf = c->warp->buf[c->parity];
    c->parity ^= 1;
    buf[lane] = v;
    pthread_barrier_wait(&c->warp->bar);
    return buf[lane ^ lane_mask];
}

static inline int __shfl_xor_sync(unsigned m, int v, int lane_mask) {
    float f;
    std::memcpy(&f, &v, 4);
    f = __shfl_xor_sync(m, f, lane_mask);
    std::memcpy(&v, &f, 4);
    return v;
}

template <class T>
static inline T __ldcg(const T* p) { return *p; }

static inline unsigned atomicAdd(unsigned* addr, unsigned v) { return __atomic_fetch_add(addr, v, __ATOMIC_ACQ_REL); }
static inline void __threadfence() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
static inline float rsqrtf(float a) { return 1.0f / std::sqrt(a); }
#define __expf(x) expf(x)      // the fast-math intrinsic: tolerance tests only
template <class T>
static inline T __ldg(const T* p) { return *p; }

static inline float atomicAdd(float* addr, float v) {      // CAS loop on the bit pattern (the emulating threads are real threads)
    unsigned* ia = reinterpret_cast<unsigned*>(addr);
    unsigned old_bits = __atomic_load_n(ia, __ATOMIC_RELAXED), new_bits;
    float old_val;
    do {
        std::memcpy(&old_val, &old_bits, 4);
        const float nv = old_val + v;
        std::memcpy(&new_bits, &nv, 4);
    } while (!__atomic_compare_exchange_n(ia, &old_bits, new_bits, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED));
    return old_val;
}

// explicitly rounded single operations (compile the emulation with -ffp-contract=off so that they stay single operations)
static inline float __fmaf_rn(float a, float b, float c) { return std::fmaf(a, b, c); }
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fsub_rn(float a, float b) { return a - b; }
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fdiv_rn(float a, float b) { return a / b; }
static inline float __fsqrt_rn(float a) { return std::sqrt(a); }
static inline float __int_as_float(int v) { float f; std::memcpy(&f, &v, 4); return f; }
// static __shared__ arrays: only for kernels launched with ONE block at a time (the array is shared by that block's threads)
#define __shared__ static

static inline void __syncwarp(unsigned = 0xffffffffu) { pthread_barrier_wait(&::cpu_emul::tctx->warp->bar); }

namespace b200trk {

static inline float warp_sum(float v) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

static inline float block_sum(float v, float* red) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) red[wid] = v;
    __syncthreads();
    const int nw = (blockDim.x + 31) >> 5;
    float t = (lane < nw) ? red[lane] : 0.f;
    t = warp_sum(t);
    return t;
}

static inline void grid_barrier(unsigned*, unsigned& epoch) {
    __syncthreads();
    epoch += 1;
    if (threadIdx.x == 0) pthread_barrier_wait(&::cpu_emul::tctx->grid->leaders);
    __syncthreads();
}

static inline void eco_grid_barrier(unsigned* c, unsigned& epoch, unsigned&) { grid_barrier(c, epoch); }

}  // namespace b200trk
