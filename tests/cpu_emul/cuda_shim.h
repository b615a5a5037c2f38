// TEST INFRASTRUCTURE ONLY.  A minimal SIMT-on-CPU shim: compiles a plain CUDA C kernel (no inline PTX) as host code and runs it
// with ONE OS THREAD PER CUDA THREAD, so that bar.sync, warp shuffles, shared memory and the cooperative grid barrier keep their
// semantics (a missing barrier is a real data race here as well, and ThreadSanitizer finds it).  Kernels without a grid barrier can also run
// cooperatively (one OS thread, the block's threads as ucontext fibers that hand over at barriers): ~10x faster, nothing for TSan to see.
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
#include <ucontext.h>

#include <algorithm>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <thread>
#include <vector>

#define B200_CPU_EMUL 1
#ifdef __SANITIZE_ADDRESS__
#define B200_EMUL_SMEM_SLACK 0       // AddressSanitizer build: the first byte past the launch's dynamic shared memory is out of bounds
#else
#define B200_EMUL_SMEM_SLACK 16
#endif

struct alignas(8) float2 { float x, y; };      // as on the device: a misaligned vector access is an error there (-fsanitize=alignment finds it here)
struct alignas(16) float4 { float x, y, z, w; };
static inline float2 make_float2(float x, float y) { return float2{x, y}; }
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
struct emul_dim3 { unsigned x, y, z; };
typedef void* cudaStream_t;                               // only named by host declarations the kernel headers pull in
using std::max;
using std::min;

namespace cpu_emul {

// ---- cooperative mode: the threads of one block as user-level contexts on ONE OS thread --------------------------------------------
// A pthread barrier over 256 OS threads costs hundreds of context switches through the kernel; kernels with thousands of blocks and several
// bar.sync each (the IoUNet linear layers, the stem) are emulated ~20x faster when the block's threads are fibers that hand over at
// barriers.  launch_blocks() uses fibers unless B200_EMUL_THREADS=1 (ThreadSanitizer runs need real threads); launch() (concurrent blocks,
// grid barriers) always uses OS threads.
struct FiberBarrier { unsigned count = 0, gen = 0; };
struct FiberSched {
    std::vector<ucontext_t> ctx;
    std::vector<char> done;
    std::vector<std::vector<char>> stacks;
    ucontext_t main_ctx;
    int cur = -1, alive = 0;
    std::function<void(int)> body;
};
inline thread_local FiberSched* fsched = nullptr;          // non-null while this OS thread runs a block's fibers
void fiber_yield();

struct Warp {
    pthread_barrier_t bar;
    FiberBarrier fbar;
    unsigned lanes = 32;
    float buf[2][32];
};
struct Block {
    pthread_barrier_t bar;
    FiberBarrier fbar;
    unsigned nthreads = 0;
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
    int fiber;
};
inline thread_local ThreadCtx* tctx = nullptr;

inline unsigned char* dyn_smem() { return tctx->blk->smem.data(); }

struct FiberCtxTable { std::vector<ThreadCtx>* ctxs = nullptr; };
inline thread_local FiberCtxTable ftable;

inline void fiber_switch(int next) {
    FiberSched* S = fsched;
    const int prev = S->cur;
    S->cur = next;
    tctx = &(*ftable.ctxs)[next];
    swapcontext(&S->ctx[prev], &S->ctx[next]);
}

inline void fiber_yield() {                                  // hand over to the next unfinished fiber (round robin)
    FiberSched* S = fsched;
    const int n = (int)S->ctx.size();
    for (int k = 1; k <= n; ++k) {
        const int j = (S->cur + k) % n;
        if (!S->done[j] && j != S->cur) { fiber_switch(j); return; }
    }
    std::abort();                                            // every other thread finished while this one waits at a barrier: a kernel bug
}

inline void fiber_barrier(FiberBarrier& b, unsigned n) {
    if (++b.count == n) { b.count = 0; ++b.gen; return; }
    const unsigned my = b.gen;
    while (b.gen == my) fiber_yield();
}

inline void fiber_entry() {
    FiberSched* S = fsched;
    const int me = S->cur;
    S->body(me);
    S->done[me] = 1;
    S->alive -= 1;
    if (S->alive == 0) { tctx = nullptr; swapcontext(&S->ctx[me], &S->main_ctx); }
    fiber_yield();                                           // never returns here
    std::abort();
}

inline void run_fibers(std::vector<ThreadCtx>& ctxs, std::function<void(int)> body) {
    FiberSched S;
    const int n = (int)ctxs.size();
    S.ctx.resize(n); S.done.assign(n, 0); S.stacks.assign(n, std::vector<char>(256 * 1024)); S.alive = n; S.body = body;
    for (int i = 0; i < n; ++i) {
        getcontext(&S.ctx[i]);
        S.ctx[i].uc_stack.ss_sp = S.stacks[i].data();
        S.ctx[i].uc_stack.ss_size = S.stacks[i].size();
        S.ctx[i].uc_link = nullptr;
        makecontext(&S.ctx[i], fiber_entry, 0);
    }
    fsched = &S;
    ftable.ctxs = &ctxs;
    S.cur = 0;
    tctx = &ctxs[0];
    swapcontext(&S.main_ctx, &S.ctx[0]);
    fsched = nullptr;
    ftable.ctxs = nullptr;
    tctx = nullptr;
}

template <class Kernel, class... Args>
void launch(Kernel kernel, unsigned grid_dim, unsigned block_dim, size_t smem_bytes, Args... params) {
    Grid g;
    g.blocks = std::vector<Block>(grid_dim);
    pthread_barrier_init(&g.leaders, nullptr, grid_dim);
    const unsigned nwarps = (block_dim + 31) / 32;
    for (auto& b : g.blocks) {
        pthread_barrier_init(&b.bar, nullptr, block_dim);
        b.smem.assign(smem_bytes + B200_EMUL_SMEM_SLACK, 0xCD);                 // poisoned: uninitialised shared memory shows up as garbage
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

// Cooperative kernels with MANY blocks, cheaply: one OS thread per block, the block's threads as fibers inside it.  Blocks run
// concurrently (the grid barrier is a pthread barrier over the blocks' leader fibers: the other fibers of a block wait at bar.sync anyway),
// ThreadSanitizer sees the cross-block traffic through global memory, and -- with B200_EMUL_COOP_FIBERS defined before this header --
// function-local `__shared__` arrays become `static thread_local`, i.e. one copy per block.
template <class Kernel, class... Args>
void launch_coop(Kernel kernel, unsigned grid_dim, unsigned block_dim, size_t smem_bytes, Args... params) {
    Grid g;
    g.blocks = std::vector<Block>(grid_dim);
    pthread_barrier_init(&g.leaders, nullptr, grid_dim);
    const unsigned nwarps = (block_dim + 31) / 32;
    for (auto& b : g.blocks) {
        b.nthreads = block_dim;
        b.smem.assign(smem_bytes + B200_EMUL_SMEM_SLACK, 0xCD);
        b.warps = std::vector<Warp>(nwarps);
        for (unsigned w = 0; w < nwarps; ++w) b.warps[w].lanes = std::min(32u, block_dim - w * 32);
    }
    std::vector<std::thread> threads;
    threads.reserve(grid_dim);
    for (unsigned bi = 0; bi < grid_dim; ++bi)
        threads.emplace_back([&, bi]() {
            std::vector<ThreadCtx> ctxs(block_dim);
            for (unsigned t = 0; t < block_dim; ++t) {
                ThreadCtx& c = ctxs[t];
                c.tid = {t, 0, 0};
                c.bid = {bi, 0, 0};
                c.bdim = {block_dim, 1, 1};
                c.gdim = {grid_dim, 1, 1};
                c.blk = &g.blocks[bi];
                c.warp = &g.blocks[bi].warps[t / 32];
                c.grid = &g;
                c.fiber = 1;
            }
            run_fibers(ctxs, [&](int) { kernel(params...); });
        });
    for (auto& th : threads) th.join();
    pthread_barrier_destroy(&g.leaders);
}

// Kernels without a grid barrier: the blocks of a (3-D) grid one after the other (fibers by default, OS threads with B200_EMUL_THREADS=1).  Function-local `__shared__` arrays (static here) are safe
// in this mode because only one block is alive at a time.
template <class Kernel, class... Args>
void launch_blocks2(Kernel kernel, unsigned gx, unsigned gy, unsigned gz, unsigned block_x, unsigned block_y, size_t smem_bytes, Args... params) {
    const unsigned block_dim = block_x * block_y;            // threads are linearised x-fastest, as on the device (warps = 32 consecutive)
    static const bool use_threads = std::getenv("B200_EMUL_THREADS") != nullptr;
    if (!use_threads) {                                      // cooperative mode: one OS thread, the block's threads as fibers
        Grid g;
        g.blocks = std::vector<Block>(1);
        Block& b = g.blocks[0];
        b.nthreads = block_dim;
        b.smem.assign(smem_bytes + B200_EMUL_SMEM_SLACK, 0xCD);
        const unsigned nwarps = (block_dim + 31) / 32;
        b.warps = std::vector<Warp>(nwarps);
        for (unsigned w = 0; w < nwarps; ++w) b.warps[w].lanes = std::min(32u, block_dim - w * 32);
        std::vector<ThreadCtx> ctxs(block_dim);
        for (unsigned t = 0; t < block_dim; ++t) {
            ThreadCtx& c = ctxs[t];
            c.tid = {t % block_x, t / block_x, 0};
            c.bdim = {block_x, block_y, 1};
            c.gdim = {gx, gy, gz};
            c.blk = &b;
            c.warp = &b.warps[t / 32];
            c.grid = &g;
            c.fiber = 1;
        }
        run_fibers(ctxs, [&](int me) {
            ThreadCtx& c = ctxs[me];
            for (unsigned bz = 0; bz < gz; ++bz)
                for (unsigned by = 0; by < gy; ++by)
                    for (unsigned bx = 0; bx < gx; ++bx) {
                        c.bid = {bx, by, bz};
                        kernel(params...);
                        fiber_barrier(b.fbar, block_dim);    // the block is finished
                    }
        });
        return;
    }
    // one pool of block_dim OS threads walks the grid block by block (a barrier closes every block), so thread creation is paid once per
    // launch; the dynamic shared memory is poisoned once and then carries over from block to block, as it may on the device
    Grid g;
    g.blocks = std::vector<Block>(1);
    pthread_barrier_init(&g.leaders, nullptr, 1);
    Block& b = g.blocks[0];
    pthread_barrier_init(&b.bar, nullptr, block_dim);
    b.smem.assign(smem_bytes + B200_EMUL_SMEM_SLACK, 0xCD);
    const unsigned nwarps = (block_dim + 31) / 32;
    b.warps = std::vector<Warp>(nwarps);
    for (unsigned w = 0; w < nwarps; ++w) pthread_barrier_init(&b.warps[w].bar, nullptr, std::min(32u, block_dim - w * 32));
    std::vector<std::thread> threads;
    threads.reserve(block_dim);
    for (unsigned t = 0; t < block_dim; ++t)
        threads.emplace_back([&, t]() {
            ThreadCtx c{};
            c.tid = {t % block_x, t / block_x, 0};
            c.bdim = {block_x, block_y, 1};
            c.gdim = {gx, gy, gz};
            c.blk = &b;
            c.warp = &b.warps[t / 32];
            c.grid = &g;
            tctx = &c;
            for (unsigned bz = 0; bz < gz; ++bz)
                for (unsigned by = 0; by < gy; ++by)
                    for (unsigned bx = 0; bx < gx; ++bx) {
                        c.bid = {bx, by, bz};
                        kernel(params...);
                        pthread_barrier_wait(&b.bar);        // the block is finished (threads that returned early arrive here too)
                    }
            tctx = nullptr;
        });
    for (auto& th : threads) th.join();
    pthread_barrier_destroy(&b.bar);
    for (auto& w : b.warps) pthread_barrier_destroy(&w.bar);
    pthread_barrier_destroy(&g.leaders);
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

static inline void __syncthreads() {
    auto* c = ::cpu_emul::tctx;
    if (c->fiber) ::cpu_emul::fiber_barrier(c->blk->fbar, c->blk->nthreads); else pthread_barrier_wait(&c->blk->bar);
}

static inline float __shfl_xor_sync(unsigned, float v, int lane_mask) {
    auto* c = ::cpu_emul::tctx;
    const int lane = c->tid.x & 31;
    float* buf = c->warp->buf[c->parity];
    c->parity ^= 1;
    buf[lane] = v;
    if (c->fiber) ::cpu_emul::fiber_barrier(c->warp->fbar, c->warp->lanes); else pthread_barrier_wait(&c->warp->bar);
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

// 32-bit "shared window" address of a pointer into the block's dynamic shared memory (cp.async destinations, csrc/corr2.cuh)
static inline size_t __cvta_generic_to_shared(const void* p) {
    return (size_t)(reinterpret_cast<const unsigned char*>(p) - ::cpu_emul::dyn_smem());
}
static inline unsigned atomicAdd(unsigned* addr, unsigned v) { return __atomic_fetch_add(addr, v, __ATOMIC_ACQ_REL); }
static inline void __threadfence() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
static inline float rsqrtf(float a) { return 1.0f / std::sqrt(a); }
static inline void sincospif(float a, float* s, float* c) { *s = (float)std::sin(3.14159265358979323846 * (double)a); *c = (float)std::cos(3.14159265358979323846 * (double)a); }
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
// static __shared__ arrays: a function-local static, i.e. only for kernels launched with ONE live block at a time (launch with one block,
// launch_blocks) -- or, with B200_EMUL_COOP_FIBERS, one copy per OS thread = per block of launch_coop
#ifdef B200_EMUL_COOP_FIBERS
#define __shared__ static thread_local
#else
#define __shared__ static
#endif

static inline void __syncwarp(unsigned = 0xffffffffu) {
    auto* c = ::cpu_emul::tctx;
    if (c->fiber) ::cpu_emul::fiber_barrier(c->warp->fbar, c->warp->lanes); else pthread_barrier_wait(&c->warp->bar);
}

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

static inline float ordered_sum_ldcg(const float* p, size_t stride, int count) {      // common.cuh: index order, batches of 8
    float s = 0.f;
    for (int k0 = 0; k0 < count; k0 += 8) {
        float v[8];
        for (int u = 0; u < 8; ++u) v[u] = (k0 + u < count) ? p[(size_t)(k0 + u) * stride] : 0.f;
        for (int u = 0; u < 8; ++u) s += v[u];
    }
    return s;
}

static inline void eco_grid_barrier(unsigned* c, unsigned& epoch, unsigned&) { grid_barrier(c, epoch); }

}  // namespace b200trk
