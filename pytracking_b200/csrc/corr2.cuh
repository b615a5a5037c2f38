// Second-generation correlation sweeps (used by the persistent SD optimisers): apply (s = A w) and its adjoint
// (g = A^T r) for one 4x4 filter over [n,C,FS,FS] sample memories.
//
// Both sweeps are ~9 FLOP per byte of sample memory: the kernel must be fp32-FMA efficient AND keep many bytes in
// flight.  Design:
//   * an "item" = 16 consecutive channel planes of one sample (16 x FS^2 floats, contiguous in HBM/L2).  Items are
//     streamed into shared memory with 8-byte cp.async (LDGSTS) straight into a zero-bordered [ROWS][28] layout
//     (no register staging), NST-deep multistage pipeline, one block barrier per item.
//   * thread = (channel slot = tid & 15, 5x4 output tile = tid >> 4).  Lanes of a quarter warp are 8 different
//     planes; the plane stride (ROWS*28 floats, ROWS odd) is an odd number of 16-byte units, so every 128-bit shared
//     load is bank-conflict free.  A thread reads its 8x8 input patch with 16 LDS.128 and issues 320 FMAs
//     (20 FMA per shared-memory instruction -> FMA-pipe bound, not LSU bound).
//   * apply: the 16 slot partials of an output are reduced with 4 xor-shuffles inside the half warp;
//     transpose: the per-tile partial gradients are reduced through shared memory once per channel pass.
#pragma once
#include "common.cuh"

// 8-byte asynchronous global -> shared copy.  The host build of the kernels (tests/cpu_emul/cuda_shim.h) performs the copy at once -- the
// earliest moment the hardware could land it, i.e. the schedule that exposes a stage reused too early.
#ifdef B200_CPU_EMUL
#define B200_CP_ASYNC_8(d, g) std::memcpy(::cpu_emul::dyn_smem() + (d), (g), 8)
#else
#define B200_CP_ASYNC_8(d, g) asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(g) : "memory")
#endif

namespace b200trk {

template <int FS>
struct Corr2 {
    static constexpr int KS = 4, PAD = 2, SLOTS = 16;
    static constexpr int OS = FS + 1;                            // 19 / 23
    static constexpr int TR = 5, TC = 4;
    static constexpr int NTY = (OS + TR - 1) / TR, NTX = (OS + TC - 1) / TC;
    static constexpr int NT = NTY * NTX;                         // 20 / 30 tiles
    static constexpr int PITCH = 28;                             // floats per padded row (>= NTX*4 + 4)
    static constexpr int ROWS0 = NTY * TR + KS - 1;              // 23 / 28 rows touched
    static constexpr int ROWS = (ROWS0 & 1) ? ROWS0 : ROWS0 + 1; // odd -> plane stride is an odd multiple of 16 B
    static constexpr int PLANE = ROWS * PITCH;                   // floats per padded plane (644 / 812)
    static constexpr int ITEM_FLOATS = SLOTS * PLANE;            // one pipeline stage
    static constexpr int NPOS = OS * OS;
    static constexpr int FPLANE = FS * FS;
    static constexpr int PW = NTX * TC;                          // width of a tile-padded map row (20 / 24)
    static constexpr int PMAP = NTY * TR * PW;                   // floats of a tile-padded map (400 / 600)
    static constexpr int NCONS = NT * SLOTS;                     // consumer (= all) threads: 320 / 480
    static constexpr int CHUNK8 = FS / 2;                        // 8-byte chunks per feature row (9 / 11)
    static constexpr int NCP = SLOTS * FS * CHUNK8;              // cp.async ops per item (2592 / 3872)
    static constexpr int CPT = (NCP + NCONS - 1) / NCONS;        // per thread (9 / 9)
    static constexpr int VEC_STRIDE = 20;                        // floats between the tap vectors of two channels in smem
    static constexpr int RED_STRIDE = 17;
    static_assert(PITCH >= NTX * TC + 4, "pitch too small for the 8-wide patch reads");
    static_assert((PLANE / 4) % 2 == 1, "plane stride must be an odd number of 16-byte units");
    static_assert(FS % 2 == 0, "8-byte copies need an even feature width");

    // ---- item geometry -------------------------------------------------------------------------------------
    struct Ctx {
        const float* feat;   // [n,C,FS,FS]
        int C, n, c0, passes, group, NG;
        int dbg_mode;        // timing experiments only: 1 = skip the FMA tiles, 2 = skip the copies
        int pstride;         // floats between two channel planes of the sample memory; 0 = dense (FS*FS)
        __device__ __forceinline__ int spc() const { return (n - group + NG - 1) / NG; }
        __device__ __forceinline__ int sample(int j) const { return group + j * NG; }
        __device__ __forceinline__ int ps() const { return pstride ? pstride : FPLANE; }
        __device__ __forceinline__ const float* src(int j, int p) const {
            return feat + ((size_t)sample(j) * C + c0 + p * SLOTS) * ps();
        }
    };

    // zero every stage once (the borders are never written again)
    __device__ static __forceinline__ void zero_stages(float* stages, int nst) {
        float4* p = reinterpret_cast<float4*>(stages);
        for (int i = threadIdx.x; i < nst * ITEM_FLOATS / 4; i += NCONS) p[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }

    // issue the asynchronous copy of one item (16 planes, dense) into the interior of a padded stage
    __device__ static __forceinline__ void issue_item(const float* __restrict__ src, float* __restrict__ stage, int ps) {
        const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(stage);
#pragma unroll
        for (int m = 0; m < CPT; ++m) {
            const int idx = threadIdx.x + m * NCONS;
            if (idx < NCP) {
                const int slot = idx / (FS * CHUNK8);
                const int rem = idx - slot * (FS * CHUNK8);
                const int row = rem / CHUNK8, ch = rem - row * CHUNK8;
                const float* g = src + slot * ps + row * FS + ch * 2;
                const uint32_t d = sbase + (uint32_t)(slot * PLANE + (row + PAD) * PITCH + PAD + ch * 2) * 4u;
                B200_CP_ASYNC_8(d, g);
            }
        }
    }
    // one of the CPT copy chunks of an item (interleaved with the FMA rows of the previous item's tile so that the
    // LSU copy traffic overlaps the FMA pipe instead of forming a separate phase)
    __device__ static __forceinline__ void issue_chunk(const float* __restrict__ src, uint32_t sbase, int m, int ps) {
        const int idx = threadIdx.x + m * NCONS;
        if (idx < NCP) {
            const int slot = idx / (FS * CHUNK8);
            const int rem = idx - slot * (FS * CHUNK8);
            const int row = rem / CHUNK8, ch = rem - row * CHUNK8;
            const float* g = src + slot * ps + row * FS + ch * 2;
            const uint32_t d = sbase + (uint32_t)(slot * PLANE + (row + PAD) * PITCH + PAD + ch * 2) * 4u;
            B200_CP_ASYNC_8(d, g);
        }
    }
#ifdef B200_CPU_EMUL
    __device__ static __forceinline__ void commit() {}
    template <int N>
    __device__ static __forceinline__ void wait_group() {}
#else
    __device__ static __forceinline__ void commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
    template <int N>
    __device__ static __forceinline__ void wait_group() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
#endif

    // ---- register-tile kernels -----------------------------------------------------------------------------
    // acc[orow*4+oc] += sum_{u,v} x[(5ty+orow+u), (4tx+oc+v)] * w[u*4+v]
    // Loop order (v outer; u, oc inner): 16 consecutive FMAs hit 16 different accumulators; the next patch row is
    // fetched from shared memory before the FMAs of the current one are issued.
    __device__ static __forceinline__ void apply_tile(const float* __restrict__ patch, const float (&w)[16], float (&acc)[20],
                                                      const float* __restrict__ nsrc, uint32_t nstage, int ps) {
        float4 a = *reinterpret_cast<const float4*>(patch);
        float4 b = *reinterpret_cast<const float4*>(patch + 4);
#pragma unroll
        for (int r = 0; r < TR + 3; ++r) {
            const float x[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
            if (r + 1 < TR + 3) {
                a = *reinterpret_cast<const float4*>(patch + (r + 1) * PITCH);
                b = *reinterpret_cast<const float4*>(patch + (r + 1) * PITCH + 4);
            }
            if (nsrc) {
                issue_chunk(nsrc, nstage, r, ps);
                if (r == TR + 2) {
#pragma unroll
                    for (int m = TR + 3; m < CPT; ++m) issue_chunk(nsrc, nstage, m, ps);
                }
            }
#pragma unroll
            for (int v = 0; v < 4; ++v)
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int orow = r - u;
                    if (orow >= 0 && orow < TR) {
#pragma unroll
                        for (int oc = 0; oc < 4; ++oc)
                            acc[orow * 4 + oc] = fmaf(x[oc + v], w[u * 4 + v], acc[orow * 4 + oc]);
                    }
                }
        }
    }
    // g[u*4+v] += sum_{orow,oc} R[orow*4+oc] * x[(5ty+orow+u), (4tx+oc+v)]
    // Loop order (oc outer; u, v inner): 16 consecutive FMAs hit the 16 different gradient taps.
    __device__ static __forceinline__ void transpose_tile(const float* __restrict__ patch, const float (&R)[20], float (&g)[16],
                                                          const float* __restrict__ nsrc, uint32_t nstage, int ps) {
        float4 a = *reinterpret_cast<const float4*>(patch);
        float4 b = *reinterpret_cast<const float4*>(patch + 4);
#pragma unroll
        for (int r = 0; r < TR + 3; ++r) {
            const float x[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
            if (r + 1 < TR + 3) {
                a = *reinterpret_cast<const float4*>(patch + (r + 1) * PITCH);
                b = *reinterpret_cast<const float4*>(patch + (r + 1) * PITCH + 4);
            }
            if (nsrc) {
                issue_chunk(nsrc, nstage, r, ps);
                if (r == TR + 2) {
#pragma unroll
                    for (int m = TR + 3; m < CPT; ++m) issue_chunk(nsrc, nstage, m, ps);
                }
            }
#pragma unroll
            for (int oc = 0; oc < 4; ++oc)
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int orow = r - u;
                    if (orow >= 0 && orow < TR) {
#pragma unroll
                        for (int v = 0; v < 4; ++v)
                            g[u * 4 + v] = fmaf(R[orow * 4 + oc], x[oc + v], g[u * 4 + v]);
                    }
                }
        }
    }

    // ---- pipelined sweeps ----------------------------------------------------------------------------------
    // Item order: apply  -> (j outer, p inner): a sample's score partial finishes after `passes` items;
    //             transpose -> (p outer, j inner): a channel pass' gradient finishes after `spc` items.
    template <bool APPLY>
    __device__ static __forceinline__ void item_jp(const Ctx& cx, int t, int& j, int& p) {
        if (APPLY) { j = t / cx.passes; p = t - j * cx.passes; }
        else { const int spc = cx.spc(); p = t / spc; j = t - p * spc; }
    }

    // prefetch the first NST-1 items of a sweep (may be called before a grid barrier: the sample memory is read-only)
    template <bool APPLY, int NST>
    __device__ static __forceinline__ void sweep_prologue(const Ctx& cx, float* stages) {
        const int nitems = cx.spc() * cx.passes;
#pragma unroll
        for (int s = 0; s < NST - 1; ++s) {
            if (s < nitems) {
                int j, p;
                item_jp<APPLY>(cx, s, j, p);
                issue_item(cx.src(j, p), stages + s * ITEM_FLOATS, cx.ps());
            }
            commit();
        }
    }

    // apply sweep: partial scores of the CTA's channels for each of its samples.
    //   vec  : smem, [passes*SLOTS][VEC_STRIDE] filter taps of the chunk
    //   part : global; sample i goes to part + i*part_stride_sample (NPOS floats, linear layout)
    template <int NST>
    __device__ static void sweep_apply(const Ctx& cx, float* stages, const float* vec, float* part, size_t part_stride_sample) {
        const int tid = threadIdx.x, slot = tid & 15, tile = tid >> 4;
        const int ty = tile / NTX, tx = tile - ty * NTX;
        const int nitems = cx.spc() * cx.passes;
        const int poff = slot * PLANE + (ty * TR) * PITCH + tx * TC;
        float acc[20];
        for (int t = 0; t < nitems; ++t) {
            wait_group<NST - 2>();
            __syncthreads();
            const float* nsrc = nullptr;
            uint32_t nstage = 0;
            {
                const int tn = t + NST - 1;
                if (tn < nitems) {
                    int jn, pn;
                    item_jp<true>(cx, tn, jn, pn);
                    nsrc = cx.src(jn, pn);
                    nstage = (uint32_t)__cvta_generic_to_shared(stages + (tn % NST) * ITEM_FLOATS);
                }
                if (cx.dbg_mode == 2) nsrc = nullptr;
            }
            const int j = t / cx.passes, p = t - j * cx.passes;
            if (p == 0) {
#pragma unroll
                for (int q = 0; q < 20; ++q) acc[q] = 0.f;
            }
            float w[16];
            {
                const float4* wp = reinterpret_cast<const float4*>(vec + (p * SLOTS + slot) * VEC_STRIDE);
                const float4 w0 = wp[0], w1 = wp[1], w2 = wp[2], w3 = wp[3];
                w[0] = w0.x; w[1] = w0.y; w[2] = w0.z; w[3] = w0.w; w[4] = w1.x; w[5] = w1.y; w[6] = w1.z; w[7] = w1.w;
                w[8] = w2.x; w[9] = w2.y; w[10] = w2.z; w[11] = w2.w; w[12] = w3.x; w[13] = w3.y; w[14] = w3.z; w[15] = w3.w;
            }
            if (cx.dbg_mode != 1) apply_tile(stages + (t % NST) * ITEM_FLOATS + poff, w, acc, nsrc, nstage, cx.ps());
            else if (nsrc) issue_item(nsrc, stages + ((t + NST - 1) % NST) * ITEM_FLOATS, cx.ps());
            commit();
            if (p == cx.passes - 1) {
                // sum over the 16 channel slots of the half warp, fixed butterfly order (deterministic)
#pragma unroll
                for (int q = 0; q < 20; ++q) {
                    float v = acc[q];
                    v += __shfl_xor_sync(0xffffffffu, v, 8);
                    v += __shfl_xor_sync(0xffffffffu, v, 4);
                    v += __shfl_xor_sync(0xffffffffu, v, 2);
                    v += __shfl_xor_sync(0xffffffffu, v, 1);
                    acc[q] = v;
                }
                if (slot == 0) {
                    float* dst = part + (size_t)cx.sample(j) * part_stride_sample;
#pragma unroll
                    for (int q = 0; q < 20; ++q) {
                        const int y = ty * TR + q / 4, x = tx * TC + (q & 3);
                        if (y < OS && x < OS) dst[y * OS + x] = acc[q];
                    }
                }
            }
        }
        wait_group<0>();
        __syncthreads();       // every stage is free again (the next sweep's prologue may overwrite them)
    }

    // transpose sweep: partial filter gradient of the CTA's channels summed over the CTA's samples.
    //   rt   : smem, [spc][PMAP] tile-padded residual maps (zero outside the OS x OS map)
    //   red  : smem scratch, NT*SLOTS*RED_STRIDE floats
    //   gout : global, [passes*SLOTS][16] destination of this CTA's partial
    template <int NST>
    __device__ static void sweep_transpose(const Ctx& cx, float* stages, float* red, const float* rt, float* gout) {
        const int tid = threadIdx.x, slot = tid & 15, tile = tid >> 4;
        const int ty = tile / NTX, tx = tile - ty * NTX;
        const int spc = cx.spc();
        const int nitems = spc * cx.passes;
        const int poff = slot * PLANE + (ty * TR) * PITCH + tx * TC;
        float g[16];
        for (int t = 0; t < nitems; ++t) {
            wait_group<NST - 2>();
            __syncthreads();
            const float* nsrc = nullptr;
            uint32_t nstage = 0;
            {
                const int tn = t + NST - 1;
                if (tn < nitems) {
                    int jn, pn;
                    item_jp<false>(cx, tn, jn, pn);
                    nsrc = cx.src(jn, pn);
                    nstage = (uint32_t)__cvta_generic_to_shared(stages + (tn % NST) * ITEM_FLOATS);
                }
                if (cx.dbg_mode == 2) nsrc = nullptr;
            }
            const int p = t / spc, j = t - p * spc;
            if (j == 0) {
#pragma unroll
                for (int q = 0; q < 16; ++q) g[q] = 0.f;
            }
            float R[20];
            {
                const float* rp = rt + j * PMAP + (ty * TR) * PW + tx * TC;
#pragma unroll
                for (int o = 0; o < TR; ++o) {
                    const float4 v = *reinterpret_cast<const float4*>(rp + o * PW);
                    R[o * 4] = v.x; R[o * 4 + 1] = v.y; R[o * 4 + 2] = v.z; R[o * 4 + 3] = v.w;
                }
            }
            if (cx.dbg_mode != 1) transpose_tile(stages + (t % NST) * ITEM_FLOATS + poff, R, g, nsrc, nstage, cx.ps());
            else if (nsrc) issue_item(nsrc, stages + ((t + NST - 1) % NST) * ITEM_FLOATS, cx.ps());
            commit();
            if (j == spc - 1) {
                float* rp = red + (tile * SLOTS + slot) * RED_STRIDE;
#pragma unroll
                for (int q = 0; q < 16; ++q) rp[q] = g[q];
                __syncthreads();
                if (tid < SLOTS * 16) {
                    const int sl = tid >> 4, q = tid & 15;
                    float s = 0.f;
#pragma unroll 4
                    for (int tl = 0; tl < NT; ++tl) s += red[(tl * SLOTS + sl) * RED_STRIDE + q];
                    gout[(p * SLOTS + sl) * 16 + q] = s;
                }
                // `red` is rewritten only after the next item's block barrier
            }
        }
        wait_group<0>();
        __syncthreads();
    }
};

}  // namespace b200trk
