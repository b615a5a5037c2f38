// Micro-benchmark: how fast can ONE SM pull L2-resident fp32 operand tiles into shared memory, and through which path?
// (both tcgen05 kernels are paced by their operand fill.)  Every CTA streams `units` 16 KB tiles of its own
// slice of a 33 MB buffer (the DiMP sample memory: [50][512][324] fp32), `sweeps` times, so that everything after the first sweep
// comes from L2.  Variants:
//   0  TMA 3-D box {32 px, 128 ch}, SWIZZLE_128B   (the adjoint-sweep box of sd_tc.cu)
//   1  TMA 3-D box {128 px, 32 ch}, no swizzle     (the apply-sweep box)
//   2  TMA 2-D box {32, 128} of a DENSE [rows][32] view (128-byte rows back to back: a fully contiguous 16 KB)
//   3  cp.async.bulk (1-D) of 16 KB contiguous
//   4  LDG.128, coalesced over the box of variant 0 (warp = 4 rows x 128 B), 8 loads in flight per thread, 256 threads
//   5  LDG.128, contiguous 16 KB, same
//   6  cp.async 16 B (LDGSTS) over the box of variant 0, 256 threads, groups of 4 per thread
//   7  as 4 with 512 threads
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o gpurun_out/fill_bench tools/fill_bench.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok = 0;
    long long t0 = clock64();
    while (true) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (ok) break;
        if (clock64() - t0 > 2000000000ll) __trap();
    }
}

constexpr int NS = 6;             // stages
constexpr int UB = 16384;         // bytes per unit
constexpr int NPX = 324, C = 512, NSMP = 50;

struct Params {
    CUtensorMap m0, m1, m2;
    const float* buf;
    long long* cyc;       // per-CTA cycles of the timed sweeps
    float* sink;
    int variant, sweeps, units_total;
};

__global__ void __launch_bounds__(512, 1) fill_kernel(const __grid_constant__ Params P) {
    extern __shared__ uint8_t raw[];
    __shared__ __align__(8) uint64_t s_full[NS];
    uint8_t* base = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
    const uint32_t base_u = smem_u32(base);
    const int tid = threadIdx.x, G = gridDim.x, b = blockIdx.x;
    const int u_lo = (int)((long long)P.units_total * b / G), u_hi = (int)((long long)P.units_total * (b + 1) / G);
    const int nun = u_hi - u_lo;
    const int KBT = 11;           // 32-pixel blocks per plane (324 -> 11, last one partly out of bounds)
    if (tid == 0) {
        for (int i = 0; i < NS; ++i) mbar_init(smem_u32(&s_full[i]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    float acc = 0.f;
    long long t0 = 0;
    uint32_t cnt = 0;
    for (int sw = 0; sw < P.sweeps + 1; ++sw) {
        if (sw == 1) { __syncthreads(); t0 = clock64(); }
        if (P.variant <= 3) {
            // one producer thread, NS boxes in flight; the consumer is the same thread (waits for the oldest, then reuses the slot)
            if (tid == 0) {
                for (int i = 0; i < nun + NS; ++i) {
                    if (i >= NS) { const uint32_t j = cnt - NS; mbar_wait(smem_u32(&s_full[j % NS]), (j / NS) & 1u); }
                    if (i < nun) {
                        const int u = u_lo + i;
                        const uint32_t s = cnt % NS, full = smem_u32(&s_full[s]), dst = base_u + s * UB;
                        mbar_expect_tx(full, UB);
                        if (P.variant == 0) {
                            // units: (chunk of 128 channels, sample, pixel block)
                            const int per_chunk = NSMP * KBT, chunk = u / per_chunk, r = u - chunk * per_chunk, smp = r / KBT, kb = r - smp * KBT;
                            asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                                         ::"r"(dst), "l"((uint64_t)&P.m0), "r"(full), "r"(kb * 32), "r"(chunk * 128), "r"(smp) : "memory");
                        } else if (P.variant == 1) {
                            const int per_s = 3 * 16, smp = u / per_s, r = u - smp * per_s, pt = r / 16, kb = r - pt * 16;
                            asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                                         ::"r"(dst), "l"((uint64_t)&P.m1), "r"(full), "r"(pt * 128), "r"(kb * 32), "r"(smp) : "memory");
                        } else if (P.variant == 2) {
                            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                                         ::"r"(dst), "l"((uint64_t)&P.m2), "r"(full), "r"(0), "r"(u * 128) : "memory");
                        } else {
                            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                         ::"r"(dst), "l"((uint64_t)(P.buf + (size_t)u * (UB / 4))), "r"(UB), "r"(full) : "memory");
                        }
                        ++cnt;
                    }
                }
            }
        } else if (P.variant == 4 || P.variant == 5 || P.variant == 7) {
            const int nth = (P.variant == 7) ? 512 : 256;
            if (tid < nth) {
                // a unit = 1024 float4; thread t takes float4 t, t + nth, ...: 8 loads in flight
                for (int i = 0; i < nun; ++i) {
                    const int u = u_lo + i;
                    const float* ub;
                    int pitch;       // floats between the 128-byte rows
                    if (P.variant == 5) { ub = P.buf + (size_t)u * (UB / 4); pitch = 32; }
                    else {
                        const int per_chunk = NSMP * KBT, chunk = u / per_chunk, r = u - chunk * per_chunk, smp = r / KBT, kb = min(r - smp * KBT, 9);
                        ub = P.buf + ((size_t)smp * C + chunk * 128) * NPX + kb * 32;
                        pitch = NPX;
                    }
                    for (int j0 = 0; j0 < 1024; j0 += 8 * nth) {
                        float4 v[8];
#pragma unroll
                        for (int k = 0; k < 8; ++k) {
                            const int e = j0 + k * nth + tid;
                            if (e < 1024) v[k] = __ldcg(reinterpret_cast<const float4*>(ub + (size_t)(e >> 3) * pitch) + (e & 7));
                            else v[k] = make_float4(0.f, 0.f, 0.f, 0.f);
                        }
#pragma unroll
                        for (int k = 0; k < 8; ++k) acc += v[k].x + v[k].w;
                    }
                }
            }
        } else if (P.variant == 6) {
            if (tid < 256) {
                for (int i = 0; i < nun; ++i) {
                    const int u = u_lo + i;
                    const int per_chunk = NSMP * KBT, chunk = u / per_chunk, r = u - chunk * per_chunk, smp = r / KBT, kb = min(r - smp * KBT, 9);
                    const float* ub = P.buf + ((size_t)smp * C + chunk * 128) * NPX + kb * 32;
                    const uint32_t dst = base_u + (uint32_t)(i % NS) * UB;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const int e = k * 256 + tid;
                        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + e * 16), "l"(reinterpret_cast<const float4*>(ub + (size_t)(e >> 3) * NPX) + (e & 7)) : "memory");
                    }
                    asm volatile("cp.async.commit_group;" ::: "memory");
                    asm volatile("cp.async.wait_group %0;" ::"n"(NS - 1) : "memory");
                }
                asm volatile("cp.async.wait_group 0;" ::: "memory");
            }
        }
    }
    __syncthreads();
    if (tid == 0) P.cyc[b] = clock64() - t0;
    if (acc == 123.456f) P.sink[0] = acc;
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                             const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    EncodeFn enc = (EncodeFn)fn;
    const size_t floats = (size_t)NSMP * C * NPX;
    float* buf; CK(cudaMalloc(&buf, floats * 4 + 65536)); CK(cudaMemset(buf, 0, floats * 4 + 65536));
    long long* cyc; CK(cudaMalloc(&cyc, 148 * 8));
    float* sink; CK(cudaMalloc(&sink, 4));
    Params P;
    P.buf = buf; P.cyc = cyc; P.sink = sink;
    cuuint32_t es[3] = {1, 1, 1};
    {
        cuuint64_t d[3] = {NPX, C, NSMP}; cuuint64_t s[2] = {NPX * 4, (cuuint64_t)C * NPX * 4};
        cuuint32_t b0[3] = {32, 128, 1}, b1[3] = {128, 32, 1};
        CUresult r = enc(&P.m0, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, buf, d, s, b0, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r) { printf("encode m0 %d\n", (int)r); return 1; }
        r = enc(&P.m1, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, buf, d, s, b1, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r) { printf("encode m1 %d\n", (int)r); return 1; }
        cuuint64_t d2[2] = {32, floats / 32}; cuuint64_t s2[1] = {128}; cuuint32_t b2[2] = {32, 128};
        r = enc(&P.m2, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, buf, d2, s2, b2, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r) { printf("encode m2 %d\n", (int)r); return 1; }
    }
    const int smem = NS * UB + 2048;
    CK(cudaFuncSetAttribute(fill_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    const char* names[8] = {"TMA box 32px x 128ch swizzle128", "TMA box 128px x 32ch", "TMA dense 2-D 16 KB", "cp.async.bulk 16 KB",
                            "LDG.128 strided rows, 256 thr", "LDG.128 contiguous, 256 thr", "cp.async 16 B strided rows, 256 thr", "LDG.128 strided rows, 512 thr"};
    for (int grid : {148, 74, 16}) {
        for (int v = 0; v < 8; ++v) {
            P.variant = v; P.sweeps = 10;
            P.units_total = (v == 1) ? NSMP * 3 * 16 : ((v == 0 || v == 4 || v == 6 || v == 7) ? 4 * NSMP * 11 : (int)(floats * 4 / UB));
            // the same number of units per CTA whatever the grid
            P.units_total = (int)((long long)P.units_total * grid / 148);
            fill_kernel<<<grid, 512, smem>>>(P);       // warm-up (also brings the buffer into L2)
            CK(cudaDeviceSynchronize());
            CK(cudaEventRecord(e0));
            fill_kernel<<<grid, 512, smem>>>(P);
            CK(cudaEventRecord(e1));
            CK(cudaDeviceSynchronize());
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
            std::vector<long long> h(grid);
            CK(cudaMemcpy(h.data(), cyc, grid * 8, cudaMemcpyDeviceToHost));
            long long mx = 0; double mean = 0;
            for (auto c : h) { mx = c > mx ? c : mx; mean += (double)c / grid; }
            const double units_per_cta = (double)P.units_total / grid;
            printf("grid %3d  v%d %-36s units/CTA %.1f  cycles/unit mean %.0f max %.0f  => %.1f B/clk/SM   kernel %.1f us (11 sweeps)\n", grid, v, names[v],
                   units_per_cta, mean / (P.sweeps * units_per_cta), (double)mx / (P.sweeps * units_per_cta), UB / (mean / (P.sweeps * units_per_cta)), ms * 1e3);
        }
    }
    return 0;
}
