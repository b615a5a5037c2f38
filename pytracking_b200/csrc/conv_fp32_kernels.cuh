// The fp32 CUDA-core kernels of stage 1 (conv_fp32.cu): preprocessing, the 7x7 stem, max-pool, the generic implicit-GEMM convolution
// with its deterministic split-K epilogue (the exact-fp32 path), layout conversion and InstanceL2Norm.  Plain SIMT CUDA C in a header of
// their own so that the SAME source also compiles as host code under tests/cpu_emul/cuda_shim.h (tests/test_conv_fp32_kernels_cpu.py).
// Included by conv_fp32.cu only.
#pragma once
#include "conv_fp32.cuh"

#ifdef B200_CPU_EMUL
#define B200_DYN_SMEM_F4(name) float4* name = reinterpret_cast<float4*>(::cpu_emul::dyn_smem())
#else
#define B200_DYN_SMEM_F4(name) extern __shared__ float4 name[]
#endif

namespace b200trk {

// --------------------------------------------------------------------------------------------------
// preprocess: NetWithBackbone.preprocess_image (pytracking/features/net_wrappers.py:55-69), NCHW 0..255 -> NHWC4
// --------------------------------------------------------------------------------------------------
__global__ void preprocess_kernel(const float* __restrict__ crop, float4* __restrict__ out, int S, int HW) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= S * HW) return;
    const int s = idx / HW, p = idx - s * HW;
    const float* base = crop + (size_t)s * 3 * HW + p;
    // same operation order as the reference: im/255; im -= mean; im /= std  (IEEE division, no fast-math)
    const float r = (base[0] / 255.f - 0.485f) / 0.229f;
    const float g = (base[HW] / 255.f - 0.456f) / 0.224f;
    const float b = (base[2 * HW] / 255.f - 0.406f) / 0.225f;
    out[idx] = make_float4(r, g, b, 0.f);
}

// --------------------------------------------------------------------------------------------------
// stem: conv 7x7 stride 2 pad 3, Cin = 3, Cout = 64, + folded BN bias + ReLU (ltr/models/backbone/resnet.py:182-184).
// CTA = 8x8 output pixels x 64 channels, 8 warps: warp = (16-channel group, 32-pixel half), lane = pixel, so every
// weight fetch is a warp-wide broadcast (one shared-memory wavefront) and every input fetch is conflict free: the
// 21x21 input patch is stored as [channel][column parity][row][12] so that the stride-2 column walk of a warp touches
// consecutive words (row pitch 12: the four pixel rows of a warp land in disjoint bank octets).
// --------------------------------------------------------------------------------------------------
constexpr int STEM_T = 8, STEM_P = (STEM_T - 1) * 2 + 7;   // 21x21 input patch
constexpr int STEM_RP = 12;                                // row pitch of a parity plane (11 used)
constexpr int STEM_PLANE = STEM_P * STEM_RP;               // floats per (channel, parity) plane
constexpr int STEM_PATCH_FLOATS = 3 * 2 * STEM_PLANE;
constexpr int STEM_W_FLOATS = 49 * 3 * 64;                 // [tap][cin][cout]

__global__ void __launch_bounds__(256)
stem_kernel(const float4* __restrict__ in, const float* __restrict__ wt, const float* __restrict__ bias,
            float* __restrict__ out, int Hin, int Win, int Hout, int Wout) {
    B200_DYN_SMEM_F4(sm4);
    float* patch = reinterpret_cast<float*>(sm4);          // [3][2][21][12]
    float* ws = patch + STEM_PATCH_FLOATS;                 // [49][3][64]
    const int s = blockIdx.z, oy0 = blockIdx.y * STEM_T, ox0 = blockIdx.x * STEM_T;
    const int iy0 = oy0 * 2 - 3, ix0 = ox0 * 2 - 3;
    for (int i = threadIdx.x; i < STEM_P * STEM_P; i += 256) {
        const int py = i / STEM_P, px = i - py * STEM_P;
        const int iy = iy0 + py, ix = ix0 + px;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (iy >= 0 && iy < Hin && ix >= 0 && ix < Win) v = in[((size_t)s * Hin + iy) * Win + ix];
        const int o = (px & 1) * STEM_PLANE + py * STEM_RP + (px >> 1);
        patch[o] = v.x; patch[2 * STEM_PLANE + o] = v.y; patch[4 * STEM_PLANE + o] = v.z;
    }
    for (int i = threadIdx.x; i < STEM_W_FLOATS / 4; i += 256)
        reinterpret_cast<float4*>(ws)[i] = __ldg(reinterpret_cast<const float4*>(wt) + i);
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int cg = (warp & 3) * 16;                        // first output channel of this warp
    const int py = (warp >> 2) * 4 + (lane >> 3), px = lane & 7;
    float acc[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) acc[j] = 0.f;
    for (int kh = 0; kh < 7; ++kh) {
        const float* prow = patch + (py * 2 + kh) * STEM_RP + px;
#pragma unroll
        for (int kw = 0; kw < 7; ++kw) {
            const float* xp = prow + (kw & 1) * STEM_PLANE + (kw >> 1);
            const float x0 = xp[0], x1 = xp[2 * STEM_PLANE], x2 = xp[4 * STEM_PLANE];
            const float4* wp = reinterpret_cast<const float4*>(ws + (kh * 7 + kw) * 192 + cg);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float4 w0 = wp[j], w1 = wp[16 + j], w2 = wp[32 + j];
                acc[4 * j + 0] = fmaf(x0, w0.x, acc[4 * j + 0]); acc[4 * j + 1] = fmaf(x0, w0.y, acc[4 * j + 1]);
                acc[4 * j + 2] = fmaf(x0, w0.z, acc[4 * j + 2]); acc[4 * j + 3] = fmaf(x0, w0.w, acc[4 * j + 3]);
                acc[4 * j + 0] = fmaf(x1, w1.x, acc[4 * j + 0]); acc[4 * j + 1] = fmaf(x1, w1.y, acc[4 * j + 1]);
                acc[4 * j + 2] = fmaf(x1, w1.z, acc[4 * j + 2]); acc[4 * j + 3] = fmaf(x1, w1.w, acc[4 * j + 3]);
                acc[4 * j + 0] = fmaf(x2, w2.x, acc[4 * j + 0]); acc[4 * j + 1] = fmaf(x2, w2.y, acc[4 * j + 1]);
                acc[4 * j + 2] = fmaf(x2, w2.z, acc[4 * j + 2]); acc[4 * j + 3] = fmaf(x2, w2.w, acc[4 * j + 3]);
            }
        }
    }
    const int oy = oy0 + py, ox = ox0 + px;
    if (oy < Hout && ox < Wout) {
        float* o = out + (((size_t)s * Hout + oy) * Wout + ox) * 64 + cg;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float4 b = __ldg(reinterpret_cast<const float4*>(bias + cg) + j);
            float4 v;
            v.x = fmaxf(acc[4 * j + 0] + b.x, 0.f); v.y = fmaxf(acc[4 * j + 1] + b.y, 0.f);
            v.z = fmaxf(acc[4 * j + 2] + b.z, 0.f); v.w = fmaxf(acc[4 * j + 3] + b.w, 0.f);
            reinterpret_cast<float4*>(o)[j] = v;
        }
    }
}

// --------------------------------------------------------------------------------------------------
// max-pool 3x3 stride 2 pad 1 (resnet.py:189), NHWC, implicit -inf padding like torch
// --------------------------------------------------------------------------------------------------
__global__ void maxpool_kernel(const float4* __restrict__ in, float4* __restrict__ out, int S, int Hin, int Win,
                               int Hout, int Wout, int C4) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const int total = S * Hout * Wout * C4;
    if (idx >= total) return;
    const int c = idx % C4;
    int p = idx / C4;
    const int ox = p % Wout; p /= Wout;
    const int oy = p % Hout;
    const int s = p / Hout;
    float4 m = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
#pragma unroll
    for (int dy = 0; dy < 3; ++dy) {
        const int iy = oy * 2 - 1 + dy;
        if (iy < 0 || iy >= Hin) continue;
#pragma unroll
        for (int dx = 0; dx < 3; ++dx) {
            const int ix = ox * 2 - 1 + dx;
            if (ix < 0 || ix >= Win) continue;
            const float4 v = in[(((size_t)s * Hin + iy) * Win + ix) * C4 + c];
            m.x = fmaxf(m.x, v.x); m.y = fmaxf(m.y, v.y); m.z = fmaxf(m.z, v.z); m.w = fmaxf(m.w, v.w);
        }
    }
    out[idx] = m;
}

// --------------------------------------------------------------------------------------------------
// generic implicit-GEMM convolution: D[m][n] = sum_k A[m][k] * W[n][k], m = output pixel, k = (kh,kw,cin)
// CTA tile 64x64x16, 64 threads, 8x8 register micro-tile (4 FMA per shared-memory word), register-prefetched
// double buffering. gridDim.z = split-K factor (partials are reduced in a fixed order by splitk_epilogue_kernel).
// --------------------------------------------------------------------------------------------------
constexpr int CB = 64, CK = 16, CLD = CB + 4, CTHREADS = 64;

template <bool FUSED>
__global__ void __launch_bounds__(CTHREADS)
conv_igemm_kernel(const float* __restrict__ in, const float* __restrict__ w, float* __restrict__ out,
                  ConvShape sh, ConvEpilogue ep, int ksteps_per_split) {
    __shared__ __align__(16) float As[2][CK][CLD];
    __shared__ __align__(16) float Bs[2][CK][CLD];
    const int M = sh.M(), Kt = sh.K();
    const int m0 = blockIdx.x * CB, n0 = blockIdx.y * CB;
    const int cblks = sh.Cin / CK;
    const int total_ks = sh.k * sh.k * cblks;
    const int ks_begin = blockIdx.z * ksteps_per_split;
    const int ks_end = min(total_ks, ks_begin + ksteps_per_split);
    const int t = threadIdx.x, kq = t & 3, r0 = t >> 2;

    int a_iy0[4], a_ix0[4];
    const float* a_base[4];
    bool a_ok[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int m = m0 + r0 + 16 * r;
        a_ok[r] = m < M;
        const int mm = a_ok[r] ? m : 0;
        const int ox = mm % sh.Wout;
        const int tmp = mm / sh.Wout;
        const int oy = tmp % sh.Hout;
        const int s = tmp / sh.Hout;
        a_iy0[r] = oy * sh.stride - sh.pad;
        a_ix0[r] = ox * sh.stride - sh.pad;
        a_base[r] = in + (size_t)s * sh.Hin * sh.Win * sh.Cin;
    }
    float4 ra[4], rb[4];
    auto load = [&](int ks) {
        const int tap = ks / cblks, cb = ks - tap * cblks;
        const int kh = tap / sh.k, kw = tap - kh * sh.k;
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int iy = a_iy0[r] + kh, ix = a_ix0[r] + kw;
            const bool ok = a_ok[r] && iy >= 0 && iy < sh.Hin && ix >= 0 && ix < sh.Win;
            ra[r] = ok ? __ldg(reinterpret_cast<const float4*>(a_base[r] + ((size_t)iy * sh.Win + ix) * sh.Cin + cb * CK + kq * 4))
                       : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int n = n0 + r0 + 16 * r;
            rb[r] = (n < sh.Cout) ? __ldg(reinterpret_cast<const float4*>(w + (size_t)n * Kt + (size_t)ks * CK + kq * 4))
                                  : make_float4(0.f, 0.f, 0.f, 0.f);
        }
    };
    auto store = [&](int stage) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int row = r0 + 16 * r;
            As[stage][kq * 4 + 0][row] = ra[r].x; As[stage][kq * 4 + 1][row] = ra[r].y;
            As[stage][kq * 4 + 2][row] = ra[r].z; As[stage][kq * 4 + 3][row] = ra[r].w;
            Bs[stage][kq * 4 + 0][row] = rb[r].x; Bs[stage][kq * 4 + 1][row] = rb[r].y;
            Bs[stage][kq * 4 + 2][row] = rb[r].z; Bs[stage][kq * 4 + 3][row] = rb[r].w;
        }
    };

    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
    const int ty = t >> 3, tx = t & 7;

    if (ks_begin < ks_end) {
        load(ks_begin);
        store(0);
    }
    __syncthreads();
    for (int ks = ks_begin; ks < ks_end; ++ks) {
        const int cur = (ks - ks_begin) & 1;
        const bool more = ks + 1 < ks_end;
        if (more) load(ks + 1);
#pragma unroll
        for (int kk = 0; kk < CK; ++kk) {
            const float4 a0 = *reinterpret_cast<const float4*>(&As[cur][kk][ty * 4]);
            const float4 a1 = *reinterpret_cast<const float4*>(&As[cur][kk][32 + ty * 4]);
            const float4 b0 = *reinterpret_cast<const float4*>(&Bs[cur][kk][tx * 4]);
            const float4 b1 = *reinterpret_cast<const float4*>(&Bs[cur][kk][32 + tx * 4]);
            const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        if (more) store(cur ^ 1);
        __syncthreads();
    }

    float* dst = FUSED ? out : out + (size_t)blockIdx.z * M * sh.Cout;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int m = m0 + ((i < 4) ? ty * 4 + i : 32 + ty * 4 + (i - 4));
        if (m >= M) continue;
#pragma unroll
        for (int jh = 0; jh < 2; ++jh) {
            const int n = n0 + jh * 32 + tx * 4;
            if (n >= sh.Cout) continue;
            float4 v = make_float4(acc[i][jh * 4 + 0], acc[i][jh * 4 + 1], acc[i][jh * 4 + 2], acc[i][jh * 4 + 3]);
            if (FUSED) {
                if (ep.bias) {
                    const float4 b = __ldg(reinterpret_cast<const float4*>(ep.bias + n));
                    v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
                }
                if (ep.residual) {
                    const float4 r = __ldg(reinterpret_cast<const float4*>(ep.residual + (size_t)m * sh.Cout + n));
                    v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w;
                }
                if (ep.relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
            }
            *reinterpret_cast<float4*>(dst + (size_t)m * sh.Cout + n) = v;
        }
    }
}

__global__ void splitk_epilogue_kernel(const float4* __restrict__ part, float4* __restrict__ out, int splits,
                                       int MN4, int Cout4, ConvEpilogue ep) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= MN4) return;
    float4 v = part[idx];
    for (int z = 1; z < splits; ++z) {
        const float4 p = part[(size_t)z * MN4 + idx];
        v.x += p.x; v.y += p.y; v.z += p.z; v.w += p.w;
    }
    if (ep.bias) {
        const float4 b = __ldg(reinterpret_cast<const float4*>(ep.bias) + (idx % Cout4));
        v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
    }
    if (ep.residual) {
        const float4 r = __ldg(reinterpret_cast<const float4*>(ep.residual) + idx);
        v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w;
    }
    if (ep.relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
    out[idx] = v;
}

// --------------------------------------------------------------------------------------------------
// layout conversion NHWC -> NCHW (32x32 smem transpose), optionally scaled per sample
// --------------------------------------------------------------------------------------------------
__global__ void nhwc_to_nchw_kernel(const float* __restrict__ in, float* __restrict__ out, int HW, int C,
                                    const float* __restrict__ partials, int nparts, float scale, float eps) {
    __shared__ float tile[32][33];
    __shared__ float s_factor;
    const int s = blockIdx.z;
    if (partials) {
        if (threadIdx.x == 0 && threadIdx.y == 0) {
            float tot = 0.f;
            for (int i = 0; i < nparts; ++i) tot += partials[s * nparts + i];
            // InstanceL2Norm (normalization.py:15-18): scale * sqrt(C*H*W / (sum + eps))
            s_factor = scale * sqrtf(((float)C * (float)HW) / (tot + eps));
        }
        __syncthreads();
    }
    const float f = partials ? s_factor : 1.f;
    const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    const float* src = in + (size_t)s * HW * C;
    float* dst = out + (size_t)s * HW * C;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int p = p0 + i, c = c0 + threadIdx.x;
        tile[i][threadIdx.x] = (p < HW && c < C) ? src[(size_t)p * C + c] : 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int c = c0 + i, p = p0 + threadIdx.x;
        if (p < HW && c < C) dst[(size_t)c * HW + p] = partials ? tile[threadIdx.x][i] * f : tile[threadIdx.x][i];
    }
}

constexpr int L2_PARTS = 64;

__global__ void sumsq_kernel(const float4* __restrict__ in, float* __restrict__ partials, int n4_per_sample) {
    __shared__ float red[32];
    const int s = blockIdx.y;
    const float4* src = in + (size_t)s * n4_per_sample;
    const int per = (n4_per_sample + L2_PARTS - 1) / L2_PARTS;
    const int b = blockIdx.x * per, e = min(n4_per_sample, b + per);
    float acc = 0.f;
    for (int i = b + threadIdx.x; i < e; i += blockDim.x) {
        const float4 v = src[i];
        acc += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
    }
    acc = block_sum(acc, red);
    if (threadIdx.x == 0) partials[s * L2_PARTS + blockIdx.x] = acc;
}

}  // namespace b200trk
