// The device kernels of atom_ops.cu (feature_normalize, softmax_reg, conv1x1, fourier_interp; references in atom_ops.cu's header comment)
// and the host computation of the Fourier interpolation tables.  Plain SIMT CUDA C in a header of their own so that the SAME source also
// compiles as host code under tests/cpu_emul/cuda_shim.h (tests/test_atom_ops_kernels_cpu.py).  Included by atom_ops.cu only.
#pragma once
#include <cmath>
#include <vector>

#ifdef B200_CPU_EMUL
#define B200_DYN_SMEM_F(name) float* name = reinterpret_cast<float*>(::cpu_emul::dyn_smem())
#else
#define B200_DYN_SMEM_F(name) extern __shared__ float name[]
#endif

namespace b200trk {

// --------------------------------------------------------------------------------------------------
// feature_normalize: x /= (sum |x|^p / (C*H*W) + 1e-10)^(1/p), one CTA per sample (in place)
// --------------------------------------------------------------------------------------------------
__global__ void feature_normalize_kernel(float* __restrict__ x, int per_sample, float p) {
    __shared__ float red[32];
    float* xs = x + (size_t)blockIdx.x * per_sample;
    float acc = 0.f;
    if (p == 2.f) {
        for (int i = threadIdx.x; i < per_sample; i += blockDim.x) { const float v = xs[i]; acc += v * v; }
    } else {
        for (int i = threadIdx.x; i < per_sample; i += blockDim.x) acc += powf(fabsf(xs[i]), p);
    }
    const float tot = block_sum(acc, red);
    const float mean = tot / (float)per_sample + 1e-10f;
    const float denom = (p == 2.f) ? sqrtf(mean) : powf(mean, 1.f / p);
    for (int i = threadIdx.x; i < per_sample; i += blockDim.x) xs[i] = xs[i] / denom;
}

// --------------------------------------------------------------------------------------------------
// softmax_reg over the last dimension with one extra constant logit in the denominator
// (ltr/models/layers/activation.py:7-16; PrDiMP score pre-processing, pytracking/tracker/dimp/dimp.py:206-210)
// --------------------------------------------------------------------------------------------------
__global__ void softmax_reg_kernel(const float* __restrict__ x, float* __restrict__ y, int L, int has_reg, float reg) {
    __shared__ float red[32];
    const float* xr = x + (size_t)blockIdx.x * L;
    float m = has_reg ? reg : -INFINITY;
    for (int i = threadIdx.x; i < L; i += blockDim.x) m = fmaxf(m, xr[i]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
    __syncthreads();
    m = red[0];
    for (int w = 1; w < (int)(blockDim.x + 31) / 32; ++w) m = fmaxf(m, red[w]);
    __syncthreads();
    float s = 0.f;
    for (int i = threadIdx.x; i < L; i += blockDim.x) s += expf(xr[i] - m);
    s = block_sum(s, red);
    const float den = s + (has_reg ? expf(reg - m) : 0.f);
    for (int i = threadIdx.x; i < L; i += blockDim.x) y[(size_t)blockIdx.x * L + i] = expf(xr[i] - m) / den;
}

// --------------------------------------------------------------------------------------------------
// conv1x1 on NCHW: out[s,co,p] = sum_ci P[co,ci] x[s,ci,p].  CTA tile: 64 output channels x 64 pixels, K step 16.
// --------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) conv1x1_kernel(const float* __restrict__ x, const float* __restrict__ P,
                                                      float* __restrict__ out, int Cin, int Cout, int HW) {
    __shared__ float Ps[16][65];   // [k][co]
    __shared__ float Xs[16][65];   // [k][pix]
    const int s = blockIdx.z, co0 = blockIdx.y * 64, p0 = blockIdx.x * 64;
    const float* xs = x + (size_t)s * Cin * HW;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;      // 4 co x 4 pix per thread
    float acc[4][4] = {};
    for (int k0 = 0; k0 < Cin; k0 += 16) {
        for (int i = threadIdx.x; i < 16 * 64; i += 256) {
            const int kk = i & 15, c = i >> 4;
            Ps[kk][c] = (co0 + c < Cout && k0 + kk < Cin) ? P[(size_t)(co0 + c) * Cin + k0 + kk] : 0.f;
        }
        for (int i = threadIdx.x; i < 16 * 64; i += 256) {
            const int pp = i & 63, kk = i >> 6;
            Xs[kk][pp] = (p0 + pp < HW && k0 + kk < Cin) ? xs[(size_t)(k0 + kk) * HW + p0 + pp] : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < 16; ++kk) {
            float a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) { a[i] = Ps[kk][ty * 4 + i]; b[i] = Xs[kk][tx * 4 + i]; }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int co = co0 + ty * 4 + i, pp = p0 + tx * 4 + j;
            if (co < Cout && pp < HW) out[((size_t)s * Cout + co) * HW + pp] = acc[i][j];
        }
}

// --------------------------------------------------------------------------------------------------
// Fourier-series upsampling: out[s,Y,X] = (1/(H*W)) sum_y Dy[Y,y] sum_x s[s,y,x] Dx[X,x]
// --------------------------------------------------------------------------------------------------
constexpr int FI_ROWS = 8;

__global__ void fourier_interp_kernel(const float* __restrict__ sc, const float* __restrict__ Dy, const float* __restrict__ Dx,
                                      float* __restrict__ out, int H, int W, int OH, int OW, float scale) {
    B200_DYN_SMEM_F(fsm);
    float* ssm = fsm;                 // [H][W] score map
    float* T = fsm + H * W;           // [FI_ROWS][W]
    const int s = blockIdx.y, Y0 = blockIdx.x * FI_ROWS;
    for (int i = threadIdx.x; i < H * W; i += blockDim.x) ssm[i] = sc[(size_t)s * H * W + i];
    __syncthreads();
    for (int i = threadIdx.x; i < FI_ROWS * W; i += blockDim.x) {
        const int r = i / W, x = i - r * W;
        float a = 0.f;
        if (Y0 + r < OH) {
            const float* d = Dy + (size_t)(Y0 + r) * H;
            for (int y = 0; y < H; ++y) a = fmaf(d[y], ssm[y * W + x], a);
        }
        T[i] = a;
    }
    __syncthreads();
    for (int X = threadIdx.x; X < OW; X += blockDim.x) {
        const float* d = Dx + (size_t)X * W;
        float acc[FI_ROWS];
#pragma unroll
        for (int r = 0; r < FI_ROWS; ++r) acc[r] = 0.f;
        for (int x = 0; x < W; ++x) {
            const float dv = d[x];
#pragma unroll
            for (int r = 0; r < FI_ROWS; ++r) acc[r] = fmaf(T[r * W + x], dv, acc[r]);
        }
#pragma unroll
        for (int r = 0; r < FI_ROWS; ++r)
            if (Y0 + r < OH) out[((size_t)s * OH + Y0 + r) * OW + X] = acc[r] * scale;
    }
}

// D[Yo, y] = sum_{k=-K..K} cos(k * (2 pi (Yo/O - y/N) + shift)),  K = floor(N/2),  shift = pi (1 - (ksz % 2)/N); double precision (host)
inline void interp_table_host(int N, int O, int ksz, std::vector<float>& h) {
    h.resize((size_t)O * N);
    const double pi = 3.14159265358979323846;
    const double shift = pi * (1.0 - (double)(ksz % 2) / N);
    const int K = N / 2;
    for (int Yo = 0; Yo < O; ++Yo)
        for (int y = 0; y < N; ++y) {
            const double ph = 2.0 * pi * ((double)Yo / O - (double)y / N) + shift;
            double a = 1.0;
            for (int k = 1; k <= K; ++k) a += 2.0 * std::cos(k * ph);
            h[(size_t)Yo * N + y] = (float)a;
        }
}

}  // namespace b200trk
