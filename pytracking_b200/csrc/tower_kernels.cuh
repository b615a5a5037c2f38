// The CUDA-core kernels of tower.cu (ToMP DenseBoxRegressor: attention-scaled NCHW -> NHWC import, GroupNorm(1, C) + ReLU, exp export;
// references in tower.cu's header comment).  Plain SIMT CUDA C in a header of their own so that the SAME source also compiles as host code
// under tests/cpu_emul/cuda_shim.h (tests/test_tomp_kernels_cpu.py).  Included by tower.cu only.
#pragma once

namespace b200trk {

// NCHW [S,C,HW] * attention [S,HW] -> NHWC [S,HW,C]
__global__ void import_scaled_kernel(const float* __restrict__ in, const float* __restrict__ att, float* __restrict__ out, int HW, int C) {
    __shared__ float tile[32][33];
    const int s = blockIdx.z, p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    const float* src = in + (size_t)s * HW * C;
    float* dst = out + (size_t)s * HW * C;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int c = c0 + i, p = p0 + threadIdx.x;
        tile[i][threadIdx.x] = (p < HW && c < C) ? src[(size_t)c * HW + p] * (att ? att[(size_t)s * HW + p] : 1.f) : 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int p = p0 + i, c = c0 + threadIdx.x;
        if (p < HW && c < C) dst[(size_t)p * C + c] = tile[threadIdx.x][i];
    }
}

// GroupNorm(1, C) + ReLU in place on NHWC x [S][HW][C]: y = relu((x - mean) * rsqrt(var + eps) * gamma[c] + beta[c]), biased variance
__global__ void __launch_bounds__(1024) groupnorm1_relu_kernel(float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                                                              int HW, int C, float eps) {
    __shared__ float red[32];
    __shared__ float s_mean, s_rstd;
    float* xs = x + (size_t)blockIdx.x * HW * C;
    const int n = HW * C;
    float a = 0.f;
    for (int i = threadIdx.x; i < n; i += blockDim.x) a += xs[i];
    a = block_sum(a, red);
    if (threadIdx.x == 0) s_mean = a / (float)n;
    __syncthreads();
    const float mean = s_mean;
    float v = 0.f;
    for (int i = threadIdx.x; i < n; i += blockDim.x) { const float d = xs[i] - mean; v += d * d; }
    v = block_sum(v, red);
    if (threadIdx.x == 0) s_rstd = rsqrtf(v / (float)n + eps);
    __syncthreads();
    const float rstd = s_rstd;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const int c = i % C;
        xs[i] = fmaxf((xs[i] - mean) * rstd * gamma[c] + beta[c], 0.f);
    }
}

// NHWC [S][HW][C] -> NCHW with exp (C = 4)
__global__ void export_exp_kernel(const float* __restrict__ in, float* __restrict__ out, int HW, int C, int S) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= S * HW * C) return;
    const int c = i % C, p = (i / C) % HW, s = i / (C * HW);
    out[((size_t)s * C + c) * HW + p] = expf(in[i]);
}

}  // namespace b200trk
