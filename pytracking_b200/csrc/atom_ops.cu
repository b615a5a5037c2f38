// ATOM stage-2 operators:
//   feature_normalize   MultiFeatureBase.get_feature p-norm normalisation (pytracking/features/featurebase.py:105-108)
//   conv1x1             operation.conv1x1 / project_sample (pytracking/libs/operation.py:35-42, atom.py:427-431)
//   conv2d_same         operation.conv2d(mode='same') with ONE 4x4 filter = ATOM.apply_filter (atom.py:301-302)
//   fourier_interp      ATOM.localize_target's cfft2 -> shift_fs -> sum_fs -> sample_fs chain (atom.py:304-316,
//                       pytracking/libs/fourier.py:20-92): the Fourier-series upsampling of the score map, evaluated
//                       directly as  out = Dy * s * Dx^T / (H*W)  with the (2K+1)-term Dirichlet-type kernels the chain
//                       implies (both Nyquist rows/columns of an even-sized map are kept, exactly as cfft2 / irfft do).
#include "common.cuh"
#include <cmath>
#include <map>
#include <mutex>
#include <tuple>
#include <vector>

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
    extern __shared__ float fsm[];
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

// D[Yo, y] = sum_{k=-K..K} cos(k * (2 pi (Yo/O - y/N) + shift)),  K = floor(N/2),  shift = pi (1 - (ksz % 2)/N)
static float* interp_table(int N, int O, int ksz) {
    static std::map<std::tuple<int, int, int, int>, float*> cache;
    static std::mutex mu;
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lk(mu);
    auto key = std::make_tuple(dev, N, O, ksz % 2);
    auto it = cache.find(key);
    if (it != cache.end()) return it->second;
    std::vector<float> h((size_t)O * N);
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
    float* d = nullptr;
    if (cudaMalloc(&d, h.size() * sizeof(float)) != cudaSuccess) return nullptr;
    if (cudaMemcpy(d, h.data(), h.size() * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess) { cudaFree(d); return nullptr; }
    cache[key] = d;
    return d;
}

}  // namespace b200trk

using namespace b200trk;

extern "C" int b200trk_feature_normalize(float* feat, int S, int C, int H, int W, float normalize_power, b200trk_stream_t stream) {
    B200_REQUIRE(feat, "feature_normalize: null pointer");
    B200_REQUIRE(S > 0 && C > 0 && H > 0 && W > 0 && normalize_power > 0.f, "feature_normalize: bad shape / power");
    feature_normalize_kernel<<<S, 512, 0, (cudaStream_t)stream>>>(feat, C * H * W, normalize_power);
    B200_LAUNCH_CHECK();
    return 0;
}

extern "C" int b200trk_conv1x1(const float* x, const float* P, float* out, int S, int Cin, int Cout, int H, int W,
                               b200trk_stream_t stream) {
    B200_REQUIRE(x && P && out, "conv1x1: null pointer");
    B200_REQUIRE(S > 0 && Cin > 0 && Cout > 0 && H > 0 && W > 0, "conv1x1: empty input");
    const int HW = H * W;
    conv1x1_kernel<<<dim3((HW + 63) / 64, (Cout + 63) / 64, S), 256, 0, (cudaStream_t)stream>>>(x, P, out, Cin, Cout, HW);
    B200_LAUNCH_CHECK();
    return 0;
}

extern "C" int b200trk_fourier_interp(const float* scores, float* out, int S, int H, int W, int ksz_h, int ksz_w, int out_h,
                                      int out_w, b200trk_stream_t stream) {
    B200_REQUIRE(scores && out, "fourier_interp: null pointer");
    B200_REQUIRE(S > 0 && H > 0 && W > 0 && H <= 64 && W <= 64, "fourier_interp: score map %dx%d not supported (<= 64x64)", H, W);
    B200_REQUIRE(out_h >= H + 1 - (H & 1) && out_w >= W + 1 - (W & 1) && out_h <= 4096 && out_w <= 4096,
                 "fourier_interp: output grid %dx%d must not be smaller than the Fourier series of a %dx%d map", out_h, out_w, H, W);
    float* Dy = interp_table(H, out_h, ksz_h);
    float* Dx = interp_table(W, out_w, ksz_w);
    B200_REQUIRE(Dy && Dx, "fourier_interp: could not build the interpolation tables");
    const size_t smem = (size_t)(H * W + FI_ROWS * W) * sizeof(float);
    fourier_interp_kernel<<<dim3((out_h + FI_ROWS - 1) / FI_ROWS, S), 256, smem, (cudaStream_t)stream>>>(
        scores, Dy, Dx, out, H, W, out_h, out_w, 1.0f / (float)(H * W));
    B200_LAUNCH_CHECK();
    return 0;
}

extern "C" int b200trk_softmax_reg(const float* x, float* out, int n, int L, int has_reg, float reg, b200trk_stream_t stream) {
    B200_REQUIRE(x && out, "softmax_reg: null pointer");
    B200_REQUIRE(n > 0 && L > 0, "softmax_reg: empty input");
    softmax_reg_kernel<<<n, 256, 0, (cudaStream_t)stream>>>(x, out, L, has_reg ? 1 : 0, reg);
    B200_LAUNCH_CHECK();
    return 0;
}
