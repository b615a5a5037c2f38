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

#include "atom_ops_kernels.cuh"      // the four kernels + interp_table_host

namespace b200trk {

// the table of interp_table_host on the current device, built once per geometry
static float* interp_table(int N, int O, int ksz) {
    static std::map<std::tuple<int, int, int, int>, float*> cache;
    static std::mutex mu;
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lk(mu);
    auto key = std::make_tuple(dev, N, O, ksz % 2);
    auto it = cache.find(key);
    if (it != cache.end()) return it->second;
    std::vector<float> h;
    interp_table_host(N, O, ksz, h);
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
