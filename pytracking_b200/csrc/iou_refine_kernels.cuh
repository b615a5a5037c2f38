// The device kernels of iou_refine.cu (IoUNet predict_iou forward, its analytic box gradient and the box ascent step; references and
// algebra in iou_refine.cu's header comment) and the host-side folding of a LinearBlock's BatchNorm.  Plain SIMT CUDA C in a header of their
// own so that the SAME source also compiles as host code under tests/cpu_emul/cuda_shim.h (tests/test_iou_kernels_cpu.py).  Included by
// iou_refine.cu only.
#pragma once
#include <cmath>
#include <vector>

#ifndef B200_DYN_SMEM_F
#ifdef B200_CPU_EMUL
#define B200_DYN_SMEM_F(name) float* name = reinterpret_cast<float*>(::cpu_emul::dyn_smem())
#else
#define B200_DYN_SMEM_F(name) extern __shared__ float name[]
#endif
#endif

constexpr int IOU_RMAX = 16;

namespace {

// boxes (x, y, w, h) -> rois (0, x0, y0, x1, y1) (atom_iou_net.py:119-125, one image)
__global__ void make_rois_kernel(const float* __restrict__ boxes, float* __restrict__ rois, int R) {
    const int r = threadIdx.x;
    if (r >= R) return;
    const float x = boxes[4 * r], y = boxes[4 * r + 1], w = boxes[4 * r + 2], h = boxes[4 * r + 3];
    rois[5 * r] = 0.f; rois[5 * r + 1] = x; rois[5 * r + 2] = y; rois[5 * r + 3] = x + w; rois[5 * r + 4] = y + h;
}

// fc3_rt / fc4_rt forward: z[r][j] = sum_k W[j][k] * mod[k / PP] * pooled[r][k].  The 8.9 MB of weights are the only real traffic:
// CTA = (8 neurons, one K slice of FC_KS elements); the slice of the R pooled vectors is staged (already modulated) in shared memory
// and each warp streams its neuron's weight slice once, coalesced, against all R boxes.  Partial sums per K slice go to `part`
// [slices][RMAX][D3 + D4]; `iou_head_kernel` adds them in slice order (deterministic), applies bias + ReLU and the final linear layer.
constexpr int FC_KS = 640;          // K slice: 6400 = 10 x 640, 2304 = 3.6 x 640
template <int RMAX>
__global__ void __launch_bounds__(256) fc_forward_kernel(const float* __restrict__ w3, const float* __restrict__ w4, const float* __restrict__ pool3,
                                                         const float* __restrict__ pool4, const float* __restrict__ mod3,
                                                         const float* __restrict__ mod4, float* __restrict__ part, int R, int K3, int PP3,
                                                         int K4, int PP4, int D3, int D4, int nb3, int ns3, int ns4) {
    __shared__ float xs[RMAX][FC_KS];
    const bool lvl4 = (int)blockIdx.x >= nb3 * ns3;
    const int bid = lvl4 ? (int)blockIdx.x - nb3 * ns3 : (int)blockIdx.x;
    const int ns = lvl4 ? ns4 : ns3;
    const int jb = bid / ns, ks = bid - jb * ns;
    const int K = lvl4 ? K4 : K3, PP = lvl4 ? PP4 : PP3, D = lvl4 ? D4 : D3;
    const float* pool = lvl4 ? pool4 : pool3;
    const float* mod = lvl4 ? mod4 : mod3;
    const int k0 = ks * FC_KS, klen = min(FC_KS, K - k0);
    for (int i = threadIdx.x; i < RMAX * FC_KS; i += blockDim.x) {
        const int r = i / FC_KS, k = i - r * FC_KS;
        xs[r][k] = (r < R && k < klen) ? pool[(size_t)r * K + k0 + k] * mod[(k0 + k) / PP] : 0.f;
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int j = jb * 8 + warp;
    if (j >= D) return;
    const float* w = (lvl4 ? w4 : w3) + (size_t)j * K + k0;
    float acc[RMAX];
#pragma unroll
    for (int r = 0; r < RMAX; ++r) acc[r] = 0.f;
    for (int k = lane; k < klen; k += 32) {
        const float wv = w[k];
#pragma unroll
        for (int r = 0; r < RMAX; ++r) acc[r] = fmaf(wv, xs[r][k], acc[r]);
    }
    const int Dall = D3 + D4, jg = (lvl4 ? D3 : 0) + j;
#pragma unroll
    for (int r = 0; r < RMAX; ++r) {
        const float v = warp_sum(acc[r]);
        if (lane == 0 && r < R) part[((size_t)ks * RMAX + r) * Dall + jg] = v;
    }
}

// a[r][j] = relu(b[j] + sum over the K slices); iou[r] = bp + sum_j wp[j] * a[r][j]  (iou_predictor, atom_iou_net.py:134); one warp per box
__global__ void iou_head_kernel(const float* __restrict__ part, const float* __restrict__ b3, const float* __restrict__ b4, float* __restrict__ act,
                                const float* __restrict__ wp, float bp, float* __restrict__ iou, int R, int D3, int D4, int ns3, int ns4, int RMAX) {
    const int r = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (r >= R) return;
    const int D = D3 + D4;
    float s = 0.f;
    for (int j = lane; j < D; j += 32) {
        const int ns = j < D3 ? ns3 : ns4;
        float z = j < D3 ? b3[j] : b4[j - D3];
        for (int k = 0; k < ns; ++k) z += part[((size_t)k * RMAX + r) * D + j];
        z = fmaxf(z, 0.f);
        act[(size_t)r * D + j] = z;
        s = fmaf(wp[j], z, s);
    }
    s = warp_sum(s);
    if (lane == 0 && iou) iou[r] = s + bp;
}

// d iou[r] / d pooled[r][k] = mod[k / PP] * sum_j wp[j] * [a[r][j] > 0] * W[j][k]; thread = k (coalesced over the rows of W)
template <int RMAX>
__global__ void __launch_bounds__(256) fc_backward_kernel(const float* __restrict__ w3, const float* __restrict__ w4, const float* __restrict__ act,
                                                          const float* __restrict__ wp, const float* __restrict__ mod3,
                                                          const float* __restrict__ mod4, float* __restrict__ gpool3,
                                                          float* __restrict__ gpool4, int R, int K3, int PP3, int K4, int PP4, int D3, int D4) {
    B200_DYN_SMEM_F(sg);                          // [D][RMAX] : wp[j] * relu'(a[r][j])
    const int nb3 = (K3 + 255) / 256;
    const bool lvl4 = (int)blockIdx.x >= nb3;
    const int D = lvl4 ? D4 : D3, joff = lvl4 ? D3 : 0, K = lvl4 ? K4 : K3, PP = lvl4 ? PP4 : PP3;
    for (int i = threadIdx.x; i < D * RMAX; i += blockDim.x) {
        const int j = i / RMAX, r = i - j * RMAX;
        sg[i] = (r < R && act[(size_t)r * (D3 + D4) + joff + j] > 0.f) ? wp[joff + j] : 0.f;
    }
    __syncthreads();
    const int k = (lvl4 ? (int)blockIdx.x - nb3 : (int)blockIdx.x) * 256 + threadIdx.x;
    if (k >= K) return;
    const float* w = lvl4 ? w4 : w3;
    float acc[RMAX];
#pragma unroll
    for (int r = 0; r < RMAX; ++r) acc[r] = 0.f;
    for (int j = 0; j < D; ++j) {
        const float wv = w[(size_t)j * K + k];
#pragma unroll
        for (int r = 0; r < RMAX; ++r) acc[r] = fmaf(wv, sg[j * RMAX + r], acc[r]);
    }
    const float m = (lvl4 ? mod4 : mod3)[k / PP];
    float* g = lvl4 ? gpool4 : gpool3;
#pragma unroll
    for (int r = 0; r < RMAX; ++r)
        if (r < R) g[(size_t)r * K + k] = acc[r] * m;
}

// rois_grad (., x0, y0, x1, y1) of both levels -> d iou / d (x, y, w, h); optionally one ascent step on the boxes
// mode 0: grad only; 1: default space (dimp.py:745); 2: relative space (dimp.py:781 with rect_to_rel / rel_to_rect)
__global__ void box_step_kernel(const float* __restrict__ g3, const float* __restrict__ g4, float* __restrict__ boxes, float* __restrict__ rel,
                                const float* __restrict__ sznorm, float* __restrict__ grad_out, float* __restrict__ step, float decay, int R,
                                int mode) {
    const int r = threadIdx.x;
    if (r < R) {
        const float gx0 = g3[5 * r + 1] + g4[5 * r + 1], gy0 = g3[5 * r + 2] + g4[5 * r + 2];
        const float gx1 = g3[5 * r + 3] + g4[5 * r + 3], gy1 = g3[5 * r + 4] + g4[5 * r + 4];
        const float gx = gx0 + gx1, gy = gy0 + gy1, gw = gx1, gh = gy1;            // x1 = x + w, y1 = y + h
        if (grad_out) { grad_out[4 * r] = gx; grad_out[4 * r + 1] = gy; grad_out[4 * r + 2] = gw; grad_out[4 * r + 3] = gh; }
        const float s = step ? step[0] : 0.f;
        if (mode == 1) {
            const float w = boxes[4 * r + 2], h = boxes[4 * r + 3];
            boxes[4 * r] += s * gx * w; boxes[4 * r + 1] += s * gy * h; boxes[4 * r + 2] += s * gw * w; boxes[4 * r + 3] += s * gh * h;
        } else if (mode == 2) {
            // bb = rel_to_rect(rel): sz = exp(rel[2:]), tl = rel[:2] * sz_norm - sz / 2
            const float sw = expf(rel[4 * r + 2]), sh = expf(rel[4 * r + 3]);
            const float grx = gx * sznorm[0], gry = gy * sznorm[1];
            const float grw = (gw - 0.5f * gx) * sw, grh = (gh - 0.5f * gy) * sh;
            rel[4 * r] += s * grx; rel[4 * r + 1] += s * gry; rel[4 * r + 2] += s * grw; rel[4 * r + 3] += s * grh;
            const float nw = expf(rel[4 * r + 2]), nh = expf(rel[4 * r + 3]);
            boxes[4 * r] = rel[4 * r] * sznorm[0] - 0.5f * nw; boxes[4 * r + 1] = rel[4 * r + 1] * sznorm[1] - 0.5f * nh;
            boxes[4 * r + 2] = nw; boxes[4 * r + 3] = nh;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0 && step && mode != 0) step[0] *= decay;
}

// rect_to_rel of the initial boxes with sz_norm = size of box 0 (dimp.py:761-762)
__global__ void to_rel_kernel(const float* __restrict__ boxes, float* __restrict__ rel, float* __restrict__ sznorm, int R) {
    const int r = threadIdx.x;
    const float nw = boxes[2], nh = boxes[3];
    if (r == 0) { sznorm[0] = nw; sznorm[1] = nh; }
    if (r >= R) return;
    const float x = boxes[4 * r], y = boxes[4 * r + 1], w = boxes[4 * r + 2], h = boxes[4 * r + 3];
    rel[4 * r] = (x + 0.5f * w) / nw; rel[4 * r + 1] = (y + 0.5f * h) / nh; rel[4 * r + 2] = logf(w); rel[4 * r + 3] = logf(h);
}

__global__ void set_scalar_kernel(float* p, float v) { p[0] = v; }

// LinearBlock = linear (+bias) -> eval BatchNorm -> ReLU : fold BN into (W, b) in double precision
void fold_linear(const b200trk_linear_block_t& L, int out, int in, std::vector<float>& w, std::vector<float>& b) {
    w.resize((size_t)out * in); b.resize(out);
    for (int j = 0; j < out; ++j) {
        double sc = 1.0, sh = 0.0;
        if (L.bn_gamma) {
            sc = (double)L.bn_gamma[j] / std::sqrt((double)L.bn_var[j] + 1e-5);
            sh = (double)L.bn_beta[j] - (double)L.bn_mean[j] * sc;
        }
        b[j] = (float)((L.bias ? (double)L.bias[j] : 0.0) * sc + sh);
        for (int k = 0; k < in; ++k) w[(size_t)j * in + k] = (float)((double)L.weight[(size_t)j * in + k] * sc);
    }
}

}  // namespace
