// IoUNet box refinement on the device:
//   AtomIoUNet.predict_iou            ltr/models/bbreg/atom_iou_net.py:96-136   (modulation, PrRoIPool 5x5 / 3x3, fc3_rt / fc4_rt, iou_predictor)
//   DiMP.optimize_boxes_default       pytracking/tracker/dimp/dimp.py:725-751   (gradient ascent on the boxes, step * grad * [w,h,w,h])
//   DiMP.optimize_boxes_relative      pytracking/tracker/dimp/dimp.py:754-793   (the same in rect_to_rel space, ltr/data/bounding_box_utils.py:4-30)
// The reference obtains d iou / d box by autograd (outputs.backward through LinearBlock, PrRoIPool2DFunction.backward); here the
// gradient is written out: relu masks -> the folded linear layers transposed -> PrRoIPool coordinate backward (prroi.cu) -> xywh.
// Everything stays on the device between the iterations; the host reads the final boxes and IoUs once.
//
// Algebra used: PrRoIPool is linear in the features, so pooling the raw IoU features and scaling the pooled channels by the
// modulation vector equals pooling the modulated features (atom_iou_net.py:109-110); eval-mode BatchNorm of a LinearBlock
// (ltr/models/layers/blocks.py:24-40) is folded into the linear layer in double precision at create time.
#include "common.cuh"
#include <cmath>
#include <vector>

using namespace b200trk;

struct b200trk_iou_predictor {
    int C3 = 0, P3 = 0, C4 = 0, P4 = 0, D3 = 0, D4 = 0;
    float *w3 = nullptr, *b3 = nullptr, *w4 = nullptr, *b4 = nullptr, *wp = nullptr;   // folded [D3][C3*P3*P3], [D3], ..., [D3+D4]
    float bp = 0.f;
    // per-call scratch (R <= RMAX)
    float *rois = nullptr, *pool3 = nullptr, *pool4 = nullptr, *act = nullptr, *gpool3 = nullptr, *gpool4 = nullptr;
    float *grois3 = nullptr, *grois4 = nullptr, *rel = nullptr, *sznorm = nullptr, *step = nullptr, *part = nullptr;
    std::vector<void*> owned;
};

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
    extern __shared__ float sg[];                 // [D][RMAX] : wp[j] * relu'(a[r][j])
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

int iou_alloc(b200trk_iou_predictor* p, float** q, size_t floats) {
    void* d = nullptr;
    B200_CHECK_CUDA(cudaMalloc(&d, floats * sizeof(float)));
    p->owned.push_back(d);
    *q = (float*)d;
    return 0;
}

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

extern "C" int b200trk_iou_predictor_create(b200trk_iou_predictor_t** out, const b200trk_linear_block_t* fc3_rt, const b200trk_linear_block_t* fc4_rt,
                                            const float* iou_predictor_weight, const float* iou_predictor_bias, int C3, int P3, int C4, int P4,
                                            int D3, int D4) {
    B200_REQUIRE(out && fc3_rt && fc4_rt && iou_predictor_weight && fc3_rt->weight && fc4_rt->weight, "iou_predictor_create: null pointer");
    B200_REQUIRE(C3 > 0 && C4 > 0 && P3 > 0 && P4 > 0 && D3 > 0 && D4 > 0 && D3 <= 1024 && D4 <= 1024, "iou_predictor_create: bad dimensions");
    b200trk_iou_predictor* p = new b200trk_iou_predictor();
    p->C3 = C3; p->P3 = P3; p->C4 = C4; p->P4 = P4; p->D3 = D3; p->D4 = D4;
    const int K3 = C3 * P3 * P3, K4 = C4 * P4 * P4;
    std::vector<float> w, b;
    int e = 0;
    auto up = [&](float** dst, const std::vector<float>& v) {
        if (e) return;
        e = iou_alloc(p, dst, v.size());
        if (!e && cudaMemcpy(*dst, v.data(), v.size() * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess) { set_error("iou_predictor_create: upload failed"); e = 1; }
    };
    fold_linear(*fc3_rt, D3, K3, w, b); up(&p->w3, w); up(&p->b3, b);
    fold_linear(*fc4_rt, D4, K4, w, b); up(&p->w4, w); up(&p->b4, b);
    std::vector<float> wp(iou_predictor_weight, iou_predictor_weight + D3 + D4);
    up(&p->wp, wp);
    p->bp = iou_predictor_bias ? iou_predictor_bias[0] : 0.f;
    if (!e) e = iou_alloc(p, &p->rois, IOU_RMAX * 5);
    if (!e) e = iou_alloc(p, &p->pool3, (size_t)IOU_RMAX * K3);
    if (!e) e = iou_alloc(p, &p->pool4, (size_t)IOU_RMAX * K4);
    if (!e) e = iou_alloc(p, &p->gpool3, (size_t)IOU_RMAX * K3);
    if (!e) e = iou_alloc(p, &p->gpool4, (size_t)IOU_RMAX * K4);
    if (!e) e = iou_alloc(p, &p->act, (size_t)IOU_RMAX * (D3 + D4));
    if (!e) e = iou_alloc(p, &p->part, (size_t)((K3 > K4 ? K3 : K4) / FC_KS + 1) * IOU_RMAX * (D3 + D4));
    if (!e) e = iou_alloc(p, &p->grois3, IOU_RMAX * 5);
    if (!e) e = iou_alloc(p, &p->grois4, IOU_RMAX * 5);
    if (!e) e = iou_alloc(p, &p->rel, IOU_RMAX * 4);
    if (!e) e = iou_alloc(p, &p->sznorm, 4);
    if (!e) e = iou_alloc(p, &p->step, 4);
    if (e) { b200trk_iou_predictor_destroy(p); return e; }
    *out = p;
    return 0;
}

extern "C" int b200trk_iou_predictor_destroy(b200trk_iou_predictor_t* p) {
    if (!p) return 0;
    for (void* q : p->owned) cudaFree(q);
    delete p;
    return 0;
}

// one evaluation of predict_iou (+ its box gradient, + an optional ascent step) for the R boxes in `boxes`
static int iou_eval(b200trk_iou_predictor* p, const float* mod3, const float* mod4, const float* feat3, int H3, int W3, const float* feat4,
                    int H4, int W4, float* boxes, int R, float* iou_out, float* grad_out, int step_mode, float decay, bool need_grad,
                    b200trk_stream_t stream) {
    cudaStream_t st = (cudaStream_t)stream;
    const int K3 = p->C3 * p->P3 * p->P3, K4 = p->C4 * p->P4 * p->P4, D = p->D3 + p->D4;
    make_rois_kernel<<<1, 32, 0, st>>>(boxes, p->rois, R);
    B200_LAUNCH_CHECK();
    if (int e = b200trk_prroi_pool_forward(feat3, p->rois, p->pool3, 1, p->C3, H3, W3, R, p->P3, p->P3, 1.f / 8.f, stream)) return e;
    if (int e = b200trk_prroi_pool_forward(feat4, p->rois, p->pool4, 1, p->C4, H4, W4, R, p->P4, p->P4, 1.f / 16.f, stream)) return e;
    const int ns3 = (K3 + FC_KS - 1) / FC_KS, ns4 = (K4 + FC_KS - 1) / FC_KS, nb3 = (p->D3 + 7) / 8, nb4 = (p->D4 + 7) / 8;
    fc_forward_kernel<IOU_RMAX><<<nb3 * ns3 + nb4 * ns4, 256, 0, st>>>(p->w3, p->w4, p->pool3, p->pool4, mod3, mod4, p->part, R, K3, p->P3 * p->P3, K4,
                                                                      p->P4 * p->P4, p->D3, p->D4, nb3, ns3, ns4);
    B200_LAUNCH_CHECK();
    iou_head_kernel<<<1, 32 * IOU_RMAX, 0, st>>>(p->part, p->b3, p->b4, p->act, p->wp, p->bp, iou_out, R, p->D3, p->D4, ns3, ns4, IOU_RMAX);
    B200_LAUNCH_CHECK();
    (void)D;
    if (!need_grad) return 0;
    const int nb = (K3 + 255) / 256 + (K4 + 255) / 256;
    const size_t smem = (size_t)(p->D3 > p->D4 ? p->D3 : p->D4) * IOU_RMAX * sizeof(float);
    B200_CHECK_CUDA(cudaFuncSetAttribute(fc_backward_kernel<IOU_RMAX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    fc_backward_kernel<IOU_RMAX><<<nb, 256, smem, st>>>(p->w3, p->w4, p->act, p->wp, mod3, mod4, p->gpool3, p->gpool4, R, K3, p->P3 * p->P3,
                                                       K4, p->P4 * p->P4, p->D3, p->D4);
    B200_LAUNCH_CHECK();
    if (int e = b200trk_prroi_pool_coor_backward(feat3, p->rois, p->pool3, p->gpool3, p->grois3, 1, p->C3, H3, W3, R, p->P3, p->P3, 1.f / 8.f, stream)) return e;
    if (int e = b200trk_prroi_pool_coor_backward(feat4, p->rois, p->pool4, p->gpool4, p->grois4, 1, p->C4, H4, W4, R, p->P4, p->P4, 1.f / 16.f, stream)) return e;
    box_step_kernel<<<1, 32, 0, st>>>(p->grois3, p->grois4, boxes, p->rel, p->sznorm, grad_out, p->step, decay, R, step_mode);
    B200_LAUNCH_CHECK();
    return 0;
}

extern "C" int b200trk_iou_predict(b200trk_iou_predictor_t* p, const float* mod3, const float* mod4, const float* feat3, int H3, int W3,
                                   const float* feat4, int H4, int W4, const float* proposals, int R, float* iou_out, float* grad_out,
                                   b200trk_stream_t stream) {
    B200_REQUIRE(p && mod3 && mod4 && feat3 && feat4 && proposals && iou_out, "iou_predict: null pointer");
    B200_REQUIRE(R >= 1 && R <= IOU_RMAX, "iou_predict: R=%d proposals (1..%d)", R, IOU_RMAX);
    return iou_eval(p, mod3, mod4, feat3, H3, W3, feat4, H4, W4, const_cast<float*>(proposals), R, iou_out, grad_out, 0, 1.f, grad_out != nullptr, stream);
}

extern "C" int b200trk_iou_refine(b200trk_iou_predictor_t* p, const float* mod3, const float* mod4, const float* feat3, int H3, int W3,
                                  const float* feat4, int H4, int W4, float* boxes, int R, int num_iter, float step_length, float step_decay,
                                  int relative, float* iou_out, b200trk_stream_t stream) {
    B200_REQUIRE(p && mod3 && mod4 && feat3 && feat4 && boxes && iou_out, "iou_refine: null pointer");
    B200_REQUIRE(R >= 1 && R <= IOU_RMAX, "iou_refine: R=%d boxes (1..%d)", R, IOU_RMAX);
    B200_REQUIRE(num_iter >= 1 && num_iter <= 64, "iou_refine: num_iter=%d", num_iter);
    cudaStream_t st = (cudaStream_t)stream;
    set_scalar_kernel<<<1, 1, 0, st>>>(p->step, step_length);
    B200_LAUNCH_CHECK();
    if (relative) {
        to_rel_kernel<<<1, 32, 0, st>>>(boxes, p->rel, p->sznorm, R);
        B200_LAUNCH_CHECK();
    }
    // the reference returns the IoUs predicted in the LAST iteration's forward pass (before its step), dimp.py:737-751
    for (int it = 0; it < num_iter; ++it)
        if (int e = iou_eval(p, mod3, mod4, feat3, H3, W3, feat4, H4, W4, boxes, R, it == num_iter - 1 ? iou_out : nullptr, nullptr,
                             relative ? 2 : 1, step_decay, true, stream)) return e;
    return 0;
}
