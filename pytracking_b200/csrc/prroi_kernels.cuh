// The device kernels of prroi.cu (Precise RoI Pooling forward / backward / coordinate backward; derivation and references in prroi.cu's
// header comment).  Plain SIMT CUDA C in a header of their own so that the SAME source also compiles as host code under
// tests/cpu_emul/cuda_shim.h (tests/test_prroi_kernels_cpu.py).  Included by prroi.cu only.
#pragma once

namespace b200trk {

constexpr int PR_MAXW = 64;     // max cells covered by one bin along one axis (+2)

__device__ __forceinline__ float hat_cdf(float t) {   // integral_{-inf}^{t} max(0,1-|u|) du
    if (t <= -1.f) return 0.f;
    if (t <= 0.f) { const float a = t + 1.f; return 0.5f * a * a; }
    if (t < 1.f) { const float a = 1.f - t; return 1.f - 0.5f * a * a; }
    return 1.f;
}
__device__ __forceinline__ float hat(float t) { t = fabsf(t); return t < 1.f ? 1.f - t : 0.f; }

struct BinGeom {
    int b;                  // batch index
    float x0, x1, y0, y1;   // bin window in feature cells
    float win;              // window area
    int ws, nw, hs, nh;     // covered integer range [ws, ws+nw), [hs, hs+nh) clipped to the map
};

__device__ __forceinline__ BinGeom bin_geom(const float* rois, int r, int ph, int pw, int PH, int PW, float scale, int H, int W) {
    const float* roi = rois + 5 * r;
    BinGeom g;
    g.b = (int)roi[0];
    const float rx0 = roi[1] * scale, ry0 = roi[2] * scale, rx1 = roi[3] * scale, ry1 = roi[4] * scale;
    const float rw = fmaxf(rx1 - rx0, 0.f), rh = fmaxf(ry1 - ry0, 0.f);
    const float bw = rw / (float)PW, bh = rh / (float)PH;
    g.x0 = rx0 + bw * pw; g.y0 = ry0 + bh * ph;
    g.x1 = g.x0 + bw; g.y1 = g.y0 + bh;
    g.win = fmaxf(0.f, bw * bh);
    int ws = (int)floorf(g.x0), we = (int)ceilf(g.x1);     // cells ws..we (inclusive) can carry weight
    int hs = (int)floorf(g.y0), he = (int)ceilf(g.y1);
    ws = max(ws, 0); hs = max(hs, 0); we = min(we, W - 1); he = min(he, H - 1);
    g.ws = ws; g.nw = max(0, we - ws + 1); g.hs = hs; g.nh = max(0, he - hs + 1);
    return g;
}

// ---- forward ------------------------------------------------------------------------------------------
__global__ void prroi_forward_kernel(const float* __restrict__ feat, const float* __restrict__ rois, float* __restrict__ out,
                                     int C, int H, int W, int PH, int PW, float scale) {
    __shared__ float sWx[PR_MAXW], sWy[PR_MAXW];
    __shared__ BinGeom sg;
    const int bin = blockIdx.x, r = blockIdx.y, ph = bin / PW, pw = bin % PW;
    if (threadIdx.x == 0) sg = bin_geom(rois, r, ph, pw, PH, PW, scale, H, W);
    __syncthreads();
    const BinGeom g = sg;
    for (int i = threadIdx.x; i < g.nw; i += blockDim.x) sWx[i] = hat_cdf(g.x1 - (float)(g.ws + i)) - hat_cdf(g.x0 - (float)(g.ws + i));
    for (int i = threadIdx.x; i < g.nh; i += blockDim.x) sWy[i] = hat_cdf(g.y1 - (float)(g.hs + i)) - hat_cdf(g.y0 - (float)(g.hs + i));
    __syncthreads();
    const float inv = g.win > 0.f ? 1.f / g.win : 0.f;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        const float* f = feat + ((size_t)g.b * C + c) * H * W;
        float acc = 0.f;
        for (int i = 0; i < g.nh; ++i) {
            const float* row = f + (size_t)(g.hs + i) * W + g.ws;
            float ra = 0.f;
            for (int j = 0; j < g.nw; ++j) ra = fmaf(row[j], sWx[j], ra);
            acc = fmaf(ra, sWy[i], acc);
        }
        out[(((size_t)r * C + c) * PH + ph) * PW + pw] = acc * inv;
    }
}

// ---- backward w.r.t. features (scatter; atomics as in the reference) -----------------------------------------
__global__ void prroi_backward_kernel(const float* __restrict__ rois, const float* __restrict__ ograd, float* __restrict__ fgrad,
                                      int C, int H, int W, int PH, int PW, float scale) {
    __shared__ float sWx[PR_MAXW], sWy[PR_MAXW];
    __shared__ BinGeom sg;
    const int bin = blockIdx.x, r = blockIdx.y, ph = bin / PW, pw = bin % PW;
    if (threadIdx.x == 0) sg = bin_geom(rois, r, ph, pw, PH, PW, scale, H, W);
    __syncthreads();
    const BinGeom g = sg;
    for (int i = threadIdx.x; i < g.nw; i += blockDim.x) sWx[i] = hat_cdf(g.x1 - (float)(g.ws + i)) - hat_cdf(g.x0 - (float)(g.ws + i));
    for (int i = threadIdx.x; i < g.nh; i += blockDim.x) sWy[i] = hat_cdf(g.y1 - (float)(g.hs + i)) - hat_cdf(g.y0 - (float)(g.hs + i));
    __syncthreads();
    if (g.win <= 0.f) return;
    const float inv = 1.f / g.win;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        const float go = ograd[(((size_t)r * C + c) * PH + ph) * PW + pw] * inv;
        float* f = fgrad + ((size_t)g.b * C + c) * H * W;
        for (int i = 0; i < g.nh; ++i)
            for (int j = 0; j < g.nw; ++j) {
                const float wgt = sWy[i] * sWx[j];
                if (wgt != 0.f) atomicAdd(f + (size_t)(g.hs + i) * W + g.ws + j, go * wgt);
            }
    }
}

// ---- backward w.r.t. the RoI coordinates ---------------------------------------------------------------------------
// stage 1: per (roi, bin) partial d/d(x1,y1,x2,y2) summed over channels -> part[r][bin][4]
__global__ void prroi_coor_backward_kernel(const float* __restrict__ feat, const float* __restrict__ rois,
                                           const float* __restrict__ outp, const float* __restrict__ ograd,
                                           float* __restrict__ part, int C, int H, int W, int PH, int PW, float scale) {
    __shared__ float sWx[PR_MAXW], sWy[PR_MAXW], sTx0[PR_MAXW], sTx1[PR_MAXW], sTy0[PR_MAXW], sTy1[PR_MAXW];
    __shared__ BinGeom sg;
    __shared__ float red[32];
    const int bin = blockIdx.x, r = blockIdx.y, ph = bin / PW, pw = bin % PW;
    if (threadIdx.x == 0) sg = bin_geom(rois, r, ph, pw, PH, PW, scale, H, W);
    __syncthreads();
    const BinGeom g = sg;
    for (int i = threadIdx.x; i < g.nw; i += blockDim.x) {
        const float w = (float)(g.ws + i);
        sWx[i] = hat_cdf(g.x1 - w) - hat_cdf(g.x0 - w);
        sTx0[i] = hat(g.x0 - w);
        sTx1[i] = hat(g.x1 - w);
    }
    for (int i = threadIdx.x; i < g.nh; i += blockDim.x) {
        const float h = (float)(g.hs + i);
        sWy[i] = hat_cdf(g.y1 - h) - hat_cdf(g.y0 - h);
        sTy0[i] = hat(g.y0 - h);
        sTy1[i] = hat(g.y1 - h);
    }
    __syncthreads();
    float gx1 = 0.f, gy1 = 0.f, gx2 = 0.f, gy2 = 0.f;
    if (g.win > 0.f) {
        const float inv = 1.f / g.win;
        const float dy = g.y1 - g.y0, dx = g.x1 - g.x0;
        for (int c = threadIdx.x; c < C; c += blockDim.x) {
            const size_t oi = (((size_t)r * C + c) * PH + ph) * PW + pw;
            const float go = ograd[oi];
            if (go * inv == 0.f) continue;                          // impl.cu:325-327
            const float top = outp[oi];
            const float* f = feat + ((size_t)g.b * C + c) * H * W;
            float e_x0 = 0.f, e_x1 = 0.f, e_y0 = 0.f, e_y1 = 0.f;    // edge line integrals
            for (int i = 0; i < g.nh; ++i) {
                const float* row = f + (size_t)(g.hs + i) * W + g.ws;
                float a0 = 0.f, a1 = 0.f, aw = 0.f;
                for (int j = 0; j < g.nw; ++j) {
                    const float v = row[j];
                    a0 = fmaf(v, sTx0[j], a0);
                    a1 = fmaf(v, sTx1[j], a1);
                    aw = fmaf(v, sWx[j], aw);
                }
                e_x0 = fmaf(a0, sWy[i], e_x0);
                e_x1 = fmaf(a1, sWy[i], e_x1);
                e_y0 = fmaf(aw, sTy0[i], e_y0);
                e_y1 = fmaf(aw, sTy1[i], e_y1);
            }
            // impl.cu:357-366
            const float px1 = (-e_x0 + dy * top) * inv * scale;
            const float py1 = (-e_y0 + dx * top) * inv * scale;
            const float px2 = (e_x1 - dy * top) * inv * scale;
            const float py2 = (e_y1 - dx * top) * inv * scale;
            // impl.cu:370-377
            const float fw0 = (float)pw / PW, fw1 = (float)(pw + 1) / PW, fh0 = (float)ph / PH, fh1 = (float)(ph + 1) / PH;
            gx1 += (px1 * (1.f - fw0) + px2 * (1.f - fw1)) * go;
            gy1 += (py1 * (1.f - fh0) + py2 * (1.f - fh1)) * go;
            gx2 += (px2 * fw1 + px1 * fw0) * go;
            gy2 += (py2 * fh1 + py1 * fh0) * go;
        }
    }
    gx1 = block_sum(gx1, red); gy1 = block_sum(gy1, red); gx2 = block_sum(gx2, red); gy2 = block_sum(gy2, red);
    if (threadIdx.x == 0) {
        float* p = part + ((size_t)r * PH * PW + bin) * 4;
        p[0] = gx1; p[1] = gy1; p[2] = gx2; p[3] = gy2;
    }
}

// stage 2: rois_grad[r] = (0, sum over bins)
__global__ void prroi_coor_reduce_kernel(const float* __restrict__ part, float* __restrict__ rgrad, int R, int nbins) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= R) return;
    float a[4] = {0.f, 0.f, 0.f, 0.f};
    for (int b = 0; b < nbins; ++b)
        for (int q = 0; q < 4; ++q) a[q] += part[((size_t)r * nbins + b) * 4 + q];
    rgrad[5 * r] = 0.f;
    for (int q = 0; q < 4; ++q) rgrad[5 * r + 1 + q] = a[q];
}

}  // namespace b200trk
