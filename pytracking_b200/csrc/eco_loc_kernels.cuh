// ECO's score computation in the Fourier domain (the second half of its citation: eco.py:244-300).
//   reference: ECO.apply_filter      pytracking/tracker/eco/eco.py:244-245   complex.mult(filter, sample_xf).sum(1, keepdim=True)
//              ECO.localize_target   eco.py:247-252                          fourier.sample_fs(fourier.sum_fs(weight * sf), output_sz)
//              fourier.sum_fs        pytracking/libs/fourier.py:95-114       blocks added into the largest, aligned at the DC row / kx = 0
//              fourier.sample_fs     fourier.py:35-61                        zero-pad the centred half spectrum, grid_sz.prod() * irfft2
// The reference pads the series to the output grid and runs an inverse real FFT (250 x 250 for 63 x 32 non-zero coefficients); what that
// computes is the trigonometric series itself,
//     out[y, x] = sum_ky Re(F[ky,0] e^{2 pi i ky y/Oh}) + 2 sum_ky sum_{kx >= 1} Re(F[ky,kx] e^{2 pi i (ky y/Oh + kx x/Ow)})
// (the C2R transform drops the imaginary part of the kx = 0 column), evaluated here directly and separably: per CTA the fused spectrum
// F = sum_b w_b sf_b of one scale in shared memory, X[r][kx] = sum_ky F[ky][kx] e^{i phi_y} for EL_ROWS output rows, then the row sums
// over kx.  Twiddles come from two shared-memory tables e^{2 pi i m/O}, m = (k * y) mod O kept as a running integer (exact phase
// reduction, no division in the loops).
#pragma once
#include <cstddef>
#include <cstdint>

#ifndef B200_DYN_SMEM_F
#ifdef B200_CPU_EMUL
#define B200_DYN_SMEM_F(name) float* name = reinterpret_cast<float*>(::cpu_emul::dyn_smem())
#else
#define B200_DYN_SMEM_F(name) extern __shared__ float name[]
#endif
#endif

namespace b200trk {

constexpr int EL_ROWS = 8;          // output rows per CTA of eco_sample_fs_kernel
constexpr int EL_MAX_BLOCKS = 8;    // feature blocks fused by one call

// sf[s, px] = sum_c hf[c, px] * xf[s, c, px]  (complex; px = ky * Wh + kx), one thread per (s, px), channels in order
__global__ void eco_apply_filter_kernel(const float2* __restrict__ hf, const float2* __restrict__ xf, float2* __restrict__ sf, int S, int C, int HW) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= S * HW) return;
    const int s = i / HW, px = i - s * HW;
    const float2* x = xf + (size_t)s * C * HW + px;
    const float2* h = hf + px;
    float re = 0.f, im = 0.f;
    for (int c = 0; c < C; ++c) {
        const float2 a = h[(size_t)c * HW], b = x[(size_t)c * HW];
        re = fmaf(a.x, b.x, re); re = fmaf(-a.y, b.y, re);
        im = fmaf(a.x, b.y, im); im = fmaf(a.y, b.x, im);
    }
    sf[i] = make_float2(re, im);
}

struct EcoLocParams {
    const float2* sf[EL_MAX_BLOCKS];   // [S,1,H_b,Wh_b] complex, in the order fourier.sum_fs adds them (descending H_b)
    int H[EL_MAX_BLOCKS], Wh[EL_MAX_BLOCKS];
    float w[EL_MAX_BLOCKS];
    int nb, S, OH, OW;
    float* out;                        // [S,1,OH,OW]
};

// host: the blocks in the order fourier.sum_fs adds them (descending number of rows, ties in the caller's order; fourier.py:101), checked
// against what the kernel claims; nullptr or the reason for rejecting the call
inline const char* eco_loc_bind(EcoLocParams& P, const float* const* sf_blocks, const int* H, const int* Wh, const float* weights, int nb, int S,
                                int out_h, int out_w, float* out) {
    if (!sf_blocks || !H || !Wh || !out) return "null pointer";
    if (nb < 1 || nb > EL_MAX_BLOCKS || S < 1 || S > 65535) return "between 1 and 8 blocks and 1 .. 65535 scales";
    int order[EL_MAX_BLOCKS];
    for (int b = 0; b < nb; ++b) order[b] = b;
    for (int i = 1; i < nb; ++i)
        for (int j = i; j > 0 && H[order[j]] > H[order[j - 1]]; --j) { const int t = order[j]; order[j] = order[j - 1]; order[j - 1] = t; }
    for (int i = 0; i < nb; ++i) {
        const int b = order[i];
        if (!sf_blocks[b] || ((uintptr_t)sf_blocks[b] & 7) != 0) return "a block is null or not 8-byte aligned";
        if (H[b] < 1 || H[b] % 2 != 1 || Wh[b] < 1) return "a centred half spectrum has an odd number of rows";
        if (Wh[b] > Wh[order[0]]) return "a block has more columns than the block with the most rows";
        P.sf[i] = reinterpret_cast<const float2*>(sf_blocks[b]); P.H[i] = H[b]; P.Wh[i] = Wh[b]; P.w[i] = weights ? weights[b] : 1.f;
    }
    // fourier.py:43-48: only grids at least as large as the series; the equal-size case is the reference's other branch (even-sized output)
    if (out_h < P.H[0] || out_w < 2 * P.Wh[0] - 1 || (out_h == P.H[0] && out_w == 2 * P.Wh[0] - 1) || out_h >= (1 << 20) || out_w >= (1 << 20))
        return "the grid must be at least as large as the series and not equal to it";
    P.nb = nb; P.S = S; P.OH = out_h; P.OW = out_w; P.out = out;
    return nullptr;
}

// shared memory (floats): F 2*H*Wh | ty 2*OH | tx 2*OW | X 2*EL_ROWS*Wh      (H, Wh of block 0 = the largest)
inline size_t eco_sample_fs_smem_floats(int H, int Wh, int OH, int OW) { return 2 * ((size_t)H * Wh + OH + OW + (size_t)EL_ROWS * Wh); }

__global__ void __launch_bounds__(256) eco_sample_fs_kernel(EcoLocParams P) {
    B200_DYN_SMEM_F(esm);
    const int H = P.H[0], Wh = P.Wh[0], OH = P.OH, OW = P.OW;
    float2* F = reinterpret_cast<float2*>(esm);
    float2* ty = F + H * Wh;
    float2* tx = ty + OH;
    float2* X = tx + OW;
    const int s = blockIdx.y, y0 = blockIdx.x * EL_ROWS;
    const int mid = (H - 1) / 2;
    // fused spectrum of this scale: the largest block, then the others added at their centred position (fourier.py:103-112)
    for (int i = threadIdx.x; i < H * Wh; i += blockDim.x) {
        const float2 v = P.sf[0][(size_t)s * H * Wh + i];
        F[i] = make_float2(P.w[0] * v.x, P.w[0] * v.y);
    }
    for (int m = threadIdx.x; m < OH; m += blockDim.x) { float sn, cs; sincospif(2.f * (float)m / (float)OH, &sn, &cs); ty[m] = make_float2(cs, sn); }
    for (int m = threadIdx.x; m < OW; m += blockDim.x) { float sn, cs; sincospif(2.f * (float)m / (float)OW, &sn, &cs); tx[m] = make_float2(cs, sn); }
    __syncthreads();
    for (int b = 1; b < P.nb; ++b) {
        const int hb = P.H[b], wb = P.Wh[b], top = mid - (hb - 1) / 2;
        for (int i = threadIdx.x; i < hb * wb; i += blockDim.x) {
            const int r = i / wb, k = i - r * wb;
            const float2 v = P.sf[b][(size_t)s * hb * wb + i];
            float2& f = F[(top + r) * Wh + k];
            f.x += P.w[b] * v.x; f.y += P.w[b] * v.y;
        }
        __syncthreads();
    }
    // X[r][kx] = sum_ky F[ky][kx] e^{2 pi i ky y / OH}
    for (int i = threadIdx.x; i < EL_ROWS * Wh; i += blockDim.x) {
        const int r = i / Wh, kx = i - r * Wh, y = y0 + r;
        float re = 0.f, im = 0.f;
        if (y < OH) {
            int m = (int)(((long long)(OH - (mid % OH)) * y) % OH);      // (-mid * y) mod OH
            for (int k = 0; k < H; ++k) {
                const float2 f = F[k * Wh + kx], t = ty[m];
                re = fmaf(f.x, t.x, re); re = fmaf(-f.y, t.y, re);
                im = fmaf(f.x, t.y, im); im = fmaf(f.y, t.x, im);
                m += y; if (m >= OH) m -= OH;
            }
        }
        X[i] = make_float2(re, im);
    }
    __syncthreads();
    // out[y][x] = Re X[r][0] + 2 sum_{kx >= 1} Re( X[r][kx] e^{2 pi i kx x / OW} )
    for (int x = threadIdx.x; x < OW; x += blockDim.x) {
        float acc[EL_ROWS];
#pragma unroll
        for (int r = 0; r < EL_ROWS; ++r) acc[r] = 0.f;
        int m = x;                                                       // kx = 1
        for (int kx = 1; kx < Wh; ++kx) {
            const float2 t = tx[m];
#pragma unroll
            for (int r = 0; r < EL_ROWS; ++r) {
                const float2 v = X[r * Wh + kx];
                acc[r] = fmaf(v.x, t.x, acc[r]); acc[r] = fmaf(-v.y, t.y, acc[r]);
            }
            m += x; if (m >= OW) m -= OW;
        }
#pragma unroll
        for (int r = 0; r < EL_ROWS; ++r)
            if (y0 + r < OH) P.out[((size_t)s * OH + y0 + r) * OW + x] = fmaf(2.f, acc[r], X[r * Wh].x);
    }
}

// ---- ECO.preprocess_sample (eco.py:297-300) for one block: x *= window (in place, as the reference), cfft2 (fourier.py:20-25),
// dcf.interpolate_dft (dcf.py:96-102).  One CTA per (sample, channel) map; the H x W map is tiny (62 x 62 at most in ECO), so the
// real-to-complex transform is evaluated directly and separably in shared memory:
//     T[y][kx] = sum_x v[y][x] e^{-2 pi i kx x / W},   xf[ky][kx] = sum_y T[y][kx] e^{-2 pi i ky y / H},   ky = -(H'-1)/2 .. H'/2
// (H' = H + (H+1)%2 rows: for an even H the Nyquist row appears at both ends, exactly what rfftshift2 produces), then
// (xf * interp_y[ky]) * interp_x[kx] in the reference's order.  Twiddle tables e^{2 pi i m / N} in shared memory, indices as running integers.
// shared memory (floats): v H*W | T 2*H*Wh' | tx 2*W | ty 2*H
inline size_t eco_preprocess_smem_floats(int H, int W) { return (size_t)H * W + (H * W) % 2 + 2 * ((size_t)H * (W / 2 + 1) + W + H); }

// x is addressed through its strides (in elements): the tracker hands over a permuted view of the projection's [H,W,S,C] result (eco.py:304-309)
__global__ void __launch_bounds__(256) eco_preprocess_kernel(float* __restrict__ x, const float* __restrict__ window, const float2* __restrict__ iy,
                                                             const float2* __restrict__ ix, float2* __restrict__ xf, int C, int H, int W,
                                                             long long st_s, long long st_c, long long st_y, long long st_x) {
    B200_DYN_SMEM_F(psm);
    const int HP = H + (H + 1) % 2, WH = W / 2 + 1;
    float* v = psm;
    float2* T = reinterpret_cast<float2*>(psm + H * W + (H * W) % 2);      // 8-byte aligned
    float2* tx = T + H * WH;
    float2* ty = tx + W;
    float* xm = x + (long long)(blockIdx.x / C) * st_s + (long long)(blockIdx.x % C) * st_c;
    for (int i = threadIdx.x; i < H * W; i += blockDim.x) {
        const int y = i / W, xx = i - y * W;
        float* e = xm + y * st_y + xx * st_x;
        const float a = *e * window[i];
        v[i] = a; *e = a;
    }
    for (int m = threadIdx.x; m < W; m += blockDim.x) { float sn, cs; sincospif(2.f * (float)m / (float)W, &sn, &cs); tx[m] = make_float2(cs, sn); }
    for (int m = threadIdx.x; m < H; m += blockDim.x) { float sn, cs; sincospif(2.f * (float)m / (float)H, &sn, &cs); ty[m] = make_float2(cs, sn); }
    __syncthreads();
    for (int i = threadIdx.x; i < H * WH; i += blockDim.x) {
        const int y = i / WH, kx = i - y * WH;
        const float* row = v + y * W;
        float re = 0.f, im = 0.f;
        int m = 0;
        for (int xx = 0; xx < W; ++xx) {
            const float2 t = tx[m];
            re = fmaf(row[xx], t.x, re); im = fmaf(-row[xx], t.y, im);
            m += kx; if (m >= W) m -= W;
        }
        T[i] = make_float2(re, im);
    }
    __syncthreads();
    float2* out = xf + (size_t)blockIdx.x * HP * WH;
    for (int i = threadIdx.x; i < HP * WH; i += blockDim.x) {
        const int q = i / WH, kx = i - q * WH;
        const int ky = q - (HP - 1) / 2;
        const int step = ((ky % H) + H) % H;
        float re = 0.f, im = 0.f;
        int m = 0;
        for (int y = 0; y < H; ++y) {                                      // T * conj(t)
            const float2 a = T[y * WH + kx], t = ty[m];
            re = fmaf(a.x, t.x, re); re = fmaf(a.y, t.y, re);
            im = fmaf(a.y, t.x, im); im = fmaf(-a.x, t.y, im);
            m += step; if (m >= H) m -= H;
        }
        const float2 wy = iy[q], wx = ix[kx];
        const float r1 = re * wy.x - im * wy.y, i1 = re * wy.y + im * wy.x;       // complex.mult (complex.py:14-32), twice
        out[i] = make_float2(r1 * wx.x - i1 * wx.y, r1 * wx.y + i1 * wx.x);
    }
}

// ---- fourier.shift_fs (fourier.py:78-92): out = (a * e^{i shift_y ky}) * e^{i shift_x kx}, one thread per coefficient; the phases formed as the
// reference forms them (the shift as a float times the float frequency), the two complex products in its order
__global__ void eco_shift_fs_kernel(const float2* __restrict__ a, float2* __restrict__ out, long long total, int H, int Wh, float sy, float sx) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int kx = (int)(i % Wh), q = (int)((i / Wh) % H);
    const float py = sy * (float)(q - (H - 1) / 2), px = sx * (float)kx;
    const float cy = cosf(py), sny = sinf(py), cx = cosf(px), snx = sinf(px);
    const float2 v = a[i];
    const float r1 = v.x * cy - v.y * sny, i1 = v.x * sny + v.y * cy;
    out[i] = make_float2(r1 * cx - i1 * snx, r1 * snx + i1 * cx);
}

}  // namespace b200trk
