// The kernels of atom_gn.cu (ATOM's first-frame GaussNewtonCG on FactorizedConvProblem as a stream of small kernels with device-resident
// CG scalars; derivation and references in atom_gn.cu's header comment).  Plain SIMT CUDA C in a header of their own so that the SAME source
// also compiles as host code under tests/cpu_emul/cuda_shim.h (tests/test_atom_gn_kernels_cpu.py).  Included by atom_gn.cu only.
#pragma once

namespace b200trk {

__device__ __forceinline__ float gn_act(float s, int kind, float a) {
    if (kind == 1) return fmaxf(s, 0.f);
    if (kind == 2) return s > 0.f ? s : (expf(s) - 1.f);
    if (kind == 3) return s >= 0.f ? s : a * (expf(s / a) - 1.f);
    return s;
}
__device__ __forceinline__ float gn_act_deriv(float s, int kind, float a) {
    if (kind == 1) return s > 0.f ? 1.f : 0.f;
    if (kind == 2) return s > 0.f ? 1.f : expf(s);
    if (kind == 3) return s >= 0.f ? 1.f : expf(s / a);
    return 1.f;
}

// maps at the linearisation point: r0 = sw phi'(s)(phi(s) - y), D = sw phi'(s)^2; r0 also zero-padded to (H+1)x(W+1)
__global__ void gn_linearise_kernel(const float* __restrict__ s, const float* __restrict__ y, const float* __restrict__ sw,
                                    float* __restrict__ r0, float* __restrict__ r0_pad, float* __restrict__ D, int n, int H, int W,
                                    int act, float ap) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * (H + 1) * (W + 1)) return;
    const int Wp = W + 1, Hp = H + 1;
    const int x = i % Wp, yy = (i / Wp) % Hp, smp = i / (Wp * Hp);
    float r = 0.f;
    if (x < W && yy < H) {
        const int j = (smp * H + yy) * W + x;
        const float sv = s[j], a = gn_act(sv, act, ap), d = gn_act_deriv(sv, act, ap);
        r = sw[smp] * d * (a - y[j]);
        r0[j] = r;
        D[j] = sw[smp] * d * d;
    }
    r0_pad[i] = r;
}
// u = D (t1 + t2), dense and zero-padded
__global__ void gn_mapu_kernel(const float* __restrict__ t1, const float* __restrict__ t2, const float* __restrict__ D,
                               float* __restrict__ u, float* __restrict__ u_pad, int n, int H, int W) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * (H + 1) * (W + 1)) return;
    const int Wp = W + 1, Hp = H + 1;
    const int x = i % Wp, yy = (i / Wp) % Hp, smp = i / (Wp * Hp);
    float v = 0.f;
    if (x < W && yy < H) {
        const int j = (smp * H + yy) * W + x;
        v = D[j] * (t1[j] + t2[j]);
        u[j] = v;
    }
    u_pad[i] = v;
}
// T[i,c,y',x'] = sum_{a,b} w[c,a,b] u[i, y'-a+k/2, x'-b+k/2]   (adjoint of conv_same w.r.t. its input, k = 4)
__global__ void gn_expand_kernel(const float* __restrict__ u, const float* __restrict__ w, float* __restrict__ T, int n, int Cc,
                                 int H, int W) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * Cc * H * W) return;
    const int xq = i % W, yq = (i / W) % H, c = (i / (W * H)) % Cc, smp = i / (W * H * Cc);
    const float* um = u + (size_t)smp * H * W;
    float acc = 0.f;
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        const int ys = yq - a + 2;
        if (ys < 0 || ys >= H) continue;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const int xs = xq - b + 2;
            if (xs < 0 || xs >= W) continue;
            acc = fmaf(w[(c * 4 + a) * 4 + b], um[ys * W + xs], acc);
        }
    }
    T[i] = acc;
}
// per-sample partial of G[c,k] = sum_pix T[i,c,pix] X[i,k,pix]: CTA = (64-column block of k, sample), tile 64 x 64, K step 16
__global__ void __launch_bounds__(256) gn_txt_kernel(const float* __restrict__ T, const float* __restrict__ X, float* __restrict__ part,
                                                     int Cc, int Cin, int HW) {
    __shared__ float Ts[16][65];
    __shared__ float Xs[16][65];
    const int smp = blockIdx.z, c0 = blockIdx.y * 64, k0 = blockIdx.x * 64;
    const float* Tm = T + (size_t)smp * Cc * HW;
    const float* Xm = X + (size_t)smp * Cin * HW;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    float acc[4][4] = {};
    for (int p0 = 0; p0 < HW; p0 += 16) {
        for (int i = threadIdx.x; i < 16 * 64; i += 256) {
            const int pp = i & 15, r = i >> 4;
            Ts[pp][r] = (c0 + r < Cc && p0 + pp < HW) ? Tm[(size_t)(c0 + r) * HW + p0 + pp] : 0.f;
            Xs[pp][r] = (k0 + r < Cin && p0 + pp < HW) ? Xm[(size_t)(k0 + r) * HW + p0 + pp] : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int pp = 0; pp < 16; ++pp) {
            float a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) { a[i] = Ts[pp][ty * 4 + i]; b[i] = Xs[pp][tx * 4 + i]; }
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
            const int c = c0 + ty * 4 + i, k = k0 + tx * 4 + j;
            if (c < Cc && k < Cin) part[((size_t)smp * Cc + c) * Cin + k] = acc[i][j];
        }
}

// ---- single-CTA vector kernels over the joint variable v = [w (nw floats) ; P (nP floats)] --------------------------
struct GnVec { float *r, *rprev, *p, *x, *q; int nw, nP; float lw, lP; };

__device__ __forceinline__ float gn_block_dot(const float* a, const float* b, int n, float* red) {
    float s = 0.f;
    for (int i = threadIdx.x; i < n; i += blockDim.x) s += a[i] * b[i];
    return block_sum(s, red);
}

// r = b = -(g + reg * var); x = 0; p = 0; scalars reset  (start of a GN iteration; g_w from feat_transpose, g_P = sum of partials)
__global__ void __launch_bounds__(1024) gn_init_kernel(GnVec V, const float* __restrict__ gw, const float* __restrict__ gP_part, int n_part,
                                                       const float* __restrict__ w, const float* __restrict__ P, float* sc) {
    for (int i = threadIdx.x; i < V.nw; i += blockDim.x) { V.r[i] = -(gw[i] + V.lw * w[i]); V.x[i] = 0.f; V.p[i] = 0.f; V.rprev[i] = 0.f; }
    for (int i = threadIdx.x; i < V.nP; i += blockDim.x) {
        float g = 0.f;
        for (int s = 0; s < n_part; ++s) g += gP_part[(size_t)s * V.nP + i];
        V.r[V.nw + i] = -(g + V.lP * P[i]); V.x[V.nw + i] = 0.f; V.p[V.nw + i] = 0.f; V.rprev[V.nw + i] = 0.f;
    }
    if (threadIdx.x == 0) { sc[0] = 1.f; sc[6] = 0.f; sc[7] = 0.f; }      // rho = 1 (reset_state), done = 0, have_p = 0
}
// z = M1(r) = r / diag_M; rho = <r,z>; beta; p = z + beta p   (optimization.py:100-125)
__global__ void __launch_bounds__(1024) gn_dir_kernel(GnVec V, float* sc, int fletcher_reeves) {
    __shared__ float red[32];
    __shared__ float s_beta, s_stop;
    const int N = V.nw + V.nP;
    float l0 = 0.f, l1 = 0.f;
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
        const float z = V.r[i] / (i < V.nw ? V.lw : V.lP);
        l0 += V.r[i] * z;
        l1 += V.rprev[i] * z;
    }
    const float rho = block_sum(l0, red);
    const float rho2 = block_sum(l1, red);
    if (threadIdx.x == 0) {
        const float rho1 = sc[0];
        float beta = 0.f;
        if (sc[7] != 0.f) {
            beta = fletcher_reeves ? rho / rho1 : (rho - rho2) / rho1;
            beta = fmaxf(beta, 0.f);
        }
        s_stop = (sc[6] != 0.f || rho == 0.f) ? 1.f : 0.f;                 // check_zero(rho): keep the current iterate
        if (s_stop == 0.f) { sc[0] = rho; sc[7] = 1.f; } else sc[6] = 1.f;
        s_beta = beta;
    }
    __syncthreads();
    if (s_stop != 0.f) return;
    for (int i = threadIdx.x; i < N; i += blockDim.x) V.p[i] = V.r[i] / (i < V.nw ? V.lw : V.lP) + s_beta * V.p[i];
}
// q = [q_w + l_w p_w ; sum_i partial_i + l_P p_P]; alpha = rho / <p,q>; r_prev = r; x += alpha p; r -= alpha q (not on the last CG iteration)
__global__ void __launch_bounds__(1024) gn_step_kernel(GnVec V, const float* __restrict__ qw, const float* __restrict__ qP_part, int n_part,
                                                       float* sc, int fletcher_reeves, int last) {
    __shared__ float red[32];
    if (sc[6] != 0.f) return;
    const int N = V.nw + V.nP;
    float l = 0.f;
    for (int i = threadIdx.x; i < V.nw; i += blockDim.x) { const float q = qw[i] + V.lw * V.p[i]; V.q[i] = q; l += V.p[i] * q; }
    for (int i = threadIdx.x; i < V.nP; i += blockDim.x) {
        float g = 0.f;
        for (int s = 0; s < n_part; ++s) g += qP_part[(size_t)s * V.nP + i];
        const float q = g + V.lP * V.p[V.nw + i];
        V.q[V.nw + i] = q;
        l += V.p[V.nw + i] * q;
    }
    const float pq = block_sum(l, red);
    const float alpha = sc[0] / pq;
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
        if (!fletcher_reeves) V.rprev[i] = V.r[i];
        V.x[i] += alpha * V.p[i];
        if (!last) V.r[i] -= alpha * V.q[i];
    }
}
__global__ void gn_apply_kernel(GnVec V, float* w, float* P) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < V.nw) w[i] += V.x[i];
    else if (i < V.nw + V.nP) P[i - V.nw] += V.x[i];
}

}  // namespace b200trk
