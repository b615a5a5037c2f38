// Stage 3 (ATOM initialisation): GaussNewtonCG.run(num_cg_iter, num_gn_iter) on FactorizedConvProblem -- the joint
// optimisation of the filter w [1,Cc,k,k] and the projection matrix P [Cc,Cin,1,1] over the augmented first-frame samples.
//   reference: pytracking/libs/optimization.py:328-421 (run / run_GN_iter / A), :72-163 (run_CG),
//              pytracking/tracker/atom/optim.py:6-68 (residuals, ip_input, M1), wiring pytracking/tracker/atom/atom.py:157-178.
// The reference differentiates through operation.conv1x1 / conv2d twice per CG iteration; here J and J^T are explicit
// (SURVEY.md 9.5), projection_activation = identity (the ATOM default, atom/default.py):
//   comp_i = P X_i ;  s_i = conv_same(comp_i, w) ;  f0 = [sqrt(sw_i)(phi(s_i) - y_i), sqrt(l_w) w, sqrt(l_P) P]
//   J (dw, dP)_i = sqrt(sw_i) phi'(s_i) [conv_same(comp_i, dw) + conv_same(dP X_i, w)]
//   J^T u        = [A_comp^T(sqrt(sw) phi' u) + sqrt(l_w) u_w ;  sum_i B_w^T(sqrt(sw_i) phi'_i u_i) X_i^T + sqrt(l_P) u_P]
// where B_w^T spreads a score-sized map back onto the Cc compressed channels through the filter taps.
// This runs once per sequence, so it is a stream of small kernels with DEVICE-resident CG scalars (no host round trip
// inside the 60 CG iterations) built from the stage-2 kernels (conv1x1, conv2d_same, apply_feat_transpose).
#include "common.cuh"
#include <cstdlib>

extern "C" {
int b200trk_conv1x1(const float*, const float*, float*, int, int, int, int, int, b200trk_stream_t);
int b200trk_conv2d_same(const float*, const float*, float*, int, int, int, int, int, b200trk_stream_t);
int b200trk_apply_feat_transpose(const float*, const float*, float*, int, int, int, int, int, b200trk_stream_t);
}

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

using namespace b200trk;

extern "C" int b200trk_atom_gn_joint(float* filter, float* proj, const float* samples, const float* y, const float* sample_weight,
                                     int n, int Cin, int Cc, int H, int W, int k, int num_cg_iter, int num_gn_iter, float filter_reg,
                                     float projection_reg, int fletcher_reeves, int activation, float act_param,
                                     b200trk_stream_t stream) {
    B200_REQUIRE(filter && proj && samples && y && sample_weight, "atom_gn_joint: null pointer");
    B200_REQUIRE(n > 0 && Cin > 0 && Cc > 0 && Cc % 16 == 0, "atom_gn_joint: bad shape (n=%d, Cin=%d, Cc=%d; Cc must be a multiple of 16)", n, Cin, Cc);
    B200_REQUIRE(k == 4 && H == W && (H == 18 || H == 22), "atom_gn_joint: only a 4x4 filter on 18x18 / 22x22 features is supported");
    B200_REQUIRE(num_cg_iter >= 0 && num_gn_iter >= 0 && activation >= 0 && activation <= 3, "atom_gn_joint: bad iteration counts / activation");
    B200_REQUIRE(filter_reg > 0.f && projection_reg > 0.f, "atom_gn_joint: the M1 preconditioner divides by the regularisation weights");
    cudaStream_t st = (cudaStream_t)stream;
    if (num_cg_iter == 0 || num_gn_iter == 0) return 0;
    const int HW = H * W, HWp = (H + 1) * (W + 1), nw = Cc * 16, nP = Cc * Cin, N = nw + nP;
    // scratch carving (slot 5)
    size_t fl = 0;
    auto take = [&](size_t cnt) { const size_t o = fl; fl += (cnt + 63) / 64 * 64; return o; };
    const size_t o_comp = take((size_t)n * Cc * HW), o_compp = take((size_t)n * Cc * HW), o_T = take((size_t)n * Cc * HW);
    const size_t o_s = take((size_t)n * HW), o_t1 = take((size_t)n * HW), o_t2 = take((size_t)n * HW), o_D = take((size_t)n * HW), o_u = take((size_t)n * HW);
    const size_t o_upad = take((size_t)n * HWp), o_gw = take(nw), o_part = take((size_t)n * nP);
    const size_t o_r = take(N), o_rp = take(N), o_p = take(N), o_x = take(N), o_q = take(N), o_sc = take(64);
    float* ws = (float*)workspace(fl * sizeof(float), 5);
    if (!ws) return 3;
    float *comp = ws + o_comp, *compp = ws + o_compp, *T = ws + o_T, *s = ws + o_s, *t1 = ws + o_t1, *t2 = ws + o_t2, *D = ws + o_D;
    float *u = ws + o_u, *upad = ws + o_upad, *gw = ws + o_gw, *part = ws + o_part, *sc = ws + o_sc;
    GnVec V{ws + o_r, ws + o_rp, ws + o_p, ws + o_x, ws + o_q, nw, nP, filter_reg, projection_reg};
    const int mapT = (n * HWp + 255) / 256, expT = (n * Cc * HW + 255) / 256;
    const dim3 txt_grid((Cin + 63) / 64, (Cc + 63) / 64, n);
    auto launched = [&]() { g_launch_count.fetch_add(1, std::memory_order_relaxed); return cudaGetLastError() == cudaSuccess; };
    for (int gn = 0; gn < num_gn_iter; ++gn) {
        // ---- linearise at (w, P) ----
        if (int e = b200trk_conv1x1(samples, proj, comp, n, Cin, Cc, H, W, stream)) return e;
        if (int e = b200trk_conv2d_same(comp, filter, s, n, Cc, H, W, k, stream)) return e;
        gn_linearise_kernel<<<mapT, 256, 0, st>>>(s, y, sample_weight, u, upad, D, n, H, W, activation, act_param);
        if (!launched()) { set_error("atom_gn_joint: linearise launch failed"); return 1; }
        if (int e = b200trk_apply_feat_transpose(comp, upad, gw, n, Cc, H, W, k, stream)) return e;
        gn_expand_kernel<<<expT, 256, 0, st>>>(u, filter, T, n, Cc, H, W);
        if (!launched()) { set_error("atom_gn_joint: expand launch failed"); return 1; }
        gn_txt_kernel<<<txt_grid, 256, 0, st>>>(T, samples, part, Cc, Cin, HW);
        if (!launched()) { set_error("atom_gn_joint: gemm launch failed"); return 1; }
        gn_init_kernel<<<1, 1024, 0, st>>>(V, gw, part, n, filter, proj, sc);
        if (!launched()) { set_error("atom_gn_joint: init launch failed"); return 1; }
        // ---- CG on J^T J dx = -J^T f0 ----
        for (int ii = 0; ii < num_cg_iter; ++ii) {
            gn_dir_kernel<<<1, 1024, 0, st>>>(V, sc, fletcher_reeves);
            if (!launched()) { set_error("atom_gn_joint: direction launch failed"); return 1; }
            if (int e = b200trk_conv1x1(samples, V.p + nw, compp, n, Cin, Cc, H, W, stream)) return e;        // dP X
            if (int e = b200trk_conv2d_same(comp, V.p, t1, n, Cc, H, W, k, stream)) return e;                 // conv_same(P X, dw)
            if (int e = b200trk_conv2d_same(compp, filter, t2, n, Cc, H, W, k, stream)) return e;             // conv_same(dP X, w)
            gn_mapu_kernel<<<mapT, 256, 0, st>>>(t1, t2, D, u, upad, n, H, W);
            if (!launched()) { set_error("atom_gn_joint: map launch failed"); return 1; }
            if (int e = b200trk_apply_feat_transpose(comp, upad, gw, n, Cc, H, W, k, stream)) return e;
            gn_expand_kernel<<<expT, 256, 0, st>>>(u, filter, T, n, Cc, H, W);
            if (!launched()) { set_error("atom_gn_joint: expand launch failed"); return 1; }
            gn_txt_kernel<<<txt_grid, 256, 0, st>>>(T, samples, part, Cc, Cin, HW);
            if (!launched()) { set_error("atom_gn_joint: gemm launch failed"); return 1; }
            gn_step_kernel<<<1, 1024, 0, st>>>(V, gw, part, n, sc, fletcher_reeves, ii == num_cg_iter - 1 ? 1 : 0);
            if (!launched()) { set_error("atom_gn_joint: step launch failed"); return 1; }
        }
        gn_apply_kernel<<<(N + 255) / 256, 256, 0, st>>>(V, filter, proj);
        if (!launched()) { set_error("atom_gn_joint: apply launch failed"); return 1; }
    }
    return 0;
}
