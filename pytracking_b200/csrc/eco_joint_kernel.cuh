// ECO first-frame joint optimisation: GaussNewtonCG.run(num_cg_iter, num_gn_iter) on FactorizedConvProblem for ONE
// feature block -- the filter hf [C,H,Wh] (complex) and the projection matrix P [Cin,C] (real) -- as ONE persistent cooperative kernel.
//   reference: pytracking/tracker/eco/optim.py:8-117 (residuals, ip_input, M1), pytracking/libs/optimization.py:328-421 (run,
//   run_GN_iter, A by double backward) + :72-163 (run_CG; Fletcher-Reeves, state reset every GN iteration), wiring eco.py:155-162.
//
// The reference differentiates the residual function twice per CG iteration; here J and J^T are explicit (oracle/eco_oracle.py
// joint_J / joint_JT, pinned to the reference's autograd results).  With c0 = X P0 (compressed samples at the linearisation point):
//   f        = [ sqrt(sw_n) (sum_c c0[n,c] h0[c] - yf),  W (*) ext(h0),  sqrt(lambda) P0 ]
//   J (dh,dP)= [ sqrt(sw_n) (sum_c c0[n,c] dh[c] + sum_i X[n,i] v[i]),  W (*) [0 | dh],  sqrt(lambda) dP ],   v[i] = sum_c dP[i,c] h0[c]
//   J^T u    : gh[c] = sum_n sqrt(sw_n) conj(c0[n,c]) u[n] + (W^T-correlation of u_reg, extension columns dropped)
//              gP[i,c] = sum_coefficients Re(conj(h0[c]) w[i]) + sqrt(lambda) u_P,   w[i] = sum_n sqrt(sw_n) conj(X[n,i]) u[n]
// The Hermitian extension of the regularisation residual is DETACHED in the reference (optim.py:56): it enters f but not J.
// J_reg^T J_reg is the correlation with the autocorrelation of W (all partial overlaps exist because the reference pads fully).
//
// Decomposition as in eco_cg_kernel.cuh: Fourier coefficients dealt to the CTAs in contiguous ranges, the [N, Cin] sample slab of a
// coefficient resident in shared memory (row pitch Cin + 1, bank-conflict free for both access directions); in the coefficient-local
// phases a coefficient is worked on by a team: one warp when the CTA owns many coefficients, the whole CTA when it owns few.  The projection-matrix gradient is a sum over all coefficients: every coefficient publishes its
// Cin-vector w, then the [Cin, C] elements are dealt to the CTAs in tiles (element x coefficient-split threads, fixed summation
// order).  Four grid barriers per CG iteration; every reduction deterministic.
// Plain SIMT CUDA C: the same source runs on the CPU under tests/cpu_emul/cuda_shim.h (tests/test_eco_cpu.py).
#pragma once

namespace b200trk {

struct EcoJointParams {
    float* hf;                 // [C,H,Wh,2] in/out
    float* proj;               // [Cin,C] in/out
    const float* samples;      // [H,Wh,N,Cin,2]
    const float* yf;           // [H,Wh]
    const float* sw_sqrt;      // [N]
    const float* reg_filter;   // [rh,rw]
    const float* dMh_in;       // [C,H,Wh] filter part of the diagonal preconditioner (FactorizedConvProblem.diag_M)
    float dMP, lambda;         // projection part of the preconditioner; projection_reg
    int H, Wh, N, Cin, C, rh, rw, num_cg, num_gn;
    // workspace
    float2 *h0w, *phw, *xhw, *rhw, *qhw;   // pixel-major [H*Wh][C]
    float* dMh;                            // pixel-major [H*Wh][C]
    float2* c0w;                           // [H*Wh][N][C]   compressed samples at the linearisation point
    float2* wv;                            // [H*Wh][Cin]    w vectors of the projection gradient
    float *pP, *xP, *rP, *qP;              // [Cin*C]
    float* dots;
    unsigned* barrier;
    int res_slabs, npx_max, EPB, SPL;      // resident slabs per CTA; elements per tile, coefficient splits per element
    int stage_pm;                          // the projection matrix / its CG direction staged in shared memory per phase
    int wide;                              // the whole CTA works on one coefficient at a time (few coefficients per CTA)
};

struct EcoJointPlan {
    int grid, block, res_slabs, npx_max, EPB, SPL, stage_pm, wide;
    size_t smem_bytes, ws_bytes, off_fields, off_dMh, off_c0, off_wv, off_P, off_dots;
};

constexpr int ECOJ_MAX_TAPS = 15 * 15;

// shared memory (floats): acorr[228], tapw[228], tapi[228], scal[8], red32[32], sw[N4], per warp {h0[2C], ph[2C], v[2Cin], u[2N4]},
// tile partials [block], staged matrix [Cin][C+1] (when it fits), slabs [res][N][(Cin+1)*2]
inline size_t ecoj_fixed_smem_floats(int N, int Cin, int C, int block) {
    const int N4 = (N + 3) & ~3;
    return 3 * (size_t)(ECOJ_MAX_TAPS + 3) + 8 + 32 + N4 + (size_t)(block / 32) * (4 * C + 2 * Cin + 2 * N4) + block;
}
inline size_t ecoj_stage_floats(int Cin, int C) { return ((size_t)Cin * (C + 1) + 1) & ~(size_t)1; }

inline EcoJointPlan eco_joint_plan(int H, int Wh, int N, int Cin, int C, int num_cg, int num_gn, int max_ctas, int block) {
    EcoJointPlan pl{};
    const int P = H * Wh;
    pl.block = block;
    pl.grid = P < max_ctas ? P : max_ctas;
    pl.npx_max = (P + pl.grid - 1) / pl.grid;
    pl.wide = pl.npx_max * 2 <= block / 32 ? 1 : 0;
    size_t fixed = ecoj_fixed_smem_floats(N, Cin, C, block) * sizeof(float);
    const size_t slab = (size_t)N * (Cin + 1) * 2 * sizeof(float);
    const size_t limit = 227 * 1024 - 1024;
    // the [Cin, C] matrix every coefficient multiplies with (P when linearising, the direction dP in the CG iterations) first ...
    pl.stage_pm = fixed + ecoj_stage_floats(Cin, C) * sizeof(float) <= limit ? 1 : 0;
    if (pl.stage_pm) fixed += ecoj_stage_floats(Cin, C) * sizeof(float);
    // ... then as many sample slabs as still fit
    size_t res = fixed < limit ? (limit - fixed) / slab : 0;
    if (res > (size_t)pl.npx_max) res = (size_t)pl.npx_max;
    pl.res_slabs = (int)res;
    pl.smem_bytes = fixed + res * slab;
    // projection-gradient tiles: EPB consecutive (i,c) elements x SPL splits of the coefficient range per CTA tile
    const int nelem = Cin * C;
    int epb = (nelem + pl.grid - 1) / pl.grid;
    epb = (epb + 15) & ~15;
    if (epb > block) epb = block;
    pl.EPB = epb;
    pl.SPL = block / epb;
    const size_t field = (size_t)P * C * 2 * sizeof(float);
    size_t off = 256;
    pl.off_fields = off; off += 5 * field;
    pl.off_dMh = off; off += field / 2;
    pl.off_c0 = off; off += (size_t)P * N * C * 2 * sizeof(float);
    pl.off_wv = off; off += (size_t)P * Cin * 2 * sizeof(float);
    pl.off_P = off; off += 4 * (size_t)nelem * sizeof(float);
    off = (off + 255) & ~(size_t)255;
    pl.off_dots = off; off += (size_t)((2 * num_cg + 2) * num_gn + 4) * pl.grid * 2 * sizeof(float);
    pl.ws_bytes = off;
    return pl;
}

#ifndef B200_DYN_SMEM
#ifdef B200_CPU_EMUL
#define B200_DYN_SMEM(name) unsigned char* name = ::cpu_emul::dyn_smem()
#else
#define B200_DYN_SMEM(name) extern __shared__ __align__(16) unsigned char name[]
#endif
#endif

__global__ void __launch_bounds__(256, 1) eco_joint_kernel(EcoJointParams P) {
    B200_DYN_SMEM(smem_raw);
    const int tid = threadIdx.x, NT = blockDim.x, nb = gridDim.x, cta = blockIdx.x;
    const int lane = tid & 31, warp = tid >> 5, NW = NT >> 5;
    const int H = P.H, Wh = P.Wh, N = P.N, Cin = P.Cin, C = P.C;
    const int NPIX = H * Wh, NE = Cin * C, N4 = (N + 3) & ~3;
    const int TW = 2 * P.rw - 1, NTAP = (2 * P.rh - 1) * TW;
    const int PITCH = Cin + 1;                               // float2 per slab row in shared memory

    float* s_ac = reinterpret_cast<float*>(smem_raw);        // autocorrelation of the regularisation filter
    float* s_tapw = s_ac + ECOJ_MAX_TAPS + 3;                // its non-zero taps, compacted: weight ...
    int* s_tapi = reinterpret_cast<int*>(s_tapw + ECOJ_MAX_TAPS + 3);   // ... and (row << 8 | column)
    float* s_scal = reinterpret_cast<float*>(s_tapi + ECOJ_MAX_TAPS + 3);
    float* s_red32 = s_scal + 8;
    float* s_sw = s_red32 + 32;
    float* s_warp = s_sw + N4;                               // per-warp scratch
    const int WARP_FLOATS = 4 * C + 2 * Cin + 2 * N4;
    float2* s_h0 = reinterpret_cast<float2*>(s_warp + (size_t)warp * WARP_FLOATS);   // {h0[C], ph[C], v[Cin], u[N4]} of this warp
    float* s_tile = s_warp + (size_t)NW * WARP_FLOATS;       // [NT] partial sums of the projection-gradient tiles
    float* s_pm = s_tile + NT;                               // [Cin][C+1] staged matrix (P.stage_pm)
    const int PMP = C + 1;
    float2* s_slab = reinterpret_cast<float2*>(s_pm + (P.stage_pm ? (((size_t)Cin * PMP + 1) & ~(size_t)1) : 0));

    const int p0 = (int)(((long long)cta * NPIX) / nb), p1 = (int)(((long long)(cta + 1) * NPIX) / nb);
    const int npx = p1 - p0;
    const int nel = npx * C;
    const size_t e0 = (size_t)p0 * C;
    unsigned epoch = 0, dead = 0;
    int dot_slot = 0;

    // ---- prologue: autocorrelation of W, sqrt sample weights, resident slabs (padded rows), state import ---------------------------
    for (int t = tid; t < NTAP; t += NT) {
        const int s = t / TW - (P.rh - 1), u = t % TW - (P.rw - 1);
        float acc = 0.f;
        for (int a = 0; a < P.rh; ++a) {
            const int a2 = a + s;
            if (a2 < 0 || a2 >= P.rh) continue;
            for (int b = 0; b < P.rw; ++b) {
                const int b2 = b + u;
                if (b2 < 0 || b2 >= P.rw) continue;
                acc += P.reg_filter[a * P.rw + b] * P.reg_filter[a2 * P.rw + b2];
            }
        }
        s_ac[t] = acc;
    }
    for (int n = tid; n < N; n += NT) s_sw[n] = P.sw_sqrt[n];
    __syncthreads();
    if (tid == 0) {                                          // the sparsified reg filters leave many taps exactly zero
        int nz = 0;
        for (int t = 0; t < NTAP; ++t)
            if (s_ac[t] != 0.f) { s_tapw[nz] = s_ac[t]; s_tapi[nz] = ((t / TW) << 8) | (t % TW); ++nz; }
        s_tapi[ECOJ_MAX_TAPS + 2] = nz;
    }
    {
        const int nres = npx < P.res_slabs ? npx : P.res_slabs;
        const float2* src = reinterpret_cast<const float2*>(P.samples) + (size_t)p0 * N * Cin;
        const int rows = nres * N;
        for (int e = tid; e < rows * Cin; e += NT) {
            const int row = e / Cin, i = e - row * Cin;
            s_slab[(size_t)row * PITCH + i] = src[e];
        }
    }
    for (int e = tid; e < nel; e += NT) {
        const int pix = p0 + e / C, c = e % C;
        const size_t ref = (size_t)c * NPIX + pix;
        P.h0w[e0 + e] = reinterpret_cast<const float2*>(P.hf)[ref];
        P.dMh[e0 + e] = P.dMh_in[ref];
    }
    eco_grid_barrier(P.barrier, epoch, dead);                // h0 of every CTA in place (shared memory is published by its bar.sync)

    auto slab_of = [&](int j, int& pitch) -> const float2* {
        if (j < P.res_slabs) { pitch = PITCH; return s_slab + (size_t)j * N * PITCH; }
        pitch = Cin;
        return reinterpret_cast<const float2*>(P.samples) + (size_t)(p0 + j) * N * Cin;
    };

    // regularisation part of J^T (.) at coefficient `pix`, channel c: correlation of `field` with the autocorrelation of W; negative kx
    // by Hermitian symmetry when `ext` (the residual value), zero otherwise (the Jacobian)
    auto reg_at = [&](const float2* field, int pix, int c, bool ext) -> float2 {
        const int y = pix / Wh, x = pix - y * Wh;
        const int NZ = s_tapi[ECOJ_MAX_TAPS + 2];
        float2 acc = make_float2(0.f, 0.f);
        for (int t0 = 0; t0 < NZ; t0 += 8) {                 // eight independent L2 requests in flight
            float2 v[8];
            float wk[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const int t = t0 + q;
                bool ok = t < NZ;
                const int code = ok ? s_tapi[t] : 0;
                int yy = y + (code >> 8) - (P.rh - 1);
                int kx = x + (code & 255) - (P.rw - 1);
                ok = ok && yy >= 0 && yy < H && kx < Wh && (ext || kx >= 0);
                const bool cj = kx < 0;
                if (cj) { yy = H - 1 - yy; kx = -kx; }
                if (!ok) { yy = y; kx = x; }
                wk[q] = ok ? s_tapw[t] : 0.f;
                v[q] = ok ? __ldcg(field + ((size_t)yy * Wh + kx) * C + c) : make_float2(0.f, 0.f);
                if (cj) v[q].y = -v[q].y;
            }
#pragma unroll
            for (int q = 0; q < 8; ++q) { acc.x += wk[q] * v[q].x; acc.y += wk[q] * v[q].y; }
        }
        return acc;
    };

    // coefficient-local part of J^T J (dh, dP) (lin == false) or of J^T f (lin == true, which also recomputes c0 = X P) for the
    // coefficients of this CTA, one warp per coefficient: writes gh -> dst and the w vector -> wv
    auto pixel_phase = [&](bool lin, float2* dst) {
        const float* pm = lin ? P.proj : P.pP;               // the matrix of this phase, L2 resident ...
        int pmp = C;
        if (P.stage_pm) {                                    // ... staged once per phase for all coefficients of the CTA
            for (int e = tid; e < NE; e += NT) s_pm[(size_t)(e / C) * PMP + e % C] = __ldcg(pm + e);
            __syncthreads();
            pm = s_pm;
            pmp = PMP;
        }
        // a coefficient is worked on by a TEAM: one warp when the CTA owns many coefficients, the whole CTA when it owns few (deep block:
        // one); the scratch vectors of a team live in shared memory, its loops are strided by the team size
        const bool wide = P.wide != 0;
        const int T = wide ? NT : 32, tl = wide ? tid : lane;
        float2* t_h0 = wide ? reinterpret_cast<float2*>(s_warp) : s_h0;
        float2* t_ph = t_h0 + C;
        float2* t_v = t_ph + C;
        float2* t_u = t_v + Cin;
        auto team_sync = [&]() { if (wide) __syncthreads(); else __syncwarp(); };
        const int NB8 = (N + 7) / 8;
        for (int j = wide ? 0 : warp; j < npx; j += wide ? 1 : NW) {
            const int pix = p0 + j;
            int pitch;
            const float2* S = slab_of(j, pitch);
            float2* c0 = P.c0w + (size_t)pix * N * C;
            for (int c = tl; c < C; c += T) {
                t_h0[c] = P.h0w[(size_t)pix * C + c];
                t_ph[c] = lin ? make_float2(0.f, 0.f) : P.phw[(size_t)pix * C + c];
            }
            team_sync();
            if (lin) {
                // c0[n,c] = sum_i X[n,i] P[i,c], eight sample rows per pass over the projection matrix
                for (int item = tl; item < NB8 * C; item += T) {
                    const int n0 = (item / C) * 8, c = item % C;
                    float2 acc[8];
#pragma unroll
                    for (int k = 0; k < 8; ++k) acc[k] = make_float2(0.f, 0.f);
                    for (int i = 0; i < Cin; ++i) {
                        const float pv = P.stage_pm ? pm[(size_t)i * pmp + c] : __ldcg(pm + (size_t)i * pmp + c);
#pragma unroll
                        for (int k = 0; k < 8; ++k) {
                            const int n = n0 + k < N ? n0 + k : N - 1;          // the tail repeats the last row (not stored)
                            const float2 xv = S[(size_t)n * pitch + i];
                            acc[k].x += xv.x * pv; acc[k].y += xv.y * pv;
                        }
                    }
#pragma unroll
                    for (int k = 0; k < 8; ++k)
                        if (n0 + k < N) c0[(size_t)(n0 + k) * C + c] = acc[k];
                }
            } else {
                // v[i] = sum_c dP[i,c] h0[c]
                for (int i = tl; i < Cin; i += T) {
                    float ar = 0.f, ai = 0.f;
                    for (int c = 0; c < C; ++c) {
                        const float pv = P.stage_pm ? pm[(size_t)i * pmp + c] : __ldcg(pm + (size_t)i * pmp + c);
                        ar += pv * t_h0[c].x; ai += pv * t_h0[c].y;
                    }
                    t_v[i] = make_float2(ar, ai);
                }
            }
            team_sync();
            // u[n] = sw_n (sum_c c0[n,c] h[c] (+ sum_i X[n,i] v[i]) (- yf))
            const float yfv = P.yf[pix];
            const float2* hv = lin ? t_h0 : t_ph;
            if (wide) {                                      // a warp per sample row, the two dot products reduced across its lanes
                for (int n = warp; n < N; n += NW) {
                    float ar = 0.f, ai = 0.f;
                    for (int c = lane; c < C; c += 32) {
                        const float2 cv = c0[(size_t)n * C + c];
                        ar += cv.x * hv[c].x - cv.y * hv[c].y;
                        ai += cv.x * hv[c].y + cv.y * hv[c].x;
                    }
                    if (!lin)
                        for (int i = lane; i < Cin; i += 32) {
                            const float2 xv = S[(size_t)n * pitch + i];
                            ar += xv.x * t_v[i].x - xv.y * t_v[i].y;
                            ai += xv.x * t_v[i].y + xv.y * t_v[i].x;
                        }
                    ar = warp_sum(ar);
                    ai = warp_sum(ai);
                    if (lane == 0) {
                        if (lin) ar -= yfv;
                        const float w2 = s_sw[n] * s_sw[n];
                        t_u[n] = make_float2(w2 * ar, w2 * ai);
                    }
                }
            } else {                                         // a sample row per lane
                for (int n = lane; n < N; n += 32) {
                    float ar = 0.f, ai = 0.f;
                    for (int c = 0; c < C; ++c) {
                        const float2 cv = c0[(size_t)n * C + c];
                        ar += cv.x * hv[c].x - cv.y * hv[c].y;
                        ai += cv.x * hv[c].y + cv.y * hv[c].x;
                    }
                    if (lin) {
                        ar -= yfv;
                    } else {
                        for (int i = 0; i < Cin; ++i) {
                            const float2 xv = S[(size_t)n * pitch + i];
                            ar += xv.x * t_v[i].x - xv.y * t_v[i].y;
                            ai += xv.x * t_v[i].y + xv.y * t_v[i].x;
                        }
                    }
                    const float w2 = s_sw[n] * s_sw[n];      // sqrt(sw) from the residual / J, sqrt(sw) again in J^T
                    t_u[n] = make_float2(w2 * ar, w2 * ai);
                }
            }
            team_sync();
            // gh[c] = sum_n conj(c0[n,c]) u[n] + regularisation
            for (int c = tl; c < C; c += T) {
                float ar = 0.f, ai = 0.f;
                for (int n = 0; n < N; ++n) {
                    const float2 cv = c0[(size_t)n * C + c];
                    ar += cv.x * t_u[n].x + cv.y * t_u[n].y;
                    ai += cv.x * t_u[n].y - cv.y * t_u[n].x;
                }
                const float2 rg = reg_at(lin ? P.h0w : P.phw, pix, c, lin);
                dst[(size_t)pix * C + c] = make_float2(ar + rg.x, ai + rg.y);
            }
            // w[i] = sum_n conj(X[n,i]) u[n]
            for (int i = tl; i < Cin; i += T) {
                float ar = 0.f, ai = 0.f;
                for (int n = 0; n < N; ++n) {
                    const float2 xv = S[(size_t)n * pitch + i];
                    ar += xv.x * t_u[n].x + xv.y * t_u[n].y;
                    ai += xv.x * t_u[n].y - xv.y * t_u[n].x;
                }
                P.wv[(size_t)pix * Cin + i] = make_float2(ar, ai);
            }
            team_sync();
        }
        __syncthreads();
    };

    // projection part of J^T: out[e] = sum_coefficients Re(conj(h0[c]) w[i]) + lambda * addv[e] for the tiles of this CTA, then `fn(e, value)`
    // on the split-0 thread of the element.  Tiles: EPB consecutive elements x SPL splits of the coefficient range.
    auto proj_phase = [&](const float* addv, auto fn) {
        const int EPB = P.EPB, SPL = P.SPL;
        const int el = tid % EPB, sp = tid / EPB;
        const int ntiles = (NE + EPB - 1) / EPB;
        for (int tile = cta; tile < ntiles; tile += nb) {
            const int e = tile * EPB + el;
            const bool act = sp < SPL && e < NE;
            float acc = 0.f;
            if (act) {
                const int i = e / C, c = e - i * C;
                const int q0 = (int)(((long long)sp * NPIX) / SPL), q1 = (int)(((long long)(sp + 1) * NPIX) / SPL);
                for (int q = q0; q < q1; q += 8) {           // sixteen independent L2 requests in flight, summed in coefficient order
                    float2 hv[8], wv[8];
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const bool ok = q + k < q1;
                        hv[k] = ok ? __ldcg(P.h0w + (size_t)(q + k) * C + c) : make_float2(0.f, 0.f);
                        wv[k] = ok ? __ldcg(P.wv + (size_t)(q + k) * Cin + i) : make_float2(0.f, 0.f);
                    }
#pragma unroll
                    for (int k = 0; k < 8; ++k) acc += hv[k].x * wv[k].x + hv[k].y * wv[k].y;
                }
            }
            s_tile[tid] = acc;
            __syncthreads();
            if (act && sp == 0) {
                for (int s2 = 1; s2 < SPL; ++s2) acc += s_tile[s2 * EPB + el];
                fn(e, acc + P.lambda * addv[e]);
            }
            __syncthreads();
        }
    };

    auto exchange = [&](float& d0) {
        d0 = block_sum(d0, s_red32);
        float* slotp = P.dots + (size_t)dot_slot * nb * 2;
        dot_slot += 1;
        if (tid == 0) slotp[cta * 2] = d0;
        eco_grid_barrier(P.barrier, epoch, dead);
        if (tid < 32) {
            float a = 0.f;
            for (int i = tid; i < nb; i += 32) a += __ldcg(slotp + 2 * i);
            a = warp_sum(a);
            if (tid == 0) s_scal[0] = a;
        }
        __syncthreads();
        d0 = s_scal[0];
        __syncthreads();
    };

    for (int gn = 0; gn < P.num_gn; ++gn) {
        // ---- linearise: c0 = X P, b = -J^T f (optimization.py:383-401) ---------------------------------------------------------------
        pixel_phase(true, P.rhw);
        eco_grid_barrier(P.barrier, epoch, dead);            // w of every coefficient in place
        float l0 = 0.f;
        proj_phase(P.proj, [&](int e, float g) {
            const float r = -g;
            P.rP[e] = r;
            l0 += r * (r / P.dMP);
        });
        for (int e = tid; e < nel; e += NT) {
            const float kw = ((p0 + e / C) % Wh == 0) ? 1.f : 2.f;
            float2 r = P.rhw[e0 + e];
            r.x = -r.x; r.y = -r.y;
            P.rhw[e0 + e] = r;
            const float dm = P.dMh[e0 + e];
            l0 += kw * (r.x * (r.x / dm) + r.y * (r.y / dm));
        }
        exchange(l0);
        float rho = l0, rho1 = 1.f;
        bool have_p = false, have_x = false;

        for (int ii = 0; ii < P.num_cg; ++ii) {
            if (rho == 0.f) break;                           // check_zero: uniform across the grid
            // ---- p = z + beta p (Fletcher-Reeves, optimization.py:111-121) -------------------------------------------------------------
            const float beta = have_p ? fmaxf(rho / rho1, 0.f) : 0.f;
            for (int e = tid; e < nel; e += NT) {
                const float2 r = P.rhw[e0 + e];
                const float dm = P.dMh[e0 + e];
                float2 p = have_p ? P.phw[e0 + e] : make_float2(0.f, 0.f);
                p.x = r.x / dm + p.x * beta; p.y = r.y / dm + p.y * beta;
                P.phw[e0 + e] = p;
            }
            {
                const int EPB = P.EPB, ntiles = (NE + EPB - 1) / EPB;
                for (int tile = cta; tile < ntiles; tile += nb) {
                    const int e = tile * EPB + tid;
                    if (tid < EPB && e < NE) P.pP[e] = P.rP[e] / P.dMP + (have_p ? P.pP[e] * beta : 0.f);
                }
            }
            have_p = true;
            eco_grid_barrier(P.barrier, epoch, dead);        // p of every CTA in place
            // ---- q = J^T J p ---------------------------------------------------------------------------------------------------------------
            pixel_phase(false, P.qhw);
            eco_grid_barrier(P.barrier, epoch, dead);        // w of every coefficient in place
            l0 = 0.f;
            proj_phase(P.pP, [&](int e, float g) {
                P.qP[e] = g;
                l0 += P.pP[e] * g;
            });
            for (int e = tid; e < nel; e += NT) {
                const float kw = ((p0 + e / C) % Wh == 0) ? 1.f : 2.f;
                const float2 p = P.phw[e0 + e], q = P.qhw[e0 + e];
                l0 += kw * (p.x * q.x + p.y * q.y);
            }
            exchange(l0);
            const float alpha = rho / l0;
            // ---- x += alpha p; r -= alpha q; next <r,z> (optimization.py:127-146) ---------------------------------------------------------------
            const bool more = ii + 1 < P.num_cg;
            l0 = 0.f;
            for (int e = tid; e < nel; e += NT) {
                const float2 p = P.phw[e0 + e];
                float2 x = have_x ? P.xhw[e0 + e] : make_float2(0.f, 0.f);
                x.x += p.x * alpha; x.y += p.y * alpha;
                P.xhw[e0 + e] = x;
                if (more) {
                    float2 r = P.rhw[e0 + e];
                    const float2 q = P.qhw[e0 + e];
                    r.x -= q.x * alpha; r.y -= q.y * alpha;
                    P.rhw[e0 + e] = r;
                    const float kw = ((p0 + e / C) % Wh == 0) ? 1.f : 2.f;
                    const float dm = P.dMh[e0 + e];
                    l0 += kw * (r.x * (r.x / dm) + r.y * (r.y / dm));
                }
            }
            {
                const int EPB = P.EPB, ntiles = (NE + EPB - 1) / EPB;
                for (int tile = cta; tile < ntiles; tile += nb) {
                    const int e = tile * EPB + tid;
                    if (tid < EPB && e < NE) {
                        P.xP[e] = (have_x ? P.xP[e] : 0.f) + P.pP[e] * alpha;
                        if (more) {
                            const float r = P.rP[e] - P.qP[e] * alpha;
                            P.rP[e] = r;
                            l0 += r * (r / P.dMP);
                        }
                    }
                }
            }
            have_x = true;
            if (more) {
                rho1 = rho;
                exchange(l0);
                rho = l0;
            }
        }
        // ---- x += delta (optimization.py:403-404) ----------------------------------------------------------------------------------------
        __syncthreads();
        if (have_x) {
            for (int e = tid; e < nel; e += NT) {
                float2 h = P.h0w[e0 + e];
                const float2 x = P.xhw[e0 + e];
                h.x += x.x; h.y += x.y;
                P.h0w[e0 + e] = h;
            }
            const int EPB = P.EPB, ntiles = (NE + EPB - 1) / EPB;
            for (int tile = cta; tile < ntiles; tile += nb) {
                const int e = tile * EPB + tid;
                if (tid < EPB && e < NE) P.proj[e] += P.xP[e];
            }
        }
        eco_grid_barrier(P.barrier, epoch, dead);            // the new linearisation point of every CTA in place
    }

    for (int e = tid; e < nel; e += NT) {
        const int pix = p0 + e / C, c = e % C;
        reinterpret_cast<float2*>(P.hf)[(size_t)c * NPIX + pix] = P.h0w[e0 + e];
    }
}

}  // namespace b200trk
