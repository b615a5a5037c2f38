// ToMP model predictor core: Transformer.forward of ltr/models/transformer/transformer.py:90-96
// = 6 post-norm encoder layers (TransformerEncoderLayer.forward_post :173-181) + 6 post-norm decoder layers
// (TransformerDecoderLayer.forward_post :224-238) + final decoder LayerNorm, nn.MultiheadAttention semantics
// (packed in_proj, scaling 1/sqrt(head_dim), key_padding_mask -> -inf), dropout = identity (eval).
//
// All token-wise linear layers (QK / V / output projections, FFN 256->2048->256, decoder K/V of the memory) run on the
// tcgen05 3xTF32 implicit-GEMM kernel of conv_tc.cu as 1x1 convolutions over the [L*B, D] token matrix (bias, ReLU and the
// residual add fused in its epilogue); attention is an fp32 online-softmax kernel (4 lanes per query, K/V tiles staged in
// shared memory); LayerNorm is one warp per token; the decoder's 1-token linear layers are warp-per-output GEMVs.
#include "net.cuh"
#include <cmath>
#include <cstring>

namespace b200trk {

int tc_conv_prepare(b200trk_net* net, Op& op, const std::vector<float>& w_khwc);   // conv_tc.cu
int tc_conv_launch(b200trk_net* net, const Op& op, int S, cudaStream_t st);
void tc_conv_free(TcConv* tc);

// ---------------------------------------------------------------------------------------------------------------
// elementwise / small kernels
// ---------------------------------------------------------------------------------------------------------------
// out[l,b,:] = x[l,b,:] + pos[l, b % Bp, :]
__global__ void add_pos_kernel(const float4* __restrict__ x, const float4* __restrict__ pos, float4* __restrict__ out,
                               int L, int B, int Bp, int D4) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= L * B * D4) return;
    const int d = i % D4, t = i / D4, b = t % B, l = t / B;
    const float4 a = x[i], p = pos[((size_t)l * Bp + (Bp == 1 ? 0 : b)) * D4 + d];
    out[i] = make_float4(a.x + p.x, a.y + p.y, a.z + p.z, a.w + p.w);
}

// y[t,:] = LayerNorm(x[t,:]) * gamma + beta, eps = 1e-5, D <= 1024 (one warp per token)
__global__ void layernorm_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                                 float* __restrict__ y, int T, int D) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= T) return;
    const float* xr = x + (size_t)warp * D;
    float s = 0.f;
    for (int d = lane; d < D; d += 32) s += xr[d];
    s = warp_sum(s);
    const float mean = s / (float)D;
    float v = 0.f;
    for (int d = lane; d < D; d += 32) { const float c = xr[d] - mean; v += c * c; }
    v = warp_sum(v);
    const float rstd = rsqrtf(v / (float)D + 1e-5f);
    for (int d = lane; d < D; d += 32) y[(size_t)warp * D + d] = (xr[d] - mean) * rstd * gamma[d] + beta[d];
}

// y[m,n] = act(sum_k x[m,k] W[n,k] + b[n]) (+ res[m,n]) for a handful of rows m (decoder tokens): one warp per (m, n)
__global__ void small_linear_kernel(const float* __restrict__ x, const float* __restrict__ W, const float* __restrict__ b,
                                    const float* __restrict__ res, float* __restrict__ y, int M, int K, int N, int relu) {
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (w >= M * N) return;
    const int m = w / N, n = w - m * N;
    const float4* xr = reinterpret_cast<const float4*>(x + (size_t)m * K);
    const float4* wr = reinterpret_cast<const float4*>(W + (size_t)n * K);
    float acc = 0.f;
    for (int k = lane; k < K / 4; k += 32) {
        const float4 a = xr[k], c = __ldg(wr + k);
        acc += a.x * c.x + a.y * c.y + a.z * c.z + a.w * c.w;
    }
    acc = warp_sum(acc);
    if (lane == 0) {
        float v = acc + (b ? b[n] : 0.f);
        if (relu) v = fmaxf(v, 0.f);
        if (res) v += res[(size_t)m * N + n];
        y[(size_t)m * N + n] = v;
    }
}

// nn.MultiheadAttention core: O[lq,b,h,:] = softmax_l(Q[lq,b,h,:] . K[l,b,h,:] / sqrt(32) + mask) V[l,b,h,:]
// head_dim = 32. Q/K/V are token-major with leading dimensions ldq/ldk/ldv (floats per token).
// CTA = 16 queries of one (b, h); 8 lanes per query split the keys of each 64-key tile and take them FOUR at a time: four
// independent 32-term dot products (the single dependent FMA chain per key was the latency bound of the first version,
// profiles/r02k_ncu_summary_all_kernels.txt), one running-maximum update per group, then the four rows of V.  Online softmax per
// lane, the 8 partial states of a query merged with shuffles at the end (fixed order => deterministic).
constexpr int AT_Q = 16, AT_LPQ = 8, AT_KT = 64, AT_HD = 32;
__global__ void __launch_bounds__(128) attention_kernel(const float* __restrict__ Q, const float* __restrict__ Kp,
                                                        const float* __restrict__ V, const unsigned char* __restrict__ mask,
                                                        float* __restrict__ O, int Lq, int L, int B, int H, int ldq, int ldk,
                                                        int ldv, int ldo, float scale) {
    __shared__ __align__(16) float Ks[AT_KT][AT_HD + 4];
    __shared__ __align__(16) float Vs[AT_KT][AT_HD + 4];
    __shared__ float Mb[AT_KT];                     // 0 for a live key, -inf for a masked / out-of-range one
    const int bh = blockIdx.y, b = bh / H, h = bh - b * H;
    const int qi = blockIdx.x * AT_Q + (threadIdx.x / AT_LPQ), part = threadIdx.x % AT_LPQ;
    const bool qv = qi < Lq;
    float q[AT_HD], acc[AT_HD];
#pragma unroll
    for (int d = 0; d < AT_HD; ++d) { q[d] = 0.f; acc[d] = 0.f; }
    if (qv) {
        const float4* qp = reinterpret_cast<const float4*>(Q + ((size_t)qi * B + b) * ldq + h * AT_HD);
#pragma unroll
        for (int d = 0; d < AT_HD / 4; ++d) {
            const float4 v4 = qp[d];
            q[4 * d] = v4.x * scale; q[4 * d + 1] = v4.y * scale; q[4 * d + 2] = v4.z * scale; q[4 * d + 3] = v4.w * scale;
        }
    }
    float mmax = -INFINITY, ssum = 0.f;
    for (int l0 = 0; l0 < L; l0 += AT_KT) {
        __syncthreads();
        for (int i = threadIdx.x; i < AT_KT * (AT_HD / 4); i += 128) {
            const int r = i / (AT_HD / 4), c = i - r * (AT_HD / 4);
            float4 kv = make_float4(0.f, 0.f, 0.f, 0.f), vv = kv;
            if (l0 + r < L) {
                kv = *reinterpret_cast<const float4*>(Kp + ((size_t)(l0 + r) * B + b) * ldk + h * AT_HD + 4 * c);
                vv = *reinterpret_cast<const float4*>(V + ((size_t)(l0 + r) * B + b) * ldv + h * AT_HD + 4 * c);
            }
            *reinterpret_cast<float4*>(&Ks[r][4 * c]) = kv;
            *reinterpret_cast<float4*>(&Vs[r][4 * c]) = vv;
        }
        if (threadIdx.x < AT_KT) {
            const int l = l0 + threadIdx.x;
            const bool dead = (l >= L) || (mask && mask[(size_t)b * L + l]);
            Mb[threadIdx.x] = dead ? -INFINITY : 0.f;
        }
        __syncthreads();
        // keys part, part + 8, part + 16, part + 24 and then the same + 32
#pragma unroll 1
        for (int g = 0; g < AT_KT / (4 * AT_LPQ); ++g) {
            const int j0 = g * 4 * AT_LPQ + part;
            float s[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int d4 = 0; d4 < AT_HD / 4; ++d4) {
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const float4 k4 = *reinterpret_cast<const float4*>(&Ks[j0 + u * AT_LPQ][4 * d4]);
                    s[u] = fmaf(q[4 * d4], k4.x, s[u]); s[u] = fmaf(q[4 * d4 + 1], k4.y, s[u]);
                    s[u] = fmaf(q[4 * d4 + 2], k4.z, s[u]); s[u] = fmaf(q[4 * d4 + 3], k4.w, s[u]);
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) s[u] += Mb[j0 + u * AT_LPQ];          // -inf removes the key
            const float gm = fmaxf(fmaxf(s[0], s[1]), fmaxf(s[2], s[3]));
            if (gm > mmax) {
                const float c = __expf(mmax - gm);       // exp(-inf) = 0 on the first live key
                ssum *= c;
#pragma unroll
                for (int d = 0; d < AT_HD; ++d) acc[d] *= c;
                mmax = gm;
            }
            if (mmax == -INFINITY) continue;             // nothing live so far
            float p[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) { p[u] = __expf(s[u] - mmax); ssum += p[u]; }
#pragma unroll
            for (int d4 = 0; d4 < AT_HD / 4; ++d4) {
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const float4 v4 = *reinterpret_cast<const float4*>(&Vs[j0 + u * AT_LPQ][4 * d4]);
                    acc[4 * d4] = fmaf(p[u], v4.x, acc[4 * d4]); acc[4 * d4 + 1] = fmaf(p[u], v4.y, acc[4 * d4 + 1]);
                    acc[4 * d4 + 2] = fmaf(p[u], v4.z, acc[4 * d4 + 2]); acc[4 * d4 + 3] = fmaf(p[u], v4.w, acc[4 * d4 + 3]);
                }
            }
        }
    }
    // merge the 8 partial softmax states of a query (lanes 8q .. 8q+7)
#pragma unroll
    for (int o = 1; o < AT_LPQ; o <<= 1) {
        const float m2 = __shfl_xor_sync(0xffffffffu, mmax, o), s2 = __shfl_xor_sync(0xffffffffu, ssum, o);
        const float mn = fmaxf(mmax, m2);
        const float c1 = (mmax == -INFINITY) ? 0.f : __expf(mmax - mn), c2 = (m2 == -INFINITY) ? 0.f : __expf(m2 - mn);
        ssum = ssum * c1 + s2 * c2;
#pragma unroll
        for (int d = 0; d < AT_HD; ++d) {
            const float a2 = __shfl_xor_sync(0xffffffffu, acc[d], o);
            acc[d] = acc[d] * c1 + a2 * c2;
        }
        mmax = mn;
    }
    if (qv) {
        const float inv = 1.f / ssum;
        float* op = O + ((size_t)qi * B + b) * ldo + h * AT_HD + part * (AT_HD / AT_LPQ);
#pragma unroll
        for (int d = 0; d < AT_HD / AT_LPQ; ++d) {
            // (acc is fully unrolled: select the lane's slice without dynamic register indexing)
            float v = 0.f;
#pragma unroll
            for (int pp = 0; pp < AT_LPQ; ++pp) v = (part == pp) ? acc[pp * (AT_HD / AT_LPQ) + d] : v;
            op[d] = v * inv;
        }
    }
}


// The decoder's cross attention has ONE query per (batch, head) (ToMP: a single foreground token): the general kernel above would
// walk the 972 keys in 16 CTAs of which 124 threads idle.  Here one CTA per (b, h): thread = key for the scores (two-pass softmax in
// shared memory), then warp w accumulates the keys w, w + 8, ... for all 32 output dims (lane = dim) and the 8 partial rows are summed
// in warp order (fixed order => deterministic).
__global__ void __launch_bounds__(256) attention_q1_kernel(const float* __restrict__ Q, const float* __restrict__ Kp, const float* __restrict__ V,
                                                           const unsigned char* __restrict__ mask, float* __restrict__ O, int L, int B, int H,
                                                           int ldq, int ldk, int ldv, int ldo, float scale) {
    extern __shared__ float sp[];                 // [L] scores / probabilities, then [8][32] partial outputs
    __shared__ float red[32];
    __shared__ float s_q[AT_HD];
    const int bh = blockIdx.x, b = bh / H, h = bh - b * H;
    if (threadIdx.x < AT_HD) s_q[threadIdx.x] = Q[(size_t)b * ldq + h * AT_HD + threadIdx.x] * scale;
    __syncthreads();
    float mx = -INFINITY;
    for (int l = threadIdx.x; l < L; l += blockDim.x) {
        float sc = -INFINITY;
        if (!(mask && mask[(size_t)b * L + l])) {
            const float4* kp = reinterpret_cast<const float4*>(Kp + ((size_t)l * B + b) * ldk + h * AT_HD);
            sc = 0.f;
#pragma unroll
            for (int d = 0; d < AT_HD / 4; ++d) {
                const float4 k4 = kp[d];
                sc = fmaf(s_q[4 * d], k4.x, sc); sc = fmaf(s_q[4 * d + 1], k4.y, sc); sc = fmaf(s_q[4 * d + 2], k4.z, sc); sc = fmaf(s_q[4 * d + 3], k4.w, sc);
            }
        }
        sp[l] = sc;
        mx = fmaxf(mx, sc);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
    __syncthreads();
    mx = red[0];
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) mx = fmaxf(mx, red[w]);
    __syncthreads();
    float sum = 0.f;
    for (int l = threadIdx.x; l < L; l += blockDim.x) {
        const float p = (sp[l] == -INFINITY) ? 0.f : __expf(sp[l] - mx);
        sp[l] = p;
        sum += p;
    }
    sum = block_sum(sum, red);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    float acc = 0.f;
    for (int l = warp; l < L; l += nw) acc = fmaf(sp[l], V[((size_t)l * B + b) * ldv + h * AT_HD + lane], acc);
    __syncthreads();
    float* part = sp + L;
    part[warp * AT_HD + lane] = acc;
    __syncthreads();
    if (warp == 0) {
        float o = 0.f;
        for (int w = 0; w < nw; ++w) o += part[w * AT_HD + lane];
        O[(size_t)b * ldo + h * AT_HD + lane] = o / sum;
    }
}

}  // namespace b200trk

using namespace b200trk;

struct b200trk_transformer {
    int D = 0, H = 0, FF = 0, L = 0, B = 0, n_enc = 0, n_dec = 0, M = 0;
    b200trk_net box;                      // container for the GEMM plan: buffers, split-K workspace, SM count
    std::vector<Op> gemms;                // tcgen05 GEMMs in execution order
    struct Lin { float *w = nullptr, *b = nullptr; };
    struct LN { float *g = nullptr, *b = nullptr; };
    struct Enc { Lin qk, v, o, f1, f2; LN n1, n2; int g_qk, g_v, g_o, g_f1, g_f2; };
    struct Dec { Lin sa_v, sa_o, ca_q, ca_k, ca_v, ca_o, f1, f2; LN n1, n2, n3; int g_k, g_v; };
    std::vector<Enc> enc;
    std::vector<Dec> dec;
    LN dec_norm;
    // activation buffer ids inside box.bufs
    int b_src = -1, b_qkin = -1, b_qk = -1, b_v = -1, b_att = -1, b_tmp = -1, b_ff = -1, b_mem_pos = -1, b_k = -1;
    float *pos_full = nullptr;            // [L,B,D] broadcast copy of pos when Bp == 1 (not needed: add_pos handles Bp)
    float *d_tgt = nullptr, *d_t1 = nullptr, *d_t2 = nullptr, *d_q = nullptr, *d_att = nullptr, *d_ff = nullptr, *d_qpos = nullptr;
};

static int t_alloc(b200trk_transformer* t, float** p, size_t floats) {
    void* q = nullptr;
    B200_CHECK_CUDA(cudaMalloc(&q, floats * sizeof(float)));
    B200_CHECK_CUDA(cudaMemset(q, 0, floats * sizeof(float)));
    t->box.owned.push_back(q);
    *p = (float*)q;
    return 0;
}
static int t_upload(b200trk_transformer* t, float** p, const float* host, size_t floats) {
    if (int e = t_alloc(t, p, floats)) return e;
    B200_CHECK_CUDA(cudaMemcpy(*p, host, floats * sizeof(float), cudaMemcpyHostToDevice));
    return 0;
}
static int t_buf(b200trk_transformer* t, size_t floats, int* id) {
    float* p = nullptr;
    if (int e = t_alloc(t, &p, floats)) return e;
    t->box.bufs.push_back(p);
    t->box.buf_floats.push_back(floats);
    *id = (int)t->box.bufs.size() - 1;
    return 0;
}
// register y[M,N] = act(x[M,K] W^T + b) (+ res) as a 1x1 convolution over an (Ht x Wt) "image" of tokens
static int t_gemm(b200trk_transformer* t, int in, int out, int res, const b200trk_transformer::Lin& lin, int K, int N, int relu, int* gid) {
    Op op;
    op.kind = OP_CONV; op.in = in; op.out = out; op.res = res;
    int Wt = 1;
    for (int w = 16; w >= 1; --w) if (t->M % w == 0) { Wt = w; break; }
    op.Win = op.Wout = Wt; op.Hin = op.Hout = t->M / Wt;
    op.Cin = K; op.Cout = N; op.k = 1; op.stride = 1; op.pad = 0; op.relu = relu;
    op.w = lin.w; op.bias = lin.b;
    std::vector<float> unused;
    if (int e = tc_conv_prepare(&t->box, op, unused)) return e;
    t->gemms.push_back(op);
    *gid = (int)t->gemms.size() - 1;
    return 0;
}

extern "C" int b200trk_transformer_destroy(b200trk_transformer_t* t) {
    if (!t) return 0;
    for (auto& op : t->gemms) if (op.tc) tc_conv_free(op.tc);
    for (void* p : t->box.owned) cudaFree(p);
    delete t;
    return 0;
}

extern "C" int b200trk_transformer_create(b200trk_transformer_t** out, const b200trk_enc_layer_t* enc, int n_enc,
                                          const b200trk_dec_layer_t* dec, int n_dec, const float* dec_norm_w,
                                          const float* dec_norm_b, int d_model, int nhead, int dim_ff, int L, int B) {
    B200_REQUIRE(out && enc && dec && dec_norm_w && dec_norm_b, "transformer_create: null pointer");
    B200_REQUIRE(d_model % 64 == 0 && d_model / nhead == 32 && dim_ff % 64 == 0, "transformer_create: d_model=%d nhead=%d dim_ff=%d not supported (head_dim must be 32, widths multiples of 64)", d_model, nhead, dim_ff);
    B200_REQUIRE(L >= 1 && B >= 1 && B <= 8 && (size_t)L * B <= 65536, "transformer_create: L=%d B=%d out of range", L, B);
    b200trk_transformer* t = new b200trk_transformer();
    t->D = d_model; t->H = nhead; t->FF = dim_ff; t->L = L; t->B = B; t->n_enc = n_enc; t->n_dec = n_dec; t->M = L * B;
    t->box.max_batch = 1; t->box.precision = 0; t->box.sms = device_sm_count();
    const int D = d_model, M = t->M;
    int e = 0;
    t->box.splitk_ws_floats = (size_t)4 << 20;
    if (!e) e = t_alloc(t, &t->box.splitk_ws, t->box.splitk_ws_floats);
    if (!e) e = t_buf(t, (size_t)M * D, &t->b_src);
    if (!e) e = t_buf(t, (size_t)M * D, &t->b_qkin);
    if (!e) e = t_buf(t, (size_t)M * 2 * D, &t->b_qk);
    if (!e) e = t_buf(t, (size_t)M * D, &t->b_v);
    if (!e) e = t_buf(t, (size_t)M * D, &t->b_att);
    if (!e) e = t_buf(t, (size_t)M * D, &t->b_tmp);
    if (!e) e = t_buf(t, (size_t)M * dim_ff, &t->b_ff);
    if (!e) e = t_buf(t, (size_t)M * D, &t->b_mem_pos);
    if (!e) e = t_buf(t, (size_t)M * D, &t->b_k);
    auto up_lin = [&](b200trk_transformer::Lin& l, const float* w, const float* b, size_t rows, size_t cols) {
        if (!e) e = t_upload(t, &l.w, w, rows * cols);
        if (!e && b) e = t_upload(t, &l.b, b, rows);
    };
    auto up_ln = [&](b200trk_transformer::LN& n, const float* g, const float* b) {
        if (!e) e = t_upload(t, &n.g, g, D);
        if (!e) e = t_upload(t, &n.b, b, D);
    };
    t->enc.resize(n_enc);
    for (int i = 0; i < n_enc && !e; ++i) {
        const b200trk_enc_layer_t& w = enc[i];
        auto& E = t->enc[i];
        up_lin(E.qk, w.self_attn.in_proj_weight, w.self_attn.in_proj_bias, 2 * D, D);                                  // rows [0, 2D): q and k
        up_lin(E.v, w.self_attn.in_proj_weight + (size_t)2 * D * D, w.self_attn.in_proj_bias + 2 * D, D, D);            // rows [2D, 3D): v
        up_lin(E.o, w.self_attn.out_proj_weight, w.self_attn.out_proj_bias, D, D);
        up_lin(E.f1, w.linear1_weight, w.linear1_bias, dim_ff, D);
        up_lin(E.f2, w.linear2_weight, w.linear2_bias, D, dim_ff);
        up_ln(E.n1, w.norm1_weight, w.norm1_bias);
        up_ln(E.n2, w.norm2_weight, w.norm2_bias);
        if (!e) e = t_gemm(t, t->b_qkin, t->b_qk, -1, E.qk, D, 2 * D, 0, &E.g_qk);
        if (!e) e = t_gemm(t, t->b_src, t->b_v, -1, E.v, D, D, 0, &E.g_v);
        if (!e) e = t_gemm(t, t->b_att, t->b_tmp, t->b_src, E.o, D, D, 0, &E.g_o);          // + residual src
        if (!e) e = t_gemm(t, t->b_src, t->b_ff, -1, E.f1, D, dim_ff, 1, &E.g_f1);          // ReLU
        if (!e) e = t_gemm(t, t->b_ff, t->b_tmp, t->b_src, E.f2, dim_ff, D, 0, &E.g_f2);    // + residual src
    }
    t->dec.resize(n_dec);
    for (int i = 0; i < n_dec && !e; ++i) {
        const b200trk_dec_layer_t& w = dec[i];
        auto& Dl = t->dec[i];
        up_lin(Dl.sa_v, w.self_attn.in_proj_weight + (size_t)2 * D * D, w.self_attn.in_proj_bias + 2 * D, D, D);
        up_lin(Dl.sa_o, w.self_attn.out_proj_weight, w.self_attn.out_proj_bias, D, D);
        up_lin(Dl.ca_q, w.cross_attn.in_proj_weight, w.cross_attn.in_proj_bias, D, D);
        up_lin(Dl.ca_k, w.cross_attn.in_proj_weight + (size_t)D * D, w.cross_attn.in_proj_bias + D, D, D);
        up_lin(Dl.ca_v, w.cross_attn.in_proj_weight + (size_t)2 * D * D, w.cross_attn.in_proj_bias + 2 * D, D, D);
        up_lin(Dl.ca_o, w.cross_attn.out_proj_weight, w.cross_attn.out_proj_bias, D, D);
        up_lin(Dl.f1, w.linear1_weight, w.linear1_bias, dim_ff, D);
        up_lin(Dl.f2, w.linear2_weight, w.linear2_bias, D, dim_ff);
        up_ln(Dl.n1, w.norm1_weight, w.norm1_bias);
        up_ln(Dl.n2, w.norm2_weight, w.norm2_bias);
        up_ln(Dl.n3, w.norm3_weight, w.norm3_bias);
        if (!e) e = t_gemm(t, t->b_mem_pos, t->b_k, -1, Dl.ca_k, D, D, 0, &Dl.g_k);        // K = W_k (memory + pos)
        if (!e) e = t_gemm(t, t->b_src, t->b_v, -1, Dl.ca_v, D, D, 0, &Dl.g_v);            // V = W_v memory
    }
    up_ln(t->dec_norm, dec_norm_w, dec_norm_b);
    if (!e) e = t_alloc(t, &t->d_tgt, (size_t)B * D);
    if (!e) e = t_alloc(t, &t->d_t1, (size_t)B * D);
    if (!e) e = t_alloc(t, &t->d_t2, (size_t)B * D);
    if (!e) e = t_alloc(t, &t->d_q, (size_t)B * D);
    if (!e) e = t_alloc(t, &t->d_att, (size_t)B * D);
    if (!e) e = t_alloc(t, &t->d_ff, (size_t)B * dim_ff);
    if (!e) e = t_alloc(t, &t->d_qpos, (size_t)B * D);
    if (e) { b200trk_transformer_destroy(t); return e; }
    *out = t;
    return 0;
}

extern "C" int b200trk_transformer_forward(b200trk_transformer_t* t, const float* src, const float* pos, int Bp,
                                           const unsigned char* key_padding_mask, const float* query_embed, float* hs,
                                           float* memory, b200trk_stream_t stream) {
    B200_REQUIRE(t && src && pos && query_embed && hs && memory, "transformer_forward: null pointer");
    B200_REQUIRE(Bp == 1 || Bp == t->B, "transformer_forward: pos batch %d must be 1 or %d", Bp, t->B);
    cudaStream_t st = (cudaStream_t)stream;
    const int D = t->D, M = t->M, L = t->L, B = t->B, H = t->H, FF = t->FF;
    auto buf = [&](int id) { return t->box.bufs[id]; };
    const float scale = 1.0f / sqrtf(32.f);
    auto gemm = [&](int gid) { return tc_conv_launch(&t->box, t->gemms[gid], 1, st); };
    auto ln = [&](const float* x, const b200trk_transformer::LN& n, float* y, int T) {
        layernorm_kernel<<<(T * 32 + 255) / 256, 256, 0, st>>>(x, n.g, n.b, y, T, D);
        g_launch_count.fetch_add(1, std::memory_order_relaxed);
        return cudaGetLastError() == cudaSuccess ? 0 : 1;
    };
    auto addpos = [&](const float* x, float* y) {
        const int n4 = M * D / 4;
        add_pos_kernel<<<(n4 + 255) / 256, 256, 0, st>>>((const float4*)x, (const float4*)pos, (float4*)y, L, B, Bp, D / 4);
        g_launch_count.fetch_add(1, std::memory_order_relaxed);
        return cudaGetLastError() == cudaSuccess ? 0 : 1;
    };
    auto small = [&](const float* x, const b200trk_transformer::Lin& lin, const float* res, float* y, int K, int N, int relu) {
        small_linear_kernel<<<(B * N * 32 + 255) / 256, 256, 0, st>>>(x, lin.w, lin.b, res, y, B, K, N, relu);
        g_launch_count.fetch_add(1, std::memory_order_relaxed);
        return cudaGetLastError() == cudaSuccess ? 0 : 1;
    };
    B200_CHECK_CUDA(cudaMemcpyAsync(buf(t->b_src), src, (size_t)M * D * sizeof(float), cudaMemcpyDeviceToDevice, st));
    // ---------------- encoder ----------------
    for (int i = 0; i < t->n_enc; ++i) {
        auto& E = t->enc[i];
        if (addpos(buf(t->b_src), buf(t->b_qkin))) { set_error("transformer: add_pos launch failed"); return 1; }
        if (int e = gemm(E.g_qk)) return e;                              // [M, 2D] = (src + pos) [Wq; Wk]^T + b
        if (int e = gemm(E.g_v)) return e;                               // [M, D]  = src Wv^T + b
        attention_kernel<<<dim3((L + AT_Q - 1) / AT_Q, B * H), 128, 0, st>>>(buf(t->b_qk), buf(t->b_qk) + D, buf(t->b_v),
                                                                              key_padding_mask, buf(t->b_att), L, L, B, H, 2 * D,
                                                                              2 * D, D, D, scale);
        B200_LAUNCH_CHECK();
        if (int e = gemm(E.g_o)) return e;                               // tmp = src + att Wo^T + b
        if (ln(buf(t->b_tmp), E.n1, buf(t->b_src), M)) { set_error("transformer: layernorm launch failed"); return 1; }
        if (int e = gemm(E.g_f1)) return e;                              // ff = relu(src W1^T + b1)
        if (int e = gemm(E.g_f2)) return e;                              // tmp = src + ff W2^T + b2
        if (ln(buf(t->b_tmp), E.n2, buf(t->b_src), M)) { set_error("transformer: layernorm launch failed"); return 1; }
    }
    B200_CHECK_CUDA(cudaMemcpyAsync(memory, buf(t->b_src), (size_t)M * D * sizeof(float), cudaMemcpyDeviceToDevice, st));
    // ---------------- decoder (one query per batch element) ----------------
    if (addpos(buf(t->b_src), buf(t->b_mem_pos))) { set_error("transformer: add_pos launch failed"); return 1; }
    B200_CHECK_CUDA(cudaMemsetAsync(t->d_tgt, 0, (size_t)B * D * sizeof(float), st));
    for (int b = 0; b < B; ++b)
        B200_CHECK_CUDA(cudaMemcpyAsync(t->d_qpos + (size_t)b * D, query_embed, D * sizeof(float), cudaMemcpyDeviceToDevice, st));
    for (int i = 0; i < t->n_dec; ++i) {
        auto& Dl = t->dec[i];
        // self attention over a single token: softmax over one key is 1, so the output is out_proj(v_proj(tgt))
        if (small(t->d_tgt, Dl.sa_v, nullptr, t->d_t1, D, D, 0)) return 1;
        if (small(t->d_t1, Dl.sa_o, t->d_tgt, t->d_t2, D, D, 0)) return 1;
        if (ln(t->d_t2, Dl.n1, t->d_tgt, B)) return 1;
        // cross attention: q = W_q (tgt + query_pos)
        add_pos_kernel<<<(B * D / 4 + 255) / 256, 256, 0, st>>>((const float4*)t->d_tgt, (const float4*)t->d_qpos, (float4*)t->d_t1, 1, B, B, D / 4);
        B200_LAUNCH_CHECK();
        if (small(t->d_t1, Dl.ca_q, nullptr, t->d_q, D, D, 0)) return 1;
        if (int e = gemm(Dl.g_k)) return e;
        if (int e = gemm(Dl.g_v)) return e;
        attention_q1_kernel<<<B * H, 256, (size_t)(L + 8 * AT_HD) * sizeof(float), st>>>(t->d_q, buf(t->b_k), buf(t->b_v), key_padding_mask,
                                                                                         t->d_att, L, B, H, D, D, D, D, scale);
        B200_LAUNCH_CHECK();
        if (small(t->d_att, Dl.ca_o, t->d_tgt, t->d_t2, D, D, 0)) return 1;
        if (ln(t->d_t2, Dl.n2, t->d_tgt, B)) return 1;
        if (small(t->d_tgt, Dl.f1, nullptr, t->d_ff, D, FF, 1)) return 1;
        if (small(t->d_ff, Dl.f2, t->d_tgt, t->d_t2, FF, D, 0)) return 1;
        if (ln(t->d_t2, Dl.n3, t->d_tgt, B)) return 1;
    }
    if (ln(t->d_tgt, t->dec_norm, hs, B)) return 1;
    return 0;
}
