"""TEST INFRASTRUCTURE ONLY -- numpy / ctypes driver that pushes one representative call through every multi-threaded kernel of the
`*_emul.cpp` harnesses (attention, single-query attention, LayerNorm; apply_filter with the last-CTA reduction, apply_feat_transpose, max2d;
the three PrRoIPool kernels incl. the atomics of the backward; feature normalisation, softmax_reg, conv1x1, Fourier interpolation; the stem,
the implicit-GEMM convolution fused and split-K, InstanceL2Norm + export; one CTA of the persistent steepest-descent optimiser and of
ATOM's CG kernel with the immediate form of their cp.async copies -- the schedule that exposes a pipeline stage reused too early; ECO's
score kernels).  tests/test_kernels_tsan_cpu.py runs it in a subprocess under
ThreadSanitizer (no torch import): every hand-over between the threads of a block must be ordered by a barrier.

    python tsan_sweep.py <dir with lib{transformer,corr,prroi,atom,conv_fp32,sd,cg,eco_loc}_tsan.so>"""
import sys
LIB = {n: "%s/lib%s_tsan.so" % (sys.argv[1], n) for n in ("transformer", "corr", "prroi", "atom", "conv_fp32", "sd", "cg", "eco_loc")}
import ctypes as C, numpy as np
rng = np.random.RandomState(0)
P = lambda a: a.ctypes.data_as(C.c_void_p) if a is not None else None
f32 = lambda *s: np.ascontiguousarray(rng.randn(*s).astype(np.float32))
# transformer kernels
L = C.CDLL(LIB["transformer"])
Ln, B, H = 40, 2, 2; D = H * 32
qk, v, out = f32(Ln, B, 2 * D), f32(Ln, B, D), np.zeros((Ln, B, D), np.float32)
m = np.zeros((B, Ln), np.uint8); m[1, 5:9] = 1
L.tr_emul_attention(P(qk), C.c_void_p(qk.ctypes.data + D * 4), P(v), P(m), P(out), Ln, Ln, B, H, 2 * D, 2 * D, D, D, C.c_float(0.17))
q1, o1 = f32(B, D), np.zeros((B, D), np.float32)
L.tr_emul_attention_q1(P(q1), C.c_void_p(qk.ctypes.data + D * 4), P(v), P(m), P(o1), Ln, B, H, D, 2 * D, D, D, C.c_float(0.17))
x, g, b, y = f32(9, 256), f32(256), f32(256), np.zeros((9, 256), np.float32)
L.tr_emul_layernorm(P(x), P(g), P(b), P(y), 9, 256)
print("transformer ok")
# corr kernels
L = C.CDLL(LIB["corr"])
feat, w = f32(2, 64, 18, 18), f32(1, 64, 4, 4)
s, mv, mi = np.zeros((2, 1, 19, 19), np.float32), np.zeros(2, np.float32), np.zeros((2, 2), np.int64)
assert L.corr_emul_apply_filter(P(feat), P(w), P(s), 2, 64, 18, 18, P(mv), P(mi), 0) == 0
r, gr = f32(2, 1, 19, 19), np.zeros((1, 64, 4, 4), np.float32)
assert L.corr_emul_feat_transpose(P(feat), P(r), P(gr), 2, 64, 18, 18, 8) == 0
L.corr_emul_max2d(P(s), 2, 19, 19, P(mv), P(mi))
print("corr ok")
# prroi
L = C.CDLL(LIB["prroi"])
feat = f32(1, 8, 12, 12); rois = np.array([[0, 1.5, 2.0, 30.0, 40.0], [0, 10, 10, 60, 50]], np.float32)
out = np.zeros((2, 8, 3, 3), np.float32)
L.prroi_emul_forward(P(feat), P(rois), P(out), 1, 8, 12, 12, 2, 3, 3, C.c_float(0.25))
og, fg, rg = f32(2, 8, 3, 3), np.zeros_like(feat), np.zeros((2, 5), np.float32)
L.prroi_emul_backward(P(rois), P(og), P(fg), 1, 8, 12, 12, 2, 3, 3, C.c_float(0.25))
L.prroi_emul_coor_backward(P(feat), P(rois), P(out), P(og), P(rg), 1, 8, 12, 12, 2, 3, 3, C.c_float(0.25))
print("prroi ok")
# atom ops
L = C.CDLL(LIB["atom"])
x = f32(2, 16, 10, 10); L.atom_emul_feature_normalize(P(x), 2, 16, 10, 10, C.c_float(2.0))
xs, ys = f32(3, 361), np.zeros((3, 361), np.float32); L.atom_emul_softmax_reg(P(xs), P(ys), 3, 361, 1, C.c_float(0.1))
xi, Pm, o = f32(1, 40, 9, 9), f32(24, 40), np.zeros((1, 24, 9, 9), np.float32); L.atom_emul_conv1x1(P(xi), P(Pm), P(o), 1, 40, 24, 9, 9)
sc, up = f32(1, 1, 18, 18), np.zeros((1, 1, 40, 40), np.float32); L.atom_emul_fourier_interp(P(sc), P(up), 1, 18, 18, 4, 4, 40, 40)
print("atom ok")
# conv_fp32
L = C.CDLL(LIB["conv_fp32"])
pre, wt, bb, st = f32(1, 20, 22, 4), f32(49, 3, 64), f32(64), np.zeros((1, 10, 11, 64), np.float32)
L.c32_emul_stem(P(pre), P(wt), P(bb), P(st), 1, 20, 22)
xin, wk, oo = f32(1, 8, 8, 32), f32(16, 3, 3, 32), np.zeros((1, 8, 8, 16), np.float32)
L.c32_emul_conv(P(xin), P(wk), P(oo), 1, 8, 8, 32, 16, 3, 1, 1, P(f32(16)), None, 1, 148, None)
L.c32_emul_conv(P(xin), P(wk), P(oo), 1, 8, 8, 32, 16, 3, 1, 1, None, None, 0, 0, None)
xe, oe = f32(1, 9, 40), np.zeros((1, 40, 9), np.float32); L.c32_emul_export(P(xe), P(oe), 1, 9, 40, 1, C.c_float(0.1), C.c_float(1e-5))
print("conv_fp32 ok")
# persistent optimiser kernels, one CTA (static __shared__ storage): DiMP steepest descent and ATOM CG on the cp.async sweeps
L = C.CDLL(LIB["sd"])
n, c, h, it = 4, 32, 18, 2
feat, w0, w = f32(n, c, h, h), f32(1, c, 4, 4) * 0.1, np.zeros((1, c, 4, 4), np.float32)
bb = np.ascontiguousarray(np.array([[100, 110, 60, 50]] * n, np.float32) + rng.rand(n, 4).astype(np.float32) * 10)
luts = [np.ascontiguousarray(rng.rand(100).astype(np.float32)) for _ in range(3)]
its, losses = np.zeros((it + 1, c, 4, 4), np.float32), np.zeros(it + 1, np.float32)
assert L.sd_emul_dimp_sd_gn(P(w0), P(w), P(feat), P(bb), None, n, c, h, h, it, P(luts[0]), P(luts[1]), P(luts[2]), 100, C.c_float(0.1), C.c_float(16.0),
                            C.c_float(0.9), C.c_float(0.01), C.c_float(0.0), P(its), P(losses)) == 0
print("sd ok")
L = C.CDLL(LIB["cg"])
y, sw, out = f32(n, 1, h, h), np.ascontiguousarray(np.full(n, 1.0 / n, np.float32)), np.zeros((1, c, 4, 4), np.float32)
assert L.cg_emul_atom_cg_filter(P(w0), P(out), P(feat), P(y), P(sw), n, c, h, h, 2, C.c_float(0.1), 0, 3, C.c_float(0.05)) == 0
print("cg ok")
# ECO score kernels: apply_filter, then sum_fs + sample_fs of two blocks on an odd grid
L = C.CDLL(LIB["eco_loc"])
blocks = [(9, 5, 6), (5, 3, 4)]
sfs = []
for (hh, wh, cc) in blocks:
    hf, xf, sf = f32(1, cc, hh, wh, 2), f32(2, cc, hh, wh, 2), np.zeros((2, 1, hh, wh, 2), np.float32)
    assert L.eco_loc_emul_apply_filter(P(hf), P(xf), P(sf), 2, cc, hh, wh) == 0
    sfs.append(sf)
sc = np.zeros((2, 1, 13, 11), np.float32)
ptrs = (C.c_void_p * 2)(*[a.ctypes.data for a in sfs])
assert L.eco_loc_emul_sample_fs(ptrs, (C.c_int * 2)(9, 5), (C.c_int * 2)(5, 3), (C.c_float * 2)(1.0, 0.5), 2, 2, 13, 11, P(sc)) == 0 and np.isfinite(sc).all()
xs, win = f32(2, 3, 6, 8), np.abs(f32(1, 1, 6, 8))
xv = np.ascontiguousarray(xs.transpose(2, 3, 0, 1)).transpose(2, 3, 0, 1)                 # the tracker's permuted view
iy, ix, xo = f32(1, 1, 7, 1, 2), f32(1, 1, 1, 5, 2), np.zeros((2, 3, 7, 5, 2), np.float32)
st = [C.c_longlong(v // 4) for v in xv.strides]
assert L.eco_loc_emul_preprocess(C.c_void_p(xv.transpose(2, 3, 0, 1).ctypes.data), st[0], st[1], st[2], st[3], P(win), P(iy), P(ix), P(xo), 2, 3, 6, 8) == 0
assert np.isfinite(xo).all()
ash, osh = f32(2, 3, 7, 4, 2), np.zeros((2, 3, 7, 4, 2), np.float32)
assert L.eco_loc_emul_shift_fs(P(ash), P(osh), 2, 3, 7, 4, C.c_float(0.3), C.c_float(-1.1)) == 0 and np.isfinite(osh).all()
print("eco_loc ok")
print("EMUL_DONE")
