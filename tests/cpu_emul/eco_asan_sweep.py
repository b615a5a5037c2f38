"""TEST INFRASTRUCTURE ONLY -- numpy / ctypes driver (no torch) that runs the two ECO kernels (tests/cpu_emul/eco_emul.cpp) at the per-CTA
configurations of ECO's real block sizes and at ragged ones, on random inputs; tests/test_eco_cpu.py runs it in a subprocess under
AddressSanitizer: an access outside a launch's dynamic shared memory or outside a global buffer (which the device may or may not fault on)
is an error there.

    python eco_asan_sweep.py <libeco_emul_asan.so>"""
import sys, numpy as np, ctypes as C
lib = C.CDLL(sys.argv[1])
rng = np.random.RandomState(0)
P = lambda a: a.ctypes.data_as(C.c_void_p) if a is not None else None
f32 = lambda *s: np.ascontiguousarray(rng.randn(*s).astype(np.float32))
reg3 = np.array([[0.0, 0.23, 0.0], [0.16, 0.78, 0.16], [0.0, 0.23, 0.0]], np.float32).reshape(1, 1, 3, 3)
for (h, wh, n, c, ctas) in [(3, 3, 200, 64, 9), (7, 8, 200, 16, 4), (3, 3, 60, 128, 9), (5, 4, 33, 32, 6)]:
    hf, S, yf = 0.01 * f32(1, c, h, wh, 2), f32(h, wh, n, c, 2), np.abs(f32(1, 1, h, wh))
    sw = np.abs(f32(n)); sw /= sw.sum()
    en, p, rp, rho, plan = np.zeros((1, c, h, wh), np.float32), np.zeros_like(hf), np.zeros_like(hf), np.ones(1, np.float32), (C.c_int * 6)()
    for r in range(2):
        nx = f32(1, c, h, wh, 2)
        rc = lib.eco_emul_filter_cg(P(hf), P(S), P(yf), P(sw), P(reg3), 3, 3, P(en), int(r > 0), P(nx), P(p), P(rp), P(rho), int(r > 0), h, wh, n, c, 3, 0, 1,
                                    C.c_float(0.5), C.c_float(0.0075), C.c_float(0.3), C.c_float(0.15), ctas, 256, -1, plan)
        assert rc == 0 and np.isfinite(hf).all()
    print("online", (h, wh, n, c), list(plan), flush=True)
for (h, wh, n, cin, c, ctas) in [(3, 3, 30, 256, 64, 9), (7, 8, 30, 96, 16, 4), (5, 5, 40, 64, 32, 3), (4, 3, 7, 24, 16, 5)]:
    S, yf = f32(h, wh, n, cin, 2), np.abs(f32(1, 1, h, wh))
    P0 = np.ascontiguousarray(np.linalg.qr(rng.randn(cin, cin))[0][:, :c].astype(np.float32))
    hf, dMh, plan = np.zeros((1, c, h, wh, 2), np.float32), (np.abs(f32(1, c, h, wh)) + 0.5), (C.c_int * 6)()
    sws = np.full(n, (1.0 / n) ** 0.5, np.float32)
    rc = lib.eco_emul_joint_gn(P(hf), P(P0), P(S), P(yf), P(sws), P(reg3), 3, 3, P(dMh), C.c_float(35.0), C.c_float(5e-8), h, wh, n, cin, c, 3, 2, ctas, 256, -1, plan)
    assert rc == 0 and np.isfinite(hf).all() and np.isfinite(P0).all()
    print("joint", (h, wh, n, cin, c), list(plan), flush=True)
print("EMUL_DONE")
