"""CPU: pins the PrRoIPool restatement (oracle/prroi_oracle.py) -- the reference's own known-answer test, quadrature of
the bilinear interpolant, adjointness of forward/backward, finite differences of the coordinate gradient."""
import numpy as np
import torch
import torch.nn.functional as F

from oracle import prroi_oracle as P


def test_known_answer_avg_pool():
    """ltr/external/PreciseRoIPooling/pytorch/tests/test_prroi_pooling2d.py:21-35."""
    feat = torch.rand(4, 16, 24, 32, generator=torch.Generator().manual_seed(0))
    rois = np.array([[0, 0, 0, 14, 14], [1, 14, 14, 28, 28]], dtype=np.float32)
    out = P.forward(feat.numpy(), rois, 7, 7, 0.5)
    ref = F.avg_pool2d(feat, kernel_size=2, stride=1).numpy()
    assert np.allclose(out[0], ref[0, :, :7, :7], atol=1e-6)
    assert np.allclose(out[1], ref[1, :, 7:14, 7:14], atol=1e-6)


def _bilinear(data, y, x):
    h0, w0 = int(np.floor(y)), int(np.floor(x))
    v = 0.0
    for dh in (0, 1):
        for dw in (0, 1):
            hh, ww = h0 + dh, w0 + dw
            if 0 <= hh < data.shape[0] and 0 <= ww < data.shape[1]:
                v += data[hh, ww] * (1 - abs(y - hh)) * (1 - abs(x - ww))
    return v


def test_forward_is_the_integral_of_the_bilinear_interpolant():
    rng = np.random.RandomState(1)
    feat = rng.rand(1, 2, 9, 11)
    rois = np.array([[0, 1.3, 2.2, 8.6, 7.9]])
    out = P.forward(feat, rois, 2, 3, 1.0)
    bw, bh = (8.6 - 1.3) / 3, (7.9 - 2.2) / 2
    m = 200
    for c in range(2):
        for i in range(2):
            for j in range(3):
                ys = 2.2 + bh * i + (np.arange(m) + 0.5) * bh / m
                xs = 1.3 + bw * j + (np.arange(m) + 0.5) * bw / m
                q = np.mean([[_bilinear(feat[0, c], y, x) for x in xs] for y in ys])
                assert abs(q - out[0, c, i, j]) < 2e-4


def test_backward_is_the_adjoint_and_coor_backward_matches_finite_differences():
    rng = np.random.RandomState(2)
    feat = rng.randn(2, 3, 10, 12)
    rois = np.array([[0, 2.1, 1.7, 9.4, 8.2], [1, 0.5, 3.0, 6.5, 9.5]])
    g = rng.randn(2, 3, 3, 3)
    out = P.forward(feat, rois, 3, 3, 0.9)
    fg = P.backward(feat, rois, g, 3, 3, 0.9)
    x2 = rng.randn(*feat.shape)
    lhs = np.sum(P.forward(x2, rois, 3, 3, 0.9) * g)
    rhs = np.sum(x2 * fg)
    assert abs(lhs - rhs) < 1e-9 * max(1.0, abs(lhs))
    rg = P.coor_backward(feat, rois, out, g, 3, 3, 0.9)
    eps = 1e-5
    for n in range(2):
        for k in range(1, 5):
            rp, rm = rois.copy(), rois.copy()
            rp[n, k] += eps
            rm[n, k] -= eps
            fd = (np.sum(P.forward(feat, rp, 3, 3, 0.9) * g) - np.sum(P.forward(feat, rm, 3, 3, 0.9) * g)) / (2 * eps)
            assert abs(fd - rg[n, k]) < 1e-5 * max(1.0, abs(fd)), (n, k, fd, rg[n, k])
    assert np.all(rg[:, 0] == 0)
