"""TEST INFRASTRUCTURE ONLY -- CPU restatement of Precise RoI Pooling (the reference's one native component).

Follows ltr/external/PreciseRoIPooling/src/prroi_pooling_gpu_impl.cu kernel by kernel (forward :149-212 with
PrRoIPoolingMatCalculation :71-106; backward :214-272 with MatDistributeDiff :115-147; coordinate backward :274-379
with SingleCoorIntegral :50-52 and Interpolation :54-69), numpy float64, vectorised over channels only.
Parity status: pinned by the reference's own known-answer test (PrRoIPool == avg_pool2d on integer-aligned RoIs,
pytorch/tests/test_prroi_pooling2d.py:21-35, replayed in tests/test_prroi_oracle.py), by quadrature of the bilinear
interpolant, by adjointness of forward/backward and by finite differences of the box coordinates.

`RefModuleCPU` / `patch_reference_function` let the UNMODIFIED reference run its PrRoIPool call sites on CPU tensors
(the reference module is CUDA-only, functional.py:62-63) when golden vectors are generated (oracle/ref_shims.py).
"""
import math

import numpy as np
import torch


def _get(data, h, w):
    """PrRoIPoolingGetData for all channels: data [C,H,W] -> [C] (zero outside)."""
    if h < 0 or w < 0 or h >= data.shape[1] or w >= data.shape[2]:
        return np.zeros(data.shape[0], dtype=np.float64)
    return data[:, h, w]


def _axis_w(lo, hi):
    return hi - 0.5 * hi * hi - lo + 0.5 * lo * lo


def _mat_weights(s_h, s_w, y0, x0, y1, x1):
    """The four corner weights of PrRoIPoolingMatCalculation for cell (s_h, s_w) .. (s_h+1, s_w+1)."""
    e_h, e_w = s_h + 1, s_w + 1
    a_lo, a_hi = x0 - s_w, x1 - s_w
    b_lo, b_hi = y0 - s_h, y1 - s_h
    ar_lo, ar_hi = e_w - x1, e_w - x0
    br_lo, br_hi = e_h - y1, e_h - y0
    return ((s_h, s_w, _axis_w(a_lo, a_hi) * _axis_w(b_lo, b_hi)),
            (s_h, e_w, _axis_w(ar_lo, ar_hi) * _axis_w(b_lo, b_hi)),
            (e_h, s_w, _axis_w(a_lo, a_hi) * _axis_w(br_lo, br_hi)),
            (e_h, e_w, _axis_w(ar_lo, ar_hi) * _axis_w(br_lo, br_hi)))


def _bins(roi, ph, pw, scale):
    x1, y1, x2, y2 = [float(v) * scale for v in roi[1:5]]
    rw, rh = max(x2 - x1, 0.0), max(y2 - y1, 0.0)
    return x1, y1, rw / pw, rh / ph


def _cells(ws_h, ws_w, we_h, we_w):
    for w_it in range(int(math.floor(ws_w)), int(math.ceil(we_w))):
        for h_it in range(int(math.floor(ws_h)), int(math.ceil(we_h))):
            yield h_it, w_it, max(ws_h, float(h_it)), max(ws_w, float(w_it)), min(we_h, h_it + 1.0), min(we_w, w_it + 1.0)


def forward(features, rois, ph, pw, scale):
    f = np.asarray(features, dtype=np.float64)
    r = np.asarray(rois, dtype=np.float64)
    out = np.zeros((r.shape[0], f.shape[1], ph, pw), dtype=np.float64)
    for n in range(r.shape[0]):
        data = f[int(r[n, 0])]
        x1, y1, bw, bh = _bins(r[n], ph, pw, scale)
        win = max(0.0, bw * bh)
        if win == 0:
            continue
        for i in range(ph):
            for j in range(pw):
                ws_w, ws_h = x1 + bw * j, y1 + bh * i
                acc = np.zeros(f.shape[1])
                for h_it, w_it, y0, x0, yy1, xx1 in _cells(ws_h, ws_w, ws_h + bh, ws_w + bw):
                    for hh, ww, wt in _mat_weights(h_it, w_it, y0, x0, yy1, xx1):
                        acc += _get(data, hh, ww) * wt
                out[n, :, i, j] = acc / win
    return out


def backward(features, rois, output_grad, ph, pw, scale):
    f = np.asarray(features, dtype=np.float64)
    r = np.asarray(rois, dtype=np.float64)
    g = np.asarray(output_grad, dtype=np.float64)
    fg = np.zeros_like(f)
    H, W = f.shape[2], f.shape[3]
    for n in range(r.shape[0]):
        b = int(r[n, 0])
        x1, y1, bw, bh = _bins(r[n], ph, pw, scale)
        win = max(0.0, bw * bh)
        if win == 0:
            continue
        for i in range(ph):
            for j in range(pw):
                ws_w, ws_h = x1 + bw * j, y1 + bh * i
                top = g[n, :, i, j] / win
                for h_it, w_it, y0, x0, yy1, xx1 in _cells(ws_h, ws_w, ws_h + bh, ws_w + bw):
                    for hh, ww, wt in _mat_weights(h_it, w_it, y0, x0, yy1, xx1):
                        if 0 <= hh < H and 0 <= ww < W:
                            fg[b, :, hh, ww] += top * wt
    return fg


def _interp(data, h, w):
    """PrRoIPoolingInterpolation for all channels."""
    ret = np.zeros(data.shape[0])
    for dh in (0, 1):
        for dw in (0, 1):
            h1, w1 = int(math.floor(h)) + dh, int(math.floor(w)) + dw
            ret += _get(data, h1, w1) * ((1.0 - abs(h - h1)) * (1.0 - abs(w - w1)))
    return ret


def _single(s, t, c1, c2):
    return 0.5 * (t * t - s * s) * c2 + (t - 0.5 * t * t - s + 0.5 * s * s) * c1


def coor_backward(features, rois, output, output_grad, ph, pw, scale):
    f = np.asarray(features, dtype=np.float64)
    r = np.asarray(rois, dtype=np.float64)
    o = np.asarray(output, dtype=np.float64)
    g = np.asarray(output_grad, dtype=np.float64)
    rg = np.zeros((r.shape[0], 5), dtype=np.float64)
    for n in range(r.shape[0]):
        data = f[int(r[n, 0])]
        x1, y1, bw, bh = _bins(r[n], ph, pw, scale)
        win = max(0.0, bw * bh)
        if win == 0:
            continue
        for i in range(ph):
            for j in range(pw):
                ws_w, ws_h = x1 + bw * j, y1 + bh * i
                we_w, we_h = ws_w + bw, ws_h + bh
                og = g[n, :, i, j]
                active = (og / win) != 0                     # the reference returns early when sum_out == 0 (:317-318)
                gx1 = np.zeros(f.shape[1]); gx2 = np.zeros(f.shape[1]); gy1 = np.zeros(f.shape[1]); gy2 = np.zeros(f.shape[1])
                for h_it in range(int(math.floor(ws_h)), int(math.ceil(we_h))):
                    s_, t_ = max(ws_h, float(h_it)) - h_it, min(we_h, h_it + 1.0) - h_it
                    gx1 += _single(s_, t_, _interp(data, h_it, ws_w), _interp(data, h_it + 1, ws_w))
                    gx2 += _single(s_, t_, _interp(data, h_it, we_w), _interp(data, h_it + 1, we_w))
                for w_it in range(int(math.floor(ws_w)), int(math.ceil(we_w))):
                    s_, t_ = max(ws_w, float(w_it)) - w_it, min(we_w, w_it + 1.0) - w_it
                    gy1 += _single(s_, t_, _interp(data, ws_h, w_it), _interp(data, ws_h, w_it + 1))
                    gy2 += _single(s_, t_, _interp(data, we_h, w_it), _interp(data, we_h, w_it + 1))
                top = o[n, :, i, j]
                px1 = (-gx1 + (we_h - ws_h) * top) / win * scale
                py1 = (-gy1 + (we_w - ws_w) * top) / win * scale
                px2 = (gx2 - (we_h - ws_h) * top) / win * scale
                py2 = (gy2 - (we_w - ws_w) * top) / win * scale
                og = np.where(active, og, 0.0)
                rg[n, 1] += np.sum((px1 * (1.0 - j / pw) + px2 * (1.0 - (j + 1) / pw)) * og)
                rg[n, 2] += np.sum((py1 * (1.0 - i / ph) + py2 * (1.0 - (i + 1) / ph)) * og)
                rg[n, 3] += np.sum((px2 * (j + 1) / pw + px1 * j / pw) * og)
                rg[n, 4] += np.sum((py2 * (i + 1) / ph + py1 * i / ph) * og)
    return rg


# ----------------------------------------------------------------------------------------------
# hooks for oracle/ref_shims.py: a CPU stand-in for the reference's `_prroi_pooling` pybind module
# ----------------------------------------------------------------------------------------------
class RefModuleCPU:
    """Same three functions as prroi_pooling_gpu.c:109-113, on CPU tensors."""

    @staticmethod
    def prroi_pooling_forward_cuda(features, rois, ph, pw, scale):
        return torch.from_numpy(forward(features.detach().numpy(), rois.detach().numpy(), ph, pw, scale)).float()

    @staticmethod
    def prroi_pooling_backward_cuda(features, rois, output, output_diff, ph, pw, scale):
        return torch.from_numpy(backward(features.detach().numpy(), rois.detach().numpy(), output_diff.detach().numpy(), ph, pw, scale)).float()

    @staticmethod
    def prroi_pooling_coor_backward_cuda(features, rois, output, output_diff, ph, pw, scale):
        return torch.from_numpy(coor_backward(features.detach().numpy(), rois.detach().numpy(), output.detach().numpy(),
                                              output_diff.detach().numpy(), ph, pw, scale)).float()


def patch_reference_function(prf):
    """PrRoIPool2DFunction.forward asserts CUDA tensors (functional.py:62-63); lift that check for the CPU oracle run."""
    fn = prf.PrRoIPool2DFunction

    def forward_cpu(ctx, features, rois, pooled_height, pooled_width, spatial_scale):
        mod = prf._import_prroi_pooling()          # resolved per call, like the reference's own forward (functional.py:44)
        pooled_height, pooled_width, spatial_scale = int(pooled_height), int(pooled_width), float(spatial_scale)
        features, rois = features.contiguous(), rois.contiguous()
        params = (pooled_height, pooled_width, spatial_scale)
        output = mod.prroi_pooling_forward_cuda(features, rois, *params)
        ctx.params = params
        ctx.save_for_backward(features, rois, output)
        return output
    fn.forward = staticmethod(forward_cpu)
