"""ToMP model-predictor core through the C ABI (`b200trk_transformer_*`): mirrors `Transformer.forward`
(ltr/models/transformer/transformer.py:90-96) for a fixed token count, built once from the module's state_dict."""
import ctypes as C

import torch

from . import _lib


class TransformerEngine:
    def __init__(self, state_dict, L, B, d_model=256, nhead=8, dim_ff=2048, n_enc=6, n_dec=6, prefix="", device=None):
        if not torch.cuda.is_available():
            raise RuntimeError("TransformerEngine: CUDA device required (the engine has no CPU path)")
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.L, self.B, self.D, self.nhead = L, B, d_model, nhead
        keep = []

        def hp(key):
            t = state_dict[prefix + key].detach().float().contiguous().cpu()
            keep.append(t)
            return t.data_ptr()

        def mha(p):
            return _lib.MhaWeights(hp(p + ".in_proj_weight"), hp(p + ".in_proj_bias"), hp(p + ".out_proj.weight"), hp(p + ".out_proj.bias"))

        enc = (_lib.EncLayer * n_enc)()
        for i in range(n_enc):
            p = "encoder.layers.%d" % i
            enc[i] = _lib.EncLayer(mha(p + ".self_attn"), hp(p + ".linear1.weight"), hp(p + ".linear1.bias"), hp(p + ".linear2.weight"),
                                   hp(p + ".linear2.bias"), hp(p + ".norm1.weight"), hp(p + ".norm1.bias"), hp(p + ".norm2.weight"),
                                   hp(p + ".norm2.bias"))
        dec = (_lib.DecLayer * n_dec)()
        for i in range(n_dec):
            p = "decoder.layers.%d" % i
            dec[i] = _lib.DecLayer(mha(p + ".self_attn"), mha(p + ".multihead_attn"), hp(p + ".linear1.weight"), hp(p + ".linear1.bias"),
                                   hp(p + ".linear2.weight"), hp(p + ".linear2.bias"), hp(p + ".norm1.weight"), hp(p + ".norm1.bias"),
                                   hp(p + ".norm2.weight"), hp(p + ".norm2.bias"), hp(p + ".norm3.weight"), hp(p + ".norm3.bias"))
        h = C.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().b200trk_transformer_create(C.byref(h), enc, n_enc, dec, n_dec, C.c_void_p(hp("decoder.norm.weight")),
                                                             C.c_void_p(hp("decoder.norm.bias")), d_model, nhead, dim_ff, L, B),
                       "transformer_create")
        self.handle = h
        del keep

    def forward(self, src, mask, query_embed, pos_embed):
        """Same arguments and return value as the reference: src [L,B,D], mask [B,L] bool or None, query_embed [1,D],
        pos_embed [L,1|B,D] -> (hs [1,B,1,D], memory [L,B,D])."""
        for name, t in (("src", src), ("query_embed", query_embed), ("pos_embed", pos_embed)):
            if not t.is_cuda or t.dtype != torch.float32:
                raise RuntimeError("TransformerEngine.forward: '%s' must be a CUDA float32 tensor" % name)
        if tuple(src.shape) != (self.L, self.B, self.D):
            raise RuntimeError("TransformerEngine.forward: src %s, expected %s" % (tuple(src.shape), (self.L, self.B, self.D)))
        src, pos_embed, query_embed = src.contiguous(), pos_embed.contiguous(), query_embed.contiguous()
        bp = pos_embed.shape[1]
        m8 = None
        if mask is not None:
            m8 = mask.to(device=src.device, dtype=torch.uint8).contiguous()
        hs = torch.empty(self.B, self.D, device=src.device, dtype=torch.float32)
        mem = torch.empty(self.L, self.B, self.D, device=src.device, dtype=torch.float32)
        _lib.check(_lib.lib().b200trk_transformer_forward(
            self.handle, C.c_void_p(src.data_ptr()), C.c_void_p(pos_embed.data_ptr()), int(bp),
            C.c_void_p(m8.data_ptr()) if m8 is not None else None, C.c_void_p(query_embed.data_ptr()), C.c_void_p(hs.data_ptr()),
            C.c_void_p(mem.data_ptr()), C.c_void_p(torch.cuda.current_stream().cuda_stream)), "transformer_forward")
        return hs.reshape(1, self.B, 1, self.D), mem

    def close(self):
        if getattr(self, "handle", None):
            _lib.lib().b200trk_transformer_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class BoxTower:
    """ToMP `DenseBoxRegressor` tower (ltr/models/transformer/heads.py:101-141) through `b200trk_tower_*`: built from the module's
    state_dict (`tower.{0,3,6,9}` conv + `tower.{1,4,7,10}` GroupNorm, `bbreg_layer`)."""

    def __init__(self, state_dict, H, W, prefix="", max_batch=1, precision=0, device=None):
        if not torch.cuda.is_available():
            raise RuntimeError("BoxTower: CUDA device required (the engine has no CPU path)")
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        keep, descs, gam, bet = [], [], [], []

        def hp(key):
            t = state_dict[prefix + key].detach().float().contiguous().cpu()
            keep.append(t)
            return t
        i = 0
        while (prefix + "tower.%d.weight" % (3 * i)) in state_dict:
            w, b = hp("tower.%d.weight" % (3 * i)), hp("tower.%d.bias" % (3 * i))
            d = _lib.ConvDesc()
            d.weight, d.bias, d.cout, d.cin, d.k, d.stride, d.pad = w.data_ptr(), b.data_ptr(), w.shape[0], w.shape[1], 3, 1, 1
            descs.append(d)
            gam.append(hp("tower.%d.weight" % (3 * i + 1)).data_ptr())
            bet.append(hp("tower.%d.bias" % (3 * i + 1)).data_ptr())
            i += 1
        w, b = hp("bbreg_layer.weight"), hp("bbreg_layer.bias")
        d = _lib.ConvDesc()
        d.weight, d.bias, d.cout, d.cin, d.k, d.stride, d.pad = w.data_ptr(), b.data_ptr(), w.shape[0], w.shape[1], 3, 1, 1
        descs.append(d)
        self.C, self.H, self.W, self.max_batch = int(descs[0].cin), H, W, max_batch
        arr = (_lib.ConvDesc * len(descs))(*descs)
        g = (C.c_void_p * len(gam))(*gam)
        bb = (C.c_void_p * len(bet))(*bet)
        h = C.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().b200trk_tower_create(C.byref(h), arr, len(descs), g, bb, self.C, H, W, max_batch, precision), "tower_create")
        self.handle = h
        self.flops = _lib.lib().b200trk_tower_flops(h)
        del keep

    def forward(self, feat, attention=None):
        """feat [S,C,H,W], attention [S,H,W] -> ltrb [S,4,H,W] (exp applied)."""
        if not feat.is_cuda or feat.dtype != torch.float32 or tuple(feat.shape[1:]) != (self.C, self.H, self.W) or feat.shape[0] > self.max_batch:
            raise RuntimeError("BoxTower.forward: expected a CUDA float32 [S<=%d,%d,%d,%d] tensor, got %s" % (self.max_batch, self.C, self.H, self.W, tuple(feat.shape)))
        feat = feat.contiguous()
        att = None
        if attention is not None:
            att = attention.reshape(feat.shape[0], self.H, self.W).contiguous().float()
        out = torch.empty(feat.shape[0], 4, self.H, self.W, device=feat.device, dtype=torch.float32)
        _lib.check(_lib.lib().b200trk_tower_forward(self.handle, C.c_void_p(feat.data_ptr()), C.c_void_p(att.data_ptr()) if att is not None else None,
                                                    feat.shape[0], C.c_void_p(out.data_ptr()), C.c_void_p(torch.cuda.current_stream().cuda_stream)),
                   "tower_forward")
        return out

    def close(self):
        if getattr(self, "handle", None):
            _lib.lib().b200trk_tower_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class TokenBuilder:
    """ToMP token assembly (`b200trk_tomp_tokens`): built from a `FilterPredictor`'s state_dict (`box_encoding.*`, `query_embed_fg`,
    optionally `query_embed_test`); the BatchNorm1d layers of the box-encoding MLP are folded here in double precision."""

    def __init__(self, state_dict, prefix="", device=None):
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        g = lambda k: state_dict[prefix + k].detach().double().cpu()

        def fold(conv, bn):
            w, b = g(conv + ".weight").squeeze(-1), g(conv + ".bias")
            sc = g(bn + ".weight") / torch.sqrt(g(bn + ".running_var") + 1e-5)
            return w * sc[:, None], (b - g(bn + ".running_mean")) * sc + g(bn + ".bias")
        w1, b1 = fold("box_encoding.0", "box_encoding.1")
        w2, b2 = fold("box_encoding.3", "box_encoding.4")
        w3, b3 = g("box_encoding.6.weight").squeeze(-1), g("box_encoding.6.bias")
        dev = lambda t: t.float().contiguous().to(self.device)
        self.w1, self.b1, self.w2t, self.b2, self.w3t, self.b3 = dev(w1), dev(b1), dev(w2.t()), dev(b2), dev(w3.t()), dev(b3)
        self.fg = dev(g("query_embed_fg.weight").reshape(-1))
        self.test = dev(g("query_embed_test.weight").reshape(-1)) if (prefix + "query_embed_test.weight") in state_dict else None
        self.D, self.D1 = int(w3.shape[0]), int(w1.shape[0])

    def build(self, train_feat, test_feat, train_label, train_ltrb, B=2, use_test_token=False):
        """train_feat [n,D,H,W], test_feat [m,D,H,W], train_label [n,H,W], train_ltrb [n,4,H,W] -> tokens [(n+m)*H*W, B, D]."""
        n, D, H, W = train_feat.shape
        m = test_feat.shape[0]
        out = torch.empty((n + m) * H * W, B, D, device=train_feat.device, dtype=torch.float32)
        p = lambda t: C.c_void_p(t.contiguous().data_ptr())
        tf, sf, lb, lt = train_feat.contiguous(), test_feat.contiguous(), train_label.contiguous().float(), train_ltrb.contiguous().float()
        _lib.check(_lib.lib().b200trk_tomp_tokens(
            p(tf), p(sf), p(lb), p(lt), p(self.fg), p(self.test) if (use_test_token and self.test is not None) else None, p(self.w1), p(self.b1),
            p(self.w2t), p(self.b2), p(self.w3t), p(self.b3), p(out), n, m, H, W, D, self.D1, B,
            C.c_void_p(torch.cuda.current_stream().cuda_stream)), "tomp_tokens")
        return out
