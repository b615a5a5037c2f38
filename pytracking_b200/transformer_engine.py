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
