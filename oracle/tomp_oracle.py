"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the ToMP model-predictor core.

Transformer.forward (ltr/models/transformer/transformer.py:90-96) with post-norm encoder / decoder layers (:173-181,
:224-238), written with explicit attention arithmetic instead of nn.MultiheadAttention.  Pinned by
oracle/gen_golden.py:gen_transformer, which runs the reference's own `Transformer` module.
"""
import math

import torch
import torch.nn.functional as F


def _mha(sd, p, q_in, k_in, v_in, nhead, key_padding_mask=None):
    """nn.MultiheadAttention forward (eval): q_in [Lq,B,D], k_in/v_in [L,B,D], mask [B,L] bool (True = ignore)."""
    D = q_in.shape[-1]
    w, b = sd[p + ".in_proj_weight"], sd[p + ".in_proj_bias"]
    q = F.linear(q_in, w[:D], b[:D])
    k = F.linear(k_in, w[D:2 * D], b[D:2 * D])
    v = F.linear(v_in, w[2 * D:], b[2 * D:])
    Lq, B, _ = q.shape
    L = k.shape[0]
    hd = D // nhead
    q = q.reshape(Lq, B, nhead, hd).permute(1, 2, 0, 3) / math.sqrt(hd)
    k = k.reshape(L, B, nhead, hd).permute(1, 2, 0, 3)
    v = v.reshape(L, B, nhead, hd).permute(1, 2, 0, 3)
    s = q @ k.transpose(-1, -2)                                   # [B,H,Lq,L]
    if key_padding_mask is not None:
        s = s.masked_fill(key_padding_mask.view(B, 1, 1, L), float("-inf"))
    a = torch.softmax(s, dim=-1)
    o = (a @ v).permute(2, 0, 1, 3).reshape(Lq, B, D)
    return F.linear(o, sd[p + ".out_proj.weight"], sd[p + ".out_proj.bias"])


def _ln(sd, p, x):
    return F.layer_norm(x, (x.shape[-1],), sd[p + ".weight"], sd[p + ".bias"], 1e-5)


def transformer_forward(sd, src, mask, query_embed, pos, nhead, n_enc, n_dec):
    """src [L,B,D], mask [B,L] bool or None, query_embed [1,D], pos [L,1 or B,D] -> (hs [1,B,1,D], memory [L,B,D])."""
    x = src
    for i in range(n_enc):
        p = "encoder.layers.%d" % i
        qk = x + pos
        x = _ln(sd, p + ".norm1", x + _mha(sd, p + ".self_attn", qk, qk, x, nhead, mask))
        ff = F.linear(F.relu(F.linear(x, sd[p + ".linear1.weight"], sd[p + ".linear1.bias"])), sd[p + ".linear2.weight"], sd[p + ".linear2.bias"])
        x = _ln(sd, p + ".norm2", x + ff)
    memory = x
    B = src.shape[1]
    qpos = query_embed.unsqueeze(1).repeat(1, B, 1)
    tgt = torch.zeros_like(qpos)
    for i in range(n_dec):
        p = "decoder.layers.%d" % i
        q = tgt + qpos
        tgt = _ln(sd, p + ".norm1", tgt + _mha(sd, p + ".self_attn", q, q, tgt, nhead))
        tgt = _ln(sd, p + ".norm2", tgt + _mha(sd, p + ".multihead_attn", tgt + qpos, memory + pos, memory, nhead, mask))
        ff = F.linear(F.relu(F.linear(tgt, sd[p + ".linear1.weight"], sd[p + ".linear1.bias"])), sd[p + ".linear2.weight"], sd[p + ".linear2.bias"])
        tgt = _ln(sd, p + ".norm3", tgt + ff)
    hs = _ln(sd, "decoder.norm", tgt).unsqueeze(0)                # [1,1,B,D]
    return hs.transpose(1, 2), memory
