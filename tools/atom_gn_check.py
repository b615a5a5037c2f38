"""Where does b200trk_atom_gn_joint leave the oracle? (fp32 oracle vs fp64 oracle differ by 7e-7 at BASELINE size, so a 2e-3 gap to the
GPU result is not conditioning.)  Sweeps (num_cg, num_gn) and sample counts."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pytracking_b200 import ops, synth                      # noqa: E402
from oracle import atom_oracle as A                         # noqa: E402

torch.set_num_threads(16)
rel = lambda a, b: float((a.double().cpu() - b.double().cpu()).abs().max() / b.double().cpu().abs().max())
for n, cin in ((30, 256), (4, 64), (30, 64), (4, 256)):
    x, y, sw = synth.make_atom_memory(111, n, cin, 18, 18)
    g = torch.Generator().manual_seed(112)
    P0 = torch.randn(64, cin, 1, 1, generator=g) * (1.0 / 16)
    w0 = torch.zeros(1, 64, 4, 4)
    for ncg, ngn in ((1, 1), (2, 1), (5, 1), (10, 1), (1, 2), (10, 2), (10, 6)):
        for fr in (True, False):
            w, P = w0.clone().cuda(), P0.clone().cuda()
            ops.atom_gn_joint_(w, P, x.cuda(), y.cuda(), sw.cuda(), 0.1, 1e-4, ncg, ngn, "mlu", 0.05, fr)
            w_ref, P_ref = A.atom_gn_joint(w0.double(), P0.double(), x.double(), y.double(), sw.double(), 0.1, 1e-4, ncg, ngn, "mlu", 0.05, fr)
            print("n=%2d cin=%3d cg=%2d gn=%d fr=%d : w %.2e  P %.2e" % (n, cin, ncg, ngn, fr, rel(w, w_ref), rel(P, P_ref)), flush=True)
