"""Layer-by-layer comparison of the tensor-core (precision 0) plan against the fp32 CUDA-core (precision 1) plan."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pytracking_b200 import synth, _lib
from pytracking_b200.engine import BackboneEngine

arch = sys.argv[1] if len(sys.argv) > 1 else "resnet50"
size = int(sys.argv[2]) if len(sys.argv) > 2 else 288
S = int(sys.argv[3]) if len(sys.argv) > 3 else 1
sd = synth.make_dimp_state_dict(arch, seed=0, lut_seed=3)
im = synth.make_crop(41, S, size).cuda()
L = _lib.lib()
engs = [BackboneEngine(sd, arch=arch, max_batch=S, crop_size=size, precision=p) for p in (1, 0)]
outs = [e.forward(im) for e in engs]
torch.cuda.synchronize()
n = L.b200trk_net_num_ops(engs[0].handle)
worst = 0.0
for i in range(n):
    info = (C.c_int * 8)()
    _lib.check(L.b200trk_net_op_info(engs[1].handle, i, C.byref(info)))
    kind, cin, cout, k, stride, ho, wo, tc = list(info)
    if kind not in (1, 2, 3):
        continue
    bufs = []
    for e in engs:
        t = torch.empty(S, ho, wo, cout, device="cuda")
        _lib.check(L.b200trk_net_op_output(e.handle, i, S, C.c_void_p(t.data_ptr()), None))
        bufs.append(t)
    torch.cuda.synchronize()
    a, b = bufs
    err = float((a.double() - b.double()).abs().max() / (a.double().abs().max() + 1e-30))
    worst = max(worst, err)
    flag = "" if err < 1e-4 else "   <<<<<< MISMATCH"
    print("op %2d kind %d cin %4d cout %4d k %d s %d out %3dx%-3d tc %d  rel err %.3e  nan %d%s" % (
        i, kind, cin, cout, k, stride, ho, wo, tc, err, int(torch.isnan(b).any()), flag))
    if err >= 1e-4 and "-v" in sys.argv:
        d = (a - b).abs()
        idx = torch.nonzero(d > 1e-4 * a.abs().max())
        print("   first mismatches (s,y,x,c):", idx[:8].tolist(), " count", idx.shape[0], "of", d.numel())
for name in outs[0]:
    a, b = outs[0][name], outs[1][name]
    print("%-16s rel err %.3e" % (name, float((a.double() - b.double()).abs().max() / a.double().abs().max())))
print("WORST", worst)
