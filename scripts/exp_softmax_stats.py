"""Cycle counters of the tensor-core softmax head (run with IA_HEAD_DEBUG=16 [+8: no stores])."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from item_alignment_b200 import functional as F_, _lib
dev = "cuda:0"
n = 65536
for dt, h in ((torch.bfloat16, 1024), (torch.bfloat16, 768), (torch.float16, 512)):
    x = torch.tanh(torch.randn(n, h, device=dev)).to(dt); y = torch.tanh(torch.randn(n, h, device=dev)).to(dt)
    l = (torch.rand(n, device=dev) < 0.5).long(); w = torch.randn(2, 2 * h, device=dev) * 0.02; b = torch.zeros(2, device=dev)
    for _ in range(3):
        F_.softmax_head_raw(x, y, w, b, l)
    torch.cuda.synchronize()
    out = (ctypes.c_uint64 * 8)()
    _lib.check(_lib.lib().ia_softmax_head_last_stats(out))
    fw, bw, ctas = 4, 8, 148
    tot_f, tot_b = out[3] / (fw * ctas), out[5] / (bw * ctas)
    print(f"h={h}: fwd warp {tot_f:9.0f} cyc: wait rows {out[0] / (fw * ctas) / tot_f * 100:5.1f} %  barrier {out[1] / (fw * ctas) / tot_f * 100:5.1f} %  "
          f"bwd wait for release {out[2] / (bw * ctas) / tot_b * 100:5.1f} %  warp0 softmax {out[6] / ctas / tot_f * 100:5.1f} %   |  "
          f"bwd warp {tot_b:9.0f} cyc: wait deltas {out[4] / (bw * ctas) / tot_b * 100:5.1f} %")
