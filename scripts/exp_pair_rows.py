"""Row-grouping experiment for the fused pair kernel (IA_PAIR_ROWS=1|2): outputs alternate between two buffer sets."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from item_alignment_b200 import functional as F_
dev = torch.device("cuda:0")
n, d = 65536, 1024
g = torch.Generator(device=dev).manual_seed(2)
for dt, d in ((torch.bfloat16, 1024), (torch.float32, 1024), (torch.bfloat16, 512), (torch.float32, 768)):
    x = torch.tanh(torch.randn(n, d, device=dev, generator=g)).to(dt); y = torch.tanh(torch.randn(n, d, device=dev, generator=g)).to(dt)
    l = (torch.rand(n, device=dev, generator=g) < 0.5).long()
    for _ in range(20): out = F_.pair_score_loss_raw("inner_product", "bce", x, y, l)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(300): out = F_.pair_score_loss_raw("inner_product", "bce", x, y, l)     # keeps the previous outputs alive: alternating buffers
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 300 * 1e3
    b = n * (4 * d * x.element_size() + 16)
    print(f"rows={os.environ.get('IA_PAIR_ROWS','default')} {str(dt)[6:]:8s} d={d}: {us:6.1f} us  {b/us/1e3:6.0f} GB/s  {b/us/1e3/6554.9*100:5.1f}%")
