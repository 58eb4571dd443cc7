import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from item_alignment_b200 import functional as F_
dev = torch.device("cuda:0")
n, d = 65536, 1024
dt = torch.bfloat16 if (len(sys.argv) < 2 or sys.argv[1] == "bf16") else torch.float32
x = torch.tanh(torch.randn(n, d, device=dev)).to(dt); y = torch.tanh(torch.randn(n, d, device=dev)).to(dt)
l = (torch.rand(n, device=dev) < 0.5).long()
w = torch.randn(2, 2 * d, device=dev) * 0.02; b = torch.zeros(2, device=dev)
for _ in range(5):
    F_.softmax_head_raw(x, y, w, b, l)
    F_.softmax_head_raw(x, y, w, b)
torch.cuda.synchronize()
