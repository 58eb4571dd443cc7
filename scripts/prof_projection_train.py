"""One training step of the head on the fused kernels (for `ncu --metrics gpu__time_duration.sum` launch lists and timing)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from item_alignment_b200 import functional as F_
dev = torch.device("cuda:0")
n, k, h = 65536, 1024, 1024
g = torch.Generator(device=dev).manual_seed(1)
f1 = torch.randn(n, k, device=dev, generator=g).to(torch.bfloat16).requires_grad_(True)
f2 = torch.randn(n, k, device=dev, generator=g).to(torch.bfloat16).requires_grad_(True)
w = (torch.randn(h, k, device=dev, generator=g) / k ** 0.5).requires_grad_(True)
b = (torch.randn(h, device=dev, generator=g) * 0.1).requires_grad_(True)
labels = (torch.rand(n, device=dev, generator=g) < 0.5).long()
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for it in range(reps):
    if it == reps - 1:
        torch.cuda.synchronize(); e0.record()
    _, _, _, _, loss = F_.project_score_loss_train("cosine", "bce", f1, f2, w, b, labels, 0.1, 7, it + 1)
    loss.backward()
    f1.grad = f2.grad = w.grad = b.grad = None
e1.record(); torch.cuda.synchronize()
print(f"train step {e0.elapsed_time(e1):.3f} ms, loss {float(loss):.5f}")
