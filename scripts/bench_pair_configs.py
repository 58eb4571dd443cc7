"""BASELINE configs 1-3 on one GPU (device-resident inputs, CUDA events): pairs/s and achieved HBM GB/s."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from item_alignment_b200 import functional as F_

dev = torch.device("cuda:0")
peak = 6554.9
try:
    peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass


def timeit(fn, iters=200, warm=20):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3   # us


def make(n, d, dt, seed):
    g = torch.Generator(device=dev).manual_seed(seed)
    x = torch.tanh(torch.randn(n, d, device=dev, generator=g)).to(dt)
    y = torch.tanh(torch.randn(n, d, device=dev, generator=g)).to(dt)
    labels = (torch.rand(n, device=dev, generator=g) < 0.5).long()
    return x, y, labels

print(f"IA_PAIR_BULK={os.environ.get('IA_PAIR_BULK', 'default')}  peak {peak} GB/s")
x, y, l = make(50000, 768, torch.float32, 1)
us = timeit(lambda: F_.pair_score_raw("cosine", x, y, threshold=0.5))
b = 50000 * (2 * 768 * 4 + 9)
print(f"C1 cosine score+threshold fwd 50000x768 fp32: {us:7.1f} us  {50000/us:8.1f} Mpairs/s  {b/us/1e3:7.0f} GB/s  {b/us/1e3/peak*100:5.1f}%")
x, y, l = make(65536, 1024, torch.bfloat16, 2)
us = timeit(lambda: F_.pair_score_loss_raw("inner_product", "bce", x, y, l))
b = 65536 * (4 * 1024 * 2 + 16)
print(f"C2 inner+bce fwd/bwd 65536x1024 bf16:        {us:7.1f} us  {65536/us:8.1f} Mpairs/s  {b/us/1e3:7.0f} GB/s  {b/us/1e3/peak*100:5.1f}%")
x, y, l = make(65536, 1024, torch.float32, 3)
for m in ("l1", "l2"):
    for lo in ("hinge", "euclidean"):
        us = timeit(lambda: F_.pair_score_loss_raw(m, lo, x, y, l))
        b = 65536 * (4 * 1024 * 4 + 16)
        print(f"C3 {m}+{lo:9s} fwd/bwd 65536x1024 fp32:   {us:7.1f} us  {65536/us:8.1f} Mpairs/s  {b/us/1e3:7.0f} GB/s  {b/us/1e3/peak*100:5.1f}%")
for m, lo, dt in (("cosine", "cosine", torch.bfloat16), ("cosine", "bce", torch.float32)):
    x, y, l = make(65536, 1024, dt, 4)
    us = timeit(lambda: F_.pair_score_loss_raw(m, lo, x, y, l))
    b = 65536 * (4 * 1024 * x.element_size() + 16)
    print(f"   {m}+{lo:9s} fwd/bwd 65536x1024 {str(dt)[6:]:8s}: {us:7.1f} us  {65536/us:8.1f} Mpairs/s  {b/us/1e3:7.0f} GB/s  {b/us/1e3/peak*100:5.1f}%")
# softmax head + CE (config "softmax"/ce)
x, y, l = make(65536, 1024, torch.bfloat16, 5)
w = (torch.randn(2, 2048, device=dev) * 0.02); bb = torch.zeros(2, device=dev)
us = timeit(lambda: F_.softmax_head_raw(x, y, w, bb, l))
b = 65536 * (4 * 1024 * 2 + 24)
print(f"   softmax head + ce fwd/bwd 65536x1024 bf16:  {us:7.1f} us  {65536/us:8.1f} Mpairs/s  {b/us/1e3:7.0f} GB/s  {b/us/1e3/peak*100:5.1f}%")
