"""e2e host pipeline (ia_pair_score_loss_host with gradients returned) for the chunk size given in IA_HOST_CHUNK_MB."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from item_alignment_b200 import _lib
lib = _lib.lib()
dev = torch.device("cuda:0")
x, y, labels = bench.make_pairs(torch, dev, 1, torch.bfloat16)
xh, yh, lh = x.cpu().pin_memory(), y.cpu().pin_memory(), labels.cpu().pin_memory()
dxh, dyh = torch.empty_like(xh).pin_memory(), torch.empty_like(yh).pin_memory()
loss = torch.zeros(1).pin_memory()
def step():
    _lib.check(lib.ia_pair_score_loss_host(0, 0, 1.0, 1, _lib.IA_BF16, xh.data_ptr(), yh.data_ptr(), lh.data_ptr(), bench.N_PAIRS, bench.DIM,
                                           loss.data_ptr(), dxh.data_ptr(), dyh.data_ptr(), 0))
for _ in range(5): step()
t0 = time.perf_counter()
for _ in range(40): step()
ms = (time.perf_counter() - t0) / 40 * 1e3
print(f"IA_HOST_CHUNK_MB={os.environ.get('IA_HOST_CHUNK_MB', 'default(16)')}: {ms:.3f} ms/step  {bench.N_PAIRS / ms / 1e3:.2f} Mpairs/s  "
      f"{2 * bench.N_PAIRS * bench.DIM * 2 / ms / 1e6:.1f} GB/s each way")
