"""Small invocations of every kernel family, for compute-sanitizer (memcheck / racecheck / synccheck)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import item_alignment_b200 as ia
from item_alignment_b200 import functional as F_
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
for dt, n, d in ((torch.bfloat16, 300, 1024), (torch.float32, 257, 768), (torch.float32, 33, 50), (torch.float16, 64, 2048)):
    x = torch.tanh(torch.randn(n, d, device=dev, generator=g)).to(dt); y = torch.tanh(torch.randn(n, d, device=dev, generator=g)).to(dt)
    l = (torch.rand(n, device=dev, generator=g) < 0.5).long()
    for m in ("inner_product", "cosine", "l1", "l2"):
        F_.pair_score_raw(m, x, y, threshold=0.5)
        F_.pair_score_loss_raw(m, "hinge", x, y, l)
        F_.pair_score_loss_raw(m, "cosine", x, y, l, grad_dtype=torch.float32)
        gs = torch.randn(n, device=dev); F_.pair_score_bwd_raw(m, x, y, gs)
    if d % 8 == 0:
        w = torch.randn(2, 2 * d, device=dev) * 0.02; b = torch.zeros(2, device=dev)
        F_.softmax_head_raw(x, y, w, b, l); F_.softmax_head_raw(x, y, w, b)
os.environ.pop("IA_PAIR_BULK", None)
n, d = 20000, 1024
x = torch.tanh(torch.randn(n, d, device=dev, generator=g)).to(torch.bfloat16); y = torch.tanh(torch.randn(n, d, device=dev, generator=g)).to(torch.bfloat16)
l = (torch.rand(n, device=dev, generator=g) < 0.5).long()
F_.pair_score_loss_raw("inner_product", "bce", x, y, l)        # ROWS=2 path
w = torch.randn(2, 2 * d, device=dev) * 0.02; b = torch.zeros(2, device=dev)
F_.softmax_head_raw(x, y, w, b, l)                             # grouped ring path
cat = torch.tanh(torch.randn(3000, 128, device=dev, generator=g)).to(torch.bfloat16)
q = torch.tanh(torch.randn(130, 128, device=dev, generator=g)).to(torch.bfloat16)
with ia.CatalogIndex(cat) as idx:
    for k in (100, 10, 1):
        for m in ("cosine", "inner_product", "l1", "l2"):
            idx.topk(q, k, m)
    keys = idx.topk_keys(q, 10, "cosine")
    bound = (keys[:, 9] >> 32) & 0xFFFFFFFF
    idx.topk_keys(q, 100, "cosine", init_tau=bound)
with ia.CatalogIndex(cat.float()) as idx:
    idx.topk(q.float(), 20, "cosine")
ia.merge_keys(torch.stack([keys, keys]), 10)
emb = cat; F_.pair_score_gather_raw("cosine", emb, emb, torch.arange(100, device=dev), torch.arange(100, 200, device=dev))
ia.threshold_sweep(torch.rand(5000, device=dev), (torch.rand(5000, device=dev) < 0.5).long())
# head projection (tcgen05 GEMM + tanh epilogue, with and without the fused score; partial tiles in every dimension)
for dt, n, k, h in ((torch.bfloat16, 300, 264, 136), (torch.float16, 129, 768, 768), (torch.bfloat16, 21, 48, 48)):
    f1 = torch.randn(n, k, device=dev, generator=g).to(dt); f2 = torch.randn(n, k, device=dev, generator=g).to(dt)
    w = (torch.randn(h, k, device=dev, generator=g) / k ** 0.5).to(dt); b = torch.randn(h, device=dev, generator=g) * 0.1
    F_.project_tanh_raw(f1, f2, w, b)
    for m in ("inner_product", "cosine", "l1", "l2"):
        F_.project_score_raw(m, f1, f2, w, b, threshold=0.5, want_embeds=(m == "cosine"))
# dissimilarity knobs + catalog file upload
with ia.CatalogIndex(cat.float()) as idx:
    idx.topk_dissimilarity(q.float(), 10, p=1); idx.topk_dissimilarity(q.float(), 10, p=2)
import tempfile
with tempfile.TemporaryDirectory() as td:
    ia.write_catalog(td + "/c.iacat", cat.cpu(), [str(i) for i in range(cat.shape[0])])
    with ia.CatalogFile(td + "/c.iacat") as f, f.index(100, 2900) as idx:
        idx.topk(q, 5, "cosine")
# round 2: tensor-core softmax head (n >= 4096, h = 512), best-F1 search, probe pass + self-probed top-k, gradient recomputation
n, d = 4100, 512
x = torch.tanh(torch.randn(n, d, device=dev, generator=g)).to(torch.float16); y = torch.tanh(torch.randn(n, d, device=dev, generator=g)).to(torch.float16)
l = (torch.rand(n, device=dev, generator=g) < 0.5).long()
w = torch.randn(2, 2 * d, device=dev) * 0.02; b = torch.zeros(2, device=dev)
F_.softmax_head_raw(x, y, w, b, l)
xg = x.clone().requires_grad_(True)
_, _, loss = F_.pair_score_loss("cosine", "bce", xg, y, l); (loss * 8.0).backward()
ia.find_best_f1_and_threshold(torch.rand(20000, device=dev), (torch.rand(20000, device=dev) < 0.3).long())
ia.find_best_f1_and_threshold(torch.rand(9000, device=dev, dtype=torch.float64), (torch.rand(9000, device=dev) < 0.3).long(), False)
big = torch.tanh(torch.randn(40000, 128, device=dev, generator=g)).to(torch.bfloat16)
with ia.CatalogIndex(big) as idx:
    idx.topk_keys(q, 100, "cosine")                 # runs its own probe pass
    idx.probe_bound(q, 13, 2, 2048, "inner_product")
torch.cuda.synchronize()
print("sanitize_small: done")
