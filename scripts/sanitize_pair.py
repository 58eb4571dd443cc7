"""Small invocations of the CTA-pair (cta_group::2) form of the tensor-core retrieval kernel, for compute-sanitizer."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["IA_RETR_PAIR"] = "1"
import torch
import item_alignment_b200 as ia
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
for dt, qn, cn, d in ((torch.bfloat16, 300, 9000, 128), (torch.float16, 256, 4100, 64)):
    cat = torch.tanh(torch.randn(cn, d, device=dev, generator=g)).to(dt)
    q = torch.tanh(torch.randn(qn, d, device=dev, generator=g)).to(dt)
    with ia.CatalogIndex(cat) as idx:
        for k in (100, 10):
            for m in ("cosine", "inner_product"):
                keys = idx.topk_keys(q, k, m)
                os.environ["IA_RETR_PAIR"] = "0"
                ref = idx.topk_keys(q, k, m)
                os.environ["IA_RETR_PAIR"] = "1"
                assert torch.equal(keys, ref), (dt, k, m)
torch.cuda.synchronize()
print("pair kernel: small invocations done, keys equal to the single-CTA form")
