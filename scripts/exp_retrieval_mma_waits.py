"""Where the MMA-issuing thread of the tensor-core retrieval kernel waits (IA_RETR_FLAGS bit 7): for the epilogue to drain an
accumulator (tempty) or for operand stages to land (full); single-CTA vs CTA-pair form, with and without the epilogue scan."""
import os, sys, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import item_alignment_b200 as ia
C, Q, D = [int(v) for v in os.environ.get('WAIT_SHAPE', '1000000,10000,1024').split(',')]
MEASURE = os.environ.get('WAIT_MEASURE', 'cosine')
k = int(sys.argv[1]) if len(sys.argv) > 1 else 10
dev = torch.device("cuda:0")
gen = torch.Generator(device=dev).manual_seed(1)
cat = torch.empty(C, D, dtype=torch.bfloat16, device=dev)
for s in range(0, C, 131072):
    cat[s:s + 131072] = torch.tanh(torch.randn(min(131072, C - s), D, device=dev, generator=gen)).to(torch.bfloat16)
q = torch.tanh(torch.randn(Q, D, device=dev, generator=gen)).to(torch.bfloat16)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
with ia.CatalogIndex(cat) as index:
    for pair in os.environ.get('WAIT_PAIR', '0,1').split(','):
        for flags in [int(x) for x in os.environ.get('WAIT_FLAGS', '142,206').split(',')]:
            os.environ["IA_RETR_PAIR"] = pair
            os.environ["IA_RETR_FLAGS"] = str(flags)
            ts = []
            for it in range(6):
                e0.record(); index.topk_keys(q, k, MEASURE); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
            st = index.last_stats()
            ms = statistics.median(ts[2:])
            tot = max(1, st["cyc_rare"])
            print(f"pair={pair} {'no scan' if flags & 64 else 'scan   '}{' L2 prefetch' if flags & 256 else ''} k={k} {MEASURE} {Q}x{C}x{D}: {ms:7.3f} ms {2.0 * Q * C * D / ms / 1e9:6.0f} TFLOP/s | MMA thread: waits for accumulators "
                  f"{100 * st['rare_groups'] / tot:5.1f} %, for operand stages {100 * st['rare_blocks'] / tot:5.1f} % of {tot / (74 if pair == '1' else 148) / 1e6:.2f} M cycles"
                  f" | epilogue warps wait {100 * st['cyc_wait'] / max(1, st['cyc_total']):.0f} %", flush=True)
