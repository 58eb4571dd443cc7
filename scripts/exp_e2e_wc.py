"""e2e host pipeline at N ranks (torchrun): regular pinned inputs vs write-combined pinned inputs."""
import ctypes, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import bench
from item_alignment_b200 import _lib
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
lib = _lib.lib()
x, y, labels = bench.make_pairs(torch, dev, 1 + rank, torch.bfloat16)
nbytes = x.numel() * 2


def wc_copy(t):
    p = ctypes.c_void_p()
    _lib.check(lib.ia_host_alloc(ctypes.byref(p), t.numel() * t.element_size(), 1))
    buf = (ctypes.c_char * (t.numel() * t.element_size())).from_address(p.value)
    out = torch.frombuffer(buf, dtype=t.dtype).view(t.shape)
    out.copy_(t.cpu())
    return out, p


for name in ("regular pinned inputs", "write-combined inputs"):
    if name.startswith("regular"):
        xh, yh, lh = x.cpu().pin_memory(), y.cpu().pin_memory(), labels.cpu().pin_memory()
    else:
        (xh, px), (yh, py), (lh, pl) = wc_copy(x), wc_copy(y), wc_copy(labels)
    dxh, dyh = torch.empty(x.shape, dtype=x.dtype).pin_memory(), torch.empty(x.shape, dtype=x.dtype).pin_memory()
    loss = torch.zeros(1).pin_memory()

    def step():
        _lib.check(lib.ia_pair_score_loss_host(0, 0, 1.0, 1, _lib.IA_BF16, xh.data_ptr(), yh.data_ptr(), lh.data_ptr(), bench.N_PAIRS, bench.DIM,
                                               loss.data_ptr(), dxh.data_ptr(), dyh.data_ptr(), local))
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(15):
        step()
    ms = torch.tensor([(time.perf_counter() - t0) / 15 * 1e3], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(f"N={world} {name:24s}: {float(ms):.2f} ms/step  {world * bench.N_PAIRS / float(ms) / 1e3:.2f} Mpairs/s  {nbytes * 2 / float(ms) / 1e6:.1f} GB/s each way per rank", flush=True)
    for _ in range(3):
        step()                 # fwd-only variant below uses the same buffers
    t0 = time.perf_counter()
    for _ in range(15):
        _lib.check(lib.ia_pair_score_loss_host(0, 0, 1.0, 1, _lib.IA_BF16, xh.data_ptr(), yh.data_ptr(), lh.data_ptr(), bench.N_PAIRS, bench.DIM,
                                               loss.data_ptr(), None, None, local))
    ms = torch.tensor([(time.perf_counter() - t0) / 15 * 1e3], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(f"N={world} {name:24s}: upload only (no gradients back) {float(ms):.2f} ms/step  {nbytes * 2 / float(ms) / 1e6:.1f} GB/s H2D per rank", flush=True)
if world > 1:
    dist.destroy_process_group()
