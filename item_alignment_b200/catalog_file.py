"""Binary catalog files: the [C, D] embedding matrix + item ids behind one mmap.

Replaces, as the input of retrieval, the embedding JSONL the reference writes (finetune_text.py:784-792) and parses
back with `eval` (model_ensemble.py:112).  The file work is native (csrc/catalog_file.cu: writer, mmap reader,
double-buffered upload, JSONL converter); this module is the ctypes mirror.

    jsonl_to_catalog("embeds.jsonl", "catalog.iacat")          # once
    with CatalogFile("catalog.iacat") as f:
        index = f.index()                                       # CatalogIndex on the current CUDA device
        scores, rows = index.topk(queries, 100, "cosine")
        ids = [f.id(r) for r in rows[0].tolist()]
"""
import ctypes
import os

import numpy as np
import torch

from . import _lib
from ._lib import check, lib

_SIDES = {"src": 0, "tgt": 1, "both": 2}
_TORCH_DT = {_lib.IA_F32: torch.float32, _lib.IA_BF16: torch.bfloat16, _lib.IA_F16: torch.float16}
_IA_DT = {v: k for k, v in _TORCH_DT.items()}


def write_catalog(path, matrix, ids=None):
    """matrix: [C, D] CPU tensor (fp32 / bf16 / fp16) or float32 numpy array; ids: optional list of C strings."""
    if isinstance(matrix, np.ndarray):
        matrix = torch.from_numpy(np.ascontiguousarray(matrix))
    if matrix.dim() != 2 or matrix.dtype not in _IA_DT:
        raise ValueError("matrix must be a 2-D fp32 / bf16 / fp16 tensor")
    m = matrix.detach().cpu().contiguous()
    rows, dim = m.shape
    arr = None
    if ids is not None:
        if len(ids) != rows:
            raise ValueError("ids must have one entry per row")
        arr = (ctypes.c_char_p * rows)(*[str(i).encode("utf-8") for i in ids])
    check(lib().ia_catalog_file_write(os.fsencode(path), _IA_DT[m.dtype], m.data_ptr(), rows, dim, arr))


def jsonl_to_catalog(jsonl_path, out_path, dtype=torch.bfloat16, side="both", threads=0):
    """Reference embedding JSONL -> catalog file (native parser, `threads` workers, 0 = all cores).  Returns (rows, dim)."""
    if side not in _SIDES:
        raise ValueError("side must be 'src', 'tgt' or 'both'")
    rows, dim = ctypes.c_int64(0), ctypes.c_int64(0)
    check(lib().ia_embedding_jsonl_to_catalog(os.fsencode(jsonl_path), os.fsencode(out_path), _IA_DT[dtype], _SIDES[side],
                                              int(threads), ctypes.byref(rows), ctypes.byref(dim)))
    return rows.value, dim.value


class CatalogFile:
    """mmap view of a catalog file."""

    def __init__(self, path):
        h = ctypes.c_void_p()
        check(lib().ia_catalog_file_open(os.fsencode(path), ctypes.byref(h)))
        self._h = h
        dt, rows, dim, has_ids = ctypes.c_int(0), ctypes.c_int64(0), ctypes.c_int64(0), ctypes.c_int(0)
        check(lib().ia_catalog_file_info(self._h, ctypes.byref(dt), ctypes.byref(rows), ctypes.byref(dim), ctypes.byref(has_ids)))
        self.dtype, self.rows, self.dim, self.has_ids = _TORCH_DT[dt.value], rows.value, dim.value, bool(has_ids.value)

    def close(self):
        """Unmaps the file; arrays returned by numpy() must not be used afterwards."""
        if getattr(self, "_h", None):
            lib().ia_catalog_file_close(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:      # interpreter shutdown: the binding may already be gone
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def id(self, row):
        n = ctypes.c_int64(0)
        p = lib().ia_catalog_file_id(self._h, int(row), ctypes.byref(n))
        if not p:
            raise IndexError(f"no id for row {row}")
        return ctypes.string_at(p, n.value).decode("utf-8")

    def numpy(self):
        """The mapped matrix as a read-only numpy view (uint16 bit patterns for bf16 / fp16, float32 otherwise)."""
        e = 4 if self.dtype == torch.float32 else 2
        buf = (ctypes.c_uint8 * (self.rows * self.dim * e)).from_address(lib().ia_catalog_file_data(self._h))
        a = np.frombuffer(buf, dtype=np.float32 if e == 4 else np.uint16).reshape(self.rows, self.dim)
        a.flags.writeable = False
        return a

    def to_device(self, row_begin=0, row_end=None, device=None):
        """Rows [row_begin, row_end) as a CUDA tensor (native double-buffered upload)."""
        row_end = self.rows if row_end is None else row_end
        device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        if device.type != "cuda":
            raise RuntimeError("item_alignment_b200 runs on CUDA tensors only (no CPU fallback)")
        out = torch.empty((row_end - row_begin, self.dim), dtype=self.dtype, device=device)
        with torch.cuda.device(device):
            check(lib().ia_catalog_file_upload(self._h, row_begin, row_end, out.data_ptr(), torch.cuda.current_stream().cuda_stream))
        return out

    def index(self, row_begin=0, row_end=None, device=None):
        """CatalogIndex over rows [row_begin, row_end); returned rows are GLOBAL file rows."""
        from .retrieval import CatalogIndex
        return CatalogIndex(self.to_device(row_begin, row_end, device), row_base=row_begin)

    def sharded_index(self, group=None, device=None):
        """This rank's shard_bounds() slice as a ShardedCatalogIndex (one process per GPU)."""
        import torch.distributed as dist
        from .retrieval import ShardedCatalogIndex, shard_bounds
        lo, hi = shard_bounds(self.rows, dist.get_world_size(group), dist.get_rank(group))
        return ShardedCatalogIndex(self.to_device(lo, hi, device), total_rows=self.rows, group=group)
