#!/usr/bin/env python
"""Host-side measurement for SURVEY 8f rank 4: embedding JSONL -> matrix.
  native   ia_embedding_jsonl_to_catalog (csrc/catalog_file.cu, one pass, writes the catalog file)
  python   what a consumer of the reference's file does today: json.loads per line + eval of the embedding string
           (model_ensemble.py:112), rows stacked into a float32 matrix
CPU only.  Usage: python scripts/bench_jsonl_converter.py [pairs] [dim]"""
import json
import os
import sys
import tempfile
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import item_alignment_b200 as ia  # noqa: E402
from oracle import torch_port  # noqa: E402

pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
dim = int(sys.argv[2]) if len(sys.argv) > 2 else 768
rng = np.random.default_rng(0)
with tempfile.TemporaryDirectory() as td:
    src = os.path.join(td, "embeds.jsonl")
    with open(src, "w") as w:
        for i in range(pairs):
            a = np.tanh(rng.standard_normal(dim)).astype(np.float32)
            b = np.tanh(rng.standard_normal(dim)).astype(np.float32)
            w.write(torch_port.embedding_jsonl_record(f"s{i:031d}", f"t{i:031d}", a, b, 0.5))
    size = os.path.getsize(src)
    t0 = time.perf_counter()
    rows, d = ia.jsonl_to_catalog(src, os.path.join(td, "c.iacat"), torch.bfloat16, "both", threads=1)
    t_native = time.perf_counter() - t0
    t0 = time.perf_counter()
    ia.jsonl_to_catalog(src, os.path.join(td, "c_mt.iacat"), torch.bfloat16, "both", threads=0)
    t_native_mt = time.perf_counter() - t0
    same_mt = open(os.path.join(td, "c.iacat"), "rb").read() == open(os.path.join(td, "c_mt.iacat"), "rb").read()
    t0 = time.perf_counter()
    ids, mats = [], []
    with open(src) as r:
        for line in r:
            rec = json.loads(line)
            for s in ("src", "tgt"):
                ids.append(rec[f"{s}_item_id"])
                mats.append(np.asarray(eval(rec[f"{s}_item_emb"]), dtype=np.float32))   # the reference's own way of reading it
    mat = np.stack(mats)
    t_py = time.perf_counter() - t0
    with ia.CatalogFile(os.path.join(td, "c.iacat")) as f:
        same = np.array_equal(f.numpy(), torch.from_numpy(mat).to(torch.bfloat16).view(torch.int16).numpy().view(np.uint16))
    print(json.dumps({"file_mb": size / 1e6, "rows": rows, "dim": d, "native_s": t_native, "native_mb_per_s": size / 1e6 / t_native,
                      "native_all_cores_s": t_native_mt, "native_all_cores_mb_per_s": size / 1e6 / t_native_mt, "cores": os.cpu_count(),
                      "all_cores_output_identical": bool(same_mt),
                      "python_eval_s": t_py, "python_mb_per_s": size / 1e6 / t_py, "speedup": t_py / t_native, "identical": bool(same)}))
