"""CPU tests of the rank-4 "next" row: the binary catalog file, the native converter from the reference's
embedding JSONL (finetune_text.py:784-792 / model_ensemble.py:112) and the oracle restatement of torchkge's candidate
ranking against vectors produced by the reference's vendored torchkge (tests/golden/kg_golden.npz).  Host-side functions
only: nothing here needs a GPU."""
import os

import numpy as np
import pytest
import torch

from oracle import ref_import, torch_port


def test_oracle_kg_ranking_matches_reference_torchkge(golden):
    g = golden("kg_golden")
    for kind in ("L1", "L2"):
        ent, rel = torch.from_numpy(g[f"{kind}/ent_emb"]), torch.from_numpy(g[f"{kind}/rel_emb"])
        ents, rels = torch.from_numpy(g[f"{kind}/ents"]), torch.from_numpy(g[f"{kind}/rels"])
        for missing in ("tails", "heads"):
            idx, top, scores = torch_port.kg_rank_entities(ent, rel, ents, rels, 10, missing, kind)
            assert np.array_equal(scores.numpy(), g[f"{kind}/{missing}/scores"])
            assert np.array_equal(idx.numpy(), g[f"{kind}/{missing}/top_idx"])
            assert np.array_equal(top.numpy(), g[f"{kind}/{missing}/top_scores"])


@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16, torch.float16])
def test_catalog_file_round_trip(tmp_path, dt):
    import item_alignment_b200 as ia
    gen = torch.Generator().manual_seed(1)
    m = torch.randn(37, 24, generator=gen).to(dt)
    ids = [f"item-{i}" for i in range(36)] + ["ü-ñ-商品"]
    path = tmp_path / "c.iacat"
    ia.write_catalog(path, m, ids)
    with ia.CatalogFile(path) as f:
        assert (f.rows, f.dim, f.dtype, f.has_ids) == (37, 24, dt, True)
        assert [f.id(i) for i in range(37)] == ids
        a = f.numpy()
        want = m.numpy() if dt == torch.float32 else m.view(torch.int16).numpy().view(np.uint16)
        assert np.array_equal(a, want)
        with pytest.raises(IndexError):
            f.id(37)
    ia.write_catalog(path, m)                      # no id table
    with ia.CatalogFile(path) as f:
        assert not f.has_ids and f.rows == 37
        with pytest.raises(IndexError):
            f.id(0)
    assert os.path.getsize(path) == 4096 + 37 * 24 * m.element_size()


def test_catalog_file_rejects_garbage(tmp_path):
    import item_alignment_b200 as ia
    bad = tmp_path / "bad.iacat"
    bad.write_bytes(b"not a catalog" * 10)
    with pytest.raises(ValueError):
        ia.CatalogFile(bad)
    with pytest.raises(ValueError):
        ia.CatalogFile(tmp_path / "missing.iacat")
    good = tmp_path / "good.iacat"
    ia.write_catalog(good, torch.zeros(4, 8))
    raw = bytearray(good.read_bytes())
    raw[16:24] = (10 ** 9).to_bytes(8, "little")    # rows beyond the end of the file
    good.write_bytes(bytes(raw))
    with pytest.raises(ValueError):
        ia.CatalogFile(good)


def test_catalog_file_rejects_crafted_headers(tmp_path):
    """header fields come from the file -- sizes that wrap around 2^64 and id-offset tables that run backwards or
    past the blob must be rejected at open, not crash a later read."""
    import item_alignment_b200 as ia
    good = tmp_path / "good.iacat"
    ia.write_catalog(good, torch.arange(32, dtype=torch.float32).view(4, 8), ids=["a", "bb", "ccc", "dddd"])
    raw0 = good.read_bytes()
    with ia.CatalogFile(good) as f:
        assert [f.id(i) for i in range(4)] == ["a", "bb", "ccc", "dddd"]
    ids_offset = int.from_bytes(raw0[40:48], "little")

    def reject(mutate):
        raw = bytearray(raw0)
        mutate(raw)
        path = tmp_path / "crafted.iacat"
        path.write_bytes(bytes(raw))
        with pytest.raises(ValueError):
            ia.CatalogFile(path)

    # rows * dim * 4 wraps to a small number: 2^61 rows x 8 columns x 4 bytes == 0 (mod 2^64)
    reject(lambda r: r.__setitem__(slice(16, 24), (1 << 61).to_bytes(8, "little")))
    # dim huge, rows * dim wraps
    reject(lambda r: r.__setitem__(slice(24, 32), ((1 << 62) + 2).to_bytes(8, "little")))
    # data_offset + data_bytes wraps
    reject(lambda r: r.__setitem__(slice(32, 40), ((1 << 64) - 16).to_bytes(8, "little")))
    # ids_offset + ids_bytes wraps
    reject(lambda r: r.__setitem__(slice(48, 56), ((1 << 64) - 8).to_bytes(8, "little")))
    # id offsets not monotone: off[1] beyond off[2] (a read of id 0 would run past the blob, id 1 gets a negative length)
    reject(lambda r: r.__setitem__(slice(ids_offset + 8, ids_offset + 16), (1 << 40).to_bytes(8, "little")))
    reject(lambda r: r.__setitem__(slice(ids_offset + 16, ids_offset + 24), (0).to_bytes(8, "little")))


def _write_reference_jsonl(path, n_pairs, dim, seed):
    rng = np.random.default_rng(seed)
    items = {f"{i:032x}": np.tanh(rng.standard_normal(dim)).astype(np.float32) * np.float32(10.0 ** rng.integers(-6, 3))
             for i in range(n_pairs)}
    keys = list(items)
    with open(path, "w") as w:
        for i in range(n_pairs):
            s, t = keys[rng.integers(0, len(keys))], keys[rng.integers(0, len(keys))]
            w.write(torch_port.embedding_jsonl_record(s, t, items[s], items[t], 0.5))
        w.write("\n")                                # a trailing blank line
    return items


@pytest.mark.parametrize("side", ["src", "tgt", "both"])
def test_jsonl_converter_matches_reference_reader(tmp_path, side):
    import item_alignment_b200 as ia
    src = tmp_path / "embeds.jsonl"
    items = _write_reference_jsonl(src, 60, 48, seed=3)
    ids, mat = torch_port.read_embedding_jsonl(src, side)
    for dt in (torch.float32, torch.bfloat16, torch.float16):
        out = tmp_path / f"embeds_{side}.iacat"
        rows, dim = ia.jsonl_to_catalog(src, out, dt, side)
        assert (rows, dim) == mat.shape
        with ia.CatalogFile(out) as f:
            assert [f.id(i) for i in range(rows)] == ids
            got = f.numpy()
            if dt == torch.float32:
                assert np.array_equal(got, mat)                       # strtof returns the float32 the reference printed
                assert all(np.array_equal(got[i], items[ids[i]]) for i in range(rows))
            else:
                want = torch.from_numpy(mat).to(dt).view(torch.int16).numpy().view(np.uint16)
                assert np.array_equal(got, want)                      # round to nearest even, like torch's .to(dtype)


def test_jsonl_number_parsing_is_eval_then_float32_bit_for_bit(tmp_path):
    """The native parser must give float32(float64(decimal)) -- what `eval` + numpy give a consumer of the reference's file --
    on every kind of decimal the reference can print: shortest float32 reprs of arbitrary bit patterns (incl. subnormals and
    the largest finite values), float64 reprs (17 digits, the ensemble's averaged scores), exponent forms, integers, signed
    zeros, and digit strings long enough to leave the parser's fast path."""
    import json
    import item_alignment_b200 as ia
    rng = np.random.default_rng(7)
    bits = rng.integers(0, 2 ** 32, size=40000, dtype=np.uint64).astype(np.uint32)
    f32 = bits.view(np.float32)
    f32 = f32[np.isfinite(f32)]
    strings = [str(v) for v in f32]                                                      # numpy's shortest round-trip float32 repr
    f64 = np.concatenate([rng.standard_normal(5000), rng.standard_normal(2000) * 1e-30, rng.standard_normal(2000) * 1e30,
                          np.float64(f32[:3000]) * (1 + 2.0 ** -25)])                    # near float32 rounding midpoints
    strings += [repr(float(v)) for v in f64]
    strings += ["0", "-0.0", "1", "-17", "1e10", "1E-10", "3.4028234663852886e+38", "1e-45", "1.401298464324817e-45", "5e-324",
                "123456789012345678901234567890", "0.000000000000000000000000000000123456789012345678901",
                "1.00000000000000000000000000001", "9007199254740993", "9007199254740992e3", "2.5e22", "2.5e23", "7e-23"]
    dim = 50
    strings += ["0.5"] * (-len(strings) % dim)
    path = tmp_path / "numbers.jsonl"
    with open(path, "w") as w:
        for r in range(len(strings) // dim):
            emb = "[" + ",".join(strings[r * dim:(r + 1) * dim]) + "]"
            w.write(json.dumps({"src_item_id": f"row{r}", "src_item_emb": emb, "tgt_item_id": f"t{r}", "tgt_item_emb": "[0.0]",
                                "threshold": 0.5}) + "\n")
    with np.errstate(over="ignore"):
        want = np.array([float(t) for t in strings], dtype=np.float64).astype(np.float32).reshape(-1, dim)
    rows, d = ia.jsonl_to_catalog(path, tmp_path / "numbers.iacat", torch.float32, "src")
    assert (rows, d) == want.shape
    with ia.CatalogFile(tmp_path / "numbers.iacat") as f:
        got = f.numpy()
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32))                # bit for bit, signed zeros included


def test_jsonl_converter_output_does_not_depend_on_the_thread_count(tmp_path):
    import item_alignment_b200 as ia
    src = tmp_path / "embeds.jsonl"
    rng = np.random.default_rng(11)
    items = [np.tanh(rng.standard_normal(96)).astype(np.float32) for _ in range(300)]
    with open(src, "w") as w:                                     # ~1.3 MB: several 64 KiB worker runs; ids repeat across runs
        for i in range(600):
            a, b = int(rng.integers(0, 300)), int(rng.integers(0, 300))
            w.write(torch_port.embedding_jsonl_record(f"id{a}", f"id{b}", items[a], items[b], 0.5))
            if i % 97 == 0:
                w.write("\n")
    ids, mat = torch_port.read_embedding_jsonl(src, "both")
    outs = []
    for t in (1, 2, 5, 16, 0):
        out = tmp_path / f"t{t}.iacat"
        assert ia.jsonl_to_catalog(src, out, torch.float32, "both", threads=t) == mat.shape
        outs.append(out.read_bytes())
    assert all(o == outs[0] for o in outs)
    with ia.CatalogFile(tmp_path / "t5.iacat") as f:
        assert np.array_equal(f.numpy(), mat) and [f.id(i) for i in range(f.rows)] == ids
    # a bad line deep in the file is reported with its line number whatever the thread count
    lines = src.read_text().splitlines(keepends=True)
    lines[450] = '{"src_item_id": "x", "src_item_emb": "[oops]", "tgt_item_id": "y", "tgt_item_emb": "[1.0]"}\n'
    bad = tmp_path / "bad.jsonl"
    bad.write_text("".join(lines))
    for t in (1, 7):
        with pytest.raises(ValueError, match=r"bad\.jsonl:451: src_item_emb is not a float list"):
            ia.jsonl_to_catalog(bad, tmp_path / "bad.iacat", torch.float32, "both", threads=t)


def test_jsonl_converter_errors(tmp_path):
    import item_alignment_b200 as ia
    p = tmp_path / "ragged.jsonl"
    with open(p, "w") as w:
        w.write(torch_port.embedding_jsonl_record("a", "b", np.ones(4, np.float32), np.ones(4, np.float32), 0.5))
        w.write(torch_port.embedding_jsonl_record("c", "d", np.ones(5, np.float32), np.ones(4, np.float32), 0.5))
    with pytest.raises(ValueError):
        ia.jsonl_to_catalog(p, tmp_path / "o.iacat")
    p2 = tmp_path / "nokey.jsonl"
    p2.write_text('{"src_item_id": "a", "threshold": 0.5}\n')
    with pytest.raises(ValueError):
        ia.jsonl_to_catalog(p2, tmp_path / "o.iacat")
    with pytest.raises(ValueError):
        ia.jsonl_to_catalog(tmp_path / "missing.jsonl", tmp_path / "o.iacat")
    with pytest.raises(ValueError):
        ia.jsonl_to_catalog(p, tmp_path / "o.iacat", side="left")


@pytest.mark.skipif(not ref_import.available(), reason="reference tree not present (build container only)")
def test_jsonl_converter_on_the_references_own_file(tmp_path):
    """submit/deepAI_result.jsonl is the reference's own sample of this wire format (15 909 pairs, 1-d embeddings)."""
    import item_alignment_b200 as ia
    src = os.path.join(ref_import.reference_dir(), "submit", "deepAI_result.jsonl")
    ids, mat = torch_port.read_embedding_jsonl(src, "tgt")
    rows, dim = ia.jsonl_to_catalog(src, tmp_path / "deepai.iacat", torch.float32, "tgt")
    assert (rows, dim) == mat.shape
    with ia.CatalogFile(tmp_path / "deepai.iacat") as f:
        assert np.array_equal(f.numpy(), mat)
        assert [f.id(i) for i in (0, 1, rows - 1)] == [ids[0], ids[1], ids[-1]]
