"""CPU tests of host-side decisions that need no GPU: when the drop-in head may take the fused projection, argument
validation of the KG ranking mirror, and the planner-facing size queries of the C ABI."""
import types

import pytest
import torch


def _head(measure="cosine", hidden=64, layers="12", p=0.1):
    import item_alignment_b200 as ia
    cfg = types.SimpleNamespace(cls_layers=layers, cls_pool="cls", hidden_size=hidden, classifier_dropout=p,
                                hidden_dropout_prob=p, similarity_measure=measure)
    return ia.VecSimClassificationHead(cfg)


def test_fused_projection_is_only_taken_where_it_is_exact():
    head = _head().eval()
    f = torch.randn(4, 64)
    assert head._fused_dtype(f, f) is None                                   # CPU tensors: never (and scoring them raises)
    with pytest.raises(RuntimeError):
        head(f, f)
    meta16 = torch.empty(4, 64, dtype=torch.bfloat16, device="meta")
    assert head._fused_dtype(meta16, meta16) is None                          # not a CUDA tensor
    with pytest.raises(ValueError):
        _head(measure="softmax")                                             # reference base.py:64 wording and type
    head2 = _head(layers="11,12")
    assert head2.dense.in_features == 128 and head2.dense.out_features == 64   # cls_layers concatenation, base.py:47-49
    assert set(head2.state_dict()) == {"dense.weight", "dense.bias"}           # checkpoint compatible


def test_rank_entities_validates_like_the_reference():
    import item_alignment_b200 as ia
    e, r = torch.randn(10, 8), torch.randn(3, 8)
    idx = torch.zeros(2, dtype=torch.long)
    with pytest.raises(ValueError, match="missing entity should either be 'heads' or 'tails'"):   # torchkge/inference.py:208
        ia.rank_entities(e, r, idx, idx, missing="both")
    with pytest.raises(ValueError):
        ia.rank_entities(e, r, idx, idx, dissimilarity_type="cosine")


def test_projection_workspace_query_is_host_side_and_consistent():
    from item_alignment_b200._lib import lib
    L = lib()
    assert L.ia_project_score_workspace_bytes(0, 1024, 1024) == 0            # invalid shapes -> 0, no device needed
    assert L.ia_project_score_workspace_bytes(100, 1023, 1024) == 0          # k_in % 8 != 0
    try:
        b = L.ia_project_score_workspace_bytes(65536, 1024, 1024)
    except Exception:                                                        # pragma: no cover
        pytest.skip("needs a CUDA runtime to read the SM count")
    if b:                                                                    # parts * 4 quarters * n * 16 bytes
        assert b % (4 * 65536 * 16) == 0 and 1 <= b // (4 * 65536 * 16) <= 8
