"""CPU oracle for the two-tower vector-similarity path -- TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is part of the product.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import it, and there only as the checker or as the timed CPU baseline.
The product package (``item_alignment_b200``) never imports this package and fails
loudly when its CUDA library is missing.

Parity pinning status
---------------------
The reference (sunzeyeah/item-alignment) ships no tests for this path
(SURVEY.md section 4).  The oracle is therefore pinned against OUTPUTS OF THE
REFERENCE ITSELF: ``tests/golden/make_golden.py`` imports the reference's own
``src/models/base.py`` and ``src/models/loss.py`` by file path (in the build
container, where ``/root/reference`` exists), runs them on seeded inputs and commits
the input/output vectors under ``tests/golden/``.  ``tests/test_oracle_golden.py``
checks both oracle layers against those vectors and against the three weak
artefacts the reference carries (``submit/deepAI_result.jsonl`` labels,
the commented softmax-head weights of ``submit/similarity.py:5-24`` and the
pure-Python inner product of ``pred_bert.py:47-52``).  The "next" rows are pinned the
same way: ``kg_golden.npz`` holds candidate rankings computed by the reference's vendored
torchkge (``TransEModel.inference_scoring_function`` + sort), checked bit for bit in
``tests/test_catalog_file.py``; the embedding-JSONL reader restatement is checked (in the
build container) on the reference's own ``submit/deepAI_result.jsonl``; the head projection
uses the ``VecSimClassificationHead`` vectors of ``head_golden.npz``.

Layers
------
``oracle.torch_port``  the reference's modules restated from the same torch ops
                       (the arithmetic of the reference lives in torch, pinned
                       torch==1.11.0 in requirements.txt:7, run here on 2.11).
``oracle.formula``     explicit numpy formulas (float64 capable) for every score,
                       probability map, loss and gradient, following SURVEY 8(a).
``oracle.ref_import``  path-import of the real reference modules (build container
                       only; never used on the GPU box).
"""
