"""The reference's head / loss modules restated from the same torch ops -- TEST INFRASTRUCTURE.

Every function cites the reference file:line (under /root/reference) it follows.  The
arithmetic of the reference lives in torch (pinned torch==1.11.0, requirements.txt:7;
run here on 2.11 CPU): this layer composes the same ops in the same order as the
reference modules, so it is both the parity checker and -- timed with all host threads
-- the "port" CPU baseline of bench.py.  It is pinned against golden vectors produced
by the real reference modules (tests/golden/make_golden.py).

bf16/fp16 policy (SURVEY 7 "bf16 oracle"): the oracle always runs on ``x.float()`` of
the already-quantised inputs and emits fp32.
"""
import numpy as np
import torch
import torch.nn.functional as F

MEASURES = ("inner_product", "cosine", "l1", "l2")
LOSSES = ("bce", "hinge", "euclidean", "cosine", "ce")


# --------------------------------------------------------------------------- scores
def inner_product(x1, x2, normalize=False):
    """InnerProduct.forward, src/models/base.py:29-34 (batched 1xD . Dx1 via torch.bmm)."""
    bs, hs = x1.shape
    if normalize:
        x1 = F.normalize(x1, p=2, dim=1)
        x2 = F.normalize(x2, p=2, dim=1)
    return torch.bmm(x1.view(bs, 1, hs), x2.view(bs, hs, 1)).reshape(-1)


def similarity(measure, x, y):
    """The ``self.similarity`` module chosen at src/models/base.py:54-64."""
    x = x.float()
    y = y.float()
    if measure == "inner_product":
        return inner_product(x, y)                              # base.py:55
    if measure == "cosine":
        return F.cosine_similarity(x, y, dim=1, eps=1e-8)       # nn.CosineSimilarity(), base.py:58
    if measure == "l1":
        return F.pairwise_distance(x, y, p=1.0, eps=1e-6)       # nn.PairwiseDistance(p=1), base.py:60
    if measure == "l2":
        return F.pairwise_distance(x, y, p=2.0, eps=1e-6)       # nn.PairwiseDistance(p=2), base.py:62
    raise ValueError(f"Unsupported similarty measure: {measure}")   # base.py:64 (sic)


def probs_of(measure, sim):
    """Probability map of VecSimClassificationHead.forward, src/models/base.py:79-86."""
    if measure == "cosine":
        return (sim + 1) / 2
    if measure in ("l1", "l2"):
        return torch.exp(-sim)
    if measure == "inner_product":
        return torch.sigmoid(sim)
    raise ValueError(f"Unsupported similarty measure: {measure}")


def vecsim_head(measure, f1, f2, dense_w=None, dense_b=None):
    """VecSimClassificationHead.forward (eval mode: dropout = identity), base.py:66-88.

    With dense_w None the projection is skipped (scores taken on the given embeddings).
    Returns (x, y, sim, probs)."""
    if dense_w is not None:
        x = torch.tanh(F.linear(f1.float(), dense_w, dense_b))   # base.py:67-70
        y = torch.tanh(F.linear(f2.float(), dense_w, dense_b))   # base.py:72-75
    else:
        x, y = f1.float(), f2.float()
    sim = similarity(measure, x, y)                              # base.py:77
    return x, y, sim, probs_of(measure, sim)


def two_tower_head(f1, f2, w, b):
    """TwoTowerClassificationHead.forward (eval mode), src/models/base.py:103-117.
    logits = cat(f1, f2) . W^T + b ; probs = Softmax over dim 1 (implicit dim for 2-D)."""
    logits = F.linear(torch.cat((f1.float(), f2.float()), dim=1), w.float(), b.float())
    return logits, torch.softmax(logits, dim=1)


# --------------------------------------------------------------------------- losses
def _reduce(loss, reduction):
    if reduction == "sum":
        return loss.sum()
    if reduction == "mean":
        return loss.mean()
    return loss


def hinge_loss(inp, target, margin=1.0, reduction="mean"):
    """HingeLoss.forward, src/models/loss.py:126-134 (torch.max(zero, .) splits the
    sub-gradient at the kink in half)."""
    zero = torch.zeros(1, device=inp.device)
    return _reduce(torch.max(zero, margin - inp * target), reduction)


def euclidean_loss(inp, target, reduction="mean"):
    """EuclideanDistanceLoss.forward, src/models/loss.py:61-68 (input ** target)."""
    return _reduce(torch.pow(inp, target), reduction)


def loss_ladder(loss_type, logits, src_embeds, tgt_embeds, labels, margin=1.0, num_labels=2):
    """The loss dispatch repeated in every model forward, e.g. src/models/text.py:1468-1477,
    with the constructors of text.py:1400-1409.  ``labels`` are torch.long {0,1}
    (src/data/data.py:237).  bce: the reference passes long labels and crashes
    (SURVEY 2a); parity target = same op with labels.float()."""
    if loss_type == "cosine":
        return F.cosine_embedding_loss(src_embeds.float(), tgt_embeds.float(),
                                       (labels * 2 - 1).view(-1), margin=margin)
    if loss_type == "ce":
        return F.cross_entropy(logits.view(-1, num_labels), labels.view(-1))
    if loss_type == "hinge":
        return hinge_loss(logits.view(-1), (labels * 2 - 1).view(-1), margin)
    if loss_type == "euclidean":
        return euclidean_loss(logits.view(-1), (labels * 2 - 1).view(-1))
    return F.binary_cross_entropy_with_logits(logits.view(-1), labels.view(-1).float())


def pair_score_loss_fwd_bwd(measure, loss_type, x, y, labels, margin=1.0):
    """One training step of the path: head score (base.py:77-86) -> ladder
    (text.py:1468-1477) -> loss.backward() (finetune_text.py:479-482).
    Returns sim, probs, loss, dx, dy (all fp32)."""
    x = x.detach().float().clone().requires_grad_(True)
    y = y.detach().float().clone().requires_grad_(True)
    sim = similarity(measure, x, y)
    probs = probs_of(measure, sim)
    loss = loss_ladder(loss_type, sim, x, y, labels, margin)
    loss.backward()
    return sim.detach(), probs.detach(), loss.detach(), x.grad, y.grad


def softmax_head_ce_fwd_bwd(x, y, w, b, labels):
    """cls/"softmax" two-tower step: base.py:103-117 then nn.CrossEntropyLoss (text.py:1408-1409,
    1473) and backward.  Returns logits, probs, loss, dx, dy, dW, db."""
    x = x.detach().float().clone().requires_grad_(True)
    y = y.detach().float().clone().requires_grad_(True)
    w = w.detach().float().clone().requires_grad_(True)
    b = b.detach().float().clone().requires_grad_(True)
    logits, probs = two_tower_head(x, y, w, b)
    loss = F.cross_entropy(logits.view(-1, 2), labels.view(-1))
    loss.backward()
    return logits.detach(), probs.detach(), loss.detach(), x.grad, y.grad, w.grad, b.grad


# --------------------------------------------------------------------------- labels
def threshold_labels(probs, threshold):
    """finetune_text.py:576-580: ``probs >= threshold`` on the fp32 probs as numpy evaluates
    it with a float64 threshold from np.arange (comparison in float64)."""
    import numpy as np
    return np.asarray(probs.detach().cpu().numpy(), dtype=np.float32) >= np.float64(threshold)


# --------------------------------------------------------------------------- retrieval
def all_pairs_scores(measure, q, c):
    """The reference's pairwise similarity (base.py:54-62) evaluated for every (query,
    catalog) pair (SURVEY 1 "Retrieval ... defined as").  inner/cosine as a dense
    contraction (the in-batch einsum precedent, multimodal.py:926-930); l1/l2 broadcast like
    torchkge (torchkge/utils/dissimilarities.py:11-25) but with PairwiseDistance's eps."""
    q = q.float()
    c = c.float()
    if measure == "inner_product":
        return q @ c.t()
    if measure == "cosine":
        qn = q / q.norm(dim=1, keepdim=True).clamp_min(1e-8)
        cn = c / c.norm(dim=1, keepdim=True).clamp_min(1e-8)
        return qn @ cn.t()
    d = q[:, None, :] - c[None, :, :] + 1e-6
    if measure == "l1":
        return d.abs().sum(-1)
    if measure == "l2":
        return d.pow(2).sum(-1).sqrt()
    raise ValueError(f"Unsupported similarty measure: {measure}")


def retrieve_topk(measure, q, c, k, block=256):
    """Per-query top-k of all_pairs_scores by a STABLE sort (ties -> lower catalog index):
    descending for inner/cosine (torchkge/inference.py:243-246 style), ascending for the
    l1/l2 distances.  Returns (scores [Q,k] fp32, idx [Q,k] int64)."""
    descending = measure in ("inner_product", "cosine")
    out_s, out_i = [], []
    for s in range(0, q.shape[0], block):
        sc = all_pairs_scores(measure, q[s:s + block], c)
        v, i = torch.sort(sc, dim=1, descending=descending, stable=True)
        out_s.append(v[:, :k].clone())
        out_i.append(i[:, :k].clone())
    return torch.cat(out_s), torch.cat(out_i)


# --------------------------------------------------------------------------- plugins
def compute_passthrough(item_emb_1, item_emb_2):
    """submit/similarity.py:27-28 (active body: ensemble pass-through)."""
    return item_emb_2[0]


def compute_softmax_head(item_emb_1, item_emb_2, w, b):
    """submit/similarity.py:19-24 (commented numpy softmax head)."""
    import numpy as np
    emb = np.array(item_emb_1 + item_emb_2)
    logits = w.dot(emb) + b
    el = np.exp(logits)
    return (el / np.sum(el))[1]


def compute_inner(item_emb_1, item_emb_2, bias=0.0):
    """pred_bert.py:47-52 (pure-Python inner product)."""
    s = 0.0
    for a, b in zip(item_emb_1, item_emb_2):
        s += a * b
    s += bias
    return s


# --------------------------------------------------------------------------- evaluation sweep / threshold search
def threshold_sweep(probs, labels, thresholds):
    """finetune_text.py:576-580: sklearn precision / recall / f1 of `model_probs >= threshold` per threshold."""
    import numpy as np
    from sklearn.metrics import f1_score, precision_score, recall_score
    import warnings
    probs = np.asarray(probs, dtype=np.float32)
    labels = np.asarray(labels)
    p, r, f = [], [], []
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for threshold in thresholds:
            pred = probs >= threshold
            p.append(precision_score(labels, pred)); r.append(recall_score(labels, pred)); f.append(f1_score(labels, pred))
    return np.array(p), np.array(r), np.array(f)


def find_best_f1_and_threshold(scores, labels, high_score_more_similar=True):
    """finetune_bert.py:72-106, restated loop for loop."""
    assert len(scores) == len(labels)
    rows = sorted(zip(scores, labels), key=lambda x: x[0], reverse=high_score_more_similar)
    best_f1 = best_precision = best_recall = best_acc = 0
    threshold = 0
    nextract = ncorrect = fneg = 0
    total_num_duplicates = sum(labels)
    neg_total = len(labels) - total_num_duplicates
    for i in range(len(rows) - 1):
        score, label = rows[i]
        nextract += 1
        if label == 1:
            ncorrect += 1
        else:
            fneg += 1
        if ncorrect > 0:
            precision = ncorrect / nextract
            recall = ncorrect / total_num_duplicates
            f1 = 2 * precision * recall / (precision + recall)
            acc = (ncorrect + neg_total - fneg) / len(labels)
            if f1 > best_f1:
                best_f1, best_precision, best_recall = f1, precision, recall
                threshold = (rows[i][0] + rows[i + 1][0]) / 2
                best_acc = acc
    return best_acc, best_f1, best_precision, best_recall, threshold


def gcn_pair_loop(measure, node_embeddings, pairs, loss_type=None, margin=1.0):
    """The per-pair loop of GCNTwoTower.forward (src/models/graph.py:87-117) with a vector-similarity head in place of
    the cls head: one head call per pair, concatenation per iteration, summed per-pair losses divided by len(pairs)."""
    sims, probs_all, loss = [], [], None
    for pair in pairs:
        x = node_embeddings[pair["src_idx"]].unsqueeze(0).float()
        y = node_embeddings[pair["tgt_idx"]].unsqueeze(0).float()
        sim = similarity(measure, x, y)
        sims.append(sim)
        probs_all.append(probs_of(measure, sim))
        if loss_type is not None and pair.get("item_label") is not None:
            lab = torch.tensor([int(pair["item_label"])], dtype=torch.long)
            li = loss_ladder(loss_type, sim, x, y, lab, margin)
            loss = li if loss is None else loss + li
    if loss is not None:
        loss = loss / len(pairs)
    return torch.cat(sims), torch.cat(probs_all), loss


# ----------------------------------------------------------------------------------------------
# Rank-4 "next" row: the embedding JSONL wire format and torchkge-style candidate ranking
def embedding_jsonl_record(src_item_id, tgt_item_id, src_embed, tgt_embed, threshold):
    """One line of the reference's embedding dump, finetune_text.py:784-792: each embedding is
    ','.join(str(v) for v in ndarray) wrapped in brackets, the record goes through json.dumps."""
    import json
    src_item_emb = ','.join([str(emb) for emb in src_embed])
    tgt_item_emb = ','.join([str(emb) for emb in tgt_embed])
    rd = {"src_item_id": src_item_id, "src_item_emb": f"[{src_item_emb}]",
          "tgt_item_id": tgt_item_id, "tgt_item_emb": f"[{tgt_item_emb}]", "threshold": threshold}
    return json.dumps(rd) + "\n"


def read_embedding_jsonl(path, side="both"):
    """What a consumer of that file sees: json.loads per line and the float list evaluated like
    model_ensemble.py:112 does (`eval(d['tgt_item_emb'])`; literal_eval here, same value for a list of
    floats), first occurrence of an item id kept.  Returns (ids, float32 matrix)."""
    import ast
    import json
    ids, rows, seen = [], [], set()
    with open(path, "r", encoding="utf-8") as r:
        for line in r:
            if not line.strip():
                continue
            d = json.loads(line.strip())
            for s in (("src", "tgt") if side == "both" else (side,)):
                key = d[f"{s}_item_id"]
                if key in seen:
                    continue
                seen.add(key)
                ids.append(key)
                rows.append(np.asarray(ast.literal_eval(d[f"{s}_item_emb"]), dtype=np.float64).astype(np.float32))
    return ids, np.stack(rows)


def kg_dissimilarity(kind, a, b):
    """torchkge/torchkge/utils/dissimilarities.py:11-25: L1 = ||a-b||_1, L2 = ||a-b||_2 ** 2."""
    if kind == "L1":
        return (a - b).norm(p=1, dim=-1)
    return (a - b).norm(p=2, dim=-1) ** 2


def kg_rank_entities(ent_emb, rel_emb, known_entities, known_relations, top_k, missing="tails", kind="L2"):
    """EntityInference.evaluate for a TransE model, torchkge/torchkge/inference.py:216-246 with
    TranslationModel.inference_scoring_function (models/interfaces.py:240-260): tails:
    -dissimilarity((h + r)[:, None, :], candidates); heads: -dissimilarity(candidates + r[:, None, :], t[:, None, :]);
    `scores.sort(descending=True)` (made stable here so ties are defined: lower entity index first), top_k."""
    cand = ent_emb.float()[None, :, :]
    e = ent_emb.float()[known_entities]
    r = rel_emb.float()[known_relations]
    if missing == "tails":
        scores = -kg_dissimilarity(kind, (e + r)[:, None, :], cand)
    else:
        scores = -kg_dissimilarity(kind, cand + r[:, None, :], e[:, None, :])
    s, idx = torch.sort(scores, dim=1, descending=True, stable=True)
    return idx[:, :top_k], s[:, :top_k], scores
