"""Drop-in heads: same constructors, attributes, state-dict keys and return tuples as the reference's
src/models/base.py, with the similarity / probability / loss arithmetic on the sm_100a kernels.

    VecSimClassificationHead(config).forward(f1, f2)      -> (x, y, sim, probs)      base.py:44-88
    TwoTowerClassificationHead(h, dropout, num_labels)     -> (x, y, logits, probs)   base.py:96-117
    InnerProduct(normalize=False).forward(x1, x2)          -> [N]                     base.py:25-34

The dense -> tanh projection of the VecSim head runs as one tcgen05 GEMM with bias + tanh (+ dropout in train() mode) in its
epilogue (csrc/projection.cu) for bf16 / fp16 features, its backward as two more tcgen05 GEMMs (csrc/projection_bwd.cu); fp32
features stay on cuBLAS/ATen (boundary-adjacent).
`forward_with_loss` is the fused single-pass entry the reference's forward() can call instead of
classifier(...) + the loss ladder.
"""
import torch
from torch import nn

from . import functional as F_


class PairSimilarity(nn.Module):
    """`self.similarity` of the reference head (nn.CosineSimilarity / nn.PairwiseDistance(p) as built at
    base.py:57-62) on the CUDA kernel."""

    def __init__(self, measure):
        super().__init__()
        if measure not in ("cosine", "l1", "l2", "inner_product"):
            raise ValueError(f"Unsupported similarty measure: {measure}")
        self.measure = measure

    def forward(self, x1, x2):
        return F_.pair_similarity(self.measure, x1, x2)

    def extra_repr(self):
        return self.measure


class InnerProduct(nn.Module):
    """reference base.py:10-34 (torch.bmm of (N,1,D) x (N,D,1)); normalize=True applies F.normalize first."""
    __constants__ = ['normalize']
    normalize: bool

    def __init__(self, normalize: bool = False) -> None:
        super().__init__()
        self.normalize = normalize

    def forward(self, x1: torch.Tensor, x2: torch.Tensor) -> torch.Tensor:
        bs, hs = x1.shape          # 2-D input required, as in the reference (base.py:30)
        if self.normalize:
            x1 = nn.functional.normalize(x1, p=2, dim=1)
            x2 = nn.functional.normalize(x2, p=2, dim=1)
        return F_.pair_similarity("inner_product", x1, x2)


class VecSimClassificationHead(nn.Module):
    """reference base.py:37-88.  Reads config.{cls_layers, cls_pool, hidden_size, classifier_dropout,
    hidden_dropout_prob, similarity_measure}; parameters are `dense.weight/bias` (checkpoint compatible)."""

    def __init__(self, config):
        super().__init__()
        self.config = config
        cls_layers = [int(i) for i in config.cls_layers.split(",")]
        length = 1 if config.cls_pool == "avg" else len(cls_layers)
        self.dense = nn.Linear(config.hidden_size * length, config.hidden_size)
        classifier_dropout = (
            config.classifier_dropout if config.classifier_dropout is not None else config.hidden_dropout_prob
        )
        self.dropout = nn.Dropout(classifier_dropout)
        if config.similarity_measure == "inner_product":
            self.similarity = InnerProduct(normalize=False)
            self.sigmoid = nn.Sigmoid()
        elif config.similarity_measure in ("cosine", "l1", "l2"):
            self.similarity = PairSimilarity(config.similarity_measure)
        else:
            raise ValueError(f"Unsupported similarty measure: {config.similarity_measure}")

    def project(self, features):
        x = self.dropout(features)
        x = self.dense(x)
        x = torch.tanh(x)
        return self.dropout(x)

    def _fused_dtype(self, f1, f2):
        """dtype the fused tcgen05 projection would run in, or None when the library path must be used: fp32 arithmetic
        (TF32 tensor cores would change the numerics) or odd shapes."""
        if not (f1.is_cuda and f2.is_cuda) or f1.dim() != 2 or f1.shape != f2.shape:
            return None
        dt = torch.get_autocast_dtype("cuda") if torch.is_autocast_enabled("cuda") else f1.dtype
        if dt not in (torch.bfloat16, torch.float16):
            return None
        if not torch.is_autocast_enabled("cuda") and self.dense.weight.dtype != dt:
            return None
        k, h = self.dense.in_features, self.dense.out_features
        return dt if (f1.shape[1] == k and k % 8 == 0 and h % 8 == 0 and h <= 4096) else None

    def _dropout_state(self):
        """(p, seed, step) of this call's dropout masks: counter-based (csrc/philox.cuh), seeded from torch's global seed, one
        step per call -- reproducible under torch.manual_seed, no host synchronisation.  p = 0 outside train()."""
        p = float(self.dropout.p) if self.training else 0.0
        if p <= 0.0:
            return 0.0, 0, 0
        self._dropout_step = getattr(self, "_dropout_step", 0) + 1
        return p, torch.initial_seed() & 0xFFFFFFFFFFFFFFFF, self._dropout_step & 0xFFFFFFFF

    def project_pair(self, features_1, features_2):
        """Both sides through dropout -> dense -> tanh -> dropout: fused GEMM launches when possible, forward
        and backward (train() mode included: the dropout masks are generated inside the kernels)."""
        dt = self._fused_dtype(features_1, features_2)
        if dt is None:
            return self.project(features_1), self.project(features_2)
        p, seed, step = self._dropout_state()
        if p > 0.0:
            return F_.project_tanh_train(features_1.to(dt), features_2.to(dt), self.dense.weight, self.dense.bias, p, seed, step)
        return F_.project_tanh(features_1.to(dt), features_2.to(dt), self.dense.weight, self.dense.bias)

    def forward(self, features_1, features_2):
        measure = self.config.similarity_measure
        if measure not in ("cosine", "l1", "l2", "inner_product"):
            raise ValueError(f"Unsupported similarty measure: {measure}")
        dt = self._fused_dtype(features_1, features_2)
        needs_grad = torch.is_grad_enabled() and (
            features_1.requires_grad or features_2.requires_grad or self.dense.weight.requires_grad)
        if dt is not None and not needs_grad and not (self.training and self.dropout.p > 0):
            # inference: projection, score and probability map in one launch
            return F_.project_score(measure, features_1.to(dt), features_2.to(dt), self.dense.weight, self.dense.bias,
                                    want_embeds=True)
        x, y = self.project_pair(features_1, features_2)
        sim, probs = F_.pair_score(measure, x, y)
        return x, y, sim, probs

    def forward_with_loss(self, features_1, features_2, labels, loss_type, margin=1.0):
        """Head + loss ladder (reference text.py:1468-1477) + their backward.  16-bit features: forward GEMM (bias + tanh +
        dropout epilogue), ONE pair launch that produces sim, probs, loss and d_pre (loss, similarity, dropout and tanh backward
        folded together), and the two gradient GEMMs in backward() -- all on this library's kernels.  Otherwise the projection
        stays on torch and the pair launch is fused as before.  Returns (x, y, sim, probs, loss)."""
        dt = self._fused_dtype(features_1, features_2)
        if dt is not None and torch.is_grad_enabled():
            p, seed, step = self._dropout_state()
            return F_.project_score_loss_train(self.config.similarity_measure, loss_type, features_1.to(dt), features_2.to(dt),
                                               self.dense.weight, self.dense.bias, labels, p, seed, step, margin)
        x, y = self.project_pair(features_1, features_2)
        sim, probs, loss = F_.pair_score_loss(self.config.similarity_measure, loss_type, x, y, labels, margin)
        return x, y, sim, probs, loss


class TwoTowerClassificationHead(nn.Module):
    """reference base.py:91-117: logits = out_proj(cat(drop(f1), drop(f2))), probs = Softmax()(logits).
    The concat copy is never made: the kernel reads both towers' rows directly."""

    def __init__(self, hidden_size, dropout=0.0, num_labels=2):
        super().__init__()
        self.dropout = nn.Dropout(dropout)
        self.out_proj = nn.Linear(hidden_size * 2, num_labels)
        self.softmax = torch.nn.Softmax(dim=1)     # the reference's implicit dim for 2-D input
        self.num_labels = num_labels

    def _library_path(self, x, y):
        logits = self.out_proj(torch.cat((x, y), dim=1))
        return logits, self.softmax(logits)

    def _kernel_ok(self, x):
        e = 4 if x.dtype == torch.float32 else 8
        h = x.shape[1]
        return self.num_labels == 2 and x.is_cuda and h % e == 0 and h // e <= 256

    def forward(self, features_1, features_2):
        x = self.dropout(features_1)
        y = self.dropout(features_2)
        if self._kernel_ok(x):
            logits, probs = F_.softmax_head(x, y, self.out_proj.weight, self.out_proj.bias)
        else:   # num_labels != 2 or exotic widths: plain GPU library ops (still no CPU path)
            if not x.is_cuda:
                raise RuntimeError("item_alignment_b200 runs on CUDA tensors only (no CPU fallback)")
            logits, probs = self._library_path(x, y)
        return x, y, logits, probs

    def forward_with_loss(self, features_1, features_2, labels):
        """Head + nn.CrossEntropyLoss (reference text.py:1408-1409,1473) + backward in one pass.
        Returns (x, y, logits, probs, loss)."""
        x = self.dropout(features_1)
        y = self.dropout(features_2)
        if self._kernel_ok(x):
            logits, probs, loss = F_.softmax_head_ce(x, y, self.out_proj.weight, self.out_proj.bias, labels)
        else:
            if not x.is_cuda:
                raise RuntimeError("item_alignment_b200 runs on CUDA tensors only (no CPU fallback)")
            logits, probs = self._library_path(x, y)
            loss = nn.functional.cross_entropy(logits.view(-1, self.num_labels), labels.view(-1))
        return x, y, logits, probs, loss
