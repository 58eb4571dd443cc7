"""Drop-in loss modules (reference src/models/loss.py) and the loss ladder every reference model repeats.

HingeLoss / EuclideanDistanceLoss keep the reference signatures `(margin, size_average, reduce, reduction)`
and `forward(input, target)` with target in {-1, +1}; the arithmetic runs in the CUDA elementwise kernel.
"""
import torch
from torch import nn
from torch.nn.modules.loss import _Loss

from . import functional as F_


class EuclideanDistanceLoss(_Loss):
    """l_n = x_n ** y_n (reference loss.py:6-68)."""
    __constants__ = ['reduction']

    def __init__(self, size_average=None, reduce=None, reduction: str = 'mean') -> None:
        super().__init__(size_average, reduce, reduction)

    def forward(self, input: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
        red = self.reduction if self.reduction in ("sum", "mean") else "none"
        out = F_.score_loss("euclidean", input, target, 1.0, red)
        return out.view(input.shape) if red == "none" else out


class HingeLoss(_Loss):
    """l_n = max(0, margin - y_n * x_n) (reference loss.py:71-134)."""
    __constants__ = ['margin', 'reduction']
    margin: float

    def __init__(self, margin: float = 1.0, size_average=None, reduce=None, reduction: str = 'mean') -> None:
        super().__init__(size_average, reduce, reduction)
        self.margin = margin

    def forward(self, input: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
        red = self.reduction if self.reduction in ("sum", "mean") else "none"
        out = F_.score_loss("hinge", input, target, self.margin, red)
        return out.view(input.shape) if red == "none" else out


def build_loss_fct(config):
    """The constructor ladder of every reference model, e.g. src/models/text.py:1400-1409."""
    if config.loss_type == "cosine":
        return nn.CosineEmbeddingLoss(margin=config.loss_margin)
    if config.loss_type == "bce":
        return nn.BCEWithLogitsLoss()
    if config.loss_type == "euclidean":
        return EuclideanDistanceLoss()
    if config.loss_type == "hinge":
        return HingeLoss(margin=config.loss_margin)
    return nn.CrossEntropyLoss()


def apply_loss_ladder(config, loss_fct, logits, src_embeds, tgt_embeds, labels, num_labels=2):
    """The dispatch ladder, src/models/text.py:1468-1477 (unfused: module by module, as the reference).
    bce accepts the collate functions' long labels (the reference crashes on them)."""
    if config.loss_type == "cosine":
        return loss_fct(src_embeds, tgt_embeds, (labels * 2 - 1).view(-1))
    if config.loss_type == "ce":
        return loss_fct(logits.view(-1, num_labels), labels.view(-1))
    if config.loss_type == "hinge" or config.loss_type == "euclidean":
        return loss_fct(logits.view(-1), (labels * 2 - 1).view(-1))
    return loss_fct(logits.view(-1), labels.view(-1).to(logits.dtype))


def two_tower_step(classifier, config, features_1, features_2, labels=None):
    """classifier(...) + ladder in the fused single-pass form.  Returns the five fields of the reference's
    SequenceClassifierOutput (base.py:160-186): dict(loss, logits, probs, src_embeds, tgt_embeds)."""
    from .heads import TwoTowerClassificationHead, VecSimClassificationHead
    loss = None
    if isinstance(classifier, VecSimClassificationHead):
        if labels is not None:
            if config.loss_type not in ("cosine", "bce", "hinge", "euclidean"):
                # the reference ladder (text.py:1468-1477) would hand the [N] similarity vector to CrossEntropyLoss and
                # fail inside it; fail here, loudly, instead of returning loss=None
                raise ValueError(f"loss_type {config.loss_type!r} is not defined for a vector-similarity head "
                                 "(use cosine, bce, hinge or euclidean; 'ce' belongs to the softmax two-tower head)")
            x, y, logits, probs, loss = classifier.forward_with_loss(features_1, features_2, labels, config.loss_type,
                                                                     getattr(config, "loss_margin", 1.0))
        else:
            x, y, logits, probs = classifier(features_1, features_2)
    elif isinstance(classifier, TwoTowerClassificationHead):
        if labels is not None and config.loss_type == "ce":
            x, y, logits, probs, loss = classifier.forward_with_loss(features_1, features_2, labels)
        else:
            x, y, logits, probs = classifier(features_1, features_2)
            if labels is not None:
                loss = apply_loss_ladder(config, build_loss_fct(config), logits, x, y, labels)
    else:
        raise TypeError(f"unsupported classifier {type(classifier).__name__}")
    return dict(loss=loss, logits=logits, probs=probs, src_embeds=x, tgt_embeds=y)
