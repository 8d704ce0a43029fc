"""Validation scores of KiteSeg.val -- drop-in for task1/kite/losses/miou.py:22-117 (MIouLoss.score / scorem, MDiceLoss.score / scores /
scorem with the reference's signatures and semantics: maps of any softness, [B,C,H,W], scores averaged over the batch).

The reference evaluates them with a handful of ATen reductions per class and three host round trips per image; here ONE kernel
(csrc/dice.cu: score_sums) reduces sum(pr*gt), sum(pr), sum(gt) for every image and class plane, and the few resulting numbers are
combined on the device.  KiteSeg.val itself works on uint8 label maps (csrc/dice.cu: label_counts -> `from_counts`): for hard one-hot
maps both routes give identical scores."""
import torch
import torch.nn as nn

from ... import _lib as L
from ...ops import _c, _check, _p, _stream


def label_counts(pred_u8, true_u8, n_class):
    """int32 [B, C, 3]: |pred==c & true==c|, |pred==c|, |true==c| per image."""
    _check(pred_u8, true_u8)
    B = pred_u8.shape[0]
    hw = pred_u8.numel() // B
    counts = torch.zeros((B, n_class, 3), dtype=torch.int32, device=pred_u8.device)
    L.label_counts(_p(pred_u8), _p(true_u8), B, n_class, hw, _p(counts), _stream())
    return counts


def score_sums(pr, gt):
    """double [B, C, 3]: sum(pr*gt), sum(pr), sum(gt) over the pixels of every image and class plane of [B,C,H,W] maps."""
    if pr.dim() == 3:
        pr, gt = pr[:, None], gt[:, None]
    if pr.shape != gt.shape or pr.dim() != 4:
        raise RuntimeError("score: pr and gt must be [B,C,H,W] maps of the same shape, got %s and %s" % (tuple(pr.shape), tuple(gt.shape)))
    pr = _c(pr if pr.dtype == torch.float32 else pr.float())
    is_i64 = gt.dtype == torch.int64
    gt = _c(gt if (is_i64 or gt.dtype == torch.float32) else gt.float())
    _check(pr, gt)
    B, C, H, W = pr.shape
    out = torch.zeros((B, C, 3), dtype=torch.float64, device=pr.device)
    L.score_sums(_p(pr), _p(gt), int(is_i64), B, C, H * W, _p(out), _stream())
    return out


def _scores(counts, smooth=1.0):
    c = counts.double()
    inter, pr, gt = c[..., 0], c[..., 1], c[..., 2]
    dice = ((2 * inter + smooth) / (pr + gt + smooth)).mean(0)              # MDiceLoss.score: mean over the batch
    iou = ((inter + smooth) / (pr + gt - inter + smooth)).mean(0)           # MIouLoss.score
    return dice, iou


class MIouLoss(nn.Module):
    """miou.py:22-62.  `forward` (the soft-IoU training loss) is not on the stc_tt path (`--los` selects MultiLoss, loss.py:101-110)."""

    def __init__(self, nb_class=4):
        super().__init__()
        self.nb_class = nb_class

    @staticmethod
    def score(pr, gt, smooth=1):
        """mean over the batch of (inter + smooth) / (sum pr + sum gt - inter + smooth), all channels flattened together (miou.py:28-38)."""
        s = score_sums(pr, gt).sum(1)                                        # the reference flattens [B, C*H*W]
        return ((s[:, 0] + smooth) / (s[:, 1] + s[:, 2] - s[:, 0] + smooth)).mean().float()

    @staticmethod
    def scorem(pr, gt, start_idx=0):
        """mean over the classes >= start_idx of the per-class score (miou.py:40-44): one kernel for all classes."""
        _, iou = _scores(score_sums(pr, gt))
        return iou[start_idx:].mean().float()

    @staticmethod
    def from_counts(counts, start_idx=0):
        _, iou = _scores(counts)
        return iou[start_idx:].mean(), iou

    def forward(self, pr, gt, smooth=1e-6):
        raise NotImplementedError("tcct_b200: MIouLoss.forward (soft-IoU training loss) is outside the stc_tt hot path; the validation scores "
                                  "score / scorem are built")


class MDiceLoss(nn.Module):
    """miou.py:64-117."""

    def __init__(self, nb_class=2, bi=False):
        super().__init__()
        self.bi = bi

    @staticmethod
    def score(pr, gt, smooth=1):
        """mean over the batch of (2 inter + smooth) / (sum pr + sum gt + smooth) (miou.py:69-79)."""
        s = score_sums(pr, gt).sum(1)
        return ((2 * s[:, 0] + smooth) / (s[:, 1] + s[:, 2] + smooth)).mean().float()

    @staticmethod
    def scores(pr, gt):
        """per-class scores as a list of Python floats (miou.py:81-84): one kernel, ONE read-back instead of one per class."""
        dice, _ = _scores(score_sums(pr, gt))
        return [float(v) for v in dice.cpu()]

    @staticmethod
    def scorem(pr, gt, start_idx=0):
        dice, _ = _scores(score_sums(pr, gt))
        return dice[start_idx:].mean().float()

    @staticmethod
    def from_counts(counts, start_idx=0):
        dice, _ = _scores(counts)
        return dice[start_idx:].mean(), dice

    def forward(self, pr, gt):
        raise NotImplementedError("tcct_b200: MDiceLoss.forward (per-image soft Dice training loss) is outside the stc_tt hot path; the "
                                  "validation scores score / scores / scorem are built")
