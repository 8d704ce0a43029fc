"""Validation scores of KiteSeg.val -- task1/kite/losses/miou.py:22-117 (MIouLoss.scorem, MDiceLoss.scorem/scores).

The reference evaluates them on one-hot argmax maps with three host round trips per image; here the
per-image, per-class counts come from one kernel (csrc/dice.cu: label_counts) on uint8 label maps and the
few resulting integers are combined on the host."""
import torch

from ... import _lib as L
from ...ops import _check, _p, _stream


def label_counts(pred_u8, true_u8, n_class):
    """int32 [B, C, 3]: |pred==c & true==c|, |pred==c|, |true==c| per image."""
    _check(pred_u8, true_u8)
    B = pred_u8.shape[0]
    hw = pred_u8.numel() // B
    counts = torch.zeros((B, n_class, 3), dtype=torch.int32, device=pred_u8.device)
    L.label_counts(_p(pred_u8), _p(true_u8), B, n_class, hw, _p(counts), _stream())
    return counts


def _scores(counts, smooth=1.0):
    c = counts.double()
    inter, pr, gt = c[..., 0], c[..., 1], c[..., 2]
    dice = ((2 * inter + smooth) / (pr + gt + smooth)).mean(0)              # MDiceLoss.score: mean over the batch
    iou = ((inter + smooth) / (pr + gt - inter + smooth)).mean(0)           # MIouLoss.score
    return dice, iou


class MDiceLoss:
    @staticmethod
    def from_counts(counts, start_idx=0):
        dice, _ = _scores(counts)
        return dice[start_idx:].mean(), dice


class MIouLoss:
    @staticmethod
    def from_counts(counts, start_idx=0):
        _, iou = _scores(counts)
        return iou[start_idx:].mean(), iou
