"""The `--los` loss registry -- drop-in for task1/kite/losses/loss.py:70-110.

`get_loss('di'|'dice')` -> MultiLoss(DiceLoss), anything else -> MultiLoss(MSELoss), called as
criterion(logits [B,C,H,W], gt) with gt the int64 one-hot [B,C,H,W] the reference's loop builds
(loop_seg.py:119) or a class-index map [B,H,W].  Softmax over classes, the per-class Dice / MSE terms
and their gradient run fused in csrc/dice.cu (one pass forward, one pass backward)."""
import torch
import torch.nn as nn

from ... import ops as O


class DiceLoss(nn.Module):
    """1 - (1 + 2 sum(p g)) / (1 + sum p + sum g), sums over the whole batch (loss.py:9-37)."""
    __name__ = 'DiceLoss'
    mode = 0

    def __init__(self, bi=False):
        super().__init__()
        if bi:
            raise NotImplementedError("tcct_b200: DiceLoss(bi=True) (dice2) is not on the stc_tt training path")


class MSELossTag(nn.MSELoss):
    mode = 1


class MultiLoss(nn.Module):
    __name__ = 'MultiLoss'

    def __init__(self, losses, weight=None):
        super().__init__()
        self.losses = losses
        self.WEIGHT = [1, ] * 40 if weight is None else weight
        self._cache = (None, None, None)

    def labels(self, gt, n_class):
        """uint8 index map of `gt`, converted once per target tensor (the deep-supervision loop calls the
        criterion four times with the same target, loopback.py:62-73)."""
        # the cache entry keeps `gt` itself alive and is matched by identity + version: a freed target whose address the
        # caching allocator hands to the next batch's one-hot (loop_seg.py:119 builds a fresh one every step) cannot hit
        ref, ver, lab = self._cache
        if ref is not gt or ver != gt._version:
            lab = O.labels_u8(gt.contiguous(), n_class)
            self._cache = (gt, gt._version, lab)
        return lab

    def forward(self, pr, gt, **args):
        if any(w != 1 for w in self.WEIGHT[:pr.shape[1]]):
            raise NotImplementedError("tcct_b200: per-class loss weights other than 1 are not implemented")
        mode = getattr(self.losses, "mode", 1)
        return O.DiceFn.apply(pr.contiguous(), self.labels(gt, pr.shape[1]), mode)


def get_loss(loss='di', **args):
    if loss == 'dice' or loss == 'di':
        print(loss, 'DiceLoss()')
        los = DiceLoss(bi=False)
    else:
        print('MSE')
        los = MSELossTag()
    return MultiLoss(los)
