"""FeatConSuper -- drop-in for task1/nets/fcs.py:25-96.

The training path (RegNet.regular_udh) never calls these methods one by one: mask-gather, sort, rank
binning, prototype similarity and the reduction run fused in csrc/fpolar.cu (ops.FeaturePolarFn).
The methods below keep the reference's public surface for callers that use them directly on small
tensors; they are thin compositions of tensor ops on whatever device the inputs live on."""
import torch
import torch.nn as nn

BINS = 32


def points_selection_bins(feat, prob, true, card=512, **args):
    """Rows of `feat` [N,L] where true>.5, ranked by prob (descending), cut into 32 equal rank bins -> bin means [32,L]."""
    assert feat.dim() == 2, 'feat should contains N*L two dims!'
    sel = true.reshape(-1) > .5
    f, p = feat[sel], prob.reshape(-1)[sel]
    order = torch.sort(p, dim=-1, descending=True)[1]
    n = f.shape[0] // BINS
    return torch.stack([f[order[i * n:(i + 1) * n]].mean(dim=0) for i in range(BINS)])


class FeatConSuper(nn.Module):
    def __init__(self, con='cos', mode='bins', *args):
        super().__init__()
        self.__name__ = con
        self.mse = nn.MSELoss(reduction='mean')
        self.con = con
        self.forward = self.cosinesim
        self.func = points_selection_bins

    def cosinesim(self, q, k):
        """-(mean of all pairwise UN-normalised dot products) / feature length (fcs.py:63-67)."""
        return -torch.einsum('nc,kc->nk', [q, k]).mean() / q.shape[-1]

    def foreach_loss(self, fts, gts):
        los = 0
        for i, (ft, gt) in enumerate(zip(fts, gts)):
            los = los + self.forward(ft, gt)
        return los

    def select1(self, feat, pred, true, mask=None, ksize=5, card_select=16):
        assert feat.shape[-2:] == true.shape[-2:], 'shape of feat & true donot match!'
        assert feat.shape[-2:] == pred.shape[-2:], 'shape of feat & pred donot match!'
        rows = feat.permute(0, 2, 3, 1).reshape(-1, feat.shape[1])
        return self.func(rows, pred, true.float().round(), card=card_select * true.shape[0])
