"""FeatConPolar -- drop-in for task1/nets/fcp.py:16-75: fixed per-class target directions
(`vec_grad` frozen parameter, `buf_grad` = its row-normalised copy, `cos_dist`)."""
import torch
import torch.nn as nn
import torch.nn.functional as F


class FeatConPolar(nn.Module):
    def __init__(self, num_cls=8, num_emb=32, init=True):
        super().__init__()
        self.num_cls = num_cls
        self.vec_grad = nn.Parameter(torch.rand(num_cls, num_emb), requires_grad=True)
        self.cls_nums = torch.arange(0, num_cls).long()
        print('FeatConPolar-Number&Length:', num_cls, num_emb)
        n_pairs = num_cls * (num_cls - 1) // 2
        print('Constrain vectors to', -1 / (num_cls - 1))
        self.register_buffer('cos_dist', torch.FloatTensor([-1 / (num_cls - 1)] * n_pairs))
        if init:
            self._spread(333)
        self.vec_grad.requires_grad = False
        self.register_buffer('buf_grad', F.normalize(self.vec_grad, p=2, dim=-1).detach())
        print('vec_grad:', self.vec_grad.min().item(), self.vec_grad.max().item())

    def _spread(self, steps):
        """Optional one-off initialisation (fcp.py:36-57): push the class directions apart by minimising
        regular_target with Adam + ReduceLROnPlateau.  Host-side, runs once at construction."""
        opt = torch.optim.Adam(self.parameters(), lr=1e-2, betas=(0.9, 0.999), weight_decay=2e-4)
        sched = torch.optim.lr_scheduler.ReduceLROnPlateau(opt, mode='min', factor=0.7, patience=2, threshold=0.0001,
                                                           threshold_mode='rel', cooldown=2, min_lr=1e-5, eps=1e-9)
        for i in range(steps):
            opt.zero_grad()
            los = self.regular_target(self.vec_grad)
            los.backward()
            opt.step()
            sched.step(los.item())
            if los.item() < 1e-5:
                break

    def regular_target(self, vec_nd):
        v = F.normalize(vec_nd, dim=-1)
        return torch.log(torch.exp(v @ v.T).mean(dim=-1)).mean()

    def choice(self, pro, i):
        """Target rows for class i: buf_grad[i] repeated pro.shape[0] times."""
        return self.buf_grad[i].unsqueeze(0).expand(pro.shape[0], -1).contiguous()
