"""KiteBack -- drop-in for task1/kite/loopback.py: seeds, checkpoint files, deep-supervision loss sum,
optimizer / LR schedule and device set-up of the training loop.  Same method names and on-disk formats
(`<root>/los.pt`, `<root>/val_top.pt` state dicts, `<root>/params.tar` = {epoch, loss, lr}); the optimizer is
the fused flat AdamW of kite/optims.py and, when launched under torchrun, gradients are averaged with one
NCCL all-reduce over the flat gradient buffer (the reference is single-GPU: loopback.py:130-139)."""
import glob
import os
import random

import numpy as np
import torch
from torch.optim import lr_scheduler

from .. import ops as O
from . import ddp
from .losses import get_loss
from .optims import FlatAdamW


def setup_seed(seed):
    """loopback.py:16-26 (the cudnn flags are kept for parity; no cuDNN kernel runs on this path)."""
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    torch.random.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(seed)
    torch.backends.cudnn.benchmark = False
    torch.backends.cudnn.deterministic = True
    torch.backends.cudnn.enabled = True


class KiteBack(object):
    lossItem = 0
    device = torch.device('cpu')
    coff_ds = 0.5

    def __init__(self, model, dataset, root=None, **args):
        super().__init__()
        print('*' * 32, 'Keras Backend Information!')
        self.model = model
        self.root = root if root else 'exp_tcct-bp'       # the reference hard-codes the folder (loopback.py:36)
        os.makedirs(self.root, exist_ok=True)
        self.dataset = dataset
        print('\tFolder for experiment:', self.root)
        if not self.weights_load('los'):
            print('weights_init_kaiming')
        print('\tParams model:', sum(p.numel() for p in self.model.parameters() if p.requires_grad))

    def cuda(self, m):
        return m.to(self.device, non_blocking=True)

    def islrLowerThan(self, thresh=1e-5):
        return self.optimG.param_groups[0]['lr'] < thresh

    def grad_dump(self, epoch):
        torch.save({'epoch': epoch, 'loss': self.lossName, 'lr': self.optimG.param_groups[0]['lr']}, self.checkpoint_grad)

    def grad_calc(self, outs, true, ds=True, criterion=None):
        """Deep supervision (loopback.py:62-73): sum_{i=last..1} coff_ds * crit(outs[i]) + crit(outs[0])."""
        criterion = criterion or self.criterion
        losSum = 0
        if (isinstance(outs, (list, tuple)) and len(outs) == 4 and ds and outs[0].is_cuda
                and any(o.shape[-2:] != outs[0].shape[-2:] for o in outs[1:])):
            # the model left its auxiliary logits at native resolution (FTC.defer_aux): the fused kernel up-samples in registers
            if getattr(getattr(criterion, 'losses', None), 'mode', 1) == 0:
                lab = criterion.labels(true, outs[0].shape[1])
                return O.DiceMultiFn.apply(outs[0].contiguous(), outs[1].contiguous(), outs[2].contiguous(), outs[3].contiguous(),
                                           lab, float(self.args.coff_ds))[0]
            H, W = outs[0].shape[-2:]
            outs = [outs[0]] + [O.ResizeNCHWFn.apply(o, H, W) for o in outs[1:]]
        if isinstance(outs, (list, tuple)):
            if ds:
                for i in range(len(outs) - 1, 0, -1):
                    losSum = losSum + criterion(outs[i], true) * self.args.coff_ds
            outs = outs[0]
        return losSum + criterion(outs, true)

    def weights_load(self, mode, desc=True):
        path = mode if mode.endswith('.pt') else os.path.join(self.root, mode + '.pt')
        try:
            pt = torch.load(path, map_location='cpu')
            self.model.load_state_dict(pt, strict=False)
            if desc:
                print('\nLoad weight:', path)
            return True
        except Exception:
            print('\nLoad weight wrong:', path)
            return False

    def weights_desc(self, key='my'):
        for n, m in self.model.named_parameters():
            if key in n:
                print(n, m.detach().cpu().numpy())

    def remove_pths(self, flag_ignore='los'):
        for path in glob.glob(self.root + '/*.pt'):
            if flag_ignore not in path:
                os.remove(path)

    def set_superes(self, loss='ce', lr=0.01, wd=2e-4, **args):
        """loopback.py:102-128: resume {epoch, loss, lr} from params.tar; AdamW(lr, wd=2e-4) + CyclicLR(1e-6..1e-4,
        up 4 / down 60, stepped once per EPOCH).  The optimizer itself is built in set_backend, once the
        parameters live in the flat device buffer."""
        print('Setting super parameters!!!')
        self.checkpoint_grad = os.path.join(self.root, 'params.tar')
        epoch = 0
        if os.path.isfile(self.checkpoint_grad):
            try:
                tar = torch.load(self.checkpoint_grad)
                epoch, loss, lr = tar['epoch'], tar['loss'], tar['lr']
                print('Load Super params for Gradients, lr:{}, los:{}'.format(lr, loss))
            except Exception:
                print('Load Super params Failed!, lr:{}, los:{}'.format(lr, loss))
        else:
            print('Init Super params for Gradients, lr:{}, los:{}'.format(lr, loss))
        self.epoch = epoch
        print('$' * 32, 'start-epoch:{}'.format(self.epoch))
        self.lossName = loss
        self._hyper = dict(lr=lr, wd=wd)

    def set_backend(self, gpu='0', parallel=False, **args):
        print('Setting backend for Pytorch!!!')
        if not torch.cuda.is_available():
            raise RuntimeError("tcct_b200 runs on a CUDA device (sm_100a); no CPU path exists")
        _, _, local = ddp.env_world()
        self.device = torch.device('cuda', local)
        torch.cuda.set_device(self.device)
        self.rank, self.world, _ = ddp.init('nccl', self.device)
        print('Using GPU:', self.device, 'world', self.world)
        self.model = self.model.to(self.device)
        self.flat, _ = self.model.flat_state(self.device)
        ddp.broadcast_replica(self.flat.buf, self.model.buffers())   # identical replicas; BN buffers follow rank 0's
        self.criterion = get_loss(self.args.los)
        # parameters of a regulariser that is switched off never see a gradient -> untouched, as in the reference
        n_active = self.flat.n_used
        if not getattr(self.args, 'reg', False) and 'lap_reg.0.weight' in self.flat.offsets:
            n_active = min(n_active, self.flat.offsets['lap_reg.0.weight'][0])
        self.optimG = FlatAdamW(self.flat, lr=self._hyper['lr'], weight_decay=self._hyper['wd'], max_norm=12.0, n_active=n_active)
        self.optimG.grad_scale = 1.0 / self.world
        self.schedG = lr_scheduler.CyclicLR(self.optimG, base_lr=1e-6, max_lr=1e-4, cycle_momentum=False,
                                            step_size_up=4, step_size_down=60)
        self.optimG.sync_lr()

    def allreduce_grads(self):
        ddp.allreduce_flat(self.flat.grad, self.optimG.n_active)
