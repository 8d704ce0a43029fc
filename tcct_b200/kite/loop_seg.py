"""KiteSeg -- drop-in for task1/kite/loop_seg.py: fit / train / val / predict / calc_loss with the reference's
flags (`args.los, lr, bs, udh, coff_udh, reg, coff_reg, epl, coff_epl, coff_ds, bug`).

What differs from the reference's loop is only *how* a step is issued: after a few eager steps the whole
iteration (forward, Dice x4 + feature-polarisation + boundary-regression losses, backward, gradient clipping,
AdamW) is captured once per input shape in a CUDA graph and replayed; losses stay on the device and are read
back once per log interval instead of 3-4 `.item()` syncs per step (loop_seg.py:134,152-169)."""
import time

import numpy as np
import torch
import torch.nn.functional as F

from .. import _lib as L
from .. import ops as O
from ..ops import _check, _p, _stream, labels_u8
from . import ddp
from .loopback import KiteBack, setup_seed
from .losses.miou import MDiceLoss, MIouLoss, label_counts


def argmax_labels(logits):
    """uint8 [B,H,W] label map = argmax over classes of NCHW logits (first maximum wins, like torch.argmax)."""
    logits = logits.contiguous()
    _check(logits)
    B, C, H, W = logits.shape
    lab = torch.empty((B, H, W), dtype=torch.uint8, device=logits.device)
    L.argmax_nchw(_p(logits), _p(lab), B, C, H * W, _stream())
    return lab


class _Graphed:
    """One captured train step for one (image shape, label shape)."""

    def __init__(self, seg, img, lab8):
        self.img = torch.empty_like(img)
        self.lab = torch.empty_like(lab8)
        self.parts = torch.zeros(4, dtype=torch.float32, device=img.device)      # los, udh, reg, total of the last step
        self.graph = None
        self.warm = 0
        self.seg = seg

    def body(self):
        seg = self.seg
        early = {}
        total, parts = seg._losses(self.img, self.lab, early=early)
        if early:       # boundary regression already ran its backward: feed its logit gradient in next to the other losses
            torch.autograd.backward([early['rest'], early['out0']], [None, early['grad']])
        else:
            total.backward()
        vals = [parts.get('los'), parts.get('udh'), parts.get('reg'), total]
        self.parts.copy_(torch.stack([v.detach().float() if v is not None else torch.zeros((), device=self.img.device) for v in vals]))

    def step(self, img, lab8):
        seg = self.seg
        self.img.copy_(img, non_blocking=True)
        self.lab.copy_(lab8, non_blocking=True)
        fused = seg.world == 1          # the gradient all-reduce stays outside the graph
        if not seg.use_graph:
            seg.optimG.zero_grad()
            self.body()
            seg.allreduce_grads()
            seg.optimG.step()
            return
        if self.graph is None and self.warm < seg.GRAPH_WARMUP:
            # eager warm-up on a side stream (torch's CUDA-graph recipe: autograd's lazily created per-parameter
            # accumulators must not be bound to the legacy default stream when the capture runs)
            self.warm += 1
            cur = torch.cuda.current_stream()
            side = seg.side_stream
            side.wait_stream(cur)
            with torch.cuda.stream(side):
                seg.optimG.zero_grad()
                self.body()
                seg.allreduce_grads()
                seg.optimG.step()
            cur.wait_stream(side)
            return
        if self.graph is None:
            seg.release_graph_refs()    # drop the previous step's autograd graph before capturing a new one
            torch.cuda.synchronize()
            self.graph = torch.cuda.CUDAGraph()
            seg.optimG.zero_grad()
            with torch.cuda.graph(self.graph, stream=seg.side_stream):
                seg.flat.grad.zero_()
                self.body()
                if fused:
                    seg.optimG.step()
        self.graph.replay()
        if not fused:
            seg.allreduce_grads()
            seg.optimG.step()


class KiteSeg(KiteBack):
    GRAPH_WARMUP = 3
    use_graph = True
    log_every = 16
    useValSet = True
    cnt_val = 0
    udh_out = None
    udh_lab = None

    def __init__(self, args, **_args):
        self.args = args
        super().__init__(**_args)
        self.set_superes(loss=args.los, lr=args.lr)          # args.wd is not forwarded (loop_seg.py:14): wd stays 2e-4
        self.set_backend(gpu=args.gpu, parallel=args.pl)
        self.NB_CLASS = self.dataset.out_channels
        self.criterion.NB_CLASS = self.NB_CLASS
        self.use_graph = bool(getattr(args, 'graph', True))
        self._graphs = {}
        self.best_dice = -1.0
        self.side_stream = torch.cuda.Stream(device=self.device, priority=O.CHAIN_PRIORITY)      # warm-up and capture stream

    def release_graph_refs(self):
        """Forget tensors that keep the last step's autograd graph (and its per-parameter accumulators) alive."""
        self.udh_out = self.udh_lab = None
        base = getattr(self.model, 'base', self.model)
        base.feats = None
        base.feats_nhwc = None

    # ------------------------------------------------------------------ inference
    def predict(self, img, softmax=True, *args):
        """loop_seg.py:21-33: eval forward, head-0 logits; softmax=True -> one-hot float argmax map."""
        with torch.no_grad():
            pred = self.model(self.cuda(img))
            if isinstance(pred, (list, tuple)):
                pred = pred[0]
            pred = pred.detach()
            if softmax:
                pred = F.one_hot(argmax_labels(pred).long(), self.NB_CLASS).permute(0, 3, 1, 2).float()
        return pred

    def predict_labels(self, img):
        """uint8 [B,H,W] label map (the same argmax without materialising the one-hot)."""
        with torch.no_grad():
            pred = self.model(self.cuda(img))
            return argmax_labels(pred[0] if isinstance(pred, (list, tuple)) else pred)

    # ------------------------------------------------------------------ training
    def fit(self, epochs=169):
        print('\n', '*' * 8, 'Fitting:' + self.root)
        t0 = time.time()
        for i in range(self.epoch, epochs):
            ts = time.time()
            self.train(i)
            self.schedG.step()
            self.optimG.sync_lr()
            if i % 10 == 0 or (i > 0.5 * epochs and i % 5 == 0):
                ddp.average_buffers(self.model)
                logs = self.val(epoch=i)
                if logs['val_f1s'] > self.best_dice:             # reference: undefined best_dice/log/static_dict (loop_seg.py:53-55)
                    self.best_dice = logs['val_f1s']
                    if self.rank == 0:
                        torch.save(self.model.state_dict(), self.root + '/val_top.pt')
            if self.rank == 0:
                self.grad_dump(i)
            te = time.time() - ts
            print('{:03}* {:.2f} mins, left {:.2f} hours to run'.format(i, te / 60, te / 60 / 60 * (epochs - i)))
        print('\nRunning {:.2f} hours for {} epochs!'.format((time.time() - t0) / 60 / 60, epochs))
        self.weights_desc()

    def val(self, epoch=0, flagDebug=False):
        """loop_seg.py:66-106: per-image Dice / IoU of the argmax map, classes 1.. averaged."""
        was = torch.is_grad_enabled()
        torch.set_grad_enabled(False)
        self.model.eval()
        counts = []
        for i, imgs in enumerate(self.dataset.valSet(bs=1)):
            (img, lab, fov, aux) = self.dataset.parse(imgs)
            pred8 = self.predict_labels(img)
            true8 = labels_u8(self.cuda(lab).reshape(pred8.shape).contiguous(), self.NB_CLASS)
            counts.append(label_counts(pred8, true8, self.NB_CLASS))
            if (self.args.bug or flagDebug) and i > 8:
                break
        counts = torch.cat(counts).cpu()                         # one read-back for the whole validation set
        f1s, ious, scores = [], [], []
        for c in counts:
            f1, per_class = MDiceLoss.from_counts(c[None], start_idx=1)
            iou, _ = MIouLoss.from_counts(c[None], start_idx=1)
            f1s.append(float(f1)); ious.append(float(iou)); scores.append(per_class.numpy().astype(np.float32))
        n = max(len(f1s), 1)
        logs = {'val_iou': sum(ious) / n, 'val_f1s': sum(f1s) / n}
        scores = np.round(np.stack(scores, axis=0).mean(axis=0), 4)
        print('Val@{:03} iou={:.4f} & f1s={:.4f}'.format(epoch, logs['val_iou'], logs['val_f1s']))
        print('*SCORES:*', scores, '->', scores[1:].mean())
        torch.set_grad_enabled(was)
        return logs

    def _label_map(self, lab):
        lab = self.cuda(lab)
        if lab.dim() == 4 and lab.shape[1] == 1:
            lab = lab[:, 0]
        return labels_u8(lab.contiguous(), self.NB_CLASS)

    def train(self, epoch, alpha=.9):
        setup_seed(epoch * 311 + 2023 + 7919 * self.rank)       # rank 0 draws the reference's stream; the other replicas their own noise / DropPath
        torch.set_grad_enabled(True)
        self.model.train()
        self.optimG.sync_lr()
        acc = torch.zeros(4, dtype=torch.float32, device=self.device)
        n = 0
        for i, imgs in enumerate(self.dataset.trainSet(bs=self.args.bs)):
            (img, lab, fov, aux) = self.dataset.parse(imgs)
            parts = self.train_step(img, lab)
            acc += parts
            n += 1
            if self.log_every and (i + 1) % self.log_every == 0:
                p = parts.tolist()
                print('\r#{:03} los={:.4f},udh={:.4f},reg={:.4f}'.format(i, p[0], p[1], p[2]), end='')
            if self.args.bug and i > 12:
                break
        losItem = float(acc[3])
        print('\r{:03}# {}={:.4f},'.format(epoch, self.lossName, losItem), end='')
        return losItem

    def prefetch(self, img, lab):
        """Start the host -> device copy of the NEXT batch on a copy stream, so that it overlaps the step in flight (what a
        data loader with pinned buffers does).  `train_step(img, lab)` with the same two host tensors then picks the device
        copies up instead of copying again; any other batch silently takes the ordinary path.  Two persistent staging
        sets alternate (no allocation, no allocator bookkeeping across streams in the steady state)."""
        if not (torch.is_tensor(img) and torch.is_tensor(lab)) or img.is_cuda or self.device.type != 'cuda':
            return
        st8 = self.__dict__.setdefault('_pf', {'stream': None, 'sets': [None, None], 'free': [None, None], 'k': 0})
        if st8['stream'] is None:
            st8['stream'] = torch.cuda.Stream(device=self.device)
        st = st8['stream']
        k = st8['k'] = st8['k'] ^ 1
        cur = st8['sets'][k]
        if cur is None or cur[0].shape != img.shape or cur[0].dtype != img.dtype or cur[1].shape != lab.shape or cur[1].dtype != lab.dtype:
            cur = st8['sets'][k] = (torch.empty(img.shape, dtype=img.dtype, device=self.device),
                                    torch.empty(lab.shape, dtype=lab.dtype, device=self.device))
            st.wait_stream(torch.cuda.current_stream(self.device))
        if st8['free'][k] is not None:
            st.wait_event(st8['free'][k])           # the step that read this staging set two steps ago has consumed it
        with torch.cuda.stream(st):
            cur[0].copy_(img, non_blocking=True)
            cur[1].copy_(lab, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(st)
        self._prefetched = (img, lab, cur[0], cur[1], ev, k)

    def train_step(self, img, lab):
        """One optimisation step on a batch; returns the device tensor [los, udh, reg, total]."""
        pf = self.__dict__.get('_prefetched')
        used = None
        if pf is not None and pf[0] is img and pf[1] is lab:
            torch.cuda.current_stream(self.device).wait_event(pf[4])
            img, lab, used = pf[2], pf[3], pf[5]
        self._prefetched = None
        img = self.cuda(img).float()
        lab8 = self._label_map(lab)
        key = (tuple(img.shape), tuple(lab8.shape))
        g = self._graphs.get(key)
        if g is None:
            g = self._graphs[key] = _Graphed(self, img, lab8)
        g.step(img, lab8)
        if used is not None:        # everything that reads the staging set is enqueued: it may be overwritten after this point
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(self.device))
            self._pf['free'][used] = ev
        return g.parts

    def _losses(self, img, lab, early=None):
        """loop_seg.py:146-171 without the per-term host syncs: returns (total, {name: device scalar}).

        `early` (a dict, train step only): the boundary-regression term runs its backward right behind its forward, on its
        own stream (its graph ends at the logits: a detached leaf stands in for them), instead of waiting for the slower
        feature-polarisation forward to finish before any backward can start.  The dict then holds the logits, their
        gradient from that term and `rest`, the sum of the other terms, for one joint backward call."""
        if getattr(self.args, 'epl', False):
            raise AttributeError("--epl=1: the reference's RegNet has no regular_epl (loop_seg.py:166-169)")
        # the deep-supervision Dice kernel reads the auxiliary logits at native resolution (ops.DiceMultiFn): ask the model not to
        # up-sample them (FTC.defer_aux; any other criterion gets them up-sampled in grad_calc)
        base = getattr(self.model, 'base', self.model)
        defer = hasattr(base, 'defer_aux') and img.is_cuda
        if defer:
            base.defer_aux = True
        try:
            out = self.model(img)
        finally:
            if defer:
                base.defer_aux = False
        out0 = out[0] if isinstance(out, (list, tuple)) else out
        self.udh_out, self.udh_lab = out0.detach(), lab
        # The three loss families are independent chains of small (latency-bound) kernels between the forward and the
        # backward of the network, when nothing else is in flight: feature polarisation and boundary regression run on
        # side streams next to the Dice terms (their backward nodes replay on the same streams).
        dev = out0.device
        s_udh = O.fork(dev, 2) if (self.args.udh and out0.is_cuda) else None
        s_reg = O.fork(dev, 3) if (self.args.reg and out0.is_cuda) else None
        parts = {'los': self.grad_calc(out, lab, ds=True, criterion=self.criterion)}
        if self.args.udh:
            with O.on(s_udh):
                if s_udh is not None:
                    for t in (out0, lab, getattr(getattr(self.model, 'base', None), 'feats_nhwc', None)):
                        if torch.is_tensor(t):
                            t.record_stream(s_udh)
                parts['udh'] = self.model.regular_udh(out0, lab) * self.args.coff_udh
            O.join(s_udh, parts['udh'])
        if self.args.reg:
            with O.on(s_reg):
                if s_reg is not None:
                    for t in (out0, lab):
                        t.record_stream(s_reg)
                if early is not None and out0.requires_grad and torch.is_grad_enabled():
                    leaf = out0.detach().requires_grad_(True)
                    reg = self.model.regular_reg(leaf, lab) * self.args.coff_reg
                    reg.backward()
                    parts['reg'] = reg.detach()
                    early['out0'], early['grad'] = out0, leaf.grad
                else:
                    parts['reg'] = self.model.regular_reg(out0, lab) * self.args.coff_reg
            O.join(s_reg, parts['reg'], early.get('grad') if early else None)
        if early:
            early['rest'] = sum(v for v in parts.values() if v.requires_grad)
        return sum(parts.values()), parts

    def calc_loss(self, img, lab):
        """Reference signature: (losSum, logStr).  Formatting the log string reads the losses back (one sync)."""
        total, parts = self._losses(img, lab)
        logStr = ','.join('{}={:.4f}'.format(k, float(v)) for k, v in parts.items())
        return total, logStr
