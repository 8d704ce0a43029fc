"""`python -m tcct_b200.kite.main --bs=8 --net=stc_tt --los=di --epochs=100 --db=goals` -- the flag set of
task1/kite/main.py:18-49 (dead flags of the reference are accepted and ignored the same way), plus
`--graph` (CUDA-graph replay of the step, default on), `--height/--width/--batches` for the synthetic
dataset that stands in for the reference's OpenCV loader (out of scope)."""
import argparse
import os

import torch

from ..nets import RegNet
from .. import nets as _nets
from ..synth import SynthOCT
from . import ddp
from .loop_seg import KiteSeg


def str2bool(v):
    if v.lower() in ('yes', 'true', 't', 'y', '1'):
        return True
    if v.lower() in ('no', 'false', 'f', 'n', '0'):
        return False
    raise argparse.ArgumentTypeError('Unsupported value encountered.')


def build_parser():
    p = argparse.ArgumentParser(description="KiteOCT Argument")
    p.add_argument('--db', type=str, default='duke1', choices=['duke', 'duke1', 'duke2', 'duke3', 'hcms', 'hcms1', 'heg', 'goals', 'odsgh'])
    p.add_argument('--lr', type=float, default=1e-2)
    p.add_argument('--wd', type=float, default=5e-4)
    p.add_argument('--inc', type=str, default='')
    p.add_argument('--gpu', type=str, default='0')
    p.add_argument('--los', type=str, default='dice')
    p.add_argument('--net', type=str, default='stc_tt')
    p.add_argument('--pth', type=str2bool, default=True)
    p.add_argument('--bs', type=int, default=2)
    p.add_argument('--epochs', type=int, default=100)
    p.add_argument('--root', type=str, default='')
    p.add_argument('--resume', type=str2bool, default=False)
    p.add_argument('--reg', type=str2bool, default=False)
    p.add_argument('--coff_reg', type=float, default=.1)
    p.add_argument('--epl', type=str2bool, default=False)
    p.add_argument('--coff_epl', type=float, default=.1)
    p.add_argument('--udh', type=str2bool, default=False)
    p.add_argument('--coff_udh', type=float, default=1)
    p.add_argument('--type_udh', type=str, default='cos', choices=['cos', 'mse'])
    p.add_argument('--ds', type=str2bool, default=False)
    p.add_argument('--coff_ds', type=float, default=1)
    p.add_argument('--pl', type=str2bool, default=False)
    p.add_argument('--bug', type=str2bool, default=False)
    # additions (new names only)
    p.add_argument('--graph', type=str2bool, default=True)
    p.add_argument('--height', type=int, default=256)
    p.add_argument('--width', type=int, default=256)
    p.add_argument('--batches', type=int, default=16)
    return p


def main(argv=None):
    args = build_parser().parse_args(argv)
    db = {'duke1': 'duke', 'duke2': 'duke', 'duke3': 'duke', 'hcms1': 'hcms', 'odsgh': 'goals'}.get(args.db, args.db)
    # data parallel (torchrun): every rank draws its own batches (weak scaling); without this all ranks would train on identical data
    rank = int(os.environ.get('RANK', 0))
    dataset = SynthOCT(db, args.height, args.width, n_batches=args.batches, seed=ddp.shard_seed(1234, 1000 * rank))
    factory = getattr(_nets, args.net, None)
    if factory is None:
        raise SystemExit("unknown --net %s" % args.net)
    net = RegNet(factory(dataset.out_channels), con=args.type_udh, out_channels=dataset.out_channels)
    print('OUT-CHANNELS:', dataset.out_channels)
    keras = KiteSeg(model=net, dataset=dataset, root=args.root, args=args)
    if args.resume:
        path = args.root + '/val_top.pt'
        keras.model.load_state_dict(torch.load(path, map_location='cpu'), strict=False)
        print('loaded model:', path)
    keras.fit(epochs=1 if args.bug else args.epochs)


if __name__ == '__main__':
    main()
