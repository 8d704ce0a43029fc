"""Golden vectors of the `cnnu` factory (task1/nets/tcct.py:1124-1129) from the UNMODIFIED reference: eval logits / labels and
the logits + a few gradients of one train-mode forward/backward with a Dice loss on head 0.   python oracle/make_golden_cnnu.py"""
import contextlib, io, os, sys
import numpy as np
import torch
import torch.nn.functional as F
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE)); sys.path.insert(0, HERE)
import refshim
refshim.install()
import nets
from tcct_b200.synth import make_bscans, synth_state
import tcct_oracle as O

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
torch.set_num_threads(8)
n_class, n_bound, batch, height, width, seed = 5, 4, 2, 64, 64, 21
img, lab = make_bscans(batch, height, width, n_class, n_bound, seed)
with contextlib.redirect_stdout(io.StringIO()):
    ref_full = nets.RegNet(nets.stc_tt(n_class), out_channels=n_class)      # same keys: the golden state builder of the tests
    model = nets.cnnu(n_class)
state = synth_state(ref_full.state_dict(), seed)
model.load_state_dict({k[5:]: v for k, v in state.items() if k.startswith("base.")}, strict=True)
model.eval()
with torch.no_grad():
    out0 = model(img)[0]
labels = torch.argmax(F.softmax(out0, 1), 1)
P = {k: v.clone() for k, v in state.items()}
o_out0, o_lab = O.predict_labels(P, img, flag_vit=False)
print("eval: oracle vs reference max|d| %.3e (max|ref| %.3e), label flips %d" % (float((o_out0 - out0).abs().max()), float(out0.abs().max()),
                                                                               int((o_lab != labels).sum())))
# train-mode forward/backward, loss = Dice(head 0) + sum of the aux heads' means (touches every decoder branch)
model.train()
outs = model(img)
onehot = F.one_hot(lab, n_class).permute(0, 3, 1, 2)
loss = O.multi_dice(outs[0], onehot) + sum(o.mean() for o in outs[1:])
loss.backward()
named = dict(model.named_parameters())
keys = ["base_cnn.path_estan.0.block12.0.weight", "base_cnn.path_estan.3.block34.1.weight", "dec2.prep.0.weight", "t323.bias", "aux0.weight"]
assert all(named[k].grad is not None for k in keys)
assert named["tran_cnn0.0.weight"].grad is None and named["base_vit.stem.0.conv.weight"].grad is None
Pt = {k: v.clone().requires_grad_(v.dtype.is_floating_point) for k, v in state.items()}
o_outs, _ = O.ftc_forward(Pt, img, O.Ctx(True), flag_vit=False)
o_loss = O.multi_dice(o_outs[0], onehot) + sum(o.mean() for o in o_outs[1:])
o_loss.backward()
for k in keys:
    d = float((Pt["base." + k].grad - named[k].grad).abs().max()) / float(named[k].grad.abs().max())
    print("train grad %-45s oracle vs reference rel %.2e" % (k, d))
    assert d < 1e-3
print("train loss", float(loss), float(o_loss))
np.savez_compressed(os.path.join(OUT, "cnnu_goals_64.npz"), meta=np.array([n_class, n_bound, batch, height, width, seed], np.int64),
                    out0=out0.numpy(), labels=labels.numpy().astype(np.uint8), train_out0=outs[0].detach().numpy(),
                    train_loss=np.float64(float(loss)),
                    **{"grad::" + k: named[k].grad.numpy() for k in keys},
                    vit_running_mean=model.state_dict()["base_vit.stem.1.bn.running_mean"].numpy())
