"""`NetWork` -- the inference wrapper of task1/onnx/onnx_infer.py:13-31 on the B200 kernels.

The reference exports the trained model with torch.onnx.export (onnx/onnx_save.py:4-15: one input "input" [batch, 3, H, W], one output
"output" = head-0 logits, dynamic batch / height / width) and runs it with onnxruntime: `NetWork(onnx_file).forward(img)` takes an
HWC uint8 image, feeds `img.transpose(2, 0, 1)[None] / 255` and returns the squeezed output array.  An ONNX graph cannot carry this
repository's operators (they are C-ABI kernels, and neither onnx nor onnxruntime is in the image), so the wrapper keeps the CALL
contract and takes the checkpoint the ONNX file was exported from (`tcct_duke.pt`, `tcct_goals.pt`, ...) instead of the graph:
same input convention, same output (float32 logits [C, H, W]), dynamic H / W (multiples of 16), batches through `forward_batch`."""
import contextlib
import io

import numpy as np
import torch


class NetWork:
    def __init__(self, onnx_file=r"tcct_duke.pt", n_class=None, variant=None, device="cuda:0"):
        from ..nets import RegNet, stc_tt, stc_tt_onnx
        state = onnx_file if isinstance(onnx_file, dict) else torch.load(onnx_file, map_location="cpu")
        if n_class is None:
            n_class = int(state["base.aux0.weight"].shape[0])
        if variant is None:       # the older decoder of onnx/tcct_{goals,hcms,heg}.py has no t321-t324 projections
            variant = "tcct" if "base.t324.weight" in state else "onnx"
        with contextlib.redirect_stdout(io.StringIO()):
            net = RegNet((stc_tt if variant == "tcct" else stc_tt_onnx)(n_class), out_channels=n_class)
        own = net.state_dict()
        # like the reference's load_state_dict(strict=False): RegNet extras of other shapes (older lap_reg / lap_map) are skipped
        net.load_state_dict({k: v for k, v in state.items() if k in own and own[k].shape == v.shape}, strict=False)
        self.device = torch.device(device)
        self.net = net.to(self.device).eval()
        self.n_class = n_class

    tmp = {}

    def forward_batch(self, imgs):
        """imgs: uint8 [B, H, W, 3] (numpy or tensor) -> float32 logits [B, C, H, W] on the host."""
        x = torch.as_tensor(np.ascontiguousarray(imgs) if isinstance(imgs, np.ndarray) else imgs)
        x = x.to(self.device).permute(0, 3, 1, 2).float().div(255)
        with torch.no_grad():
            out = self.net(x.contiguous())[0]
        return out.float().cpu().numpy()

    def forward(self, img):
        """onnx_infer.py:18-31: HWC uint8 image -> squeezed network output (logits [C, H, W])."""
        h, w, c = img.shape
        out = self.forward_batch(np.asarray(img).reshape(1, h, w, c)).squeeze()
        print('shape-output:', out.shape)
        return out
