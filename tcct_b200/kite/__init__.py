from .loopback import KiteBack, setup_seed  # noqa: F401
from .loop_seg import KiteSeg, argmax_labels  # noqa: F401
from .losses import get_loss, MultiLoss, DiceLoss  # noqa: F401
