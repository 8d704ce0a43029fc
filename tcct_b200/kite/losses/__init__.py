from .loss import *  # noqa: F401,F403
from .loss import DiceLoss, MultiLoss, get_loss  # noqa: F401
