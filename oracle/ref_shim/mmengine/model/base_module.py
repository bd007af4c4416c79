"""mmengine BaseModule stand-in: nn.Module carrying an `init_cfg` attribute."""
import torch.nn as nn


class BaseModule(nn.Module):
    def __init__(self, init_cfg=None):
        super().__init__()
        self.init_cfg = init_cfg
