"""ORACLE shim: mmcv.runner.BaseModule == nn.Module + an ignored init_cfg."""
import torch.nn as nn


class BaseModule(nn.Module):
    def __init__(self, init_cfg=None):
        super().__init__()
        self.init_cfg = init_cfg
