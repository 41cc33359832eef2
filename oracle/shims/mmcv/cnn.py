"""ORACLE shim: mmcv.cnn.ConvModule restricted to what the reference uses
(norm_cfg=None, conv_cfg=None, act_cfg in {None, ReLU, Sigmoid}); sub-module names `conv`, `activate`."""
import torch.nn as nn


class ConvModule(nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1,
                 bias="auto", conv_cfg=None, norm_cfg=None, act_cfg=dict(type="ReLU"), inplace=True, **kw):
        super().__init__()
        assert conv_cfg is None and norm_cfg is None
        self.conv = nn.Conv2d(in_channels, out_channels, kernel_size, stride=stride, padding=padding,
                              dilation=dilation, groups=groups, bias=True if bias == "auto" else bias)
        self.with_activation = act_cfg is not None
        if self.with_activation:
            kind = act_cfg["type"]
            if kind == "ReLU":
                self.activate = nn.ReLU(inplace=inplace)
            elif kind == "Sigmoid":
                self.activate = nn.Sigmoid()
            else:
                raise NotImplementedError(kind)

    def forward(self, x):
        x = self.conv(x)
        if self.with_activation:
            x = self.activate(x)
        return x
