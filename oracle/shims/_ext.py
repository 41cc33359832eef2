"""ORACLE shim for the reference's pybind module `_ext` (reference main/utils/dcnv2/src/vision.cpp:4-9).

The reference's native op needs THC headers that torch 2.11 no longer ships, so it cannot be built here
(SURVEY.md 3.4).  `dcn_v2_forward` is served by torchvision.ops.deform_conv2d, which shares the MXNet
lineage: same offset/mask channel order and the same (-1,H)x(-1,W) validity rule
(reference main/utils/dcnv2/src/cuda/dcn_v2_im2col_cuda.cu:125-195).  oracle/dcn_naive.py restates the
kernel element by element and tests/test_oracle.py checks the two against each other and against the
reference's zero-offset known-answer test (reference main/utils/dcnv2/testcuda.py:36-71).
"""
import torchvision.ops


def dcn_v2_forward(input, weight, bias, offset, mask, kh, kw, sh, sw, ph, pw, dh, dw, deformable_group):
    assert weight.shape[2] == kh and weight.shape[3] == kw
    return torchvision.ops.deform_conv2d(input, offset, weight, bias, stride=(sh, sw), padding=(ph, pw),
                                         dilation=(dh, dw), mask=mask)


def dcn_v2_backward(*a, **k):
    raise NotImplementedError("training rows are 'next' (SURVEY.md 8f)")
