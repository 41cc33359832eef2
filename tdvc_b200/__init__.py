"""tdvc_b200 — B200-native (sm_100a) implementation of TDVC's P-frame coding forward pass.

Public surface (mirrors the reference):
  * tdvc_b200.VideoCompressor  — drop-in for reference main/model/pnet.py::VideoCompressor
  * tdvc_b200.dcn_v2_forward   — drop-in for the reference's `_ext.dcn_v2_forward`
  * tdvc_b200.gop              — GOP evaluation driver mirroring reference tools/predict.py:43-100
The compute path is hand-written CUDA behind the C-ABI in include/tdvc_b200.h; there is no CPU fallback.
"""
__all__ = ["VideoCompressor", "dcn_v2_forward"]


def __getattr__(name):  # lazy: importing the package must not need the GPU library
    if name == "VideoCompressor":
        from tdvc_b200.model import VideoCompressor
        return VideoCompressor
    if name == "dcn_v2_forward":
        from tdvc_b200.ops import dcn_v2_forward
        return dcn_v2_forward
    raise AttributeError(name)
