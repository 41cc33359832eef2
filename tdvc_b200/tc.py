"""Host-side packing for the tcgen05 (3xBF16 split) kernels: fp32 weights -> (hi, lo) bf16 operand tiles.
Filled in together with csrc/conv_tc.cu."""


def attach_bf16(packed):
    return packed
