from oracle.compressai_port import (GDN, MaskedConv2d, ResidualBlock, ResidualBlockUpsample,  # noqa: F401
                                    ResidualBlockWithStride, conv1x1, conv3x3, subpel_conv3x3)
