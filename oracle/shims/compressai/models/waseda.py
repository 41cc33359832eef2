from oracle.compressai_port import Cheng2020Anchor  # noqa: F401
