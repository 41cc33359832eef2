#!/bin/bash
python tools/tma_check.py 2>&1 | tail -9
timeout 1500 python -m pytest tests -q -m gpu --maxfail=8 2>&1 | tail -4
TDVC_B200_CONV_TMA=1 timeout 900 python -m pytest tests -q -m gpu --maxfail=8 2>&1 | tail -4
python tools/ab_frame.py env:TDVC_B200_CONV_LDG 4 44 2>&1 | tail -2
