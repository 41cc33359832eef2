"""ORACLE shim (test infrastructure): the two mmcv symbols the reference imports
(reference main/model/flownet.py:5, main/model/inflate.py:7-8). mmcv-full 1.x API."""
