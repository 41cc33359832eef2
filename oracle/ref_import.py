"""ORACLE (test infrastructure): import the reference's OWN, UNMODIFIED model code from /root/reference
on top of the three shims in oracle/shims/ (mmcv, compressai, _ext) — SURVEY.md 8c.

Only usable in the build container (/root/reference does not exist on the GPU box).  It is used
 * by tests/test_oracle.py to pin oracle/model.py (our restatement) against the reference code, and
 * by oracle/make_golden.py to generate the fixtures under tests/golden/.
Nothing is copied: the reference files are imported from where they lie.
"""
import importlib
import os
import sys

REFERENCE_ROOT = os.environ.get("TDVC_REFERENCE_ROOT", "/root/reference")
_SHIMS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "shims")
_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "main", "model", "pnet.py"))


def load_reference_pnet():
    """Returns the reference module `main.model.pnet` (reference main/model/pnet.py), imported verbatim.

    The only intervention: SPyNet is constructed with `pretrained=None` semantics, because the reference
    downloads weights at construction (reference main/model/pnet.py:126-127, flownet.py:68-70) and the
    container has no network.
    """
    if not available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    for p in (_REPO, _SHIMS, REFERENCE_ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)
    flownet = importlib.import_module("main.model.flownet")
    if not getattr(flownet.SPyNet, "_tdvc_no_download", False):
        orig_init = flownet.SPyNet.__init__

        def init_no_download(self, pretrained=None):
            orig_init(self, None)

        flownet.SPyNet.__init__ = init_no_download
        flownet.SPyNet._tdvc_no_download = True
    return importlib.import_module("main.model.pnet")


def reference_video_compressor():
    return load_reference_pnet().VideoCompressor()


def load_reference_dataset():
    """Returns the reference module `main.dataloader.dataset` (reference main/dataloader/dataset.py), imported verbatim, to pin the
    SAMPLE LISTS of tdvc_b200.data against the reference's own code.  Its imports that are absent here are stubbed: `cv2` and
    `albumentations` (used only when an item is read), `natsort` (natsorted = sort by the digit runs as numbers) and
    `main.model.basics` (drags in the unused TensorFlow-era helpers; only `CalcuPSNR` is referenced, at item-read time)."""
    import re
    import types
    if not available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    for p in (_REPO, _SHIMS, REFERENCE_ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)

    def stub(name, **attrs):
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.__dict__.update(attrs)
            sys.modules[name] = m
        return sys.modules[name]

    key = lambda s: [int(t) if t.isdigit() else t for t in re.split(r"(\d+)", str(s))]
    stub("cv2")
    stub("albumentations")
    stub("natsort", natsorted=lambda seq: sorted(seq, key=key))
    importlib.import_module("main.model")
    stub("main.model.basics", CalcuPSNR=None)
    return importlib.import_module("main.dataloader.dataset")
