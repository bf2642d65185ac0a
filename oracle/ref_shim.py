"""Import the UNMODIFIED reference (/root/reference) on CPU.

TEST INFRASTRUCTURE ONLY.  Used in the build container to (a) pin the numpy
restatement in roreg_oracle.py against the reference's own code and (b) generate
the golden fixtures under tests/golden/.  /root/reference does not exist on the
GPU box, so nothing imported at bench / smoke / `-m gpu` time may import this.

The four shims are the ones SURVEY.md section 8(c) lists:
  1. np.int / np.float aliases (removed in NumPy >= 1.24)
  2. .cuda() -> identity on Tensor and Module (constructors call it unconditionally)
  3. stub `open3d`   (imported by test/extractor.py:6, never used on the path)
  4. stub `tensorboardX` (utils/utils.py:7)
plus cwd=/root/reference (rot_coh_match.py:322 hard-codes ./utils/group_related)
and the reference root first on sys.path (its package `test` shadows the stdlib one).
"""
import os
import sys
import types

REF_ROOT = os.environ.get("ROREG_REFERENCE_ROOT", "/root/reference")


def available():
    return os.path.isdir(os.path.join(REF_ROOT, "network"))


_installed = False


def install():
    global _installed
    if _installed:
        return
    if not available():
        raise RuntimeError(f"reference not found at {REF_ROOT}")
    import numpy as np
    import torch

    if not hasattr(np, "int"):
        np.int = int
    if not hasattr(np, "float"):
        np.float = float
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.nn.Module.cuda = lambda self, *a, **k: self
    _orig_load = torch.load

    def _load(f, *a, **k):
        k.setdefault("map_location", "cpu")
        k.setdefault("weights_only", False)
        return _orig_load(f, *a, **k)

    torch.load = _load
    o3d = types.ModuleType("open3d")
    o3d.io = types.SimpleNamespace(read_point_cloud=None)
    sys.modules.setdefault("open3d", o3d)
    tbx = types.ModuleType("tensorboardX")
    tbx.SummaryWriter = type("SummaryWriter", (), {"__init__": lambda self, *a, **k: None})
    sys.modules.setdefault("tensorboardX", tbx)
    # the reference's `test` package must win over CPython's stdlib `test`
    for m in [m for m in sys.modules if m == "test" or m.startswith("test.")]:
        del sys.modules[m]
    sys.path.insert(0, REF_ROOT)
    os.chdir(REF_ROOT)
    _installed = True


def cfg(**over):
    """A cfg namespace with the defaults of parses/parses_test.py:24-54."""
    c = types.SimpleNamespace(
        base_dir="./data", origin_data_dir="./data/origin_data", backbone="FCGF",
        output_cache_fn="./data/YOHO_FCGF/Testset", model_fn=f"{REF_ROOT}/checkpoints/FCGF",
        SO3_related_files=f"{REF_ROOT}/utils/group_related",
        GF="yoho_des", RD=False, RM=False, ET="yohoc", testset="3dmatch", keynum=5000,
        max_iter=1000, ransac_ird=0.1, tau_1=0.05, tau_2=0.1, tau_3=0.2, match_n=0.5,
        bs_GF=1250, bs_ET=1000)
    for k, v in over.items():
        setattr(c, k, v)
    return c
