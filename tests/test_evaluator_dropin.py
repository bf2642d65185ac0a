"""INTEGRATION.md's claim, executed (CPU, build container): the reference's own `yoho_evaluator` (test/evaluator.py:13-101),
unmodified, drives a scene once with the reference's plugins and once with `from test import name2...` resolving to
roreg_b200.test through the one-file stub INTEGRATION.md prescribes; the files it leaves and the FMR / IR / RR the reference's
metric code computes from them must be the same.  Each arm runs in its own process (tests/_evaluator_driver.py)."""
import os
import re
import subprocess
import sys
import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
REF = os.environ.get("ROREG_REFERENCE_ROOT", "/root/reference")

pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "test")), reason="the reference tree is not present on this machine")


def _stub_from_integration_md():
    """The replacement test/__init__.py, taken from INTEGRATION.md's code block so that the document and this test cannot drift."""
    md = open(os.path.join(REPO, "INTEGRATION.md")).read()
    block = re.search(r"```python\n(# test/__init__\.py.*?)```", md, re.S).group(1)
    assert "from roreg_b200.test import" in block
    return block.replace('"/path/to/roreg_b200_repo"', repr(REPO))


def _overlay(root):
    """The reference tree with ONE file changed: symlinks to everything, test/__init__.py = the stub."""
    os.makedirs(f"{root}/test")
    for e in os.listdir(REF):
        if e != "test":
            os.symlink(os.path.join(REF, e), f"{root}/{e}")
    for e in os.listdir(f"{REF}/test"):
        if e not in ("__init__.py", "__pycache__"):
            os.symlink(f"{REF}/test/{e}", f"{root}/test/{e}")
    open(f"{root}/test/__init__.py", "w").write(_stub_from_integration_md())


def _run(arm, root, cache, out, rd):
    env = dict(os.environ, ROREG_REFERENCE_ROOT=root, PYTHONDONTWRITEBYTECODE="1")
    r = subprocess.run([sys.executable, os.path.join(HERE, "_evaluator_driver.py"), "--arm", arm, "--cache", cache, "--out", out, "--rd", str(rd)],
                       env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-3000:]
    return np.load(out)


@pytest.mark.parametrize("rd", [0, 1], ids=["default_cli", "RD"])
def test_reference_evaluator_runs_unchanged_on_the_b200_plugins(tmp_path, rd):
    overlay = str(tmp_path / "overlay")
    _overlay(overlay)
    a = _run("ref", REF, str(tmp_path / "cache_ref"), str(tmp_path / "ref.npz"), rd)
    b = _run("b200", overlay, str(tmp_path / "cache_b200"), str(tmp_path / "b200.npz"), rd)
    assert str(a["plugin_module"]) == "test.matcher" and str(a["registry_module"]) == "test.matcher"
    assert str(b["plugin_module"]) == "roreg_b200.test.matcher" and str(b["registry_module"]) == "roreg_b200.test.matcher"
    pairs = [k[6:] for k in a.files if k.startswith("match_")]
    assert len(pairs) == 3
    for p in pairs:                                                # files of the deterministic stages: identical
        for key in ("match", "scores", "dr"):
            assert a[f"{key}_{p}"].dtype == b[f"{key}_{p}"].dtype and np.array_equal(a[f"{key}_{p}"], b[f"{key}_{p}"]), (key, p)
    assert float(a["fmr"]) == float(b["fmr"]) and float(a["ir"]) == float(b["ir"])
    # yohoc: the reference forks one Pool worker per pair, each starting from the parent's RNG state (test/estimator.py:258); the
    # mirror restores that state before every pair.  Pool does not promise one pair per worker, so a pair whose worker had
    # already served another one may draw differently in the reference - then both poses must still register the pair.
    same = 0
    for p in pairs:
        if int(a[f"recall_{p}"]) == int(b[f"recall_{p}"]):
            assert np.abs(a[f"trans_{p}"] - b[f"trans_{p}"]).max() < 1e-9
            same += 1
        else:
            for z in (a, b):
                assert np.abs(z[f"trans_{p}"][:3] - z[f"gt_{p}"]).max() < 2e-2
    assert same >= 1
    assert float(a["rr"]) == float(b["rr"]) == 1.0
    assert abs(float(a["rre"]) - float(b["rre"])) < 1e-3 and abs(float(a["rte"]) - float(b["rte"])) < 1e-3
    assert np.array_equal(a["rng_after"], b["rng_after"])          # the parent's global RNG ends in the same state
    la = bytes(a["pre_log"]).decode().splitlines(); lb = bytes(b["pre_log"]).decode().splitlines()
    assert len(la) == len(lb) == 15 and la[0::5] == lb[0::5] and la[4::5] == lb[4::5]
