"""Pack the icosahedral group tables (data inputs, not code) into roreg_b200/data/group/icosa60.npz.

Run in the build container only (reads /root/reference/utils/group_related/*.npy, the same
files the reference loads via cfg.SO3_related_files, e.g. test/estimator.py:78).  The SHA-256
prefixes are the ones SURVEY.md section 8(c) records; the algebraic identities are re-checked in
tests/test_group_tables.py on every run.
"""
import hashlib
import os
import numpy as np

SRC = "/root/reference/utils/group_related"
EXPECT = {"60_60.npy": "adc09e66277819dd", "Nei_Index_in_SO3_ordered_13.npy": "a0960735360807ea",
          "Rotation.npy": "21b781b2aab5869f"}
out = os.path.join(os.path.dirname(__file__), "..", "..", "roreg_b200", "data", "group", "icosa60.npz")
arrs = {}
for fn, pre in EXPECT.items():
    raw = open(f"{SRC}/{fn}", "rb").read()
    assert hashlib.sha256(raw).hexdigest().startswith(pre), fn
    arrs[fn] = np.load(f"{SRC}/{fn}")
perm = arrs["60_60.npy"]
nei = arrs["Nei_Index_in_SO3_ordered_13.npy"]
assert np.array_equal(perm, perm.astype(np.int32)) and np.array_equal(nei, nei.astype(np.int32))
np.savez_compressed(out, perm=perm.astype(np.int32), nei=nei.astype(np.int32), rot=arrs["Rotation.npy"])
print("wrote", os.path.abspath(out), os.path.getsize(out), "bytes")
