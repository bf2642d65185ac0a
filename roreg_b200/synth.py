"""Synthetic cloud pairs for parity tests and bench.py.

There is no network for the 3DMatch / ETH datasets and the FCGF backbone (MinkowskiEngine)
cannot run here, so every test and the bench use seeded synthetic pairs that have the
shapes, dtypes, value ranges and group structure of the reference's cached features:
  * keypoints  [N,3] float64 in metres                       (dataops/dataset.py:109-116)
  * descriptors [N,32,60] float32, unit L2 norm over the 32 axis per (n,g)
                                                               (network/group_feat.py:42)
Convention (dataops/dataset.py:27-30):  R_gt . pts(id1) + t_gt = pts(id0).
Cloud id0 is cloud id1 rotated by R_gt = R_res . Rgroup[a]; by the equivariance law
F(R_a x)[:,:,g] = F(x)[:,:,P[a][g]] its descriptors are the id1 descriptors with the group
axis permuted by P[a] (plus noise), so Des2R(feats1, feats0) -> a  (test/estimator.py:110).
"""
import numpy as np
from . import group as _group


def _unit(x, axis):
    return x / np.maximum(np.linalg.norm(x, axis=axis, keepdims=True), 1e-12)


def small_rotation(rng, max_deg):
    ax = _unit(rng.standard_normal(3), 0)
    th = np.deg2rad(max_deg) * rng.random()
    K = np.array([[0, -ax[2], ax[1]], [ax[2], 0, -ax[0]], [-ax[1], ax[0], 0]])
    return np.eye(3) + np.sin(th) * K + (1 - np.cos(th)) * (K @ K)


def make_pair(seed, n=5000, overlap=0.6, sigma_desc=0.08, sigma_xyz=0.01, max_res_deg=15.0,
              tables=None, with_fcgf=False):
    """Returns dict: keys0, keys1 [n,3] f64; feats0, feats1 [n,32,60] f32; gt [3,4] f64;
    a (coarse rotation index); corr0 [n] = index in cloud1 of the true partner or -1;
    optionally fcgf0, fcgf1 (ET inputs, an independent draw with the same structure)."""
    tb = tables or _group.load()
    rng = np.random.default_rng(seed)
    a = int(rng.integers(0, 60))
    R = small_rotation(rng, max_res_deg) @ tb.rot[a]
    t = rng.uniform(-1.0, 1.0, 3)
    keys1 = rng.uniform(0.0, 3.0, (n, 3))
    feats1 = _unit(rng.standard_normal((n, 32, 60), dtype=np.float32), 1)
    n_ov = int(round(overlap * n))
    src = rng.permutation(n)[:n_ov]                       # rows of cloud1 that have a partner
    keys0 = np.empty((n, 3)); feats0 = np.empty((n, 32, 60), np.float32)
    keys0[:n_ov] = keys1[src] @ R.T + t + sigma_xyz * rng.standard_normal((n_ov, 3))
    feats0[:n_ov] = _unit(feats1[src][:, :, tb.perm[a]]
                          + sigma_desc * rng.standard_normal((n_ov, 32, 60), dtype=np.float32), 1)
    keys0[n_ov:] = rng.uniform(0.0, 3.0, (n - n_ov, 3)) @ R.T + t
    feats0[n_ov:] = _unit(rng.standard_normal((n - n_ov, 32, 60), dtype=np.float32), 1)
    corr0 = np.full(n, -1, np.int64); corr0[:n_ov] = src
    order = rng.permutation(n)                            # hide the identity ordering
    out = dict(keys0=np.ascontiguousarray(keys0[order]), keys1=keys1,
               feats0=np.ascontiguousarray(feats0[order]), feats1=feats1,
               gt=np.concatenate([R, t[:, None]], 1), a=a, corr0=corr0[order], seed=seed)
    if with_fcgf:
        f1 = _unit(rng.standard_normal((n, 32, 60), dtype=np.float32), 1)
        f0 = np.empty_like(f1)
        f0[:n_ov] = _unit(f1[src][:, :, tb.perm[a]]
                          + sigma_desc * rng.standard_normal((n_ov, 32, 60), dtype=np.float32), 1)
        f0[n_ov:] = _unit(rng.standard_normal((n - n_ov, 32, 60), dtype=np.float32), 1)
        out["fcgf0"] = np.ascontiguousarray(f0[order]); out["fcgf1"] = f1
    return out


class SynthDataset:
    """Duck type of dataops/dataset.py:41-129 (ThrDMatchPartDataset) over synthetic pairs:
    `.name`, `.pc_ids`, `.pair_ids`, `.get_kps(id)`, `.get_transform(id0,id1)`.
    Pair p uses clouds (2p, 2p+1) = (id0, id1)."""

    def __init__(self, seeds, n=256, name="synth/scene", tables=None, with_fcgf=True, **kw):
        self.name = name
        self.pairs = [make_pair(s, n=n, tables=tables, with_fcgf=with_fcgf, **kw) for s in seeds]
        self.pc_ids = [str(i) for i in range(2 * len(seeds))]
        self.pair_ids = [(str(2 * p), str(2 * p + 1)) for p in range(len(seeds))]

    def _cloud(self, cid):
        cid = int(cid)
        return self.pairs[cid // 2], cid % 2

    def get_kps(self, cid):
        pr, s = self._cloud(cid)
        return pr[f"keys{s}"]

    def get_feats(self, cid, kind="feats"):
        pr, s = self._cloud(cid)
        return pr[f"{kind}{s}"]

    def get_transform(self, id0, id1):
        return self.pairs[int(id0) // 2]["gt"].astype(np.float32)

    def write_cache(self, cache_root, backbone="FCGF", yoho=True):
        """Lay the descriptors out as the reference's on-disk contract expects
        (testset.py:180, test/extractor.py:60)."""
        import os
        d_in = f"{cache_root}/{self.name}/{backbone}_Input_Group_feature"
        d_out = f"{cache_root}/{self.name}/YOHO_Output_Group_feature"
        os.makedirs(d_in, exist_ok=True)
        for cid in self.pc_ids:
            if self.pairs[0].get("fcgf0") is not None:
                np.save(f"{d_in}/{cid}.npy", self.get_feats(cid, "fcgf"))
        if yoho:
            os.makedirs(d_out, exist_ok=True)
            for cid in self.pc_ids:
                np.save(f"{d_out}/{cid}.npy", self.get_feats(cid))


class SynthScene:
    """A whole synthetic SCENE with the access pattern of the reference's test sets (3DMatch: 433 clouds / 1623 pairs, a cloud takes
    part in ~7.5 pairs): `n_clouds` views of one world of keypoints, every view = the world rotated by R_c = R_res . Rgroup[a_c]
    (descriptors permuted by P[a_c], the equivariance law of the module docstring), translated, sub-sampled to n keypoints and
    perturbed; pairs = each cloud with its next `span` neighbours.  Same duck type as SynthDataset / dataops/dataset.py:41-129.
    For pair (id0 = c, id1 = d):  R_c R_d^T . pts(d) + (t_c - R_c R_d^T t_d) = pts(c)   (dataops/dataset.py:27-30)."""

    def __init__(self, seed, n_clouds=60, n_pairs=225, n=5000, name="synth/scene", overlap=0.67, sigma_desc=0.08, sigma_xyz=0.01,
                 max_res_deg=7.0, tables=None):
        tb = tables or _group.load()
        rng = np.random.default_rng(seed)
        self.name, self.n = name, n
        n_world = int(round(n / overlap))
        world_xyz = rng.uniform(0.0, 3.0, (n_world, 3))
        world_desc = _unit(rng.standard_normal((n_world, 32, 60), dtype=np.float32), 1)
        self.pc_ids = [str(c) for c in range(n_clouds)]
        self.keys, self.feats, self.pose, self.rows = [], [], [], []
        for c in range(n_clouds):
            a = int(rng.integers(0, 60))
            R = small_rotation(rng, max_res_deg) @ tb.rot[a]
            t = rng.uniform(-1.0, 1.0, 3)
            rows = rng.permutation(n_world)[:n]
            self.keys.append(world_xyz[rows] @ R.T + t + sigma_xyz * rng.standard_normal((n, 3)))
            self.feats.append(_unit(world_desc[rows][:, :, tb.perm[a]] + sigma_desc * rng.standard_normal((n, 32, 60), dtype=np.float32), 1))
            self.pose.append((R, t)); self.rows.append(rows)
        span = -(-n_pairs // n_clouds)
        self.pair_ids = [(str(c), str((c + g) % n_clouds)) for g in range(1, span + 1) for c in range(n_clouds)][:n_pairs]
        self.pair_ids = [(a_, b_) if int(a_) < int(b_) else (b_, a_) for a_, b_ in self.pair_ids]

    def get_kps(self, cid):
        return self.keys[int(cid)]

    def get_feats(self, cid):
        return self.feats[int(cid)]

    def get_transform64(self, id0, id1):
        (Rc, tc), (Rd, td) = self.pose[int(id0)], self.pose[int(id1)]
        R = Rc @ Rd.T
        return np.concatenate([R, (tc - R @ td)[:, None]], 1)

    def get_transform(self, id0, id1):
        return self.get_transform64(id0, id1).astype(np.float32)

    def write_cache(self, cache_root):
        """YOHO_Output_Group_feature/{pc}.npy as test/extractor.py:60 leaves them."""
        import os
        d_out = f"{cache_root}/{self.name}/YOHO_Output_Group_feature"
        os.makedirs(d_out, exist_ok=True)
        for cid in self.pc_ids:
            np.save(f"{d_out}/{cid}.npy", self.feats[int(cid)])


# ------------------------------------------------------------------------------------------------------------------------------
# Random-initialised networks with the parameter names and shapes of the reference's checkpoints (checkpoints/FCGF/{GF,ET,RD,RM}/
# model_best.pth, key `network_state_dict`): the GPU box has neither the checkpoints nor a network, and BASELINE asks for
# "random-init weights of that architecture" in the bench.  Layer tables: network/group_feat.py:7-18, network/ops.py:11-63,
# network/eqv_trans.py:78-101, network/rot_detect.py:35-42, network/rot_coh_match.py:14-32,95-104,123-131,176-186,214-274,323-337.
def _gconv_block(prefix, cin, mid, cout):
    """Residual_Comb_Conv: (BN, conv 1x13) x 2 and a (BN, conv) shortcut when the widths differ."""
    spec = [("bn", f"{prefix}.comb_layer_in.0", cin), ("conv", f"{prefix}.comb_layer_in.2", mid, cin, 13),
            ("bn", f"{prefix}.comb_layer_out.0", mid), ("conv", f"{prefix}.comb_layer_out.2", cout, mid, 13)]
    if cin != cout:
        spec += [("bn", f"{prefix}.short_cut_layer.0", cin), ("conv", f"{prefix}.short_cut_layer.2", cout, cin, 13)]
    return spec


def _mlp2(prefix, cin, mid, cout):
    spec = [("conv", f"{prefix}.net.0", mid, cin, 1), ("conv", f"{prefix}.net.3", cout, mid, 1)]
    return spec + ([("conv", f"{prefix}.res", cout, cin, 1)] if cin != cout else [])


def _attention(prefix):
    return [("conv", f"{prefix}.merge", 32, 32, 1)] + [("conv", f"{prefix}.proj.{i}", 32, 32, 1) for i in range(3)]


def _network_spec(kind):
    if kind == "GF":
        return ([("conv", "PartI_net.Conv_in.0", 256, 32, 13)] + _gconv_block("PartI_net.SO3_Conv_layers.0", 256, 512, 256)
                + [("bn", "PartI_net.Conv_out.comb_layer.0", 256), ("conv", "PartI_net.Conv_out.comb_layer.2", 32, 256, 13)])
    if kind == "ET":
        return ([("bn", "Conv_init.comb_layer.0", 128), ("conv", "Conv_init.comb_layer.2", 256, 128, 13)]
                + _gconv_block("PartII_SO3_Conv_layers.0", 256, 512, 256)
                + [("conv", "PartII_To_R_FC.0", 512, 256, 1), ("bn", "PartII_To_R_FC.1", 512), ("conv", "PartII_To_R_FC.3", 128, 512, 1),
                   ("bn", "PartII_To_R_FC.4", 128), ("conv", "PartII_To_R_FC.6", 4, 128, 1)])
    if kind == "RD":
        return _gconv_block("eqv_encoder.0", 32, 64, 16)
    if kind == "RM":
        spec = []
        for b in range(2):
            for side in ("s2t", "t2s"):
                q = f"Graph.merge_blocks.{b}.cross_graph_{side}"
                spec += _attention(q + ".cross_attn") + _mlp2(q + ".merge", 96, 64, 32)
            for side in ("s", "t"):
                q = f"Graph.merge_blocks.{b}.self_graph_{side}"
                spec += (_attention(q + ".self_attn") + _mlp2(q + ".pos_en", 3, 64, 32) + _mlp2(q + ".ambiguity", 120, 128, 32)
                         + _mlp2(q + ".val_en", 96, 64, 32) + _mlp2(q + ".merge", 96, 64, 32))
        return spec + _mlp2("final_mlp", 64, 64, 32)
    raise KeyError(kind)


def random_weights(kind, seed):
    """state-dict-like {name: float32 array} for kind in GF / ET / RD / RM: fan-in-scaled normal weights, small biases, eval-mode
    BatchNorm statistics near (0, 1).  ET's last layer is biased towards the identity quaternion so that its hypotheses
    quat2mat(q) @ Rgroup[coarse] land near a synthetic pair's planted pose (a random head never registers anything)."""
    rng = np.random.default_rng(seed)
    sd = {}
    for entry in _network_spec(kind):
        if entry[0] == "conv":
            _, name, cout, cin, taps = entry
            sd[name + ".weight"] = (rng.standard_normal((cout, cin, 1, taps)) / np.sqrt(cin * taps)).astype(np.float32)
            sd[name + ".bias"] = (0.1 * rng.standard_normal(cout)).astype(np.float32)
        else:
            _, name, c = entry
            sd[name + ".weight"] = (1 + 0.1 * rng.standard_normal(c)).astype(np.float32)
            sd[name + ".bias"] = (0.1 * rng.standard_normal(c)).astype(np.float32)
            sd[name + ".running_mean"] = (0.1 * rng.standard_normal(c)).astype(np.float32)
            sd[name + ".running_var"] = (1 + 0.2 * rng.random(c)).astype(np.float32)
    if kind == "ET":
        sd["PartII_To_R_FC.6.weight"] *= np.float32(0.05)
        sd["PartII_To_R_FC.6.bias"] = (np.array([3.0, 0, 0, 0]) + 0.02 * rng.standard_normal(4)).astype(np.float32)
    if kind == "RM":
        sd["ot_layer.bin_score"] = np.array(1.0, np.float32)
    return sd
