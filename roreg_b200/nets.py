"""Group-convolution networks on the B200 library: GF (network/group_feat.py:7-45), ET head
(network/eqv_trans.py:78-138) and the RD detector (network/rot_detect.py:35-55), inference only.

Host side = weight packing (once per checkpoint) and the layer schedule; every FLOP runs in
libroreg_b200.so (pack / implicit group-convolution tcgen05 GEMM with fused bias+residual+BN+ReLU epilogue / tails).
State-dict key names are the reference checkpoints' (checkpoints/FCGF/{GF,ET,RD}/model_best.pth)."""
import ctypes as C
import numpy as np
import torch

from . import _lib
from .ops import _ptr, _stream


def tf32_split(x):
    """hi = x rounded to 10 mantissa bits (round-to-nearest, ties away - cvt.rna.tf32.f32), lo = x - hi."""
    x = np.ascontiguousarray(x, np.float32)
    b = x.view(np.uint32).astype(np.uint64)
    hi = ((b + 0x1000) & 0xFFFFE000).astype(np.uint32).view(np.float32)
    return hi, (x - hi).astype(np.float32)


def bn_fold(sd, prefix):
    """eval-mode BatchNorm as y = x*scale + shift (float32, eps = 1e-5)."""
    w, b = sd[prefix + ".weight"], sd[prefix + ".bias"]
    m, v = sd[prefix + ".running_mean"], sd[prefix + ".running_var"]
    s = (w / np.sqrt(v + np.float32(1e-5))).astype(np.float32)
    return s, (b - m * s).astype(np.float32)


class Layer:
    """One dense layer's device-resident operands: W_flat[o][(k,c)] hi/lo (rows padded to the N tile), bias."""

    def __init__(self, ctx, W, bias):
        W = np.asarray(W, np.float32)
        O, Cin = W.shape[0], W.shape[1]
        k = W.shape[3] if W.ndim == 4 else 1
        flat = W.reshape(O, Cin, -1).transpose(0, 2, 1).reshape(O, k * Cin)      # [(k, c)] order of the gathered operand rows
        self.O, self.Kdim = O, k * Cin
        self.NT = min(256, -(-O // 16) * 16)
        self.rows = -(-O // self.NT) * self.NT
        pad = np.zeros((self.rows, self.Kdim), np.float32); pad[:O] = flat
        hi, lo = tf32_split(pad)
        self.hi, self.lo = ctx.dev(hi), ctx.dev(lo)
        self.bias = ctx.dev(np.asarray(bias, np.float32))


class GroupNets:
    def __init__(self, ctx, npass=3):
        assert npass in (1, 3)
        self.ctx, self.npass = ctx, npass
        self.lib = ctx.lib

    # ---- thin wrappers over the C ABI ------------------------------------------------------------
    def _buf(self, *shape):
        return torch.empty(shape, dtype=torch.float32, device=self.ctx.device)

    def pack(self, srcs, rows, permute, pre_idx, n_items, bn=None, relu=False):
        n_src = len(srcs)
        hi = self._buf(n_items * 60, n_src * 32); lo = self._buf(n_items * 60, n_src * 32)
        sp = (C.c_void_p * n_src)(*[s.data_ptr() for s in srcs])
        rp = (C.c_void_p * n_src)(*[(r.data_ptr() if r is not None else None) for r in rows])
        pf = (C.c_int32 * n_src)(*[int(p) for p in permute])
        sc = sh = None
        if bn is not None:
            sc, sh = self.ctx.dev(bn[0]), self.ctx.dev(bn[1])
        rc = self.lib.roreg_pack_descriptors(self.ctx.h, n_src, sp, rp, pf, _ptr(pre_idx), n_items, _ptr(sc), _ptr(sh), int(relu),
                                             _ptr(hi), _ptr(lo), _stream())
        _lib.check(self.ctx.h, rc, "roreg_pack_descriptors")
        return hi, lo

    def gconv(self, act, n_items, Cch, L, gset=None, residual=None, res_ld=0, want_raw=False, bn=None, relu=False, want_act=True):
        """Group convolution of the channel-last activation `act` = (hi, lo) [n_items*60][Cch] with layer L as an implicit GEMM
        (roreg_gconv_gemm): rows (item, g) for g in gset (all 60 when None); same epilogue options as gemm()."""
        n_g = 60 if gset is None else int(gset.shape[0])
        R = n_items * n_g
        raw = self._buf(R, L.O) if want_raw else None
        ahi = self._buf(R, L.O) if want_act else None
        alo = self._buf(R, L.O) if (want_act and self.npass == 3) else None
        sc = sh = None
        if bn is not None:
            sc, sh = self.ctx.dev(bn[0]), self.ctx.dev(bn[1])
        assert L.Kdim == 13 * Cch
        rc = self.lib.roreg_gconv_gemm(self.ctx.h, _ptr(act[0]), _ptr(act[1]), n_items, Cch, _ptr(gset), n_g, _ptr(L.hi), _ptr(L.lo), L.rows,
                                       L.O, L.NT, self.npass, _ptr(L.bias), _ptr(residual), res_ld, _ptr(raw), L.O, _ptr(ahi), _ptr(alo), L.O,
                                       _ptr(sc), _ptr(sh), int(relu), _stream())
        _lib.check(self.ctx.h, rc, "roreg_gconv_gemm")
        return raw, (ahi, alo)

    def gemm(self, A, R, L, residual=None, res_ld=0, want_raw=False, bn=None, relu=False, want_act=True):
        raw = self._buf(R, L.O) if want_raw else None
        ahi = self._buf(R, L.O) if want_act else None
        alo = self._buf(R, L.O) if want_act else None
        sc = sh = None
        if bn is not None:
            sc, sh = self.ctx.dev(bn[0]), self.ctx.dev(bn[1])
        rc = self.lib.roreg_gemm(self.ctx.h, _ptr(A[0]), _ptr(A[1]), R, L.Kdim, _ptr(L.hi), _ptr(L.lo), L.rows, L.O, L.NT, self.npass,
                                 _ptr(L.bias), _ptr(residual), res_ld, _ptr(raw), L.O, _ptr(ahi), _ptr(alo), L.O, _ptr(sc), _ptr(sh),
                                 int(relu), _stream())
        _lib.check(self.ctx.h, rc, "roreg_gemm")
        return raw, (ahi, alo)


class GFNet(GroupNets):
    """Group_feat_network.forward (network/group_feat.py:26-45): [n,32,60] -> eqv [n,32,60]."""

    def __init__(self, ctx, sd, prefix="PartI_net.", npass=3, chunk=5000):
        super().__init__(ctx, npass)
        p = prefix
        self.chunk = chunk
        self.L_in = Layer(ctx, sd[p + "Conv_in.0.weight"], sd[p + "Conv_in.0.bias"])
        r = p + "SO3_Conv_layers.0"
        self.bn_a = bn_fold(sd, r + ".comb_layer_in.0"); self.L_a = Layer(ctx, sd[r + ".comb_layer_in.2.weight"], sd[r + ".comb_layer_in.2.bias"])
        self.bn_b = bn_fold(sd, r + ".comb_layer_out.0"); self.L_b = Layer(ctx, sd[r + ".comb_layer_out.2.weight"], sd[r + ".comb_layer_out.2.bias"])
        self.bn_c = bn_fold(sd, p + "Conv_out.comb_layer.0"); self.L_out = Layer(ctx, sd[p + "Conv_out.comb_layer.2.weight"], sd[p + "Conv_out.comb_layer.2.bias"])

    def forward(self, x):
        n = x.shape[0]
        out = torch.empty_like(x)
        for s in range(0, n, self.chunk):
            xs = x[s:s + self.chunk].contiguous(); m = xs.shape[0]; R = m * 60
            a0 = self.pack([xs], [None], [0], None, m)                                  # Conv_in has no BN/ReLU (group_feat.py:16)
            raw0, act1 = self.gconv(a0, m, 32, self.L_in, want_raw=True, bn=self.bn_a, relu=True)
            _, act2 = self.gconv(act1, m, 256, self.L_a, bn=self.bn_b, relu=True)
            _, act3 = self.gconv(act2, m, 512, self.L_b, residual=raw0, res_ld=256, bn=self.bn_c, relu=True)   # identity shortcut (ops.py:60-63)
            raw3, _ = self.gconv(act3, m, 256, self.L_out, want_raw=True, want_act=False)
            rc = self.lib.roreg_gf_finalize(self.ctx.h, _ptr(raw3), _ptr(xs), m, _ptr(out[s:s + m]), _stream())
            _lib.check(self.ctx.h, rc, "roreg_gf_finalize")
        return out


class ETNet(GroupNets):
    """ET_test.forward (network/eqv_trans.py:119-138) -> unit quaternions [K,4].  Only group element 0 of the FC
    head is used (:136), so the last residual conv is evaluated at g = 0 and the middle one at its 13 neighbours."""

    def __init__(self, ctx, sd, npass=3, chunk=5000):
        super().__init__(ctx, npass)
        self.chunk = chunk
        self.bn0 = bn_fold(sd, "Conv_init.comb_layer.0"); self.L0 = Layer(ctx, sd["Conv_init.comb_layer.2.weight"], sd["Conv_init.comb_layer.2.bias"])
        r = "PartII_SO3_Conv_layers.0"
        self.bn_a = bn_fold(sd, r + ".comb_layer_in.0"); self.L_a = Layer(ctx, sd[r + ".comb_layer_in.2.weight"], sd[r + ".comb_layer_in.2.bias"])
        self.bn_b = bn_fold(sd, r + ".comb_layer_out.0"); self.L_b = Layer(ctx, sd[r + ".comb_layer_out.2.weight"], sd[r + ".comb_layer_out.2.bias"])
        self.F0 = Layer(ctx, sd["PartII_To_R_FC.0.weight"], sd["PartII_To_R_FC.0.bias"]); self.bnf1 = bn_fold(sd, "PartII_To_R_FC.1")
        self.F3 = Layer(ctx, sd["PartII_To_R_FC.3.weight"], sd["PartII_To_R_FC.3.bias"]); self.bnf4 = bn_fold(sd, "PartII_To_R_FC.4")
        self.F6 = Layer(ctx, sd["PartII_To_R_FC.6.weight"], sd["PartII_To_R_FC.6.bias"])
        self.g13 = ctx.dev(ctx.tables.nei[0].astype(np.int32))            # N[0,:]: the 13 inputs of output g = 0

    def forward(self, before0, rows_b0, before1, rows_b1, after0, rows_a0, after1, rows_a1, pre_idx):
        """before0/after0 = FCGF-in / GF-out descriptors of cloud id1 (permuted by P[pre_idx]), before1/after1 =
        cloud id0 (test/estimator.py:293-306 swaps the sides); rows_* int32 keypoint indices [K]."""
        K = pre_idx.shape[0]
        quat = torch.empty((K, 4), dtype=torch.float32, device=self.ctx.device)
        for s in range(0, K, self.chunk):
            e = min(K, s + self.chunk); m = e - s
            sl = lambda t: t[s:e].contiguous()
            a0 = self.pack([before0, before1, after0, after1], [sl(rows_b0), sl(rows_b1), sl(rows_a0), sl(rows_a1)], [1, 0, 1, 0],
                           sl(pre_idx), m, bn=self.bn0, relu=True)
            raw0, act1 = self.gconv(a0, m, 128, self.L0, want_raw=True, bn=self.bn_a, relu=True)
            _, act2 = self.gconv(act1, m, 256, self.L_a, gset=self.g13, bn=self.bn_b, relu=True)
            # rows (item, j) of act2 are the 13 inputs of output g = 0 in tap order: [m, 13*512] IS the gathered operand row
            A2 = (act2[0].view(m, 13 * 512), act2[1].view(m, 13 * 512) if act2[1] is not None else None)
            _, f = self.gemm(A2, m, self.L_b, residual=raw0, res_ld=60 * 256)          # + shortcut at g = 0; FC input is the raw sum
            _, f = self.gemm(f, m, self.F0, bn=self.bnf1, relu=True)
            _, f = self.gemm(f, m, self.F3, bn=self.bnf4, relu=True)
            q, _ = self.gemm(f, m, self.F6, want_raw=True, want_act=False)
            rc = self.lib.roreg_quat_normalize(self.ctx.h, _ptr(q), 4, m, _ptr(quat[s:e]), _stream())
            _lib.check(self.ctx.h, rc, "roreg_quat_normalize")
        return quat


class RDNet(GroupNets):
    """detector_eqv_test.forward (network/rot_detect.py:43-55): GF-out descriptors [n,32,60] -> saliency [n]."""

    def __init__(self, ctx, sd, npass=3, chunk=5000):
        super().__init__(ctx, npass)
        self.chunk = chunk
        r = "eqv_encoder.0"
        self.bn_in = bn_fold(sd, r + ".comb_layer_in.0"); self.L_in = Layer(ctx, sd[r + ".comb_layer_in.2.weight"], sd[r + ".comb_layer_in.2.bias"])
        self.bn_out = bn_fold(sd, r + ".comb_layer_out.0"); self.L_out = Layer(ctx, sd[r + ".comb_layer_out.2.weight"], sd[r + ".comb_layer_out.2.bias"])
        self.bn_sc = bn_fold(sd, r + ".short_cut_layer.0"); self.L_sc = Layer(ctx, sd[r + ".short_cut_layer.2.weight"], sd[r + ".short_cut_layer.2.bias"])

    def forward(self, x):
        n = x.shape[0]
        scores = torch.empty(n, dtype=torch.float32, device=self.ctx.device)
        for s in range(0, n, self.chunk):
            xs = x[s:s + self.chunk].contiguous(); m = xs.shape[0]; R = m * 60
            a_sc = self.pack([xs], [None], [0], None, m, bn=self.bn_sc, relu=True)
            raw_sc, _ = self.gconv(a_sc, m, 32, self.L_sc, want_raw=True, want_act=False)
            a_in = self.pack([xs], [None], [0], None, m, bn=self.bn_in, relu=True)
            _, act1 = self.gconv(a_in, m, 32, self.L_in, bn=self.bn_out, relu=True)
            raw, _ = self.gconv(act1, m, 64, self.L_out, residual=raw_sc, res_ld=16, want_raw=True, want_act=False)
            feat = torch.empty((m, 32, 60), dtype=torch.float32, device=self.ctx.device)
            rc = self.lib.roreg_rd_finalize(self.ctx.h, _ptr(raw), m, _ptr(feat), _stream())
            _lib.check(self.ctx.h, rc, "roreg_rd_finalize")
            cor, _ = self.ctx.group_corr(feat, feat, variant=1, want_argmax=False)     # autocorrelation (rot_detect.py:50-51)
            rc = self.lib.roreg_row_std60(self.ctx.h, _ptr(cor), m, _ptr(scores[s:s + m]), _stream())
            _lib.check(self.ctx.h, rc, "roreg_row_std60")
        return scores
