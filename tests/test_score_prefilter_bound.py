"""Host-side check of the error bound behind score mode 1 (roreg_set_score_mode, kernels_ransac.cuh ransac_score_pre_kernel).

The kernel decides a point test in float32 only when its float32 squared distance is outside [r2 - m, r2 + m]; this file
replays that arithmetic with NumPy float32 (fma emulated through float64, which differs from a hardware fma by at most one
float32 ulp - far inside the factor 2 the kernel puts on m) on adversarial inputs - points planted within 1e-9 .. 1e-2 of
the inlier sphere, coordinates up to +-50, poses with large translations - and asserts that every DECIDED test agrees with the
float64 evaluation of test/estimator.py:377-382 (overlap_cal), and that the undecided band stays thin.  The GPU side
(mode 1 == mode 0 bit for bit) is tests/test_gpu_parity.py::test_score_mode1_equals_float64_scoring.
"""
import numpy as np

U = 2.0 ** -24


def f32(x):
    return np.asarray(x, np.float64).astype(np.float32)


def fma32(a, b, c):
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(np.float32)


def prefilter(T, a, b, r):
    """T [H,3,4] f64, a/b [K,3] f64 -> (decided [H,K] bool, inlier32 [H,K] bool, band half-width m [H])."""
    r2 = r * r
    Amax = float(np.nextafter(np.float32(np.abs(a).max()), np.float32(np.inf)))
    Bmax = float(np.nextafter(np.float32(np.abs(b).max()), np.float32(np.inf)))
    rrow = np.abs(T[:, :, :3]).sum(2).max(1)
    tmax = np.abs(T[:, :, 3]).max(1)
    E = U * (Amax + 6.0 * (rrow * Bmax + tmax))
    m = 2.0 * (2.0 * np.sqrt(3.0) * E * r + 4.0 * E * E + 6.0 * U * r2)
    lo = np.nextafter(f32(r2 - m), np.float32(-np.inf)); hi = np.nextafter(f32(r2 + m), np.float32(np.inf))
    Tf = f32(T); af = f32(a); bf = f32(b)
    d = []
    for c in range(3):
        x = fma32(Tf[:, c, 0:1], bf[None, :, 0], fma32(Tf[:, c, 1:2], bf[None, :, 1],
                  fma32(Tf[:, c, 2:3], bf[None, :, 2], np.broadcast_to(Tf[:, c, 3:4], (T.shape[0], b.shape[0])))))
        d.append((af[None, :, c] - x).astype(np.float32))
    s2 = fma32(d[2], d[2], fma32(d[1], d[1], (d[0] * d[0]).astype(np.float32)))
    inl = s2 < lo[:, None]
    out = s2 > hi[:, None]
    return inl | out, inl, m


def exact(T, a, b, r):
    x = np.einsum('hcj,kj->hkc', T[:, :, :3], b) + T[:, None, :, 3]
    d = a[None] - x
    return (d * d).sum(2) < r * r, (d * d).sum(2)


def rand_rot(rng, n):
    q = rng.normal(size=(n, 4)); q /= np.linalg.norm(q, axis=1, keepdims=True)
    w, x, y, z = q.T
    return np.stack([1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w),
                     2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w),
                     2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)], 1).reshape(n, 3, 3)


def case(rng, scale, tscale, r, H=48, K=4096):
    R = rand_rot(rng, H); t = rng.uniform(-tscale, tscale, (H, 3))
    T = np.concatenate([R, t[:, :, None]], 2)
    b = rng.uniform(-scale, scale, (K, 3))
    # plant the k0 points on / around the inlier sphere of hypothesis (k mod H): |k0 - (R k1 + t)| = r (1 + eps)
    hsel = np.arange(K) % H
    dirs = rng.normal(size=(K, 3)); dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
    eps = rng.choice([-1, 1], K) * 10.0 ** rng.uniform(-9, -2, K)
    eps[::7] = 0.0
    a = np.einsum('kcj,kj->kc', R[hsel], b) + t[hsel] + dirs * (r * (1 + eps))[:, None]
    return T, a, b


def test_decided_tests_agree_with_float64_and_the_band_is_thin():
    rng = np.random.RandomState(7)
    tot = amb = 0
    for scale, tscale, r in [(3.0, 3.0, 0.1), (3.0, 8.0, 0.1), (50.0, 50.0, 0.5), (1.0, 0.5, 0.02), (10.0, 100.0, 0.1)]:
        T, a, b = case(rng, scale, tscale, r)
        dec, inl, m = prefilter(T, a, b, r)
        ex, d2 = exact(T, a, b, r)
        assert np.array_equal(inl[dec], ex[dec]), (scale, tscale, r)
        # every undecided test really is close to the sphere: the band is a few float32 ulps of the coordinate magnitudes wide
        assert np.all(np.abs(d2[~dec] - r * r) <= 2.1 * m[:, None].repeat(d2.shape[1], 1)[~dec])
        # planted on-sphere points (1/H of all tests) can be undecided; the random rest must essentially never be
        K, H = a.shape[0], T.shape[0]
        planted = (np.arange(K)[None, :] % H) == np.arange(H)[:, None]
        tot += (~planted).sum(); amb += (~dec & ~planted).sum()
    assert amb <= 1e-3 * tot


def test_synthetic_workload_band_fraction():
    """At the bench workload's geometry (clouds in [0,3)^3, |t| < ~5, ird 0.1) the float64 re-check is taken by < 1e-3 of the tests."""
    rng = np.random.RandomState(3)
    H, K, r = 64, 3400, 0.1
    Rg = rand_rot(rng, 1)[0]; tg = rng.uniform(-1, 1, 3)
    b = rng.uniform(0, 3, (K, 3))
    a = b @ Rg.T + tg + rng.normal(0, 0.01, (K, 3))
    a[K // 2:] = rng.uniform(0, 3, (K - K // 2, 3)) @ Rg.T + tg           # half the matches are wrong
    # hypotheses: the true pose perturbed by 0 .. 3 degrees / 0 .. 5 cm (the ones that matter), and random ones
    T = []
    for i in range(H):
        if i % 2:
            Rp = rand_rot(rng, 1)[0]; tp = rng.uniform(-3, 3, 3)
        else:
            w = rng.normal(size=3); w *= np.deg2rad(rng.uniform(0, 3)) / np.linalg.norm(w)
            Kx = np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]])
            Rp = (np.eye(3) + Kx + Kx @ Kx / 2) @ Rg; tp = tg + rng.normal(0, 0.02, 3)
        T.append(np.concatenate([Rp, tp[:, None]], 1))
    T = np.stack(T)
    dec, inl, m = prefilter(T, a, b, r)
    ex, _ = exact(T, a, b, r)
    assert np.array_equal(inl[dec], ex[dec])
    assert (~dec).mean() < 1e-3


def _host_prefilter(T, a, b, r):
    """The kernel's own expressions (roreg_b200/csrc/math3.cuh: prefilter_band, prefilter_dist2, f32_at_or_below / above) compiled
    for the host - real fmaf, real float32 - through libmath3_host.so."""
    import ctypes
    from test_math3_host import _lib, dp
    L = _lib()
    H, K = T.shape[0], a.shape[0]
    T = np.ascontiguousarray(T.reshape(H, 12)); a = np.ascontiguousarray(a); b = np.ascontiguousarray(b)
    dec = np.empty((H, K), np.int8); band = np.empty((H, (K + 255) // 256))
    L.rr_host_prefilter(T.ctypes.data_as(dp), H, a.ctypes.data_as(dp), b.ctypes.data_as(dp), K, ctypes.c_double(r),
                        dec.ctypes.data_as(ctypes.POINTER(ctypes.c_int8)), band.ctypes.data_as(dp))
    return dec, band


def test_kernel_source_on_the_host_never_decides_wrongly():
    """Same adversarial cases through the HOST BUILD of the kernel's arithmetic: every decided test equals the float64 decision,
    the undecided ones are within the band, and the per-tile band is no wider than the whole-cloud band of the NumPy replay."""
    rng = np.random.RandomState(11)
    total = undecided = 0
    for scale, tscale, r in [(3.0, 3.0, 0.1), (3.0, 8.0, 0.1), (50.0, 50.0, 0.5), (1.0, 0.5, 0.02), (10.0, 100.0, 0.1), (3.0, 3.0, 0.1)]:
        T, a, b = case(rng, scale, tscale, r, H=40, K=3000)
        dec, band = _host_prefilter(T, a, b, r)
        ex, d2 = exact(T, a, b, r)
        decided = dec != 2
        assert np.array_equal(dec[decided] == 1, ex[decided]), (scale, tscale, r)
        m_full = np.repeat(band, 256, axis=1)[:, :a.shape[0]]
        assert np.all(np.abs(d2[~decided] - r * r) <= 1.5 * m_full[~decided] + 1e-12)
        _, _, m_numpy = prefilter(T, a, b, r)
        assert np.all(band <= m_numpy[:, None] * (1 + 1e-12))
        planted = (np.arange(a.shape[0])[None, :] % T.shape[0]) == np.arange(T.shape[0])[:, None]
        total += (~planted).sum(); undecided += (~decided & ~planted).sum()
    assert undecided <= 1e-3 * total
    # hypotheses with NaN / inf entries never decide anything (the kernel sends them to float64)
    T, a, b = case(rng, 3.0, 3.0, 0.1, H=4, K=300)
    T[0, 1, 1] = np.nan; T[1, 2, 3] = np.inf; T[2, 0, 0] = -np.inf
    dec, _ = _host_prefilter(T, a, b, 0.1)
    assert np.all(dec[:3] == 2)
