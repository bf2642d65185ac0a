"""Known-answer identities of the icosahedral group tables."""
import numpy as np

INV = [0,4,3,2,1,5,16,25,20,10,9,15,29,24,14,11,6,17,26,21,8,19,28,23,13,7,18,27,22,12,35,31,46,55,40,30,45,
       59,44,39,34,49,58,43,38,36,32,47,56,41,50,51,52,53,54,33,48,57,42,37]


def test_perm_is_group_table(tables):
    P = tables.perm
    assert (P[:, 0] == np.arange(60)).all() and (P[0] == np.arange(60)).all()
    for a in range(60):
        assert sorted(P[a]) == list(range(60)) and sorted(P[:, a]) == list(range(60))
    a, b, c = np.meshgrid(np.arange(60), np.arange(60), np.arange(60), indexing="ij")
    assert (P[P[a, b], c] == P[a, P[b, c]]).all()            # associativity


def test_perm_matches_rotations(tables):
    R, P = tables.rot, tables.perm
    assert np.abs(R[0] - np.eye(3)).max() < 1e-12
    for a in range(60):
        for b in range(60):
            assert np.abs(R[b] @ R[a] - R[P[a, b]]).max() < 1e-3


def test_nei_and_inverse(tables):
    P, N = tables.perm, tables.nei
    assert list(N[0]) == [0, 1, 4, 7, 8, 11, 12, 15, 19, 20, 21, 25, 29]
    assert (N == P[:, N[0]]).all() and (N[:, 0] == np.arange(60)).all()
    assert list(tables.inv) == INV
