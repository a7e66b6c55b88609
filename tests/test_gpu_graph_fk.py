"""GPU parity: set-graph kernel (K6) and FK kernel (K7) through the C ABI vs the oracle."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def geo():
    import torch

    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from boundplanner_b200 import geometry

    return geometry


def _random_sets(rng, n, m_max=24):
    box = np.vstack((np.eye(3), -np.eye(3)))
    A = np.zeros((n, m_max, 3))
    b = np.full((n, m_max), 10.0)
    m = np.zeros(n, np.int32)
    for s in range(n):
        k = rng.integers(3, m_max - 6)
        c = rng.uniform(-0.6, 0.6, 3)
        An = rng.normal(size=(k, 3))
        An /= np.linalg.norm(An, axis=1)[:, None]
        A[s, : 6 + k] = np.vstack((box, An))
        b[s, : 6 + k] = np.concatenate((np.array([1, 1, 1.2, 1, 1, 0.0]), An @ c + rng.uniform(0.05, 0.5, k)))
        m[s] = 6 + k
    return A, b, m


def test_pair_feasible_bit_exact_vs_highs(geo):
    """Adjacency bits == the reference's own linprog/HiGHS answer on every pair
    (BoundPlanner.py:779-784), near-ties (|margin| < 1e-6) excluded and counted."""
    import torch
    from oracle.set_graph import intersection_margin, set_intersection

    rng = np.random.default_rng(21)
    n = 70
    A, b, m = _random_sets(rng, n)
    bits = geo.pair_feasible(torch.as_tensor(A).cuda(), torch.as_tensor(b).cuda(),
                             torch.as_tensor(m).cuda(), tol=0.01)
    adj = geo.unpack_adjacency(bits, n).cpu().numpy()
    sets = [[A[s, : m[s]], b[s, : m[s]]] for s in range(n)]
    near = 0
    for i in range(n):
        for j in range(n):
            if j <= i:
                assert not adj[i, j]
                continue
            ok = bool(set_intersection(sets[i], sets[j], 0.01)[2])
            if adj[i, j] != ok:
                mg = intersection_margin(sets[i], sets[j], 0.01)
                assert abs(mg) < 1e-6, f"pair ({i},{j}): gpu {adj[i, j]} vs HiGHS {ok}, margin {mg}"
                near += 1
    assert near <= 2
    assert 0.05 < adj.sum() / (n * (n - 1) / 2) < 0.95      # the case exercises both answers


def test_pair_feasible_row_blocks_and_padding(geo):
    """Row-block partition (multi-GPU sharding) gives the same bits as one call;
    padded rows (A=0, b=10) are ignored."""
    import torch

    rng = np.random.default_rng(22)
    n = 45
    A, b, m = _random_sets(rng, n)
    At, bt, mt = torch.as_tensor(A).cuda(), torch.as_tensor(b).cuda(), torch.as_tensor(m).cuda()
    full = geo.pair_feasible(At, bt, mt, tol=0.01).cpu().numpy()
    parts = [geo.pair_feasible(At, bt, mt, tol=0.01, row_begin=r0, row_end=r1).cpu().numpy()
             for r0, r1 in ((0, 7), (7, 30), (30, 45))]
    assert np.array_equal(np.vstack(parts), full)
    # use the padded row count instead of m: same answer
    mfull = torch.full_like(mt, A.shape[1])
    padded = geo.pair_feasible(At, bt, mfull, tol=0.01).cpu().numpy()
    assert np.array_equal(padded, full)


def test_pair_feasible_boxes_known_answer(geo):
    """Two boxes intersect with tol iff they overlap by more than 2*tol on every axis."""
    import torch

    box = np.vstack((np.eye(3), -np.eye(3)))
    def mk(lb, ub):
        return box, np.concatenate((ub, -np.asarray(lb)))
    cases = [([0, 0, 0], [1, 1, 1]), ([0.97, 0, 0], [2, 1, 1]), ([0.99, 0, 0], [2, 1, 1]), ([0.5, 0.5, 0.5], [0.6, 0.6, 0.6]),
             ([1.5, 0, 0], [2, 1, 1])]
    n = len(cases)
    A = np.zeros((n, 6, 3)); b = np.zeros((n, 6)); m = np.full(n, 6, np.int32)
    for s, (lb, ub) in enumerate(cases):
        A[s], b[s] = mk(np.array(lb, float), np.array(ub, float))
    bits = geo.pair_feasible(torch.as_tensor(A).cuda(), torch.as_tensor(b).cuda(), torch.as_tensor(m).cuda(), tol=0.01)
    adj = geo.unpack_adjacency(bits, n).cpu().numpy()
    assert adj[0, 1] and not adj[0, 2] and adj[0, 3] and not adj[0, 4]
    assert adj[1, 2] and not adj[1, 3] and adj[1, 4]


def test_fk_matches_oracle_and_anchors(geo):
    from oracle import fk_iiwa14 as ofk

    rng = np.random.default_rng(4)
    B = 1000
    q = rng.uniform(ofk.Q_LOWER, ofk.Q_UPPER, (B, 7))
    q[0] = 0.0
    q[1] = [0, 0, 0, -np.pi / 2, 0, np.pi / 2, 0]
    q[2] = [0.1, -0.2, 0.3, -0.4, 0.5, -0.6, 0.7]
    p_ee, p_col, T, J = geo.fk_iiwa14(q, want_pose=True, want_jacobian=True)
    p_ee, p_col, T, J = p_ee.cpu().numpy(), p_col.cpu().numpy(), T.cpu().numpy(), J.cpu().numpy()
    # anchors derived from iiwa.urdf
    assert np.abs(p_ee[0] - [0, 0, 1.4696]).max() < 1e-12
    assert np.abs(p_col[0][:, 2] - [0.5925, 0.78, 0.9925, 1.18, 1.2596, 1.08, 1.3896]).max() < 1e-12
    assert np.abs(p_ee[1] - [0.4, 0, 0.4904]).max() < 1e-12
    assert np.abs(p_ee[2] - [-0.075099282696, -0.048154013641, 1.429984381972]).max() < 1e-11
    for i in range(0, B, 7):
        assert np.abs(p_ee[i] - ofk.fk_pos(q[i])).max() < 1e-12
        assert np.abs(p_col[i] - ofk.fk_pos_col_all(q[i])).max() < 1e-12
        assert np.abs(T[i] - ofk.hom_transform_endeffector(q[i])).max() < 1e-12
        assert np.abs(J[i] - ofk.jacobian_fk(q[i])).max() < 1e-12
    # ragged tail (B not a multiple of the tile) and the no-pose path
    p2, c2, T2, J2 = geo.fk_iiwa14(q[:130])
    assert T2 is None and J2 is None
    assert np.array_equal(p2.cpu().numpy(), p_ee[:130])
    assert np.array_equal(c2.cpu().numpy(), p_col[:130])
