"""GPU: end-effector fit check (K9) and projection onto intersection sets (K10) vs the oracle."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def c2sets():
    import torch

    assert torch.cuda.is_available()
    from boundplanner_b200 import geometry as geo, scenes

    boxes, inflate, seeds, ws_min, ws_max = scenes.config_c2(1000, 64)
    sc = geo.Scene(boxes, inflate)
    out = geo.build_sets_point(sc, seeds, ws_min, ws_max, fixed_mid=True, optimize=True)
    Ar, br, mr, _, _ = geo.reduce_ineqs(out.A, out.b, out.m)          # the planner stores reduced sets
    bits, x = geo.pair_feasible(Ar, br, mr, 0.01, want_points=True)
    adj = geo.unpack_adjacency(bits, 64).cpu().numpy()
    pairs = np.argwhere(adj).astype(np.int32)
    return geo, Ar, br, mr, pairs, x.cpu().numpy()


def test_check_fit_matches_oracle(c2sets):
    from oracle.planner_graph import check_intersection

    geo, Ar, br, mr, pairs, x = c2sets
    A, b, m = Ar.cpu().numpy(), br.cpu().numpy(), mr.cpu().numpy()
    l_ee = np.array([0.0, 0.0, 0.05])
    omega_hat = np.array([0.0, 1.0, 0.0])
    for omega_norm in (0.0, 1.2):
        x0 = np.array([x[i, j] for i, j in pairs])
        fits, omega = geo.check_fit(Ar, br, mr, pairs, l_ee, omega_hat, omega_norm, x0=x0)
        fits, omega = fits.cpu().numpy(), omega.cpu().numpy()
        n_fit = 0
        for p, (i, j) in enumerate(pairs):
            a_set = np.vstack((A[i, : m[i]], A[j, : m[j]]))
            b_set = np.concatenate((b[i, : m[i]], b[j, : m[j]]))
            ok, p_inside = check_intersection(a_set, b_set, l_ee, x[i, j], omega_hat, omega_norm)
            assert bool(fits[p]) == ok, f"pair {i},{j}"
            if ok:
                assert abs(omega[p] - p_inside[3]) < 1e-15
                n_fit += 1
        assert 0 < n_fit
    assert len(pairs) > 20


def test_projection_matches_oracle(c2sets):
    from oracle.planner_graph import project_point

    geo, Ar, br, mr, pairs, x = c2sets
    A, b, m = Ar.cpu().numpy(), br.cpu().numpy(), mr.cpu().numpy()
    rng = np.random.default_rng(3)
    xd = rng.uniform([-1, -1, 0], [1, 1, 1.2], (len(pairs), 3))
    xd[0] = x[pairs[0][0], pairs[0][1]]                       # a point already inside: projection = itself
    xp, status = geo.project_points(Ar, br, mr, pairs, xd)
    xp = xp.cpu().numpy()
    assert status.cpu().numpy().tolist() == [0] * len(pairs)
    assert np.abs(xp[0] - xd[0]).max() < 1e-12
    for p, (i, j) in enumerate(pairs[:60]):
        a_set = np.vstack((A[i, : m[i]], A[j, : m[j]]))
        b_set = np.concatenate((b[i, : m[i]], b[j, : m[j]]))
        xo = project_point(a_set, b_set, xd[p])
        assert np.abs(xp[p] - xo).max() < 1e-9, f"pair {i},{j}"
        assert np.max(a_set @ xp[p] - b_set) < 1e-9
