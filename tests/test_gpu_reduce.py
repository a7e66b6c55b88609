"""GPU: redundancy removal (K8) vs the oracle's restatement of cddlib's algorithm."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_reduce_ineqs_matches_oracle_on_c2_sets():
    import torch

    assert torch.cuda.is_available()
    from boundplanner_b200 import geometry as geo, scenes
    from oracle.reduce_ineqs import redundant_row_mask

    boxes, inflate, seeds, ws_min, ws_max = scenes.config_c2(1000, 96)
    sc = geo.Scene(boxes, inflate)
    out = geo.build_sets_point(sc, seeds, ws_min, ws_max, fixed_mid=True, optimize=True)
    Ar, br, mr, keep, status = geo.reduce_ineqs(out.A, out.b, out.m)
    assert status.cpu().numpy().tolist() == [0] * 96
    A, b, m = out.A.cpu().numpy(), out.b.cpu().numpy(), out.m.cpu().numpy()
    Ar, br, mr, keep = Ar.cpu().numpy(), br.cpu().numpy(), mr.cpu().numpy(), keep.cpu().numpy()
    removed = 0
    for s in range(96):
        red = redundant_row_mask(A[s, : m[s]], b[s, : m[s]])
        assert np.array_equal(~red, keep[s, : m[s]]), f"set {s}"           # index work: exact
        assert mr[s] == (~red).sum()
        assert np.array_equal(Ar[s, : mr[s]], A[s, : m[s]][~red]) and np.array_equal(br[s, : mr[s]], b[s, : m[s]][~red])
        assert np.all(Ar[s, mr[s]:] == 0) and np.all(br[s, mr[s]:] == 10.0)
        removed += int(red.sum())
    assert removed > 96          # the workspace rows are mostly redundant
    # the reduced description is the same polytope: same intersection graph
    bits_full = geo.pair_feasible(out.A, out.b, out.m, 0.01).cpu().numpy()
    bits_red = geo.pair_feasible(torch.as_tensor(Ar).cuda(), torch.as_tensor(br).cuda(), torch.as_tensor(mr).cuda(),
                                 0.01).cpu().numpy()
    assert np.array_equal(bits_full, bits_red)


def test_reduce_ineqs_dropin_and_degenerate_cases():
    import boundplanner_b200 as bp
    from oracle.reduce_ineqs import reduce_ineqs as ref_reduce

    box = np.vstack((np.eye(3), -np.eye(3)))
    # unit cube + a cutting plane + a plane touching in a vertex + a far plane + a duplicated face
    A = np.vstack((box, [[1, 1, 1], [1, 1, 1], [1, 0, 0], [1, 0, 0]]))
    b = np.concatenate((np.ones(6), [2.5, 3.0, 5.0, 1.0]))
    got = bp.reduce_ineqs(A, b)
    want = ref_reduce(A, b)
    assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1])
    assert got[0].shape[0] == 7          # six faces + the cut; touching, far and duplicate rows go
    sets = bp.normalize_set_size([[got[0], got[1]]], 15)
    assert sets[0][0].shape == (15, 3) and np.all(sets[0][1][7:] == 10)
