"""CPU, world_size 2 over gloo: the multi-GPU sharding path (all-gather of the
halfspace tensors, balanced row blocks, all-gather of adjacency bit rows) gives
the same graph as a single process.  The pair kernel is injected as a callable;
here it is the reference's linprog call packed into bit words."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

BOX = np.vstack((np.eye(3), -np.eye(3)))


def _sets(n, seed, m_max=16):
    rng = np.random.default_rng(seed)
    A = np.zeros((n, m_max, 3))
    b = np.full((n, m_max), 10.0)
    m = np.zeros(n, np.int32)
    for s in range(n):
        k = rng.integers(3, m_max - 6)
        c = rng.uniform(-0.6, 0.6, 3)
        An = rng.normal(size=(k, 3))
        An /= np.linalg.norm(An, axis=1)[:, None]
        A[s, : 6 + k] = np.vstack((BOX, An))
        b[s, : 6 + k] = np.concatenate((np.array([1, 1, 1.2, 1, 1, 0.0]), An @ c + rng.uniform(0.05, 0.5, k)))
        m[s] = 6 + k
    return A, b, m


def cpu_pair_fn(A, b, m, tol, r0, r1):
    from oracle.set_graph import set_intersection

    A, b, m = A.numpy(), b.numpy(), m.numpy()
    S = A.shape[0]
    words = (S + 31) // 32
    bits = np.zeros((r1 - r0, words), np.uint32)
    for i in range(r0, r1):
        for j in range(i + 1, S):
            if set_intersection([A[i, : m[i]], b[i, : m[i]]], [A[j, : m[j]], b[j, : m[j]]], tol)[2]:
                bits[i - r0, j >> 5] |= np.uint32(1 << (j & 31))
    return torch.from_numpy(bits.view(np.int32).copy())


def _worker(rank, world, port, n_local, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from boundplanner_b200 import distributed as bpd

    A, b, m = _sets(n_local * world, 77)
    lo, hi = bpd.shard_range(n_local * world, rank, world)
    bits, (Ag, bg, mg) = bpd.sharded_adjacency(torch.from_numpy(A[lo:hi]), torch.from_numpy(b[lo:hi]),
                                               torch.from_numpy(m[lo:hi]), cpu_pair_fn, 0.01)
    assert np.array_equal(Ag.numpy(), A) and np.array_equal(bg.numpy(), b) and np.array_equal(mg.numpy(), m)
    np.save(os.path.join(out_dir, f"bits_{rank}.npy"), bits.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_adjacency_world2(tmp_path):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    n_local = 9
    mp.spawn(_worker, args=(2, port, n_local, str(tmp_path)), nprocs=2, join=True)
    A, b, m = _sets(2 * n_local, 77)
    single = cpu_pair_fn(torch.from_numpy(A), torch.from_numpy(b), torch.from_numpy(m), 0.01, 0, 2 * n_local).numpy()
    for r in range(2):
        assert np.array_equal(np.load(tmp_path / f"bits_{r}.npy"), single)
    assert single.any()
