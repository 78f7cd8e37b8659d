"""World-size-2 `gloo` tests of the multi-GPU sharding logic, runnable without a GPU: the per-rank
evaluator is the CPU oracle standing in for a rank's GPU workspace, so what is tested is the partitioning, the
collectives and the reassembly order -- the part of the N > 1 path that is host logic."""
import os
import socket
import sys

import numpy as np
import pytest
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for _p in (ROOT, os.path.join(ROOT, "gaussianmarkovrandomfields.jl_b200")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

from gmrf_b200 import sharding  # noqa: E402


def test_shard_ranges_cover_and_are_contiguous():
    for n in (0, 1, 2, 7, 256, 1024):
        for world in (1, 2, 3, 8):
            counts = sharding.shard_counts(n, world)
            assert sum(counts) == n and max(counts) - min(counts) <= 1
            edges = [sharding.shard_range(n, world, r) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            for a, b in zip(edges[:-1], edges[1:]):
                assert a[1] == b[0]
    assert sharding.shard_counts(256, 8) == [32] * 8          # BASELINE config 3: 256 theta points on 8 GPUs


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _tridiag(n, tau):
    return (tau * sp.diags([np.full(n - 1, -1.0), np.full(n, 2.01), np.full(n - 1, -1.0)], [-1, 0, 1])).tocsc()


def _worker(rank, world, port, tmp):
    import torch.distributed as dist
    import oracle
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n = 40
        perm = np.arange(n, dtype=np.int64)
        F = oracle.OracleFactor(_tridiag(n, 1.0), perm)       # stands in for this rank's GPU workspace
        z = np.linspace(-1.0, 1.0, n)
        thetas = [0.5 + 0.25 * i for i in range(7)]           # 7 points on 2 ranks: ragged blocks (4 + 3)
        seen = []

        def evaluate(tau):                                    # theta -> refactorize -> (logdet, z'Qz)
            Q = _tridiag(n, tau)
            F.refactorize(Q.data)
            seen.append(tau)
            return np.array([F.logdet(), float(z @ (Q @ z))])

        out = sharding.sharded_map(evaluate, thetas)
        lo, hi = sharding.shard_range(len(thetas), world, rank)
        assert seen == thetas[lo:hi]                          # each rank evaluated only its own contiguous block
        # scalar-valued evaluator -> 1-D result
        out1 = sharding.sharded_map(lambda t: 2.0 * t, thetas)
        # more ranks than items: rank 1 has an empty block
        out_small = sharding.sharded_map(lambda t: np.array([t, -t]), [3.0])
        # right-hand-side blocks: 5 columns on 2 ranks (3 + 2) against ONE factorization
        F.refactorize(_tridiag(n, 1.0).data)
        B = np.random.default_rng(0).standard_normal((n, 5))
        ncols = []

        def solve(Bblk):
            ncols.append(Bblk.shape[1])
            return np.column_stack([F.solve(Bblk[:, j]) for j in range(Bblk.shape[1])])

        X = sharding.sharded_columns(solve, B)
        assert ncols == [sharding.shard_counts(5, world)[rank]]
        np.savez(os.path.join(tmp, f"rank{rank}.npz"), out=out, out1=out1, out_small=out_small, X=X)
    finally:
        dist.destroy_process_group()


def test_sharded_sweep_and_rhs_blocks_world2(tmp_path):
    import torch.multiprocessing as mp
    world, port = 2, _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    r0, r1 = (np.load(tmp_path / f"rank{r}.npz") for r in range(world))
    for k in ("out", "out1", "out_small", "X"):
        assert np.array_equal(r0[k], r1[k])                   # every rank holds the same gathered result
    n = 40
    z = np.linspace(-1.0, 1.0, n)
    thetas = [0.5 + 0.25 * i for i in range(7)]
    want = np.array([[np.linalg.slogdet(_tridiag(n, t).toarray())[1], z @ (_tridiag(n, t) @ z)] for t in thetas])
    assert r0["out"].shape == (7, 2)
    assert np.allclose(r0["out"], want, rtol=1e-12, atol=0)   # item order preserved across the gather
    assert np.array_equal(r0["out1"], 2.0 * np.array(thetas))
    assert np.array_equal(r0["out_small"], np.array([[3.0, -3.0]]))
    B = np.random.default_rng(0).standard_normal((n, 5))
    Xw = np.linalg.solve(_tridiag(n, 1.0).toarray(), B)
    assert np.linalg.norm(r0["X"] - Xw) <= 1e-12 * np.linalg.norm(Xw)


@pytest.mark.gpu
def test_factor_broadcast_two_gpus():
    """Rank 0 factorizes, the factor is broadcast over NCCL, rank 1 solves its block without factorizing.
    Needs two GPUs on the box; the single-GPU round-end run skips it."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import subprocess
    script = os.path.join(ROOT, "tests", "multigpu_factor_broadcast.py")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", str(_free_port()), script], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "broadcast ok" in r.stdout


def _grad_worker(rank, world, port, tmp):
    """A sharded hyperparameter sweep that returns the log-density AND its gradient with respect to the coefficients of the
    value basis: model(ws; theta...) with the values assembled from the resident basis, then the contracted pullbacks."""
    import torch.distributed as dist
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from dense_backend import DenseBackend
    from latent_stand_ins import MaternModel
    from gmrf_b200 import spde
    from gmrf_b200.autodiff import logpdf_basis_gradient
    from gmrf_b200.latent_model_integration import evaluate_with_workspace, make_workspace
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        model = MaternModel(spde.MaternSPDE(*spde.mesh2d(6), 1))
        ws = make_workspace(model, {"backend_type": DenseBackend}, tau=1.0, range_=0.5)     # this rank's replica
        z = 0.1 * np.random.default_rng(0).standard_normal(model.n)
        thetas = [(0.5 + 0.3 * i, 0.4 + 0.05 * i) for i in range(5)]                         # 5 points on 2 ranks (3 + 2)

        def evaluate(th):
            d = evaluate_with_workspace(model, ws, on_device=True, tau=th[0], range_=th[1])
            return np.concatenate([[d.logpdf(z)], logpdf_basis_gradient(d, z, model.basis())])

        out = sharding.sharded_map(evaluate, thetas)
        np.save(os.path.join(tmp, f"grad_rank{rank}.npy"), out)
    finally:
        dist.destroy_process_group()


def test_sharded_gradient_sweep_world2(tmp_path):
    import torch.multiprocessing as mp
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from latent_stand_ins import MaternModel
    from gmrf_b200 import spde
    world, port = 2, _free_port()
    mp.spawn(_grad_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    r0, r1 = (np.load(tmp_path / f"grad_rank{r}.npy") for r in range(world))
    assert np.array_equal(r0, r1)
    model = MaternModel(spde.MaternSPDE(*spde.mesh2d(6), 1))
    n = model.n
    z = 0.1 * np.random.default_rng(0).standard_normal(n)
    basis = model.basis()
    assert r0.shape == (5, 1 + basis.shape[0])
    for i in range(5):
        tau, rho = 0.5 + 0.3 * i, 0.4 + 0.05 * i
        Q = model.precision_matrix(tau, rho)
        Qd = Q.toarray()
        lp = -0.5 * z @ Qd @ z + 0.5 * np.linalg.slogdet(Qd)[1] - 0.5 * n * np.log(2 * np.pi)
        assert abs(r0[i, 0] - lp) <= 1e-10 * abs(lp)
        Sigma = np.linalg.inv(Qd)
        cols = np.repeat(np.arange(n), np.diff(Q.indptr))
        want = 0.5 * (basis @ Sigma[Q.indices, cols] - basis @ (z[Q.indices] * z[cols]))
        assert np.allclose(r0[i, 1:], want, rtol=1e-8, atol=1e-8 * np.max(np.abs(want)))
