"""torchrun target (2+ GPUs): rank 0 factorizes, NCCL-broadcasts the factor, every rank solves its own block of
right-hand sides; the gathered solution is checked against Q on every rank."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "gaussianmarkovrandomfields.jl_b200")]
from gmrf_b200 import sharding, spde  # noqa: E402
from gmrf_b200.backend import B200Backend, ordering_permutation  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
cells = int(sys.argv[1]) if len(sys.argv) > 1 else 20
Q = spde.MaternSPDE(*spde.mesh3d(cells), 0).precision(1.0, 0.4)
perm = ordering_permutation(Q, "nd")                       # resolved once, identical on every rank
be = B200Backend(Q, ordering=perm, device=local, factorize=(rank == 0))
sharding.broadcast_factor(be, src=0)
ld = sharding.sharded_map(lambda _: be.compute_logdet(), list(range(world)))
assert np.all(ld == ld[0]), ld
B = np.random.default_rng(1).standard_normal((Q.shape[0], 24))
X = sharding.sharded_columns(be.backend_solve, B)
res = np.linalg.norm(Q @ X - B) / np.linalg.norm(B)
assert res < 1e-10, res
dist.barrier()
if rank == 0:
    print(f"broadcast ok: world={world} n={Q.shape[0]} residual={res:.2e} logdet={ld[0]:.12g}")
dist.destroy_process_group()
