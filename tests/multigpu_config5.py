"""BASELINE config 5 at 101^2 x 100 = 1,020,100 latent dofs (half of the ~2 M target; the largest size whose factor + update
pool + selected inverse fit one B200), posterior samples sharded over the GPUs of one node:
   python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 tests/multigpu_config5.py [nt]
Reuses bench.py's config-5 leg (rank 0 factorizes, NCCL broadcast of the factor, 1024 / N draws per rank; and the
redundant-factorization variant)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "gaussianmarkovrandomfields.jl_b200")]
import torch
import torch.distributed as dist
import bench

world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
nt = int(sys.argv[1]) if len(sys.argv) > 1 else 100
out = bench.sampling_config5(world, rank, local, nt=nt)
if rank == 0:
    print(json.dumps({"n_gpus": world, **out}), flush=True)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
