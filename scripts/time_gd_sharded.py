"""Element-sharded PINN-GD on the C5 lattice under torchrun: ms per iteration on the peer-memory transport and on NCCL.
torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/time_gd_sharded.py [--nx 578] [--iters 100]"""
import argparse, json, os, sys
from pathlib import Path
import torch, torch.distributed as dist
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from pinn_fem_b200 import sharding as S
from bench_gd import gd_element_sharded_iterations_per_second
from pinn_fem_b200.meshes import lattice_truss

ap = argparse.ArgumentParser()
ap.add_argument("--nx", type=int, default=578)
ap.add_argument("--iters", type=int, default=100)
ap.add_argument("--transports", default="peer,nccl", help="comma list; a ':graph' suffix replays the iteration as a CUDA graph")
a = ap.parse_args()
rank, local, world = S.env_rank_world()
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
nodes, el, fixed = lattice_truss(a.nx)
for item in a.transports.split(","):
    tr, _, mode = item.partition(":")
    os.environ["PF_COMM_TRANSPORT"] = tr
    os.environ["PF_GD_GRAPH"] = "1" if mode == "graph" else "0"
    res = gd_element_sharded_iterations_per_second(dev, nodes, el, fixed, world, iters=a.iters)
    if rank == 0:
        print(json.dumps({"requested": item, "transport": res["transport"], "world": world,
                          "ms_per_iteration": res["ms_per_iteration"],
                          "loop_only_ms_per_iteration": res["loop_only_ms_per_iteration"]}), flush=True)
dist.barrier()
dist.destroy_process_group()
