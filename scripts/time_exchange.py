"""torchrun xN: times the slab halo exchange alone (pack_many, grouped NCCL send/recv ring, unpack_many) and the 7-double
all-reduce, at the bench's slab shape."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ocb_loader

ocb = ocb_loader.load()
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
grid = ocb.RectilinearGrid((n, n, n), x=(0, 1), y=(rank / world, (rank + 1) / world), z=(-1, 0), dtype=torch.float64, device=dev)
m = ocb.NonhydrostaticModel(grid, tracers=("b",), closure=ocb.Closure(nu=1e-6, kappa=1e-7))
fields = [m.velocities.u, m.velocities.v, m.velocities.w, m.tracers.b, m.pressures.pNHS]
from oceanostics_b200.distributed import SlabHaloExchanger
ex = SlabHaloExchanger(grid, fields, depth=2, rank=rank, world=world)
t = torch.zeros(7, dtype=torch.float64, device=dev)


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


e, r = timed(ex.exchange), timed(lambda: dist.all_reduce(t))
if rank == 0:
    print(f"exchange {e:.4f} ms ({ex.bytes_per_exchange / 1e6:.1f} MB per rank), all_reduce(7 doubles) {r:.4f} ms; env "
          + " ".join(f"{k}={v}" for k, v in os.environ.items() if k.startswith("NCCL_")), flush=True)
dist.destroy_process_group()
