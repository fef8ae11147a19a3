"""A/B: all-reduce of ONE double, the size of the CG dot products, inside a CUDA graph --
(a) the library's peer-memory mailboxes (post kernel + wait kernel per all-reduce, smle_distctl.cuh),
(b) ncclAllReduce through torch.distributed, captured in a torch.cuda.CUDAGraph.
Run under torchrun on N GPUs; rank 0 prints one JSON line.  The row-partitioned CG uses (a) fused
into its kernels (no extra launches); this measures the exchange itself."""
import json
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "sparse-matrix-linear-equations_b200" / "python"))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
import smle_b200 as S  # noqa: E402
from smle_b200 import dist as D  # noqa: E402

world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
torch.cuda.set_device(local)
S.init(local)
st = torch.cuda.Stream()
S.set_stream(st.cuda_stream)


def gather(obj):
    out = [None] * world
    dist.all_gather_object(out, obj)
    return out


A = D.RowPartitionedCsr.grid3d(32, rank, world, gather)
with torch.cuda.stream(st):
    A.allreduce_bench(640)
    torch.cuda.synchronize(); dist.barrier()
    mail_us = A.allreduce_bench(6400)

# NCCL inside a CUDA graph
t = torch.ones(1, dtype=torch.float64, device="cuda")
s2 = torch.cuda.Stream()
with torch.cuda.stream(s2):
    for _ in range(5):
        dist.all_reduce(t)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    t.fill_(1.0)
    with torch.cuda.graph(g, stream=s2):
        for _ in range(64):
            dist.all_reduce(t)
    g.replay()
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(s2)
    for _ in range(100):
        g.replay()
    e1.record(s2)
    torch.cuda.synchronize()
    nccl_us = 1e3 * e0.elapsed_time(e1) / 6400
res = gather((mail_us, nccl_us))
if rank == 0:
    print(json.dumps({"what": "all-reduce of one fp64 inside a CUDA graph, microseconds each, max over ranks", "n_gpus": world,
                      "mailbox_post_plus_wait_kernels_us": max(r[0] for r in res), "nccl_allreduce_us": max(r[1] for r in res),
                      "nccl": ".".join(map(str, torch.cuda.nccl.version()))}), flush=True)
dist.barrier()
A.close()
dist.destroy_process_group()
