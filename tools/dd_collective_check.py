"""torchrun --nproc-per-node N tools/dd_collective_check.py [n]: pstDomainDecomp across N GPUs with the trial answers through
the host (torch.distributed all-gathers, 3-4 per trial) and with ONE gg_orb_bisect_all per level (in-stream NCCL all-gather
per trial): same domains, decomposition time of both."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

from gasoline_b200 import domain, ics

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
p = ics.plummer(n, seed=12345)
w = np.random.default_rng(3).uniform(0.5, 8.0, p.n)
out = {}
for weights, tag in ((None, "counts"), (w, "weights")):
    res = {}
    for coll in (False, True, True):
        tm = {}
        dist.barrier()
        idx = domain.device_orb_share(p, rank, world, local, "cuda", weights=weights, timing=tm, collective=coll)
        res[coll] = (idx, tm)
    same = bool(np.array_equal(res[False][0], res[True][0]))
    flag = torch.tensor([1 if same else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    out[tag] = dict(same_domains=bool(flag.item()), n_rank0=int(len(res[True][0])), trials=res[True][1]["trials"],
                    host_ms=res[False][1]["decomp_ms"], collective_ms=res[True][1]["decomp_ms"])
if rank == 0:
    print(json.dumps(dict(what="pstDomainDecomp across GPUs", n=n, world=world, **out)))
dist.barrier()
dist.destroy_process_group()
