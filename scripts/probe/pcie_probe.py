"""Raw pinned host<->device copy bandwidth at N ranks (one per GPU): the ceiling under HostSession's e2e.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/probe/pcie_probe.py

Per rank: a 648 MB pinned upload buffer and a 599 MB pinned download buffer (one frame's traffic of
bench.py's e2e), copied H2D only, D2H only and both directions at once on two streams; CUDA events,
max over ranks; rank 0 prints one JSON line with per-rank and aggregate GB/s (NUMA-bound and not)."""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from deepsvc_b200 import shard  # noqa: E402


def main():
    rank, local_rank, world = shard.init_distributed()
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    bound = shard.bind_to_gpu_numa_node(local_rank) if os.environ.get("DSVC_PROBE_NUMA") == "1" else False
    up_b, down_b = 648_034_560, 598_682_896
    h_up = torch.empty(up_b, dtype=torch.uint8, pin_memory=True)
    h_dn = torch.empty(down_b, dtype=torch.uint8, pin_memory=True)
    h_up.fill_(1)
    d_up = torch.empty(up_b, dtype=torch.uint8, device=dev)
    d_dn = torch.ones(down_b, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)

    def barrier():
        if world > 1:
            dist.barrier(device_ids=[local_rank])
        torch.cuda.synchronize(dev)

    def run(up, down, n=10):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        s1.wait_event(e0)
        s2.wait_event(e0)
        for _ in range(n):
            if up:
                with torch.cuda.stream(s1):
                    d_up.copy_(h_up, non_blocking=True)
            if down:
                with torch.cuda.stream(s2):
                    h_dn.copy_(d_dn, non_blocking=True)
        torch.cuda.current_stream(dev).wait_stream(s1)
        torch.cuda.current_stream(dev).wait_stream(s2)
        e1.record()
        barrier()
        ms = shard.max_over_ranks(e0.elapsed_time(e1), dev)
        return ((up_b if up else 0) + (down_b if down else 0)) * n / (ms * 1e-3) / 1e9

    run(True, True, 2)
    res = {"h2d_only": run(True, False), "d2h_only": run(False, True), "both": run(True, True)}
    if rank == 0:
        print(json.dumps({"ranks": world, "numa_bound": bool(bound), "cpu_count": os.cpu_count(),
                          "per_rank_gbs": res, "aggregate_gbs": {k: v * world for k, v in res.items()},
                          "frames_per_s_ceiling_full_copy": res["both"] * 1e9 / (up_b + down_b) * world}), flush=True)
    if world > 1:
        dist.barrier(device_ids=[local_rank])
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
