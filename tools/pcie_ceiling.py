#!/usr/bin/env python
"""Host-link ceiling of the end-to-end step at N GPUs of one box: bare pinned cudaMemcpyAsync of the e2e
step's buffers (H2D 2 x [B, L] float32, D2H 3 x [B, 15, 80, 20] + [B, L] float32) -- H2D alone, D2H alone, both directions at
once -- on all ranks simultaneously, one process per GPU bound to the GPU's NUMA node exactly like bench.py.

    python tools/pcie_ceiling.py                                  # 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/pcie_ceiling.py

Prints one JSON line per run; bench.py reports the same measurement live as e2e.ceiling / e2e.frac_of_ceiling."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    numa = bench.bind_to_gpu_numa_node(local_rank)
    policy = os.environ.get("AVSE_HOST_MEM", "local")           # "interleave": pinned buffers spread over all NUMA nodes
    nodes = bench.set_host_memory_policy(policy)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    B, L = 1000, 48000
    h_in = [torch.empty((B, L), dtype=torch.float32, pin_memory=True) for _ in range(2)]
    h_out = [torch.empty((B, 15, 80, 20), dtype=torch.float32, pin_memory=True) for _ in range(3)] + [torch.empty((B, L), dtype=torch.float32, pin_memory=True)]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    c = bench.copy_ceiling(torch, dist, world, device, h_in, h_out, 8, 10, barrier)
    c["host_cpus_bound_to_gpu_numa_node"] = numa
    c["host_memory_policy"] = policy
    c["numa_nodes"] = nodes
    c["aggregate_bidirectional_gbs"] = c["bidirectional_gbs"] * world
    c["e2e_ceiling_audio_s_per_s"] = world * B * 3.0 / (c["ms_per_step"] * 1e-3)
    if rank == 0:
        print(json.dumps(c), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
