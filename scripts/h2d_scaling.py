"""Concurrent pinned host->device copy bandwidth per GPU at N = 1/2/4/8 ranks on ONE box -- what bounds bench.py's `e2e`
arm once several ranks copy at the same time (e2e scaling 0.98 / 0.51 / 0.40 at N = 2 / 4 / 8).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 \
        scripts/h2d_scaling.py [--mb 1024] [--streams 1|2] [--chunk-mb 0]

Every rank copies `mb` MiB from its own pinned buffer to its GPU `reps` times, all ranks bracketed by barriers; rank 0
prints one JSON line with the per-GPU and aggregate GB/s (max time over ranks).  --streams 2 splits the buffer over two
copy streams, --chunk-mb splits each copy into chunks: the two knobs bench.py could use if the DMA engines, not the host
memory system, were the limit."""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mb", type=int, default=1024)
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--streams", type=int, default=1)
    ap.add_argument("--chunk-mb", type=int, default=0)
    ap.add_argument("--numa", action="store_true", help="bind each rank to the NUMA node of its GPU before allocating")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    bound = None
    if args.numa:
        from comet_pose_estimation_b200.launch import bind_to_gpu_numa_node

        bound = bind_to_gpu_numa_node(local) is not None
    n = args.mb * (1 << 20) // 4
    host = torch.empty(n, dtype=torch.float32).pin_memory()
    host.fill_(1.0)
    dst = torch.empty(n, dtype=torch.float32, device=dev)
    streams = [torch.cuda.Stream(device=dev) for _ in range(args.streams)]
    chunk = (args.chunk_mb * (1 << 20) // 4) or n
    pieces = [(o, min(o + chunk, n)) for o in range(0, n, chunk)]

    def copy_once():
        for i, (a, b) in enumerate(pieces):
            with torch.cuda.stream(streams[i % len(streams)]):
                dst[a:b].copy_(host[a:b], non_blocking=True)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(2):
        copy_once()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    cur = torch.cuda.current_stream(dev)
    e0.record(cur)
    for s in streams:
        s.wait_event(e0)
    for _ in range(args.reps):
        copy_once()
    for s in streams:
        cur.wait_stream(s)
    e1.record(cur)
    barrier()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    gbs = args.reps * n * 4 / (ms * 1e-3) / 1e9
    if rank == 0:
        print(json.dumps({"n_gpus": world, "mb_per_copy": args.mb, "streams": args.streams, "chunk_mb": args.chunk_mb,
                          "numa_bound": bound, "h2d_GBps_per_gpu": gbs, "h2d_GBps_aggregate": gbs * world,
                          "cpus": os.cpu_count(), "numa_nodes": len([d for d in os.listdir("/sys/devices/system/node") if d.startswith("node")])
                          if os.path.isdir("/sys/devices/system/node") else None}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
