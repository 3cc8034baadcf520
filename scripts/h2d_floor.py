"""The box's host->device floor: every rank copies the C2-sized pinned buffer (5 GB) to its GPU, nothing else, all
ranks at once.  The end-to-end scaling of bench.py is bounded by this number (all GPUs share the host side).
    torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P scripts/h2d_floor.py [pieces]"""
import os
import sys

import torch
import torch.distributed as dist

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
n = 2_499_999_600
pieces = int(sys.argv[1]) if len(sys.argv) > 1 else 17
host = torch.empty(n, dtype=torch.uint16, pin_memory=True)
host.zero_()
devbuf = torch.empty(n, dtype=torch.uint16, device=dev)
cuts = [n * i // pieces for i in range(pieces + 1)]


def copy():
    for a, b in zip(cuts[:-1], cuts[1:]):
        devbuf[a:b].copy_(host[a:b], non_blocking=True)


res = {}
for name, fn in (("one copy", lambda: devbuf.copy_(host, non_blocking=True)), (f"{pieces} pieces", copy)):
    fn(); torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(3):
        fn()
    e1.record(); torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / 3], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    res[name] = float(ms.item())
if rank == 0:
    for name, ms in res.items():
        print(f"h2d floor, {world} GPU(s), {name}: {ms:.1f} ms per 5 GB per GPU (max over ranks) = {2 * n / ms / 1e6:.1f} GB/s per GPU, "
              f"{world * 2 * n / ms / 1e6:.1f} GB/s aggregate = {world * n / ms / 1e3:.0f} Msamples/s ceiling for the end-to-end path")
    try:
        import subprocess
        print(subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True).stdout[-3000:])
        print(subprocess.run(["numactl", "-H"], capture_output=True, text=True).stdout[:1500])
    except Exception as e:
        print("topology unavailable:", e)
if world > 1:
    dist.destroy_process_group()
