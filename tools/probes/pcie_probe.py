"""Raw host<->device copy rates with every rank copying at once (torchrun): what the end-to-end arm can hope for at N GPUs.
   python -m torch.distributed.run --nproc-per-node N tools/probes/pcie_probe.py"""
import os, time, torch, torch.distributed as dist
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 512 << 20
h = torch.empty(n, dtype=torch.uint8).pin_memory(); h2 = torch.empty(n // 2, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device="cuda"); d2 = torch.empty(n // 2, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(kind, reps=6):
    torch.cuda.synchronize()
    if world > 1: dist.barrier()
    t = time.perf_counter()
    for _ in range(reps):
        if kind in ("d2h", "both"):
            with torch.cuda.stream(s1): h.copy_(d, non_blocking=True)
        if kind in ("h2d", "both"):
            with torch.cuda.stream(s2): d2.copy_(h2, non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t
    if world > 1: dist.barrier()
    return dt / reps
try:
    node = open("/sys/bus/pci/devices/%s/numa_node" % torch.cuda.get_device_properties(local).pci_bus_id.lower()).read().strip()
except Exception as e:
    node = "?"
for kind in ("d2h", "h2d", "both"):
    run(kind, 2)
    dt = run(kind)
    gb = (n if kind != "h2d" else n // 2) / dt / 1e9
    print(f"rank {rank} gpu-node {node} cpus {sorted(os.sched_getaffinity(0))[:4]}.. {kind}: {dt*1e3:.2f} ms  ({gb:.1f} GB/s {'d2h' if kind!='h2d' else 'h2d'})", flush=True)
