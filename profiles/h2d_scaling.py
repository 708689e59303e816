"""Host->device copy bandwidth of the box with 1 / 2 / 4 / 8 ranks copying at once (what bounds bench.py's e2e leg).
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 profiles/h2d_scaling.py
Every rank copies a 1 GiB page-locked buffer to its GPU in a loop for ~1.5 s per phase; phases with 1, 2, 4, 8 ranks
active.  Two allocations are compared: plain cudaHostAlloc, and cudaHostAlloc under set_mempolicy(MPOL_BIND, the GPU's
NUMA node from sysfs) -- when the container's cpuset allows that node."""
import ctypes
import json
import os
import time

import torch
import torch.distributed as dist

rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); lr = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(lr)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
libc = ctypes.CDLL(None, use_errno=True)
SYS_set_mempolicy = 238            # x86_64
MPOL_DEFAULT, MPOL_BIND = 0, 2


def gpu_numa_node(i):
    p = torch.cuda.get_device_properties(i)
    path = "/sys/bus/pci/devices/%04x:%02x:%02x.0/numa_node" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
    try:
        return int(open(path).read().strip())
    except (OSError, ValueError):
        return -1


def set_policy(node):
    if node < 0:
        return libc.syscall(SYS_set_mempolicy, MPOL_DEFAULT, None, 0)
    mask = ctypes.c_ulong(1 << node)
    return libc.syscall(SYS_set_mempolicy, MPOL_BIND, ctypes.byref(mask), 65)


def alloc(nbytes, node):
    rc = set_policy(node) if node >= 0 else 0
    err = ctypes.get_errno() if rc else 0
    t = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    t.fill_(7)                     # touch every page under the policy
    set_policy(-1)
    return t, (rc, err)


def phase(host, dev, active):
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    gbs = 0.0
    if rank < active:
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            t0 = time.perf_counter(); n = 0
            while time.perf_counter() - t0 < 1.5:
                for _ in range(4):
                    dev.copy_(host, non_blocking=True)
                s.synchronize(); n += 4
            gbs = n * host.numel() / (time.perf_counter() - t0) / 1e9
    out = torch.tensor([gbs], dtype=torch.float64, device="cuda")
    allv = [torch.zeros_like(out) for _ in range(world)]
    if world > 1:
        dist.all_gather(allv, out)
    else:
        allv = [out]
    return [float(x.item()) for x in allv]


node = gpu_numa_node(lr)
nodes = [None] * world
if world > 1:
    dist.all_gather_object(nodes, (lr, node, sorted(os.sched_getaffinity(0))[:4], len(os.sched_getaffinity(0))))
else:
    nodes = [(lr, node, sorted(os.sched_getaffinity(0))[:4], len(os.sched_getaffinity(0)))]
N = 1 << 30
dev = torch.empty(N, dtype=torch.uint8, device="cuda")
res = {"gpu_numa_nodes": nodes, "online_nodes": open("/sys/devices/system/node/online").read().strip() if os.path.exists("/sys/devices/system/node/online") else None}
for label, nd in (("default", -1), ("bound_to_gpu_node", node)):
    host, rc = alloc(N, nd)
    res[label] = {"set_mempolicy": rc}
    for active in [a for a in (1, 2, 4, 8) if a <= world]:
        v = phase(host, dev, active)
        res[label]["%d_ranks" % active] = {"per_rank_gbs": [round(x, 1) for x in v[:active]], "total_gbs": round(sum(v), 1)}
    del host
if rank == 0:
    print(json.dumps(res))
if world > 1:
    dist.destroy_process_group()
