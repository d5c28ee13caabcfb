# H2D copy time of one DVS message (3 MB) and one packet (16 MB) from pinned memory placed on each NUMA node
import os, glob, sys, time
import torch
def cpulist(s):
    out = []
    for part in s.strip().split(","):
        if "-" in part:
            a, b = part.split("-"); out += list(range(int(a), int(b) + 1))
        elif part: out.append(int(part))
    return out
p = torch.cuda.get_device_properties(0)
bdf = "%04x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
try:
    print("GPU0", bdf, "numa_node", open(f"/sys/bus/pci/devices/{bdf}/numa_node").read().strip(), "local_cpulist", open(f"/sys/bus/pci/devices/{bdf}/local_cpulist").read().strip())
except Exception as e:
    print("sysfs:", e)
print("affinity now:", len(os.sched_getaffinity(0)), "cpus")
nodes = sorted(glob.glob("/sys/devices/system/node/node[0-9]*"))
all_cpus = os.sched_getaffinity(0)
s = torch.cuda.Stream()
dst = torch.empty(16 << 20, dtype=torch.uint8, device="cuda")
for nd in nodes + [None]:
    if nd is not None:
        cpus = set(cpulist(open(nd + "/cpulist").read())) & all_cpus
        if not cpus: continue
        os.sched_setaffinity(0, cpus)
    else:
        os.sched_setaffinity(0, all_cpus)
    src = torch.empty(16 << 20, dtype=torch.uint8).pin_memory()
    src.numpy()[:] = 1
    for nbytes in (3 << 20, 16 << 20):
        with torch.cuda.stream(s):
            for _ in range(3): dst[:nbytes].copy_(src[:nbytes], non_blocking=True)
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(s)
            for _ in range(20): dst[:nbytes].copy_(src[:nbytes], non_blocking=True)
            e1.record(s)
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / 20 * 1e3
        print(os.path.basename(nd) if nd else "unbound", nbytes >> 20, "MB: %.1f us  %.1f GB/s" % (us, nbytes / us / 1e3))
