"""On an N-GPU box: host->device bandwidth of all GPUs copying at once from pinned memory, with and without binding every
process to the CPUs of its GPU's NUMA node before the pinned buffer is allocated (first touch decides where the pages live)."""
import os, sys, time, subprocess
import multiprocessing as mp


def gpu_numa_cpus(idx):
    import torch
    p = torch.cuda.get_device_properties(idx)
    bdf = "%04x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
    try:
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % bdf).read())
    except Exception:
        return bdf, -1, None
    if node < 0:
        return bdf, node, None
    cpus = set()
    for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
        a, _, b = part.partition("-")
        cpus.update(range(int(a), int(b or a) + 1))
    return bdf, node, cpus


def worker(idx, bind, q, go):
    import torch
    torch.cuda.set_device(idx)
    bdf, node, cpus = gpu_numa_cpus(idx)
    allowed = os.sched_getaffinity(0)
    if bind and cpus and (cpus & allowed):
        os.sched_setaffinity(0, cpus & allowed)
    n = 1 << 29
    h = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    h.fill_(1)
    d = torch.empty(n, dtype=torch.uint8, device="cuda")
    d.copy_(h, non_blocking=True); torch.cuda.synchronize()
    q.put(("ready", idx))
    go.wait()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    reps = 6
    for _ in range(reps):
        d.copy_(h, non_blocking=True)
    e1.record(); torch.cuda.synchronize()
    q.put(("done", idx, reps * n / (e0.elapsed_time(e1) * 1e-3) / 1e9, bdf, node, len(cpus & allowed) if cpus else 0))


if __name__ == "__main__":
    import torch
    n = torch.cuda.device_count()
    print(subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True).stdout[:3000])
    print(subprocess.run("lscpu | grep -i 'numa\\|socket\\|model name\\|^CPU(s)'", shell=True, capture_output=True, text=True).stdout)
    print("affinity of this process:", len(os.sched_getaffinity(0)), "cpus")
    ctx = mp.get_context("spawn")
    for bind in (0, 1):
        q, go = ctx.Queue(), ctx.Event()
        ps = [ctx.Process(target=worker, args=(i, bind, q, go)) for i in range(n)]
        [p.start() for p in ps]
        for _ in range(n): q.get()
        go.set()
        res = sorted(q.get() for _ in range(n))
        [p.join() for p in ps]
        print("bind=%d:" % bind, " ".join("gpu%d %.1f GB/s (numa %d, %d cpus)" % (r[1], r[2], r[4], r[5]) for r in res), "| total %.1f GB/s" % sum(r[2] for r in res), flush=True)
