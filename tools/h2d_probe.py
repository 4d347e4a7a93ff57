"""What bounds the host -> device copy of one batch on this box: PCIe / NUMA probe (diagnostic, 1 GPU).
Prints the box's topology as the container sees it, then the time of back-to-back cudaMemcpyAsync H2D copies
(CUDA events) for several sizes and pinned-allocation flavours (default, write-combined), from every NUMA node's CPUs
when more than one node is visible, and split over two streams."""
import ctypes
import glob
import os
import subprocess
import sys

import torch

rt = ctypes.CDLL([p for p in glob.glob(os.path.join(os.path.dirname(torch.__file__), "..", "nvidia", "cuda_runtime", "lib",
                                                    "libcudart.so*"))][0])
rt.cudaHostAlloc.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_size_t, ctypes.c_uint]
rt.cudaFreeHost.argtypes = [ctypes.c_void_p]
rt.cudaMemcpyAsync.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p]


def sh(cmd):
    try:
        return subprocess.run(cmd, shell=True, capture_output=True, text=True, timeout=20).stdout.strip()
    except Exception as e:  # noqa: BLE001
        return f"<{e}>"


def host_alloc(nbytes, flags):
    p = ctypes.c_void_p()
    rc = rt.cudaHostAlloc(ctypes.byref(p), nbytes, flags)
    assert rc == 0, rc
    ctypes.memset(p, 1, nbytes)            # first touch
    return p


def time_copies(dst, src, nbytes, n=40, streams=1):
    ss = [torch.cuda.Stream() for _ in range(streams)]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    part = nbytes // streams
    torch.cuda.synchronize()
    main = torch.cuda.current_stream()
    for it in range(n + 5):
        if it == 5:
            e0.record(main)
        for k, s in enumerate(ss):
            s.wait_stream(main)
            rt.cudaMemcpyAsync(ctypes.c_void_p(dst + k * part), ctypes.c_void_p(src.value + k * part), part, 1,
                               ctypes.c_void_p(s.cuda_stream))
            main.wait_stream(s)
    e1.record(main)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


def main():
    torch.cuda.init()
    print("== topology");
    print(sh("nvidia-smi topo -m | head -20"))
    print(sh("nvidia-smi --query-gpu=name,pci.bus_id,pcie.link.gen.current,pcie.link.gen.max,pcie.link.width.current,pcie.link.width.max --format=csv"))
    print("numa nodes:", sh("ls -d /sys/devices/system/node/node* | tr '\\n' ' '"), "| cpus:", sh("nproc"), "| affinity:", sorted(os.sched_getaffinity(0)))
    for n in glob.glob("/sys/bus/pci/devices/*/numa_node"):
        v = open(n).read().strip()
        cls = open(os.path.join(os.path.dirname(n), "class")).read().strip()
        if cls.startswith("0x0302") or cls.startswith("0x0300"):
            print("gpu", n.split("/")[-2], "numa_node", v)
    print(sh("lscpu | grep -iE 'model name|socket|numa|hypervisor|^cpu\\(s\\)'"))
    dev = torch.empty(64 << 20, dtype=torch.uint8, device="cuda")
    nodes = sorted(glob.glob("/sys/devices/system/node/node[0-9]*"))
    aff0 = os.sched_getaffinity(0)
    cpusets = [("current affinity", aff0)]
    if len(nodes) > 1:
        for nd in nodes:
            cpus = set()
            for part in open(os.path.join(nd, "cpulist")).read().strip().split(","):
                if part:
                    a, _, b = part.partition("-")
                    cpus |= set(range(int(a), int(b or a) + 1))
            cpus &= aff0
            if cpus:
                cpusets.append((os.path.basename(nd), cpus))
    for name, cpus in cpusets:
        os.sched_setaffinity(0, cpus)
        for flag_name, flags in (("default", 0), ("write-combined", 4)):
            buf = host_alloc(64 << 20, flags)
            row = []
            for size in (64 << 10, 256 << 10, 1 << 20, 4 << 20, 16 << 20, 64 << 20):
                us = time_copies(dev.data_ptr(), buf, size, n=40 if size <= (4 << 20) else 10)
                row.append(f"{size >> 10} KB: {us:.1f} us ({size / us / 1e3:.1f} GB/s)")
            us2 = time_copies(dev.data_ptr(), buf, 1 << 20, streams=2)
            us4 = time_copies(dev.data_ptr(), buf, 1 << 20, streams=4)
            print(f"[{name}] pinned {flag_name}: " + " | ".join(row) + f" | 1 MB over 2 streams {us2:.1f} us, over 4 {us4:.1f} us", flush=True)
            rt.cudaFreeHost(buf)
    os.sched_setaffinity(0, aff0)
    # torch's own pinned allocator, as bench.py uses it
    h = torch.empty(1 << 20, dtype=torch.uint8).pin_memory()
    d = torch.empty(1 << 20, dtype=torch.uint8, device="cuda")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(5):
        d.copy_(h, non_blocking=True)
    e0.record()
    for _ in range(40):
        d.copy_(h, non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    print(f"torch pin_memory 1 MB copy_: {e0.elapsed_time(e1) / 40 * 1e3:.1f} us")
    # device -> host for comparison
    buf = host_alloc(16 << 20, 0)
    s = torch.cuda.current_stream()
    e0.record()
    for _ in range(20):
        rt.cudaMemcpyAsync(buf, ctypes.c_void_p(dev.data_ptr()), 1 << 20, 2, ctypes.c_void_p(s.cuda_stream))
    e1.record()
    torch.cuda.synchronize()
    print(f"D2H 1 MB: {e0.elapsed_time(e1) / 20 * 1e3:.1f} us")


if __name__ == "__main__":
    main()
