"""Host-side placement of a rank next to its GPU, and the raw host-to-device ceiling of the box.

The end-to-end figure of the path is bound by the pinned cudaMemcpyAsync that stages the PointCloud2 batch
(BASELINE.json north_star: "PointCloud2 staged to device through pinned cudaMemcpyAsync").  With one rank per GPU
the copy engines of 8 GPUs pull from host DRAM at the same time, so which socket a rank's pinned buffer lives on
decides whether the copy crosses the socket link.  Containers often hide /sys/bus/pci/.../numa_node, so the
placement is found empirically when the topology is not readable: a 256 MB pinned buffer is allocated (and
therefore first-touched) with the thread bound to each candidate CPU group in turn, one H2D copy of it is timed,
and the rank stays on the best group.  Nothing here touches the data path; it only decides where buffers live.
"""
from __future__ import annotations

import ctypes
import os

_SYS_MOVE_PAGES = 279  # x86_64


def allowed_cpus() -> list[int]:
    return sorted(os.sched_getaffinity(0))


def _parse_cpulist(text: str) -> list[int]:
    cpus: list[int] = []
    for part in text.strip().split(","):
        if not part:
            continue
        a, _, b = part.partition("-")
        cpus.extend(range(int(a), int(b or a) + 1))
    return cpus


def numa_nodes() -> dict[int, list[int]]:
    """NUMA node -> allowed CPUs, from /sys (empty when the container hides it)."""
    base = "/sys/devices/system/node"
    out: dict[int, list[int]] = {}
    try:
        allowed = set(allowed_cpus())
        for name in sorted(os.listdir(base)):
            if name.startswith("node") and name[4:].isdigit():
                cpus = [c for c in _parse_cpulist(open(f"{base}/{name}/cpulist").read()) if c in allowed]
                if cpus:
                    out[int(name[4:])] = cpus
    except OSError:
        pass
    return out


def gpu_numa_node(gpu_index: int) -> int | None:
    """NUMA node of the GPU from NVML bus id + /sys (None when hidden or -1)."""
    try:
        import pynvml as nv
        nv.nvmlInit()
        try:
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = gpu_index
            if vis:
                try:
                    idx = int(vis.split(",")[gpu_index])
                except (ValueError, IndexError):
                    idx = gpu_index
            bus = nv.nvmlDeviceGetPciInfo(nv.nvmlDeviceGetHandleByIndex(idx)).busId
        finally:
            nv.nvmlShutdown()
        bus = bus.decode() if isinstance(bus, bytes) else bus
        node = int(open("/sys/bus/pci/devices/" + bus.lower()[-12:] + "/numa_node").read())
        return node if node >= 0 else None
    except Exception:
        return None


def candidate_groups(max_groups: int = 8) -> list[tuple[str, list[int]]]:
    """CPU groups a rank could live on: the NUMA nodes when visible, else contiguous slices of the allowed CPUs
    (sockets own contiguous CPU id ranges on every x86 server board, hyper-thread siblings in a second range)."""
    nodes = numa_nodes()
    if len(nodes) > 1:
        return [(f"node{n}", c) for n, c in sorted(nodes.items())]
    cpus = allowed_cpus()
    n = len(cpus)
    g = 1
    while g * 2 <= max_groups and n // (g * 2) >= 4:
        g *= 2
    if g <= 1:
        return [("all", cpus)]
    per = n // g
    return [(f"cpus{cpus[i * per]}-{cpus[(i + 1) * per - 1]}", cpus[i * per:(i + 1) * per]) for i in range(g)]


def page_nodes(addr: int, nbytes: int, samples: int = 64) -> dict[int, int]:
    """Which NUMA node the pages of [addr, addr+nbytes) sit on (move_pages query; {} when not permitted)."""
    try:
        libc = ctypes.CDLL(None, use_errno=True)
        page = os.sysconf("SC_PAGE_SIZE")
        n = max(1, min(samples, nbytes // page))
        stride = max(page, (nbytes // n) // page * page)
        pages = (ctypes.c_void_p * n)(*[(addr // page * page) + i * stride for i in range(n)])
        status = (ctypes.c_int * n)()
        rc = libc.syscall(_SYS_MOVE_PAGES, 0, ctypes.c_ulong(n), pages, None, status, 0)
        if rc != 0:
            return {}
        out: dict[int, int] = {}
        for s in status:
            out[int(s)] = out.get(int(s), 0) + 1
        return out
    except Exception:
        return {}


def h2d_gbps(host, dev, reps: int = 3, stream=None) -> float:
    """Best-of-`reps` bandwidth of one pinned cudaMemcpyAsync host->dev of the whole tensor (CUDA events)."""
    import torch
    s = stream or torch.cuda.current_stream()
    best = 0.0
    with torch.cuda.stream(s):
        dev.copy_(host, non_blocking=True)
        s.synchronize()
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(s)
            dev.copy_(host, non_blocking=True)
            e1.record(s)
            e1.synchronize()
            best = max(best, host.numel() * host.element_size() / (e0.elapsed_time(e1) * 1e-3) / 1e9)
    return best


def bind_rank(gpu_index: int, probe_bytes: int = 256 << 20) -> dict:
    """Bind this process to the CPUs closest to `gpu_index` BEFORE it allocates its pinned batch.
    Returns what was done: {"how": "sysfs"|"probe"|"none", "group": name, "cpus": n, "table": {group: GB/s}}."""
    import torch
    before = allowed_cpus()
    node = gpu_numa_node(gpu_index)
    nodes = numa_nodes()
    if node is not None and node in nodes:
        os.sched_setaffinity(0, nodes[node])
        return {"how": "sysfs", "group": f"node{node}", "cpus": len(nodes[node]), "table": {}}
    groups = candidate_groups()
    if len(groups) <= 1:
        return {"how": "none", "group": groups[0][0] if groups else "?", "cpus": len(before), "table": {},
                "why": "one CPU group only"}
    table: dict[str, float] = {}
    dev = torch.empty(probe_bytes, dtype=torch.uint8, device=torch.device("cuda", gpu_index))
    placed: dict[str, dict] = {}
    for name, cpus in groups:
        try:
            os.sched_setaffinity(0, cpus)
            host = torch.empty(probe_bytes, dtype=torch.uint8, pin_memory=True)
            host.fill_(1)
            table[name] = h2d_gbps(host, dev, reps=3)
            placed[name] = page_nodes(host.data_ptr(), probe_bytes)
            del host
        except Exception as e:  # a group the cgroup does not allow
            table[name] = 0.0
            placed[name] = {"error": type(e).__name__}
    del dev
    best = max(table, key=lambda k: table[k])
    spread = max(table.values()) / max(min(v for v in table.values() if v > 0), 1e-9) if any(table.values()) else 1.0
    if spread < 1.03:
        # no measurable difference between the groups: stay unbound (the scheduler may still move us)
        os.sched_setaffinity(0, before)
        return {"how": "probe", "group": "all (groups within 3%)", "cpus": len(before), "table": table,
                "pages": placed}
    os.sched_setaffinity(0, dict(groups)[best])
    return {"how": "probe", "group": best, "cpus": len(dict(groups)[best]), "table": table, "pages": placed}


def h2d_probe(gpu_index: int, nbytes: int = 1 << 30, reps: int = 6, barrier=None) -> dict:
    """The box's raw host->device ceiling for this rank: a bare loop of `reps` pinned cudaMemcpyAsync copies of
    `nbytes`, every rank at the same time when `barrier` (a callable) is given.  Buffers are allocated here, i.e.
    after bind_rank()."""
    import torch
    dev = torch.empty(nbytes, dtype=torch.uint8, device=torch.device("cuda", gpu_index))
    host = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    host.fill_(1)
    s = torch.cuda.Stream(device=torch.device("cuda", gpu_index))
    with torch.cuda.stream(s):
        dev.copy_(host, non_blocking=True)
    s.synchronize()
    if barrier:
        barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(s):
        e0.record(s)
        for _ in range(reps):
            dev.copy_(host, non_blocking=True)
        e1.record(s)
    e1.synchronize()
    gbps = reps * nbytes / (e0.elapsed_time(e1) * 1e-3) / 1e9
    pages = page_nodes(host.data_ptr(), nbytes)
    del host, dev
    return {"GBps": gbps, "bytes": nbytes, "reps": reps, "pages_by_node": pages}
