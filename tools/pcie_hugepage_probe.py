"""Concurrent D2H/H2D rate of the box with every rank copying at once, for three kinds of page-locked host memory:
cudaHostAlloc (what hm_host_alloc / torch.pin_memory give), cudaHostRegister over transparent huge pages, and over MAP_HUGETLB
pages (if the box has any reserved).  torchrun --nproc-per-node N tools/pcie_hugepage_probe.py"""
import ctypes, mmap, os, time
import torch, torch.distributed as dist

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
SIZE = 1 << 30
dev = torch.empty(SIZE, dtype=torch.uint8, device="cuda")
libc = ctypes.CDLL("libc.so.6", use_errno=True)
libc.mmap.restype = ctypes.c_void_p
libc.mmap.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_long]
libc.madvise.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int]
rt = torch.cuda.cudart()


def buf_hostalloc():
    return torch.empty(SIZE, dtype=torch.uint8).pin_memory()


def buf_registered(hugetlb):
    flags = mmap.MAP_PRIVATE | mmap.MAP_ANONYMOUS | (0x40000 if hugetlb else 0)  # MAP_HUGETLB
    p = libc.mmap(None, SIZE, mmap.PROT_READ | mmap.PROT_WRITE, flags, -1, 0)
    if p in (None, ctypes.c_void_p(-1).value):
        return None
    if not hugetlb:
        libc.madvise(p, SIZE, 14)  # MADV_HUGEPAGE
    ctypes.memset(p, 1, SIZE)  # fault the pages in (as huge pages when THP is on)
    rc = rt.cudaHostRegister(p, SIZE, 0)
    if int(rc) != 0:
        return None
    t = torch.frombuffer((ctypes.c_uint8 * SIZE).from_address(p), dtype=torch.uint8)
    assert t.is_pinned()
    return t


def measure(host, direction):
    s = torch.cuda.Stream()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    def copy():
        with torch.cuda.stream(s):
            if direction == "d2h":
                host.copy_(dev, non_blocking=True)
            else:
                dev.copy_(host, non_blocking=True)
    copy(); s.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ev0.record(s)
    for _ in range(4):
        copy()
    ev1.record(s); s.synchronize()
    t = torch.tensor([ev0.elapsed_time(ev1) * 1e-3], device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return world * 4 * SIZE / t.item() / 1e9


for name, make in (("cudaHostAlloc", buf_hostalloc), ("cudaHostRegister over THP (madvise)", lambda: buf_registered(False)),
                   ("cudaHostRegister over MAP_HUGETLB", lambda: buf_registered(True))):
    ptr = make()
    ok = torch.tensor([1 if ptr is not None else 0], device="cuda")
    if world > 1:
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    if not ok.item():
        if rank == 0:
            print(f"{name:40s} unavailable on this box", flush=True)
        continue
    d2h, h2d = measure(ptr, "d2h"), measure(ptr, "h2d")
    if rank == 0:
        thp = open("/sys/kernel/mm/transparent_hugepage/enabled").read().strip() if os.path.exists("/sys/kernel/mm/transparent_hugepage/enabled") else "?"
        print(f"{name:40s} {world} ranks at once: D2H {d2h:7.1f} GB/s  H2D {h2d:7.1f} GB/s   (THP: {thp})", flush=True)
if world > 1:
    dist.destroy_process_group()
