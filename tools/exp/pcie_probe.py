"""Host<->device copy probe (developer tool): 201 MB H2D / D2H alone and concurrently, default pinned vs write-combined
pinned source memory.  Tells how far bench.py's e2e leg is from the link."""
import ctypes, time, torch
rt = ctypes.CDLL("libcudart.so")
N = 64 * 3 * 512 * 512 * 4
dev = torch.device("cuda")
def host_alloc(flags):
    p = ctypes.c_void_p()
    assert rt.cudaHostAlloc(ctypes.byref(p), ctypes.c_size_t(N), ctypes.c_uint(flags)) == 0
    buf = (ctypes.c_float * (N // 4)).from_address(p.value)
    return torch.frombuffer(buf, dtype=torch.float32), p
d_in, d_out = torch.empty(N // 4, device=dev), torch.rand(N // 4, device=dev)
h_out = torch.empty(N // 4).pin_memory()
s_up, s_dn = torch.cuda.Stream(), torch.cuda.Stream()
def timeit(fn, n=10):
    fn(); torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t) / n
for name, flags in (("default pinned", 0), ("write-combined", 4)):
    h_in, keep = host_alloc(flags)
    h_in[:1024] = 1.0
    print(name, "is_pinned:", h_in.is_pinned())
    def up():
        with torch.cuda.stream(s_up): d_in.copy_(h_in, non_blocking=True)
    def dn():
        with torch.cuda.stream(s_dn): h_out.copy_(d_out, non_blocking=True)
    def both(): up(); dn()
    tu, td, tb = timeit(up), timeit(dn), timeit(both)
    print(f"  H2D alone {N/tu/1e9:6.1f} GB/s   D2H alone {N/td/1e9:6.1f} GB/s   both at once {N/tb/1e9:6.1f} GB/s each way ({tb*1e3:.2f} ms per 201 MB pair)")
