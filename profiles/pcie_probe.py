"""Pinned-memory H2D / D2H / duplex bandwidth of the box (context for the e2e number of bench.py)."""
import torch, time
d=torch.device('cuda')
h=torch.empty(64<<20,dtype=torch.uint8).pin_memory(); g=torch.empty(64<<20,dtype=torch.uint8,device=d)
s1=torch.cuda.Stream(); s2=torch.cuda.Stream()
for name,fn in (('h2d',lambda: g.copy_(h,non_blocking=True)),('d2h',lambda: h.copy_(g,non_blocking=True))):
    fn(); torch.cuda.synchronize(); t0=time.perf_counter()
    for _ in range(10): fn()
    torch.cuda.synchronize(); dt=(time.perf_counter()-t0)/10
    print(name, 64/1024/dt, 'GiB/s')
h2=torch.empty(64<<20,dtype=torch.uint8).pin_memory(); g2=torch.empty(64<<20,dtype=torch.uint8,device=d)
torch.cuda.synchronize(); t0=time.perf_counter()
for _ in range(10):
    with torch.cuda.stream(s1): g.copy_(h,non_blocking=True)
    with torch.cuda.stream(s2): h2.copy_(g2,non_blocking=True)
torch.cuda.synchronize(); dt=(time.perf_counter()-t0)/10
print('duplex each', 64/1024/dt, 'GiB/s')

# ---- write-combined pinned memory as the H2D source (cudaHostAllocWriteCombined)
import ctypes
rt = None
for name in ("libcudart.so", "libcudart.so.12", "libcudart.so.13"):
    try:
        rt = ctypes.CDLL(name)
        break
    except OSError:
        pass
if rt is None:
    import glob, os
    import torch as _t
    cands = glob.glob(os.path.join(os.path.dirname(_t.__file__), "..", "nvidia", "cuda_runtime", "lib", "libcudart.so*"))
    rt = ctypes.CDLL(cands[0]) if cands else None
if rt is not None:
    n = 64 << 20
    for flag, label in ((0, "pinned default"), (4, "pinned write-combined")):
        ptr = ctypes.c_void_p()
        assert rt.cudaHostAlloc(ctypes.byref(ptr), ctypes.c_size_t(n), ctypes.c_uint(flag)) == 0
        ctypes.memset(ptr, 1, n)
        torch.cuda.synchronize()
        rt.cudaMemcpyAsync.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p]
        st = torch.cuda.current_stream().cuda_stream
        for _ in range(2):
            rt.cudaMemcpyAsync(ctypes.c_void_p(g.data_ptr()), ptr, n, 1, ctypes.c_void_p(st))
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(10):
            rt.cudaMemcpyAsync(ctypes.c_void_p(g.data_ptr()), ptr, n, 1, ctypes.c_void_p(st))
        torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 10
        print("h2d", label, 64 / 1024 / dt, "GiB/s")
        rt.cudaFreeHost(ptr)
