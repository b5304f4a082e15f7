"""pinned vs write-combined pinned host memory: H2D bandwidth (run on the GPU box)"""
import ctypes as C, time, torch
torch.cuda.init()
rt = C.CDLL("libcudart.so")
dev = torch.empty(1 << 30, dtype=torch.uint8, device="cuda")
for flags, name in ((0, "default pinned"), (4, "write-combined"), (1, "portable")):
    p = C.c_void_p()
    assert rt.cudaHostAlloc(C.byref(p), C.c_size_t(1 << 30), C.c_uint(flags)) == 0
    C.memset(p, 1, 1 << 30)
    for n in (186 << 20, 16 << 20):
        reps = (1 << 30) // n
        for rep in range(2):
            torch.cuda.synchronize(); t = time.perf_counter()
            for i in range(reps):
                rt.cudaMemcpyAsync(C.c_void_p(dev.data_ptr() + i * n), C.c_void_p(p.value + i * n), C.c_size_t(n), 1, None)
            torch.cuda.synchronize(); dt = time.perf_counter() - t
        print(f"{name:16s} copy {n >> 20:4d} MB x {reps}: {reps * n / dt / 1e9:6.1f} GB/s")
    rt.cudaFreeHost(p)
