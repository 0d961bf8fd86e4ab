import time, torch, os
print("cpus", os.cpu_count(), "threads", torch.get_num_threads())
n = 1 << 28  # 268M entries -> 2 GiB of int64
src8 = torch.randint(0, 8, (n,), dtype=torch.uint8)
src16 = torch.randint(0, 1000, (n,), dtype=torch.int16)
dst = torch.empty(n, dtype=torch.int64).pin_memory()
for name, src in (("u8->i64", src8), ("i16->i64", src16)):
    for _ in range(3):
        t = time.perf_counter(); dst.copy_(src); dt = time.perf_counter() - t
    print(name, f"{dt*1e3:.1f} ms  write {n*8/dt/1e9:.1f} GB/s")
d = torch.empty(n, dtype=torch.int64, device="cuda")
for _ in range(3):
    torch.cuda.synchronize(); t = time.perf_counter(); dst.copy_(d); torch.cuda.synchronize(); dt = time.perf_counter() - t
print("D2H pinned", f"{dt*1e3:.1f} ms {n*8/dt/1e9:.1f} GB/s")
d32 = torch.empty(n, dtype=torch.int32, device="cuda"); h32 = torch.empty(n, dtype=torch.int32).pin_memory()
for _ in range(3):
    torch.cuda.synchronize(); t = time.perf_counter(); h32.copy_(d32, non_blocking=True); torch.cuda.synchronize(); dst.copy_(h32); dt = time.perf_counter() - t
print("D2H i32 + widen", f"{dt*1e3:.1f} ms  effective {n*8/dt/1e9:.1f} GB/s")
