import time, torch, numpy as np, os, threading
dev = torch.device("cuda", 0)
n = 512 * 1024 * 1024 // 4
pageable = torch.empty(n, dtype=torch.float32); pageable.fill_(1.0)
pinned = torch.empty(n, dtype=torch.float32).pin_memory(); pinned.fill_(1.0)
d = torch.empty(n, dtype=torch.float32, device=dev)
def t(f, reps=3):
    f(); torch.cuda.synchronize()
    b = time.perf_counter()
    for _ in range(reps): f()
    torch.cuda.synchronize()
    return (time.perf_counter() - b) / reps
print("cores", os.cpu_count())
print("H2D pinned   GB/s", 0.5368 / t(lambda: d.copy_(pinned, non_blocking=True)))
print("H2D pageable GB/s", 0.5368 / t(lambda: d.copy_(pageable)))
print("D2H pinned   GB/s", 0.5368 / t(lambda: pinned.copy_(d, non_blocking=True)))
print("D2H pageable GB/s", 0.5368 / t(lambda: pageable.copy_(d)))
src = pageable.numpy(); dst = pinned.numpy()
def mt_copy(threads):
    parts = np.array_split(np.arange(0, n + 1, max(1, n // threads))[:threads + 1], 1)[0]
    bounds = np.linspace(0, n, threads + 1).astype(np.int64)
    ts = [threading.Thread(target=lambda a, b: np.copyto(dst[a:b], src[a:b]), args=(bounds[i], bounds[i + 1])) for i in range(threads)]
    b = time.perf_counter()
    for x in ts: x.start()
    for x in ts: x.join()
    return time.perf_counter() - b
for th in (1, 2, 4, 8, 16):
    mt_copy(th)
    print("memcpy pageable->pinned threads", th, "GB/s", round(0.5368 / min(mt_copy(th) for _ in range(3)), 1))
