"""GPU-box experiment: is mmap + cudaHostRegister + DMA straight out of the page cache cheaper than read() into a pinned ring?
(scene loader: one 38.4 MB descriptor file per cloud).  python scripts/hostreg_test.py"""
import mmap, os, sys, threading, time
import numpy as np
import torch

rt = torch.cuda.cudart()
d = "/dev/shm/roreg_hostreg"; os.makedirs(d, exist_ok=True)
nb = 38_400_000; nf = 16
src = np.random.default_rng(0).integers(0, 255, nb, dtype=np.uint8)
for i in range(nf):
    with open(f"{d}/{i}.bin", "wb") as f:
        f.write(src)
dev = torch.empty((nf, nb), dtype=torch.uint8, device="cuda")
torch.cuda.synchronize()
side = torch.cuda.Stream()


def one(i, flags, prot_write=True, check=False):
    t0 = time.perf_counter()
    fd = os.open(f"{d}/{i}.bin", os.O_RDWR if prot_write else os.O_RDONLY)
    m = mmap.mmap(fd, nb, mmap.MAP_SHARED | getattr(mmap, "MAP_POPULATE", 0), mmap.PROT_READ | (mmap.PROT_WRITE if prot_write else 0))
    a = np.frombuffer(m, np.uint8)
    t1 = time.perf_counter()
    rc = rt.cudaHostRegister(a.ctypes.data, nb, flags)
    t2 = time.perf_counter()
    if int(rc) != 0:
        del a; m.close(); os.close(fd)
        return None
    t = torch.from_numpy(a) if prot_write else torch.frombuffer(m, dtype=torch.uint8)
    with torch.cuda.stream(side):
        dev[i].copy_(t, non_blocking=True)
    side.synchronize()
    t3 = time.perf_counter()
    rt.cudaHostUnregister(a.ctypes.data)
    t4 = time.perf_counter()
    ok = bool((dev[i].cpu().numpy() == src).all()) if check else None
    del t, a; m.close(); os.close(fd)
    return (t1 - t0, t2 - t1, t3 - t2, t4 - t3, ok)


for flags, pw, name in ((0, True, "RDWR mapping, default flags"), (8, False, "read-only mapping, cudaHostRegisterReadOnly")):
    r = one(0, flags, pw, check=True)
    print(name, "->", "register FAILED" if r is None else f"mmap {r[0]*1e3:.2f} ms, register {r[1]*1e3:.2f} ms, H2D {r[2]*1e3:.2f} ms ({nb/r[2]/1e9:.1f} GB/s), unregister {r[3]*1e3:.2f} ms, data ok {r[4]}")
    if r is None:
        continue
    for nt in (1, 2, 4, 8):
        idx = list(range(nf)); lock = threading.Lock()
        def w():
            while True:
                with lock:
                    if not idx:
                        return
                    i = idx.pop()
                one(i, flags, pw)
        t0 = time.perf_counter(); ts = [threading.Thread(target=w) for _ in range(nt)]
        [t.start() for t in ts]; [t.join() for t in ts]; dt = time.perf_counter() - t0
        print(f"   {nt} threads, {nf} files: {nf * nb / dt / 1e9:.1f} GB/s end to end")

# baseline: read() into pinned buffers + H2D
pin = [torch.empty(nb, dtype=torch.uint8).pin_memory() for _ in range(4)]
for nt in (1, 4, 8, 12):
    idx = list(range(nf)); lock = threading.Lock(); sem = threading.Semaphore(4); free = list(range(4))
    def w():
        while True:
            with lock:
                if not idx:
                    return
                i = idx.pop()
            sem.acquire()
            with lock:
                b = free.pop()
            with open(f"{d}/{i}.bin", "rb", buffering=0) as f:
                f.readinto(memoryview(pin[b].numpy()))
            s = torch.cuda.Stream()
            with torch.cuda.stream(s):
                dev[i].copy_(pin[b], non_blocking=True)
            s.synchronize()
            with lock:
                free.append(b)
            sem.release()
    t0 = time.perf_counter(); ts = [threading.Thread(target=w) for _ in range(nt)]
    [t.start() for t in ts]; [t.join() for t in ts]; dt = time.perf_counter() - t0
    print(f"read() into a 4-buffer pinned ring + H2D, {nt} threads: {nf * nb / dt / 1e9:.1f} GB/s end to end")
import shutil; shutil.rmtree(d)
