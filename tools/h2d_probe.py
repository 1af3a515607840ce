"""Host-to-device copy floor of the e2e step: one pinned block vs the ~25 arrays of gfpl_compact_input (B200: 55.5 vs 54.8 GB/s,
7.5 ms for 417 MB — the e2e step takes 8.1 ms).  python tools/h2d_probe.py"""
import torch, time
torch.cuda.set_device(0)
total = 417_000_000
sizes_many = [32768]*6 + [19_660_800, 19_660_800, 2_457_600, 2_457_600, 78_643_200, 78_643_200, 16_384_000, 16_384_000, 819_200, 819_200, 26_214_400, 26_214_400, 26_542_080, 2_211_840, 70_778_880, 10_747_904, 335_872, 10_747_904, 6_946_816]
def bench(sizes, reps=20):
    hs = [torch.empty(s, dtype=torch.uint8).pin_memory() for s in sizes]
    ds = [torch.empty(s, dtype=torch.uint8, device="cuda") for s in sizes]
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        for _ in range(3):
            for h, d in zip(hs, ds): d.copy_(h, non_blocking=True)
        st.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            for h, d in zip(hs, ds): d.copy_(h, non_blocking=True)
        st.synchronize()
        dt = (time.perf_counter() - t0) / reps
    return sum(sizes) / dt / 1e9, dt * 1e3
print("one block  : %.1f GB/s %.2f ms" % bench([sum(sizes_many)]))
print("%d blocks  : %.1f GB/s %.2f ms" % ((len(sizes_many),) + bench(sizes_many)))
