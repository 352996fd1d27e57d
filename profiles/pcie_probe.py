"""Probe: pinned-host <-> device copy bandwidth on this box (the e2e leg's roofline)."""
import torch
dev = 'cuda:0'
n = 700 * 1000 * 1000 // 4
h1 = torch.empty(n).pin_memory(); h2 = torch.empty(n).pin_memory()
d1 = torch.empty(n, device=dev); d2 = torch.empty(n, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def t(f, it=5):
    f(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(it): f()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / it
gb = n * 4 / 1e9
ms = t(lambda: d1.copy_(h1, non_blocking=True)); print(f'H2D  {gb/ms*1e3:6.1f} GB/s ({ms:.2f} ms for {gb:.2f} GB)')
ms = t(lambda: h2.copy_(d2, non_blocking=True)); print(f'D2H  {gb/ms*1e3:6.1f} GB/s')
def both():
    cur = torch.cuda.current_stream()
    s1.wait_stream(cur); s2.wait_stream(cur)
    with torch.cuda.stream(s1): d1.copy_(h1, non_blocking=True)
    with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
    cur.wait_stream(s1); cur.wait_stream(s2)
ms = t(both); print(f'both directions concurrently: {2*gb/ms*1e3:6.1f} GB/s total ({ms:.2f} ms)')
