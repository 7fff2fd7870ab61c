"""Pinned host -> device copy bandwidth and pin_memory cost of one step's inputs (776 MB): explains box-to-box differences of bench.py's `e2e`
(55 GB/s on a healthy box; the first batch's copy is exposed, the following ones overlap the previous step)."""
import torch, time
x = torch.empty(194_000_000, dtype=torch.float32).pin_memory()
d = torch.empty_like(x, device="cuda")
for _ in range(2): d.copy_(x, non_blocking=True)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): d.copy_(x, non_blocking=True)
e1.record(); torch.cuda.synchronize()
print("pinned H2D GB/s:", 5 * x.numel() * 4 / e0.elapsed_time(e1) / 1e6)
t = time.perf_counter(); y = torch.empty(194_000_000, dtype=torch.float32).pin_memory(); print("pin_memory 776MB s:", time.perf_counter() - t)
