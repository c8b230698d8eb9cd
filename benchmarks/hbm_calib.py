"""Calibrates the box: device-to-device copy bandwidth (read + write bytes) like MEASURED_PEAKS.json."""
import torch, json
n = 1 << 30
a = torch.empty(n, dtype=torch.bfloat16, device="cuda"); b = torch.empty_like(a)
best = 0
for _ in range(10):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); b.copy_(a); e1.record(); torch.cuda.synchronize()
    best = max(best, 2 * n * 2 / (e0.elapsed_time(e1) * 1e-3) / 1e9)
print(json.dumps({"d2d_copy_gbs": best}))
