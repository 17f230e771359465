"""What connects the GPUs of this box, and what a small peer write costs (run with >= 2 GPUs visible)."""
import subprocess, sys, time
import torch
print(subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True).stdout[:2500])
n = torch.cuda.device_count()
print("devices", n, "peer access 0->1", torch.cuda.can_device_access_peer(0, 1) if n > 1 else None)
if n > 1:
    a = torch.zeros(61440 // 2, dtype=torch.float16, device="cuda:0")
    b = torch.zeros(61440 // 2, dtype=torch.float16, device="cuda:1")
    big0 = torch.zeros(64 << 20, dtype=torch.uint8, device="cuda:0"); big1 = torch.zeros(64 << 20, dtype=torch.uint8, device="cuda:1")
    for src, dst, tag in ((a, b, "61 KB"), (big0, big1, "64 MB")):
        for _ in range(5): dst.copy_(src)
        torch.cuda.synchronize(0); torch.cuda.synchronize(1)
        with torch.cuda.device(0):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(50): dst.copy_(src)
            e1.record(); torch.cuda.synchronize(0); torch.cuda.synchronize(1)
        us = e0.elapsed_time(e1) / 50 * 1e3
        print("peer copy %s: %.1f us each (%.1f GB/s)" % (tag, us, src.numel() * src.element_size() / us / 1e3))
