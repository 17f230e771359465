"""Times the pointwise companions (SRTM, SRTM inverse, LFGA, TEPD, square) at 3840x2160 against the HBM roofline.
Algorithmic bytes per pixel = bytes read + bytes written (in-format + out-format).  Ring of frames > L2.
Usage: python tools/pointwise_time.py [out.json]"""
import json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fsr1_b200 as F
api = F.api
W, H, R = 3840, 2160, 6
peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
PEAK = float(peaks.get("hbm_gbs", 6587.7))


def timeit(fn, n=100):
    for i in range(10): fn(i)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(n): fn(i)
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3


rows = []
grain16 = (torch.rand((128, 128, 4), device="cuda") - 0.5).half()
HALF_ONLY = "--half-only" in sys.argv
sys.argv = [a for a in sys.argv if a != "--half-only"]
for dt, bpp in (((torch.float16, 8),) if HALF_ONLY else ((torch.float16, 8), (torch.float32, 16))):
    ins = [torch.rand((H, W, 4), device="cuda").to(dt) for _ in range(R)]
    outs = [torch.empty_like(ins[0]) for _ in range(R)]
    grain = grain16.to(dt)
    u8 = [torch.empty((H, W, 4), dtype=torch.uint8, device="cuda") for _ in range(R)]
    u10 = [torch.empty((H, W), dtype=torch.int32, device="cuda") for _ in range(R)]
    cases = [("srtm", lambda i: api.srtm(ins[i % R], outs[i % R]), 2 * bpp),
             ("srtm_inv", lambda i: api.srtm(ins[i % R], outs[i % R], inverse=True), 2 * bpp),
             ("lfga", lambda i: api.lfga(ins[i % R], grain, outs[i % R], 0.3), 2 * bpp),
             ("tepd8", lambda i: api.tepd(ins[i % R], outs[i % R], 8, frame=i), 2 * bpp),
             ("tepd8->rgba8", lambda i: api.tepd(ins[i % R], u8[i % R], 8, frame=i), bpp + 4),
             ("tepd10->rgb10a2", lambda i: api.tepd(ins[i % R], u10[i % R], 10, frame=i), bpp + 4)]
    for name, fn, bytes_px in cases:
        us = timeit(fn)
        gbs = W * H * bytes_px / us * 1e-3
        rows.append({"op": name, "storage": str(dt).split(".")[-1], "us": round(us, 2), "bytes_per_px": bytes_px,
                     "achieved_gbs": round(gbs, 1), "frac_of_hbm_peak": round(gbs / PEAK, 3), "kernel": api.last_kernel()})
        print(rows[-1])
    del ins, outs, u8, u10
if len(sys.argv) > 1:
    json.dump({"size": [W, H], "hbm_peak_gbs": PEAK, "rows": rows}, open(sys.argv[1], "w"), indent=1)
