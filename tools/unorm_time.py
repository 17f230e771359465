"""Times EASU and RCAS on R8G8B8A8_UNORM images at 1080p -> 4K (the formats the sample renders into).
   python tools/unorm_time.py"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fsr1_b200 as F
api = F.api
iw, ih, ow, oh = 1920, 1080, 3840, 2160
R = 8
ins = [torch.randint(0, 256, (ih, iw, 4), dtype=torch.uint8, device="cuda") for _ in range(R)]
tmps = [torch.empty((oh, ow, 4), dtype=torch.uint8, device="cuda") for _ in range(R)]
outs = [torch.empty((oh, ow, 4), dtype=torch.uint8, device="cuda") for _ in range(R)]
econ, rcon = api.easu_con(iw, ih, iw, ih, ow, oh), api.rcas_con(0.25)


def timeit(fn, n=100):
    for i in range(10): fn(i)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(n): fn(i)
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3


te = timeit(lambda i: api.easu(ins[i % R], tmps[i % R], econ)); ke = api.last_kernel()
tr = timeit(lambda i: api.rcas(tmps[i % R], outs[i % R], rcon)); kr = api.last_kernel()
tb = timeit(lambda i: api.upscale(ins[i % R], tmps[i % R], outs[i % R], econ, rcon))
print("RGBA8 1080p->4K | %s %.1f us | %s %.1f us | both %.1f us (%.0f Mpix/s)" % (ke, te, kr, tr, tb, ow * oh / tb))
# agreement with the exact (fp32, direct) kernels: code values
ref_t, ref_o = torch.empty_like(tmps[0]), torch.empty_like(outs[0])
api.upscale(ins[0], ref_t, ref_o, econ, rcon, flags=api.FLAG_EXACT)
api.upscale(ins[0], tmps[0], outs[0], econ, rcon)
torch.cuda.synchronize()
d = (outs[0].int() - ref_o.int()).abs()
print("vs FSR1_FLAG_EXACT end to end: max %d code values, %.2f %% of values differ" % (int(d.max()), 100.0 * float((d > 0).float().mean())))
