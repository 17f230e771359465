"""Times EASU and RCAS separately (CUDA events, ring of frames > L2) and reports per-stage max error.
Usage: python tools/variant_time.py [2x|1.5x|1.3x]"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import fsr1_b200 as F, oracle_lib as ol
api = F.api
wl = sys.argv[1] if len(sys.argv) > 1 else "2x"
iw, ih, ow, oh = {"2x": (1920, 1080, 3840, 2160), "1.5x": (2560, 1440, 3840, 2160), "1.3x": (2953, 1661, 3840, 2160)}[wl]
R = 8
ins = [torch.from_numpy(F.to_half(F.uniform(iw, ih, 12345 + t))).cuda() for t in range(R)]
iw_p = (iw + 1) & ~1
if iw_p != iw:
    padded = []
    for t in ins:
        b = torch.zeros((ih, iw_p, 4), dtype=torch.float16, device="cuda"); b[:, :iw] = t; padded.append(b[:, :iw])
    ins = padded
tmps = [torch.empty((oh, ow, 4), dtype=torch.float16, device="cuda") for _ in range(R)]
outs = [torch.empty((oh, ow, 4), dtype=torch.float16, device="cuda") for _ in range(R)]
econ, rcon = api.easu_con(iw, ih, iw, ih, ow, oh), api.rcas_con(0.25)
def timeit(fn, n=200):
    for i in range(20): fn(i)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(n): fn(i)
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3
EF = int(os.environ.get("FSR1_TEST_EASU_FLAGS", "0"))
te = timeit(lambda i: api.easu(ins[i % R], tmps[i % R], econ, flags=EF)); ke = api.last_kernel()
tr = timeit(lambda i: api.rcas(tmps[i % R], outs[i % R], rcon)); kr = api.last_kernel()
tb = timeit(lambda i: api.upscale(ins[i % R], tmps[i % R], outs[i % R], econ, rcon, flags=EF))
src = ins[0].cpu().numpy()
e_want = ol.easu(np.ascontiguousarray(src).astype(np.float32), ow, oh)
e_got = tmps[0].cpu().numpy()
r_want = ol.rcas(e_got.astype(np.float32), ol.rcas_con(0.25))
e2e = ol.rcas(e_want, ol.rcas_con(0.25))
print("e2e max err %.5f" % np.abs(outs[0].cpu().numpy().astype(np.float32) - e2e).max())
print("%s | %s %.1f us | %s %.1f us | both %.1f us (%.0f Mpix/s) | err easu %.4f rcas %.4f" % (
    wl, ke, te, kr, tr, tb, ow * oh / tb, np.abs(e_got.astype(np.float32) - e_want).max(),
    np.abs(outs[0].cpu().numpy().astype(np.float32) - r_want).max()))
