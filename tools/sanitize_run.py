"""Small invocations of every kernel family, for compute-sanitizer (memcheck / racecheck / initcheck):
   compute-sanitizer --tool memcheck python tools/sanitize_run.py"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fsr1_b200 as F
api = F.api
def run(iw, ih, ow, oh, dt, flags=0):
    src = F.structured(iw, ih, 5)
    src = src.astype(np.float16) if dt == torch.float16 else src
    wp = (iw + 1) & ~1
    a = torch.zeros((ih, wp, 4), dtype=dt, device="cuda"); a[:, :iw] = torch.from_numpy(src).cuda()
    owp = (ow + 1) & ~1
    t = torch.zeros((oh, owp, 4), dtype=dt, device="cuda"); o = torch.zeros((oh, owp, 4), dtype=dt, device="cuda")
    api.upscale(a[:, :iw], t[:, :ow], o[:, :ow], api.easu_con(iw, ih, iw, ih, ow, oh), api.rcas_con(0.25), flags=flags)
    torch.cuda.synchronize()
    return api.last_kernel()
names = set()
for dt in (torch.float16, torch.float32):
    for shape in ((150, 70, 300, 140), (150, 70, 225, 105), (151, 71, 197, 93), (33, 9, 66, 18)):
        for fl in (0, api.FLAG_RCAS_CLAMP, api.FLAG_FORCE_DIRECT, api.FLAG_PRECISE):
            names.add(run(*shape, dt, fl))
names.add(run(150, 70, 300, 140, torch.float16, api.FLAG_H_REFERENCE))
names.add(run(150, 70, 300, 140, torch.float32, api.FLAG_EXACT | api.FLAG_RCAS_DENOISE | api.FLAG_RCAS_PASSTHROUGH_ALPHA))
print("ran", api.launch_count(), "launches;", sorted(names))
