"""A few launches of the two production kernels at the bench configuration, for ncu:
   ncu --set full --clock-control none --import-source on -o gpurun_out/prof python tools/profile_run.py 2x 2
"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fsr1_b200 as F
api = F.api
wl = sys.argv[1] if len(sys.argv) > 1 else "2x"
iw, ih, ow, oh = {"2x": (1920, 1080, 3840, 2160), "1.5x": (2560, 1440, 3840, 2160), "1.3x": (2953, 1661, 3840, 2160)}[wl]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 4
sets = []
for t in range(n):
    sets.append((torch.from_numpy(F.to_half(F.uniform(iw, ih, 12345 + t))).cuda(),
                 torch.empty((oh, ow, 4), dtype=torch.float16, device="cuda"),
                 torch.empty((oh, ow, 4), dtype=torch.float16, device="cuda")))
econ, rcon = api.easu_con(iw, ih, iw, ih, ow, oh), api.rcas_con(0.25)
for a, t, b in sets:
    api.upscale(a, t, b, econ, rcon)
if wl == "2x":
    for a, t, b in sets:
        api.upscale(a, t, b, econ, rcon, flags=api.FLAG_FUSED)          # the fused EASU->RCAS kernel
    u8 = [torch.randint(0, 256, (ih, iw, 4), dtype=torch.uint8, device="cuda") for _ in range(2)]
    t8, o8 = torch.empty((oh, ow, 4), dtype=torch.uint8, device="cuda"), torch.empty((oh, ow, 4), dtype=torch.uint8, device="cuda")
    for a in u8:
        api.upscale(a, t8, o8, econ, rcon)                                # R8G8B8A8: TMA-tiled EASU + packed RCAS
torch.cuda.synchronize()
print("done", api.launch_count(), api.last_kernel())
