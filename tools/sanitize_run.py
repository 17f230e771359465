"""Small invocations of every kernel family, for compute-sanitizer (memcheck / racecheck / initcheck):
   compute-sanitizer --tool memcheck python tools/sanitize_run.py"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
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
# UNORM images (TMA-tiled EASU + packed RCAS at 2x, direct kernels otherwise)
for shape in ((150, 70, 300, 140), (150, 70, 225, 105)):
    iw, ih, ow, oh = shape
    a = torch.randint(0, 256, (ih, iw + (-iw) % 4, 4), dtype=torch.uint8, device="cuda")[:, :iw]
    t = torch.zeros((oh, ow, 4), dtype=torch.uint8, device="cuda"); o = torch.zeros_like(t)
    api.upscale(a, t, o, api.easu_con(iw, ih, iw, ih, ow, oh), api.rcas_con(0.25)); torch.cuda.synchronize()
    names.add(api.last_kernel())
# the packed Hx2 calling convention: widths ending inside a strip, both out-of-image rules, options, tiles that wrap
for (w, h) in ((300, 11), (37, 9), (5, 4)):
    img = torch.from_numpy(F.to_half(F.structured(w, h, 7))).cuda()
    out = torch.zeros_like(img)
    for fl in (0, api.FLAG_RCAS_CLAMP, api.FLAG_RCAS_DENOISE | api.FLAG_RCAS_PASSTHROUGH_ALPHA):
        api.rcas(img, out, api.rcas_con(0.25), flags=api.FLAG_RCAS_HX2 | fl)
    names.add(api.last_kernel())
    grain = torch.from_numpy(F.to_half(F.uniform(12, 5, 3) - 0.5)).cuda()
    api.srtm_h(img, out); api.srtm_h(out, out, inverse=True); api.lfga_h(img, grain, out, 0.5)
    api.tepd_h(img, out, 8, frame=3); api.tepd_h(img, out, 10, dither=grain)
    torch.cuda.synchronize()
    names.add(api.last_kernel())
# row-slab windows whose height is not a multiple of the 4 rows an RCAS lane walks (the 8-GPU split of 2160 rows is 270),
# through the sharded data plane: 3 ranks on this device, halo by direct stores, two frames per slot
for dt in (torch.float16, torch.float32):
    iw, ih, ow, oh, world = 160, 90, 320, 180 + 2, 3
    ups = [F.ShardedUpscaler(iw, ih, ow, oh, world, r, dtype=dt, slots=2, attach=False) for r in range(world)]
    for r, u in enumerate(ups):
        u.attach_local(ups[r - 1] if r else None, ups[r + 1] if r + 1 < world else None)
    s = torch.cuda.current_stream()
    for i in range(4):
        for r, u in enumerate(ups):
            o0, o1 = u.plan.owned_in_rows(r)
            u.wait(i % 2, s)
            u.input(i % 2).copy_(torch.from_numpy(F.structured(iw, ih, 40 + i)[o0:o1]).to(dt).cuda())
        for u in ups:
            u.submit(i % 2, s)
    for u in ups:
        u.wait(0, s); u.wait(1, s)
    torch.cuda.synchronize()
    for u in ups:
        u.status(); u.close()
    names.add("fsr1_shard x%d (%s)" % (world, str(dt)))
print("ran", api.launch_count(), "launches;", sorted(names))
