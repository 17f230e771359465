"""Where do the TMA-tiled R8G8B8A8 kernels differ from quantise(oracle(dequantise(input))) by more than one code?
   python tools/dbg_u8.py [iw ih]      (GPU; prints the error positions and their place inside the 64x16 tiles)"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import fsr1_b200 as F, oracle_lib as ol
api = F.api
iw, ih = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (1920, 1080)
ow, oh = 2 * iw, 2 * ih
raw = np.floor(F.uniform(iw, ih, 12345) * 255.0 + 0.5).astype(np.uint8)
fin = raw.astype(np.float32) / np.float32(255.0)
q = lambda x: (np.clip(x, 0, 1).astype(np.float32) * np.float32(255) + np.float32(0.5)).astype(np.int64)
want = q(ol.easu(fin, ow, oh)[..., :3])
din = torch.from_numpy(raw).cuda()
for flags, tag in ((0, "tiled"), (api.FLAG_FORCE_DIRECT, "direct")):
    tmp = torch.zeros((oh, ow, 4), dtype=torch.uint8, device="cuda")
    api.easu(din, tmp, api.easu_con(iw, ih, iw, ih, ow, oh), flags=flags)
    torch.cuda.synchronize()
    got = tmp.cpu().numpy()[..., :3].astype(np.int64)
    d = np.abs(got - want)
    ys, xs, cs = np.where(d > 1)
    print(tag, api.last_kernel(), "max", int(d.max()), "count>1", len(ys), "frac>0 %.4f" % float((d > 0).mean()))
    for y, x, c in list(zip(ys, xs, cs))[:12]:
        print("   y %d x %d c %d got %d want %d | cell row %d (tile row %d, in-tile %d)  cell col %d (tile col %d, lane %d)" % (
            y, x, c, got[y, x, c], want[y, x, c], (y - 1) // 2, ((y - 1) // 2 + 1) // 8, ((y - 1) // 2 + 1) % 8, (x - 1) // 2,
            ((x - 1) // 2 + 1) // 32, ((x - 1) // 2 + 1) % 32))
    if len(ys):
        print("   distinct rows:", sorted(set(ys.tolist()))[:20], " distinct cols mod 64:", sorted(set(((xs - 1) % 64).tolist()))[:20])
