import sys, os, numpy as np, torch
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import fsr1_b200 as F, oracle_lib as ol
api = F.api
def stats(name, got, want):
    d = np.abs(got.astype(np.float32) - want)[..., :3]
    i = np.unravel_index(np.argmax(d), d.shape)
    print("%-28s max %.5f at (y=%d,x=%d,c=%d) got %.5f want %.5f | >1e-2: %d  >5e-3: %d  mean %.2e" % (
        name, d.max(), i[0], i[1], i[2], got[i[0], i[1], i[2]], want[i[0], i[1], i[2]], (d > 1e-2).sum(), (d > 5e-3).sum(), d.mean()))
for gen in ("uniform", "structured"):
    for (iw, ih, ow, oh) in [(192, 108, 384, 216), (192, 108, 288, 162), (192, 108, 250, 141), (1920, 1080, 3840, 2160)]:
        src = F.to_half(getattr(F, gen)(iw, ih, 7 if iw < 1000 else 2024))
        din = torch.from_numpy(src).cuda()
        ow_p = (ow + 1) & ~1
        tmp = torch.zeros((oh, ow_p, 4), dtype=torch.float16, device="cuda")[:, :ow]
        out = torch.zeros((oh, ow_p, 4), dtype=torch.float16, device="cuda")[:, :ow]
        econ, rcon = api.easu_con(iw, ih, iw, ih, ow, oh), api.rcas_con(0.25)
        api.easu(din, tmp, econ); k1 = api.last_kernel()
        api.rcas(tmp, out, rcon); k2 = api.last_kernel()
        torch.cuda.synchronize()
        e_want = ol.easu(src.astype(np.float32), ow, oh)
        e_got = tmp.cpu().numpy()
        print(gen, (iw, ih, ow, oh), k1, k2)
        stats("  easu vs f32 oracle", e_got, e_want)
        stats("  rcas stage (same input)", out.cpu().numpy(), ol.rcas(np.ascontiguousarray(e_got).astype(np.float32), ol.rcas_con(0.25)))
        stats("  e2e vs f32 e2e", out.cpu().numpy(), ol.rcas(e_want, ol.rcas_con(0.25)))
