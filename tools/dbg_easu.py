import sys, os, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import fsr1_b200 as F, oracle_lib as ol
api = F.api
iw, ih, ow, oh = map(int, sys.argv[1:5])
src = F.to_half(F.uniform(iw, ih, 33))
out = torch.zeros((oh, ow, 4), dtype=torch.float16, device="cuda")
api.easu(torch.from_numpy(src).cuda(), out, api.easu_con(iw, ih, iw, ih, ow, oh))
torch.cuda.synchronize()
want = ol.easu(src.astype(np.float32), ow, oh)
print(sys.argv[1:5], api.last_kernel(), "maxerr", np.abs(out.cpu().numpy().astype(np.float32) - want).max())
