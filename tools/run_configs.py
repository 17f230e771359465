"""Runs the single-GPU configurations of BASELINE.json (configs[0..3] + the true 1.3x case) and writes one JSON table.
   python tools/run_configs.py > profiles/rNN_configs.json        (on the GPU box)"""
import json, os, subprocess, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import fsr1_b200 as F, oracle_lib as ol
import torch
api = F.api
out = {}

# configs[0]: 540p -> 1080p EASU+RCAS fp32, single synthetic frame, CPU only (the oracle / reference build)
iw, ih, ow, oh = 960, 540, 1920, 1080
frame = F.uniform(iw, ih, 12345)
rows = {}
for name, lib in (("reference_source_on_host", ol.ref()), ("oracle_port", ol.oracle())):
    if lib is None:
        continue
    ol.rcas(ol.easu(frame, ow, oh, lib=lib), ol.rcas_con(0.25), lib=lib)
    t = time.perf_counter(); e = ol.easu(frame, ow, oh, lib=lib); te = time.perf_counter() - t
    t = time.perf_counter(); r = ol.rcas(e, ol.rcas_con(0.25), lib=lib); tr = time.perf_counter() - t
    rows[name] = {"easu_ms": te * 1e3, "rcas_ms": tr * 1e3, "mpix_per_s": ow * oh / (te + tr) / 1e6, "threads": os.cpu_count(),
                  "checksum_rgb": float(r[..., :3].astype(np.float64).sum())}
# the same frame on the GPU, exact mode: bit-identical to the oracle
g = torch.zeros((oh, ow, 4), dtype=torch.float32, device="cuda"); g2 = torch.zeros_like(g)
api.easu(torch.from_numpy(frame).cuda(), g, api.easu_con(iw, ih, iw, ih, ow, oh), flags=api.FLAG_EXACT)
api.rcas(g, g2, api.rcas_con(0.25), flags=api.FLAG_EXACT); torch.cuda.synchronize()
rows["gpu_exact_bit_identical_to_oracle"] = bool(np.array_equal(g2.cpu().numpy().view(np.uint32), ol.rcas(ol.easu(frame, ow, oh), ol.rcas_con(0.25)).view(np.uint32)))
out["configs[0] 540p->1080p fp32 single frame (CPU)"] = rows

def bench(*a):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--no-cpu"] + list(a), capture_output=True, text=True)
    d = json.loads(r.stdout.strip().splitlines()[-1])
    return {"mpix_per_s": d["value"], "us_per_frame": d["ms_per_step"] * 1e3, "kernels": {k: {"kernel": v.get("kernel"), "us": v["us"], "frac_of_hbm_peak": v["frac_of_hbm_peak"]} for k, v in d["kernels"].items()},
            "e2e_mpix_per_s": d["e2e"]["value"], "clocks": d["clocks"], "steps": d["steps"]}
out["configs[1] 1080p->4K fp16"] = bench("--workload", "1080p-4k-fp16", "--steps", "200", "--warmup", "20")
out["configs[2] 1440p->4K fp16, 120-frame stream"] = bench("--workload", "1440p-4k-fp16", "--frames", "120", "--steps", "120", "--warmup", "120")
out["configs[2'] 2953x1661->4K (true 1.3x) fp16"] = bench("--workload", "uq-4k-fp16", "--steps", "100", "--warmup", "10")
out["configs[3] 1080p->4K fp32"] = bench("--workload", "1080p-4k-fp32", "--steps", "100", "--warmup", "10")

# configs[3] tolerance + sharpness sweep: fp16 kernels against the fp32 oracle on the same (quantised) frame
iw, ih, ow, oh = 1920, 1080, 3840, 2160
src = F.to_half(F.uniform(iw, ih, 12345))
e_want = ol.easu(src.astype(np.float32), ow, oh)
sweep = {}
din = torch.from_numpy(src).cuda()
tmp = torch.zeros((oh, ow, 4), dtype=torch.float16, device="cuda"); o = torch.zeros_like(tmp)
for sharp in (0.0, 0.25, 1.0, 2.0):
    api.upscale(din, tmp, o, api.easu_con(iw, ih, iw, ih, ow, oh), api.rcas_con(sharp)); torch.cuda.synchronize()
    eg = tmp.cpu().numpy().astype(np.float32)
    sweep["sharpness_%g" % sharp] = {
        "easu_max_abs": float(np.abs(eg - e_want).max()),
        "rcas_max_abs_same_input": float(np.abs(o.cpu().numpy().astype(np.float32) - ol.rcas(eg, ol.rcas_con(sharp))).max()),
        "end_to_end_max_abs": float(np.abs(o.cpu().numpy().astype(np.float32) - ol.rcas(e_want, ol.rcas_con(sharp))).max())}
out["configs[3] fp16 vs fp32-oracle tolerance, sharpness sweep (1080p->4K, LCG noise)"] = sweep
print(json.dumps(out, indent=1))
