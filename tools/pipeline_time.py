"""Frame pipelining experiment: RCAS of frame i on stream B overlaps EASU of frame i+1 on stream A.
   FSR1_EASU_CTAS_PER_SM=k python tools/pipeline_time.py"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fsr1_b200 as F
api = F.api
iw, ih, ow, oh = 1920, 1080, 3840, 2160
R = 8
ins = [torch.from_numpy(F.to_half(F.uniform(iw, ih, 12345 + t))).cuda() for t in range(R)]
tmps = [torch.empty((oh, ow, 4), dtype=torch.float16, device="cuda") for _ in range(R)]
outs = [torch.empty((oh, ow, 4), dtype=torch.float16, device="cuda") for _ in range(R)]
econ, rcon = api.easu_con(iw, ih, iw, ih, ow, oh), api.rcas_con(0.25)
prio = os.environ.get("FSR1_PIPE_PRIO")     # e.g. "0,-1": RCAS stream more urgent than the EASU stream
if prio:
    pa, pb = (int(v) for v in prio.split(","))
    sa, sb = torch.cuda.Stream(priority=pa), torch.cuda.Stream(priority=pb)
else:
    sa, sb = torch.cuda.Stream(), torch.cuda.Stream()
easu_done = [torch.cuda.Event() for _ in range(R)]
rcas_done = [torch.cuda.Event() for _ in range(R)]
def run(n):
    for i in range(n):
        j = i % R
        if i >= R:
            sa.wait_event(rcas_done[j])          # tmp[j] is free again
        api.easu(ins[j], tmps[j], econ, stream=sa)
        easu_done[j].record(sa)
        sb.wait_event(easu_done[j])
        api.rcas(tmps[j], outs[j], rcon, stream=sb)
        rcas_done[j].record(sb)
run(40); torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 400
a.record(sa); sb.wait_event(a)
run(n)
sa.wait_stream(sb); b.record(sa); torch.cuda.synchronize()
us = a.elapsed_time(b) / n * 1e3
print("EASU ctas/sm cap %s: pipelined %.1f us/frame (%.0f Mpix/s)" % (os.environ.get("FSR1_EASU_CTAS_PER_SM", "-"), us, ow * oh / us))
