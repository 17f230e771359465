"""The metrics DESIGN.md and bench.py quote, out of an `.ncu-rep` (no GPU needed; `ncu` must be on PATH).

    python tools/ncu_summary.py REPORT.ncu-rep [--json-dir profiles] [--out-px 8294400]

Prints a per-kernel summary (the text committed as profiles/rNN_ncu_summary.txt).  With --json-dir also writes
  ncu_traffic.json  {display name: dram bytes read + written per launch}          -> bench.py roofline.traffic
  ncu_issue.json    {display name: {inst_per_px, inst_per_cycle_sm, issue_active_pct, fma_pipe_pct, ...}} -> roofline.issue
keyed by the kernel names the library reports (fsr1_last_kernel_name), taking the LAST launch of each kernel in the report
(the first ones are cold)."""
import csv
import json
import os
import re
import subprocess
import sys

# ncu's demangled function name -> the name the library reports
DISPLAY = [
    (r"easu_h_quad2x_kernel<4, 7>", "easu_h_quad2x<4w,7/sm,tma2>"),
    (r"easu_u_quad2x_kernel<4, 6, 8>", "easu_u8_quad2x<4w,6/sm,tma2>"),
    (r"easu_h_pairs_kernel", "easu_h_vpairs<64x32,persistent,tma2>"),
    (r"rcas_packed_kernel<fsr1::FmtHalf, 0, 0>", "rcas_h_packed<2px,4rows,shfl60>"),
    (r"rcas_packed_kernel<fsr1::FmtUnorm<8>, 0, 0>", "rcas_u8_packed<2px,4rows,shfl60>"),
    (r"fused_h_quad2x_kernel", "fused_easu_rcas_h_quad2x<4w,6/sm,tma2,strips>"),
    (r"easu_f32_quad2x_kernel<float", "easu_f32_quad2x<4w,4/sm,tma2,ffma2>"),
    (r"easu_f32_pairs_kernel<float", "easu_f32_vpairs<64x32,persistent,tma2,ffma2>"),
    (r"rcas_f32_packed_kernel<0>", "rcas_f32_packed<2px,4rows,shfl60>"),
]
WANT = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_bytes.sum',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'smsp__inst_executed.sum',
        'sm__inst_executed.avg.per_cycle_elapsed', 'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__cycles_elapsed.max',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma_type_fp16.avg.pct_of_peak_sustained_active', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed']


def to_bytes(value, unit):
    v = float(value.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


def to_us(value, unit):
    v = float(value.replace(",", ""))
    return v * {"ns": 1e-3, "us": 1, "usecond": 1, "ms": 1e3, "msecond": 1e3, "nsecond": 1e-3, "second": 1e6}.get(unit, 1)


def main():
    rep = sys.argv[1]
    json_dir = sys.argv[sys.argv.index("--json-dir") + 1] if "--json-dir" in sys.argv else None
    out_px = float(sys.argv[sys.argv.index("--out-px") + 1]) if "--out-px" in sys.argv else 3840.0 * 2160.0
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    want = WANT + [h for h in hdr if 'average_warps_issue_stalled' in h and 'per_issue_active.ratio' in h]
    traffic, issue = {}, {}
    for r in rows[2:]:
        if 'fsr1::' not in r[idx['Kernel Name']]:
            continue  # torch's own fill / random kernels of the driver script
        print('=====')
        for w in want:
            if w in idx:
                v = r[idx[w]]
                label = w
                if 'stalled' in w:
                    try:
                        if float(v) < 0.15:
                            continue
                    except ValueError:
                        pass
                    label = w.replace('smsp__average_warps_issue_stalled_', 'stall:').replace('_per_issue_active.ratio', '')
                print(' ', label, '=', v, units[idx[w]] if label == w and units[idx[w]] else '')
        name = r[idx['Kernel Name']]
        disp = next((d for pat, d in DISPLAY if pat in name), None)
        if disp is None:
            continue

        def num(key):
            return float(r[idx[key]].replace(",", ""))
        dur_us = to_us(r[idx['gpu__time_duration.sum']], units[idx['gpu__time_duration.sum']])
        traffic[disp] = to_bytes(r[idx['dram__bytes_read.sum']], units[idx['dram__bytes_read.sum']]) + \
            to_bytes(r[idx['dram__bytes_write.sum']], units[idx['dram__bytes_write.sum']])
        inst = num('smsp__inst_executed.sum')
        issue_pct = num('smsp__issue_active.avg.pct_of_peak_sustained_active')
        issue[disp] = {
            "warp_inst": inst, "inst_per_px": inst * 32.0 / out_px,
            "inst_per_cycle_sm": num('sm__inst_executed.avg.per_cycle_elapsed'), "issue_active_pct": issue_pct,
            "fma_pipe_cycles_active_pct": num('sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active'),
            "alu_pipe_pct": num('sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active'),
            "xu_pipe_pct": num('sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active'),
            "duration_us_under_ncu": dur_us, "issue_ceiling_us": dur_us * issue_pct / 100.0,
            "fma_pipe_ceiling_us": dur_us * num('sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active') / 100.0,
            "registers": num('launch__registers_per_thread'),
            "what": "ncu --set full of one launch at 1080p->4K; inst_per_px = thread-instructions per output pixel (warp instructions x 32 / "
                    "pixels); *_ceiling_us = the launch's time if that resource were 100 % busy",
        }
    if json_dir:
        for fn, d in (("ncu_traffic.json", traffic), ("ncu_issue.json", issue)):
            path = os.path.join(json_dir, fn)
            old = {}
            if os.path.exists(path):
                try:
                    old = json.load(open(path))
                except Exception:
                    old = {}
            old.update(d)
            json.dump(old, open(path, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
