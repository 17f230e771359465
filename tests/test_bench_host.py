"""Host-side pieces of bench.py that run without a GPU: the clock sampler's life cycle (armed early, released at the first timed
step, NVML or the nvidia-smi fallback — neither exists in the build container, so this is the 'unavailable' path, which must not
raise), the NUMA binding's best-effort contract, the workload table, and the reference arm's JSON line on a tiny budget."""
import importlib.util
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench():
    spec = importlib.util.spec_from_file_location("fsr1_bench", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_clock_sampler_life_cycle_without_a_gpu():
    b = _bench()
    s = b.ClockSampler(0)
    s.start()
    s.trigger()
    out = s.stop()
    assert isinstance(out, dict) and "sm_mhz" in out and "reasons" in out
    json.dumps(out)                                   # goes into the JSON line as is
    s2 = b.ClockSampler(0)                            # stop() without start() / trigger(): the FSR1_BENCH_NO_SAMPLER path
    assert "sm_mhz" in s2._smi_once()


def test_numa_binding_is_best_effort():
    b = _bench()
    before = os.sched_getaffinity(0)
    assert b.bind_to_gpu_numa_node(0) is None         # no GPU here: nothing read, nothing changed
    assert os.sched_getaffinity(0) == before


def test_workloads_name_the_baseline_configurations():
    b = _bench()
    base = json.load(open(os.path.join(ROOT, "BASELINE.json")))
    assert "Mpixels" in b.METRIC and "pix" in base["metric"].lower()
    for name, (iw, ih, ow, oh, dt, text) in b.WORKLOADS.items():
        assert dt in ("f16", "f32", "u8") and ow >= iw and oh >= ih and "%dx%d" % (iw, ih) in text
    assert b.WORKLOADS["1080p-4k-fp16"][:4] == (1920, 1080, 3840, 2160)


def test_reference_arm_prints_one_json_line():
    env = dict(os.environ, OMP_NUM_THREADS="4")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "Mpix/s" and line["value"] > 0 and line["higher_is_better"] is True
    assert line["cpu_baseline"]["kind"] in ("reference", "port") and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert line["n_gpus"] == 1 and line["metric"] == _bench().METRIC
