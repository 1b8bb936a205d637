"""Host-side pieces of bench.py that need no GPU: the clock sampler's parsing / fall-back and the CPU reference arm's
JSON line (the contract the driver reads)."""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench():
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_under_test", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_clock_sampler_keeps_only_samples_inside_the_timed_region(tmp_path):
    b = _bench()
    s = b.ClockSampler.__new__(b.ClockSampler)
    s.proc, s.fallback = None, None
    now = time.time()
    s.t0, s.t1 = now + 1.0, now + 2.0

    def stamp(t):
        lt = time.localtime(t)
        return time.strftime("%Y/%m/%d %H:%M:%S", lt) + ".%03d" % int((t % 1) * 1000)
    log = tmp_path / "smi.csv"
    log.write_text("\n".join([
        f"{stamp(now + 0.2)}, 1965, 1965, Not Active, Not Active, Not Active, Not Active",    # before the region
        f"{stamp(now + 1.2)}, 1590, 1965, Not Active, Not Active, Not Active, Active",
        f"{stamp(now + 1.6)}, 1530, 1965, Not Active, Not Active, Not Active, Active",
        "garbage line",
        f"{stamp(now + 2.8)}, 1965, 1965, Not Active, Active, Not Active, Not Active",        # after the region
    ]) + "\n")

    class _Log:
        name = str(log)

        def flush(self):
            pass
    s.log = _Log()
    mhz, reasons, max_mhz = s._parse()
    assert mhz == [1590.0, 1530.0] and reasons == {"sw_power_cap"} and max_mhz == 1965.0


def test_clock_sampler_falls_back_without_nvidia_smi(monkeypatch):
    b = _bench()
    monkeypatch.setenv("MMR_BENCH_SAMPLER", "nvml")
    s = b.ClockSampler(0)
    s.start()
    out = s.stop()           # no GPU / NVML here: an empty record, never an exception
    assert s.fallback is not None and out["samples"] >= 0 and "reasons" in out


def test_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` = the fp32 CPU port of the same workload on the host cores (a bounded sample)."""
    env = dict(os.environ, OMP_NUM_THREADS=str(min(8, os.cpu_count() or 1)))
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                        "--warmup", "0"], capture_output=True, text=True, timeout=900, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "pairs_scored_per_sec" and line["unit"] == "pairs/s"
    assert line["value"] > 0 and line["higher_is_better"] is True and line["gpu_launches"] == 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"] == {"value": line["value"], "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in line["config"] and "model" not in line["config"]
    assert line["config"]["pairs_per_step_per_gpu"] == 256 and line["steps"] == 1 and line["warmup"] == 0
