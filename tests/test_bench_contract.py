"""bench.py's reference arm runs without a GPU: its JSON line must carry the keys the driver reads."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _line(*args):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", *args],
                         capture_output=True, text=True, timeout=600, check=True).stdout
    lines = [l for l in out.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    return json.loads(lines[0])


def test_reference_arm_line():
    d = _line("--steps", "1", "--warmup", "0")
    assert d["impl"] == "reference" and d["metric"] == "grid-point updates/s" and d["unit"] == d["metric"]
    assert d["higher_is_better"] is True and d["value"] > 0 and d["n_gpus"] == 1 and d["steps"] == 1
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0,
                        "d2h_bytes_per_step": 0}
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("port", "reference") and cb["cores"] >= 1 and cb["value"] == d["value"]
    assert "1024x1024x80" in d["config"]["workload"] and "sample" in cb


def test_reference_arm_other_ranks_stay_silent():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                        "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=120, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def _fake_smi(tmp_path, delay):
    """stand-in for `nvidia-smi --query-gpu=... -lms 10`: a line every 10 ms after `delay` seconds of start-up"""
    p = tmp_path / "fake-smi"
    p.write_text("#!%s\nimport sys, time, datetime\ntime.sleep(%r)\nwhile True:\n"
                 "    t = datetime.datetime.now().strftime('%%Y/%%m/%%d %%H:%%M:%%S.%%f')[:-3]\n"
                 "    print(t + ', 1965, 1965, 600.0, Not Active, Not Active, Not Active, Active', flush=True)\n"
                 "    time.sleep(0.01)\n" % (sys.executable, delay))
    p.chmod(0o755)
    return str(p)


def test_clock_sampler_short_region_keeps_the_load_until_a_sample_exists(tmp_path):
    import time
    sys.path.insert(0, ROOT)
    import bench
    s = bench.ClockSampler(0, command=_fake_smi(tmp_path, 0.6))
    s.start()
    s.mark_begin()
    time.sleep(0.005)                     # a timed region that ends long before the first line is printed
    s.mark_end()
    batches = 0
    while s.wants_more_load():
        time.sleep(0.02)                  # "another batch of the same steps"
        batches += 1
    c = s.stop()
    assert 5 <= batches <= 60 and c["samples"] >= 1 and c["sm_mhz"] == 1965.0 and c["reasons"] == ["sw_power_cap"]
    assert "kept running" in c["window"]


def test_clock_sampler_prefers_samples_of_the_region_and_never_extends_then(tmp_path):
    import time
    sys.path.insert(0, ROOT)
    import bench
    s = bench.ClockSampler(0, command=_fake_smi(tmp_path, 0.0))
    s.start()
    time.sleep(0.3)
    s.mark_begin()
    time.sleep(0.1)
    s.mark_end()
    assert not s.wants_more_load()
    c = s.stop()
    assert c["window"] == "timed region" and 3 <= c["samples"] <= 12
    # a region between two samples falls back to its neighbours, still without extra load
    s = bench.ClockSampler(0, command=_fake_smi(tmp_path, 0.0))
    s.start()
    time.sleep(0.3)
    s.mark_begin()
    s.mark_end()
    time.sleep(0.03)
    assert not s.wants_more_load()
    assert "same load" in s.stop()["window"]


def test_clock_sampler_gives_up_when_nvidia_smi_stays_silent(tmp_path):
    import time
    sys.path.insert(0, ROOT)
    import bench
    s = bench.ClockSampler(0, command=_fake_smi(tmp_path, 30.0))
    s.start()
    s.mark_begin()
    s.mark_end()
    t = time.time()
    while s.wants_more_load(limit_s=0.3):
        time.sleep(0.02)
    assert time.time() - t < 1.0
    c = s.stop()
    assert c["samples"] == 0 and c["sm_mhz"] is None and c["window"] == "no sample"
    assert bench.ClockSampler(0, command="/nonexistent/nvidia-smi").wants_more_load() is False
