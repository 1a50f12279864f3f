"""bench.py's host-side contract pieces that need no GPU: one JSON line on stdout whatever libraries print, the
roofline's per-ray byte floor, the clock sampler's parsing."""

import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_stdout_carries_exactly_one_json_line():
    code = ("import os, sys; sys.path.insert(0, %r); import bench; bench.claim_stdout(); "
            "os.write(1, b'NCCL version 2.28.9+cuda12.9\\n'); print('library chatter'); "
            "bench.emit({'metric': 'x', 'value': 1.5})" % ROOT)
    p = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120)
    assert p.returncode == 0, p.stderr
    lines = p.stdout.splitlines()
    assert len(lines) == 1 and json.loads(lines[0]) == {"metric": "x", "value": 1.5}
    assert "NCCL version" in p.stderr and "library chatter" in p.stderr


def test_byte_floor_per_ray_follows_the_survey():
    import bench
    # B_floor = 32 (k + 3), k = quadtree depth: 416 / 352 / 288 bytes
    assert [32 * (bench.quadtree_depth(w) + 3) for w in (92160, 23040, 5760)] == [416, 352, 288]


def test_clock_sampler_uses_only_rows_inside_the_timed_region():
    import bench
    s = bench.ClockSampler(enabled=False)
    s.proc = type("P", (), {"terminate": lambda self: None, "wait": lambda self, timeout=None: 0})()
    s.rows = [(10.0, ["1200", "1965", "Not Active", "Not Active", "Not Active", "Not Active"]),      # before
              (20.0, ["1965", "1965", "Not Active", "Not Active", "Not Active", "Active"]),
              (20.5, ["1950", "1965", "Not Active", "Not Active", "Not Active", "Not Active"]),
              (30.0, ["900", "1965", "Active", "Not Active", "Not Active", "Not Active"])]          # after
    s.t0, s.t1 = 19.5, 21.0
    out = s.stop()
    assert out["samples"] == 2 and out["sm_mhz"] == 1957.5 and out["sm_max_mhz"] == 1965.0
    assert out["reasons"] == ["sw_power_cap"]
