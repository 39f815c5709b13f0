"""The exact jump-ahead of the fp32 ``alpha += step`` accumulation (drr_device.cuh: alpha_jump) restated in C and compared
with the plain loop on 460 000 random (alpha, step, n), ties and powers of two included."""
import os
import subprocess
import sys


def test_alpha_jump_equals_sequential_additions(tmp_path):
    src = os.path.join(os.path.dirname(os.path.abspath(__file__)), "alpha_jump_check.c")
    exe = str(tmp_path / "alpha_jump_check")
    subprocess.run(["gcc", "-O2", "-ffp-contract=off", "-o", exe, src, "-lm"], check=True)
    r = subprocess.run([exe], capture_output=True, text=True)
    sys.stdout.write(r.stdout)
    assert r.returncode == 0 and "bad 0" in r.stdout


def test_device_version_is_the_same_algorithm():
    """Guards against the two copies drifting apart: the device function must contain the same steps."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    dev = open(os.path.join(root, "deepdrr_b200", "csrc", "drr_device.cuh")).read()
    body = dev[dev.index("float alpha_jump("):]
    body = body[:body.index("\n}\n")]
    for needle in ("(ua >> 23) != (un >> 23)", "un - ua", "0x7FFFFFu - (ua & 0x7FFFFFu)", "ua + m * D", "== ulp"):
        assert needle in body, needle
