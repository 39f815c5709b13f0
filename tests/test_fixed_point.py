"""The integer tricks of the FMA-pipe sampler (drr_march_warp.cu: march_core, drr_device.cuh: hw_trilinear_cell2q) restated in C
and checked on the host's IEEE arithmetic: the texture unit's 1.8 fixed-point coordinate from one round-down FMA (3 M cases
incl. exact ties and fractions that round up into the next cell), the staging index from byte permutes + dp4a, the
multiply-shift cell decomposition, all 2 x 256^3 filter weights against the unit's integer model, and the filter value."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_fixed_point_tricks_are_exact(tmp_path):
    src = os.path.join(ROOT, "tests", "fixed_point_check.c")
    exe = str(tmp_path / "fixed_point_check")
    subprocess.run(["gcc", "-O2", "-ffp-contract=off", "-frounding-math", "-o", exe, src, "-lm"], check=True)
    r = subprocess.run([exe], capture_output=True, text=True)
    sys.stdout.write(r.stdout)
    assert r.returncode == 0 and "all ok" in r.stdout


def test_device_code_is_the_same_arithmetic():
    """Guards against the C restatement and the device code drifting apart."""
    dev = open(os.path.join(ROOT, "deepdrr_b200", "csrc", "drr_device.cuh")).read()
    body = dev[dev.index("float hw_trilinear_cell2q("):]
    body = body[:body.index("\n}\n")]
    for needle in ("(qx & 0xFFu)", "__fmaf_rn(bf, 0x1p-8f, 0x1p-17f)", "__fmaf_rn(bf, -0x1p-8f, 1.0f + 0x1p-17f)", "__fsub_rn(256.0f, cf)",
                   "__ffma2_rn(X1, make_float2(-1.0f, -1.0f), wz)", "__fmul2_rn(wz, make_float2(A.x, A.y))"):
        assert needle in body, needle
    march = open(os.path.join(ROOT, "deepdrr_b200", "csrc", "drr_march_warp.cu")).read()
    for needle in ("__fmaf_rn(-256.0f, b1x, 8388608.0f)", "__fadd_rn(kfx, 0.5f)", "__fmaf_rd(x, 256.0f, kqx)", "__fmaf_rd(x, 256.0f, kfx)",
                   "__byte_perm(__byte_perm(qx, qy, 0x0051), qz, 0x0510)", "(unsigned)(65536.0f / (float)nx) + 2u", "< 21845",
                   "max(bx, max(by, bz)) <= 255 && (bx * by <= 255 || bz == 1)"):
        assert needle in march, needle
