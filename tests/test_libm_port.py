"""aprilgrid-rs_b200/csrc/ag_libm.h restates glibc's acosf / atanf / atan2f operation for operation so
that the kernels' theta / phi carry the bits of the `acosf` / `atan2f` the reference's f32::acos /
f32::atan2 resolve to here (src/detector.rs:348-349).  This pins the restatement to the machine's
libm: a strided sweep over every binade (the full sweeps -- atanf over all positive floats, acosf
over 164 M arguments, atan2f over 100 M pairs -- were run once with strides 1 / 13: no difference)."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_libm_port_matches_platform_libm(tmp_path):
    exe = str(tmp_path / "libm_port_check")
    subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-I" + os.path.join(ROOT, "aprilgrid-rs_b200", "csrc"),
                           "-o", exe, os.path.join(ROOT, "tests", "libm_port_check.c"), "-lm"])
    out = subprocess.check_output([exe, "499", "997", "4000000"], text=True).split()
    assert out == ["0", "0", "0"], "ag_libm.h differs from this platform's libm (acosf, atanf, atan2f): %s" % out
