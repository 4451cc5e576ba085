"""The kernel bodies' index arithmetic under AddressSanitizer + UBSan (host emulation, odd shapes): an out-of-bounds
index here would be an out-of-bounds access on the GPU.  tests/emu/sanitize_main.cpp drives both emulations."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++")
def test_emulated_kernel_bodies_are_sanitizer_clean(tmp_path):
    exe = str(tmp_path / "sanitize")
    emu = os.path.join(ROOT, "tests", "emu")
    cmd = ["g++", "-O1", "-g", "-std=c++17", "-fsanitize=address,undefined", "-fno-sanitize-recover=all",
           "-fno-omit-frame-pointer", "-I", os.path.join(ROOT, "ddp_b200", "csrc"), os.path.join(emu, "sanitize_main.cpp"),
           os.path.join(emu, "neck_emu.cpp"), os.path.join(emu, "bev_emu.cpp"), "-o", exe]
    build = subprocess.run(cmd, capture_output=True, text=True)
    if build.returncode != 0 and "asan" in (build.stderr + build.stdout).lower():
        pytest.skip("libasan is not installed here")
    assert build.returncode == 0, build.stderr[-2000:]
    run = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert run.returncode == 0, (run.stdout + run.stderr)[-3000:]
    assert run.stdout.count("rc 0") == 4
