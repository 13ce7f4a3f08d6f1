"""Memory-safety pass without compute-sanitizer (closed on the GPU boxes of this project): the
library's own guard zones (SB_GUARD, include/specbleach_b200.h) around every device allocation,
driven by tools/sanitize.py over every kernel family."""
import os
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
pytestmark = pytest.mark.gpu


def run_py(code_or_path, guard, is_path=False):
    env = dict(os.environ)
    env.pop("SB_GUARD", None)
    if guard:
        env["SB_GUARD"] = guard
    cmd = [sys.executable, str(code_or_path)] if is_path else [sys.executable, "-c", code_or_path]
    r = subprocess.run(cmd, capture_output=True, text=True, cwd=str(ROOT), env=env, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    return r.stdout.strip().splitlines()


def test_guard_selftest_detects_a_store_behind_an_allocation(gpu):
    code = ("import numpy as np, libspecbleach_b200 as sb\n"
            "from libspecbleach_b200 import api\n"
            "d = sb.Denoiser2D(48000, 50.0)\n"
            "d.load_parameters(learn_noise=1)\n"
            "d.process(np.zeros(4800, dtype=np.float32))\n"
            "print(api.lib().specbleach_b200_debug_check_guards())\n")
    assert int(run_py(code, "selftest")[-1]) >= 1      # the detector sees a one-byte overrun
    assert int(run_py(code, "1")[-1]) == 0
    assert int(run_py(code, None)[-1]) == -1           # guarding off: the call says so


def test_every_kernel_family_stays_inside_its_buffers(gpu):
    plain = run_py(ROOT / "tools" / "sanitize.py", None, is_path=True)
    guarded = run_py(ROOT / "tools" / "sanitize.py", "1", is_path=True)
    assert plain[-1] == "guard zones damaged: -1"
    assert guarded[-1] == "guard zones damaged: 0"
    # identical checksums: no kernel loaded a float from a guard zone or from memory it had not
    # written (fresh allocations hold NaNs under SB_GUARD)
    assert plain[:-1] == guarded[:-1] and len(plain) >= 19
