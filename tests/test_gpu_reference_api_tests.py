"""The reference's OWN API test programs against the product library (pytest -m gpu).

oracle/Makefile (target api_tests) compiles /root/reference/tests/test_specbleach_2d_denoiser.c,
test_specbleach_denoiser.c, test_integration.c and test_audio_regression.c -- unchanged, from
where they lie -- and links them with `-lspecbleach`, which resolves to this build's
libspecbleach.so.0.  They only use the public headers (plus configurations.h), so they are the
reference maintainers' own statement of the API contract: NULL handling, profile load / reset /
mode validation, adaptive method switching, in-place processing in learn mode, RMS bounds of the
1D / adaptive / 2D pipelines, run-to-run determinism (1e-10), power reduction, Martin != SPP."""
import subprocess
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu

BIN = Path(__file__).resolve().parent.parent / "oracle" / "_ref" / "api_tests"
PROGRAMS = ["test_specbleach_2d_denoiser", "test_specbleach_denoiser", "test_integration",
            "test_audio_regression"]


@pytest.mark.parametrize("prog", PROGRAMS)
def test_reference_api_test_program_passes_against_the_cuda_library(gpu, prog):
    exe = BIN / prog
    if not exe.exists():
        pytest.skip(f"{exe} not built (needs /root/reference at build time: make -C oracle api_tests)")
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=900)
    tail = (r.stdout[-1500:] + "\n" + r.stderr[-1500:]).strip()
    print(tail)
    assert r.returncode == 0, f"{prog} exited {r.returncode}:\n{tail}"
