"""CPU tests of the drop-in boundary: the shared library loads, exports every symbol
declared in include/*.h, the parameter struct has the reference's layout, the
headers are valid C, and without a GPU the product fails loudly (no CPU fallback)."""
import ctypes as C
import re
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
INCLUDE = ROOT / "include"


def declared_functions(header: Path):
    text = re.sub(r"/\*.*?\*/", "", header.read_text(), flags=re.S)
    return sorted(set(re.findall(r"\b(specbleach_\w+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from libspecbleach_b200 import api
    L = api.lib()
    names = []
    for h in ("specbleach_denoiser.h", "specbleach_2d_denoiser.h", "specbleach_b200.h"):
        names += declared_functions(INCLUDE / h)
    assert len(names) >= 22 + 10
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing
    assert set(api.PUBLIC_SYMBOLS) <= set(names) and set(api.EXTENSION_SYMBOLS) <= set(names)


def test_public_api_is_the_references_22_symbols():
    a = declared_functions(INCLUDE / "specbleach_denoiser.h")
    b = declared_functions(INCLUDE / "specbleach_2d_denoiser.h")
    assert len(a) == 11 and len(b) == 11
    assert [n.replace("specbleach_2d_", "specbleach_") for n in b] == a


def test_parameter_struct_layout():
    from libspecbleach_b200.api import DenoiserParameters
    assert C.sizeof(DenoiserParameters) == 44          # 11 fields, C layout
    offs = [getattr(DenoiserParameters, f).offset for f, _ in DenoiserParameters._fields_]
    assert offs == [0, 4, 8, 12, 16, 20, 24, 28, 32, 36, 40]


def test_headers_compile_as_c_and_struct_is_44_bytes(tmp_path):
    src = tmp_path / "t.c"
    src.write_text(
        '#include "specbleach_denoiser.h"\n#include "specbleach_2d_denoiser.h"\n'
        '#include "specbleach_b200.h"\n'
        "_Static_assert(sizeof(SpectralBleachDenoiserParameters) == 44, \"1D params\");\n"
        "_Static_assert(sizeof(SpectralBleach2DDenoiserParameters) == 44, \"2D params\");\n"
        "int main(void) { return specbleach_get_latency(0) + specbleach_2d_get_latency(0); }\n")
    r = subprocess.run(["gcc", "-std=c17", "-Wall", "-Werror", "-fsyntax-only", f"-I{INCLUDE}", str(src)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_null_handle_behaviour_needs_no_gpu():
    """NULL / bad-argument paths of the reference API (test_specbleach_2d_denoiser.c:34-77)."""
    from libspecbleach_b200 import api
    L = api.lib()
    for pre in ("specbleach_", "specbleach_2d_"):
        g = lambda n: getattr(L, pre + n)
        assert g("get_latency")(None) == 0
        assert g("get_noise_profile_size")(None) == 0
        assert not g("process")(None, 16, None, None)
        assert not g("load_parameters")(None, api.make_params())
        assert not g("reset_noise_profile")(None)
        assert not g("noise_profile_available_for_mode")(None, 1)
        assert g("get_noise_profile_block_count_for_mode")(None, 1) == 0
        assert not g("get_noise_profile_for_mode")(None, 1)
        g("free")(None)
    assert not L.specbleach_initialize(0, 50.0)
    assert not L.specbleach_initialize(3999, 50.0)        # specbleach_denoiser.c:41
    assert not L.specbleach_initialize(192001, 50.0)
    assert not L.specbleach_initialize(48000, 0.0)
    assert not L.specbleach_2d_initialize(0, 50.0)
    assert not L.specbleach_2d_initialize(48000, -1.0)


def test_fails_loudly_without_gpu():
    import libspecbleach_b200 as sb
    if sb.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        sb.Denoiser2D(48000, 50.0)


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under libspecbleach_b200/ may use it."""
    pkg = ROOT / "libspecbleach_b200"
    pat = re.compile(r"(from\s+oracle|import\s+oracle|oracle/|specbleach_np|ref_binding|libspecbleach_ref)")
    for f in list(pkg.rglob("*.py")) + list(pkg.rglob("*.cu")) + list(pkg.rglob("*.h")):
        assert not pat.search(f.read_text()), f


def test_extension_io_entry_points_reject_bad_arguments_without_a_gpu(tmp_path):
    """persistence / PCM front end (include/specbleach_b200.h): NULL handles and buffers."""
    from libspecbleach_b200 import api
    L = api.lib()
    assert not L.specbleach_b200_profile_save(None, str(tmp_path / "x").encode())
    assert not L.specbleach_b200_profile_load(None, str(tmp_path / "x").encode())
    assert L.specbleach_b200_state_size(None) == 0
    assert not L.specbleach_b200_state_save(None, None, 0)
    assert not L.specbleach_b200_state_load(None, None, 0)
    assert not L.specbleach_b200_process_interleaved(None, 2, 16, 0, None, None)


def test_wav_tool_is_built_and_prints_usage():
    tool = ROOT / "tools" / "sb_denoise"
    assert tool.exists(), "run __graft_entry__.build()"
    r = subprocess.run([str(tool)], capture_output=True, text=True)
    assert r.returncode == 2 and "Usage" in r.stderr


def test_configure_validates_its_arguments_without_a_gpu():
    """engine tuning (include/specbleach_b200.h): 0 = automatic sub-call size, otherwise
    [32, 32768] frames; 1..32 lanes; refusals carry a message.  Run in a child process so
    the settings do not leak into other tests."""
    code = (
        "from libspecbleach_b200 import api\n"
        "L = api.lib()\n"
        "assert not L.specbleach_b200_configure(16, 4) and 'chunk_frames' in api.last_error()\n"
        "assert not L.specbleach_b200_configure(40000, 4)\n"
        "assert not L.specbleach_b200_configure(0, 0) and 'lanes' in api.last_error()\n"
        "assert not L.specbleach_b200_configure(0, 33)\n"
        "assert L.specbleach_b200_configure(0, 2)\n"
        "assert L.specbleach_b200_configure(4096, 8)\n"
        "print('ok')\n")
    import sys
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=str(ROOT))
    assert r.returncode == 0 and r.stdout.strip() == "ok", r.stderr


def test_launch_share_summary_reads_the_committed_ncu_list():
    """tools/ncu_summary.py --launches is what turns the ncu launch list under profiles/ into
    the share table DESIGN.md quotes; the committed pair must stay consistent."""
    import sys
    lists = sorted((ROOT / "profiles").glob("r0*_launches_ncu.csv"))
    assert lists, "no launch list under profiles/"
    csv_path = lists[-1]
    r = subprocess.run([sys.executable, str(ROOT / "tools" / "ncu_summary.py"), "--launches", str(csv_path), "cmd"],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    rows = [ln.split() for ln in r.stdout.splitlines()[2:] if ln.strip()]
    assert rows[0][0] == "nlm3_kernel"                       # the dominant kernel
    shares = {row[0]: float(row[-4].rstrip("%")) for row in rows}
    assert abs(sum(shares.values()) - 100.0) < 1.0
    committed = (ROOT / "profiles" / csv_path.name.replace("launches_ncu.csv", "launch_shares.txt")).read_text()
    assert f"{shares['nlm3_kernel']:.1f}%" in committed


def test_bench_line_kernel_shares_agree_with_the_ncu_launch_list():
    """north_star: every kernel reported against its roofline.  The committed config-2 bench line
    (profiles/r02_bench_config2.json: roofline_by_class from CUDA events around every launch) and
    the committed ncu launch list of the same command are two independent measurements of where the
    step goes; the NLM share and the order of the big classes must agree (ncu's per-launch times are
    cold-cache and serialised: shares, not absolutes)."""
    import json
    import sys
    line = json.loads((ROOT / "profiles" / "r02_bench_config2.json").read_text().strip().splitlines()[-1])
    by_class = line["roofline_by_class"]
    for name in ("forward_fft", "inverse_fft", "nlm", "snr", "gain_pre", "final", "band", "nmr", "ola"):
        e = by_class[name]
        assert e["bound"] in ("hbm", "fp32", "smem/fp32", "mufu") and 0.0 < e["frac"] < 2.0, (name, e)
    ev = {k: v["ms"] for k, v in by_class.items()}
    ev_total = sum(ev.values())
    r = subprocess.run([sys.executable, str(ROOT / "tools" / "ncu_summary.py"), "--launches",
                        str(ROOT / "profiles" / "r02_launches_ncu.csv"), "cmd"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    rows = [ln.split() for ln in r.stdout.splitlines()[2:] if ln.strip()]
    ncu = {row[0]: float(row[-4].rstrip("%")) for row in rows}
    ncu_nlm = ncu["nlm3_kernel"] + ncu["nlm3_paste_kernel"]
    ev_nlm = 100.0 * ev["nlm"] / ev_total
    assert abs(ncu_nlm - ev_nlm) < 8.0, (ncu_nlm, ev_nlm)
    assert abs((ncu["forward_kernel<1>"] + ncu["inverse_kernel<1>"]) -
               100.0 * (ev["forward_fft"] + ev["inverse_fft"]) / ev_total) < 4.0
    # the NLM roofline object: algorithmic rate above the FP32 peak only through row-SSD sharing
    rf = line["roofline"]
    assert rf["bound"] == "fp32" and 0.9 < rf["frac"] < 1.5 and rf["traffic"] > 0
    assert 0.25 < rf["executed"]["frac_of_peak"] < 0.45


def test_library_identity_matches_the_reference_install():
    """src/meson.build:27-44: library('specbleach', soversion '0') + pkg-config 'libspecbleach'.
    The build gives the product the reference's SONAME, the soname / development links next to it
    and a libspecbleach.pc, so `-lspecbleach` and an existing DT_NEEDED libspecbleach.so.0 resolve
    to it."""
    import re
    import shutil
    import subprocess
    libdir = ROOT / "libspecbleach_b200"
    so = libdir / "libspecbleach_b200.so"
    for name in ("libspecbleach.so.0", "libspecbleach.so"):
        assert (libdir / name).exists(), name
        assert (libdir / name).stat().st_size == so.stat().st_size
    pc = (libdir / "libspecbleach.pc").read_text()
    assert re.search(r"^Name: libspecbleach$", pc, re.M) and re.search(r"^Version: 0\.3\.0$", pc, re.M)
    assert "-lspecbleach" in pc
    if shutil.which("readelf"):
        dyn = subprocess.run(["readelf", "-d", str(so)], capture_output=True, text=True).stdout
        assert "soname: [libspecbleach.so.0]" in dyn
        exe = ROOT / "oracle" / "_ref" / "api_tests" / "test_integration"
        if exe.exists():
            dyn = subprocess.run(["readelf", "-d", str(exe)], capture_output=True, text=True).stdout
            assert "[libspecbleach.so.0]" in dyn


def test_public_headers_carry_the_lgpl_attribution():
    for name in ("specbleach_denoiser.h", "specbleach_2d_denoiser.h"):
        text = (ROOT / "include" / name).read_text()
        assert "Luciano Dato" in text and "Lesser General Public" in text
