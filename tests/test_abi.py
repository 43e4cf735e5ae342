"""The C-ABI library loads without a GPU, exports every symbol include/htm_b200.h declares,
its struct layout matches the Python/Fortran mirrors, and compute entry points fail LOUDLY
(no CPU fallback) when no CUDA device is present."""
import ctypes
import os
import re

import numpy as np
import pytest

import hypotremormcmc_b200 as H
from conftest import have_gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "htm_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\bint32_t\s+(htm_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = H.load_library()
    names = header_functions()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), "libhtm_b200.so does not export %s" % n
    assert sorted(H.api.EXPORTS) == names


def test_header_is_plain_c():
    """The boundary is a C ABI: the header must compile as C99 (what cgo / ISO_C_BINDING tooling reads) and as C++."""
    import shutil
    import subprocess
    hdr = os.path.join(ROOT, "include", "htm_b200.h")
    if not shutil.which("gcc"):
        pytest.skip("no gcc")
    for lang, std in (("c", "c99"), ("c++", "c++11")):
        r = subprocess.run(["gcc", "-x", lang, "-std=" + std, "-Wall", "-Wextra", "-pedantic", "-fsyntax-only", hdr],
                           capture_output=True, text=True)
        assert r.returncode == 0 and not r.stderr.strip(), r.stderr


def test_config_struct_matches_c_defaults():
    lib = H.load_library()
    c = H.HtmConfig()
    assert lib.htm_config_default(ctypes.byref(c)) == 0
    d = H.default_config()
    # every field of the sample-file defaults agrees between the C side and the Python mirror;
    # a layout mismatch would scramble them
    for name, _ in H.HtmConfig._fields_:
        if name in ("n_sta", "n_events", "device", "shard_rank", "max_samples", "hist_bins", "lane_slots",
                    "gibbs_shard_events", "summary"):
            continue
        assert getattr(c, name) == getattr(d, name), name
    assert ctypes.sizeof(H.HtmConfig) == 21 * 8 + 28 * 4
    assert c.abi_version == 2 and c.temp_high == 200.0 and c.step_size_a_corr == 0.005


def test_fortran_binding_mirrors_the_struct():
    f90 = open(os.path.join(ROOT, "hypotremormcmc_b200", "fortran", "htm_b200_binding.f90")).read().lower()
    body = f90[f90.index("type, bind(c) :: htm_config"):f90.index("end type htm_config")]
    names = re.findall(r"::\s*([a-z0-9_, ]+)", body)
    flat = [n.strip() for group in names[1:] for n in group.split(",")]
    assert flat == [n for n, _ in H.HtmConfig._fields_]
    for fn in header_functions():
        assert 'name="%s"' % fn in f90.replace(" ", ""), "no bind(C) interface for %s" % fn


def test_argument_validation_without_device():
    lib = H.load_library()
    h = ctypes.c_void_p()
    bad = H.default_config(n_sta=0, n_events=1)
    assert lib.htm_create(ctypes.byref(h), ctypes.byref(bad)) == H.config.HTM_ERR_ARG
    bad = H.default_config(n_sta=4, n_events=2, mode=H.MODE_FACTORISED)  # solve_* = T
    rc = lib.htm_create(ctypes.byref(h), ctypes.byref(bad))
    buf = ctypes.create_string_buffer(300)
    lib.htm_last_error(None, buf, 300)
    assert rc == H.config.HTM_ERR_ARG and b"factorised" in buf.value


@pytest.mark.skipif(have_gpu(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback():
    cfg = H.default_config(n_sta=4, n_events=2, mode=H.MODE_FACTORISED, solve_vs=0, solve_t_corr=0, solve_qs=0,
                           solve_a_corr=0)
    with pytest.raises(H.HtmError) as ei:
        H.HypoTremorB200(cfg)
    assert ei.value.code == H.config.HTM_ERR_CUDA and "no CPU fallback" in str(ei.value)
    with pytest.raises(H.HtmError):
        H.api.measure_fp32_peak(0)


def test_upstream_entry_points_validate_before_they_need_a_device():
    """htm_detect_windows / htm_measure_windows / htm_select_events: argument errors are reported as such on any box; with
    valid arguments and no GPU the answer is HTM_ERR_CUDA, never a host computation"""
    env = np.abs(np.random.default_rng(0).normal(0, 1, (4, 400)))
    bad = [lambda: H.api.measure_windows(env, 1.0, 100, 50, [9]),            # window 9 = samples 400 .. 499: outside
           lambda: H.api.measure_windows(env[:2], 1.0, 100, 50, [1]),        # fewer than three stations
           lambda: H.api.measure_windows(env, 0.0, 100, 50, [1]),            # dt
           lambda: H.api.detect_windows(env, 101, 50, 0.98, 2),              # odd window (src/cls_correlator.f90:180-183)
           lambda: H.api.detect_windows(env, 100, 50, 1.0, 2),               # alpha < 1
           lambda: H.api.detect_windows(env, 100, 50, 1e-9, 2),              # int(n n_win alpha) = 0: no such element
           lambda: H.api.detect_windows(env, 100, 50, 0.98, 2, n_win=8)]     # windows beyond the data
    for f in bad:
        with pytest.raises(H.HtmError) as ei:
            f()
        assert ei.value.code == H.config.HTM_ERR_ARG, str(ei.value)
    import torch
    if not torch.cuda.is_available():
        for f in (lambda: H.api.measure_windows(env, 1.0, 100, 50, [1, 2]), lambda: H.api.detect_windows(env, 100, 50, 0.98, 2)):
            with pytest.raises(H.HtmError) as ei:
                f()
            assert ei.value.code == H.config.HTM_ERR_CUDA and "no CPU fallback" in str(ei.value)


def test_product_does_not_touch_the_oracle():
    # the oracle is test infrastructure: nothing under the product package or the driver sources
    # may import, include or link it
    for base in ("hypotremormcmc_b200", "include", "drivers"):
        for dp, _, files in os.walk(os.path.join(ROOT, base)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".hpp", ".h", ".cpp", ".f90", "Makefile")):
                    txt = open(os.path.join(dp, f), errors="ignore").read()
                    assert "htm_oracle" not in txt and "pyoracle" not in txt and "libhtm_oracle" not in txt, \
                        "%s references the oracle" % os.path.join(dp, f)


def _build_c_client(tmp_path):
    import shutil
    import subprocess
    if not shutil.which("gcc"):
        pytest.skip("no gcc")
    exe = str(tmp_path / "minimal_c_client")
    libdir = os.path.join(ROOT, "hypotremormcmc_b200", "csrc")
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-O2", "-I" + os.path.join(ROOT, "include"),
                        os.path.join(ROOT, "examples", "minimal_c_client.c"), "-L" + libdir, "-lhtm_b200",
                        "-Wl,-rpath," + libdir, "-lm", "-o", exe], capture_output=True, text=True)
    assert r.returncode == 0 and not r.stderr.strip(), r.stderr
    return exe


def test_plain_c_client_links_and_fails_loudly_without_a_device(tmp_path):
    """examples/minimal_c_client.c: the ABI is usable from plain C; with no GPU the first compute call stops it."""
    import subprocess
    if have_gpu():
        pytest.skip("GPU present: covered by the gpu-marked twin")
    r = subprocess.run([_build_c_client(tmp_path)], capture_output=True, text=True)
    assert r.returncode == 1 and "no CPU fallback" in r.stderr


@pytest.mark.gpu
def test_plain_c_client_runs_on_the_gpu(tmp_path):
    import subprocess
    r = subprocess.run([_build_c_client(tmp_path)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert r.stdout.startswith("15 records of virtual rank 0; last: iteration 1901"), r.stdout
    assert "cold x/y/z proposals 256000" in r.stdout   # 2000 iterations x 64 events x 2 cold chains
