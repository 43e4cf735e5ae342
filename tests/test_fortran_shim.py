"""(f)-2 readiness: the Fortran drop-in driver against a real compiler, for the day one exists.

This image has no Fortran compiler, so both tests skip here.  With `gfortran` on the PATH and the
reference sources present, `make -C hypotremormcmc_b200/fortran fortran` compiles htm_b200_binding.f90 and
hypo_tremor_mcmc_b200.f90 against the reference's UNMODIFIED cls_line_text / cls_param / cls_obs_data
(src/Makefile:57-60 lists them for hypo_tremor_mcmc) and links libhtm_b200.so; on a GPU box the driver then runs
BASELINE configs[0]'s dataset and must write exactly the files the C++ twin driver writes for the same seed.
"""
import filecmp
import os
import shutil
import subprocess

import pytest

import hypotremormcmc_b200 as H
from hypotremormcmc_b200 import io as hio

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FDIR = os.path.join(ROOT, "hypotremormcmc_b200", "fortran")
HAVE = bool(shutil.which(os.environ.get("FC", "gfortran"))) and os.path.isdir("/root/reference/src")
needs_fortran = pytest.mark.skipif(not HAVE, reason="no Fortran compiler / reference sources: the shim cannot be built here")


def test_make_target_reports_a_missing_compiler_instead_of_failing():
    r = subprocess.run(["make", "-s", "-C", FDIR, "fortran"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    if not HAVE:
        assert "not built" in r.stdout


@needs_fortran
def test_fortran_shim_compiles_against_the_reference_boundary_sources():
    r = subprocess.run(["make", "-C", FDIR, "fortran"], capture_output=True, text=True)
    assert r.returncode == 0 and os.path.exists(os.path.join(FDIR, "_build", "hypo_tremor_mcmc")), r.stdout + r.stderr


@needs_fortran
@pytest.mark.gpu
def test_fortran_driver_writes_what_the_cpp_twin_writes(tmp_path):
    subprocess.run(["make", "-C", FDIR, "fortran"], check=True, capture_output=True)
    syn = H.Synthetic(1, 10, 20231001)
    cfg = H.default_config(n_sta=10, n_events=1, n_procs=4, n_chains=5, n_cool=1, n_iter=20000, n_burn=5000, n_interval=50)
    runs = {}
    for name, exe in (("f90", os.path.join(FDIR, "_build", "hypo_tremor_mcmc")),
                      ("cpp", os.path.join(ROOT, "drivers", "hypo_tremor_mcmc_b200"))):
        d = tmp_path / name
        hio.write_dataset(str(d), syn, cfg)
        r = subprocess.run([exe, "hypo_tremor.in"], cwd=d, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
        runs[name] = d
    for rank in range(cfg.n_procs):
        for pat in ("vs.%02d.out", "hypo.%02d.out", "t_corr.%02d.out", "qs.%02d.out", "a_corr.%02d.out", "likelihood%02d.out"):
            assert filecmp.cmp(runs["f90"] / (pat % rank), runs["cpp"] / (pat % rank), shallow=False), pat % rank
    assert open(runs["f90"] / "proposal_count.txt").read() == open(runs["cpp"] / "proposal_count.txt").read()
