"""The C++ twin of the drop-in driver: file formats in (parameter file, station file,
selected_win.dat, opt_data.*.dat) and out (six .out families, proposal_count.txt)."""
import json
import os
import subprocess

import numpy as np
import pytest

import hypotremormcmc_b200 as H
from hypotremormcmc_b200 import io as hio

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DRIVER = os.path.join(ROOT, "drivers", "hypo_tremor_mcmc_b200")
NOSOLVE = dict(solve_vs=0, solve_t_corr=0, solve_qs=0, solve_a_corr=0)


def run_driver(cwd, *args):
    return subprocess.run([DRIVER, "hypo_tremor.in", *args], cwd=cwd, capture_output=True, text=True)


def test_dry_run_parses_the_reference_formats(tmp_path):
    syn = H.Synthetic(3, 5, 4)
    cfg = H.default_config(n_sta=5, n_events=3, n_procs=2, n_chains=3, n_iter=100, n_burn=10, n_interval=5)
    hio.write_dataset(str(tmp_path), syn, cfg)
    # rewrite the parameter file the way a user would: comments, blanks everywhere, D exponents, .true.
    lines = open(tmp_path / "hypo_tremor.in").read().splitlines()
    messy = ["#  comment line", "   "]
    for ln in lines:
        if ln.startswith("temp_high"):
            ln = "temp_high   =  2 0 0 . d 0   # blanks inside a value are removed too"
        if ln.startswith("solve_vs"):
            ln = "solve_vs=.true."
        if ln.startswith("prior_qs"):
            ln = "prior_qs = 250"
        messy.append("  " + ln.replace("=", "  =  ") + "   # trailing comment")
    messy.append("prior_t_corr = 1.5D-1")
    messy.append("alpha = 0.5      # keys of the other programs are accepted")
    open(tmp_path / "hypo_tremor.in", "w").write("\n".join(messy) + "\n")
    r = run_driver(tmp_path, "--dry-run")
    assert r.returncode == 0, r.stderr
    d = json.loads(r.stdout)
    assert d["n_sta"] == 5 and d["n_events"] == 3 and d["n_procs"] == 2 and d["n_chains"] == 3
    assert d["temp_high"] == 200.0 and d["prior_qs"] == 250.0 and d["prior_t_corr"] == 0.15
    assert d["solve_vs"] == 1 and d["use_amp"] == 1 and d["mode"] == H.MODE_BLOCKED_GIBBS
    assert d["x_mu0"] == pytest.approx(syn.x_mu[0], abs=1e-8) and d["y_mu0"] == pytest.approx(syn.y_mu[0], abs=1e-8)
    assert d["t_obs00"] == syn.t_obs[0, 0] and d["a_stdv_last"] == syn.a_stdv[-1, -1]
    assert d["sta_z_last"] == pytest.approx(syn.sta_z[-1], abs=1e-8)


def test_observation_files_are_read_one_record_per_station_line(tmp_path):
    """`read(io,*)` of 7 items per station (src/cls_obs_data.f90:92-99): extra columns and trailing text on a line are
    ignored, a record may continue on the next line, `r*c` repeats count -- and a short file is an error, not a shift."""
    syn = H.Synthetic(2, 4, 9)
    cfg = H.default_config(n_sta=4, n_events=2, n_procs=1, n_chains=2, n_iter=10, n_burn=1, n_interval=5)
    hio.write_dataset(str(tmp_path), syn, cfg)
    ref = json.loads(run_driver(tmp_path, "--dry-run").stdout)
    rows = open(tmp_path / "opt_data.000001.dat").read().splitlines()
    v = rows[0].split()
    rows[0] = " ".join(v) + "   99.0 extra columns are ignored"          # an 8th column must not shift the next station
    v = rows[1].split()
    rows[1] = " ".join(v[:4]) + "\n   " + ", ".join(v[4:])               # a record continued on the next line, commas
    open(tmp_path / "opt_data.000001.dat", "w").write("\n".join(rows) + "\n")
    got = json.loads(run_driver(tmp_path, "--dry-run").stdout)
    assert got == ref
    rows2 = open(tmp_path / "opt_data.000002.dat").read().splitlines()
    v = rows2[-1].split()
    rows2[-1] = " ".join(v[:5]) + " 2*" + v[5]                            # a_obs and a_stdv := 2 copies of one value
    open(tmp_path / "opt_data.000002.dat", "w").write("\n".join(rows2) + "\n")
    got = json.loads(run_driver(tmp_path, "--dry-run").stdout)
    assert got["a_stdv_last"] == float(v[5])
    open(tmp_path / "opt_data.000002.dat", "w").write("\n".join(rows2[:-1]) + "\n")   # one station short
    r = run_driver(tmp_path, "--dry-run")
    assert r.returncode != 0 and "short obs file" in r.stderr


def test_missing_and_unknown_keys_are_fatal(tmp_path):
    syn = H.Synthetic(2, 4, 1)
    cfg = H.default_config(n_sta=4, n_events=2)
    hio.write_dataset(str(tmp_path), syn, cfg)
    text = open(tmp_path / "hypo_tremor.in").read()
    open(tmp_path / "hypo_tremor.in", "w").write(text.replace("n_cool = 1\n", ""))
    r = run_driver(tmp_path, "--dry-run")
    assert r.returncode != 0 and "n_cool is not given" in r.stderr
    open(tmp_path / "hypo_tremor.in", "w").write(text + "n_colo = 1\n")
    r = run_driver(tmp_path, "--dry-run")
    assert r.returncode != 0 and "Invalid parameter name" in r.stderr
    r = subprocess.run([DRIVER], cwd=tmp_path, capture_output=True, text=True)
    assert r.returncode != 0 and "USAGE: hypo_tremor_mcmc [parameter file]" in r.stderr


def test_stream_reader_and_quantiles(tmp_path):
    rec = np.dtype([("i", ">i4"), ("v", ">f8", (3,))])
    a = np.zeros(5, dtype=rec)
    a["i"] = [1, 11, 21, 31, 41]
    a["v"] = np.arange(15).reshape(5, 3)
    a.tofile(tmp_path / "x.out")
    it, v = hio.read_stream(str(tmp_path / "x.out"), 3)
    assert list(it) == [1, 11, 21, 31, 41] and np.array_equal(v, np.arange(15.0).reshape(5, 3))
    with pytest.raises(ValueError):
        hio.read_stream(str(tmp_path / "x.out"), 4)
    s = np.arange(1, 1001, dtype=float)[::-1]
    med, lo, hi = hio.quantile_summary(s)
    assert (med, lo, hi) == (500.0, 25.0, 975.0)      # elements int(0.5 n), int(0.025 n), int(0.975 n), 1-based


def test_stat_files_layout(tmp_path):
    # (I9,9F13.6) / (A12,6F13.6) / (6F13.6) records of src/cls_statistics.f90:245-252,379-385,420-425
    rng = np.random.default_rng(0)
    out = dict(hypo=rng.normal(0, 1, (400, 6)), t_corr=rng.normal(0, 1, (400, 3)), a_corr=rng.normal(0, 1, (400, 3)),
               vs=rng.normal(3, 0.1, (400, 1)), qs=rng.normal(250, 10, (400, 1)))
    hio.write_stat_files(str(tmp_path), out, [7, 12], ["AAA", "BB", "C"])
    lines = open(tmp_path / "hypo.stat").read().splitlines()
    assert lines[0].startswith("# window ID") and len(lines) == 3 and len(lines[1]) == 9 + 9 * 13
    v = [float(lines[2][9 + 13 * k: 22 + 13 * k]) for k in range(9)]
    s = np.sort(out["hypo"][:, 3])
    assert abs(v[0] - s[199]) < 1e-6 and abs(v[1] - s[9]) < 1e-6 and abs(v[2] - s[389]) < 1e-6   # 1-based 200, 10, 390
    lines = open(tmp_path / "station_corrections.stat").read().splitlines()
    assert len(lines) == 4 and len(lines[1]) == 12 + 6 * 13 and lines[3][:12].strip() == "C"
    lines = open(tmp_path / "uniform_structure.stat").read().splitlines()
    assert len(lines) == 2 and len(lines[1]) == 6 * 13 and abs(float(lines[1][:13]) - 3.0) < 0.05


@pytest.mark.gpu
@pytest.mark.parametrize("solve", [0, 1])
def test_driver_end_to_end(tmp_path, solve):
    E, S, R, K = 6, 10, 2, 4
    syn = H.Synthetic(E, S, 20231002)
    kw = NOSOLVE if not solve else {}
    cfg = H.default_config(n_sta=S, n_events=E, n_procs=R, n_chains=K, n_cool=1, n_iter=2000, n_burn=500, n_interval=10,
                           **kw)
    hio.write_dataset(str(tmp_path), syn, cfg)
    r = run_driver(tmp_path, "--chunk", "37", "--summary")
    assert r.returncode == 0, r.stderr
    # the statistics stage's three tables straight from the device-side store
    dev_stat = {f: open(tmp_path / f).read() for f in ("hypo.stat", "station_corrections.stat", "uniform_structure.stat")}
    out = hio.read_outputs(str(tmp_path), R, E, S)
    n_mod = (cfg.n_iter - cfg.n_burn) * R * cfg.n_cool // cfg.n_interval     # src/cls_statistics.f90:65
    for k in ("hypo", "t_corr", "vs", "a_corr", "qs"):
        assert out[k].shape[0] == n_mod, k
        assert np.all(out["iter"][k] % 10 == 1) and out["iter"][k].min() > 500
    assert out["lik"].shape[0] == cfg.n_iter * R * cfg.n_cool // cfg.n_interval and out["iter"]["lik"].min() == 1
    assert np.all(np.isfinite(out["hypo"])) and np.all(out["hypo"][:, 2::3] > cfg.prior_z)
    if not solve:
        assert np.all(out["vs"] == cfg.prior_vs) and np.all(out["t_corr"] == 0.0)
    else:
        assert out["vs"].std() > 0 and out["t_corr"].std() > 0
    rows = hio.read_proposal_count(str(tmp_path / "proposal_count.txt"))
    assert [r_[0] for r_ in rows] == H.PROPOSAL_LABELS
    assert all(a <= p for _, p, a in rows) and sum(p for _, p, _ in rows[4:]) == cfg.n_iter * E * R
    # the same run through the Python mirror of the ABI gives byte-identical samples
    c2 = H.copy_config(cfg, mode=H.MODE_BLOCKED_GIBBS if solve else H.MODE_FACTORISED, precision=32, max_samples=256)
    with H.HypoTremorB200(c2) as g:
        g.load(syn)
        g.init_chains()
        g.run(1, cfg.n_iter)
        s = [g.fetch_samples(rank) for rank in range(R)]
    assert np.array_equal(np.concatenate([x["hypo"] for x in s]), out["hypo"])
    assert np.array_equal(np.concatenate([x["vs"] for x in s]), out["vs"][:, 0])
    # the statistics stage's tables from these files: medians bracketed by the 2.5 / 97.5 % bounds
    hio.write_stat_files(str(tmp_path), out, list(range(1, E + 1)), ["ST%02d" % j for j in range(S)])
    # ... equal, character for character, the tables computed from the .out files the way hypo_tremor_statistics does
    for f, text in dev_stat.items():
        assert open(tmp_path / f).read() == text, f
    rows = [ln for ln in open(tmp_path / "hypo.stat").read().splitlines()[1:]]
    assert len(rows) == E
    for ln in rows:
        v = [float(ln[9 + 13 * k: 22 + 13 * k]) for k in range(9)]
        assert v[1] <= v[0] <= v[2] and v[4] <= v[3] <= v[5] and v[7] <= v[6] <= v[8]
    # little-endian switch
    r = run_driver(tmp_path, "--chunk", "64", "--little-endian")
    assert r.returncode == 0
    out_le = hio.read_outputs(str(tmp_path), R, E, S, big_endian=False)
    assert np.array_equal(out_le["hypo"], out["hypo"])
