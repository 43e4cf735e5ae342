"""File-level twins of the upstream programs hypo_tremor_measure and hypo_tremor_select (drivers/*.cpp): formats in
(parameter file, station file, STA.merged.env, detected_win.dat, opt_data.*.dat) and out (detected_win.dat, cc_thred.dat,
opt_data.*.dat, regress.dat, selected_win.dat), and the chain measure -> select -> mcmc on the files they hand over."""
import json
import os
import subprocess

import numpy as np
import pytest

import hypotremormcmc_b200 as H
from hypotremormcmc_b200 import io as hio

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MEASURE = os.path.join(ROOT, "drivers", "hypo_tremor_measure_b200")
SELECT = os.path.join(ROOT, "drivers", "hypo_tremor_select_b200")
MCMC = os.path.join(ROOT, "drivers", "hypo_tremor_mcmc_b200")


def dataset(tmp_path, S=8, n_total=4000, dt=1.0, big_endian=True, seed=3):
    """a source under the network radiating bursts: envelopes delayed by distance / vs and scaled by spreading and
    attenuation, plus station noise"""
    rng = np.random.default_rng(seed)
    names = ["N.ST%02d" % i for i in range(S)]
    sx, sy, sz = rng.uniform(-30, 30, S), rng.uniform(-30, 30, S), rng.uniform(0, 1, S)
    d = np.sqrt(sx ** 2 + sy ** 2 + (sz - 7.0) ** 2)
    kern = np.hanning(15)
    src = np.convolve(rng.normal(0, 1, n_total + 200) ** 2, kern, mode="same") * (np.sin(np.arange(n_total + 200) * 2 * np.pi / 1500.0) > 0.3)
    env = np.empty((S, n_total))
    for i in range(S):
        lag = int(round(d[i] / 3.0 / dt))
        own = np.convolve(rng.normal(0, 1, n_total) ** 2, kern, mode="same")
        env[i] = 5.0 * np.exp(-0.02 * d[i]) / d[i] * src[100 - lag:100 - lag + n_total] + 0.02 * own
    hio.write_envelopes(str(tmp_path), names, env, dt, big_endian=big_endian)
    hio.write_upstream_params(str(tmp_path), names, sx, sy, sz, t_win=300.0, t_step=150.0, alpha=0.98, n_pair_thred=12)
    return names, (sx, sy, sz), env


def run(prog, cwd, *args):
    return subprocess.run([prog, "hypo_tremor.in", *args], cwd=cwd, capture_output=True, text=True)


def test_measure_driver_reads_the_reference_formats(tmp_path):
    names, _, env = dataset(tmp_path)
    r = run(MEASURE, tmp_path, "--dry-run")
    assert r.returncode == 0, r.stderr
    d = json.loads(r.stdout)
    assert d["n_sta"] == 8 and d["n_total"] == 4000 and d["dt"] == 1.0 and d["n"] == 300 and d["n_step"] == 150
    assert d["n_win"] == (4000 - 300) // 150 and d["n_pair"] == 28 and d["alpha"] == 0.98 and d["n_pair_thred"] == 12
    assert d["env_first"] == env[0, 0] and d["env_last"] == env[-1, -1]
    # little-endian files need the switch; with the wrong byte order the sampling interval is nonsense
    hio.write_envelopes(str(tmp_path), names, env, 1.0, big_endian=False)
    assert json.loads(run(MEASURE, tmp_path, "--dry-run", "--little-endian").stdout) == d
    # a missing envelope, a different length and a different sampling interval are fatal (src/cls_measurer.f90:121-146)
    hio.write_envelopes(str(tmp_path), names[:1], env[:1, :-10], 1.0)
    r = run(MEASURE, tmp_path, "--dry-run")
    assert r.returncode != 0 and "invalid" in r.stderr
    hio.write_envelopes(str(tmp_path), names[:1], env[:1], 0.5)
    r = run(MEASURE, tmp_path, "--dry-run")
    assert r.returncode != 0 and "invalid delta" in r.stderr
    os.remove(tmp_path / (names[0] + ".merged.env"))
    r = run(MEASURE, tmp_path, "--dry-run")
    assert r.returncode != 0 and "does not exist" in r.stderr
    # the parameter keys of this program are required, the MCMC ones are not
    text = open(tmp_path / "hypo_tremor.in").read()
    open(tmp_path / "hypo_tremor.in", "w").write(text.replace("alpha = 0.98\n", ""))
    r = run(MEASURE, tmp_path, "--dry-run")
    assert r.returncode != 0 and "alpha is not given" in r.stderr
    r = subprocess.run([SELECT], cwd=tmp_path, capture_output=True, text=True)
    assert r.returncode != 0 and "USAGE" in r.stderr


@pytest.mark.gpu
def test_measure_select_mcmc_chain_on_files(tmp_path):
    names, (sx, sy, sz), env = dataset(tmp_path)
    S = len(names)
    r = run(MEASURE, tmp_path)
    assert r.returncode == 0, r.stderr
    # the files equal what the library returns for the same arrays
    g = H.api.detect_windows(env, 300, 150, 0.98, 12)
    det = hio.read_table(tmp_path / "detected_win.dat")
    assert [int(v[0]) for v in det] == list(g["win_id"]) and len(det) > 2
    assert [float(v[1]) for v in det] == [(i - 1) * 150.0 + 150.0 for i in g["win_id"]]
    thr = hio.read_table(tmp_path / "cc_thred.dat")
    assert [v[0] for v in thr[:S - 1]] == [names[0]] * (S - 1) and [v[1] for v in thr[:S - 1]] == names[1:]
    assert np.array_equal(np.array([float(v[2]) for v in thr]), g["cc_thred"])
    m = H.api.measure_windows(env, 1.0, 300, 150, g["win_id"])
    for k, wid in enumerate(g["win_id"]):
        rows = np.array(hio.read_table(tmp_path / ("opt_data.%06d.dat" % wid)), dtype=float)
        assert rows.shape == (S, 7)
        assert np.array_equal(rows[:, 0], sx) and np.array_equal(rows[:, 2], sz)
        assert np.array_equal(rows[:, 3], m["t"][k]) and np.array_equal(rows[:, 4], m["t_stdv"][k])
        assert np.array_equal(rows[:, 5], m["amp"][k]) and np.array_equal(rows[:, 6], m["amp_stdv"][k])
    # ---- select on those files ----
    r = run(SELECT, tmp_path)
    assert r.returncode == 0, r.stderr
    s = H.api.select_events(sx, sy, sz, 7.0, m["t"], m["t_stdv"], m["amp"], m["amp_stdv"])
    reg = np.array(hio.read_table(tmp_path / "regress.dat"), dtype=float)
    assert np.array_equal(reg[:, 0], g["win_id"])
    for col, key in ((1, "vs"), (2, "b"), (3, "t0"), (4, "a0"), (5, "cc_t"), (6, "cc_a")):     # src/hypo_tremor_select.f90:119
        # (a window whose lags are perfectly consistent has zero scatter, hence infinite weights and NaN, as in the reference)
        assert np.array_equal(reg[:, col], s[key], equal_nan=True), key
    sel = hio.read_table(tmp_path / "selected_win.dat")
    assert [int(v[0]) for v in sel] == [int(w) for w, ok in zip(g["win_id"], s["selected"]) if ok]
    # the planted propagation speed comes back for the windows inside a burst
    assert np.isfinite(s["vs"]).sum() >= 3 and np.abs(np.nanmedian(s["vs"]) - 3.0) < 1.0
    # ---- and the MCMC driver reads what was handed over ----
    if len(sel) > 0:
        cfg = H.default_config(n_sta=S, n_events=len(sel), n_procs=2, n_chains=3, n_iter=100, n_burn=10, n_interval=5)
        text = open(tmp_path / "hypo_tremor.in").read()
        syn = H.Synthetic(len(sel), S, 1)
        sub = tmp_path / "mcmc_in"
        hio.write_dataset(str(sub), syn, cfg, station_file="unused.list")
        extra = "".join(ln + "\n" for ln in open(sub / "hypo_tremor.in").read().splitlines()
                        if ln.split("=")[0].strip() not in ("n_procs", "station_file") and not ln.startswith("#"))
        open(tmp_path / "hypo_tremor.in", "w").write(text.replace("n_procs = 1\n", "n_procs = 2\n") + extra)
        r = subprocess.run([MCMC, "hypo_tremor.in", "--dry-run"], cwd=tmp_path, capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        d = json.loads(r.stdout)
        first = np.array(hio.read_table(tmp_path / ("opt_data.%06d.dat" % int(sel[0][0]))), dtype=float)
        assert d["n_events"] == len(sel) and d["n_sta"] == S and d["t_obs00"] == first[0, 3]
