"""TEST INFRASTRUCTURE -- an independent writer of the reference's output files from an oracle run.

What the reference driver writes (src/hypo_tremor_mcmc.f90:216-233,270-280, src/cls_mcmc.f90:371-420,
src/cls_parallel.f90:244-281), restated with `struct`, not shared with drivers/htm_files.hpp nor with
hypotremormcmc_b200/io.py's readers:
  per virtual rank RR: vs.RR.out, hypo.RR.out, t_corr.RR.out, qs.RR.out, a_corr.RR.out -- stream records
  `int32 iteration, float64 values...` of the cold chains after burn-in -- and likelihoodRR.out (sic: no dot)
  with `int32 iteration, float64 L` including the burn-in; big-endian under the shipped -fconvert flag
  (src/Makefile:9,14,18);
  proposal_count.txt: seven lines '(A,2I10)' of '"label"' // n_propose // n_accept, labels character(5).
"""
import os
import struct

LABELS = ["vs   ", "t_cor", "qs   ", "a_cor", "x    ", "y    ", "z    "]  # character(5), src/cls_mcmc.f90:83


def _records(path, iters, rows, big_endian):
    e = ">" if big_endian else "<"
    with open(path, "wb") as f:
        for it, row in zip(iters, rows):
            vals = [float(v) for v in (row if hasattr(row, "__len__") else [row])]
            f.write(struct.pack(e + "i", int(it)))
            f.write(struct.pack(e + "%dd" % len(vals), *vals))


def write_reference_outputs(dirname, oracle, n_procs, big_endian=True, count_width=10):
    """The six families of every virtual rank and proposal_count.txt from a finished oracle run."""
    os.makedirs(dirname, exist_ok=True)
    for r in range(n_procs):
        s = oracle.fetch_samples(r)
        it_l, lik = oracle.fetch_likelihood(r)
        _records(os.path.join(dirname, "vs.%02d.out" % r), s["iter"], s["vs"], big_endian)
        _records(os.path.join(dirname, "hypo.%02d.out" % r), s["iter"], s["hypo"], big_endian)
        _records(os.path.join(dirname, "t_corr.%02d.out" % r), s["iter"], s["t_corr"], big_endian)
        _records(os.path.join(dirname, "qs.%02d.out" % r), s["iter"], s["qs"], big_endian)
        _records(os.path.join(dirname, "a_corr.%02d.out" % r), s["iter"], s["a_corr"], big_endian)
        _records(os.path.join(dirname, "likelihood%02d.out" % r), it_l, lik, big_endian)
    p, a = oracle.get_counts()
    with open(os.path.join(dirname, "proposal_count.txt"), "w") as f:
        for k in range(7):
            f.write('"%s"%*d%*d\n' % (LABELS[k], count_width, int(p[k]), count_width, int(a[k])))


FAMILIES = ["vs.%02d.out", "hypo.%02d.out", "t_corr.%02d.out", "qs.%02d.out", "a_corr.%02d.out", "likelihood%02d.out"]
