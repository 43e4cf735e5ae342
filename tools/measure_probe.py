"""Times the measure stage (htm_measure_windows, csrc/htm_measure.cu) on synthetic envelopes and reports the DFMA rate of
the pair correlation against the measured FP64 peak:   python tools/measure_probe.py [n_win n_sta n_smp]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hypotremormcmc_b200 as H  # noqa: E402


def main():
    W, S, n = (int(v) for v in sys.argv[1:4]) if len(sys.argv) >= 4 else (2000, 50, 300)
    n_step = n // 2
    rng = np.random.default_rng(1)
    kern = np.hanning(21)
    env = np.stack([np.convolve(rng.normal(0, 1, n_step * (W + 1)) ** 2, kern, mode="same") for _ in range(S)])
    win = np.arange(1, W + 1)
    ms = min(H.api.measure_windows(env, 1.0, n, n_step, win)["kernel_ms"] for _ in range(4))
    flop = 2.0 * W * (S * (S - 1) // 2) * n * n
    peak = H.api.measure_fp64_peak()
    print("measure_kernel %d windows x %d stations x %d samples: %.2f ms, %.2f us per window, %.2f TFLOP/s float64 = %.2f of "
          "the measured DFMA peak %.1f" % (W, S, n, ms, 1e3 * ms / W, flop / ms / 1e9, flop / ms / 1e9 / peak, peak))


if __name__ == "__main__":
    main()
