"""Synthetic tremor events for parity tests and benchmarks.

No data ships with the reference, so the generator is ours.  The forward model is the one
of src/cls_forward.f90:155-160,242-248 (straight ray, f = 5 Hz); observations are made
zero-mean per event like the measurer's outputs (src/cls_measurer.f90:440-446,503-509).
"""
import numpy as np

FREQ = 5.0
TRUE_VS = 3.0
TRUE_QS = 250.0


class Synthetic:
    """Arrays are float64, laid out as the C ABI wants them: (n_events, n_sta) C-order ==
    Fortran (n_sta, n_events) column-major."""

    def __init__(self, n_events, n_sta, seed):
        rng = np.random.default_rng(seed)
        self.n_events, self.n_sta, self.seed = n_events, n_sta, seed
        self.sta_x = rng.uniform(-50.0, 50.0, n_sta)
        self.sta_y = rng.uniform(-50.0, 50.0, n_sta)
        self.sta_z = rng.uniform(0.0, 3.0, n_sta)
        self.true_x = rng.uniform(-30.0, 30.0, n_events)
        self.true_y = rng.uniform(-30.0, 30.0, n_events)
        self.true_z = rng.uniform(5.0, 15.0, n_events)
        self.t_stdv = rng.uniform(0.2, 0.6, (n_events, n_sta))
        self.a_stdv = rng.uniform(0.1, 0.3, (n_events, n_sta))
        d = np.sqrt((self.true_x[:, None] - self.sta_x[None, :]) ** 2
                    + (self.true_y[:, None] - self.sta_y[None, :]) ** 2
                    + (self.true_z[:, None] - self.sta_z[None, :]) ** 2)
        b = np.pi * FREQ / (TRUE_QS * TRUE_VS)
        t = d / TRUE_VS + rng.standard_normal((n_events, n_sta)) * self.t_stdv
        a = -d * b - np.log(d) + rng.standard_normal((n_events, n_sta)) * self.a_stdv
        self.t_obs = np.ascontiguousarray(t - t.mean(axis=1, keepdims=True))
        self.a_obs = np.ascontiguousarray(a - a.mean(axis=1, keepdims=True))
        # obs%make_initial_guess, src/cls_obs_data.f90:120-134: station of max a_obs
        # (maxloc returns the first maximum, as np.argmax does)
        ista = np.argmax(self.a_obs, axis=1)
        self.x_mu = self.sta_x[ista].copy()
        self.y_mu = self.sta_y[ista].copy()

    def shard(self, rank, count):
        """Contiguous block of events for shard `rank` of `count` (same split as the C ABI)."""
        lo, hi = shard_bounds(self.n_events, rank, count)
        s = object.__new__(Synthetic)
        s.__dict__.update(self.__dict__)
        s.n_events = hi - lo
        for name in ("true_x", "true_y", "true_z", "x_mu", "y_mu"):
            setattr(s, name, getattr(self, name)[lo:hi].copy())
        for name in ("t_obs", "t_stdv", "a_obs", "a_stdv"):
            setattr(s, name, np.ascontiguousarray(getattr(self, name)[lo:hi]))
        s.event_offset = lo
        return s


def shard_bounds(n_events, rank, count):
    base, rem = divmod(n_events, count)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)
