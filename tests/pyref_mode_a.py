"""TEST INFRASTRUCTURE.  A second, independent restatement of the reference's hypo_tremor_mcmc main loop in
plain Python, written from the Fortran with the Fortran's OWN structure (type-bound objects, deep copies of
the five models at every get_mc / propose / set_mc, `partially_update` adding station terms one by one) --
deliberately unlike the in-place C++ oracle, so that a transcription slip in either shows up as a mismatch.
Small cases only (pure-Python loops)."""
import copy
import math

EPS = 2.220446049250313e-16
M32 = 0xFFFFFFFF
LOG_2PI_HALF = 0.5 * math.log(2.0 * math.acos(-1.0))


class Random:  # src/mod_random.f90
    def __init__(self, rank):
        j = rank + 1
        self.s = [((i * j ** 4 + 1000 * i * j ** 2 + i) & M32) for i in (5551111, 453222, 4444431, 6765)]

    def _w(self):
        x, y, z, w = self.s
        t = (x ^ ((x << 11) & M32)) & M32
        w2 = ((w ^ (w >> 19)) ^ (t ^ (t >> 8))) & M32
        self.s = [y, z, w, w2]
        return w2 - (1 << 32) if w2 & 0x80000000 else w2

    def rand_u(self):
        return (float(self._w()) + 2.0 ** 31) / 2.0 ** 32

    def rand_u2(self):
        return (float(self._w()) + 2.0 ** 31 + 0.5) / 2.0 ** 32

    def rand_g(self):
        v1 = self.rand_u2()
        v2 = self.rand_u2()
        return math.sqrt(-2.0 * math.log(v1)) * math.cos(2.0 * math.acos(-1.0) * v2)

    def rand_r(self):
        return math.sqrt(-2.0 * math.log(self.rand_u2()))


class Model:  # src/cls_model.f90
    def __init__(self, nx):
        self.nx = nx
        self.prior_type = [0] * nx
        self.x = [0.0] * nx
        self.mu = [0.0] * nx
        self.sigma = [0.0] * nx
        self.step = [0.0] * nx

    def generate(self, rng):
        for i in range(self.nx):
            if self.prior_type[i] == 0:
                self.x[i] = self.mu[i] + rng.rand_g() * self.sigma[i]
            else:
                self.x[i] = self.mu[i] + rng.rand_r() * self.sigma[i]

    def perturb(self, i, rng):
        x_old = self.x[i]
        x_new = x_old + rng.rand_g() * self.step[i]
        self.x[i] = x_new
        lpr = -((x_new - self.mu[i]) ** 2 - (x_old - self.mu[i]) ** 2) / (2.0 * self.sigma[i] * self.sigma[i])
        ok = True
        if self.prior_type[i] == 1:
            if x_new <= self.mu[i]:
                ok = False
            else:
                lpr = lpr + math.log(x_new - self.mu[i]) - math.log(x_old - self.mu[i])
        return lpr, ok


class Forward:  # src/cls_forward.f90
    def __init__(self, syn, use_time, use_amp):
        self.S, self.E = syn.n_sta, syn.n_events
        self.sx, self.sy, self.sz = list(syn.sta_x), list(syn.sta_y), list(syn.sta_z)
        self.use_time, self.use_amp = use_time, use_amp
        self.t_obs = [list(r) for r in syn.t_obs]
        self.a_obs = [list(r) for r in syn.a_obs]
        self.t_stdv = [list(r) for r in syn.t_stdv]
        self.a_stdv = [list(r) for r in syn.a_stdv]
        self.t_prec = [[0.0] * self.S for _ in range(self.E)]
        self.a_prec = [[0.0] * self.S for _ in range(self.E)]
        self.log_t = [[0.0] * self.S for _ in range(self.E)]
        self.log_a = [[0.0] * self.S for _ in range(self.E)]
        for i in range(self.E):
            for j in range(self.S):
                if self.t_stdv[i][j] > 1.e-16:
                    self.log_t[i][j] = math.log(self.t_stdv[i][j])
                    self.t_prec[i][j] = 1.0 / self.t_stdv[i][j] ** 2
                    self.log_a[i][j] = math.log(self.a_stdv[i][j])
                    self.a_prec[i][j] = 1.0 / self.a_stdv[i][j] ** 2
                else:
                    self.log_t[i][j] = 1.0
                    self.t_stdv[i][j] = 1.0
                    self.t_prec[i][j] = 1.0
                    self.log_a[i][j] = 1.0
                    self.a_stdv[i][j] = 1.0
                    self.a_prec[i][j] = 1.0

    def _dist(self, hypo, i, j):
        return math.sqrt((hypo.x[3 * i] - self.sx[j]) ** 2 + (hypo.x[3 * i + 1] - self.sy[j]) ** 2
                         + (hypo.x[3 * i + 2] - self.sz[j]) ** 2)

    def t_single(self, i, hypo, t_corr, vs):
        beta = vs.x[0]
        t = [self._dist(hypo, i, j) / beta - t_corr.x[j] for j in range(self.S)]
        num = 0.0
        for j in range(self.S):
            num += self.t_prec[i][j] * (t[j] - self.t_obs[i][j])
        den = 0.0
        for j in range(self.S):
            den += self.t_prec[i][j]
        m = num / den
        return [v - m for v in t]

    def a_single(self, i, hypo, a_corr, qs, vs):
        q, beta = qs.x[0], vs.x[0]
        pi = math.acos(-1.0)
        a = []
        for j in range(self.S):
            d = self._dist(hypo, i, j)
            a.append(-d * pi * 5.0 / (q * beta) - math.log(d) - a_corr.x[j])
        num = 0.0
        for j in range(self.S):
            num += self.a_prec[i][j] * (a[j] - self.a_obs[i][j])
        den = 0.0
        for j in range(self.S):
            den += self.a_prec[i][j]
        m = num / den
        return [v - m for v in a]

    def full(self, hypo, t_corr, vs, a_corr, qs):
        ll = 0.0
        if self.use_time:
            syn = [self.t_single(i, hypo, t_corr, vs) for i in range(self.E)]
            for i in range(self.E):
                for j in range(self.S):
                    ll = ll - (self.t_obs[i][j] - syn[i][j]) ** 2 / (2.0 * self.t_stdv[i][j] ** 2) - LOG_2PI_HALF - self.log_t[i][j]
        if self.use_amp:
            syn = [self.a_single(i, hypo, a_corr, qs, vs) for i in range(self.E)]
            for i in range(self.E):
                for j in range(self.S):
                    ll = ll - (self.a_obs[i][j] - syn[i][j]) ** 2 / (2.0 * self.a_stdv[i][j] ** 2) - LOG_2PI_HALF - self.log_a[i][j]
        return ll

    def partial(self, evt_id, hypo_old, ll_old, hypo, t_corr, vs, a_corr, qs):
        i = evt_id - 1
        ll = ll_old
        if self.use_time:
            s = self.t_single(i, hypo_old, t_corr, vs)
            for j in range(self.S):
                ll = ll + (self.t_obs[i][j] - s[j]) ** 2 / (2.0 * self.t_stdv[i][j] ** 2) + LOG_2PI_HALF + self.log_t[i][j]
            s = self.t_single(i, hypo, t_corr, vs)
            for j in range(self.S):
                ll = ll - (self.t_obs[i][j] - s[j]) ** 2 / (2.0 * self.t_stdv[i][j] ** 2) - LOG_2PI_HALF - self.log_t[i][j]
        if self.use_amp:
            s = self.a_single(i, hypo_old, a_corr, qs, vs)
            for j in range(self.S):
                ll = ll + (self.a_obs[i][j] - s[j]) ** 2 / (2.0 * self.a_stdv[i][j] ** 2) + LOG_2PI_HALF + self.log_a[i][j]
            s = self.a_single(i, hypo, a_corr, qs, vs)
            for j in range(self.S):
                ll = ll - (self.a_obs[i][j] - s[j]) ** 2 / (2.0 * self.a_stdv[i][j] ** 2) - LOG_2PI_HALF - self.log_a[i][j]
        return ll


class Mcmc:  # src/cls_mcmc.f90
    def __init__(self, hypo, t_corr, vs, a_corr, qs, cfg):
        self.hypo, self.t_corr, self.vs, self.a_corr, self.qs = (copy.deepcopy(m) for m in (hypo, t_corr, vs, a_corr, qs))
        self.n_events, self.n_sta = hypo.nx // 3, t_corr.nx
        self.n_propose, self.n_accept = [0] * 7, [0] * 7
        self.log_likelihood = -9.0e300
        self.temp = 1.0
        self.p_vs = 0.025 if cfg.solve_vs else 0.0
        self.p_t_corr = 0.025 if cfg.solve_t_corr else 0.0
        self.p_qs = 0.025 if cfg.solve_qs else 0.0
        self.p_a_corr = 0.025 if cfg.solve_a_corr else 0.0
        self.i_proposal_type = 0
        self.is_accepted = False

    def propose(self, rng):
        h, tc, vs, ac, qs = (copy.deepcopy(m) for m in (self.hypo, self.t_corr, self.vs, self.a_corr, self.qs))
        a_select = rng.rand_u()
        evt_id = -999
        if a_select < self.p_vs:
            idx = 1
            lpr, ok = vs.perturb(0, rng)
            self.i_proposal_type = 1
        elif a_select < self.p_vs + self.p_t_corr:
            idx = int(rng.rand_u() * self.n_sta) + 1
            lpr, ok = tc.perturb(idx - 1, rng)
            self.i_proposal_type = 2
        elif a_select < self.p_vs + self.p_t_corr + self.p_qs:
            idx = 1
            lpr, ok = qs.perturb(0, rng)
            self.i_proposal_type = 3
        elif a_select < self.p_vs + self.p_t_corr + self.p_qs + self.p_a_corr:
            idx = int(rng.rand_u() * self.n_sta) + 1
            lpr, ok = ac.perturb(idx - 1, rng)
            self.i_proposal_type = 4
        else:
            ident = int(rng.rand_u() * self.n_events) + 1
            icmp = int(rng.rand_u() * 3)
            idx = 3 * ident - icmp
            lpr, ok = h.perturb(idx - 1, rng)
            self.i_proposal_type = 5 + icmp
            evt_id = ident
        return h, tc, vs, ac, qs, lpr, ok, evt_id, idx

    def judge(self, h, tc, vs, ac, qs, ll, lpr, ok, rng):
        if self.temp < 1.0 + EPS:
            self.n_propose[self.i_proposal_type - 1] += 1
        self.is_accepted = False
        if ok:
            ratio = (ll - self.log_likelihood) / self.temp
            ratio = ratio + lpr
            r = rng.rand_u()
            if r >= EPS:
                if math.log(r) <= ratio:
                    self.is_accepted = True
        if self.is_accepted:
            self.hypo, self.t_corr, self.vs, self.a_corr, self.qs = (copy.deepcopy(m) for m in (h, tc, vs, ac, qs))
            self.log_likelihood = ll
            if self.temp < 1.0 + EPS:
                self.n_accept[self.i_proposal_type - 1] += 1


def run(syn, cfg, x_mu, y_mu, n_iter):
    """src/hypo_tremor_mcmc.f90:72-284 with virtual ranks.  Returns (trace, swaps, samples_per_rank)."""
    R, K, E, S = cfg.n_procs, cfg.n_chains, syn.n_events, syn.n_sta
    fwd = Forward(syn, bool(cfg.use_time), bool(cfg.use_amp))
    rngs = [Random(r) for r in range(R)]
    pt = [[None] * K for _ in range(R)]
    for r in range(R):
        for j in range(K):
            tc = Model(S)
            for i in range(S):
                tc.mu[i], tc.sigma[i], tc.step[i], tc.x[i] = cfg.prior_t_corr, cfg.prior_width_t_corr, cfg.step_size_t_corr, cfg.prior_t_corr
            if cfg.solve_t_corr:
                tc.generate(rngs[r])
            ac = Model(S)
            for i in range(S):
                ac.mu[i], ac.sigma[i], ac.step[i], ac.x[i] = cfg.prior_a_corr, cfg.prior_width_a_corr, cfg.step_size_a_corr, cfg.prior_a_corr
            if cfg.solve_a_corr:
                ac.generate(rngs[r])
            h = Model(3 * E)
            for i in range(E):
                h.mu[3 * i], h.sigma[3 * i], h.step[3 * i] = x_mu[i], cfg.prior_width_xy, cfg.step_size_xy
                h.mu[3 * i + 1], h.sigma[3 * i + 1], h.step[3 * i + 1] = y_mu[i], cfg.prior_width_xy, cfg.step_size_xy
                h.mu[3 * i + 2], h.sigma[3 * i + 2], h.step[3 * i + 2] = cfg.prior_z, cfg.prior_width_z, cfg.step_size_z
                h.prior_type[3 * i + 2] = 1
            h.generate(rngs[r])
            vs = Model(1)
            vs.mu[0], vs.sigma[0], vs.step[0], vs.x[0] = cfg.prior_vs, cfg.prior_width_vs, cfg.step_size_vs, cfg.prior_vs
            qs = Model(1)
            qs.mu[0], qs.sigma[0], qs.step[0], qs.x[0] = cfg.prior_qs, cfg.prior_width_qs, cfg.step_size_qs, cfg.prior_qs
            mc = Mcmc(h, tc, vs, ac, qs, cfg)
            if j + 1 <= cfg.n_cool:
                mc.temp = 1.0
            else:
                mc.temp = math.exp((rngs[r].rand_u() * (1.0 - EPS) + EPS) * math.log(cfg.temp_high))
            pt[r][j] = mc
    trace, swaps = [], []
    samples = [[] for _ in range(R)]
    for i in range(1, n_iter + 1):
        for r in range(R):
            for j in range(K):
                mc = copy.deepcopy(pt[r][j])                       # pt%get_mc: a deep copy
                h, tc, vs, ac, qs, lpr, ok, evt_id, idx = mc.propose(rngs[r])
                ll = 0.0
                if ok:
                    if evt_id > 0 and i > 1:
                        ll = fwd.partial(evt_id, copy.deepcopy(mc.hypo), mc.log_likelihood, h, tc, vs, ac, qs)
                    else:
                        ll = fwd.full(h, tc, vs, ac, qs)
                mc.judge(h, tc, vs, ac, qs, ll, lpr, ok, rngs[r])
                pt[r][j] = copy.deepcopy(mc)                       # pt%set_mc
                trace.append((mc.i_proposal_type, idx, int(ok), int(mc.is_accepted), mc.log_likelihood))
                if mc.temp < 1.0 + EPS and i % cfg.n_interval == 1 and i > cfg.n_burn:
                    samples[r].append((i, mc.vs.x[0], list(mc.hypo.x), list(mc.t_corr.x), mc.qs.x[0], list(mc.a_corr.x)))
        # parallel_swap_temperature
        i1 = int(rngs[0].rand_u() * R * K)
        while True:
            i2 = int(rngs[0].rand_u() * R * K)
            if i1 != i2:
                break
        r1, r2, c1, c2 = i1 // K, i2 // K, i1 % K + 1, i2 % K + 1
        m1, m2 = pt[r1][c1 - 1], pt[r2][c2 - 1]
        del_s = (m2.log_likelihood - m1.log_likelihood) * (1.0 / m1.temp - 1.0 / m2.temp)
        rr = rngs[r1].rand_u()
        acc = rr >= EPS and math.log(rr) <= del_s
        if acc:
            m1.temp, m2.temp = m2.temp, m1.temp
        swaps.append((r1, c1, r2, c2, int(acc)))
    counts = ([sum(pt[r][j].n_propose[k] for r in range(R) for j in range(K)) for k in range(7)],
              [sum(pt[r][j].n_accept[k] for r in range(R) for j in range(K)) for k in range(7)])
    return trace, swaps, samples, counts
