import sys, os, numpy as np
sys.path.insert(0, ".")
import hypotremormcmc_b200 as H
from oracle.pyoracle import Oracle
NOSOLVE = dict(solve_vs=0, solve_t_corr=0, solve_qs=0, solve_a_corr=0)
for S in (20, 50):
  for kernel, slots in ((2,1),(2,2),(1,0)):
    E, R, K = 24, 4, 8
    syn = H.Synthetic(E, S, 17)
    cfg = H.default_config(n_sta=S, n_events=E, n_procs=R, n_chains=K, n_cool=1, n_iter=40, n_burn=0, n_interval=10, mode=H.MODE_FACTORISED, precision=32, kernel=kernel, lane_slots=slots, **NOSOLVE)
    o = Oracle(cfg, syn)
    with H.HypoTremorB200(cfg) as g:
        g.load(syn); g.init_chains(); tr,_ = g.run_traced(1, 40)
        errs=[]; Ls=[]
        for r in range(R):
            for k in range(K):
                st = g.get_chain_state(r, k)
                _, pe = o.loglik(st["hypo"][None,:], np.zeros((1,S)), np.zeros((1,S)), [cfg.prior_vs],[cfg.prior_qs], per_event=True)
                errs.append(np.abs(pe[0]-tr["log_likelihood"][-1,:,r,k])); Ls.append(np.abs(pe[0]))
        errs=np.array(errs).ravel(); Ls=np.array(Ls).ravel()
        i=np.argmax(errs)
        print(os.environ.get("HTM_B200_LIB","new")[-12:], S, kernel, slots, "max abs %.3g at |L|=%.3g ; max rel %.3g ; median abs %.3g ; max abs for |L|<500: %.3g" % (errs[i], Ls[i], np.max(errs/np.maximum(1,Ls)), np.median(errs), errs[Ls<500].max()))
