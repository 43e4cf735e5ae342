/* Minimal plain-C client of the C ABI (include/htm_b200.h): the calls a host program makes, in order, for a
 * factorised run on synthetic numbers.  It is what a cgo / ISO_C_BINDING / JNI shim boils down to.
 *
 *   gcc -std=c99 -O2 -I../include minimal_c_client.c -L../hypotremormcmc_b200/csrc -lhtm_b200 \
 *       -Wl,-rpath,'$ORIGIN/../hypotremormcmc_b200/csrc' -lm -o minimal_c_client
 *
 * Without a CUDA device htm_create fails loudly (the library has no CPU fallback) and the program exits 1. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "htm_b200.h"

static void check(htm_handle h, int32_t rc, const char* where) {
  if (rc == HTM_OK) return;
  char msg[512];
  htm_last_error(h, msg, (int32_t)sizeof msg);
  fprintf(stderr, "ERROR: %s: %s\n", where, msg);
  exit(1);
}

int main(void) {
  enum { S = 12, E = 64 };
  static double sta_x[S], sta_y[S], sta_z[S], t_obs[E * S], t_sd[E * S], a_obs[E * S], a_sd[E * S], x_mu[E], y_mu[E];
  for (int j = 0; j < S; ++j) {
    sta_x[j] = 40.0 * cos(0.52 * j);
    sta_y[j] = 40.0 * sin(0.52 * j);
    sta_z[j] = 0.1 * j;
  }
  for (int e = 0; e < E; ++e) {
    const double x = 10.0 * cos(0.1 * e), y = 10.0 * sin(0.1 * e), z = 8.0 + 0.05 * e;
    double mt = 0.0, ma = 0.0;
    for (int j = 0; j < S; ++j) {  /* [E][S] in C == (n_sta, n_events) column-major in Fortran */
      const double d = sqrt((x - sta_x[j]) * (x - sta_x[j]) + (y - sta_y[j]) * (y - sta_y[j]) + (z - sta_z[j]) * (z - sta_z[j]));
      t_obs[e * S + j] = d / 3.0;
      a_obs[e * S + j] = -d * 3.14159265358979 * 5.0 / (250.0 * 3.0) - log(d);
      t_sd[e * S + j] = 0.4;
      a_sd[e * S + j] = 0.2;
      mt += t_obs[e * S + j] / S;
      ma += a_obs[e * S + j] / S;
    }
    int best = 0;
    for (int j = 0; j < S; ++j) {
      t_obs[e * S + j] -= mt;  /* relative data */
      a_obs[e * S + j] -= ma;
      if (a_obs[e * S + j] > a_obs[e * S + best]) best = j;
    }
    x_mu[e] = sta_x[best];
    y_mu[e] = sta_y[best];
  }

  htm_config cfg;
  htm_handle h = NULL;
  check(NULL, htm_config_default(&cfg), "htm_config_default");
  cfg.n_sta = S;
  cfg.n_events = E;
  cfg.n_procs = 2;
  cfg.n_chains = 8;
  cfg.n_iter = 2000;
  cfg.n_burn = 500;
  cfg.n_interval = 100;
  cfg.mode = HTM_MODE_FACTORISED;
  cfg.solve_vs = cfg.solve_qs = cfg.solve_t_corr = cfg.solve_a_corr = 0;
  cfg.max_samples = 32;
  check(h, htm_create(&h, &cfg), "htm_create");
  check(h, htm_set_stations(h, sta_x, sta_y, sta_z), "htm_set_stations");
  check(h, htm_set_observations(h, t_obs, t_sd, a_obs, a_sd), "htm_set_observations");
  check(h, htm_set_xy_prior(h, x_mu, y_mu), "htm_set_xy_prior");
  check(h, htm_init_chains(h), "htm_init_chains");
  check(h, htm_run(h, 1, cfg.n_iter), "htm_run");
  check(h, htm_synchronize(h), "htm_synchronize");

  static int32_t iter[32];
  static double hypo[32 * 3 * E];
  int32_t n = 0;
  check(h, htm_fetch_samples(h, 0, 32, &n, iter, NULL, NULL, hypo, NULL, NULL), "htm_fetch_samples");
  int64_t n_prop[7], n_acc[7];
  check(h, htm_get_counts(h, n_prop, n_acc), "htm_get_counts");
  printf("%d records of virtual rank 0; last: iteration %d, event 0 at (%.2f, %.2f, %.2f); cold x/y/z proposals %lld, accepted %lld\n",
         (int)n, n ? (int)iter[n - 1] : 0, n ? hypo[(n - 1) * 3 * E] : 0.0, n ? hypo[(n - 1) * 3 * E + 1] : 0.0,
         n ? hypo[(n - 1) * 3 * E + 2] : 0.0, (long long)(n_prop[4] + n_prop[5] + n_prop[6]),
         (long long)(n_acc[4] + n_acc[5] + n_acc[6]));
  htm_destroy(h);
  return 0;
}
