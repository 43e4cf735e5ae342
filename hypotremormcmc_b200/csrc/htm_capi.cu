// C ABI of libhtm_b200.so (include/htm_b200.h): handle, host-side table building, launches,
// result fetches.  All compute is on the device; nothing here evaluates the forward model or
// steps a chain on the CPU.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "htm_kernels.hpp"

using namespace htm;

namespace {
thread_local std::string g_create_error;
}

struct htm_handle_s {
  htm_config cfg;
  std::string err;
  int E = 0, E_total = 0, ev_off = 0, rank_off = 0, S = 0, R = 0, K = 0, C = 0;
  size_t rs = 4;  // sizeof(real)
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  bool have_sta = false, have_obs = false, have_prior = false, tables_ok = false, chains_ready = false;
  // raw float64 inputs: pinned host staging (the library copies inputs during the call and keeps no caller
  // pointer) -> ONE async H2D per set_* call -> device-side table build (htm_tables.cu).  Layout in doubles:
  // obs [4][E][S] | sta [3][S] | xy_mu [2][E] | fixed station terms [2][S]
  double* pin_in = nullptr;
  double* d_in = nullptr;
  bool pin_in_pinned = false;
  cudaEvent_t ev_in = nullptr;  // last H2D out of pin_in: the next set_* call waits for it before overwriting
  size_t off_obs() const { return 0; }
  size_t off_sta() const { return static_cast<size_t>(4) * E * S; }
  size_t off_xy() const { return off_sta() + static_cast<size_t>(3) * S; }
  size_t off_g() const { return off_xy() + static_cast<size_t>(2) * E; }
  size_t n_in() const { return off_g() + static_cast<size_t>(2) * S; }
  double g_vs = 0, g_qs = 0;
  std::vector<double> g_tc, g_ac;
  // device tables
  void *d_sta4 = nullptr, *d_obs4 = nullptr, *d_obs4_raw = nullptr, *d_evc4 = nullptr, *d_prior_xy = nullptr;
  double* d_prior_xy64 = nullptr;
  void* d_obsx = nullptr;
  // mode B state
  void *d_x = nullptr, *d_y = nullptr, *d_z = nullptr, *d_L = nullptr, *d_T = nullptr;
  unsigned long long* d_counts = nullptr;
  uint32_t* d_hist = nullptr;
  void* d_samples = nullptr;
  int rec_cap = 0, rec_origin = 0, rec_pending = 0;
  std::vector<int> cur_samp, cur_lik;  // per rank, in recorded iterations consumed
  // Sample output: every htm_run queues, on a second stream behind its kernel, the device-to-host copy of the
  // ring slots it fills into a pinned host mirror of the ring -- the copy overlaps the next htm_run chunk and
  // htm_fetch_* only wait for the copy event.
  cudaStream_t cstream = nullptr;
  cudaEvent_t ev_run = nullptr, ev_copy = nullptr;
  char* pin_samples = nullptr;  // mode B: mirror of d_samples; mode C: mirror of the seven record arrays
  bool pin_samples_pinned = false;
  bool host_samples_valid = false;
  // mode A state
  double *d_hypo = nullptr, *d_tc = nullptr, *d_ac = nullptr, *d_vs = nullptr, *d_qs = nullptr, *d_temp = nullptr,
         *d_Lc = nullptr;
  unsigned long long* d_chain_counts = nullptr;
  long long* d_cursor = nullptr;
  int32_t* d_status = nullptr;
  // mode C state (device pointers live in gl)
  GibbsLaunch gl;
  std::vector<void*> gibbs_bufs;
  int n_cold_total = 0;
  char* host_hypo_rec = nullptr;  // carved out of pin_samples (alloc_state)
  int* host_rec_chain = nullptr;
  double *host_rec_vs = nullptr, *host_rec_qs = nullptr, *host_rec_L = nullptr, *host_rec_tc = nullptr, *host_rec_ac = nullptr;
  // multi-GPU (event shards): NCCL communicator, created by htm_comm_init
  void* comm = nullptr;
  // event-sharded blocked Gibbs: per-iteration exchange through peer memory (htm_comm_p2p_export / _import)
  double* xch_buf = nullptr;               // this shard's exchange buffer (cudaMalloc: IPC-exportable)
  void* xch_peer[kMaxPeers] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  bool xch_on = false;
  uint32_t xch_epoch = 1;                  // number of the next exchange; advances identically on every shard
  int* d_xch_status = nullptr;
  // device-side posterior store (cfg.summary): hypocentre marginals [3E][store_cap] in the handle's precision,
  // shared-parameter marginals [2 + 2S][store_cap] float64 (blocked-Gibbs mode), filled by every htm_run
  void* d_store_hypo = nullptr;
  double* d_store_shared = nullptr;
  size_t store_cap = 0, store_n = 0;
  int last_partials = 0, last_parity = 0;  // float32 mode C: which accumulator set holds the sums of the last iteration
  // stats
  bool timed = false;
  int64_t last_launches = 0, last_proposals = 0;
};

namespace {

int32_t fail(htm_handle h, int32_t code, const std::string& msg) {
  if (h)
    h->err = msg;
  else
    g_create_error = msg;
  return code;
}
#define HTM_CK(h, call)                                                                      \
  do {                                                                                       \
    cudaError_t e__ = (call);                                                                \
    if (e__ != cudaSuccess)                                                                  \
      return fail(h, HTM_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__));     \
  } while (0)

void free_dev(void* p) {
  if (p) cudaFree(p);
}

void shard_bounds(int n, int rank, int count, int* lo, int* hi) {
  const int base = n / count, rem = n % count;
  *lo = rank * base + (rank < rem ? rank : rem);
  *hi = *lo + base + (rank < rem ? 1 : 0);
}

// Event-sharded blocked Gibbs: after a peer failed to answer an exchange (htm_gibbs_decide.cuh: peer_allreduce)
// the kernels stop taking decisions and the shards have diverged -- every entry point that returns results
// calls this after synchronising the stream and fails hard instead of handing out stale numbers.
int32_t check_exchange(htm_handle h) {
  if (!h->d_xch_status) return HTM_OK;
  int st = 0;
  HTM_CK(h, cudaMemcpy(&st, h->d_xch_status, sizeof(int), cudaMemcpyDeviceToHost));
  if (st != 0)
    return fail(h, HTM_ERR_CUDA,
                "peer-memory exchange timed out: a shard did not reach the same iteration (results of this run are "
                "invalid on every shard; raise HTM_XCH_TIMEOUT_S if the shards are legitimately that far apart)");
  return HTM_OK;
}

// pinned when possible (asynchronous copies), plain host memory otherwise (copies still work, just synchronously)
void* host_alloc(size_t bytes, bool* pinned) {
  void* p = nullptr;
  if (cudaHostAlloc(&p, bytes ? bytes : 8, cudaHostAllocDefault) == cudaSuccess) {
    *pinned = true;
    return p;
  }
  (void)cudaGetLastError();
  *pinned = false;
  return std::malloc(bytes ? bytes : 8);
}
void host_free(void* p, bool pinned) {
  if (!p) return;
  if (pinned)
    cudaFreeHost(p);
  else
    std::free(p);
}

// Stage `n` doubles of one input segment: caller memory -> pinned staging -> device, asynchronously.  The
// previous copy out of the staging buffer must have finished before it is overwritten.
int32_t stage_input(htm_handle h, size_t off, const double* const* src, const size_t* len, int n_src) {
  HTM_CK(h, cudaSetDevice(h->cfg.device));
  HTM_CK(h, cudaEventSynchronize(h->ev_in));
  size_t o = off;
  for (int i = 0; i < n_src; ++i) {
    std::memcpy(h->pin_in + o, src[i], len[i] * sizeof(double));
    o += len[i];
  }
  HTM_CK(h, cudaMemcpyAsync(h->d_in + off, h->pin_in + off, (o - off) * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  HTM_CK(h, cudaEventRecord(h->ev_in, h->stream));
  h->tables_ok = false;
  return HTM_OK;
}

Tables tables_of(htm_handle h);

// sta4 / obs4 / obs4_raw / evc4 / prior_xy (layouts: htm_forward.cuh) are built ON THE DEVICE from the raw
// float64 arrays (htm_tables.cu: init_forward's rules, src/cls_forward.f90:71-92, in float64): no host pass
// over the observations, no synchronisation.
int32_t build_tables(htm_handle h) {
  const double* gsrc[2] = {h->g_tc.data(), h->g_ac.data()};
  const size_t glen[2] = {static_cast<size_t>(h->S), static_cast<size_t>(h->S)};
  const int32_t rc = stage_input(h, h->off_g(), gsrc, glen, 2);
  if (rc != HTM_OK) return rc;
  TableBuild b;
  b.E = h->E;
  b.S = h->S;
  b.use_time = h->cfg.use_time != 0;
  b.use_amp = h->cfg.use_amp != 0;
  b.obs_in = h->d_in + h->off_obs();
  b.sta_xyz = h->d_in + h->off_sta();
  b.xy_mu = h->have_prior ? h->d_in + h->off_xy() : nullptr;
  b.g_tc_ac = h->d_in + h->off_g();
  b.sta4 = h->d_sta4;
  b.obs4 = h->d_obs4;
  b.obs4_raw = h->d_obs4_raw;
  b.evc4 = h->d_evc4;
  b.prior_xy = h->d_prior_xy;
  b.prior_xy64 = h->d_prior_xy64;
  HTM_CK(h, launch_build_tables(h->cfg.precision, b, h->stream));
  h->tables_ok = true;
  return HTM_OK;
}

Tables tables_of(htm_handle h);

int32_t ensure_tables(htm_handle h) {
  if (h->tables_ok) return HTM_OK;
  if (!h->have_sta) return fail(h, HTM_ERR_STATE, "stations not set (htm_set_stations)");
  if (!h->have_obs) return fail(h, HTM_ERR_STATE, "observations not set (htm_set_observations)");
  const int32_t rc = build_tables(h);
  if (rc != HTM_OK) return rc;
  if (h->cfg.mode == HTM_MODE_BLOCKED_GIBBS && h->cfg.precision == HTM_PRECISION_F32) {
    // expanded station-pair rows for the float32 joint-chain sweep (layout: htm_forward.cuh)
    const int xrow = 4 + 4 * (h->S / 2);  // htm_gibbs_f32.cu: f32_xrow
    HTM_CK(h, launch_expand_obs(tables_of(h), h->E, h->S, h->d_obsx, h->stream));
    h->gl.obsx = h->d_obsx;
    h->gl.xrow = xrow;
  }
  return HTM_OK;
}

Tables tables_of(htm_handle h) {
  Tables t;
  t.sta4 = h->d_sta4;
  t.obs4 = h->d_obs4;
  t.obs4_raw = h->d_obs4_raw;
  t.evc4 = h->d_evc4;
  t.prior_xy = h->d_prior_xy;
  return t;
}

FactLaunch fact_launch_of(htm_handle h) {
  FactLaunch a;
  a.precision = h->cfg.precision;
  a.kernel = h->cfg.kernel == HTM_KERNEL_AUTO ? HTM_KERNEL_LANE_PER_CHAIN : h->cfg.kernel;
  a.slots = h->cfg.lane_slots;  // 0 = choose
  a.tab = tables_of(h);
  a.x = h->d_x;
  a.y = h->d_y;
  a.z = h->d_z;
  a.L = h->d_L;
  a.T = h->d_T;
  a.E = h->E;
  a.S = h->S;
  a.R = h->R;
  a.K = h->K;
  a.n_cool = h->cfg.n_cool;
  a.n_burn = h->cfg.n_burn;
  a.n_interval = h->cfg.n_interval;
  a.seed = h->cfg.seed;
  a.event_offset = static_cast<uint32_t>(h->ev_off);
  a.vs = h->g_vs;
  a.qs = h->g_qs;
  a.prior_z = h->cfg.prior_z;
  a.width_z = h->cfg.prior_width_z;
  a.width_xy = h->cfg.prior_width_xy;
  a.step_xy = h->cfg.step_size_xy;
  a.step_z = h->cfg.step_size_z;
  a.counts = h->d_counts;
  a.samples = h->d_samples;
  a.rec_origin = h->rec_origin;
  a.rec_cap = h->rec_cap;
  a.hist = h->d_hist;
  a.hist_bins = h->cfg.hist_bins;
  a.hist_hw = h->cfg.hist_xy_halfwidth;
  a.hist_zmax = h->cfg.hist_z_max;
  return a;
}

bool record_ids(int first, int last, int n_interval, int* m_lo, int* m_hi);

int32_t alloc_state(htm_handle h) {
  const size_t nB = static_cast<size_t>(h->E) * h->R * h->K;
  {  // input staging + device tables (sizes are fixed by the configuration)
    const size_t ES = static_cast<size_t>(h->E) * h->S, r4 = 4 * h->rs;
    h->pin_in = static_cast<double*>(host_alloc(h->n_in() * sizeof(double), &h->pin_in_pinned));
    if (!h->pin_in) return fail(h, HTM_ERR_CUDA, "out of host memory for the input staging buffer");
    HTM_CK(h, cudaMalloc(&h->d_in, h->n_in() * sizeof(double)));
    HTM_CK(h, cudaMalloc(&h->d_sta4, h->S * r4));
    HTM_CK(h, cudaMalloc(&h->d_obs4_raw, ES * r4));
    // the table with the fixed station terms folded in serves the factorised mode only
    if (h->cfg.mode == HTM_MODE_FACTORISED) HTM_CK(h, cudaMalloc(&h->d_obs4, ES * r4));
    HTM_CK(h, cudaMalloc(&h->d_evc4, h->E * r4));
    HTM_CK(h, cudaMalloc(&h->d_prior_xy, h->E * 2 * h->rs));
    if (h->cfg.mode == HTM_MODE_REPLAY) HTM_CK(h, cudaMalloc(&h->d_prior_xy64, h->E * 2 * sizeof(double)));
    if (h->cfg.mode == HTM_MODE_BLOCKED_GIBBS && h->cfg.precision == HTM_PRECISION_F32)
      HTM_CK(h, cudaMalloc(&h->d_obsx, static_cast<size_t>(h->E) * (4 + 4 * (h->S / 2)) * 16));
  }
  if (h->cfg.mode == HTM_MODE_FACTORISED) {
    for (void** p : {&h->d_x, &h->d_y, &h->d_z, &h->d_L, &h->d_T}) {
      HTM_CK(h, cudaMalloc(p, nB * h->rs));
      HTM_CK(h, cudaMemset(*p, 0, nB * h->rs));
    }
    if (h->cfg.hist_bins > 0) {
      const size_t nh = static_cast<size_t>(h->E) * 3 * h->cfg.hist_bins * sizeof(uint32_t);
      HTM_CK(h, cudaMalloc(&h->d_hist, nh));
      HTM_CK(h, cudaMemset(h->d_hist, 0, nh));
    }
    if (h->cfg.max_samples > 0) {
      h->rec_cap = h->cfg.max_samples;
      const size_t ns = static_cast<size_t>(h->rec_cap) * h->R * h->cfg.n_cool * h->E * 4 * h->rs;
      HTM_CK(h, cudaMalloc(&h->d_samples, ns));
      h->pin_samples = static_cast<char*>(host_alloc(ns, &h->pin_samples_pinned));
      if (!h->pin_samples) return fail(h, HTM_ERR_CUDA, "out of host memory for the sample ring mirror");
    }
    h->cur_samp.assign(h->R, 0);
    h->cur_lik.assign(h->R, 0);
  } else if (h->cfg.mode == HTM_MODE_REPLAY) {
    const size_t C = h->C;
    HTM_CK(h, cudaMalloc(&h->d_hypo, C * 3 * h->E * sizeof(double)));
    HTM_CK(h, cudaMalloc(&h->d_tc, C * h->S * sizeof(double)));
    HTM_CK(h, cudaMalloc(&h->d_ac, C * h->S * sizeof(double)));
    HTM_CK(h, cudaMalloc(&h->d_vs, C * sizeof(double)));
    HTM_CK(h, cudaMalloc(&h->d_qs, C * sizeof(double)));
    HTM_CK(h, cudaMalloc(&h->d_temp, C * sizeof(double)));
    HTM_CK(h, cudaMalloc(&h->d_Lc, C * sizeof(double)));
    HTM_CK(h, cudaMalloc(&h->d_chain_counts, C * 14 * sizeof(unsigned long long)));
    HTM_CK(h, cudaMemset(h->d_chain_counts, 0, C * 14 * sizeof(unsigned long long)));
    HTM_CK(h, cudaMalloc(&h->d_cursor, h->R * sizeof(long long)));
    HTM_CK(h, cudaMalloc(&h->d_status, sizeof(int32_t)));
  }
  HTM_CK(h, cudaMalloc(&h->d_counts, 14 * sizeof(unsigned long long)));
  HTM_CK(h, cudaMemset(h->d_counts, 0, 14 * sizeof(unsigned long long)));
  if (h->cfg.mode == HTM_MODE_BLOCKED_GIBBS) {
    GibbsLaunch& g = h->gl;
    const size_t J = h->C, E = h->E, S = h->S;
    // partial sums per chain: one per 32-event tile, or one per CTA of a one-wave grid (at most 8 CTAs per SM)
    const size_t nt = std::max<size_t>((E + 31) / 32, 8 * 160);
    g.part_tiles = static_cast<int>(nt);
    h->n_cold_total = h->R * h->cfg.n_cool;
    auto grab = [&](void** p, size_t bytes) -> cudaError_t {
      cudaError_t e = cudaMalloc(p, bytes ? bytes : 8);
      if (e == cudaSuccess) {
        h->gibbs_bufs.push_back(*p);
        e = cudaMemset(*p, 0, bytes ? bytes : 8);
      }
      return e;
    };
    if (h->cfg.precision == HTM_PRECISION_F32) {  // htm_gibbs_f32.cu: float4 state arrays
      HTM_CK(h, grab(&g.sH, J * E * 16));
      HTM_CK(h, grab(&g.sM, J * E * 16));
      HTM_CK(h, grab(&g.sQ, J * E * 16));
      HTM_CK(h, grab(&g.sP, J * E * 16));
    } else {
      HTM_CK(h, grab(&g.hx, J * E * h->rs));
      HTM_CK(h, grab(&g.hy, J * E * h->rs));
      HTM_CK(h, grab(&g.hz, J * E * h->rs));
      HTM_CK(h, grab(&g.hLe, J * E * h->rs));
    }
    HTM_CK(h, grab(&g.hLp, J * E * h->rs));
    HTM_CK(h, grab(reinterpret_cast<void**>(&g.g_vs), J * 8));
    HTM_CK(h, grab(reinterpret_cast<void**>(&g.g_qs), J * 8));
    HTM_CK(h, grab(reinterpret_cast<void**>(&g.g_tc), J * S * 8));
    HTM_CK(h, grab(reinterpret_cast<void**>(&g.g_ac), J * S * 8));
    HTM_CK(h, grab(reinterpret_cast<void**>(&g.g_T), J * 8));
    HTM_CK(h, grab(reinterpret_cast<void**>(&g.g_L), J * 8));
    HTM_CK(h, grab(reinterpret_cast<void**>(&g.prop_which), J * 4));
    HTM_CK(h, grab(reinterpret_cast<void**>(&g.prop_idx), J * 4));
    HTM_CK(h, grab(reinterpret_cast<void**>(&g.a_prev), J * 4));
    HTM_CK(h, grab(reinterpret_cast<void**>(&g.slot_of), J * 4));
    HTM_CK(h, grab(reinterpret_cast<void**>(&g.prop_xnew), J * 8));
    HTM_CK(h, grab(reinterpret_cast<void**>(&g.prop_lpr), J * 8));
    HTM_CK(h, grab(reinterpret_cast<void**>(&g.part_cur), 4 * J * nt * 8));  // [2 buffers][cur, prop][J][tiles]
    g.part_prop = g.part_cur + J * nt;
    HTM_CK(h, grab(reinterpret_cast<void**>(&g.done_counter), 8));
    HTM_CK(h, grab(reinterpret_cast<void**>(&g.totals), 4 * J * 8));  // float32: [cur, prop][J][2 limbs]
    if (h->cfg.max_samples > 0) {
      h->rec_cap = h->cfg.max_samples;
      const size_t n = static_cast<size_t>(h->rec_cap) * h->n_cold_total;
      HTM_CK(h, grab(&g.hypo_rec, n * E * 4 * h->rs));
      HTM_CK(h, grab(reinterpret_cast<void**>(&g.rec_chain), n * 4));
      HTM_CK(h, grab(reinterpret_cast<void**>(&g.rec_vs), n * 8));
      HTM_CK(h, grab(reinterpret_cast<void**>(&g.rec_qs), n * 8));
      HTM_CK(h, grab(reinterpret_cast<void**>(&g.rec_L), n * 8));
      HTM_CK(h, grab(reinterpret_cast<void**>(&g.rec_tc), n * S * 8));
      HTM_CK(h, grab(reinterpret_cast<void**>(&g.rec_ac), n * S * 8));
      // host mirror (pinned): hypo records, then the 8-byte arrays, then the chain ids
      const size_t b_h = n * E * 4 * h->rs, b_d = n * 8, b_s = n * S * 8;
      h->pin_samples = static_cast<char*>(host_alloc(b_h + 3 * b_d + 2 * b_s + n * 4, &h->pin_samples_pinned));
      if (!h->pin_samples) return fail(h, HTM_ERR_CUDA, "out of host memory for the sample ring mirror");
      char* q = h->pin_samples;
      h->host_hypo_rec = q; q += b_h;
      h->host_rec_vs = reinterpret_cast<double*>(q); q += b_d;
      h->host_rec_qs = reinterpret_cast<double*>(q); q += b_d;
      h->host_rec_L = reinterpret_cast<double*>(q); q += b_d;
      h->host_rec_tc = reinterpret_cast<double*>(q); q += b_s;
      h->host_rec_ac = reinterpret_cast<double*>(q); q += b_s;
      h->host_rec_chain = reinterpret_cast<int*>(q);
    }
    h->cur_samp.assign(h->R, 0);
    h->cur_lik.assign(h->R, 0);
  }
  if (h->cfg.summary) {
    // recorded iterations after the burn-in: it = m n_interval + 1 with n_burn < it <= n_iter
    int m_lo = 0, m_hi = -1;
    long n_rec = 0;
    if (record_ids(h->cfg.n_burn + 1, h->cfg.n_iter, h->cfg.n_interval, &m_lo, &m_hi)) n_rec = m_hi - m_lo + 1;
    h->store_cap = static_cast<size_t>(n_rec > 0 ? n_rec : 1) * h->R * h->cfg.n_cool;
    const size_t bytes = static_cast<size_t>(3) * h->E * h->store_cap * h->rs;
    if (bytes > (static_cast<size_t>(96) << 30))
      return fail(h, HTM_ERR_UNSUPPORTED, "summary store would exceed 96 GB: thin more (n_interval) or summarise from the files");
    HTM_CK(h, cudaMalloc(&h->d_store_hypo, bytes));
    if (h->cfg.mode == HTM_MODE_BLOCKED_GIBBS)
      HTM_CK(h, cudaMalloc(&h->d_store_shared, static_cast<size_t>(2 + 2 * h->S) * h->store_cap * sizeof(double)));
  }
  return HTM_OK;
}

// fill the launch-invariant part of the mode-C launch description
void gibbs_launch_of(htm_handle h) {
  GibbsLaunch& g = h->gl;
  g.precision = h->cfg.precision;
  g.tab = tables_of(h);
  g.E = h->E;
  g.S = h->S;
  g.J = h->C;
  g.K = h->K;
  g.n_cool_total = h->n_cold_total;
  g.n_burn = h->cfg.n_burn;
  g.n_interval = h->cfg.n_interval;
  g.seed = h->cfg.seed;
  g.event_offset = static_cast<uint32_t>(h->ev_off);
  g.chain_offset = static_cast<uint32_t>(h->rank_off) * h->K;
  g.J_total = static_cast<uint32_t>(h->cfg.n_procs) * h->K;
  // rank shards are independent ensembles (own swap stream each); event shards replicate ONE ensemble
  g.swap_stream = h->cfg.gibbs_shard_events ? 0u : static_cast<uint32_t>(h->cfg.shard_rank);
  g.prior_z = h->cfg.prior_z;
  g.width_z = h->cfg.prior_width_z;
  g.width_xy = h->cfg.prior_width_xy;
  g.step_xy = h->cfg.step_size_xy;
  g.step_z = h->cfg.step_size_z;
  g.solve[0] = h->cfg.solve_vs;
  g.solve[1] = h->cfg.solve_t_corr;
  g.solve[2] = h->cfg.solve_qs;
  g.solve[3] = h->cfg.solve_a_corr;
  const double pr[4] = {h->cfg.prior_vs, h->cfg.prior_t_corr, h->cfg.prior_qs, h->cfg.prior_a_corr};
  const double wd[4] = {h->cfg.prior_width_vs, h->cfg.prior_width_t_corr, h->cfg.prior_width_qs, h->cfg.prior_width_a_corr};
  const double st[4] = {h->cfg.step_size_vs, h->cfg.step_size_t_corr, h->cfg.step_size_qs, h->cfg.step_size_a_corr};
  for (int t = 0; t < 4; ++t) {
    g.g_prior[t] = pr[t];
    g.g_width[t] = wd[t];
    g.g_step[t] = st[t];
  }
  g.counts = h->d_counts;
  g.rec_origin = h->rec_origin;
  g.rec_cap = h->rec_cap;
  // event-sharded joint chains (or HTM_GIBBS_FORCE_ALLREDUCE with a communicator, for single-GPU tests)
  const bool ev_sharded = h->cfg.gibbs_shard_events && h->cfg.shard_count > 1;
  g.comm = (ev_sharded || (h->comm && std::getenv("HTM_GIBBS_FORCE_ALLREDUCE"))) ? h->comm : nullptr;
  g.count_globals = (!ev_sharded || h->cfg.shard_rank == 0) ? 1 : 0;
  g.xch = PeerExchange();
  const char* force = std::getenv("HTM_GIBBS_EXCHANGE");  // "nccl": keep the all-reduce although peer memory is mapped
  if (ev_sharded && h->xch_on && !(force && std::string(force) == "nccl" && h->comm)) {
    for (int r = 0; r < h->cfg.shard_count; ++r) g.xch.peer[r] = static_cast<double*>(h->xch_peer[r]);
    g.xch.n = h->cfg.shard_count;
    g.xch.rank = h->cfg.shard_rank;
    g.xch.status = h->d_xch_status;
    if (const char* ts = std::getenv("HTM_XCH_TIMEOUT_S")) {  // wall-clock budget of one wait (default 120 s)
      const double sec = std::atof(ts);
      if (sec > 0.0) g.xch.timeout_ns = static_cast<unsigned long long>(sec * 1e9);
    }
  }
  g.xch_epoch0 = h->xch_epoch;
}

// recorded iterations (mod(it, n_interval) == 1) inside [first, last]: ids m = (it-1)/n_interval
bool record_ids(int first, int last, int n_interval, int* m_lo, int* m_hi) {
  if (n_interval <= 1) return false;  // mod(i,1) == 1 is never true (reference quirk Q6)
  const int lo = (first - 1 + n_interval - 1) / n_interval;
  const int hi = (last - 1) / n_interval;
  if (hi < lo) return false;
  *m_lo = lo;
  *m_hi = hi;
  return true;
}

// Queue, behind the kernel(s) just launched on h->stream, the device-to-host copy of ring slots [first, end)
// into the pinned mirror, on the copy stream: it overlaps whatever htm_run queues next.
int32_t queue_sample_copy(htm_handle h, int first, int end) {
  if (end <= first) return HTM_OK;
  HTM_CK(h, cudaEventRecord(h->ev_run, h->stream));
  HTM_CK(h, cudaStreamWaitEvent(h->cstream, h->ev_run, 0));
  auto cp = [&](void* dst, const void* src, size_t per_slot) -> cudaError_t {
    return cudaMemcpyAsync(static_cast<char*>(dst) + first * per_slot, static_cast<const char*>(src) + first * per_slot,
                           (end - first) * per_slot, cudaMemcpyDeviceToHost, h->cstream);
  };
  if (h->cfg.mode == HTM_MODE_BLOCKED_GIBBS) {
    const size_t nc = h->n_cold_total, E = h->E, S = h->S;
    HTM_CK(h, cp(h->host_hypo_rec, h->gl.hypo_rec, nc * E * 4 * h->rs));
    HTM_CK(h, cp(h->host_rec_chain, h->gl.rec_chain, nc * 4));
    HTM_CK(h, cp(h->host_rec_vs, h->gl.rec_vs, nc * 8));
    HTM_CK(h, cp(h->host_rec_qs, h->gl.rec_qs, nc * 8));
    HTM_CK(h, cp(h->host_rec_L, h->gl.rec_L, nc * 8));
    HTM_CK(h, cp(h->host_rec_tc, h->gl.rec_tc, nc * S * 8));
    HTM_CK(h, cp(h->host_rec_ac, h->gl.rec_ac, nc * S * 8));
  } else {
    HTM_CK(h, cp(h->pin_samples, h->d_samples, static_cast<size_t>(h->R) * h->cfg.n_cool * h->E * 4 * h->rs));
  }
  HTM_CK(h, cudaEventRecord(h->ev_copy, h->cstream));
  return HTM_OK;
}

// cfg.summary: the post-burn-in records among ring slots [first, end) (a suffix: slots are in iteration order) go
// into the device-side store, behind the kernel that wrote them, on the same stream
int32_t append_to_store(htm_handle h, int first, int end, int origin) {
  if (!h->cfg.summary || end <= first) return HTM_OK;
  int s0 = first;
  while (s0 < end && (origin + s0) * h->cfg.n_interval + 1 <= h->cfg.n_burn) ++s0;
  const int per_slot = h->cfg.mode == HTM_MODE_BLOCKED_GIBBS ? h->n_cold_total : h->R * h->cfg.n_cool;
  const int n_new = (end - s0) * per_slot;
  if (n_new <= 0) return HTM_OK;
  if (h->store_n + n_new > h->store_cap)
    return fail(h, HTM_ERR_STATE, "summary store full: more post-burn-in records than n_iter / n_burn / n_interval announce");
  const void* ring = h->cfg.mode == HTM_MODE_BLOCKED_GIBBS ? h->gl.hypo_rec : h->d_samples;
  HTM_CK(h, launch_store_append_hypo(h->cfg.precision, ring, s0 * per_slot, n_new, h->E, h->d_store_hypo, h->store_cap,
                                     h->store_n, h->stream));
  if (h->cfg.mode == HTM_MODE_BLOCKED_GIBBS)
    HTM_CK(h, launch_store_append_shared(h->gl.rec_vs, h->gl.rec_qs, h->gl.rec_tc, h->gl.rec_ac, s0 * per_slot, n_new, h->S,
                                         h->d_store_shared, h->store_cap, h->store_n, h->stream));
  h->store_n += n_new;
  return HTM_OK;
}

// the pending records are in the pinned mirror once the last queued copy has completed
int32_t pull_samples(htm_handle h) {
  if (h->host_samples_valid) return HTM_OK;
  HTM_CK(h, cudaEventSynchronize(h->ev_copy));
  if (const int32_t rcx = check_exchange(h)) return rcx;
  h->host_samples_valid = true;
  return HTM_OK;
}
int32_t pull_samples_gibbs(htm_handle h) { return pull_samples(h); }
inline double sample_at(htm_handle h, int rec, int rank, int m, int e, int comp) {
  const size_t i = (((static_cast<size_t>(rec) * h->R + rank) * h->cfg.n_cool + m) * h->E + e) * 4 + comp;
  return h->rs == 8 ? reinterpret_cast<const double*>(h->pin_samples)[i]
                    : static_cast<double>(reinterpret_cast<const float*>(h->pin_samples)[i]);
}

}  // namespace

extern "C" {

int32_t htm_config_default(htm_config* c) {
  if (!c) return HTM_ERR_ARG;
  std::memset(c, 0, sizeof(*c));
  c->abi_version = HTM_ABI_VERSION;
  c->seed = 20231001ull;
  c->temp_high = 200.0;  // sample/hypo_tremor.in:148
  c->prior_z = 0.0;
  c->prior_width_z = 10.0;
  c->prior_width_xy = 30.0;
  c->prior_vs = 3.0;
  c->prior_width_vs = 1.0;
  c->prior_qs = 250.0;
  c->prior_width_qs = 100.0;
  c->prior_t_corr = 0.0;
  c->prior_width_t_corr = 0.5;
  c->prior_a_corr = 0.0;
  c->prior_width_a_corr = 0.02;
  c->step_size_z = 0.4;
  c->step_size_xy = 2.0;
  c->step_size_vs = 0.2;
  c->step_size_qs = 5.0;
  c->step_size_t_corr = 0.03;
  c->step_size_a_corr = 0.005;
  c->hist_xy_halfwidth = 100.0;
  c->hist_z_max = 60.0;
  c->n_procs = 1;
  c->n_chains = 5;
  c->n_cool = 1;
  c->n_iter = 4000000;
  c->n_burn = 2000000;
  c->n_interval = 1000;
  c->solve_vs = c->solve_t_corr = c->solve_qs = c->solve_a_corr = 1;
  c->use_time = c->use_amp = 1;
  c->mode = HTM_MODE_BLOCKED_GIBBS;
  c->precision = HTM_PRECISION_F32;
  c->ladder = HTM_LADDER_RANDOM;
  c->kernel = HTM_KERNEL_AUTO;
  c->shard_count = 1;
  return HTM_OK;
}

int32_t htm_create(htm_handle* out, const htm_config* cfg) {
  if (!out || !cfg) return fail(nullptr, HTM_ERR_ARG, "null argument");
  *out = nullptr;
  if (cfg->abi_version != HTM_ABI_VERSION) return fail(nullptr, HTM_ERR_ARG, "abi_version mismatch");
  if (cfg->n_sta < 1 || cfg->n_events < 1) return fail(nullptr, HTM_ERR_ARG, "n_sta and n_events must be >= 1");
  if (cfg->n_procs < 1 || cfg->n_chains < 1) return fail(nullptr, HTM_ERR_ARG, "n_procs and n_chains must be >= 1");
  if (cfg->n_cool < 1 || cfg->n_cool > cfg->n_chains)
    return fail(nullptr, HTM_ERR_ARG, "n_cool must be in 1..n_chains");
  if (cfg->n_interval < 1) return fail(nullptr, HTM_ERR_ARG, "n_interval must be >= 1");
  if (cfg->precision != HTM_PRECISION_F64 && cfg->precision != HTM_PRECISION_F32)
    return fail(nullptr, HTM_ERR_ARG, "precision must be 32 or 64");
  if (cfg->shard_count < 1 || cfg->shard_rank < 0 || cfg->shard_rank >= cfg->shard_count)
    return fail(nullptr, HTM_ERR_ARG, "bad shard_rank / shard_count");
  if (cfg->temp_high < 1.0) return fail(nullptr, HTM_ERR_ARG, "temp_high must be >= 1");
  const bool any_solve = cfg->solve_vs || cfg->solve_t_corr || cfg->solve_qs || cfg->solve_a_corr;
  if (cfg->mode == HTM_MODE_FACTORISED && any_solve)
    return fail(nullptr, HTM_ERR_ARG,
                "factorised mode needs solve_vs = solve_t_corr = solve_qs = solve_a_corr = F "
                "(the posterior only factorises over events when the shared parameters are fixed)");
  if (cfg->mode == HTM_MODE_REPLAY && cfg->precision != HTM_PRECISION_F64)
    return fail(nullptr, HTM_ERR_ARG, "replay mode is float64 only");
  if (cfg->mode == HTM_MODE_REPLAY && cfg->shard_count != 1)
    return fail(nullptr, HTM_ERR_ARG, "replay mode does not shard (replicas only)");
  if (cfg->mode != HTM_MODE_REPLAY && cfg->mode != HTM_MODE_FACTORISED && cfg->mode != HTM_MODE_BLOCKED_GIBBS)
    return fail(nullptr, HTM_ERR_ARG, "unknown mode");
  if (cfg->hist_bins < 0 || cfg->max_samples < 0) return fail(nullptr, HTM_ERR_ARG, "negative hist_bins/max_samples");
  if (cfg->summary && (cfg->max_samples < 1 || cfg->mode == HTM_MODE_REPLAY))
    return fail(nullptr, HTM_ERR_ARG, "summary = 1 needs max_samples > 0 (the store is fed from the sample ring) and mode B or C");
  if (cfg->mode == HTM_MODE_BLOCKED_GIBBS && !cfg->gibbs_shard_events && cfg->shard_count > cfg->n_procs)
    return fail(nullptr, HTM_ERR_ARG, "blocked-Gibbs mode shards the virtual ranks: shard_count must be <= n_procs");
  if (cfg->gibbs_shard_events && cfg->mode != HTM_MODE_BLOCKED_GIBBS)
    return fail(nullptr, HTM_ERR_ARG, "gibbs_shard_events applies to the blocked-Gibbs mode only");

  int n_dev = 0;
  cudaError_t ce = cudaGetDeviceCount(&n_dev);
  if (ce != cudaSuccess || n_dev == 0)
    return fail(nullptr, HTM_ERR_CUDA,
                std::string("no CUDA device (libhtm_b200 has no CPU fallback): ") +
                    (ce != cudaSuccess ? cudaGetErrorString(ce) : "device count is 0"));
  if (cfg->device < 0 || cfg->device >= n_dev) return fail(nullptr, HTM_ERR_ARG, "device ordinal out of range");
  ce = cudaSetDevice(cfg->device);
  if (ce != cudaSuccess) return fail(nullptr, HTM_ERR_CUDA, cudaGetErrorString(ce));

  htm_handle h = new htm_handle_s();
  h->cfg = *cfg;
  h->E_total = cfg->n_events;
  int lo, hi;
  if (cfg->mode == HTM_MODE_BLOCKED_GIBBS && !cfg->gibbs_shard_events) {
    // joint chains couple all events, so this mode shards the VIRTUAL RANKS by default (the reference's own
    // decomposition, src/hypo_tremor_mcmc.f90:114-118): every shard holds all events and an independent
    // ensemble of its ranks' chains; swaps stay inside the shard; no collective.
    shard_bounds(cfg->n_procs, cfg->shard_rank, cfg->shard_count, &lo, &hi);
    h->rank_off = lo;
    h->ev_off = 0;
    h->E = cfg->n_events;
    h->R = hi - lo;
  } else {
    shard_bounds(cfg->n_events, cfg->shard_rank, cfg->shard_count, &lo, &hi);
    h->ev_off = lo;
    h->E = hi - lo;
    h->R = cfg->n_procs;
  }
  h->S = cfg->n_sta;
  h->K = cfg->n_chains;
  h->C = h->R * h->K;
  h->rs = cfg->precision == HTM_PRECISION_F64 ? 8 : 4;
  h->g_vs = cfg->prior_vs;
  h->g_qs = cfg->prior_qs;
  h->g_tc.assign(h->S, cfg->prior_t_corr);
  h->g_ac.assign(h->S, cfg->prior_a_corr);
  if (h->E < 1 || h->R < 1) {
    delete h;
    return fail(nullptr, HTM_ERR_ARG, "this shard holds no events / no virtual ranks");
  }
  ce = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking);
  if (ce == cudaSuccess) ce = cudaStreamCreateWithFlags(&h->cstream, cudaStreamNonBlocking);
  if (ce == cudaSuccess) ce = cudaEventCreate(&h->ev0);
  if (ce == cudaSuccess) ce = cudaEventCreate(&h->ev1);
  if (ce == cudaSuccess) ce = cudaEventCreateWithFlags(&h->ev_in, cudaEventDisableTiming);
  if (ce == cudaSuccess) ce = cudaEventCreateWithFlags(&h->ev_run, cudaEventDisableTiming);
  if (ce == cudaSuccess) ce = cudaEventCreateWithFlags(&h->ev_copy, cudaEventDisableTiming);
  if (ce != cudaSuccess) {
    g_create_error = cudaGetErrorString(ce);
    delete h;
    return HTM_ERR_CUDA;
  }
  const int32_t rc = alloc_state(h);
  if (rc != HTM_OK) {
    g_create_error = h->err;
    htm_destroy(h);
    return rc;
  }
  *out = h;
  return HTM_OK;
}

int32_t htm_destroy(htm_handle h) {
  if (!h) return HTM_OK;
  cudaSetDevice(h->cfg.device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  if (h->cstream) cudaStreamSynchronize(h->cstream);
  free_dev(h->d_in);
  free_dev(h->d_store_hypo);
  free_dev(h->d_store_shared);
  host_free(h->pin_in, h->pin_in_pinned);
  host_free(h->pin_samples, h->pin_samples_pinned);
  for (cudaEvent_t ev : {h->ev_in, h->ev_run, h->ev_copy})
    if (ev) cudaEventDestroy(ev);
  if (h->cstream) cudaStreamDestroy(h->cstream);
  for (void* p : {h->d_sta4, h->d_obs4, h->d_obs4_raw, h->d_evc4, h->d_prior_xy, static_cast<void*>(h->d_prior_xy64),
                  h->d_x, h->d_y, h->d_z, h->d_L, h->d_T, static_cast<void*>(h->d_counts),
                  static_cast<void*>(h->d_hist), h->d_samples, static_cast<void*>(h->d_hypo),
                  static_cast<void*>(h->d_tc), static_cast<void*>(h->d_ac), static_cast<void*>(h->d_vs),
                  static_cast<void*>(h->d_qs), static_cast<void*>(h->d_temp), static_cast<void*>(h->d_Lc),
                  static_cast<void*>(h->d_chain_counts), static_cast<void*>(h->d_cursor),
                  static_cast<void*>(h->d_status)})
    free_dev(p);
  for (void* p : h->gibbs_bufs) free_dev(p);
  free_dev(h->d_obsx);
  nccl_destroy(h->comm);
  for (int r = 0; r < kMaxPeers; ++r)
    if (h->xch_peer[r] && h->xch_peer[r] != static_cast<void*>(h->xch_buf)) cudaIpcCloseMemHandle(h->xch_peer[r]);
  free_dev(h->xch_buf);
  free_dev(h->d_xch_status);
  if (h->ev0) cudaEventDestroy(h->ev0);
  if (h->ev1) cudaEventDestroy(h->ev1);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
  return HTM_OK;
}

int32_t htm_last_error(htm_handle h, char* buf, int32_t len) {
  if (!buf || len <= 0) return HTM_ERR_ARG;
  const std::string& s = h ? h->err : g_create_error;
  std::snprintf(buf, static_cast<size_t>(len), "%s", s.c_str());
  return HTM_OK;
}

int32_t htm_set_stations(htm_handle h, const double* sx, const double* sy, const double* sz) {
  if (!h || !sx || !sy || !sz) return fail(h, HTM_ERR_ARG, "null argument");
  const double* src[3] = {sx, sy, sz};
  const size_t len[3] = {static_cast<size_t>(h->S), static_cast<size_t>(h->S), static_cast<size_t>(h->S)};
  const int32_t rc = stage_input(h, h->off_sta(), src, len, 3);
  if (rc == HTM_OK) h->have_sta = true;
  return rc;
}

int32_t htm_set_observations(htm_handle h, const double* t_obs, const double* t_stdv, const double* a_obs,
                             const double* a_stdv) {
  if (!h || !t_obs || !t_stdv || !a_obs || !a_stdv) return fail(h, HTM_ERR_ARG, "null argument");
  const size_t n = static_cast<size_t>(h->E) * h->S;
  const double* src[4] = {t_obs, t_stdv, a_obs, a_stdv};
  const size_t len[4] = {n, n, n, n};
  const int32_t rc = stage_input(h, h->off_obs(), src, len, 4);
  if (rc == HTM_OK) h->have_obs = true;
  return rc;
}

int32_t htm_set_xy_prior(htm_handle h, const double* x_mu, const double* y_mu) {
  if (!h || !x_mu || !y_mu) return fail(h, HTM_ERR_ARG, "null argument");
  const double* src[2] = {x_mu, y_mu};
  const size_t len[2] = {static_cast<size_t>(h->E), static_cast<size_t>(h->E)};
  const int32_t rc = stage_input(h, h->off_xy(), src, len, 2);
  if (rc == HTM_OK) h->have_prior = true;
  return rc;
}

int32_t htm_set_globals(htm_handle h, double vs, double qs, const double* t_corr, const double* a_corr) {
  if (!h) return HTM_ERR_ARG;
  if (!(vs > 0.0) || !(qs > 0.0)) return fail(h, HTM_ERR_ARG, "vs and qs must be positive");
  h->g_vs = vs;
  h->g_qs = qs;
  if (t_corr) h->g_tc.assign(t_corr, t_corr + h->S);
  if (a_corr) h->g_ac.assign(a_corr, a_corr + h->S);
  h->tables_ok = false;
  return HTM_OK;
}

int32_t htm_init_chains(htm_handle h) {
  if (!h) return HTM_ERR_ARG;
  HTM_CK(h, cudaSetDevice(h->cfg.device));
  if (h->cfg.mode == HTM_MODE_REPLAY)
    return fail(h, HTM_ERR_UNSUPPORTED,
                "replay mode: chains are initialised by the host's mod_random (htm_set_chain_state)");
  if (!h->have_prior) return fail(h, HTM_ERR_STATE, "xy prior not set (htm_set_xy_prior)");
  int32_t rc = ensure_tables(h);
  if (rc != HTM_OK) return rc;
  if (h->cfg.mode == HTM_MODE_BLOCKED_GIBBS) {
    gibbs_launch_of(h);
    {
      const cudaError_t ei = launch_gibbs_init(h->gl, h->cfg.temp_high, h->cfg.ladder, h->cfg.n_cool, h->stream);
      if (ei == cudaErrorInvalidConfiguration || ei == cudaErrorCooperativeLaunchTooLarge)
        return fail(h, HTM_ERR_UNSUPPORTED,
                    "blocked-Gibbs mode (float32) stages event rows and chain-level state in shared memory: n_sta is "
                    "too large (limit about 180) or there are more than ~9000 joint chains per GPU");
      HTM_CK(h, ei);
    }
    HTM_CK(h, cudaMemsetAsync(h->gl.a_prev, 0, h->C * sizeof(int), h->stream));
    HTM_CK(h, cudaMemsetAsync(h->d_counts, 0, 14 * sizeof(unsigned long long), h->stream));
    h->rec_pending = 0;
    h->store_n = 0;
    h->host_samples_valid = false;
    h->chains_ready = true;
    return HTM_OK;
  }
  h->store_n = 0;
  FactLaunch a = fact_launch_of(h);
  HTM_CK(h, launch_factorised_init(a, h->cfg.temp_high, h->cfg.ladder, h->stream));
  HTM_CK(h, cudaMemsetAsync(h->d_counts, 0, 14 * sizeof(unsigned long long), h->stream));
  if (h->d_hist)
    HTM_CK(h, cudaMemsetAsync(h->d_hist, 0, static_cast<size_t>(h->E) * 3 * h->cfg.hist_bins * sizeof(uint32_t),
                              h->stream));
  h->rec_pending = 0;
  h->host_samples_valid = false;
  h->chains_ready = true;
  return HTM_OK;
}

int32_t htm_set_chain_state(htm_handle h, int32_t rank, int32_t chain, const double* hypo, const double* t_corr,
                            const double* a_corr, double vs, double qs, double temp, double log_likelihood) {
  if (!h || !hypo) return fail(h, HTM_ERR_ARG, "null argument");
  if (rank < 0 || rank >= h->R || chain < 0 || chain >= h->K) return fail(h, HTM_ERR_ARG, "rank/chain out of range");
  HTM_CK(h, cudaSetDevice(h->cfg.device));
  if (h->cfg.mode == HTM_MODE_REPLAY) {
    if (!t_corr || !a_corr) return fail(h, HTM_ERR_ARG, "t_corr / a_corr required in replay mode");
    const size_t c = static_cast<size_t>(rank) * h->K + chain;
    HTM_CK(h, cudaMemcpy(h->d_hypo + c * 3 * h->E, hypo, 3 * h->E * sizeof(double), cudaMemcpyHostToDevice));
    HTM_CK(h, cudaMemcpy(h->d_tc + c * h->S, t_corr, h->S * sizeof(double), cudaMemcpyHostToDevice));
    HTM_CK(h, cudaMemcpy(h->d_ac + c * h->S, a_corr, h->S * sizeof(double), cudaMemcpyHostToDevice));
    HTM_CK(h, cudaMemcpy(h->d_vs + c, &vs, sizeof(double), cudaMemcpyHostToDevice));
    HTM_CK(h, cudaMemcpy(h->d_qs + c, &qs, sizeof(double), cudaMemcpyHostToDevice));
    HTM_CK(h, cudaMemcpy(h->d_temp + c, &temp, sizeof(double), cudaMemcpyHostToDevice));
    HTM_CK(h, cudaMemcpy(h->d_Lc + c, &log_likelihood, sizeof(double), cudaMemcpyHostToDevice));
    h->chains_ready = true;
    return HTM_OK;
  }
  if (h->cfg.mode == HTM_MODE_FACTORISED) {
    // hypocentres and temperature of chain (rank, chain) for every event; the per-event
    // log-likelihoods are recomputed on the device by htm_refresh (next run)
    return fail(h, HTM_ERR_UNSUPPORTED, "htm_set_chain_state: factorised mode initialises with htm_init_chains");
  }
  return fail(h, HTM_ERR_UNSUPPORTED,
              "htm_set_chain_state: blocked-Gibbs mode initialises with htm_init_chains");
}

int32_t htm_get_chain_state(htm_handle h, int32_t rank, int32_t chain, double* hypo, double* t_corr, double* a_corr,
                            double* vs, double* qs, double* temp, double* log_likelihood) {
  if (!h) return HTM_ERR_ARG;
  if (rank < 0 || rank >= h->R || chain < 0 || chain >= h->K) return fail(h, HTM_ERR_ARG, "rank/chain out of range");
  if (!h->chains_ready) return fail(h, HTM_ERR_STATE, "chains not initialised");
  HTM_CK(h, cudaSetDevice(h->cfg.device));
  HTM_CK(h, cudaStreamSynchronize(h->stream));
  if (const int32_t rcx = check_exchange(h)) return rcx;
  if (h->cfg.mode == HTM_MODE_REPLAY) {
    const size_t c = static_cast<size_t>(rank) * h->K + chain;
    if (hypo) HTM_CK(h, cudaMemcpy(hypo, h->d_hypo + c * 3 * h->E, 3 * h->E * sizeof(double), cudaMemcpyDeviceToHost));
    if (t_corr) HTM_CK(h, cudaMemcpy(t_corr, h->d_tc + c * h->S, h->S * sizeof(double), cudaMemcpyDeviceToHost));
    if (a_corr) HTM_CK(h, cudaMemcpy(a_corr, h->d_ac + c * h->S, h->S * sizeof(double), cudaMemcpyDeviceToHost));
    if (vs) HTM_CK(h, cudaMemcpy(vs, h->d_vs + c, sizeof(double), cudaMemcpyDeviceToHost));
    if (qs) HTM_CK(h, cudaMemcpy(qs, h->d_qs + c, sizeof(double), cudaMemcpyDeviceToHost));
    if (temp) HTM_CK(h, cudaMemcpy(temp, h->d_temp + c, sizeof(double), cudaMemcpyDeviceToHost));
    if (log_likelihood) HTM_CK(h, cudaMemcpy(log_likelihood, h->d_Lc + c, sizeof(double), cudaMemcpyDeviceToHost));
    return HTM_OK;
  }
  if (h->cfg.mode == HTM_MODE_FACTORISED) {
    const size_t n = static_cast<size_t>(h->E) * h->R * h->K;
    std::vector<char> bx(n * h->rs), by(n * h->rs), bz(n * h->rs), bL(n * h->rs), bT(n * h->rs);
    HTM_CK(h, cudaMemcpy(bx.data(), h->d_x, n * h->rs, cudaMemcpyDeviceToHost));
    HTM_CK(h, cudaMemcpy(by.data(), h->d_y, n * h->rs, cudaMemcpyDeviceToHost));
    HTM_CK(h, cudaMemcpy(bz.data(), h->d_z, n * h->rs, cudaMemcpyDeviceToHost));
    HTM_CK(h, cudaMemcpy(bL.data(), h->d_L, n * h->rs, cudaMemcpyDeviceToHost));
    HTM_CK(h, cudaMemcpy(bT.data(), h->d_T, n * h->rs, cudaMemcpyDeviceToHost));
    auto at = [&](const std::vector<char>& b, size_t i) -> double {
      return h->rs == 8 ? reinterpret_cast<const double*>(b.data())[i]
                        : static_cast<double>(reinterpret_cast<const float*>(b.data())[i]);
    };
    double ls = 0.0;
    for (int e = 0; e < h->E; ++e) {
      const size_t i = (static_cast<size_t>(e) * h->R + rank) * h->K + chain;
      if (hypo) {
        hypo[3 * e] = at(bx, i);
        hypo[3 * e + 1] = at(by, i);
        hypo[3 * e + 2] = at(bz, i);
      }
      ls += at(bL, i);
    }
    if (t_corr) std::memcpy(t_corr, h->g_tc.data(), h->S * sizeof(double));
    if (a_corr) std::memcpy(a_corr, h->g_ac.data(), h->S * sizeof(double));
    if (vs) *vs = h->g_vs;
    if (qs) *qs = h->g_qs;
    if (temp) *temp = at(bT, (static_cast<size_t>(0) * h->R + rank) * h->K + chain);
    if (log_likelihood) *log_likelihood = ls;
    return HTM_OK;
  }
  // blocked Gibbs: chain c = rank*n_chains + chain
  {
    const size_t c = static_cast<size_t>(rank) * h->K + chain, E = h->E, S = h->S;
    if (hypo && h->rs == 4) {
      std::vector<float> b4(E * 4);
      HTM_CK(h, cudaMemcpy(b4.data(), static_cast<char*>(h->gl.sH) + c * E * 16, E * 16, cudaMemcpyDeviceToHost));
      for (size_t e = 0; e < E; ++e)
        for (int k = 0; k < 3; ++k) hypo[3 * e + k] = static_cast<double>(b4[4 * e + k]);
    } else if (hypo) {
      std::vector<double> bx(E), by(E), bz(E);
      HTM_CK(h, cudaMemcpy(bx.data(), static_cast<char*>(h->gl.hx) + c * E * 8, E * 8, cudaMemcpyDeviceToHost));
      HTM_CK(h, cudaMemcpy(by.data(), static_cast<char*>(h->gl.hy) + c * E * 8, E * 8, cudaMemcpyDeviceToHost));
      HTM_CK(h, cudaMemcpy(bz.data(), static_cast<char*>(h->gl.hz) + c * E * 8, E * 8, cudaMemcpyDeviceToHost));
      for (size_t e = 0; e < E; ++e) {
        hypo[3 * e] = bx[e];
        hypo[3 * e + 1] = by[e];
        hypo[3 * e + 2] = bz[e];
      }
    }
    if (t_corr) HTM_CK(h, cudaMemcpy(t_corr, h->gl.g_tc + c * S, S * 8, cudaMemcpyDeviceToHost));
    if (a_corr) HTM_CK(h, cudaMemcpy(a_corr, h->gl.g_ac + c * S, S * 8, cudaMemcpyDeviceToHost));
    if (vs) HTM_CK(h, cudaMemcpy(vs, h->gl.g_vs + c, 8, cudaMemcpyDeviceToHost));
    if (qs) HTM_CK(h, cudaMemcpy(qs, h->gl.g_qs + c, 8, cudaMemcpyDeviceToHost));
    if (temp) HTM_CK(h, cudaMemcpy(temp, h->gl.g_T + c, 8, cudaMemcpyDeviceToHost));
    if (log_likelihood) HTM_CK(h, cudaMemcpy(log_likelihood, h->gl.g_L + c, 8, cudaMemcpyDeviceToHost));
    return HTM_OK;
  }
}

int32_t htm_loglik(htm_handle h, int32_t n_models, const double* hypo, const double* t_corr, const double* a_corr,
                   const double* vs, const double* qs, double* log_likelihood, double* per_event) {
  if (!h || !hypo || !t_corr || !a_corr || !vs || !qs || !log_likelihood) return fail(h, HTM_ERR_ARG, "null argument");
  if (n_models < 1) return fail(h, HTM_ERR_ARG, "n_models must be >= 1");
  HTM_CK(h, cudaSetDevice(h->cfg.device));
  int32_t rc = ensure_tables(h);
  if (rc != HTM_OK) return rc;
  const size_t M_ = n_models, E = h->E, S = h->S;
  double *d_h = nullptr, *d_tc = nullptr, *d_ac = nullptr, *d_vs = nullptr, *d_qs = nullptr, *d_pe = nullptr,
         *d_L = nullptr;
  auto cleanup = [&]() {
    for (double* p : {d_h, d_tc, d_ac, d_vs, d_qs, d_pe, d_L}) free_dev(p);
  };
  cudaError_t e = cudaSuccess;
  auto up = [&](double** d, const double* src, size_t n) {
    if (e != cudaSuccess) return;
    e = cudaMalloc(d, n * sizeof(double));
    if (e == cudaSuccess && src) e = cudaMemcpyAsync(*d, src, n * sizeof(double), cudaMemcpyHostToDevice, h->stream);
  };
  up(&d_h, hypo, M_ * 3 * E);
  up(&d_tc, t_corr, M_ * S);
  up(&d_ac, a_corr, M_ * S);
  up(&d_vs, vs, M_);
  up(&d_qs, qs, M_);
  up(&d_pe, nullptr, M_ * E);
  up(&d_L, nullptr, M_);
  if (e == cudaSuccess)
    e = launch_loglik(h->cfg.precision, tables_of(h), h->E, h->S, n_models, d_h, d_tc, d_ac, d_vs, d_qs, d_pe, d_L,
                      h->stream);
  if (e == cudaSuccess)
    e = cudaMemcpyAsync(log_likelihood, d_L, M_ * sizeof(double), cudaMemcpyDeviceToHost, h->stream);
  if (e == cudaSuccess && per_event)
    e = cudaMemcpyAsync(per_event, d_pe, M_ * E * sizeof(double), cudaMemcpyDeviceToHost, h->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
  cleanup();
  if (e != cudaSuccess) return fail(h, HTM_ERR_CUDA, std::string("htm_loglik: ") + cudaGetErrorString(e));
  return HTM_OK;
}

static int32_t run_impl(htm_handle h, int32_t iter_first, int32_t iter_last, htm_step_trace* d_trace,
                        htm_swap_trace* d_swaps) {
  if (!h) return HTM_ERR_ARG;
  if (iter_first < 1 || iter_last < iter_first) return fail(h, HTM_ERR_ARG, "need 1 <= iter_first <= iter_last");
  if (h->cfg.mode == HTM_MODE_REPLAY) return fail(h, HTM_ERR_UNSUPPORTED, "replay mode runs through htm_replay");
  if (!h->chains_ready) return fail(h, HTM_ERR_STATE, "chains not initialised (htm_init_chains)");
  HTM_CK(h, cudaSetDevice(h->cfg.device));
  int32_t rc = ensure_tables(h);
  if (rc != HTM_OK) return rc;
  // sample ring bookkeeping: computed here, COMMITTED only after the launch went through (a failed launch must
  // not leave pending records that were never written, nor -- on event shards -- an exchange number that no
  // longer matches the peers')
  int m_lo = 0, m_hi = -1;
  const bool recs = record_ids(iter_first, iter_last, h->cfg.n_interval, &m_lo, &m_hi);
  const bool have_ring = h->cfg.mode == HTM_MODE_BLOCKED_GIBBS ? h->gl.hypo_rec != nullptr : h->d_samples != nullptr;
  int new_origin = h->rec_origin, new_pending = h->rec_pending;
  bool reset_cursors = false;
  if (have_ring && recs) {
    bool all_consumed = true;
    for (int r = 0; r < h->R; ++r)
      if (h->cur_samp[r] < h->rec_pending || h->cur_lik[r] < h->rec_pending) all_consumed = false;
    if (h->rec_pending == 0 || all_consumed) {
      new_origin = m_lo;
      new_pending = 0;
      reset_cursors = true;
    }
    if (m_lo != new_origin + new_pending)
      return fail(h, HTM_ERR_STATE, "iterations must continue where the previous htm_run stopped while samples are pending");
    if (m_hi - new_origin + 1 > h->rec_cap)
      return fail(h, HTM_ERR_STATE,
                  "sample ring full: fetch samples and likelihood of every rank (or htm_discard_samples) "
                  "before running further, or raise max_samples");
    new_pending = m_hi - new_origin + 1;
  }
  // the ring restarts at slot 0: copies of the previous contents must have left the device first
  if (have_ring && recs && reset_cursors) HTM_CK(h, cudaStreamWaitEvent(h->stream, h->ev_copy, 0));
  const int copy_first = (have_ring && recs) ? new_pending - (m_hi - m_lo + 1) : 0;
  auto commit = [&]() {
    if (!(have_ring && recs)) return;
    if (reset_cursors) {
      h->cur_samp.assign(h->R, 0);
      h->cur_lik.assign(h->R, 0);
    }
    h->rec_origin = new_origin;
    h->rec_pending = new_pending;
    h->host_samples_valid = false;
  };
  if (h->cfg.mode == HTM_MODE_BLOCKED_GIBBS) {
    if (h->cfg.gibbs_shard_events && h->cfg.shard_count > 1 && !h->comm && !h->xch_on)
      return fail(h, HTM_ERR_STATE,
                  "event-sharded blocked-Gibbs run needs htm_comm_p2p_export/_import or htm_comm_init first "
                  "(one exchange of the per-chain sums per iteration)");
    const int keep_origin = h->rec_origin;
    h->rec_origin = new_origin;  // the launch description reads the ring origin
    gibbs_launch_of(h);
    h->rec_origin = keep_origin;
    h->gl.iter_first = iter_first;
    h->gl.iter_last = iter_last;
    h->gl.trace = d_trace;
    h->gl.swaps = d_swaps;
    h->gl.out_partials = &h->last_partials;
    h->last_parity = iter_last % 3;
    int nlg = 0;
    HTM_CK(h, cudaEventRecord(h->ev0, h->stream));
    {
      const cudaError_t eg = launch_gibbs(h->gl, h->stream, &nlg);
      if (eg == cudaErrorInvalidConfiguration)
        return fail(h, HTM_ERR_UNSUPPORTED,
                    "blocked-Gibbs mode stages event rows and chain-level state in shared memory: n_sta is too large "
                    "(limit about 180), or -- float64 only -- n_procs*n_chains*n_sta exceeds about 10^4");
      if (eg == cudaErrorCooperativeLaunchTooLarge)
        return fail(h, HTM_ERR_UNSUPPORTED,
                    "the cooperative grid does not fit on the device at once (HTM_GIBBS_PERSIST=1 with too many tiles, "
                    "or more than ~9000 joint chains per GPU in float32)");
      if (eg == cudaErrorNotSupported)
        return fail(h, HTM_ERR_UNSUPPORTED,
                    "float32 event-sharded blocked Gibbs exchanges the per-chain sums through peer memory only: call "
                    "htm_comm_p2p_export / htm_comm_p2p_import (the NCCL all-reduce per iteration serves float64)");
      HTM_CK(h, eg);
    }
    HTM_CK(h, cudaEventRecord(h->ev1, h->stream));
    commit();
    if (have_ring && recs) {
      int32_t rcq = append_to_store(h, copy_first, new_pending, new_origin);
      if (rcq == HTM_OK) rcq = queue_sample_copy(h, copy_first, new_pending);
      if (rcq != HTM_OK) return rcq;
    }
    if (h->gl.xch.n > 1) h->xch_epoch += static_cast<uint32_t>(iter_last - iter_first + 1);
    h->timed = true;
    h->last_launches = nlg;
    h->last_proposals = static_cast<int64_t>(iter_last - iter_first + 1) * (static_cast<int64_t>(h->E) + 1) * h->C;
    return HTM_OK;
  }
  FactLaunch a = fact_launch_of(h);
  a.rec_origin = new_origin;
  a.iter_first = iter_first;
  a.iter_last = iter_last;
  a.trace = d_trace;
  a.swaps = d_swaps;
  int nl = 0;
  const char* why = "";
  HTM_CK(h, cudaEventRecord(h->ev0, h->stream));
  cudaError_t e = launch_factorised(a, h->stream, &nl, &why);
  if (e != cudaSuccess)
    return fail(h, e == cudaErrorInvalidValue && why[0] ? HTM_ERR_UNSUPPORTED : HTM_ERR_CUDA,
                why[0] ? std::string(why) : std::string("launch_factorised: ") + cudaGetErrorString(e));
  HTM_CK(h, cudaEventRecord(h->ev1, h->stream));
  commit();
  if (have_ring && recs) {
    int32_t rcq = append_to_store(h, copy_first, new_pending, new_origin);
    if (rcq == HTM_OK) rcq = queue_sample_copy(h, copy_first, new_pending);
    if (rcq != HTM_OK) return rcq;
  }
  h->timed = true;
  h->last_launches = nl;
  h->last_proposals = static_cast<int64_t>(iter_last - iter_first + 1) * h->E * h->R * h->K;
  return HTM_OK;
}

int32_t htm_run(htm_handle h, int32_t iter_first, int32_t iter_last) {
  return run_impl(h, iter_first, iter_last, nullptr, nullptr);
}

// Validation entry point: htm_run that also returns the per-step and per-swap records
// ([n_it][E][R][K] and [n_it][E][R]) so the kernels can be compared step by step with the
// oracle's statement of the same schedule.  Synchronous.
int32_t htm_run_traced(htm_handle h, int32_t iter_first, int32_t iter_last, htm_step_trace* trace,
                       htm_swap_trace* swaps) {
  if (!h) return HTM_ERR_ARG;
  if (iter_first < 1 || iter_last < iter_first) return fail(h, HTM_ERR_ARG, "need 1 <= iter_first <= iter_last");
  HTM_CK(h, cudaSetDevice(h->cfg.device));
  const size_t n_it = static_cast<size_t>(iter_last - iter_first + 1);
  const bool gibbs = h->cfg.mode == HTM_MODE_BLOCKED_GIBBS;
  // blocked Gibbs: trace [n_it][E+1][J] (row E = shared-parameter step), swaps [n_it]
  const size_t nt = gibbs ? n_it * (static_cast<size_t>(h->E) + 1) * h->C : n_it * h->E * h->R * h->K;
  const size_t ns = gibbs ? n_it : n_it * h->E * h->R;
  htm_step_trace* d_t = nullptr;
  htm_swap_trace* d_s = nullptr;
  cudaError_t ea = cudaSuccess;
  if (trace) {
    ea = cudaMalloc(&d_t, nt * sizeof(htm_step_trace));
    if (ea == cudaSuccess) ea = cudaMemset(d_t, 0, nt * sizeof(htm_step_trace));
  }
  if (ea == cudaSuccess && swaps) {
    ea = cudaMalloc(&d_s, ns * sizeof(htm_swap_trace));
    if (ea == cudaSuccess) ea = cudaMemset(d_s, 0, ns * sizeof(htm_swap_trace));
  }
  if (ea != cudaSuccess) {
    free_dev(d_t);
    free_dev(d_s);
    return fail(h, HTM_ERR_CUDA, std::string("htm_run_traced: ") + cudaGetErrorString(ea));
  }
  int32_t rc = run_impl(h, iter_first, iter_last, d_t, d_s);
  cudaError_t e = cudaSuccess;
  if (rc == HTM_OK) e = cudaStreamSynchronize(h->stream);
  if (rc == HTM_OK && e == cudaSuccess && trace)
    e = cudaMemcpy(trace, d_t, nt * sizeof(htm_step_trace), cudaMemcpyDeviceToHost);
  if (rc == HTM_OK && e == cudaSuccess && swaps)
    e = cudaMemcpy(swaps, d_s, ns * sizeof(htm_swap_trace), cudaMemcpyDeviceToHost);
  free_dev(d_t);
  free_dev(d_s);
  if (rc != HTM_OK) return rc;
  if (e != cudaSuccess) return fail(h, HTM_ERR_CUDA, std::string("htm_run_traced: ") + cudaGetErrorString(e));
  return check_exchange(h);
}

int32_t htm_synchronize(htm_handle h) {
  if (!h) return HTM_ERR_ARG;
  HTM_CK(h, cudaSetDevice(h->cfg.device));
  HTM_CK(h, cudaStreamSynchronize(h->stream));
  return check_exchange(h);
}

int32_t htm_replay(htm_handle h, int32_t iter_first, int32_t iter_last, const int32_t* const* draws,
                   const int64_t* n_draws, htm_step_trace* trace, htm_swap_trace* swaps, int64_t* n_used) {
  if (!h || !draws || !n_draws) return fail(h, HTM_ERR_ARG, "null argument");
  if (h->cfg.mode != HTM_MODE_REPLAY) return fail(h, HTM_ERR_STATE, "handle was not created in replay mode");
  if (iter_first < 1 || iter_last < iter_first) return fail(h, HTM_ERR_ARG, "need 1 <= iter_first <= iter_last");
  if (!h->chains_ready) return fail(h, HTM_ERR_STATE, "chains not initialised (htm_set_chain_state)");
  if (!h->have_prior) return fail(h, HTM_ERR_STATE, "xy prior not set (htm_set_xy_prior)");
  HTM_CK(h, cudaSetDevice(h->cfg.device));
  int32_t rc = ensure_tables(h);
  if (rc != HTM_OK) return rc;
  const int R = h->R;
  std::vector<long long> off(R + 1, 0);
  for (int r = 0; r < R; ++r) {
    if (n_draws[r] < 0 || (n_draws[r] > 0 && !draws[r])) return fail(h, HTM_ERR_ARG, "bad draw stream");
    off[r + 1] = off[r] + n_draws[r];
  }
  const size_t n_it = static_cast<size_t>(iter_last - iter_first + 1);
  int32_t* d_draws = nullptr;
  long long* d_off = nullptr;
  htm_step_trace* d_t = nullptr;
  htm_swap_trace* d_s = nullptr;
  cudaError_t e = cudaMalloc(&d_draws, (off[R] > 0 ? off[R] : 1) * sizeof(int32_t));
  if (e == cudaSuccess) e = cudaMalloc(&d_off, (R + 1) * sizeof(long long));
  for (int r = 0; r < R && e == cudaSuccess; ++r)
    if (n_draws[r] > 0)
      e = cudaMemcpyAsync(d_draws + off[r], draws[r], n_draws[r] * sizeof(int32_t), cudaMemcpyHostToDevice, h->stream);
  if (e == cudaSuccess)
    e = cudaMemcpyAsync(d_off, off.data(), (R + 1) * sizeof(long long), cudaMemcpyHostToDevice, h->stream);
  if (e == cudaSuccess) e = cudaMemsetAsync(h->d_cursor, 0, R * sizeof(long long), h->stream);
  if (e == cudaSuccess && trace) e = cudaMalloc(&d_t, n_it * h->C * sizeof(htm_step_trace));
  if (e == cudaSuccess && swaps) {
    e = cudaMalloc(&d_s, n_it * sizeof(htm_swap_trace));
    if (e == cudaSuccess) e = cudaMemsetAsync(d_s, 0, n_it * sizeof(htm_swap_trace), h->stream);
  }
  int32_t status = 0;
  if (e == cudaSuccess) {
    ReplayLaunch a;
    a.tab = tables_of(h);
    a.E = h->E;
    a.S = h->S;
    a.R = R;
    a.K = h->K;
    a.iter_first = iter_first;
    a.iter_last = iter_last;
    // proposal probabilities, src/cls_mcmc.f90:91-108
    a.p_vs = h->cfg.solve_vs ? 0.025 : 0.0;
    a.p_t_corr = h->cfg.solve_t_corr ? 0.025 : 0.0;
    a.p_qs = h->cfg.solve_qs ? 0.025 : 0.0;
    a.p_a_corr = h->cfg.solve_a_corr ? 0.025 : 0.0;
    a.prior_xy = h->d_prior_xy64;
    a.prior_z = h->cfg.prior_z;
    a.width_z = h->cfg.prior_width_z;
    a.width_xy = h->cfg.prior_width_xy;
    a.step_xy = h->cfg.step_size_xy;
    a.step_z = h->cfg.step_size_z;
    a.prior_vs = h->cfg.prior_vs;
    a.width_vs = h->cfg.prior_width_vs;
    a.step_vs = h->cfg.step_size_vs;
    a.prior_qs = h->cfg.prior_qs;
    a.width_qs = h->cfg.prior_width_qs;
    a.step_qs = h->cfg.step_size_qs;
    a.prior_tc = h->cfg.prior_t_corr;
    a.width_tc = h->cfg.prior_width_t_corr;
    a.step_tc = h->cfg.step_size_t_corr;
    a.prior_ac = h->cfg.prior_a_corr;
    a.width_ac = h->cfg.prior_width_a_corr;
    a.step_ac = h->cfg.step_size_a_corr;
    a.hypo = h->d_hypo;
    a.tc = h->d_tc;
    a.ac = h->d_ac;
    a.vs = h->d_vs;
    a.qs = h->d_qs;
    a.temp = h->d_temp;
    a.L = h->d_Lc;
    a.chain_counts = h->d_chain_counts;
    a.draws = d_draws;
    a.draw_off = reinterpret_cast<const int64_t*>(d_off);
    a.cursor = reinterpret_cast<int64_t*>(h->d_cursor);
    a.trace = d_t;
    a.swaps = d_s;
    a.status = h->d_status;
    cudaEventRecord(h->ev0, h->stream);
    e = launch_replay(a, h->stream);
    cudaEventRecord(h->ev1, h->stream);
    h->timed = true;
    h->last_launches = 1;
    h->last_proposals = static_cast<int64_t>(n_it) * h->C;
  }
  if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
  if (e == cudaSuccess) e = cudaMemcpy(&status, h->d_status, sizeof(int32_t), cudaMemcpyDeviceToHost);
  if (e == cudaSuccess && trace) e = cudaMemcpy(trace, d_t, n_it * h->C * sizeof(htm_step_trace), cudaMemcpyDeviceToHost);
  if (e == cudaSuccess && swaps) e = cudaMemcpy(swaps, d_s, n_it * sizeof(htm_swap_trace), cudaMemcpyDeviceToHost);
  if (e == cudaSuccess && n_used) {
    std::vector<long long> cur(R);
    e = cudaMemcpy(cur.data(), h->d_cursor, R * sizeof(long long), cudaMemcpyDeviceToHost);
    for (int r = 0; r < R; ++r) n_used[r] = cur[r];
  }
  free_dev(d_draws);
  free_dev(d_off);
  free_dev(d_t);
  free_dev(d_s);
  if (e != cudaSuccess) return fail(h, HTM_ERR_CUDA, std::string("htm_replay: ") + cudaGetErrorString(e));
  if (status != 0) return fail(h, HTM_ERR_DRAWS, "htm_replay: the supplied draw stream was exhausted");
  return HTM_OK;
}

int32_t htm_fetch_samples(htm_handle h, int32_t rank, int32_t max_records, int32_t* n_records, int32_t* iter,
                          double* vs, double* qs, double* hypo, double* t_corr, double* a_corr) {
  if (!h || !n_records) return fail(h, HTM_ERR_ARG, "null argument");
  *n_records = 0;
  if (rank < 0 || rank >= h->R) return fail(h, HTM_ERR_ARG, "rank out of range");
  if (h->cfg.mode == HTM_MODE_REPLAY) return fail(h, HTM_ERR_UNSUPPORTED, "samples: not recorded in replay mode");
  HTM_CK(h, cudaSetDevice(h->cfg.device));
  if (h->cfg.mode == HTM_MODE_BLOCKED_GIBBS) {
    // records of the cold chains that sat on virtual rank `rank`, in loop order (iteration, chain)
    if (!h->gl.hypo_rec) return HTM_OK;
    int32_t rcg = pull_samples_gibbs(h);
    if (rcg != HTM_OK) return rcg;
    const int E = h->E, S = h->S, nc = h->n_cold_total;
    int n = 0, rec = h->cur_samp[rank];
    for (; rec < h->rec_pending; ++rec) {
      const int it = (h->rec_origin + rec) * h->cfg.n_interval + 1;
      if (it <= h->cfg.n_burn) continue;
      int here = 0;
      for (int sl = 0; sl < nc; ++sl)
        if (h->host_rec_chain[static_cast<size_t>(rec) * nc + sl] / h->K == rank) ++here;
      if (n + here > max_records) break;
      for (int sl = 0; sl < nc; ++sl) {
        const size_t o = static_cast<size_t>(rec) * nc + sl;
        if (h->host_rec_chain[o] / h->K != rank) continue;
        if (iter) iter[n] = it;
        if (vs) vs[n] = h->host_rec_vs[o];
        if (qs) qs[n] = h->host_rec_qs[o];
        if (hypo)
          for (int e = 0; e < E; ++e)
            for (int cc = 0; cc < 3; ++cc) {
              const size_t i = (o * E + e) * 4 + cc;
              hypo[static_cast<size_t>(n) * 3 * E + 3 * e + cc] =
                  h->rs == 8 ? reinterpret_cast<const double*>(h->host_hypo_rec)[i]
                             : static_cast<double>(reinterpret_cast<const float*>(h->host_hypo_rec)[i]);
            }
        if (t_corr) std::memcpy(t_corr + static_cast<size_t>(n) * S, h->host_rec_tc + o * S, S * sizeof(double));
        if (a_corr) std::memcpy(a_corr + static_cast<size_t>(n) * S, h->host_rec_ac + o * S, S * sizeof(double));
        ++n;
      }
    }
    h->cur_samp[rank] = rec;
    *n_records = n;
    return HTM_OK;
  }
  if (!h->d_samples) return HTM_OK;  // max_samples = 0: nothing is recorded
  int32_t rc = pull_samples(h);
  if (rc != HTM_OK) return rc;
  const int nc = h->cfg.n_cool, E = h->E, S = h->S;
  int n = 0;
  int rec = h->cur_samp[rank];
  for (; rec < h->rec_pending; ++rec) {
    const int it = (h->rec_origin + rec) * h->cfg.n_interval + 1;
    if (it <= h->cfg.n_burn) continue;  // samples start after burn-in, :272
    if (n + nc > max_records) break;
    for (int m = 0; m < nc; ++m, ++n) {
      if (iter) iter[n] = it;
      if (vs) vs[n] = h->g_vs;
      if (qs) qs[n] = h->g_qs;
      if (hypo)
        for (int e = 0; e < E; ++e)
          for (int c = 0; c < 3; ++c) hypo[static_cast<size_t>(n) * 3 * E + 3 * e + c] = sample_at(h, rec, rank, m, e, c);
      if (t_corr) std::memcpy(t_corr + static_cast<size_t>(n) * S, h->g_tc.data(), S * sizeof(double));
      if (a_corr) std::memcpy(a_corr + static_cast<size_t>(n) * S, h->g_ac.data(), S * sizeof(double));
    }
  }
  h->cur_samp[rank] = rec;
  *n_records = n;
  return HTM_OK;
}

int32_t htm_fetch_likelihood(htm_handle h, int32_t rank, int32_t max_records, int32_t* n_records, int32_t* iter,
                             double* log_likelihood) {
  if (!h || !n_records) return fail(h, HTM_ERR_ARG, "null argument");
  *n_records = 0;
  if (rank < 0 || rank >= h->R) return fail(h, HTM_ERR_ARG, "rank out of range");
  if (h->cfg.mode == HTM_MODE_REPLAY) return fail(h, HTM_ERR_UNSUPPORTED, "likelihood: not recorded in replay mode");
  HTM_CK(h, cudaSetDevice(h->cfg.device));
  if (h->cfg.mode == HTM_MODE_BLOCKED_GIBBS) {
    if (!h->gl.hypo_rec) return HTM_OK;
    int32_t rcg = pull_samples_gibbs(h);
    if (rcg != HTM_OK) return rcg;
    const int nc = h->n_cold_total;
    int n = 0, rec = h->cur_lik[rank];
    for (; rec < h->rec_pending; ++rec) {
      const int it = (h->rec_origin + rec) * h->cfg.n_interval + 1;
      int here = 0;
      for (int sl = 0; sl < nc; ++sl)
        if (h->host_rec_chain[static_cast<size_t>(rec) * nc + sl] / h->K == rank) ++here;
      if (n + here > max_records) break;
      for (int sl = 0; sl < nc; ++sl) {
        const size_t o = static_cast<size_t>(rec) * nc + sl;
        if (h->host_rec_chain[o] / h->K != rank) continue;
        if (iter) iter[n] = it;
        if (log_likelihood) log_likelihood[n] = h->host_rec_L[o];
        ++n;
      }
    }
    h->cur_lik[rank] = rec;
    *n_records = n;
    return HTM_OK;
  }
  if (!h->d_samples) return HTM_OK;
  int32_t rc = pull_samples(h);
  if (rc != HTM_OK) return rc;
  const int nc = h->cfg.n_cool, E = h->E;
  int n = 0;
  int rec = h->cur_lik[rank];
  for (; rec < h->rec_pending; ++rec) {
    if (n + nc > max_records) break;
    const int it = (h->rec_origin + rec) * h->cfg.n_interval + 1;
    for (int m = 0; m < nc; ++m, ++n) {
      double s = 0.0;
      for (int e = 0; e < E; ++e) s += sample_at(h, rec, rank, m, e, 3);
      if (iter) iter[n] = it;
      if (log_likelihood) log_likelihood[n] = s;
    }
  }
  h->cur_lik[rank] = rec;
  *n_records = n;
  return HTM_OK;
}

// Drop every pending record (for callers that only want histograms / counts).
int32_t htm_discard_samples(htm_handle h) {
  if (!h) return HTM_ERR_ARG;
  h->rec_pending = 0;
  h->host_samples_valid = false;
  h->cur_samp.assign(h->R, 0);
  h->cur_lik.assign(h->R, 0);
  return HTM_OK;
}

int32_t htm_get_counts(htm_handle h, int64_t n_propose[7], int64_t n_accept[7]) {
  if (!h || !n_propose || !n_accept) return fail(h, HTM_ERR_ARG, "null argument");
  HTM_CK(h, cudaSetDevice(h->cfg.device));
  HTM_CK(h, cudaStreamSynchronize(h->stream));
  if (const int32_t rcx = check_exchange(h)) return rcx;
  for (int k = 0; k < 7; ++k) n_propose[k] = n_accept[k] = 0;
  if (h->cfg.mode == HTM_MODE_REPLAY) {
    std::vector<unsigned long long> c(static_cast<size_t>(h->C) * 14);
    HTM_CK(h, cudaMemcpy(c.data(), h->d_chain_counts, c.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    for (int i = 0; i < h->C; ++i)
      for (int k = 0; k < 7; ++k) {
        n_propose[k] += static_cast<int64_t>(c[static_cast<size_t>(i) * 14 + k]);
        n_accept[k] += static_cast<int64_t>(c[static_cast<size_t>(i) * 14 + 7 + k]);
      }
    return HTM_OK;
  }
  unsigned long long c[14];
  HTM_CK(h, cudaMemcpy(c, h->d_counts, sizeof(c), cudaMemcpyDeviceToHost));
  for (int k = 0; k < 7; ++k) {
    n_propose[k] = static_cast<int64_t>(c[k]);
    n_accept[k] = static_cast<int64_t>(c[7 + k]);
  }
  return HTM_OK;
}

int32_t htm_get_histograms(htm_handle h, uint32_t* hist) {
  if (!h || !hist) return fail(h, HTM_ERR_ARG, "null argument");
  if (!h->d_hist) return fail(h, HTM_ERR_STATE, "histograms are off (hist_bins = 0)");
  HTM_CK(h, cudaSetDevice(h->cfg.device));
  HTM_CK(h, cudaStreamSynchronize(h->stream));
  HTM_CK(h, cudaMemcpy(hist, h->d_hist, static_cast<size_t>(h->E) * 3 * h->cfg.hist_bins * sizeof(uint32_t),
                       cudaMemcpyDeviceToHost));
  return HTM_OK;
}

int32_t htm_comm_unique_id(char id[128]) {
  if (!id) return fail(nullptr, HTM_ERR_ARG, "null argument");
  std::string why;
  if (!nccl_unique_id(id, &why)) return fail(nullptr, HTM_ERR_CUDA, why);
  return HTM_OK;
}

int32_t htm_comm_init(htm_handle h, const char id[128]) {
  if (!h || !id) return fail(h, HTM_ERR_ARG, "null argument");
  if (h->comm) return fail(h, HTM_ERR_STATE, "communicator already initialised");
  HTM_CK(h, cudaSetDevice(h->cfg.device));
  std::string why;
  if (!nccl_init(&h->comm, id, h->cfg.shard_rank, h->cfg.shard_count, &why)) return fail(h, HTM_ERR_CUDA, why);
  return HTM_OK;
}

int32_t htm_comm_p2p_export(htm_handle h, unsigned char handle[64]) {
  if (!h || !handle) return fail(h, HTM_ERR_ARG, "null argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  if (h->cfg.mode != HTM_MODE_BLOCKED_GIBBS || !h->cfg.gibbs_shard_events || h->cfg.shard_count < 2)
    return fail(h, HTM_ERR_STATE, "peer-memory exchange serves event-sharded blocked-Gibbs runs (gibbs_shard_events, shard_count >= 2)");
  if (h->cfg.shard_count > kMaxPeers) return fail(h, HTM_ERR_UNSUPPORTED, "peer-memory exchange supports up to 8 shards");
  HTM_CK(h, cudaSetDevice(h->cfg.device));
  if (!h->xch_buf) {
    const size_t bytes = peer_exchange_bytes(h->cfg.shard_count, h->C);
    HTM_CK(h, cudaMalloc(&h->xch_buf, bytes));
    HTM_CK(h, cudaMemset(h->xch_buf, 0, bytes));
    HTM_CK(h, cudaMalloc(&h->d_xch_status, sizeof(int)));
    HTM_CK(h, cudaMemset(h->d_xch_status, 0, sizeof(int)));
    HTM_CK(h, cudaDeviceSynchronize());  // zeroed before any peer can see the handle
  }
  cudaIpcMemHandle_t ipc;
  HTM_CK(h, cudaIpcGetMemHandle(&ipc, h->xch_buf));
  std::memcpy(handle, &ipc, 64);
  return HTM_OK;
}

int32_t htm_comm_p2p_import(htm_handle h, const unsigned char* handles) {
  if (!h || !handles) return fail(h, HTM_ERR_ARG, "null argument");
  if (!h->xch_buf) return fail(h, HTM_ERR_STATE, "call htm_comm_p2p_export first");
  if (h->xch_on) return fail(h, HTM_ERR_STATE, "peer memory already mapped");
  HTM_CK(h, cudaSetDevice(h->cfg.device));
  for (int r = 0; r < h->cfg.shard_count; ++r) {
    if (r == h->cfg.shard_rank) {
      h->xch_peer[r] = h->xch_buf;
      continue;
    }
    cudaIpcMemHandle_t ipc;
    std::memcpy(&ipc, handles + static_cast<size_t>(64) * r, 64);
    HTM_CK(h, cudaIpcOpenMemHandle(&h->xch_peer[r], ipc, cudaIpcMemLazyEnablePeerAccess));
  }
  h->xch_on = true;
  return HTM_OK;
}

int32_t htm_gather(htm_handle h, uint32_t* hist_all, int64_t n_propose[7], int64_t n_accept[7]) {
  if (!h) return HTM_ERR_ARG;
  if (!h->comm) return fail(h, HTM_ERR_STATE, "htm_comm_init was not called");
  if (h->cfg.mode == HTM_MODE_REPLAY) return fail(h, HTM_ERR_UNSUPPORTED, "replay mode does not shard");
  HTM_CK(h, cudaSetDevice(h->cfg.device));
  const int W = h->cfg.shard_count;
  std::string why;
  // ---- counters: all-reduce (the MPI_Reduce of src/cls_parallel.f90:265-268) ----
  if (n_propose && n_accept) {
    unsigned long long* d_sum = nullptr;
    HTM_CK(h, cudaMalloc(&d_sum, 14 * sizeof(unsigned long long)));
    bool ok = nccl_allreduce_u64(h->comm, h->d_counts, d_sum, 14, h->stream, &why);
    unsigned long long c[14];
    cudaError_t e = ok ? cudaMemcpyAsync(c, d_sum, sizeof(c), cudaMemcpyDeviceToHost, h->stream) : cudaSuccess;
    if (ok && e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    free_dev(d_sum);
    if (!ok) return fail(h, HTM_ERR_CUDA, why);
    HTM_CK(h, e);
    for (int k = 0; k < 7; ++k) {
      n_propose[k] = static_cast<int64_t>(c[k]);
      n_accept[k] = static_cast<int64_t>(c[7 + k]);
    }
  }
  // ---- histograms: all-gather of the event blocks (mode B shards events; blocks padded to the largest) ----
  if (hist_all) {
    if (!h->d_hist) return fail(h, HTM_ERR_STATE, "histograms are off (hist_bins = 0)");
    if (h->cfg.mode != HTM_MODE_FACTORISED) return fail(h, HTM_ERR_UNSUPPORTED, "histograms exist in the factorised mode only");
    const size_t per_ev = static_cast<size_t>(3) * h->cfg.hist_bins;
    int biggest = 0;
    for (int r = 0; r < W; ++r) {
      int lo, hi;
      shard_bounds(h->E_total, r, W, &lo, &hi);
      if (hi - lo > biggest) biggest = hi - lo;
    }
    uint32_t *d_send = nullptr, *d_recv = nullptr;
    HTM_CK(h, cudaMalloc(&d_send, biggest * per_ev * sizeof(uint32_t)));
    HTM_CK(h, cudaMalloc(&d_recv, static_cast<size_t>(W) * biggest * per_ev * sizeof(uint32_t)));
    cudaError_t e = cudaMemsetAsync(d_send, 0, biggest * per_ev * sizeof(uint32_t), h->stream);
    if (e == cudaSuccess)
      e = cudaMemcpyAsync(d_send, h->d_hist, h->E * per_ev * sizeof(uint32_t), cudaMemcpyDeviceToDevice, h->stream);
    bool ok = e == cudaSuccess && nccl_allgather_u32(h->comm, d_send, d_recv, biggest * per_ev, h->stream, &why);
    for (int r = 0; r < W && ok && e == cudaSuccess; ++r) {
      int lo, hi;
      shard_bounds(h->E_total, r, W, &lo, &hi);
      e = cudaMemcpyAsync(hist_all + lo * per_ev, d_recv + static_cast<size_t>(r) * biggest * per_ev,
                          (hi - lo) * per_ev * sizeof(uint32_t), cudaMemcpyDeviceToHost, h->stream);
    }
    if (ok && e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    free_dev(d_send);
    free_dev(d_recv);
    if (!ok && !why.empty()) return fail(h, HTM_ERR_CUDA, why);
    HTM_CK(h, e);
  }
  return check_exchange(h);  // after the collectives, so that no shard is left waiting in them
}

int32_t htm_gather_samples(htm_handle h, int32_t rank, int32_t max_records, int32_t* n_records, int32_t* iter,
                           double* vs, double* qs, double* hypo_all, double* t_corr, double* a_corr) {
  if (!h || !n_records) return fail(h, HTM_ERR_ARG, "null argument");
  if (!h->comm) return fail(h, HTM_ERR_STATE, "htm_comm_init was not called");
  if (h->cfg.mode != HTM_MODE_FACTORISED)
    return fail(h, HTM_ERR_UNSUPPORTED, "sample gather serves the event-sharded factorised mode (the other modes hold every event on every shard)");
  HTM_CK(h, cudaSetDevice(h->cfg.device));
  const int W = h->cfg.shard_count, E = h->E, nc = h->cfg.n_cool;
  int biggest = 0;
  for (int r = 0; r < W; ++r) {
    int lo, hi;
    shard_bounds(h->E_total, r, W, &lo, &hi);
    if (hi - lo > biggest) biggest = hi - lo;
  }
  // Which ring records go out is decided exactly as htm_fetch_samples does (the cursors advance identically on
  // every shard because the recorded iterations are the same everywhere); the scalars come from that call, the
  // hypocentres do NOT pass through the host: ring -> packed [n][3*biggest] doubles (pack_hypo_kernel) ->
  // ncclAllGather over NVLink -> one strided device-to-host copy per shard block into hypo_all.
  const int rec0 = (rank >= 0 && rank < h->R) ? h->cur_samp[rank] : 0;
  int32_t rc = htm_fetch_samples(h, rank, max_records, n_records, iter, vs, qs, nullptr, t_corr, a_corr);
  const int n = rc == HTM_OK ? *n_records : 0;
  // every shard enters the collectives, whatever happened locally: first agree on (failure, record count)
  unsigned long long st_host[2] = {rc == HTM_OK ? 0ull : 1ull, static_cast<unsigned long long>(n)};
  unsigned long long* d_st = nullptr;
  std::string why;
  cudaError_t e = cudaMalloc(&d_st, 2 * sizeof(unsigned long long));
  if (e == cudaSuccess) e = cudaMemcpyAsync(d_st, st_host, sizeof(st_host), cudaMemcpyHostToDevice, h->stream);
  bool ok = e == cudaSuccess && nccl_allreduce_u64(h->comm, d_st, d_st, 2, h->stream, &why);
  if (ok) e = cudaMemcpyAsync(st_host, d_st, sizeof(st_host), cudaMemcpyDeviceToHost, h->stream);
  if (ok && e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
  free_dev(d_st);
  if (!ok && !why.empty()) return fail(h, HTM_ERR_CUDA, why);
  HTM_CK(h, e);
  if (rc != HTM_OK) return rc;
  if (st_host[0] != 0) return fail(h, HTM_ERR_STATE, "htm_gather_samples: another shard failed to fetch its records");
  if (st_host[1] != static_cast<unsigned long long>(n) * W)
    return fail(h, HTM_ERR_STATE, "htm_gather_samples: the shards hold different numbers of records (call with the same arguments everywhere)");
  if (!hypo_all || n == 0) return HTM_OK;
  // ring slots of the n records just consumed: recorded iterations after burn-in, n_cool records each
  std::vector<int> slots;
  for (int rec = rec0; rec < h->cur_samp[rank]; ++rec) {
    const int it = (h->rec_origin + rec) * h->cfg.n_interval + 1;
    if (it <= h->cfg.n_burn) continue;
    for (int m = 0; m < nc; ++m) slots.push_back(rec * nc + m);  // (ring record, cold slot) flattened
  }
  if (static_cast<int>(slots.size()) != n) return fail(h, HTM_ERR_STATE, "htm_gather_samples: record bookkeeping mismatch");
  const size_t blk = static_cast<size_t>(n) * 3 * biggest;  // doubles per shard
  double *d_send = nullptr, *d_recv = nullptr;
  int* d_slots = nullptr;
  auto cleanup = [&]() {
    free_dev(d_send);
    free_dev(d_recv);
    free_dev(d_slots);
  };
  e = cudaMalloc(&d_send, blk * sizeof(double));
  if (e == cudaSuccess) e = cudaMalloc(&d_recv, blk * W * sizeof(double));
  if (e == cudaSuccess) e = cudaMalloc(&d_slots, n * sizeof(int));
  // the flattened (record, cold slot) index addresses rows of E entries when the rank dimension is fixed:
  // ring layout [slot][R][n_cool][E]  ->  row = (slot * R + rank) * n_cool + m
  std::vector<int> rows(n);
  for (int k = 0; k < n; ++k) rows[k] = (slots[k] / nc * h->R + rank) * nc + slots[k] % nc;
  if (e == cudaSuccess) e = cudaMemcpyAsync(d_slots, rows.data(), n * sizeof(int), cudaMemcpyHostToDevice, h->stream);
  if (e == cudaSuccess && biggest != E) e = cudaMemsetAsync(d_send, 0, blk * sizeof(double), h->stream);
  if (e == cudaSuccess)
    e = launch_pack_hypo(h->cfg.precision, h->d_samples, d_slots, n, E, static_cast<size_t>(E), 0, d_send,
                         static_cast<size_t>(3) * biggest, h->stream);
  ok = e == cudaSuccess && nccl_allgather_u32(h->comm, d_send, d_recv, blk * 2, h->stream, &why);
  for (int r = 0; r < W && ok && e == cudaSuccess; ++r) {
    int lo, hi;
    shard_bounds(h->E_total, r, W, &lo, &hi);
    e = cudaMemcpy2DAsync(hypo_all + static_cast<size_t>(3) * lo, static_cast<size_t>(3) * h->E_total * sizeof(double),
                          d_recv + static_cast<size_t>(r) * blk, static_cast<size_t>(3) * biggest * sizeof(double),
                          static_cast<size_t>(3) * (hi - lo) * sizeof(double), n, cudaMemcpyDeviceToHost, h->stream);
  }
  if (ok && e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
  cleanup();
  if (!ok && !why.empty()) return fail(h, HTM_ERR_CUDA, why);
  HTM_CK(h, e);
  return HTM_OK;
}

int32_t htm_posterior_quantiles(htm_handle h, int32_t* n_samples, double* hypo_q, double* vs_q, double* qs_q,
                                double* t_corr_q, double* a_corr_q) {
  if (!h || !n_samples) return fail(h, HTM_ERR_ARG, "null argument");
  if (!h->cfg.summary) return fail(h, HTM_ERR_STATE, "the posterior store is off (cfg.summary = 0)");
  HTM_CK(h, cudaSetDevice(h->cfg.device));
  const int n = static_cast<int>(h->store_n), S = h->S;
  *n_samples = n;
  if (n == 0) return fail(h, HTM_ERR_STATE, "no post-burn-in sample has been recorded yet");
  // src/cls_statistics.f90:230-232: il = 0.025 * n_mod etc. -- default REAL products, truncated; 1-based
  const int il = static_cast<int>(0.025f * static_cast<float>(n)), im = static_cast<int>(0.5f * static_cast<float>(n)),
            iu = static_cast<int>(0.975f * static_cast<float>(n));
  const int r_m = (im > 1 ? im : 1) - 1, r_l = (il > 1 ? il : 1) - 1, r_u = (iu > 1 ? iu : 1) - 1;
  const bool gibbs = h->cfg.mode == HTM_MODE_BLOCKED_GIBBS;
  const int n_h = 3 * h->E, n_s = gibbs ? 2 + 2 * S : 0;
  double* d_out = nullptr;
  HTM_CK(h, cudaMalloc(&d_out, static_cast<size_t>(n_h + n_s) * 3 * sizeof(double)));
  cudaError_t e = launch_quantile_select(h->cfg.precision, h->d_store_hypo, h->store_cap, n_h, n, r_m, r_l, r_u, d_out, h->stream);
  if (e == cudaSuccess && gibbs)
    e = launch_quantile_select(64, h->d_store_shared, h->store_cap, n_s, n, r_m, r_l, r_u, d_out + static_cast<size_t>(n_h) * 3,
                               h->stream);
  std::vector<double> sh(static_cast<size_t>(n_s) * 3);
  if (e == cudaSuccess && hypo_q)
    e = cudaMemcpyAsync(hypo_q, d_out, static_cast<size_t>(n_h) * 3 * sizeof(double), cudaMemcpyDeviceToHost, h->stream);
  if (e == cudaSuccess && gibbs)
    e = cudaMemcpyAsync(sh.data(), d_out + static_cast<size_t>(n_h) * 3, sh.size() * sizeof(double), cudaMemcpyDeviceToHost, h->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
  free_dev(d_out);
  HTM_CK(h, e);
  for (int k = 0; k < 3; ++k) {
    if (vs_q) vs_q[k] = gibbs ? sh[k] : h->g_vs;
    if (qs_q) qs_q[k] = gibbs ? sh[3 + k] : h->g_qs;
    for (int j = 0; j < S; ++j) {
      if (t_corr_q) t_corr_q[3 * j + k] = gibbs ? sh[3 * (2 + j) + k] : h->g_tc[j];
      if (a_corr_q) a_corr_q[3 * j + k] = gibbs ? sh[3 * (2 + S + j) + k] : h->g_ac[j];
    }
  }
  return check_exchange(h);
}

// Validation entry points of the float32 blocked-Gibbs kernel (tests/): the shared-parameter proposal that the
// NEXT iteration will judge, and the sums over this shard's events of the per-event log-likelihoods, current and
// under the proposal, that the LAST iteration judged -- so a test can redo that difference in float64.
int32_t htm_gibbs_pending(htm_handle h, int32_t* which, int32_t* idx, double* x_new) {
  if (!h || !which || !idx || !x_new) return fail(h, HTM_ERR_ARG, "null argument");
  if (h->cfg.mode != HTM_MODE_BLOCKED_GIBBS) return fail(h, HTM_ERR_STATE, "blocked-Gibbs mode only");
  if (!h->chains_ready) return fail(h, HTM_ERR_STATE, "chains not initialised");
  HTM_CK(h, cudaSetDevice(h->cfg.device));
  HTM_CK(h, cudaStreamSynchronize(h->stream));
  static_assert(sizeof(int) == sizeof(int32_t), "int32");
  HTM_CK(h, cudaMemcpy(which, h->gl.prop_which, h->C * sizeof(int), cudaMemcpyDeviceToHost));
  HTM_CK(h, cudaMemcpy(idx, h->gl.prop_idx, h->C * sizeof(int), cudaMemcpyDeviceToHost));
  HTM_CK(h, cudaMemcpy(x_new, h->gl.prop_xnew, h->C * sizeof(double), cudaMemcpyDeviceToHost));
  return check_exchange(h);
}

int32_t htm_gibbs_last_sums(htm_handle h, double* cur, double* prop) {
  if (!h || !cur || !prop) return fail(h, HTM_ERR_ARG, "null argument");
  if (h->cfg.mode != HTM_MODE_BLOCKED_GIBBS || h->cfg.precision != HTM_PRECISION_F32)
    return fail(h, HTM_ERR_UNSUPPORTED, "float32 blocked-Gibbs mode only");
  if (h->last_partials <= 0) return fail(h, HTM_ERR_STATE, "nothing was run yet");
  HTM_CK(h, cudaSetDevice(h->cfg.device));
  HTM_CK(h, cudaStreamSynchronize(h->stream));
  // the accumulator set of the last iteration: [cur, prop][J][2 limbs] fixed-point words (htm_gibbs_f32.cu)
  const size_t J = h->C;
  std::vector<unsigned long long> buf(4 * J);
  HTM_CK(h, cudaMemcpy(buf.data(), reinterpret_cast<const unsigned long long*>(h->gl.part_cur) + static_cast<size_t>(h->last_parity) * 4 * J,
                       4 * J * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
  for (size_t c = 0; c < J; ++c) {
    cur[c] = gibbs_f32_sum_to_double(buf[2 * c], buf[2 * c + 1]);
    prop[c] = gibbs_f32_sum_to_double(buf[2 * (J + c)], buf[2 * (J + c) + 1]);
  }
  return check_exchange(h);
}

int32_t htm_device_ptr(htm_handle h, int32_t what, void** ptr, int64_t* n_bytes) {
  if (!h || !ptr || !n_bytes) return fail(h, HTM_ERR_ARG, "null argument");
  if (what == 0) {
    if (!h->d_hist) return fail(h, HTM_ERR_STATE, "histograms are off (hist_bins = 0)");
    *ptr = h->d_hist;
    *n_bytes = static_cast<int64_t>(h->E) * 3 * h->cfg.hist_bins * sizeof(uint32_t);
    return HTM_OK;
  }
  if (what == 1) {
    *ptr = h->d_counts;
    *n_bytes = 14 * sizeof(unsigned long long);
    return HTM_OK;
  }
  return fail(h, HTM_ERR_ARG, "unknown device buffer id");
}

int32_t htm_last_run_stats(htm_handle h, double* ms, int64_t* n_launches, int64_t* n_proposals) {
  if (!h) return HTM_ERR_ARG;
  if (!h->timed) return fail(h, HTM_ERR_STATE, "nothing was run yet");
  HTM_CK(h, cudaSetDevice(h->cfg.device));
  HTM_CK(h, cudaEventSynchronize(h->ev1));
  float t = 0;
  HTM_CK(h, cudaEventElapsedTime(&t, h->ev0, h->ev1));
  if (ms) *ms = t;
  if (n_launches) *n_launches = h->last_launches;
  if (n_proposals) *n_proposals = h->last_proposals;
  return HTM_OK;
}

int32_t htm_measure_fp32_peak(int32_t device, double* tflops, double* mufu_gops) {
  int n_dev = 0;
  if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0) return fail(nullptr, HTM_ERR_CUDA, "no CUDA device");
  cudaError_t e = measure_fp32_peak(device, tflops, mufu_gops);
  if (e != cudaSuccess) return fail(nullptr, HTM_ERR_CUDA, cudaGetErrorString(e));
  return HTM_OK;
}

int32_t htm_select_events(int32_t device, int32_t n_sta, int32_t n_events, const double* sta_x, const double* sta_y,
                          const double* sta_z, double z_guess, const double* t, const double* t_err, const double* a,
                          const double* a_err, double vs_min, double vs_max, double b_min, double b_max, double* vs,
                          double* t0, double* b, double* a0, double* cc_t, double* cc_a, int32_t* selected,
                          double* kernel_ms) {
  if (!sta_x || !sta_y || !sta_z || !t || !t_err || !a || !a_err || !vs || !t0 || !b || !a0 || !cc_t || !cc_a || !selected)
    return fail(nullptr, HTM_ERR_ARG, "null argument");
  if (n_sta < 3 || n_events < 1) return fail(nullptr, HTM_ERR_ARG, "need n_sta >= 3 and n_events >= 1");
  int n_dev = 0;
  if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0)
    return fail(nullptr, HTM_ERR_CUDA, "no CUDA device (libhtm_b200 has no CPU fallback)");
  if (device < 0 || device >= n_dev) return fail(nullptr, HTM_ERR_ARG, "device ordinal out of range");
  cudaError_t e = cudaSetDevice(device);
  const size_t S = n_sta, ES = static_cast<size_t>(n_events) * n_sta, E = n_events;
  double *d_in = nullptr, *d_out = nullptr;
  int32_t* d_sel = nullptr;
  cudaStream_t st = nullptr;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaEventCreate(&e0);
  if (e == cudaSuccess) e = cudaEventCreate(&e1);
  if (e == cudaSuccess) e = cudaMalloc(&d_in, (3 * S + 4 * ES) * sizeof(double));
  if (e == cudaSuccess) e = cudaMalloc(&d_out, 6 * E * sizeof(double));
  if (e == cudaSuccess) e = cudaMalloc(&d_sel, E * sizeof(int32_t));
  const double* src[7] = {sta_x, sta_y, sta_z, t, t_err, a, a_err};
  size_t off = 0;
  for (int k = 0; k < 7 && e == cudaSuccess; ++k) {
    const size_t n = k < 3 ? S : ES;
    e = cudaMemcpyAsync(d_in + off, src[k], n * sizeof(double), cudaMemcpyHostToDevice, st);
    off += n;
  }
  SelectArgs sa;
  sa.E = n_events;
  sa.S = n_sta;
  sa.sta_x = d_in;
  sa.sta_y = d_in + S;
  sa.sta_z = d_in + 2 * S;
  sa.t = d_in + 3 * S;
  sa.t_err = sa.t + ES;
  sa.a = sa.t_err + ES;
  sa.a_err = sa.a + ES;
  sa.z_guess = z_guess;
  sa.vs_min = vs_min;
  sa.vs_max = vs_max;
  sa.b_min = b_min;
  sa.b_max = b_max;
  sa.vs = d_out;
  sa.t0 = d_out + E;
  sa.b = d_out + 2 * E;
  sa.a0 = d_out + 3 * E;
  sa.cc_t = d_out + 4 * E;
  sa.cc_a = d_out + 5 * E;
  sa.selected = d_sel;
  if (e == cudaSuccess) e = cudaEventRecord(e0, st);
  if (e == cudaSuccess) e = launch_select(sa, st);
  if (e == cudaSuccess) e = cudaEventRecord(e1, st);
  double* dst[6] = {vs, t0, b, a0, cc_t, cc_a};
  for (int k = 0; k < 6 && e == cudaSuccess; ++k)
    e = cudaMemcpyAsync(dst[k], d_out + k * E, E * sizeof(double), cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(selected, d_sel, E * sizeof(int32_t), cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  if (e == cudaSuccess && kernel_ms) {
    float ms = 0;
    e = cudaEventElapsedTime(&ms, e0, e1);
    *kernel_ms = ms;
  }
  free_dev(d_in);
  free_dev(d_out);
  free_dev(d_sel);
  if (e0) cudaEventDestroy(e0);
  if (e1) cudaEventDestroy(e1);
  if (st) cudaStreamDestroy(st);
  if (e != cudaSuccess) return fail(nullptr, HTM_ERR_CUDA, std::string("htm_select_events: ") + cudaGetErrorString(e));
  return HTM_OK;
}

int32_t htm_measure_windows(int32_t device, int32_t n_sta, int64_t n_total, const double* env, double dt, int32_t n_smp,
                            int32_t n_step, int32_t n_win, const int32_t* win_id, double* t, double* t_stdv, double* amp,
                            double* amp_stdv, int32_t* lag, double* kernel_ms) {
  if (!env || !win_id || !t || !t_stdv || !amp || !amp_stdv) return fail(nullptr, HTM_ERR_ARG, "null argument");
  if (n_sta < 3 || n_smp < 2 || n_step < 1 || n_win < 1 || n_total < n_smp || !(dt > 0.0))
    return fail(nullptr, HTM_ERR_ARG, "need n_sta >= 3, n_smp >= 2, n_step >= 1, n_win >= 1, n_total >= n_smp, dt > 0");
  for (int32_t w = 0; w < n_win; ++w) {  // src/cls_measurer.f90:331-333: samples (id - 1) n_step + 1 ... + n_smp
    const int64_t j1 = static_cast<int64_t>(win_id[w] - 1) * n_step;
    if (win_id[w] < 1 || j1 + n_smp > n_total) return fail(nullptr, HTM_ERR_ARG, "window outside the envelopes");
  }
  int n_dev = 0;
  if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0)
    return fail(nullptr, HTM_ERR_CUDA, "no CUDA device (libhtm_b200 has no CPU fallback)");
  if (device < 0 || device >= n_dev) return fail(nullptr, HTM_ERR_ARG, "device ordinal out of range");
  cudaError_t e = cudaSetDevice(device);
  const size_t S = n_sta, W = n_win, P = S * (S - 1) / 2, WS = W * S;
  double *d_env = nullptr, *d_out = nullptr;
  int32_t *d_id = nullptr, *d_lag = nullptr;
  cudaStream_t st = nullptr;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaEventCreate(&e0);
  if (e == cudaSuccess) e = cudaEventCreate(&e1);
  if (e == cudaSuccess) e = cudaMalloc(&d_env, S * static_cast<size_t>(n_total) * sizeof(double));
  if (e == cudaSuccess) e = cudaMalloc(&d_out, 4 * WS * sizeof(double));
  if (e == cudaSuccess) e = cudaMalloc(&d_id, W * sizeof(int32_t));
  if (e == cudaSuccess && lag) e = cudaMalloc(&d_lag, W * P * sizeof(int32_t));
  if (e == cudaSuccess)
    e = cudaMemcpyAsync(d_env, env, S * static_cast<size_t>(n_total) * sizeof(double), cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(d_id, win_id, W * sizeof(int32_t), cudaMemcpyHostToDevice, st);
  MeasureArgs ma;
  ma.S = n_sta;
  ma.n = n_smp;
  ma.n_step = n_step;
  ma.n_win = n_win;
  ma.n_total = static_cast<long>(n_total);
  ma.dt = dt;
  ma.env = d_env;
  ma.win_id = d_id;
  ma.t = d_out;
  ma.t_stdv = d_out + WS;
  ma.amp = d_out + 2 * WS;
  ma.amp_stdv = d_out + 3 * WS;
  ma.lag = d_lag;
  if (e == cudaSuccess) e = cudaEventRecord(e0, st);
  bool unsupported = false;
  if (e == cudaSuccess) {
    e = launch_measure(ma, st);
    unsupported = e == cudaErrorNotSupported;
  }
  if (e == cudaSuccess) e = cudaEventRecord(e1, st);
  double* dst[4] = {t, t_stdv, amp, amp_stdv};
  for (int k = 0; k < 4 && e == cudaSuccess; ++k)
    e = cudaMemcpyAsync(dst[k], d_out + k * WS, WS * sizeof(double), cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess && lag) e = cudaMemcpyAsync(lag, d_lag, W * P * sizeof(int32_t), cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  if (e == cudaSuccess && kernel_ms) {
    float ms = 0;
    e = cudaEventElapsedTime(&ms, e0, e1);
    *kernel_ms = ms;
  }
  free_dev(d_env);
  free_dev(d_out);
  free_dev(d_id);
  free_dev(d_lag);
  if (e0) cudaEventDestroy(e0);
  if (e1) cudaEventDestroy(e1);
  if (st) cudaStreamDestroy(st);
  if (unsupported) {
    cudaGetLastError();
    return fail(nullptr, HTM_ERR_UNSUPPORTED,
                "htm_measure_windows: n_sta x n_smp does not fit one CTA's shared memory (about n_sta (n_smp + 48) 8 B + "
                "n_sta^2 8 B <= 227 KB)");
  }
  if (e != cudaSuccess) return fail(nullptr, HTM_ERR_CUDA, std::string("htm_measure_windows: ") + cudaGetErrorString(e));
  return HTM_OK;
}

int32_t htm_detect_windows(int32_t device, int32_t n_sta, int64_t n_total, const double* env, int32_t n_smp, int32_t n_step,
                           double alpha, int32_t n_pair_thred, int32_t n_win, double* cc_thred, double* cc_max,
                           int32_t* detected, int32_t* n_pairs_above, double* kernel_ms) {
  if (!env || !cc_thred || !detected) return fail(nullptr, HTM_ERR_ARG, "null argument");
  if (n_sta < 2 || n_smp < 2 || (n_smp & 1) || n_step < 1 || n_win < 1 || !(alpha >= 0.0) || !(alpha < 1.0))
    return fail(nullptr, HTM_ERR_ARG, "need n_sta >= 2, an even n_smp >= 2, n_step >= 1, n_win >= 1, 0 <= alpha < 1");
  if (static_cast<int64_t>(n_win - 1) * n_step + n_smp > n_total) return fail(nullptr, HTM_ERR_ARG, "windows outside the envelopes");
  const size_t S = n_sta, W = n_win, P = S * (S - 1) / 2, N = W * static_cast<size_t>(n_smp);
  if (N > 0x7fffffffull) return fail(nullptr, HTM_ERR_UNSUPPORTED, "n_win * n_smp exceeds 2^31");
  // src/cls_measurer.f90:223: cc_thred = sorted(int(n * n_win * alpha)), 1-based
  const long rank1 = static_cast<long>(static_cast<double>(N) * alpha);
  if (rank1 < 1) return fail(nullptr, HTM_ERR_ARG, "int(n_smp * n_win * alpha) must be >= 1");
  int n_dev = 0;
  if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0)
    return fail(nullptr, HTM_ERR_CUDA, "no CUDA device (libhtm_b200 has no CPU fallback)");
  if (device < 0 || device >= n_dev) return fail(nullptr, HTM_ERR_ARG, "device ordinal out of range");
  cudaError_t e = cudaSetDevice(device);
  size_t free_b = 0, total_b = 0;
  if (e == cudaSuccess) e = cudaMemGetInfo(&free_b, &total_b);
  const size_t need = (P * N + P * W + 3 * P + S * static_cast<size_t>(n_total)) * sizeof(double) + 2 * W * sizeof(int32_t);
  if (e == cudaSuccess && need + (static_cast<size_t>(1) << 28) > free_b)
    return fail(nullptr, HTM_ERR_UNSUPPORTED,
                "htm_detect_windows: the correlation functions of all pairs and windows (n_pair * n_win * n_smp * 8 B) do not "
                "fit the device memory; split the time range");
  double *d_env = nullptr, *d_cc = nullptr, *d_max = nullptr, *d_thr = nullptr;
  int32_t *d_det = nullptr, *d_cnt = nullptr;
  cudaStream_t st = nullptr;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaEventCreate(&e0);
  if (e == cudaSuccess) e = cudaEventCreate(&e1);
  if (e == cudaSuccess) e = cudaMalloc(&d_env, S * static_cast<size_t>(n_total) * sizeof(double));
  if (e == cudaSuccess) e = cudaMalloc(&d_cc, P * N * sizeof(double));
  if (e == cudaSuccess) e = cudaMalloc(&d_max, P * W * sizeof(double));
  if (e == cudaSuccess) e = cudaMalloc(&d_thr, 3 * P * sizeof(double));
  if (e == cudaSuccess) e = cudaMalloc(&d_det, W * sizeof(int32_t));
  if (e == cudaSuccess) e = cudaMalloc(&d_cnt, W * sizeof(int32_t));
  if (e == cudaSuccess)
    e = cudaMemcpyAsync(d_env, env, S * static_cast<size_t>(n_total) * sizeof(double), cudaMemcpyHostToDevice, st);
  MeasureArgs ma;
  ma.S = n_sta;
  ma.n = n_smp;
  ma.n_step = n_step;
  ma.n_win = n_win;
  ma.n_total = static_cast<long>(n_total);
  ma.dt = 1.0;
  ma.env = d_env;
  ma.mode = 1;
  ma.cc = d_cc;
  ma.cc_max = d_max;
  if (e == cudaSuccess) e = cudaEventRecord(e0, st);
  bool unsupported = false;
  if (e == cudaSuccess) {
    e = launch_measure(ma, st);
    unsupported = e == cudaErrorNotSupported;
  }
  const int rank0 = static_cast<int>(rank1 - 1);
  if (e == cudaSuccess) e = launch_quantile_select(64, d_cc, N, static_cast<int>(P), static_cast<int>(N), rank0, rank0, rank0, d_thr, st);
  if (e == cudaSuccess) e = launch_detect(d_max, d_thr, static_cast<int>(P), n_win, n_pair_thred, d_det, d_cnt, st);
  if (e == cudaSuccess) e = cudaEventRecord(e1, st);
  if (e == cudaSuccess) e = cudaMemcpy2DAsync(cc_thred, sizeof(double), d_thr, 3 * sizeof(double), sizeof(double), P, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess && cc_max) e = cudaMemcpyAsync(cc_max, d_max, P * W * sizeof(double), cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(detected, d_det, W * sizeof(int32_t), cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess && n_pairs_above) e = cudaMemcpyAsync(n_pairs_above, d_cnt, W * sizeof(int32_t), cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  if (e == cudaSuccess && kernel_ms) {
    float ms = 0;
    e = cudaEventElapsedTime(&ms, e0, e1);
    *kernel_ms = ms;
  }
  free_dev(d_env);
  free_dev(d_cc);
  free_dev(d_max);
  free_dev(d_thr);
  free_dev(d_det);
  free_dev(d_cnt);
  if (e0) cudaEventDestroy(e0);
  if (e1) cudaEventDestroy(e1);
  if (st) cudaStreamDestroy(st);
  if (unsupported) {
    cudaGetLastError();
    return fail(nullptr, HTM_ERR_UNSUPPORTED,
                "htm_detect_windows: n_sta x n_smp does not fit one CTA's shared memory (about n_sta (n_smp + 48) 8 B + "
                "n_sta^2 8 B <= 227 KB)");
  }
  if (e != cudaSuccess) return fail(nullptr, HTM_ERR_CUDA, std::string("htm_detect_windows: ") + cudaGetErrorString(e));
  return HTM_OK;
}

int32_t htm_measure_fp64_peak(int32_t device, double* tflops) {
  int n_dev = 0;
  if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0) return fail(nullptr, HTM_ERR_CUDA, "no CUDA device");
  cudaError_t e = measure_fp64_peak(device, tflops);
  if (e != cudaSuccess) return fail(nullptr, HTM_ERR_CUDA, cudaGetErrorString(e));
  return HTM_OK;
}

}  // extern "C"
