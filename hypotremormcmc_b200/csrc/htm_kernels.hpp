// Host-visible launch interface of the CUDA translation units of libhtm_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

#include "../../include/htm_b200.h"

namespace htm {

// device tables of one shard (see htm_forward.cuh for the layouts); pointers are real4 /
// real2 arrays of the handle's precision
struct Tables {
  const void* sta4 = nullptr;      // [S]
  const void* obs4 = nullptr;      // [E][S], fixed globals folded in (mode B)
  const void* obs4_raw = nullptr;  // [E][S], no station terms folded (loglik, modes A and C)
  const void* evc4 = nullptr;      // [E]
  const void* prior_xy = nullptr;  // [E] real2 {x_mu, y_mu}
};

// device-side table build (htm_tables.cu): raw float64 inputs as the driver hands them over -> the tables above
struct TableBuild {
  int E = 0, S = 0, use_time = 1, use_amp = 1;
  const double* obs_in = nullptr;    // [4][E][S]: t_obs, t_stdv, a_obs, a_stdv (station index fastest)
  const double* sta_xyz = nullptr;   // [3][S]
  const double* xy_mu = nullptr;     // [2][E] or null (no prior centre set: zeros)
  const double* g_tc_ac = nullptr;   // [2][S] fixed station terms folded into obs4 (mode B)
  void *sta4 = nullptr, *obs4 = nullptr, *obs4_raw = nullptr, *evc4 = nullptr, *prior_xy = nullptr;
  double* prior_xy64 = nullptr;      // replay mode only
};
cudaError_t launch_build_tables(int precision, const TableBuild& b, cudaStream_t stream);
// ring records {x, y, z, L_e} -> hypo[record][3E] doubles, device to device (sample gathers)
cudaError_t launch_pack_hypo(int precision, const void* ring, const int* slot_of_rec, int n_rec, int E, size_t ring_stride,
                             size_t row_offset, double* out, size_t out_stride, cudaStream_t stream);

// device-side posterior store and order statistics (htm_summary.cu)
cudaError_t launch_store_append_hypo(int precision, const void* ring, int row0, int n_new, int E, void* store,
                                     size_t cap, size_t pos0, cudaStream_t stream);
cudaError_t launch_store_append_shared(const double* rec_vs, const double* rec_qs, const double* rec_tc, const double* rec_ac,
                                       int row0, int n_new, int S, double* store, size_t cap, size_t pos0,
                                       cudaStream_t stream);
cudaError_t launch_quantile_select(int precision_bits, const void* store, size_t cap, int n_marginals, int n, int r0, int r1,
                                   int r2, double* out, cudaStream_t stream);

// everything a factorised-mode launch needs
struct FactLaunch {
  int precision = 32;
  int kernel = HTM_KERNEL_LANE_PER_CHAIN;
  int slots = 0;  // lane kernel: chains per thread (1, 2 or 4); 0 = choose
  Tables tab;
  void *x = nullptr, *y = nullptr, *z = nullptr, *L = nullptr, *T = nullptr;  // real[E*R*K]
  int E = 0, S = 0, R = 0, K = 0, n_cool = 0;
  int iter_first = 0, iter_last = 0, n_burn = 0, n_interval = 1;
  uint64_t seed = 0;
  uint32_t event_offset = 0;
  double vs = 0, qs = 0;
  double prior_z = 0, width_z = 0, width_xy = 0, step_xy = 0, step_z = 0;
  unsigned long long* counts = nullptr;  // [14]: n_propose[7], n_accept[7]
  void* samples = nullptr;               // real4 [cap][R][n_cool][E] or null
  int rec_origin = 0;                    // record id ((it-1)/n_interval) stored in slot 0
  int rec_cap = 0;
  uint32_t* hist = nullptr;  // [E][3][bins] or null
  int hist_bins = 0;
  double hist_hw = 0, hist_zmax = 0;
  htm_step_trace* trace = nullptr;  // debug: [n_it][E][R][K]
  htm_swap_trace* swaps = nullptr;  // debug: [n_it][E][R]
};

// returns the number of kernels launched through *n_launches
cudaError_t launch_factorised(const FactLaunch& a, cudaStream_t stream, int* n_launches, const char** why);
// generate_model + temperatures + initial log-likelihood of every chain (Philox)
cudaError_t launch_factorised_init(const FactLaunch& a, double temp_high, int ladder, cudaStream_t stream);

// batched full log-likelihood: hypo[M][3E], tc[M][S], ac[M][S], vs[M], qs[M] (device, double)
// -> per_event[M][E] (device, double), L[M] (device, double)
cudaError_t launch_loglik(int precision, const Tables& tab, int E, int S, int M, const double* hypo,
                          const double* tc, const double* ac, const double* vs, const double* qs,
                          double* per_event, double* L, cudaStream_t stream);

// mode A replay (float64)
struct ReplayLaunch {
  Tables tab;  // double tables, obs4_raw
  int E = 0, S = 0, R = 0, K = 0;
  int iter_first = 0, iter_last = 0;
  double p_vs = 0, p_t_corr = 0, p_qs = 0, p_a_corr = 0;
  // per-parameter prior / step tables
  const double* prior_xy = nullptr;  // [E][2]
  double prior_z = 0, width_z = 0, width_xy = 0, step_xy = 0, step_z = 0;
  double prior_vs = 0, width_vs = 0, step_vs = 0, prior_qs = 0, width_qs = 0, step_qs = 0;
  double prior_tc = 0, width_tc = 0, step_tc = 0, prior_ac = 0, width_ac = 0, step_ac = 0;
  // chain state [R*K]...
  double* hypo = nullptr;  // [C][3E]
  double* tc = nullptr;    // [C][S]
  double* ac = nullptr;    // [C][S]
  double *vs = nullptr, *qs = nullptr, *temp = nullptr, *L = nullptr;  // [C]
  unsigned long long* chain_counts = nullptr;                           // [C][14]
  const int32_t* draws = nullptr;   // concatenated per-rank streams
  const int64_t* draw_off = nullptr;  // [R+1] offsets into draws
  int64_t* cursor = nullptr;        // [R] consumed words (in/out)
  htm_step_trace* trace = nullptr;
  htm_swap_trace* swaps = nullptr;
  int32_t* status = nullptr;  // 0 ok, 1 draws exhausted
};
cudaError_t launch_replay(const ReplayLaunch& a, cudaStream_t stream);

// Event-sharded joint chains: all-reduce of the per-chain sums through peer memory (NVLink), inside the last
// CTA of the sweep.  Every shard owns one buffer, mapped into the others with CUDA IPC:
//   double   data [2][n][2*J]   slot [parity][writer shard]; parity = exchange number & 1
//   uint32_t flag [2][n]        exchange number published by `writer shard`
constexpr int kMaxPeers = 8;
struct PeerExchange {
  double* peer[kMaxPeers] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  int n = 0, rank = 0;    // n <= 1: no exchange
  uint32_t epoch = 0;     // exchange number of this launch, the same on every shard, never reused
  int* status = nullptr;  // device flag: 1 = a peer never answered
  unsigned long long timeout_ns = 120ull * 1000000000ull;  // wall-clock budget of ONE wait (globaltimer); the host
                                                           // sets it from HTM_XCH_TIMEOUT_S
};
inline size_t peer_exchange_bytes(int n, int J) {
  // two exchange numbers in flight x n shards x (float64: 2 J doubles; float32: 4 J integer limbs), then the flags
  return static_cast<size_t>(2) * n * 4 * J * sizeof(double) + static_cast<size_t>(2) * n * sizeof(uint32_t);
}

// mode C (blocked Gibbs): joint chains with solved shared parameters
struct GibbsLaunch {
  int precision = 32;
  Tables tab;
  void *hx = nullptr, *hy = nullptr, *hz = nullptr, *hLe = nullptr, *hLp = nullptr;  // real [J][E] (float32: hLp only)
  // float32 state (htm_gibbs_f32.cu), float4 [J][E] each: {x,y,z,L_e}, moments {A1t,A3t,A1a,A3a},
  // sums {S1t,S1a,A2t,A2a}, and the sums under the chain's pending shared-parameter proposal
  void *sH = nullptr, *sM = nullptr, *sQ = nullptr, *sP = nullptr;
  double *g_vs = nullptr, *g_qs = nullptr, *g_tc = nullptr, *g_ac = nullptr, *g_T = nullptr, *g_L = nullptr;
  int *prop_which = nullptr, *prop_idx = nullptr, *a_prev = nullptr, *slot_of = nullptr;
  double *prop_xnew = nullptr, *prop_lpr = nullptr;
  double *part_cur = nullptr, *part_prop = nullptr;  // one allocation of 4 * J * part_tiles doubles
  int part_tiles = 0;                                // >= ceil(E/32)
  unsigned int* done_counter = nullptr;              // CTAs finished in the current sweep
  const void* obsx = nullptr;                        // float32: expanded station-pair rows [E][xrow] float4
  int xrow = 0;
  // event-sharded joint chains: all-reduce of the per-chain sums every iteration
  void* comm = nullptr;       // NCCL communicator (null = single shard)
  PeerExchange xch;           // peer-memory exchange (preferred over the NCCL all-reduce when set up)
  uint32_t xch_epoch0 = 0;    // exchange number of iteration iter_first
  double* totals = nullptr;   // [2][J] scratch
  int count_globals = 1;      // shared-parameter counters are replicated on every shard: only shard 0 counts
  int E = 0, S = 0, J = 0, K = 0, n_cool_total = 0;
  int iter_first = 0, iter_last = 0, n_burn = 0, n_interval = 1;
  uint64_t seed = 0;
  uint32_t event_offset = 0;
  uint32_t chain_offset = 0, J_total = 0, swap_stream = 0;  // shards of virtual ranks (global Philox ids)
  double prior_z = 0, width_z = 0, width_xy = 0, step_xy = 0, step_z = 0;
  int solve[4] = {0, 0, 0, 0};  // vs, t_corr, qs, a_corr
  double g_prior[4] = {0, 0, 0, 0}, g_width[4] = {0, 0, 0, 0}, g_step[4] = {0, 0, 0, 0};
  unsigned long long* counts = nullptr;
  void* hypo_rec = nullptr;  // real4 [cap][n_cool_total][E]
  int* rec_chain = nullptr;  // [cap][n_cool_total]
  double *rec_vs = nullptr, *rec_qs = nullptr, *rec_L = nullptr;  // [cap][n_cool_total]
  double *rec_tc = nullptr, *rec_ac = nullptr;                    // [cap][n_cool_total][S]
  int rec_origin = 0, rec_cap = 0;
  htm_step_trace* trace = nullptr;  // debug: [n_it][E+1][J]
  htm_swap_trace* swaps = nullptr;  // debug: [n_it]
  int* out_partials = nullptr;      // float32: receives the number of partial sums per chain of this launch
};
cudaError_t launch_gibbs(const GibbsLaunch& a, cudaStream_t stream, int* n_launches);
double gibbs_f32_sum_to_double(unsigned long long l0, unsigned long long l1);  // htm_gibbs_f32.cu: limbs_to_double
// float32: build the expanded rows from the raw tables (once per table upload)
cudaError_t launch_expand_obs(const Tables& tab, int E, int S, void* obsx, cudaStream_t stream);
cudaError_t launch_gibbs_init(const GibbsLaunch& a, double temp_high, int ladder, int n_cool, cudaStream_t stream);

// NCCL, loaded with dlopen (htm_comm.cu)
struct NcclId {
  char internal[128];
};
bool nccl_unique_id(char id[128], std::string* why);
bool nccl_init(void** comm, const char id[128], int rank, int nranks, std::string* why);
void nccl_destroy(void* comm);
bool nccl_allgather_u32(void* comm, const void* send, void* recv, size_t count, cudaStream_t s, std::string* why);
bool nccl_allreduce_f64(void* comm, const void* send, void* recv, size_t count, cudaStream_t s, std::string* why);
bool nccl_allreduce_u64(void* comm, const void* send, void* recv, size_t count, cudaStream_t s, std::string* why);

// upstream QC stage hypo_tremor_select, batched (htm_select.cu); all pointers are device memory
struct SelectArgs {
  int E = 0, S = 0;
  const double *sta_x = nullptr, *sta_y = nullptr, *sta_z = nullptr;                   // [S]
  const double *t = nullptr, *t_err = nullptr, *a = nullptr, *a_err = nullptr;        // [E][S]
  double z_guess = 0, vs_min = 0, vs_max = 0, b_min = 0, b_max = 0;
  double *vs = nullptr, *t0 = nullptr, *b = nullptr, *a0 = nullptr, *cc_t = nullptr, *cc_a = nullptr;  // [E]
  int32_t* selected = nullptr;                                                         // [E]
};
cudaError_t launch_select(const SelectArgs& a, cudaStream_t stream);

// upstream stage hypo_tremor_measure, lag / amplitude optimisation of all detected windows (htm_measure.cu); all
// pointers are device memory
struct MeasureArgs {
  int S = 0, n = 0, n_step = 0, n_win = 0;  // stations, samples per window, samples between window starts, windows
  long n_total = 0;                         // samples per station in env
  double dt = 0;
  const double* env = nullptr;              // [S][n_total] merged envelopes
  const int32_t* win_id = nullptr;          // [n_win] 1-based window numbers
  double *t = nullptr, *t_stdv = nullptr, *amp = nullptr, *amp_stdv = nullptr;  // [n_win][S]
  int32_t* lag = nullptr;                   // optional [n_win][S (S - 1) / 2]: sample index of each pair's maximum
  // mode 1: the correlation functions of hypo_tremor_correlate instead (win_id null: windows 1 .. n_win)
  int mode = 0;
  double* cc = nullptr;                     // optional [S (S - 1) / 2][n_win][n]: correlation value at circular lag k
  double* cc_max = nullptr;                 // optional [S (S - 1) / 2][n_win]
};
cudaError_t launch_measure(const MeasureArgs& a, cudaStream_t stream);
cudaError_t launch_detect(const double* cc_max, const double* thr, int n_pair, int n_win, int n_pair_thred, int32_t* detected,
                          int32_t* count, cudaStream_t stream);

// FFMA / MUFU microbenchmark (roofline denominators)
cudaError_t measure_fp32_peak(int device, double* tflops, double* mufu_gops);
cudaError_t measure_fp64_peak(int device, double* tflops);

}  // namespace htm
