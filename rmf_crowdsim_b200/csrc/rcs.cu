// rcs.cu -- host side of the C ABI declared in include/rcs.h.
//
// Owns all device memory of a simulation handle and enqueues the per-step kernel pipeline of
// rcs_kernels.cuh on the handle's stream.  There is no CPU implementation of any compute entry
// point in this file: without a CUDA device every such call fails.
//
// Compile: nvcc -std=c++17 -O3 --fmad=false -gencode arch=compute_100a,code=sm_100a -lineinfo
#include "../../include/rcs.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <limits>
#include <numeric>
#include <string>
#include <vector>

#include "rcs_kernels.cuh"
#include "rcs_step_warp.cuh"

using namespace rcs;

namespace {

thread_local std::string g_create_error;

struct LPDesc {
  uint32_t kind;
  double agent_scale, obstacle_scale, reaction_time, force_distance, agent_mass, agent_radius;
};
struct HLDesc {
  uint32_t kind;
  double vx, vy;
};
struct GroupKey {
  uint32_t hl, lp;
  double eyesight;
  int32_t source_sink;
};

struct PendingStep {
  AgentArrays cur, srt;   // pointer roles when the step was enqueued
  bool snapshot_in_srt;   // true: pre-step state is in srt (sorted path); false: it is in cur (streaming path)
  uint32_t n;
};

}  // namespace

struct rcs_sim {
  rcs_sim_desc desc{};
  GridDev grid{};
  int device = 0;
  cudaStream_t stream = nullptr;
  uint64_t cap = 0;
  uint32_t n = 0;
  AgentArrays cur{}, srt{};
  uint32_t *cellid = nullptr, *perm = nullptr, *cell_count = nullptr, *cell_start = nullptr, *cursor = nullptr;
  uint32_t *tile_sums = nullptr, *scan_total = nullptr, *big_list = nullptr;
  uint64_t tile_sums_cap = 0;
  GroupDev* d_groups = nullptr;
  uint32_t d_groups_cap = 0;
  std::vector<GroupDev> groups;
  std::vector<GroupKey> group_keys;
  bool groups_dirty = false;
  std::vector<LPDesc> lps;
  std::vector<HLDesc> hls;
  bool any_zanlungo = false;
  bool have_host_hl = false;
  DevStatus* d_status = nullptr;
  DevStatus* h_status = nullptr;  // pinned
  uint64_t last_alloc_agent_id = 0;
  uint64_t max_id_plus1 = 0;
  bool index_valid = false;  // srt + cell_start describe the current positions
  // id-addressed access
  uint32_t *slot_of_id = nullptr, *id_rank = nullptr, *order_by_id = nullptr, *presence = nullptr;
  uint64_t slot_table_cap = 0;
  bool slot_valid = false;
  // trace
  bool trace = false;
  double *tr_ti = nullptr, *tr_fx = nullptr, *tr_fy = nullptr;
  uint32_t *tr_nbc = nullptr, *tr_nbo = nullptr;
  uint64_t* tr_nbids = nullptr;
  uint64_t tr_nbids_cap = 0, tr_nb_total = 0;
  uint32_t tr_n = 0;
  bool tr_valid = false;
  // staging
  void* stage = nullptr;
  uint64_t stage_bytes = 0;
  void* flush_buf = nullptr;
  uint64_t flush_bytes = 0;
  unsigned int* d_bad = nullptr;
  std::vector<PendingStep> pending;
  uint64_t steps_enqueued = 0;
  uint64_t* d_steps_done = nullptr;
  uint64_t steps_done_at_sync = 0;
  std::string err;
  uint64_t launches = 0;
  rcs_stats stats{};
  cudaEvent_t events[RCS_NUM_EVENTS]{};
  uint32_t opt_step_kernel = 0;
  // dominant-kernel timing
  bool ktiming = false;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> kevents;  // pending pairs
  std::vector<cudaEvent_t> kevent_pool;
  double ktime_ms = 0.0;
  uint64_t ktime_n = 0;
};

namespace {

#define CU_TRY(sim, call)                                                                         \
  do {                                                                                            \
    cudaError_t e__ = (call);                                                                     \
    if (e__ != cudaSuccess) {                                                                     \
      (sim)->err = std::string("CUDA error: ") + cudaGetErrorString(e__) + " at " #call;          \
      return RCS_ERR_CUDA;                                                                        \
    }                                                                                             \
  } while (0)

inline uint32_t blocks_for(uint64_t n, uint32_t threads) { return (uint32_t)((n + threads - 1) / threads); }

// Rust `f64 as usize` on the host (same rule as rcs_math.cuh)
uint64_t host_f64_as_usize(double v) {
  if (!(v > 0.0)) return 0;
  if (v >= 18446744073709551616.0) return std::numeric_limits<uint64_t>::max();
  return (uint64_t)v;
}

bool host_location_to_index(const GridDev& g, double px, double py, uint64_t& idx) {
  uint64_t x_idx = host_f64_as_usize((px - g.offx) / g.res);
  uint64_t y_idx = host_f64_as_usize((py - g.offy) / g.res);
  idx = x_idx * g.nx + y_idx;
  return idx < g.len;
}

// Smallest double T such that sqrt(T) >= R (correctly rounded sqrt is monotone), so that the
// reference's strict test `norm < radius` (location_hash_2d.rs:251) is exactly `norm_squared < T`.
double radius_threshold(double R) {
  if (R != R) return R;           // NaN: every comparison false, as in the reference
  if (!(R > 0.0)) return 0.0;     // sqrt(d2) >= 0 is never < R
  if (std::isinf(R)) return R;    // d2 < inf  <=>  sqrt(d2) < inf
  double t = R * R;
  if (std::isinf(t)) {            // R*R overflows: walk down from the largest finite double
    t = std::numeric_limits<double>::max();
    if (std::sqrt(t) < R) return std::numeric_limits<double>::infinity();
  }
  while (std::sqrt(t) >= R) t = std::nextafter(t, 0.0);
  while (std::sqrt(t) < R) t = std::nextafter(t, std::numeric_limits<double>::infinity());
  return t;
}

template <class T>
cudaError_t dalloc(T** p, uint64_t count) {
  return cudaMalloc(reinterpret_cast<void**>(p), std::max<uint64_t>(count, 1) * sizeof(T));
}

int alloc_agent_arrays(rcs_sim* s, AgentArrays& a, uint64_t cap) {
  CU_TRY(s, dalloc(&a.x, cap));
  CU_TRY(s, dalloc(&a.y, cap));
  CU_TRY(s, dalloc(&a.vx, cap));
  CU_TRY(s, dalloc(&a.vy, cap));
  CU_TRY(s, dalloc(&a.id, cap));
  CU_TRY(s, dalloc(&a.grp, cap));
  CU_TRY(s, dalloc(&a.wp, cap));
  a.pvx = a.pvy = nullptr;
  return RCS_OK;
}

void free_agent_arrays(AgentArrays& a) {
  cudaFree(a.x); cudaFree(a.y); cudaFree(a.vx); cudaFree(a.vy);
  cudaFree(a.id); cudaFree(a.grp); cudaFree(a.wp); cudaFree(a.pvx); cudaFree(a.pvy);
  a = AgentArrays{};
}

int ensure_stage(rcs_sim* s, uint64_t bytes) {
  if (bytes <= s->stage_bytes) return RCS_OK;
  CU_TRY(s, cudaStreamSynchronize(s->stream));
  if (s->stage) cudaFree(s->stage);
  s->stage = nullptr;
  s->stage_bytes = 0;
  uint64_t want = bytes + bytes / 4 + 4096;
  CU_TRY(s, cudaMalloc(&s->stage, want));
  s->stage_bytes = want;
  return RCS_OK;
}

// exclusive scan of in[0..len) into out[0..len] (len+1 entries), optional cursor copy
int exclusive_scan(rcs_sim* s, const uint32_t* in, uint64_t len, uint32_t* out, uint32_t* cursor) {
  if (len == 0) {
    CU_TRY(s, cudaMemsetAsync(out, 0, sizeof(uint32_t), s->stream));
    return RCS_OK;
  }
  uint64_t tiles = (len + SCAN_TILE - 1) / SCAN_TILE;
  if (tiles > s->tile_sums_cap) {
    CU_TRY(s, cudaStreamSynchronize(s->stream));
    cudaFree(s->tile_sums);
    s->tile_sums = nullptr;
    CU_TRY(s, dalloc(&s->tile_sums, tiles + 1024));
    s->tile_sums_cap = tiles + 1024;
  }
  scan_reduce_kernel<<<(uint32_t)tiles, SCAN_THREADS, 0, s->stream>>>(in, len, s->tile_sums);
  scan_tile_sums_kernel<<<1, SCAN_THREADS, 0, s->stream>>>(s->tile_sums, (uint32_t)tiles, s->scan_total);
  scan_apply_kernel<<<(uint32_t)tiles, SCAN_THREADS, 0, s->stream>>>(in, len, s->tile_sums, out, cursor);
  s->launches += 3;
  CU_TRY(s, cudaGetLastError());
  return RCS_OK;
}

int upload_groups(rcs_sim* s) {
  if (!s->groups_dirty) return RCS_OK;
  if (s->groups.size() > s->d_groups_cap) {
    CU_TRY(s, cudaStreamSynchronize(s->stream));
    cudaFree(s->d_groups);
    s->d_groups = nullptr;
    uint32_t cap = (uint32_t)s->groups.size() * 2 + 16;
    CU_TRY(s, dalloc(&s->d_groups, cap));
    s->d_groups_cap = cap;
  }
  // groups is a host vector that may be reallocated later: synchronous copy
  CU_TRY(s, cudaStreamSynchronize(s->stream));
  CU_TRY(s, cudaMemcpy(s->d_groups, s->groups.data(), s->groups.size() * sizeof(GroupDev), cudaMemcpyHostToDevice));
  s->groups_dirty = false;
  return RCS_OK;
}

uint32_t find_or_add_group(rcs_sim* s, uint32_t hl, uint32_t lp, double eyesight, int32_t source_sink) {
  for (size_t k = 0; k < s->group_keys.size(); ++k) {
    const GroupKey& g = s->group_keys[k];
    if (g.hl == hl && g.lp == lp && g.source_sink == source_sink &&
        std::memcmp(&g.eyesight, &eyesight, sizeof(double)) == 0)
      return (uint32_t)k;
  }
  const LPDesc& L = s->lps[lp];
  const HLDesc& H = s->hls[hl];
  GroupDev g{};
  g.eyesight = eyesight;
  g.thr2 = radius_threshold(eyesight);
  g.hl_vx = H.vx;
  g.hl_vy = H.vy;
  g.hl_kind = H.kind;
  g.lp_kind = L.kind;
  g.source_sink = source_sink;
  if (L.kind == LP_ZANLUNGO) {
    g.agent_scale = L.agent_scale;
    g.force_distance = L.force_distance;
    g.inv_mass = 1.0 / L.agent_mass;
    g.rr = L.agent_radius * L.agent_radius;
    g.two_r = L.agent_radius * 2.0;
    // weight-0 pairs can be skipped only if 0*agent_scale == 0 and exp(-(dist - 2r)/D) cannot overflow
    bool ok = std::isfinite(L.agent_scale) && L.force_distance > 0.0 && std::isfinite(g.two_r) &&
              (g.two_r / L.force_distance) < 700.0;
    g.w0_fast = ok ? 1u : 0u;
    s->any_zanlungo = true;
  }
  s->groups.push_back(g);
  s->group_keys.push_back(GroupKey{hl, lp, eyesight, source_sink});
  s->groups_dirty = true;
  return (uint32_t)(s->groups.size() - 1);
}

int ensure_pref_arrays(rcs_sim* s) {
  if (s->cur.pvx) return RCS_OK;
  CU_TRY(s, cudaStreamSynchronize(s->stream));
  const double nan = std::numeric_limits<double>::quiet_NaN();
  for (AgentArrays* a : {&s->cur, &s->srt}) {
    CU_TRY(s, dalloc(&a->pvx, s->cap));
    CU_TRY(s, dalloc(&a->pvy, s->cap));
    fill_f64_kernel<<<blocks_for(s->cap, 256), 256, 0, s->stream>>>(s->cap, a->pvx, nan);
    fill_f64_kernel<<<blocks_for(s->cap, 256), 256, 0, s->stream>>>(s->cap, a->pvy, nan);
    s->launches += 2;
  }
  CU_TRY(s, cudaGetLastError());
  // pointer roles recorded for pending steps are stale now, but ensure_pref_arrays only runs right after a sync
  return RCS_OK;
}

void invalidate(rcs_sim* s) {
  s->index_valid = false;
  s->slot_valid = false;
  s->tr_valid = false;
}

// (Re)build the canonical (cell, id) sorted copy `srt` of `cur` and cell_start.
int build_index(rcs_sim* s, bool count_oob_as_dead) {
  (void)count_oob_as_dead;
  const uint64_t len = s->grid.len;
  const uint32_t n = s->n;
  CU_TRY(s, cudaMemsetAsync(s->cell_count, 0, (len + 1) * sizeof(uint32_t), s->stream));
  if (n) {
    bin_count_kernel<<<blocks_for(n, 256), 256, 0, s->stream>>>(s->grid, n, s->cur.x, s->cur.y, s->cellid,
                                                                s->cell_count, s->d_status);
    s->launches += 1;
  }
  int rc = exclusive_scan(s, s->cell_count, len, s->cell_start, s->cursor);
  if (rc) return rc;
  if (n) {
    scatter_perm_kernel<<<blocks_for(n, 256), 256, 0, s->stream>>>(n, s->cellid, s->cursor, s->perm, s->d_status);
    if (len)
      sort_cells_by_id_kernel<<<blocks_for(len, 128), 128, 0, s->stream>>>(len, s->cell_start, s->cur.id, s->perm,
                                                                           s->big_list, 4096, s->d_status);
    sort_big_cells_kernel<<<64, 256, 0, s->stream>>>(s->cell_start, s->cur.id, s->perm, s->cellid, s->big_list, 4096,
                                                     s->d_status);
    gather_sorted_kernel<<<blocks_for(n, 256), 256, 0, s->stream>>>(n, s->perm, s->cur, s->srt,
                                                                    s->cell_start + len, s->d_status);
    s->launches += 4;
  }
  CU_TRY(s, cudaGetLastError());
  return RCS_OK;
}

__global__ void begin_step_kernel(DevStatus* st) {
  if (st->failed) return;
  st->oob_count = 0;
  st->nonfinite_count = 0;
  st->big_cells = 0;
  st->first_oob_id = ~0ull;
  st->finite_tti = 0;
  st->neighbour_total = 0;
  st->candidate_total = 0;
}

__global__ void end_step_kernel(DevStatus* st, uint64_t* steps_done, int may_fail) {
  if (st->failed) return;
  if (may_fail && st->oob_count) {
    st->failed = 1;
    return;
  }
  *steps_done += 1;
}

cudaEvent_t kevent_get(rcs_sim* s) {
  if (!s->kevent_pool.empty()) {
    cudaEvent_t e = s->kevent_pool.back();
    s->kevent_pool.pop_back();
    return e;
  }
  cudaEvent_t e = nullptr;
  cudaEventCreate(&e);
  return e;
}

// launch the dominant kernel, optionally bracketed by events on the launching stream
void launch_step_kernel(rcs_sim* s, const StepArgs& a, uint32_t n, bool sorted_input = true) {
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  if (s->ktiming) {
    e0 = kevent_get(s);
    e1 = kevent_get(s);
    cudaEventRecord(e0, s->stream);
  }
  if (sorted_input && s->opt_step_kernel != 1)
    step_warp_kernel<<<blocks_for(n, 32 * SW_WARPS), 32 * SW_WARPS, 0, s->stream>>>(a);
  else
    step_kernel<<<blocks_for(n, 128), 128, 0, s->stream>>>(a);
  s->launches += 1;
  if (s->ktiming) {
    cudaEventRecord(e1, s->stream);
    s->kevents.push_back({e0, e1});
  }
}

int drain_kevents(rcs_sim* s) {
  for (auto& pr : s->kevents) {
    CU_TRY(s, cudaEventSynchronize(pr.second));
    float ms = 0.f;
    CU_TRY(s, cudaEventElapsedTime(&ms, pr.first, pr.second));
    s->ktime_ms += ms;
    s->ktime_n += 1;
    s->kevent_pool.push_back(pr.first);
    s->kevent_pool.push_back(pr.second);
  }
  s->kevents.clear();
  return RCS_OK;
}

StepArgs make_step_args(rcs_sim* s, const AgentArrays& in, const AgentArrays& out, double dt) {
  StepArgs a{};
  a.grid = s->grid;
  a.n = s->n;
  a.n_sorted = s->cell_start + s->grid.len;
  a.in = in;
  a.cell_start = s->cell_start;
  a.groups = s->d_groups;
  a.dt = dt;
  a.ox = out.x;
  a.oy = out.y;
  a.ovx = out.vx;
  a.ovy = out.vy;
  a.status = s->d_status;
  a.collect_stats = 1;
  if (s->trace) {
    a.t_i = s->tr_ti;
    a.fx = s->tr_fx;
    a.fy = s->tr_fy;
    a.nb_count = s->tr_nbc;
  }
  return a;
}

__global__ void copy_meta_kernel(uint32_t n, const uint32_t* n_sorted, AgentArrays in, AgentArrays out,
                                 const DevStatus* st) {
  if (st->failed) return;
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || i >= *n_sorted) return;
  out.id[i] = in.id[i];
  out.grp[i] = in.grp[i];
  out.wp[i] = in.wp[i];
  if (in.pvx) {
    out.pvx[i] = in.pvx[i];
    out.pvy[i] = in.pvy[i];
  }
}

int do_sync(rcs_sim* s) {
  CU_TRY(s, cudaMemcpyAsync(s->h_status, s->d_status, sizeof(DevStatus), cudaMemcpyDeviceToHost, s->stream));
  uint64_t steps_done = 0;
  CU_TRY(s, cudaMemcpyAsync(&steps_done, s->d_steps_done, sizeof(uint64_t), cudaMemcpyDeviceToHost, s->stream));
  CU_TRY(s, cudaStreamSynchronize(s->stream));
  const DevStatus& st = *s->h_status;
  s->stats.oob_count = st.oob_count;
  s->stats.first_oob_id = st.first_oob_id;
  s->stats.nonfinite_count = st.nonfinite_count;
  s->stats.finite_tti_count = st.finite_tti;
  s->stats.neighbour_total = st.neighbour_total;
  s->stats.candidate_total = st.candidate_total;
  s->stats.n_agents = s->n;
  s->stats.steps = steps_done;
  int rc = drain_kevents(s);
  if (rc) return rc;
  if (st.failed) {
    // the step with index k (since the last sync) failed; every later one was skipped on the device
    uint64_t k = steps_done - s->steps_done_at_sync;
    if (k < s->pending.size()) {
      const PendingStep& p = s->pending[k];
      if (p.snapshot_in_srt) {
        s->cur = p.srt;
        s->srt = p.cur;
      } else {
        s->cur = p.cur;
        s->srt = p.srt;
      }
      s->n = p.n;
    }
    invalidate(s);
    CU_TRY(s, cudaMemsetAsync(&s->d_status->failed, 0, sizeof(unsigned int), s->stream));
    CU_TRY(s, cudaStreamSynchronize(s->stream));
    s->err = "Index out of bounds";
    rc = RCS_ERR_OUT_OF_BOUNDS;
  }
  s->pending.clear();
  s->steps_done_at_sync = steps_done;
  return rc;
}

int build_slot_table(rcs_sim* s) {
  if (s->slot_valid) return RCS_OK;
  uint64_t L = std::max<uint64_t>(s->max_id_plus1, 1);
  if (L > s->slot_table_cap) {
    CU_TRY(s, cudaStreamSynchronize(s->stream));
    cudaFree(s->slot_of_id); cudaFree(s->id_rank); cudaFree(s->presence);
    s->slot_of_id = s->id_rank = s->presence = nullptr;
    uint64_t cap = L + L / 2 + 1024;
    CU_TRY(s, dalloc(&s->slot_of_id, cap));
    CU_TRY(s, dalloc(&s->id_rank, cap + 1));
    CU_TRY(s, dalloc(&s->presence, cap + 16));
    s->slot_table_cap = cap;
  }
  CU_TRY(s, cudaMemsetAsync(s->slot_of_id, 0xff, L * sizeof(uint32_t), s->stream));
  if (s->n)
    build_slot_of_id_kernel<<<blocks_for(s->n, 256), 256, 0, s->stream>>>(s->n, s->cur.id, s->slot_of_id, L);
  presence_kernel<<<blocks_for(L, 256), 256, 0, s->stream>>>(L, s->slot_of_id, s->presence);
  s->launches += 2;
  int rc = exclusive_scan(s, s->presence, L, s->id_rank, nullptr);
  if (rc) return rc;
  order_by_id_kernel<<<blocks_for(L, 256), 256, 0, s->stream>>>(L, s->slot_of_id, s->id_rank, s->order_by_id);
  s->launches += 1;
  CU_TRY(s, cudaGetLastError());
  s->slot_valid = true;
  return RCS_OK;
}

template <class T>
int read_array(rcs_sim* s, const T* src, const uint32_t* order, uint32_t n, T* host_out, uint64_t stage_off) {
  T* st = reinterpret_cast<T*>(static_cast<char*>(s->stage) + stage_off);
  gather_kernel<T><<<blocks_for(n, 256), 256, 0, s->stream>>>(n, order, src, st);
  s->launches += 1;
  CU_TRY(s, cudaMemcpyAsync(host_out, st, (uint64_t)n * sizeof(T), cudaMemcpyDeviceToHost, s->stream));
  return RCS_OK;
}

}  // namespace

extern "C" {

uint32_t rcs_abi_version(void) { return RCS_ABI_VERSION; }

const char* rcs_last_error(const rcs_sim* sim) { return sim ? sim->err.c_str() : g_create_error.c_str(); }

int rcs_sim_create(const rcs_sim_desc* desc, rcs_sim** out) {
  if (!desc || !out) {
    g_create_error = "null argument";
    return RCS_ERR_ARG;
  }
  *out = nullptr;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    g_create_error = std::string("no CUDA device: ") + cudaGetErrorString(e) +
                     " (this library has no CPU fallback)";
    return RCS_ERR_NO_DEVICE;
  }
  if (desc->device < 0 || desc->device >= ndev) {
    g_create_error = "device ordinal out of range";
    return RCS_ERR_ARG;
  }
  rcs_sim* s = new rcs_sim();
  s->desc = *desc;
  s->device = desc->device;
  s->cap = std::max<uint64_t>(desc->capacity, 1);
  // LocationHash2D::new, location_hash_2d.rs:33-51
  GridDev& g = s->grid;
  g.offx = desc->offset_x;
  g.offy = desc->offset_y;
  g.res = desc->cell_size;
  g.nx = host_f64_as_usize(desc->width / desc->cell_size);
  uint64_t ny = host_f64_as_usize(desc->height / desc->cell_size);
  if (g.nx != 0 && ny > (0xfffffff0ull / g.nx)) {
    g_create_error = "grid has more than 2^32 cells";
    delete s;
    return RCS_ERR_ARG;
  }
  g.len = g.nx * ny;
  g.x_max = (g.len == 0 || g.nx == 0) ? -1 : (int64_t)((g.len - 1) / g.nx);
  if (s->cap >= 0xfffffff0ull) {
    g_create_error = "capacity must be < 2^32";
    delete s;
    return RCS_ERR_ARG;
  }
  auto fail = [&](int rc) {
    g_create_error = s->err;
    rcs_sim_destroy(s);
    return rc;
  };
#define CR_TRY(call)                                                           \
  do {                                                                         \
    cudaError_t e__ = (call);                                                  \
    if (e__ != cudaSuccess) {                                                  \
      s->err = std::string("CUDA error: ") + cudaGetErrorString(e__) + " at " #call; \
      return fail(RCS_ERR_CUDA);                                               \
    }                                                                          \
  } while (0)
  CR_TRY(cudaSetDevice(s->device));
  CR_TRY(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking));
  if (alloc_agent_arrays(s, s->cur, s->cap) || alloc_agent_arrays(s, s->srt, s->cap)) return fail(RCS_ERR_CUDA);
  CR_TRY(dalloc(&s->cellid, s->cap + 16));
  CR_TRY(dalloc(&s->perm, s->cap + 16));
  CR_TRY(dalloc(&s->order_by_id, s->cap + 16));
  CR_TRY(dalloc(&s->cell_count, g.len + 16));
  CR_TRY(dalloc(&s->cell_start, g.len + 16));
  CR_TRY(dalloc(&s->cursor, g.len + 16));
  CR_TRY(dalloc(&s->scan_total, 4));
  CR_TRY(dalloc(&s->big_list, 4096));
  CR_TRY(dalloc(&s->d_status, 1));
  CR_TRY(dalloc(&s->d_steps_done, 1));
  CR_TRY(dalloc(&s->d_bad, 1));
  CR_TRY(cudaMemset(s->d_status, 0, sizeof(DevStatus)));
  CR_TRY(cudaMemset(s->d_steps_done, 0, sizeof(uint64_t)));
  CR_TRY(cudaMemset(s->scan_total, 0xff, 4 * sizeof(uint32_t)));
  CR_TRY(cudaMemset(s->cell_start, 0, (g.len + 16) * sizeof(uint32_t)));
  CR_TRY(cudaMallocHost(reinterpret_cast<void**>(&s->h_status), sizeof(DevStatus)));
  for (uint32_t k = 0; k < RCS_NUM_EVENTS; ++k) CR_TRY(cudaEventCreate(&s->events[k]));
#undef CR_TRY
  s->stats.first_oob_id = ~0ull;
  *out = s;
  return RCS_OK;
}

void rcs_sim_destroy(rcs_sim* s) {
  if (!s) return;
  cudaSetDevice(s->device);
  if (s->stream) cudaStreamSynchronize(s->stream);
  free_agent_arrays(s->cur);
  free_agent_arrays(s->srt);
  cudaFree(s->cellid); cudaFree(s->perm); cudaFree(s->order_by_id); cudaFree(s->cell_count);
  cudaFree(s->cell_start); cudaFree(s->cursor); cudaFree(s->tile_sums); cudaFree(s->scan_total);
  cudaFree(s->big_list); cudaFree(s->d_groups); cudaFree(s->d_status); cudaFree(s->d_steps_done);
  cudaFree(s->d_bad); cudaFree(s->slot_of_id); cudaFree(s->id_rank); cudaFree(s->presence);
  cudaFree(s->tr_ti); cudaFree(s->tr_fx); cudaFree(s->tr_fy); cudaFree(s->tr_nbc); cudaFree(s->tr_nbo);
  cudaFree(s->tr_nbids); cudaFree(s->stage); cudaFree(s->flush_buf);
  if (s->h_status) cudaFreeHost(s->h_status);
  for (uint32_t k = 0; k < RCS_NUM_EVENTS; ++k)
    if (s->events[k]) cudaEventDestroy(s->events[k]);
  for (auto& pr : s->kevents) {
    cudaEventDestroy(pr.first);
    cudaEventDestroy(pr.second);
  }
  for (auto e : s->kevent_pool) cudaEventDestroy(e);
  if (s->stream) cudaStreamDestroy(s->stream);
  delete s;
}

int rcs_lp_none(rcs_sim* s, uint32_t* out_lp) {
  if (!s || !out_lp) return RCS_ERR_ARG;
  s->lps.push_back(LPDesc{LP_NONE, 0, 0, 0, 0, 0, 0});
  *out_lp = (uint32_t)s->lps.size() - 1;
  return RCS_OK;
}

int rcs_lp_zanlungo(rcs_sim* s, double agent_scale, double obstacle_scale, double reaction_time, double force_distance,
                    double agent_mass, double agent_radius, uint32_t* out_lp) {
  if (!s || !out_lp) return RCS_ERR_ARG;
  s->lps.push_back(LPDesc{LP_ZANLUNGO, agent_scale, obstacle_scale, reaction_time, force_distance, agent_mass,
                          agent_radius});
  *out_lp = (uint32_t)s->lps.size() - 1;
  return RCS_OK;
}

static int push_hl(rcs_sim* s, uint32_t kind, double vx, double vy, uint32_t* out_hl) {
  if (!s || !out_hl) return RCS_ERR_ARG;
  s->hls.push_back(HLDesc{kind, vx, vy});
  *out_hl = (uint32_t)s->hls.size() - 1;
  return RCS_OK;
}
int rcs_hl_constant(rcs_sim* s, double vx, double vy, uint32_t* out_hl) { return push_hl(s, HL_CONSTANT, vx, vy, out_hl); }
int rcs_hl_parity(rcs_sim* s, double vx, double vy, uint32_t* out_hl) { return push_hl(s, HL_PARITY, vx, vy, out_hl); }
int rcs_hl_none(rcs_sim* s, uint32_t* out_hl) { return push_hl(s, HL_NONE, 0, 0, out_hl); }
int rcs_hl_host(rcs_sim* s, uint32_t* out_hl) {
  if (!s || !out_hl) return RCS_ERR_ARG;
  CU_TRY(s, cudaSetDevice(s->device));
  int rc = do_sync(s);
  if (rc) return rc;
  rc = ensure_pref_arrays(s);
  if (rc) return rc;
  s->have_host_hl = true;
  return push_hl(s, HL_HOST, 0, 0, out_hl);
}

static int add_agents_impl(rcs_sim* s, uint64_t n, const uint64_t* ids, const double* xy, const double* vxy,
                           uint32_t hl, uint32_t lp, double eyesight, int32_t source_sink, uint64_t* out_ids) {
  if (!s || (n && !xy)) return RCS_ERR_ARG;
  if (hl >= s->hls.size() || lp >= s->lps.size()) {
    s->err = "unknown planner handle";
    return RCS_ERR_ARG;
  }
  CU_TRY(s, cudaSetDevice(s->device));
  if ((uint64_t)s->n + n > s->cap) {
    s->err = "capacity exceeded";
    return RCS_ERR_CAPACITY;
  }
  // location_to_index of every spawn position first (lib.rs:146-149)
  for (uint64_t k = 0; k < n; ++k) {
    uint64_t idx;
    if (!host_location_to_index(s->grid, xy[2 * k], xy[2 * k + 1], idx)) {
      s->err = "Index out of bounds";
      return RCS_ERR_OUT_OF_BOUNDS;
    }
  }
  if (n == 0) return RCS_OK;
  int rc = do_sync(s);
  if (rc) return rc;
  uint32_t grp = find_or_add_group(s, hl, lp, eyesight, source_sink);
  std::vector<double> hx(n), hy(n), hvx(n, 0.0), hvy(n, 0.0);
  std::vector<uint64_t> hid(n);
  std::vector<uint32_t> hgrp(n, grp), hwp(n, 0u);
  for (uint64_t k = 0; k < n; ++k) {
    hx[k] = xy[2 * k];
    hy[k] = xy[2 * k + 1];
    if (vxy) {
      hvx[k] = vxy[2 * k];
      hvy[k] = vxy[2 * k + 1];
    }
    if (ids) {
      hid[k] = ids[k];
      s->max_id_plus1 = std::max(s->max_id_plus1, ids[k] + 1);
    } else {
      hid[k] = s->last_alloc_agent_id++;  // lib.rs:128-129
    }
    if (out_ids) out_ids[k] = hid[k];
  }
  if (!ids) s->max_id_plus1 = std::max(s->max_id_plus1, s->last_alloc_agent_id);
  const uint32_t o = s->n;
  CU_TRY(s, cudaMemcpy(s->cur.x + o, hx.data(), n * sizeof(double), cudaMemcpyHostToDevice));
  CU_TRY(s, cudaMemcpy(s->cur.y + o, hy.data(), n * sizeof(double), cudaMemcpyHostToDevice));
  CU_TRY(s, cudaMemcpy(s->cur.vx + o, hvx.data(), n * sizeof(double), cudaMemcpyHostToDevice));
  CU_TRY(s, cudaMemcpy(s->cur.vy + o, hvy.data(), n * sizeof(double), cudaMemcpyHostToDevice));
  CU_TRY(s, cudaMemcpy(s->cur.id + o, hid.data(), n * sizeof(uint64_t), cudaMemcpyHostToDevice));
  CU_TRY(s, cudaMemcpy(s->cur.grp + o, hgrp.data(), n * sizeof(uint32_t), cudaMemcpyHostToDevice));
  CU_TRY(s, cudaMemcpy(s->cur.wp + o, hwp.data(), n * sizeof(uint32_t), cudaMemcpyHostToDevice));
  if (s->cur.pvx) {
    const double nan = std::numeric_limits<double>::quiet_NaN();
    fill_f64_kernel<<<blocks_for(n, 256), 256, 0, s->stream>>>(n, s->cur.pvx + o, nan);
    fill_f64_kernel<<<blocks_for(n, 256), 256, 0, s->stream>>>(n, s->cur.pvy + o, nan);
    s->launches += 2;
    CU_TRY(s, cudaStreamSynchronize(s->stream));
  }
  s->n += (uint32_t)n;
  invalidate(s);
  return RCS_OK;
}

int rcs_add_agents(rcs_sim* s, uint64_t n, const double* xy, uint32_t hl, uint32_t lp, double eyesight,
                   uint64_t* out_ids) {
  return add_agents_impl(s, n, nullptr, xy, nullptr, hl, lp, eyesight, -1, out_ids);
}

int rcs_dist_add_agents(rcs_sim* s, uint64_t n, const uint64_t* ids, const double* xy, const double* vxy, uint32_t hl,
                        uint32_t lp, double eyesight) {
  if (n && !ids) return RCS_ERR_ARG;
  return add_agents_impl(s, n, ids, xy, vxy, hl, lp, eyesight, -1, nullptr);
}

int rcs_agent_count(rcs_sim* s, uint64_t* out_n) {
  if (!s || !out_n) return RCS_ERR_ARG;
  CU_TRY(s, cudaSetDevice(s->device));
  int rc = do_sync(s);
  *out_n = s->n;
  return rc;
}

// keep[i] for the removal compaction
__global__ void mark_remove_kernel(uint32_t m, const uint64_t* __restrict__ ids, const uint32_t* __restrict__ slot_of_id,
                                   uint64_t table_len, uint32_t* __restrict__ keep, unsigned int* bad) {
  uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= m) return;
  uint64_t v = ids[k];
  uint32_t sl = v < table_len ? slot_of_id[v] : 0xffffffffu;
  if (sl == 0xffffffffu) {
    atomicAdd(bad, 1u);
    return;
  }
  keep[sl] = 0u;
}

__global__ void compact_kernel(uint32_t n, const uint32_t* __restrict__ keep, const uint32_t* __restrict__ pos,
                               AgentArrays in, AgentArrays out) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || !keep[i]) return;
  uint32_t k = pos[i];
  out.x[k] = in.x[i];
  out.y[k] = in.y[i];
  out.vx[k] = in.vx[i];
  out.vy[k] = in.vy[i];
  out.id[k] = in.id[i];
  out.grp[k] = in.grp[i];
  out.wp[k] = in.wp[i];
  if (in.pvx) {
    out.pvx[k] = in.pvx[i];
    out.pvy[k] = in.pvy[i];
  }
}

int rcs_remove_agents(rcs_sim* s, uint64_t m, const uint64_t* ids) {
  if (!s || (m && !ids)) return RCS_ERR_ARG;
  if (m == 0) return RCS_OK;
  CU_TRY(s, cudaSetDevice(s->device));
  int rc = do_sync(s);
  if (rc) return rc;
  rc = build_slot_table(s);
  if (rc) return rc;
  rc = ensure_stage(s, m * sizeof(uint64_t));
  if (rc) return rc;
  uint64_t* d_ids = static_cast<uint64_t*>(s->stage);
  CU_TRY(s, cudaMemcpyAsync(d_ids, ids, m * sizeof(uint64_t), cudaMemcpyHostToDevice, s->stream));
  CU_TRY(s, cudaMemsetAsync(s->d_bad, 0, sizeof(unsigned int), s->stream));
  uint32_t n = s->n;
  fill_u32_kernel<<<blocks_for(n, 256), 256, 0, s->stream>>>(n, s->cellid, 1u);
  mark_remove_kernel<<<blocks_for(m, 256), 256, 0, s->stream>>>((uint32_t)m, d_ids, s->slot_of_id,
                                                                std::max<uint64_t>(s->max_id_plus1, 1), s->cellid,
                                                                s->d_bad);
  s->launches += 2;
  unsigned int bad = 0;
  CU_TRY(s, cudaMemcpyAsync(&bad, s->d_bad, sizeof(bad), cudaMemcpyDeviceToHost, s->stream));
  CU_TRY(s, cudaStreamSynchronize(s->stream));
  if (bad) {
    s->err = "unknown agent id";
    return RCS_ERR_ARG;
  }
  rc = exclusive_scan(s, s->cellid, n, s->perm, nullptr);
  if (rc) return rc;
  compact_kernel<<<blocks_for(n, 256), 256, 0, s->stream>>>(n, s->cellid, s->perm, s->cur, s->srt);
  s->launches += 1;
  uint32_t kept = 0;
  CU_TRY(s, cudaMemcpyAsync(&kept, s->perm + n, sizeof(uint32_t), cudaMemcpyDeviceToHost, s->stream));
  CU_TRY(s, cudaStreamSynchronize(s->stream));
  std::swap(s->cur, s->srt);
  s->n = kept;
  invalidate(s);
  return RCS_OK;
}

int rcs_set_state(rcs_sim* s, uint64_t m, const uint64_t* ids, const double* x, const double* y, const double* vx,
                  const double* vy) {
  if (!s) return RCS_ERR_ARG;
  if (m == 0) return RCS_OK;
  CU_TRY(s, cudaSetDevice(s->device));
  int rc = do_sync(s);
  if (rc) return rc;
  if (!ids && m != s->n) {
    s->err = "ids == NULL requires n == agent count";
    return RCS_ERR_ARG;
  }
  // the reference's index rejects out-of-grid positions (add_or_update, location_hash_2d.rs:126-130)
  if (x && y) {
    for (uint64_t k = 0; k < m; ++k) {
      uint64_t idx;
      if (!host_location_to_index(s->grid, x[k], y[k], idx)) {
        s->err = "Index out of bounds";
        return RCS_ERR_OUT_OF_BOUNDS;
      }
    }
  }
  rc = build_slot_table(s);
  if (rc) return rc;
  rc = ensure_stage(s, m * (sizeof(uint64_t) + sizeof(double)));
  if (rc) return rc;
  uint64_t* d_ids = nullptr;
  double* d_val = reinterpret_cast<double*>(static_cast<char*>(s->stage));
  if (ids) {
    d_ids = reinterpret_cast<uint64_t*>(static_cast<char*>(s->stage) + m * sizeof(double));
    CU_TRY(s, cudaMemcpyAsync(d_ids, ids, m * sizeof(uint64_t), cudaMemcpyHostToDevice, s->stream));
  }
  CU_TRY(s, cudaMemsetAsync(s->d_bad, 0, sizeof(unsigned int), s->stream));
  const double* srcs[4] = {x, y, vx, vy};
  double* dsts[4] = {s->cur.x, s->cur.y, s->cur.vx, s->cur.vy};
  for (int a = 0; a < 4; ++a) {
    if (!srcs[a]) continue;
    CU_TRY(s, cudaMemcpyAsync(d_val, srcs[a], m * sizeof(double), cudaMemcpyHostToDevice, s->stream));
    scatter_by_id_kernel<double><<<blocks_for(m, 256), 256, 0, s->stream>>>(
        (uint32_t)m, s->order_by_id, d_ids, s->slot_of_id, std::max<uint64_t>(s->max_id_plus1, 1), d_val, 1, dsts[a],
        s->d_bad);
    s->launches += 1;
  }
  unsigned int bad = 0;
  CU_TRY(s, cudaMemcpyAsync(&bad, s->d_bad, sizeof(bad), cudaMemcpyDeviceToHost, s->stream));
  CU_TRY(s, cudaStreamSynchronize(s->stream));
  s->index_valid = false;
  s->tr_valid = false;
  if (bad) {
    s->err = "unknown agent id";
    return RCS_ERR_ARG;
  }
  return RCS_OK;
}

int rcs_set_preferred_velocity(rcs_sim* s, uint64_t m, const uint64_t* ids, const double* vxy) {
  if (!s || (m && !vxy)) return RCS_ERR_ARG;
  if (m == 0) return RCS_OK;
  CU_TRY(s, cudaSetDevice(s->device));
  if (!s->cur.pvx) {
    s->err = "no rcs_hl_host planner exists on this handle";
    return RCS_ERR_ARG;
  }
  if (!ids && m != s->n) {
    s->err = "ids == NULL requires n == agent count";
    return RCS_ERR_ARG;
  }
  int rc = build_slot_table(s);
  if (rc) return rc;
  rc = ensure_stage(s, m * (sizeof(uint64_t) + 2 * sizeof(double)));
  if (rc) return rc;
  double* d_val = reinterpret_cast<double*>(static_cast<char*>(s->stage));
  uint64_t* d_ids = nullptr;
  CU_TRY(s, cudaMemcpyAsync(d_val, vxy, 2 * m * sizeof(double), cudaMemcpyHostToDevice, s->stream));
  if (ids) {
    d_ids = reinterpret_cast<uint64_t*>(static_cast<char*>(s->stage) + 2 * m * sizeof(double));
    CU_TRY(s, cudaMemcpyAsync(d_ids, ids, m * sizeof(uint64_t), cudaMemcpyHostToDevice, s->stream));
  }
  CU_TRY(s, cudaMemsetAsync(s->d_bad, 0, sizeof(unsigned int), s->stream));
  const uint64_t L = std::max<uint64_t>(s->max_id_plus1, 1);
  scatter_by_id_kernel<double><<<blocks_for(m, 256), 256, 0, s->stream>>>((uint32_t)m, s->order_by_id, d_ids,
                                                                          s->slot_of_id, L, d_val, 2, s->cur.pvx,
                                                                          s->d_bad);
  scatter_by_id_kernel<double><<<blocks_for(m, 256), 256, 0, s->stream>>>((uint32_t)m, s->order_by_id, d_ids,
                                                                          s->slot_of_id, L, d_val + 1, 2, s->cur.pvy,
                                                                          s->d_bad);
  s->launches += 2;
  CU_TRY(s, cudaGetLastError());
  if (ids) {
    unsigned int bad = 0;
    CU_TRY(s, cudaMemcpyAsync(&bad, s->d_bad, sizeof(bad), cudaMemcpyDeviceToHost, s->stream));
    CU_TRY(s, cudaStreamSynchronize(s->stream));
    if (bad) {
      s->err = "unknown agent id";
      return RCS_ERR_ARG;
    }
  }
  // the stage buffer is reused by later calls on the same stream: stream order keeps this safe
  return RCS_OK;
}

int rcs_read_agents(rcs_sim* s, uint32_t order, uint64_t cap, uint64_t* ids, double* x, double* y, double* vx,
                    double* vy, uint32_t* next_waypoint, uint64_t* out_n) {
  if (!s) return RCS_ERR_ARG;
  CU_TRY(s, cudaSetDevice(s->device));
  int rc = do_sync(s);
  if (out_n) *out_n = s->n;
  if (rc) return rc;
  uint32_t n = s->n;
  if (n == 0) return RCS_OK;
  if (cap < n) {
    s->err = "output capacity too small";
    return RCS_ERR_CAPACITY;
  }
  const uint32_t* ord = nullptr;
  if (order == RCS_ORDER_ID) {
    rc = build_slot_table(s);
    if (rc) return rc;
    ord = s->order_by_id;
  }
  if (!ord) {
    // storage order: straight copies
    if (ids) CU_TRY(s, cudaMemcpyAsync(ids, s->cur.id, n * sizeof(uint64_t), cudaMemcpyDeviceToHost, s->stream));
    if (x) CU_TRY(s, cudaMemcpyAsync(x, s->cur.x, n * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
    if (y) CU_TRY(s, cudaMemcpyAsync(y, s->cur.y, n * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
    if (vx) CU_TRY(s, cudaMemcpyAsync(vx, s->cur.vx, n * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
    if (vy) CU_TRY(s, cudaMemcpyAsync(vy, s->cur.vy, n * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
    if (next_waypoint)
      CU_TRY(s, cudaMemcpyAsync(next_waypoint, s->cur.wp, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, s->stream));
  } else {
    rc = ensure_stage(s, (uint64_t)n * 48 + 256);
    if (rc) return rc;
    uint64_t off = 0;
    if (ids) { rc = read_array<uint64_t>(s, s->cur.id, ord, n, ids, off); off += (uint64_t)n * 8; if (rc) return rc; }
    if (x) { rc = read_array<double>(s, s->cur.x, ord, n, x, off); off += (uint64_t)n * 8; if (rc) return rc; }
    if (y) { rc = read_array<double>(s, s->cur.y, ord, n, y, off); off += (uint64_t)n * 8; if (rc) return rc; }
    if (vx) { rc = read_array<double>(s, s->cur.vx, ord, n, vx, off); off += (uint64_t)n * 8; if (rc) return rc; }
    if (vy) { rc = read_array<double>(s, s->cur.vy, ord, n, vy, off); off += (uint64_t)n * 8; if (rc) return rc; }
    if (next_waypoint) { rc = read_array<uint32_t>(s, s->cur.wp, ord, n, next_waypoint, off); if (rc) return rc; }
  }
  CU_TRY(s, cudaStreamSynchronize(s->stream));
  return RCS_OK;
}

int rcs_step_async(rcs_sim* s, uint64_t secs, uint32_t nanos, uint32_t flags) {
  if (!s) return RCS_ERR_ARG;
  CU_TRY(s, cudaSetDevice(s->device));
  const double dt = (double)secs + (double)nanos / 1000000000.0;  // Duration::as_secs_f64
  int rc = upload_groups(s);
  if (rc) return rc;
  const bool no_commit = (flags & RCS_STEP_NO_COMMIT) != 0;
  const uint32_t n = s->n;
  PendingStep p{s->cur, s->srt, true, n};
  begin_step_kernel<<<1, 1, 0, s->stream>>>(s->d_status);
  s->launches += 1;
  if (n) {
    if (s->any_zanlungo || s->trace) {
      // A1-A4 rebuild, then the fused query + planner + integrate kernel reading srt, writing cur
      rc = build_index(s, false);
      if (rc) return rc;
      StepArgs a = make_step_args(s, s->srt, s->cur, dt);
      launch_step_kernel(s, a, n);
      copy_meta_kernel<<<blocks_for(n, 256), 256, 0, s->stream>>>(n, a.n_sorted, s->srt, s->cur, s->d_status);
      s->launches += 1;
      p.snapshot_in_srt = true;
      if (no_commit) std::swap(s->cur, s->srt);  // pre-step snapshot (sorted) becomes current again
    } else {
      // NoLocalPlan only: the radius query cannot influence the result (no_local_plan.rs:10-17), so the
      // step is a pure stream over the agents in storage order; new x,y,vx,vy go to the spare buffers.
      StepArgs a = make_step_args(s, s->cur, s->srt, dt);
      a.n_sorted = s->scan_total + 1;  // constant 0xffffffff: every slot below n is live in storage order
      launch_step_kernel(s, a, n, false);
      p.snapshot_in_srt = false;
      if (!no_commit) {
        std::swap(s->cur.x, s->srt.x);
        std::swap(s->cur.y, s->srt.y);
        std::swap(s->cur.vx, s->srt.vx);
        std::swap(s->cur.vy, s->srt.vy);
      }
    }
  }
  end_step_kernel<<<1, 1, 0, s->stream>>>(s->d_status, s->d_steps_done, no_commit ? 0 : 1);
  s->launches += 1;
  CU_TRY(s, cudaGetLastError());
  s->pending.push_back(p);
  s->steps_enqueued += 1;
  s->index_valid = false;
  s->slot_valid = false;
  s->tr_valid = false;
  if (s->trace && n) {
    // neighbour lists of this step (debug path: synchronous)
    rc = exclusive_scan(s, s->tr_nbc, n, s->tr_nbo, nullptr);
    if (rc) return rc;
    uint32_t total = 0;
    CU_TRY(s, cudaMemcpyAsync(&total, s->tr_nbo + n, sizeof(uint32_t), cudaMemcpyDeviceToHost, s->stream));
    CU_TRY(s, cudaStreamSynchronize(s->stream));
    if (total > s->tr_nbids_cap) {
      cudaFree(s->tr_nbids);
      s->tr_nbids = nullptr;
      CU_TRY(s, dalloc(&s->tr_nbids, (uint64_t)total + total / 4 + 1024));
      s->tr_nbids_cap = (uint64_t)total + total / 4 + 1024;
    }
    // the sorted snapshot of this step: srt normally, cur after a no-commit swap
    const AgentArrays& snap = no_commit ? s->cur : s->srt;
    StepArgs a = make_step_args(s, snap, snap, dt);
    trace_neighbours_kernel<<<blocks_for(n, 128), 128, 0, s->stream>>>(a, s->tr_nbo, s->tr_nbids);
    s->launches += 1;
    CU_TRY(s, cudaGetLastError());
    s->tr_nb_total = total;
    s->tr_n = n;
    s->tr_valid = true;
  }
  return RCS_OK;
}

int rcs_sync(rcs_sim* s) {
  if (!s) return RCS_ERR_ARG;
  CU_TRY(s, cudaSetDevice(s->device));
  return do_sync(s);
}

int rcs_step(rcs_sim* s, uint64_t secs, uint32_t nanos) {
  int rc = rcs_step_async(s, secs, nanos, RCS_STEP_DEFAULT);
  if (rc) return rc;
  return rcs_sync(s);
}

int rcs_step_stats(rcs_sim* s, rcs_stats* out) {
  if (!s || !out) return RCS_ERR_ARG;
  CU_TRY(s, cudaSetDevice(s->device));
  int rc = do_sync(s);
  *out = s->stats;
  return rc;
}

int rcs_poll_events(rcs_sim* s, uint64_t, uint64_t*, double*, uint64_t* n_spawned, uint64_t, uint64_t*,
                    uint64_t* n_destroyed) {
  if (!s) return RCS_ERR_ARG;
  if (n_spawned) *n_spawned = 0;
  if (n_destroyed) *n_destroyed = 0;
  return RCS_OK;
}

int rcs_add_source_sink(rcs_sim* s, const rcs_source_sink_desc*, uint64_t*) {
  if (!s) return RCS_ERR_ARG;
  s->err = "source sinks are not implemented yet";
  return RCS_ERR_ARG;
}
int rcs_remove_source_sink(rcs_sim* s, uint64_t) {
  if (!s) return RCS_ERR_ARG;
  s->err = "source sinks are not implemented yet";
  return RCS_ERR_ARG;
}

int rcs_cell_of(rcs_sim* s, uint64_t n, const double* xy, int64_t* out_idx) {
  if (!s || (n && (!xy || !out_idx))) return RCS_ERR_ARG;
  if (n == 0) return RCS_OK;
  CU_TRY(s, cudaSetDevice(s->device));
  int rc = ensure_stage(s, n * 24);
  if (rc) return rc;
  double* d_xy = static_cast<double*>(s->stage);
  long long* d_out = reinterpret_cast<long long*>(static_cast<char*>(s->stage) + n * 16);
  CU_TRY(s, cudaMemcpyAsync(d_xy, xy, n * 16, cudaMemcpyHostToDevice, s->stream));
  cell_of_kernel<<<blocks_for(n, 256), 256, 0, s->stream>>>(s->grid, (uint32_t)n, d_xy, d_out);
  s->launches += 1;
  CU_TRY(s, cudaMemcpyAsync(out_idx, d_out, n * 8, cudaMemcpyDeviceToHost, s->stream));
  CU_TRY(s, cudaStreamSynchronize(s->stream));
  return RCS_OK;
}

static int ensure_index(rcs_sim* s) {
  int rc = do_sync(s);
  if (rc) return rc;
  if (s->index_valid) return RCS_OK;
  begin_step_kernel<<<1, 1, 0, s->stream>>>(s->d_status);
  s->launches += 1;
  rc = build_index(s, true);
  if (rc) return rc;
  s->index_valid = true;
  return RCS_OK;
}

int rcs_query_radius(rcs_sim* s, uint64_t nq, const double* qxy, const double* radius, uint64_t* offsets,
                     uint64_t* out_ids, uint64_t ids_cap) {
  if (!s || (nq && (!qxy || !radius || !offsets))) return RCS_ERR_ARG;
  CU_TRY(s, cudaSetDevice(s->device));
  if (nq == 0) return RCS_OK;
  int rc = ensure_index(s);
  if (rc) return rc;
  std::vector<double> thr(nq);
  for (uint64_t q = 0; q < nq; ++q) thr[q] = radius_threshold(radius[q]);
  // stage: qxy (16 nq) | radius (8 nq) | thr (8 nq) | counts (4 nq)
  rc = ensure_stage(s, nq * 40 + 64);
  if (rc) return rc;
  char* base = static_cast<char*>(s->stage);
  double* d_q = reinterpret_cast<double*>(base);
  double* d_r = reinterpret_cast<double*>(base + nq * 16);
  double* d_t = reinterpret_cast<double*>(base + nq * 24);
  uint32_t* d_c = reinterpret_cast<uint32_t*>(base + nq * 32);
  CU_TRY(s, cudaMemcpyAsync(d_q, qxy, nq * 16, cudaMemcpyHostToDevice, s->stream));
  CU_TRY(s, cudaMemcpyAsync(d_r, radius, nq * 8, cudaMemcpyHostToDevice, s->stream));
  CU_TRY(s, cudaMemcpyAsync(d_t, thr.data(), nq * 8, cudaMemcpyHostToDevice, s->stream));
  query_radius_kernel<<<blocks_for(nq, 128), 128, 0, s->stream>>>(s->grid, s->cell_start, s->srt.x, s->srt.y, s->srt.id,
                                                                  (uint32_t)nq, d_q, d_r, d_t, d_c, nullptr, nullptr, 0);
  s->launches += 1;
  std::vector<uint32_t> counts(nq);
  CU_TRY(s, cudaMemcpyAsync(counts.data(), d_c, nq * 4, cudaMemcpyDeviceToHost, s->stream));
  CU_TRY(s, cudaStreamSynchronize(s->stream));
  uint64_t total = 0;
  for (uint64_t q = 0; q < nq; ++q) {
    offsets[q] = total;
    total += counts[q];
  }
  offsets[nq] = total;
  if (total > ids_cap || (total && !out_ids)) {
    s->err = "ids_cap too small";
    return RCS_ERR_CAPACITY;
  }
  if (total == 0) return RCS_OK;
  uint64_t* d_off = nullptr;
  uint64_t* d_ids = nullptr;
  CU_TRY(s, cudaMalloc(reinterpret_cast<void**>(&d_off), (nq + 1) * 8));
  CU_TRY(s, cudaMalloc(reinterpret_cast<void**>(&d_ids), total * 8));
  cudaMemcpyAsync(d_off, offsets, (nq + 1) * 8, cudaMemcpyHostToDevice, s->stream);
  query_radius_kernel<<<blocks_for(nq, 128), 128, 0, s->stream>>>(s->grid, s->cell_start, s->srt.x, s->srt.y, s->srt.id,
                                                                  (uint32_t)nq, d_q, d_r, d_t, d_c, d_off, d_ids, 1);
  s->launches += 1;
  cudaMemcpyAsync(out_ids, d_ids, total * 8, cudaMemcpyDeviceToHost, s->stream);
  cudaError_t e = cudaStreamSynchronize(s->stream);
  cudaFree(d_off);
  cudaFree(d_ids);
  CU_TRY(s, e);
  return RCS_OK;
}

int rcs_query_knn(rcs_sim* s, uint64_t, const double*, uint64_t, uint64_t*, uint64_t*) {
  if (!s) return RCS_ERR_ARG;
  s->err = "kNN is not implemented yet";
  return RCS_ERR_ARG;
}

int rcs_index_add_or_update(rcs_sim* s, uint64_t n, const uint64_t* ids, const double* xy) {
  if (!s || (n && (!ids || !xy))) return RCS_ERR_ARG;
  if (n == 0) return RCS_OK;
  CU_TRY(s, cudaSetDevice(s->device));
  int rc = do_sync(s);
  if (rc) return rc;
  if (s->lps.empty()) { uint32_t t; rcs_lp_none(s, &t); }
  if (s->hls.empty()) { uint32_t t; rcs_hl_none(s, &t); }
  // split into updates of known ids and inserts of new ones (location_hash_2d.rs:134-145)
  rc = build_slot_table(s);
  if (rc) return rc;
  std::vector<uint32_t> slots(std::max<uint64_t>(s->max_id_plus1, 1));
  CU_TRY(s, cudaMemcpy(slots.data(), s->slot_of_id, slots.size() * sizeof(uint32_t), cudaMemcpyDeviceToHost));
  std::vector<uint64_t> upd_ids, new_ids;
  std::vector<double> ux, uy, nxy;
  for (uint64_t k = 0; k < n; ++k) {
    if (ids[k] >= (1ull << 31)) {
      s->err = "index ids must be < 2^31";
      return RCS_ERR_ARG;
    }
    uint64_t idx;
    if (!host_location_to_index(s->grid, xy[2 * k], xy[2 * k + 1], idx)) {
      s->err = "Index out of bounds";
      return RCS_ERR_OUT_OF_BOUNDS;
    }
    bool known = ids[k] < s->max_id_plus1 && slots[ids[k]] != 0xffffffffu;
    // an id inserted earlier in this very call counts as known for later entries
    if (!known) {
      auto it = std::find(new_ids.begin(), new_ids.end(), ids[k]);
      if (it != new_ids.end()) {
        size_t j = it - new_ids.begin();
        nxy[2 * j] = xy[2 * k];
        nxy[2 * j + 1] = xy[2 * k + 1];
        continue;
      }
      new_ids.push_back(ids[k]);
      nxy.push_back(xy[2 * k]);
      nxy.push_back(xy[2 * k + 1]);
    } else {
      upd_ids.push_back(ids[k]);
      ux.push_back(xy[2 * k]);
      uy.push_back(xy[2 * k + 1]);
    }
  }
  if (!upd_ids.empty()) {
    rc = rcs_set_state(s, upd_ids.size(), upd_ids.data(), ux.data(), uy.data(), nullptr, nullptr);
    if (rc) return rc;
  }
  if (!new_ids.empty()) {
    rc = add_agents_impl(s, new_ids.size(), new_ids.data(), nxy.data(), nullptr, 0, 0, 0.0, -1, nullptr);
    if (rc) return rc;
  }
  return RCS_OK;
}

int rcs_index_remove(rcs_sim* s, uint64_t n, const uint64_t* ids) {
  if (!s || (n && !ids)) return RCS_ERR_ARG;
  if (n == 0) return RCS_OK;
  CU_TRY(s, cudaSetDevice(s->device));
  // remove_agent of an unknown id is a no-op in the reference (location_hash_2d.rs:260-267)
  int rc = do_sync(s);
  if (rc) return rc;
  rc = build_slot_table(s);
  if (rc) return rc;
  std::vector<uint32_t> slots(std::max<uint64_t>(s->max_id_plus1, 1));
  CU_TRY(s, cudaMemcpy(slots.data(), s->slot_of_id, slots.size() * sizeof(uint32_t), cudaMemcpyDeviceToHost));
  std::vector<uint64_t> known;
  for (uint64_t k = 0; k < n; ++k)
    if (ids[k] < s->max_id_plus1 && slots[ids[k]] != 0xffffffffu &&
        std::find(known.begin(), known.end(), ids[k]) == known.end())
      known.push_back(ids[k]);
  return rcs_remove_agents(s, known.size(), known.data());
}

int rcs_set_trace(rcs_sim* s, int32_t on) {
  if (!s) return RCS_ERR_ARG;
  CU_TRY(s, cudaSetDevice(s->device));
  int rc = do_sync(s);
  if (rc) return rc;
  if (on && !s->tr_ti) {
    CU_TRY(s, dalloc(&s->tr_ti, s->cap));
    CU_TRY(s, dalloc(&s->tr_fx, s->cap));
    CU_TRY(s, dalloc(&s->tr_fy, s->cap));
    CU_TRY(s, dalloc(&s->tr_nbc, s->cap + 16));
    CU_TRY(s, dalloc(&s->tr_nbo, s->cap + 16));
  }
  s->trace = on != 0;
  s->tr_valid = false;
  return RCS_OK;
}

int rcs_trace_sizes(rcs_sim* s, uint64_t* n_agents, uint64_t* n_neighbours) {
  if (!s) return RCS_ERR_ARG;
  if (!s->tr_valid) {
    s->err = "no trace recorded (enable with rcs_set_trace, then step)";
    return RCS_ERR_ARG;
  }
  if (n_agents) *n_agents = s->tr_n;
  if (n_neighbours) *n_neighbours = s->tr_nb_total;
  return RCS_OK;
}

int rcs_read_trace(rcs_sim* s, uint64_t* ids, double* t_i, double* fx, double* fy, uint64_t* nb_offsets,
                   uint64_t* nb_ids) {
  if (!s) return RCS_ERR_ARG;
  if (!s->tr_valid) {
    s->err = "no trace recorded (enable with rcs_set_trace, then step)";
    return RCS_ERR_ARG;
  }
  CU_TRY(s, cudaSetDevice(s->device));
  CU_TRY(s, cudaStreamSynchronize(s->stream));
  const uint32_t n = s->tr_n;
  // the trace arrays are in the canonical sorted order of the traced step; the ids of that order
  // are in whichever buffer holds the sorted snapshot: after a committed step cur has the same order
  std::vector<uint64_t> sid(n);
  std::vector<double> hti(n), hfx(n), hfy(n);
  std::vector<uint32_t> hoff(n + 1);
  std::vector<uint64_t> hnb(s->tr_nb_total);
  CU_TRY(s, cudaMemcpy(sid.data(), s->cur.id, n * sizeof(uint64_t), cudaMemcpyDeviceToHost));
  CU_TRY(s, cudaMemcpy(hti.data(), s->tr_ti, n * sizeof(double), cudaMemcpyDeviceToHost));
  CU_TRY(s, cudaMemcpy(hfx.data(), s->tr_fx, n * sizeof(double), cudaMemcpyDeviceToHost));
  CU_TRY(s, cudaMemcpy(hfy.data(), s->tr_fy, n * sizeof(double), cudaMemcpyDeviceToHost));
  CU_TRY(s, cudaMemcpy(hoff.data(), s->tr_nbo, (n + 1) * sizeof(uint32_t), cudaMemcpyDeviceToHost));
  if (s->tr_nb_total)
    CU_TRY(s, cudaMemcpy(hnb.data(), s->tr_nbids, s->tr_nb_total * sizeof(uint64_t), cudaMemcpyDeviceToHost));
  std::vector<uint32_t> order(n);
  std::iota(order.begin(), order.end(), 0u);
  std::sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return sid[a] < sid[b]; });
  uint64_t off = 0;
  for (uint32_t r = 0; r < n; ++r) {
    uint32_t k = order[r];
    if (ids) ids[r] = sid[k];
    if (t_i) t_i[r] = hti[k];
    if (fx) fx[r] = hfx[k];
    if (fy) fy[r] = hfy[k];
    if (nb_offsets) nb_offsets[r] = off;
    for (uint32_t j = hoff[k]; j < hoff[k + 1]; ++j) {
      if (nb_ids) nb_ids[off] = hnb[j];
      off++;
    }
  }
  if (nb_offsets) nb_offsets[n] = off;
  return RCS_OK;
}

int rcs_set_option(rcs_sim* s, uint32_t option, uint64_t value) {
  if (!s) return RCS_ERR_ARG;
  if (option == RCS_OPT_STEP_KERNEL && value <= 2) {
    s->opt_step_kernel = (uint32_t)value;
    return RCS_OK;
  }
  s->err = "unknown option or value";
  return RCS_ERR_ARG;
}

int rcs_event_record(rcs_sim* s, uint32_t slot) {
  if (!s || slot >= RCS_NUM_EVENTS) return RCS_ERR_ARG;
  CU_TRY(s, cudaSetDevice(s->device));
  CU_TRY(s, cudaEventRecord(s->events[slot], s->stream));
  return RCS_OK;
}

int rcs_event_elapsed_ms(rcs_sim* s, uint32_t a, uint32_t b, float* out_ms) {
  if (!s || a >= RCS_NUM_EVENTS || b >= RCS_NUM_EVENTS || !out_ms) return RCS_ERR_ARG;
  CU_TRY(s, cudaSetDevice(s->device));
  CU_TRY(s, cudaEventSynchronize(s->events[b]));
  CU_TRY(s, cudaEventElapsedTime(out_ms, s->events[a], s->events[b]));
  return RCS_OK;
}

int rcs_host_alloc(uint64_t bytes, void** out) {
  if (!out) return RCS_ERR_ARG;
  cudaError_t e = cudaMallocHost(out, std::max<uint64_t>(bytes, 1));
  if (e != cudaSuccess) {
    g_create_error = std::string("CUDA error: ") + cudaGetErrorString(e) + " at cudaMallocHost";
    return RCS_ERR_CUDA;
  }
  return RCS_OK;
}

int rcs_host_free(void* p) {
  if (p) cudaFreeHost(p);
  return RCS_OK;
}

int rcs_flush_l2(rcs_sim* s, uint64_t bytes) {
  if (!s) return RCS_ERR_ARG;
  CU_TRY(s, cudaSetDevice(s->device));
  if (bytes > s->flush_bytes) {
    CU_TRY(s, cudaStreamSynchronize(s->stream));
    cudaFree(s->flush_buf);
    s->flush_buf = nullptr;
    CU_TRY(s, cudaMalloc(&s->flush_buf, bytes));
    s->flush_bytes = bytes;
  }
  flush_l2_kernel<<<148 * 8, 256, 0, s->stream>>>(static_cast<uint4*>(s->flush_buf), bytes / 16);
  s->launches += 1;
  CU_TRY(s, cudaGetLastError());
  return RCS_OK;
}

int rcs_kernel_timing(rcs_sim* s, int32_t on) {
  if (!s) return RCS_ERR_ARG;
  CU_TRY(s, cudaSetDevice(s->device));
  int rc = drain_kevents(s);
  if (rc) return rc;
  if (on && !s->ktiming) {
    s->ktime_ms = 0.0;
    s->ktime_n = 0;
  }
  s->ktiming = on != 0;
  return RCS_OK;
}

int rcs_kernel_time_ms(rcs_sim* s, double* out_ms, uint64_t* out_launches) {
  if (!s) return RCS_ERR_ARG;
  CU_TRY(s, cudaSetDevice(s->device));
  int rc = drain_kevents(s);
  if (rc) return rc;
  if (out_ms) *out_ms = s->ktime_ms;
  if (out_launches) *out_launches = s->ktime_n;
  return RCS_OK;
}

int rcs_launch_count(rcs_sim* s, uint64_t* out) {
  if (!s || !out) return RCS_ERR_ARG;
  *out = s->launches;
  return RCS_OK;
}

int rcs_fp64_peak(int32_t device, double* out_tflops, double* out_dadd_tops) {
  cudaError_t e = cudaSetDevice(device);
  if (e != cudaSuccess) {
    g_create_error = std::string("no CUDA device: ") + cudaGetErrorString(e);
    return RCS_ERR_NO_DEVICE;
  }
  const int blocks = 148 * 8, threads = 256, iters = 1 << 14;
  double* d = nullptr;
  if (cudaMalloc(reinterpret_cast<void**>(&d), (size_t)blocks * threads * sizeof(double)) != cudaSuccess)
    return RCS_ERR_CUDA;
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  double res[2] = {0, 0};
  for (int mode = 0; mode < 2; ++mode) {
    fp64_peak_kernel<<<blocks, threads>>>(d, iters, mode);  // warm-up
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
      cudaEventRecord(a);
      fp64_peak_kernel<<<blocks, threads>>>(d, iters, mode);
      cudaEventRecord(b);
      cudaEventSynchronize(b);
      float ms = 0;
      cudaEventElapsedTime(&ms, a, b);
      best = std::min(best, ms);
    }
    double ops = (double)blocks * threads * (double)iters * 8.0;
    res[mode] = ops / (best * 1e-3) / 1e12;
  }
  cudaEventDestroy(a);
  cudaEventDestroy(b);
  e = cudaDeviceSynchronize();
  cudaFree(d);
  if (e != cudaSuccess) return RCS_ERR_CUDA;
  if (out_dadd_tops) *out_dadd_tops = res[0];
  if (out_tflops) *out_tflops = res[1] * 2.0;
  return RCS_OK;
}

int rcs_nccl_unique_id(uint8_t*) {
  g_create_error = "multi-GPU strips are not implemented yet";
  return RCS_ERR_NCCL;
}
int rcs_dist_init(rcs_sim* s, int32_t, int32_t, const uint8_t*) {
  if (!s) return RCS_ERR_ARG;
  s->err = "multi-GPU strips are not implemented yet";
  return RCS_ERR_NCCL;
}
int rcs_dist_strip(rcs_sim* s, int32_t rank, int32_t world, uint64_t* c0, uint64_t* c1) {
  if (!s || !c0 || !c1 || world <= 0 || rank < 0 || rank >= world) return RCS_ERR_ARG;
  uint64_t cols = s->grid.nx;
  *c0 = cols * (uint64_t)rank / (uint64_t)world;
  *c1 = cols * (uint64_t)(rank + 1) / (uint64_t)world;
  return RCS_OK;
}

}  // extern "C"
