// rcs_host_core.inl -- device memory, groups, the index rebuild and rcs_sync of a simulation handle.
// Part of rcs.cu (single translation unit).

namespace rcs_host {

thread_local std::string g_create_error;

// Rust `f64 as usize` on the host (same rule as rcs_math.cuh)
static uint64_t host_f64_as_usize(double v) {
  if (!(v > 0.0)) return 0;
  if (v >= 18446744073709551616.0) return std::numeric_limits<uint64_t>::max();
  return (uint64_t)v;
}

static int64_t host_floor_as_i64(double v) {
  v = std::floor(v);
  if (v != v) return 0;
  if (v >= 9223372036854775808.0) return std::numeric_limits<int64_t>::max();
  if (v <= -9223372036854775808.0) return std::numeric_limits<int64_t>::min();
  return (int64_t)v;
}

static bool host_location_to_index(const GridDev& g, double px, double py, uint64_t& idx) {
  uint64_t x_idx = host_f64_as_usize((px - g.offx) / g.res);
  uint64_t y_idx = host_f64_as_usize((py - g.offy) / g.res);
  idx = x_idx * g.nx + y_idx;
  return idx < g.len;
}

// Smallest double T such that sqrt(T) >= R (correctly rounded sqrt is monotone), so that the
// reference's strict test `norm < radius` (location_hash_2d.rs:251) is exactly `norm_squared < T`.
static double radius_threshold(double R) {
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

static int alloc_agent_arrays(rcs_sim* s, AgentArrays& a, uint64_t cap) {
  // two spare rows: the bulk-copy staging of step_tile_kernel rounds a row range up to 16 bytes
  CU_TRY(s, dalloc(&a.pos, cap + 2));
  CU_TRY(s, dalloc(&a.vel, cap + 2));
  CU_TRY(s, dalloc(&a.id, cap + 2));
  CU_TRY(s, dalloc(&a.grp, cap + 2));
  CU_TRY(s, dalloc(&a.wp, cap + 2));
  a.pv = nullptr;
  return RCS_OK;
}

static void free_agent_arrays(AgentArrays& a) {
  cudaFree(a.pos); cudaFree(a.vel);
  cudaFree(a.id); cudaFree(a.grp); cudaFree(a.wp); cudaFree(a.pv);
  a = AgentArrays{};
}

static int ensure_stage(rcs_sim* s, uint64_t bytes) {
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
// Launch on the handle's stream -- with RCS_OPT_PDL as a programmatic dependent of the kernel enqueued before it: its
// blocks may be dispatched while that kernel still runs, and the kernel's first statement, pdl_enter(), holds them
// until it has completed (rcs_kernels.cuh).  Only kernels that start with pdl_enter() are launched through here.
template <class... P, class... A>
static void launch_dep(rcs_sim* s, void (*kernel)(P...), dim3 grid, dim3 block, size_t smem, A&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s->stream;
  cudaLaunchAttribute attr{};
  attr.id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr.val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = &attr;
  cfg.numAttrs = s->opt_pdl ? 1u : 0u;
  cudaLaunchKernelEx(&cfg, kernel, static_cast<P>(args)...);
}

static int exclusive_scan(rcs_sim* s, const uint32_t* in, uint64_t len, uint32_t* out, uint32_t* cursor) {
  if (len == 0) {
    CU_TRY(s, cudaMemsetAsync(out, 0, sizeof(uint32_t), s->stream));
    return RCS_OK;
  }
  uint64_t tiles = (len + SCAN_TILE - 1) / SCAN_TILE;
  if (tiles > s->tile_sums_cap) {
    s->graph_epoch += 1;
    CU_TRY(s, cudaStreamSynchronize(s->stream));
    cudaFree(s->tile_sums);
    s->tile_sums = nullptr;
    CU_TRY(s, dalloc(&s->tile_sums, tiles + 1024));
    s->tile_sums_cap = tiles + 1024;
  }
  if (tiles <= 8192) {
    // the cell histogram of a step (4096 tiles at 2^24 cells): every block of the apply pass sums the tiles before
    // its own, a single tile needs no sums at all
    if (tiles > 1) launch_dep(s, scan_reduce_kernel, (uint32_t)tiles, SCAN_THREADS, 0, in, len, s->tile_sums);
    launch_dep(s, scan_apply_kernel<true>, (uint32_t)tiles, SCAN_THREADS, 0, in, len, s->tile_sums, out, cursor);
    s->launches += tiles > 1 ? 2 : 1;
  } else {
    launch_dep(s, scan_reduce_kernel, (uint32_t)tiles, SCAN_THREADS, 0, in, len, s->tile_sums);
    scan_tile_sums_kernel<<<1, SCAN_THREADS, 0, s->stream>>>(s->tile_sums, (uint32_t)tiles, s->scan_total);
    launch_dep(s, scan_apply_kernel<false>, (uint32_t)tiles, SCAN_THREADS, 0, in, len, s->tile_sums, out, cursor);
    s->launches += 3;
  }
  CU_TRY(s, cudaGetLastError());
  return RCS_OK;
}

static int upload_groups(rcs_sim* s) {
  if (!s->groups_dirty) return RCS_OK;
  s->graph_epoch += 1;
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

static uint32_t find_or_add_group(rcs_sim* s, uint32_t hl, uint32_t lp, double eyesight, int32_t source_sink) {
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
  g.route_off = H.route_off;
  g.route_n = H.route_n;
  if (H.kind == HL_ROUTE) s->any_route = true;
  g.lp_kind = L.kind;
  g.source_sink = source_sink;
  if (L.kind == LP_ZANLUNGO) {
    g.agent_scale = L.agent_scale;
    g.force_distance = L.force_distance;
    g.inv_mass = 1.0 / L.agent_mass;
    g.rr = L.agent_radius * L.agent_radius;
    g.two_r = L.agent_radius * 2.0;
    g.inv_fd = 1.0 / L.force_distance;
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

static int ensure_pref_arrays(rcs_sim* s) {
  if (s->cur.pv) return RCS_OK;
  s->graph_epoch += 1;
  CU_TRY(s, cudaStreamSynchronize(s->stream));
  const double nan = std::numeric_limits<double>::quiet_NaN();
  for (AgentArrays* a : {&s->cur, &s->srt}) {
    CU_TRY(s, dalloc(&a->pv, s->cap));
    fill_f64_kernel<<<blocks_for(2 * s->cap, 256), 256, 0, s->stream>>>(2 * s->cap, reinterpret_cast<double*>(a->pv), nan);
    s->launches += 1;
  }
  CU_TRY(s, cudaGetLastError());
  // pointer roles recorded for pending steps are stale now, but ensure_pref_arrays only runs right after a sync
  return RCS_OK;
}

static void invalidate(rcs_sim* s) {
  s->graph_epoch += 1;
  s->index_valid = false;
  s->binned_ahead = false;
  s->slot_valid = false;
  s->tr_valid = false;
}

// The host changed the agent count (add / remove / rollback): rewrite the device counters.
static int upload_counts(rcs_sim* s) {
  if (!s->cnt_dirty) return RCS_OK;
  s->graph_epoch += 1;
  set_counts_kernel<<<1, 1, 0, s->stream>>>(s->cnt, s->n);
  s->launches += 1;
  CU_TRY(s, cudaGetLastError());
  s->cnt_dirty = false;
  return RCS_OK;
}

// A1/A2 of SURVEY.md section 8a.  Bins the agents [first, cnt[CNT_TOT]) of `cur` into the histogram
// (which the caller has zeroed, or which already holds the owned agents of a strip).
static int bin_agents(rcs_sim* s, uint32_t n_ub, const uint32_t* first, uint32_t launch_n = 0, bool pack = false,
                      bool pack_only = false) {
  if (!n_ub) return RCS_OK;
  if (!launch_n) launch_n = n_ub;  // threads to launch: fewer than n_ub when only the tail behind *first is binned
  PackArgs pk{};
  if (pack) {  // strips, owned pass: the boundary columns go to the send buffers on the way
    pk.enabled = 1u;
    pk.nx = (uint32_t)s->grid.nx;
    pk.width = s->halo_width;
    pk.has_left = s->rank > 0;
    pk.has_right = s->rank + 1 < s->world;
    pk.st = s->strip;
    pk.left = s->send_l.buf;
    pk.right = s->send_r.buf;
    pk.cur = s->cur;
    if (s->peer.enabled) {  // rows go straight into the neighbours' receive buffers; the counts stay local
      pk.left = s->peer.remote[0];
      pk.right = s->peer.remote[1];
      pk.left.count = s->send_l.buf.count;
      pk.right.count = s->send_r.buf.count;
      pk.xseq = s->peer.xseq;
    }
  }
  if (pack_only) {  // the owned agents were binned by the previous step's epilogue
    halo_pack_kernel<<<blocks_for(launch_n, 256), 256, 0, s->stream>>>(n_ub, s->cnt + CNT_TOT, s->cellid, pk, s->d_status);
    s->launches += 1;
    CU_TRY(s, cudaGetLastError());
    return RCS_OK;
  }
  const uint32_t* dead = s->cur_has_dead ? s->keep : nullptr;
  if (pack)
    launch_dep(s, bin_count_kernel<true>, blocks_for(launch_n, BIN_THREADS), BIN_THREADS, 0, 
        s->grid, n_ub, first, s->cnt + CNT_TOT, s->cur.pos, dead, s->cellid, s->cell_count, s->cell_lo, s->cell_hi, pk,
        s->d_status);
  else
    launch_dep(s, bin_count_kernel<false>, blocks_for(launch_n, BIN_THREADS), BIN_THREADS, 0, 
        s->grid, n_ub, first, s->cnt + CNT_TOT, s->cur.pos, dead, s->cellid, s->cell_count, s->cell_lo, s->cell_hi, pk,
        s->d_status);
  s->launches += 1;
  CU_TRY(s, cudaGetLastError());
  return RCS_OK;
}

// Histogram -> exclusive scan -> permutation scatter -> ascending-id order inside each cell -> physical
// reorder of `cur` into `srt`.  cell_start[len] is the number of sorted (in-bounds) agents.
static const uint32_t* n_sorted_ptr(const rcs_sim* s) { return s->cell_start + s->cell_hi; }

static int clear_histogram(rcs_sim* s) {
  CU_TRY(s, cudaMemsetAsync(s->cell_count + s->cell_lo, 0, (s->cell_hi - s->cell_lo + 1) * sizeof(uint32_t), s->stream));
  return RCS_OK;
}

static int sort_into_srt(rcs_sim* s, uint32_t n_ub, bool clear_after_scan = false) {
  const uint64_t lo = s->cell_lo, hi = s->cell_hi, len = hi - lo;
  int rc = exclusive_scan(s, s->cell_count + lo, len, s->cell_start + lo, s->cursor + lo);
  if (rc) return rc;
  if (clear_after_scan) {  // the step's epilogue refills the histogram for the next step
    rc = clear_histogram(s);
    if (rc) return rc;
  }
  if (n_ub) {
    launch_dep(s, scatter_perm_kernel, blocks_for(n_ub, 256), 256, 0, n_ub, s->cnt + CNT_TOT, s->cellid, s->cursor,
                                                                      s->perm, s->d_status);
    if (len)
      launch_dep(s, sort_cells_by_id_kernel, blocks_for(len, 128), 128, 0, lo, hi, s->cell_start, s->cur.id, s->perm,
                                                                           s->big_list, 4096, s->d_status);
    launch_dep(s, sort_big_cells_kernel, 148, 1024, 0, lo, hi, s->cell_start, s->cur.id, s->perm, s->slow_list,
                                                       s->wide_list,
                                                       s->big_list, 4096, s->d_status);
    launch_dep(s, gather_sorted_kernel, blocks_for(n_ub, GATHER_THREADS), GATHER_THREADS, 0, 
        n_ub, s->perm, s->cur, s->srt, s->cellid, s->strip.enabled ? s->srt_cell : nullptr, n_sorted_ptr(s),
        s->grid, s->cell_start, s->d_groups, s->d_groups ? s->slices : nullptr,
        s->d_groups ? s->tile_ranges : nullptr, s->d_status);
    s->launches += 4;
  }
  CU_TRY(s, cudaGetLastError());
  return RCS_OK;
}

// (Re)build the canonical (cell, id) sorted copy `srt` of `cur` and cell_start (no step in flight).
static int build_index(rcs_sim* s) {
  s->binned_ahead = false;  // the histogram is rebuilt (and consumed) here
  int rc = upload_groups(s);  // gather_sorted_kernel reads the group table
  if (rc) return rc;
  rc = upload_counts(s);
  if (rc) return rc;
  rc = clear_histogram(s);
  if (rc) return rc;
  rc = bin_agents(s, s->n, nullptr);
  if (rc) return rc;
  return sort_into_srt(s, s->n);
}

static cudaEvent_t kevent_get(rcs_sim* s) {
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
static void launch_step_kernel(rcs_sim* s, const StepArgs& a, uint32_t n_ub, bool sorted_input) {
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  if (s->ktiming) {
    e0 = kevent_get(s);
    e1 = kevent_get(s);
    cudaEventRecord(e0, s->stream);
  }
  if (sorted_input && (s->opt_step_kernel == 0 || s->opt_step_kernel == 3)) {
    // stencil staged in shared memory (rcs_step_tile.cuh); agents it cannot stage go to step_aside_kernel's lists
    if (!s->tile_attr_set) {  // per device context
      cudaFuncSetAttribute(step_tile_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(TileShared));
      cudaFuncSetAttribute(step_tile_kernel<false>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
      cudaFuncSetAttribute(step_tile_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(TileShared));
      cudaFuncSetAttribute(step_tile_kernel<true>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
      s->tile_attr_set = true;
    }
    if (a.strip.enabled)
      launch_dep(s, step_tile_kernel<true>, blocks_for(n_ub, 32 * ST_WARPS), 32 * ST_WARPS, sizeof(TileShared), a);
    else
      launch_dep(s, step_tile_kernel<false>, blocks_for(n_ub, 32 * ST_WARPS), 32 * ST_WARPS, sizeof(TileShared), a);
    launch_dep(s, step_aside_kernel, 148 * 4, 32 * SW_WARPS, 0, a);
    s->launches += 2;
  } else if (sorted_input && s->opt_step_kernel != 1) {
    step_warp_kernel<<<blocks_for(n_ub, 32 * SW_WARPS), 32 * SW_WARPS, 0, s->stream>>>(a);
    // the agents it left aside (device-side lists): stencils wider than three columns or crowded columns go to
    // the chunked cooperative kernel, ids >= 2^53 and planners without the weight-0 proof to the sequential one
    launch_dep(s, step_aside_kernel, 148 * 4, 32 * SW_WARPS, 0, a);
    s->launches += 2;
  } else if (!sorted_input && s->opt_step_kernel != 1) {
    step_stream_kernel<<<blocks_for((n_ub + 1) / 2, 256), 256, 0, s->stream>>>(a);  // NoLocalPlan only, no churn
    s->launches += 1;
  } else {
    step_kernel<<<blocks_for(n_ub, 128), 128, 0, s->stream>>>(a);
    s->launches += 1;
  }
  if (s->ktiming) {
    cudaEventRecord(e1, s->stream);
    s->kevents.push_back({e0, e1});
  }
}

static int drain_kevents(rcs_sim* s) {
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

// Undo a failed step on a strip: the snapshot in `cur` (sorted, ghosts included) is reduced to the agents
// this rank owns.
static int rollback_owned(rcs_sim* s, uint32_t n_tot) {
  const uint32_t* n_sorted = n_sorted_ptr(s);
  role_keep_kernel<<<blocks_for(n_tot, 256), 256, 0, s->stream>>>(n_tot, n_sorted, s->srt_cell, (uint32_t)s->grid.nx,
                                                                  s->strip, s->cellid);
  s->launches += 1;
  int rc = exclusive_scan(s, s->cellid, n_tot, s->perm, nullptr);
  if (rc) return rc;
  compact_keep_kernel<<<blocks_for(n_tot, 256), 256, 0, s->stream>>>(n_tot, n_sorted, s->cellid, s->perm, s->cur,
                                                                     s->srt, nullptr);
  s->launches += 1;
  uint32_t kept = 0;
  CU_TRY(s, cudaMemcpyAsync(&kept, s->perm + n_tot, sizeof(uint32_t), cudaMemcpyDeviceToHost, s->stream));
  CU_TRY(s, cudaStreamSynchronize(s->stream));
  std::swap(s->cur, s->srt);
  s->n = kept;
  return RCS_OK;
}

// Drop the entries flagged keep = 0 from `cur` (stream compaction: scan of the flags, then a gather).  Steps do
// not need this -- the next step's counting sort skips those entries -- so it only runs when the host looks.
static int compact_cur(rcs_sim* s) {
  if (!s->cur_has_dead) return RCS_OK;
  s->binned_ahead = false;  // entries move: cellid[] no longer matches
  const uint32_t n_ub = s->n;  // entries in cur, dead ones included
  uint32_t kept = 0;
  if (n_ub) {
    int rc = exclusive_scan(s, s->keep, n_ub, s->perm, nullptr);
    if (rc) return rc;
    compact_keep_kernel<<<blocks_for(n_ub, 256), 256, 0, s->stream>>>(n_ub, s->cnt + CNT_CUR, s->keep, s->perm, s->cur,
                                                                      s->srt, nullptr);
    s->launches += 1;
    CU_TRY(s, cudaMemcpyAsync(&kept, s->perm + n_ub, sizeof(uint32_t), cudaMemcpyDeviceToHost, s->stream));
    CU_TRY(s, cudaStreamSynchronize(s->stream));
    std::swap(s->cur, s->srt);
  }
  s->n = kept;
  s->n_ub = kept;
  s->cnt_dirty = true;
  s->cur_has_dead = false;
  s->stats.n_agents = kept;
  return RCS_OK;
}

static int do_sync(rcs_sim* s) {
  CU_TRY(s, cudaMemcpyAsync(s->h_status, s->d_status, sizeof(DevStatus), cudaMemcpyDeviceToHost, s->stream));
  CU_TRY(s, cudaMemcpyAsync(s->h_cnt, s->cnt, CNT_N * sizeof(uint32_t), cudaMemcpyDeviceToHost, s->stream));
  unsigned long long steps_done = 0;
  CU_TRY(s, cudaMemcpyAsync(&steps_done, s->d_steps_done, sizeof(steps_done), cudaMemcpyDeviceToHost, s->stream));
  unsigned long long next_id = s->last_alloc_agent_id;
  if (s->ever_had_sources)
    CU_TRY(s, cudaMemcpyAsync(&next_id, s->d_next_id, sizeof(next_id), cudaMemcpyDeviceToHost, s->stream));
  CU_TRY(s, cudaStreamSynchronize(s->stream));
  const DevStatus& st = *s->h_status;
  if (!s->cnt_dirty && !s->pending.empty()) s->n = s->h_cnt[CNT_CUR];
  if (next_id > s->last_alloc_agent_id) {  // source sinks allocated ids on the device (lib.rs:128-129)
    s->last_alloc_agent_id = next_id;
    s->max_id_plus1 = std::max<uint64_t>(s->max_id_plus1, next_id);
  }
  s->n_ub = s->n;
  s->stats.oob_count = st.oob_count;
  s->stats.first_oob_id = st.first_oob_id;
  s->stats.nonfinite_count = st.nonfinite_count;
  s->stats.finite_tti_count = st.finite_tti;
  s->stats.neighbour_total = st.neighbour_total;
  s->stats.candidate_total = st.candidate_total;
  s->stats.spawned = st.spawned;
  s->stats.destroyed = st.destroyed;
  s->stats.n_agents = s->n;
  s->stats.steps = steps_done;
  int rc = drain_kevents(s);
  if (rc) return rc;
  if (st.failed) {
    // the step with index k (since the last sync) failed; every later one was skipped on the device
    uint64_t k = steps_done - s->steps_done_at_sync;
    // the sorted snapshot of the failed step holds cell_start[len] entries (all live: the sort dropped the rest)
    uint32_t n_snapshot = 0;
    CU_TRY(s, cudaMemcpy(&n_snapshot, n_sorted_ptr(s), sizeof(uint32_t), cudaMemcpyDeviceToHost));
    if (k < s->pending.size()) {
      const PendingStep& p = s->pending[k];
      if (p.snapshot_in_srt) {
        s->cur = p.srt;
        s->srt = p.cur;
      } else {
        s->cur = p.cur;
        s->srt = p.srt;
        n_snapshot = p.n;
      }
    }
    if (st.local_failed && st.oob_count) {
      s->err = "Index out of bounds";
      rc = RCS_ERR_OUT_OF_BOUNDS;
    } else if (st.local_failed && st.capacity_err) {
      s->err = "capacity exceeded (agents, halo or event buffers)";
      rc = RCS_ERR_CAPACITY;
    } else if (st.local_failed && st.halo_err) {
      s->err = "an agent moved further than the halo ring in one step";
      rc = RCS_ERR_HALO;
    } else {
      s->err = "a neighbouring rank failed its step";
      rc = RCS_ERR_HALO;
    }
    // the snapshot holds the agents that were spawned at the start of the failed step (their events stand);
    // destroy events of the failed step are dropped
    s->n = n_snapshot;
    s->h_cnt[CNT_EV_DESTROY] = s->h_cnt[CNT_EV_DESTROY_SAVE];
    CU_TRY(s, cudaMemcpyAsync(s->cnt + CNT_EV_DESTROY, s->h_cnt + CNT_EV_DESTROY_SAVE, sizeof(uint32_t),
                              cudaMemcpyHostToDevice, s->stream));
    CU_TRY(s, cudaMemsetAsync(&s->d_status->failed, 0, sizeof(unsigned int), s->stream));
    CU_TRY(s, cudaStreamSynchronize(s->stream));
    if (s->strip.enabled && k < s->pending.size() && s->pending[k].snapshot_in_srt) {
      int rc2 = rollback_owned(s, n_snapshot);
      if (rc2) return rc2;
    }
    s->n_ub = s->n;
    s->cnt_dirty = true;
    s->cur_has_dead = false;
    s->stats.n_agents = s->n;
    invalidate(s);
  }
  s->pending.clear();
  s->steps_done_at_sync = steps_done;
  if (rc == RCS_OK && s->cur_has_dead) rc = compact_cur(s);
  return rc;
}

static int build_slot_table(rcs_sim* s) {
  if (s->slot_valid) return RCS_OK;
  if (s->strip.enabled && s->n) {
    // agents migrate between ranks: the id range of this handle is whatever currently lives here
    unsigned long long* d_max = reinterpret_cast<unsigned long long*>(s->d_bad2);
    CU_TRY(s, cudaMemsetAsync(d_max, 0, sizeof(unsigned long long), s->stream));
    max_id_kernel<<<blocks_for(s->n, 256), 256, 0, s->stream>>>(s->n, s->cur.id, d_max);
    s->launches += 1;
    unsigned long long mx = 0;
    CU_TRY(s, cudaMemcpyAsync(&mx, d_max, sizeof(mx), cudaMemcpyDeviceToHost, s->stream));
    CU_TRY(s, cudaStreamSynchronize(s->stream));
    s->max_id_plus1 = std::max<uint64_t>(s->max_id_plus1, mx + 1);
  }
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

// src: first element of the component to read; stride 2 for one component of a double2 array
template <class T>
static int read_array(rcs_sim* s, const T* src, int stride, const uint32_t* order, uint32_t n, T* host_out,
                      uint64_t stage_off) {
  T* st = reinterpret_cast<T*>(static_cast<char*>(s->stage) + stage_off);
  gather_kernel<T><<<blocks_for(n, 256), 256, 0, s->stream>>>(n, order, src, stride, st);
  s->launches += 1;
  CU_TRY(s, cudaMemcpyAsync(host_out, st, (uint64_t)n * sizeof(T), cudaMemcpyDeviceToHost, s->stream));
  return RCS_OK;
}

}  // namespace rcs_host
