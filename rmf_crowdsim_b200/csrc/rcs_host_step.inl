// rcs_host_step.inl -- Simulation::step (lib.rs:195-383) as a kernel pipeline, source sinks and events.
// Part of rcs.cu (single translation unit).
//
// One step on the sorted path (any Zanlungo group, tracing, source sinks or strips):
//
//   phase A   begin_step                              reset per-step counters
//             ss_probe + ss_spawn          [sources]  lib.rs:199-254: spawn at most one agent per source
//             bin owned (+ halo pack)       [strips]  boundary columns -> send buffers, fused into the binning pass
//   exchange  NCCL send/recv | peer copies  [strips]
//   phase B   halo unpack + bin ghosts      [strips]  ghosts appended behind the owned agents
//             bin -> scan -> scatter -> sort cells by id -> gather      LocationHash2D rebuild (A1, A2)
//             step_warp (+ step_aside)                radius query + Zanlungo + Euler (A3-A11), keep flags (A12)
//             end_step                                out of bounds / halo / capacity => the step does not stand;
//                                                     [churn] leaving agents are only FLAGGED (keep = 0): the next
//                                                     step's counting sort drops them, rcs_sync compacts on demand
//
// NoLocalPlan-only crowds without churn skip the index: the step is one streaming kernel in storage order.

namespace rcs_host {

static int upload_sources(rcs_sim* s) {
  if (!s->sources_dirty) return RCS_OK;
  s->graph_epoch += 1;
  CU_TRY(s, cudaStreamSynchronize(s->stream));
  const size_t ns = s->sources.size();
  cudaFree(s->d_sources); cudaFree(s->d_ss_wp); cudaFree(s->d_blocked); cudaFree(s->d_sg_start); cudaFree(s->d_sg_items);
  s->d_sources = nullptr; s->d_ss_wp = nullptr; s->d_blocked = nullptr; s->d_sg_start = nullptr; s->d_sg_items = nullptr;
  CU_TRY(s, dalloc(&s->d_sources, ns));
  CU_TRY(s, dalloc(&s->d_ss_wp, s->ss_wp.size()));
  CU_TRY(s, dalloc(&s->d_blocked, ns));
  CU_TRY(s, cudaMemcpy(s->d_sources, s->sources.data(), ns * sizeof(SourceSinkDev), cudaMemcpyHostToDevice));
  CU_TRY(s, cudaMemcpy(s->d_ss_wp, s->ss_wp.data(), s->ss_wp.size() * sizeof(double), cudaMemcpyHostToDevice));
  CU_TRY(s, cudaMemset(s->d_blocked, 0, std::max<size_t>(ns, 1) * sizeof(uint32_t)));
  if (s->strip.enabled) {
    cudaFree(s->d_ss_bits_local); cudaFree(s->d_ss_bits_parts); cudaFree(s->d_ss_bits);
    s->d_ss_bits_local = s->d_ss_bits_parts = s->d_ss_bits = nullptr;
    s->ss_words = (uint32_t)((ns + 31) / 32);
    CU_TRY(s, dalloc(&s->d_ss_bits_local, s->ss_words));
    CU_TRY(s, dalloc(&s->d_ss_bits, s->ss_words));
    CU_TRY(s, dalloc(&s->d_ss_bits_parts, (uint64_t)s->ss_words * std::max(s->world, 1)));
  }
  // lookup grid over the alive sources (cells >= 1 m, at most 1024 x 1024 of them)
  double x0 = 0, y0 = 0, x1 = 0, y1 = 0;
  bool any = false;
  for (const SourceSinkDev& q : s->sources) {
    if (!q.alive || !std::isfinite(q.sx) || !std::isfinite(q.sy)) continue;
    if (!any) { x0 = x1 = q.sx; y0 = y1 = q.sy; any = true; }
    x0 = std::min(x0, q.sx); x1 = std::max(x1, q.sx);
    y0 = std::min(y0, q.sy); y1 = std::max(y1, q.sy);
  }
  SourceGridDev sg{};
  sg.cell = std::max({1.0, (x1 - x0) / 1024.0, (y1 - y0) / 1024.0});
  sg.x0 = x0 - 0.5 * sg.cell;
  sg.y0 = y0 - 0.5 * sg.cell;
  sg.nx = (uint32_t)std::floor((x1 - sg.x0) / sg.cell) + 1;
  sg.ny = (uint32_t)std::floor((y1 - sg.y0) / sg.cell) + 1;
  std::vector<uint32_t> start((size_t)sg.nx * sg.ny + 1, 0u), items;
  auto cell_of = [&](const SourceSinkDev& q) {
    uint32_t cx = (uint32_t)std::floor((q.sx - sg.x0) / sg.cell), cy = (uint32_t)std::floor((q.sy - sg.y0) / sg.cell);
    return (size_t)std::min(cx, sg.nx - 1) * sg.ny + std::min(cy, sg.ny - 1);
  };
  for (const SourceSinkDev& q : s->sources)
    if (q.alive && std::isfinite(q.sx) && std::isfinite(q.sy)) start[cell_of(q) + 1]++;
  for (size_t c = 0; c + 1 < start.size(); ++c) start[c + 1] += start[c];
  items.resize(start.back());
  std::vector<uint32_t> cursor(start.begin(), start.end() - 1);
  for (size_t k = 0; k < ns; ++k) {
    const SourceSinkDev& q = s->sources[k];
    if (q.alive && std::isfinite(q.sx) && std::isfinite(q.sy)) items[cursor[cell_of(q)]++] = (uint32_t)k;
  }
  CU_TRY(s, dalloc(&s->d_sg_start, start.size()));
  CU_TRY(s, dalloc(&s->d_sg_items, items.size()));
  CU_TRY(s, cudaMemcpy(s->d_sg_start, start.data(), start.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
  if (!items.empty())
    CU_TRY(s, cudaMemcpy(s->d_sg_items, items.data(), items.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
  sg.start = s->d_sg_start;
  sg.items = s->d_sg_items;
  s->sgrid = sg;
  s->sources_dirty = false;
  return RCS_OK;
}

static StepArgs make_step_args(rcs_sim* s, const AgentArrays& in, const AgentArrays& out, double dt, uint32_t n_ub,
                               bool full_out, uint32_t flags) {
  StepArgs a{};
  a.grid = s->grid;
  a.n = n_ub;
  a.n_sorted = n_sorted_ptr(s);
  a.in = in;
  a.cell_start = s->cell_start;
  a.groups = s->d_groups;
  a.dt = dt;
  a.opos = out.pos;
  a.ovel = out.vel;
  if (full_out) {
    a.oid = out.id;
    a.ogrp = out.grp;
    a.owp = out.wp;
    a.opv = in.pv ? out.pv : nullptr;
  }
  a.status = s->d_status;
  a.collect_stats = 1;
  a.no_commit = (flags & RCS_STEP_NO_COMMIT) ? 1u : 0u;
  a.steps_done = s->d_steps_done;
  if (s->trace) {
    a.t_i = s->tr_ti;
    a.fx = s->tr_fx;
    a.fy = s->tr_fy;
    a.nb_count = s->tr_nbc;
    a.tr_id = s->tr_id;
    a.tr_own = s->tr_own;
  }
  a.cnt = s->cnt;
  a.slow_list = s->slow_list;
  a.wide_list = s->wide_list;
  a.slices = full_out ? s->slices : nullptr;
  a.tile_ranges = full_out ? s->tile_ranges : nullptr;
  if (full_out && !s->trace && s->opt_bin_ahead) {  // the epilogue bins for the next step
    a.next_cellid = s->cellid;
    a.next_count = s->cell_count;
    a.cell_lo = s->cell_lo;
    a.cell_hi = s->cell_hi;
  }
  a.routes = s->d_routes;
  a.route_thr2 = radius_threshold(1e-1);
  if (full_out && churn(s)) {
    a.keep = s->keep;
    a.cell = s->srt_cell;
    a.strip = s->strip;
    if (s->ever_had_sources) {
      a.ss = s->d_sources;
      a.ss_wp = s->d_ss_wp;
      a.ev_destroyed = s->ev_destroyed;
      a.ev_cap = s->ev_cap;
    }
  }
  return a;
}

static int strip_halo_width(rcs_sim* s);  // rcs_host_dist.inl

static bool sorted_path(const rcs_sim* s) { return s->any_zanlungo || s->any_route || s->trace || churn(s); }

// ---- phase A ------------------------------------------------------------------------------------
static int upload_routes(rcs_sim* s) {
  if (!s->routes_dirty) return RCS_OK;
  s->graph_epoch += 1;
  CU_TRY(s, cudaStreamSynchronize(s->stream));
  cudaFree(s->d_routes);
  s->d_routes = nullptr;
  CU_TRY(s, dalloc(&s->d_routes, s->routes.size()));
  CU_TRY(s, cudaMemcpy(s->d_routes, s->routes.data(), s->routes.size() * sizeof(double), cudaMemcpyHostToDevice));
  s->routes_dirty = false;
  return RCS_OK;
}

// Upper bound on the entries of `cur` after this step's spawns (at most one per source, lib.rs:207-219).  The exact
// bound grows by the number of sources every step; the launch bound derived from it moves in coarse granules, so
// that consecutive steps launch the same grids (and can replay one CUDA graph).
static uint32_t spawn_launch_bound(rcs_sim* s) {
  // n_ub was reset from an exact count (sync, add, remove) -- or happens to equal the last coarse bound
  if (s->n_ub != s->n_coarse || s->n_tight < s->n) s->n_tight = std::max(s->n_ub, s->n);
  s->n_tight = (uint32_t)std::min<uint64_t>(s->cap, (uint64_t)s->n_tight + s->n_sources_alive);
  const uint64_t gran = std::max<uint64_t>(32ull * s->n_sources_alive, 4096);
  s->n_coarse = (uint32_t)std::min<uint64_t>(s->cap, ((uint64_t)s->n_tight + gran - 1) / gran * gran);
  return s->n_coarse;
}

static int step_spawn_set_nccl(rcs_sim* s);  // rcs_host_dist.inl
static void strip_range(const rcs_sim* s, int rank, int world, uint64_t& c0, uint64_t& c1);

// Phase A, first half: counters reset; with source sinks the spawn probe (lib.rs:212-214) -- and on a strip this
// rank's part of the step's spawn set, which is then summed over the ranks (step_spawn_set_nccl or the peer copies of
// rcs_dist_step_local) before the second half.
static int step_phase_a1(rcs_sim* s, double dt) {
  int rc = upload_groups(s);
  if (rc) return rc;
  rc = upload_routes(s);
  if (rc) return rc;
  rc = upload_sources(s);
  if (rc) return rc;
  rc = upload_counts(s);
  if (rc) return rc;
  // single-process transport: the neighbours must have fetched the previous step's send buffers before their
  // headers are reset (with NCCL the sends are ordered on this very stream)
  for (rcs_sim* nb : s->local_group)
    if (nb && nb != s && std::abs(nb->rank - s->rank) == 1 && nb->ev_copied)
      CU_TRY(s, cudaStreamWaitEvent(s->stream, nb->ev_copied, 0));
  launch_dep(s, begin_step_kernel, 1, 1, 0, s->d_status, s->cnt, s->strip.enabled ? s->send_l.buf.count : nullptr,
                                            s->strip.enabled ? s->send_r.buf.count : nullptr,
                                            s->peer.enabled ? s->peer.xseq : nullptr);
  s->launches += 1;
  if (s->n_sources_alive) {
    const uint32_t n_before = s->strip.enabled ? (uint32_t)s->cap : s->n_ub;
    // On a strip the probe looks at the owned agents only: a source sink is accepted only where its whole probe
    // stencil lies in its owner's columns (rcs_add_source_sink), and a rank owns exactly the agents in its columns.
    if (n_before)
      ss_probe_kernel<<<blocks_for(n_before, 256), 256, 0, s->stream>>>(
          s->grid, s->sgrid, s->d_sources, radius_threshold(0.4), n_before, s->cnt + CNT_CUR, s->cur.pos,
          s->cur_has_dead ? s->keep : nullptr, s->d_blocked, s->d_status);
    s->launches += 1;
    if (s->strip.enabled) {
      CU_TRY(s, cudaMemsetAsync(s->d_ss_bits_local, 0, s->ss_words * sizeof(uint32_t), s->stream));
      ss_flags_kernel<<<blocks_for(s->sources.size(), 256), 256, 0, s->stream>>>(
          s->d_sources, (uint32_t)s->sources.size(), dt, s->d_blocked, s->d_ss_bits_local, s->d_status);
      s->launches += 1;
    }
  }
  CU_TRY(s, cudaGetLastError());
  return RCS_OK;
}

// Phase A, second half: spawn at most one agent per source (lib.rs:199-254); strips: bin the owned agents and pack
// the boundary columns for the neighbours.
static int step_phase_a2(rcs_sim* s, double dt) {
  int rc = RCS_OK;
  if (s->n_sources_alive) {
    ss_spawn_kernel<<<1, SCAN_THREADS, 0, s->stream>>>(s->grid, s->d_sources, s->d_groups, (uint32_t)s->sources.size(), dt,
                                                       s->strip.enabled ? s->d_ss_bits : nullptr,
                                                       s->d_blocked, s->cur, s->keep, (uint32_t)s->cap, s->cnt, s->d_next_id,
                                                       s->ev_spawn_id, s->ev_spawn_xy, s->ev_cap, s->d_status);
    s->launches += 1;
    s->n_ub = spawn_launch_bound(s);
  }
  if (s->strip.enabled) {
    s->n_ub = (uint32_t)s->cap;
    rc = strip_halo_width(s);  // also narrows [cell_lo, cell_hi) to the strip and its halo
    if (rc) return rc;
    if (s->binned_ahead) {
      rc = bin_agents(s, s->n_ub, nullptr, 0, true, true);  // binned by the last step's epilogue: halo pack only
      if (rc) return rc;
      if (s->n_sources_alive)  // ... except this step's spawns, behind the pre-spawn count
        rc = bin_agents(s, s->n_ub, s->cnt + CNT_SAVE, s->n_sources_alive, true);
    } else {
      rc = clear_histogram(s);
      if (rc) return rc;
      rc = bin_agents(s, s->n_ub, nullptr, 0, true);  // owned agents: histogram + halo pack in one pass
    }
    if (rc) return rc;
    if (s->ev_packed) CU_TRY(s, cudaEventRecord(s->ev_packed, s->stream));
  }
  CU_TRY(s, cudaGetLastError());
  return RCS_OK;
}

static int step_phase_a(rcs_sim* s, double dt) {
  int rc = step_phase_a1(s, dt);
  if (rc) return rc;
  if (s->strip.enabled && s->n_sources_alive) {
    rc = step_spawn_set_nccl(s);
    if (rc) return rc;
  }
  return step_phase_a2(s, dt);
}

// ---- exchange: NCCL point-to-point with the two strip neighbours -------------------------------------
static int step_exchange_nccl(rcs_sim* s);
// ---- exchange: single-process transport (all ranks' handles live in this process) ---------------------
static int step_exchange_local(rcs_sim* s) {
  for (int side = 0; side < 2; ++side) {
    const int nb_rank = side == 0 ? s->rank - 1 : s->rank + 1;
    if (nb_rank < 0 || nb_rank >= s->world) continue;
    rcs_sim* nb = s->local_group[nb_rank];
    CU_TRY(s, cudaStreamWaitEvent(s->stream, nb->ev_packed, 0));
    const HaloMem& src = side == 0 ? nb->send_r : nb->send_l;
    const HaloMem& dst = side == 0 ? s->recv_l : s->recv_r;
    CU_TRY(s, cudaMemcpyAsync(dst.base, src.base, src.bytes, cudaMemcpyDefault, s->stream));
  }
  CU_TRY(s, cudaEventRecord(s->ev_copied, s->stream));
  return RCS_OK;
}

// ---- phase B ------------------------------------------------------------------------------------
static int step_phase_b(rcs_sim* s, double dt, uint32_t flags) {
  const bool no_commit = (flags & RCS_STEP_NO_COMMIT) != 0;
  const uint32_t n_ub = s->n_ub;
  PendingStep p{s->cur, s->srt, true, s->n};
  int rc = RCS_OK;
  bool churned = false;
  if (n_ub && sorted_path(s)) {
    if (s->strip.enabled) {
      const int has_l = s->rank > 0, has_r = s->rank + 1 < s->world;
      const HaloBuf& rl = s->peer.enabled ? s->peer.local[0] : s->recv_l.buf;
      const HaloBuf& rr = s->peer.enabled ? s->peer.local[1] : s->recv_r.buf;
      const uint32_t ghosts_ub = rl.cap + rr.cap;
      launch_dep(s, halo_unpack_kernel, blocks_for(ghosts_ub, 256), 256, 0, 
          s->cur, s->keep, (uint32_t)s->cap, rl, rr, has_l, has_r, s->cnt, s->d_status,
          s->peer.enabled ? s->peer.xseq : nullptr, s->grid, s->cellid, s->cell_count, s->cell_lo, s->cell_hi,
          s->send_l.buf.count, s->send_r.buf.count, s->peer.remote_hdr[0], s->peer.remote_hdr[1]);
      s->launches += 1;  // (the ghosts are binned on the way: no second pass over them)
    } else if (!s->binned_ahead) {
      rc = clear_histogram(s);
      if (rc) return rc;
      rc = bin_agents(s, n_ub, nullptr);
    } else if (s->n_sources_alive) {
      // binned by the last step's epilogue; only this step's spawns (behind the pre-spawn count) are new
      rc = bin_agents(s, n_ub, s->cnt + CNT_SAVE, s->n_sources_alive);
    }
    if (rc) return rc;
    const bool bin_ahead = !s->trace && s->opt_bin_ahead;
    rc = sort_into_srt(s, n_ub, bin_ahead);
    if (rc) return rc;
    if (s->trace) {
      CU_TRY(s, cudaMemsetAsync(s->tr_nbc, 0, (uint64_t)n_ub * sizeof(uint32_t), s->stream));
      CU_TRY(s, cudaMemsetAsync(s->tr_own, 0, (uint64_t)n_ub * sizeof(uint32_t), s->stream));
    }
    StepArgs a = make_step_args(s, s->srt, s->cur, dt, n_ub, true, flags);
    launch_step_kernel(s, a, n_ub, true);
    CU_TRY(s, cudaGetLastError());
    if (s->trace) {
      // neighbour lists of this step (debug path: synchronous), read from the sorted snapshot in srt
      rc = exclusive_scan(s, s->tr_nbc, n_ub, s->tr_nbo, nullptr);
      if (rc) return rc;
      uint32_t total = 0, n_sorted = 0;
      CU_TRY(s, cudaMemcpyAsync(&total, s->tr_nbo + n_ub, sizeof(uint32_t), cudaMemcpyDeviceToHost, s->stream));
      CU_TRY(s, cudaMemcpyAsync(&n_sorted, n_sorted_ptr(s), sizeof(uint32_t), cudaMemcpyDeviceToHost,
                                s->stream));
      CU_TRY(s, cudaStreamSynchronize(s->stream));
      if (total > s->tr_nbids_cap) {
        cudaFree(s->tr_nbids);
        s->tr_nbids = nullptr;
        CU_TRY(s, dalloc(&s->tr_nbids, (uint64_t)total + total / 4 + 1024));
        s->tr_nbids_cap = (uint64_t)total + total / 4 + 1024;
      }
      StepArgs t = make_step_args(s, s->srt, s->srt, dt, n_ub, true, flags);
      trace_neighbours_kernel<<<blocks_for(n_ub, 128), 128, 0, s->stream>>>(t, s->tr_nbo, s->tr_nbids);
      s->launches += 1;
      CU_TRY(s, cudaGetLastError());
      s->tr_nb_total = total;
      s->tr_n = std::min(n_sorted, n_ub);
    }
    if (churn(s)) {
      // The new state (or, without commit, the sorted snapshot) holds the agents that stay AND the ones that
      // leave, told apart by the keep flags the step kernel wrote.  Nothing is moved: the next step's counting
      // sort skips the flagged entries, and rcs_sync compacts when the host wants to look (compact_cur).
      if (no_commit) std::swap(s->cur, s->srt);
      churned = true;
    } else if (no_commit) {
      std::swap(s->cur, s->srt);  // the sorted pre-step snapshot becomes current again
    }
  } else if (n_ub) {
    // NoLocalPlan only: the radius query cannot influence the result (no_local_plan.rs:10-17), so the
    // step is a pure stream over the agents in storage order; new x,y,vx,vy go to the spare buffers.
    StepArgs a = make_step_args(s, s->cur, s->srt, dt, n_ub, false, flags);
    a.n_sorted = s->cnt + CNT_CUR;
    launch_step_kernel(s, a, n_ub, false);
    p.snapshot_in_srt = false;
    if (!no_commit) {
      std::swap(s->cur.pos, s->srt.pos);
      std::swap(s->cur.vel, s->srt.vel);
    }
  }
  launch_dep(s, end_step_kernel, 1, 1, 0, s->d_status, no_commit ? 0 : 1, s->d_steps_done,
                                          churned ? s->cnt + CNT_CUR : nullptr, n_sorted_ptr(s));
  s->launches += 1;
  CU_TRY(s, cudaGetLastError());
  if (churned) s->cur_has_dead = true;
  s->binned_ahead = n_ub && sorted_path(s) && !s->trace && s->opt_bin_ahead;
  s->pending.push_back(p);
  s->steps_enqueued += 1;
  s->index_valid = false;
  s->slot_valid = false;
  s->tr_valid = s->trace && n_ub && sorted_path(s);
  return RCS_OK;
}

}  // namespace rcs_host

extern "C" {

// One step, enqueued kernel by kernel on the handle's stream (which may be capturing).  `from_b`: phase A and the
// exchange have been enqueued already.
static int step_enqueue(rcs_sim* s, double dt, uint32_t flags, bool from_b = false) {
  if (!from_b) {
    int rc = rcs_host::step_phase_a(s, dt);
    if (rc) return rc;
    if (s->strip.enabled) {
      rc = rcs_host::step_exchange_nccl(s);
      if (rc) return rc;
    }
  }
  return rcs_host::step_phase_b(s, dt, flags);
}

static void step_graphs_clear(rcs_sim* s) {
  for (auto& g : s->step_graphs)
    if (g.exec) cudaGraphExecDestroy(g.exec);
  s->step_graphs.clear();
}

// The host-side half of a step whose device-side half is a graph launch: exactly what step_phase_a / step_phase_b
// change on the handle in the steady state the key describes (no uploads, no trace, no kernel timing).
static void step_replay_host(rcs_sim* s, uint32_t flags, uint64_t launches, bool from_b) {
  using namespace rcs_host;
  const bool no_commit = (flags & RCS_STEP_NO_COMMIT) != 0;
  if (!from_b) {
    if (s->n_sources_alive) s->n_ub = spawn_launch_bound(s);
    if (s->strip.enabled) s->n_ub = (uint32_t)s->cap;
  }
  const uint32_t n_ub = s->n_ub;
  PendingStep p{s->cur, s->srt, true, s->n};
  bool churned = false;
  if (n_ub && sorted_path(s)) {
    if (churn(s)) {
      if (no_commit) std::swap(s->cur, s->srt);
      churned = true;
    } else if (no_commit) {
      std::swap(s->cur, s->srt);
    }
  } else if (n_ub) {
    p.snapshot_in_srt = false;
    if (!no_commit) {
      std::swap(s->cur.pos, s->srt.pos);
      std::swap(s->cur.vel, s->srt.vel);
    }
  }
  if (churned) s->cur_has_dead = true;
  s->binned_ahead = n_ub && sorted_path(s) && s->opt_bin_ahead;
  s->pending.push_back(p);
  s->steps_enqueued += 1;
  s->index_valid = false;
  s->slot_valid = false;
  s->tr_valid = false;
  s->launches += launches;
}

int rcs_step_async(rcs_sim* s, uint64_t secs, uint32_t nanos, uint32_t flags) {
  using namespace rcs_host;
  if (!s) return RCS_ERR_ARG;
  CU_TRY(s, cudaSetDevice(s->device));
  if (!s->local_group.empty()) {
    s->err = "handles of a single-process strip group are stepped together with rcs_dist_step_local";
    return RCS_ERR_ARG;
  }
  const double dt = (double)secs + (double)nanos / 1000000000.0;  // Duration::as_secs_f64
  // A step in the steady state -- nothing to upload, no trace, no kernel timing -- launches the same kernels with the
  // same arguments as the last step with the same key did: the second time a key comes up the step is captured into
  // a CUDA graph, from then on it is one graph launch (a 10 000-agent step is a dozen kernels of a few microseconds,
  // the launch gaps between them are most of its time; the same holds for a rank of an 8-GPU run).
  const bool steady = s->opt_graphs && !s->trace && !s->ktiming && !s->groups_dirty && !s->routes_dirty &&
                      !s->sources_dirty && !s->cnt_dirty && s->pending.size() < 4096;
  // Between processes the NCCL exchange stays outside the graph: captured send / recv pairs were measured 3 x slower
  // than eager ones on 8 ranks (1.85 against 0.59 ms per step; on 2 ranks there is no difference).  Phase A and the
  // exchange are then enqueued kernel by kernel and only phase B -- the rebuild and the step kernels -- replays.
  // With the peer-store transport (rcs_dist_peer_connect) the exchange is kernels of this library: the whole step is
  // captured, as on a single handle -- unless source sinks need the spawn-set all-reduce of phase A.
  const bool nccl_in_step = !s->peer.enabled || s->n_sources_alive != 0;
  const bool from_b = steady && s->strip.enabled && s->world > 1 && nccl_in_step;
  if (from_b) {
    int rc = step_phase_a(s, dt);
    if (rc) return rc;
    rc = step_exchange_nccl(s);
    if (rc) return rc;
  }
  StepKey key;
  if (steady) {
    key.epoch = s->graph_epoch;
    std::memcpy(&key.dt_bits, &dt, sizeof(double));
    key.cur_pos = s->cur.pos;
    key.flags = flags;
    key.n_ub = s->n_ub;
    key.n = s->n;
    key.cur_has_dead = s->cur_has_dead;
    key.binned_ahead = s->binned_ahead;
    if (!s->step_graphs.empty() && s->step_graphs.front().key.epoch != key.epoch) step_graphs_clear(s);
    for (const StepGraph& g : s->step_graphs) {
      if (g.key == key) {
        CU_TRY(s, cudaGraphLaunch(g.exec, s->stream));
        step_replay_host(s, flags, g.launches, from_b);
        s->graph_launches += 1;
        return RCS_OK;
      }
    }
    if ((s->recent_keys[0] == key || s->recent_keys[1] == key) && s->step_graphs.size() < 8) {
      const uint64_t launches0 = s->launches;
      CU_TRY(s, cudaStreamBeginCapture(s->stream, cudaStreamCaptureModeThreadLocal));
      int rc = step_enqueue(s, dt, flags, from_b);
      cudaGraph_t graph = nullptr;
      cudaError_t e = cudaStreamEndCapture(s->stream, &graph);
      if (rc) {
        if (graph) cudaGraphDestroy(graph);
        return rc;
      }
      if (e != cudaSuccess || !graph) {
        s->err = std::string("CUDA error: ") + cudaGetErrorString(e) + " at cudaStreamEndCapture";
        return RCS_ERR_CUDA;
      }
      StepGraph g;
      g.key = key;
      g.launches = s->launches - launches0;
      e = cudaGraphInstantiate(&g.exec, graph, 0);
      cudaGraphDestroy(graph);
      if (e != cudaSuccess) {
        s->err = std::string("CUDA error: ") + cudaGetErrorString(e) + " at cudaGraphInstantiate";
        return RCS_ERR_CUDA;
      }
      CU_TRY(s, cudaGraphLaunch(g.exec, s->stream));  // the host-side half was done by step_enqueue itself
      s->step_graphs.push_back(g);
      s->graph_captures += 1;
      return RCS_OK;
    }
  }
  int rc = step_enqueue(s, dt, flags, from_b);
  if (steady && rc == RCS_OK) {
    s->recent_keys[1] = s->recent_keys[0];
    s->recent_keys[0] = key;
  }
  return rc;
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

// ---- source sinks (lib.rs:159-168) ------------------------------------------------------------------
int rcs_add_source_sink(rcs_sim* s, const rcs_source_sink_desc* d, uint64_t* out_id) {
  if (!s || !d) return RCS_ERR_ARG;
  CU_TRY(s, cudaSetDevice(s->device));
  if (d->hl >= s->hls.size() || d->lp >= s->lps.size()) {
    s->err = "unknown planner handle";
    return RCS_ERR_ARG;
  }
  if (d->n_waypoints > WP_MASK) {
    s->err = "a source sink can have at most 65535 waypoints";
    return RCS_ERR_ARG;
  }
  if (d->n_waypoints == 0 || !d->waypoints_xy) {
    s->err = "a source sink needs at least one waypoint (the reference indexes waypoints[0], lib.rs:244)";
    return RCS_ERR_ARG;
  }
  if (s->hls[d->hl].kind == HL_ROUTE && d->n_waypoints > 1) {
    // RMFPlanner::set_target plans a new route to every waypoint (rmf/mod.rs:217-237); the device-side follower has
    // one caller-supplied polyline per planner, so it can serve the leg to the sink only
    s->err = "a route-follower source sink takes exactly one waypoint (the sink): one polyline per planner";
    return RCS_ERR_ARG;
  }
  uint64_t idx;
  if (!host_location_to_index(s->grid, d->source_x, d->source_y, idx)) {
    // the reference fails at the first spawn (lib.rs:146-149 -> :252); reported when the source is added
    s->err = "Failed to add agents from source";
    return RCS_ERR_SPAWN;
  }
  int rc = do_sync(s);
  if (rc) return rc;
  if (!s->ev_spawn_id) {
    s->ev_cap = (uint32_t)std::min<uint64_t>(s->cap + 4096, 0xfffffff0ull);
    CU_TRY(s, dalloc(&s->ev_spawn_id, s->ev_cap));
    CU_TRY(s, dalloc(&s->ev_spawn_xy, 2ull * s->ev_cap));
    CU_TRY(s, dalloc(&s->ev_destroyed, 2ull * s->ev_cap));
  }
  const uint64_t id = s->sources.size();
  SourceSinkDev q{};
  q.sx = d->source_x;
  q.sy = d->source_y;
  q.rate = d->monotonic_rate;
  q.thr2_sink = radius_threshold(d->radius_sink);
  // get_bounds(0.4, source), location_hash_2d.rs:103-122
  const GridDev& g = s->grid;
  q.pr = host_floor_as_i64(((d->source_x + 0.4) - g.offx) / g.res);
  q.pl = host_floor_as_i64(((d->source_x - 0.4) - g.offx) / g.res);
  q.pt = host_floor_as_i64(((d->source_y + 0.4) - g.offy) / g.res);
  q.pb = host_floor_as_i64(((d->source_y - 0.4) - g.offy) / g.res);
  q.wp_off = (uint32_t)(s->ss_wp.size() / 2);
  q.n_wp = (uint32_t)d->n_waypoints;
  q.grp = find_or_add_group(s, d->hl, d->lp, d->agent_eyesight_range, (int32_t)id);
  q.loop_forever = d->loop_forever ? 1u : 0u;
  q.alive = 1u;
  q.owned = 1u;
  if (s->strip.enabled) {
    // Every rank holds every source sink (same calls in the same order: the group tables and the source ids agree);
    // the rank whose strip holds the source's cell column spawns for it.  Its spawn probe (lib.rs:212-214) looks at
    // the agents that rank owns, so the probe's whole stencil of cell columns must lie inside that strip -- the same
    // verdict on every rank.
    const uint64_t cx = host_f64_as_usize((d->source_x - g.offx) / g.res);
    int owner = -1;
    uint64_t oc0 = 0, oc1 = 0;
    for (int r = 0; r < s->world; ++r) {
      strip_range(s, r, s->world, oc0, oc1);
      if (cx >= oc0 && cx < oc1) {
        owner = r;
        break;
      }
    }
    const bool left_ok = owner == 0 || q.pl >= (long long)oc0;
    const bool right_ok = owner == s->world - 1 || q.pr < (long long)oc1;
    if (owner < 0 || !left_ok || !right_ok) {
      s->err = "on strips a source sink must lie far enough inside its strip for its 0.4 m spawn probe to stay in the "
               "strip's cell columns";
      return RCS_ERR_ARG;
    }
    q.owned = owner == s->rank ? 1u : 0u;
  }
  s->ss_wp.insert(s->ss_wp.end(), d->waypoints_xy, d->waypoints_xy + 2 * d->n_waypoints);
  s->sources.push_back(q);
  s->sources_dirty = true;
  s->n_sources_alive += 1;
  s->ever_had_sources = true;
  if (out_id) *out_id = id;
  return RCS_OK;
}

int rcs_remove_source_sink(rcs_sim* s, uint64_t id) {
  if (!s) return RCS_ERR_ARG;
  CU_TRY(s, cudaSetDevice(s->device));
  if (id >= s->sources.size() || !s->sources[id].alive) {
    s->err = "unknown source sink id";
    return RCS_ERR_ARG;
  }
  int rc = do_sync(s);
  if (rc) return rc;
  // The reference panics at the next step if agents of the removed source sink are still alive
  // (lib.rs:164-168 leaves source_sink_agent_correspondence dangling, :309).  Here such agents simply stop
  // being tested against the waypoints and live on.
  s->sources[id].alive = 0u;
  s->sources_dirty = true;
  s->n_sources_alive -= 1;
  return RCS_OK;
}

int rcs_poll_events(rcs_sim* s, uint64_t spawned_cap, uint64_t* spawned_ids, double* spawned_xy, uint64_t* n_spawned,
                    uint64_t destroyed_cap, uint64_t* destroyed_ids, uint64_t* n_destroyed) {
  if (!s) return RCS_ERR_ARG;
  CU_TRY(s, cudaSetDevice(s->device));
  int rc = do_sync(s);
  const uint64_t ns = s->ev_spawn_id ? s->h_cnt[CNT_EV_SPAWN] : 0, nd = s->ev_spawn_id ? s->h_cnt[CNT_EV_DESTROY] : 0;
  if (n_spawned) *n_spawned = ns;
  if (n_destroyed) *n_destroyed = nd;
  if (rc) return rc;
  const bool take_s = ns == 0 || (spawned_ids && spawned_cap >= ns);
  const bool take_d = nd == 0 || (destroyed_ids && destroyed_cap >= nd);
  if (!take_s || !take_d || (ns == 0 && nd == 0)) return RCS_OK;  // counts only: nothing is consumed
  if (ns) {
    static_assert(sizeof(unsigned long long) == sizeof(uint64_t), "id width");
    CU_TRY(s, cudaMemcpy(spawned_ids, s->ev_spawn_id, ns * sizeof(uint64_t), cudaMemcpyDeviceToHost));
    if (spawned_xy) CU_TRY(s, cudaMemcpy(spawned_xy, s->ev_spawn_xy, 2 * ns * sizeof(double), cudaMemcpyDeviceToHost));
  }
  if (nd) {
    std::vector<uint64_t> rec(2 * nd);
    CU_TRY(s, cudaMemcpy(rec.data(), s->ev_destroyed, 2 * nd * sizeof(uint64_t), cudaMemcpyDeviceToHost));
    std::vector<uint32_t> order(nd);
    std::iota(order.begin(), order.end(), 0u);
    // canonical order: by step, ascending id inside a step (the reference's order is HashMap-random)
    std::sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) {
      if (rec[2 * a + 1] != rec[2 * b + 1]) return rec[2 * a + 1] < rec[2 * b + 1];
      return rec[2 * a] < rec[2 * b];
    });
    for (uint64_t k = 0; k < nd; ++k) destroyed_ids[k] = rec[2 * order[k]];
  }
  CU_TRY(s, cudaMemset(s->cnt + CNT_EV_SPAWN, 0, 2 * sizeof(uint32_t)));
  s->h_cnt[CNT_EV_SPAWN] = s->h_cnt[CNT_EV_DESTROY] = 0;
  return RCS_OK;
}

}  // extern "C"
