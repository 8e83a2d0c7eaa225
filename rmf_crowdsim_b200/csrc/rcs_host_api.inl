// rcs_host_api.inl -- handle life cycle, planner descriptors, agents and the pub `agents` view.
// Part of rcs.cu (single translation unit).

using namespace rcs_host;

static void dist_teardown(rcs_sim* s);

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
  g.inv_res = 1.0 / desc->cell_size;
  s->cell_lo = 0;
  s->cell_hi = g.len;
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
  CR_TRY(dalloc(&s->slow_list, s->cap + 16));
  CR_TRY(dalloc(&s->wide_list, s->cap + 16));
  CR_TRY(dalloc(&s->keep, s->cap + 16));
  CR_TRY(dalloc(&s->slices, s->cap + 16));
  CR_TRY(dalloc(&s->tile_ranges, s->cap / GATHER_THREADS + 2));
  CR_TRY(dalloc(&s->cnt, CNT_N));
  CR_TRY(dalloc(&s->d_next_id, 1));
  CR_TRY(cudaMemset(s->cnt, 0, CNT_N * sizeof(uint32_t)));
  CR_TRY(cudaMemset(s->d_next_id, 0, sizeof(unsigned long long)));
  CR_TRY(cudaMallocHost(reinterpret_cast<void**>(&s->h_cnt), CNT_N * sizeof(uint32_t)));
  std::memset(s->h_cnt, 0, CNT_N * sizeof(uint32_t));
  CR_TRY(dalloc(&s->d_status, 1));
  CR_TRY(dalloc(&s->d_steps_done, 1));
  CR_TRY(dalloc(&s->d_bad, 1));
  CR_TRY(dalloc(&s->d_bad2, 1));
  CR_TRY(cudaMemset(s->d_status, 0, sizeof(DevStatus)));
  CR_TRY(cudaMemset(s->d_steps_done, 0, sizeof(unsigned long long)));
  CR_TRY(cudaMemset(s->scan_total, 0xff, 4 * sizeof(uint32_t)));
  CR_TRY(cudaMemset(s->cell_start, 0, (g.len + 16) * sizeof(uint32_t)));
  CR_TRY(cudaMallocHost(reinterpret_cast<void**>(&s->h_status), sizeof(DevStatus)));
  for (uint32_t k = 0; k < RCS_NUM_EVENTS; ++k) CR_TRY(cudaEventCreate(&s->events[k]));
#undef CR_TRY
  s->stats.first_oob_id = ~0ull;
  if (const char* e = std::getenv("RCS_GRAPHS")) s->opt_graphs = std::atoi(e) != 0 ? 1u : 0u;  // default of RCS_OPT_GRAPHS
  if (const char* e = std::getenv("RCS_PDL")) s->opt_pdl = std::atoi(e) != 0 ? 1u : 0u;        // default of RCS_OPT_PDL
  *out = s;
  return RCS_OK;
}

void rcs_sim_destroy(rcs_sim* s) {
  if (!s) return;
  cudaSetDevice(s->device);
  if (s->stream) cudaStreamSynchronize(s->stream);
  step_graphs_clear(s);
  free_agent_arrays(s->cur);
  free_agent_arrays(s->srt);
  cudaFree(s->cellid); cudaFree(s->perm); cudaFree(s->order_by_id); cudaFree(s->cell_count);
  cudaFree(s->cell_start); cudaFree(s->cursor); cudaFree(s->tile_sums); cudaFree(s->scan_total);
  cudaFree(s->big_list); cudaFree(s->d_groups); cudaFree(s->d_routes); cudaFree(s->d_status); cudaFree(s->d_steps_done);
  cudaFree(s->d_bad); cudaFree(s->d_bad2); cudaFree(s->slot_of_id); cudaFree(s->id_rank); cudaFree(s->presence);
  cudaFree(s->tr_ti); cudaFree(s->tr_fx); cudaFree(s->tr_fy); cudaFree(s->tr_nbc); cudaFree(s->tr_nbo);
  cudaFree(s->tr_nbids); cudaFree(s->tr_id); cudaFree(s->tr_own); cudaFree(s->stage); cudaFree(s->flush_buf);
  cudaFree(s->slow_list); cudaFree(s->wide_list); cudaFree(s->keep); cudaFree(s->slices); cudaFree(s->tile_ranges); cudaFree(s->cnt); cudaFree(s->d_next_id); cudaFree(s->srt_cell);
  cudaFree(s->d_sources); cudaFree(s->d_ss_wp); cudaFree(s->d_blocked); cudaFree(s->d_sg_start);
  cudaFree(s->d_ss_bits_local); cudaFree(s->d_ss_bits_parts); cudaFree(s->d_ss_bits);
  cudaFree(s->d_sg_items); cudaFree(s->ev_spawn_id); cudaFree(s->ev_destroyed); cudaFree(s->ev_spawn_xy);
  dist_teardown(s);
  if (s->h_status) cudaFreeHost(s->h_status);
  if (s->h_cnt) cudaFreeHost(s->h_cnt);
  for (uint32_t k = 0; k < RCS_NUM_EVENTS; ++k)
    if (s->events[k]) cudaEventDestroy(s->events[k]);
  for (auto& pr : s->kevents) {
    cudaEventDestroy(pr.first);
    cudaEventDestroy(pr.second);
  }
  for (auto e : s->kevent_pool) cudaEventDestroy(e);
  if (s->copy_stream) {
    cudaStreamSynchronize(s->copy_stream);
    cudaStreamDestroy(s->copy_stream);
    cudaEventDestroy(s->ev_gathered);
    cudaEventDestroy(s->ev_read_done);
    cudaEventDestroy(s->ev_half_done[0]);
    cudaEventDestroy(s->ev_half_done[1]);
  }
  cudaFree(s->stage2);
  if (s->up_stream) {
    cudaStreamSynchronize(s->up_stream);
    cudaStreamDestroy(s->up_stream);
    cudaEventDestroy(s->ev_uploaded);
    cudaEventDestroy(s->ev_pv_scattered);
  }
  cudaFree(s->pv_stage);
  inloop_free(s->inloop);
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
  s->hls.push_back(HLDesc{kind, vx, vy, 0u, 0u});
  *out_hl = (uint32_t)s->hls.size() - 1;
  return RCS_OK;
}
int rcs_hl_constant(rcs_sim* s, double vx, double vy, uint32_t* out_hl) { return push_hl(s, HL_CONSTANT, vx, vy, out_hl); }
int rcs_hl_parity(rcs_sim* s, double vx, double vy, uint32_t* out_hl) { return push_hl(s, HL_PARITY, vx, vy, out_hl); }
int rcs_hl_none(rcs_sim* s, uint32_t* out_hl) { return push_hl(s, HL_NONE, 0, 0, out_hl); }
int rcs_hl_route(rcs_sim* s, uint64_t n_points, const double* xy, uint32_t* out_hl) {
  if (!s || !out_hl || !xy || n_points == 0 || n_points > WP_MASK - 1) {
    if (s) s->err = "a route needs 1..65534 points";
    return RCS_ERR_ARG;
  }
  HLDesc h{HL_ROUTE, 0.0, 0.0, (uint32_t)(s->routes.size() / 2), (uint32_t)n_points};
  s->routes.insert(s->routes.end(), xy, xy + 2 * n_points);
  s->routes_dirty = true;
  s->hls.push_back(h);
  *out_hl = (uint32_t)s->hls.size() - 1;
  return RCS_OK;
}

int rcs_hl_route_set_target(rcs_sim* s, uint64_t m, const uint64_t* ids) {
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
  route_set_target_kernel<<<blocks_for(m, 256), 256, 0, s->stream>>>((uint32_t)m, d_ids, s->slot_of_id,
                                                                      std::max<uint64_t>(s->max_id_plus1, 1), s->cur.wp,
                                                                      s->d_bad);
  s->launches += 1;
  unsigned int bad = 0;
  CU_TRY(s, cudaMemcpyAsync(&bad, s->d_bad, sizeof(bad), cudaMemcpyDeviceToHost, s->stream));
  CU_TRY(s, cudaStreamSynchronize(s->stream));
  if (bad) {
    s->err = "unknown agent id";
    return RCS_ERR_ARG;
  }
  return RCS_OK;
}

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
  // Steps that are still in flight may have spawned agents (counted on the device only): s->n is exact -- and the
  // capacity check below meaningful -- only after the sync.
  int rc = do_sync(s);
  if (rc) return rc;
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
  if (s->strip.enabled) {
    for (uint64_t k = 0; k < n; ++k) {
      uint64_t cx = host_f64_as_usize((xy[2 * k] - s->grid.offx) / s->grid.res);
      if (cx < s->strip.c0 || cx >= s->strip.c1) {
        s->err = "agent position is outside this rank's strip";
        return RCS_ERR_ARG;
      }
    }
  }
  if (ids) {
    // id -> slot tables are dense arrays over the id range (12 bytes per id up to the largest one)
    for (uint64_t k = 0; k < n; ++k) {
      if (ids[k] >> 32) {
        s->err = "caller-supplied agent ids must be < 2^32";
        return RCS_ERR_ARG;
      }
    }
  }
  if (n == 0) {
    // strips: ghosts carry their group number, so every rank must hold the same group table in the same order --
    // also a rank whose strip is empty at this moment
    if (s->strip.enabled) find_or_add_group(s, hl, lp, eyesight, source_sink);
    return RCS_OK;
  }
  uint32_t grp = find_or_add_group(s, hl, lp, eyesight, source_sink);
  std::vector<double> hv;  // initial velocities: zero unless given (lib.rs:139)
  if (!vxy) hv.assign(2 * n, 0.0);
  std::vector<uint64_t> hid(n);
  std::vector<uint32_t> hgrp(n, grp), hwp(n, 0u);
  for (uint64_t k = 0; k < n; ++k) {
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
  // the caller's interleaved (x, y) rows are the device layout
  CU_TRY(s, cudaMemcpy(s->cur.pos + o, xy, n * sizeof(double2), cudaMemcpyHostToDevice));
  CU_TRY(s, cudaMemcpy(s->cur.vel + o, vxy ? vxy : hv.data(), n * sizeof(double2), cudaMemcpyHostToDevice));
  CU_TRY(s, cudaMemcpy(s->cur.id + o, hid.data(), n * sizeof(uint64_t), cudaMemcpyHostToDevice));
  CU_TRY(s, cudaMemcpy(s->cur.grp + o, hgrp.data(), n * sizeof(uint32_t), cudaMemcpyHostToDevice));
  CU_TRY(s, cudaMemcpy(s->cur.wp + o, hwp.data(), n * sizeof(uint32_t), cudaMemcpyHostToDevice));
  if (s->cur.pv) {
    const double nan = std::numeric_limits<double>::quiet_NaN();
    fill_f64_kernel<<<blocks_for(2 * n, 256), 256, 0, s->stream>>>(2 * n, reinterpret_cast<double*>(s->cur.pv + o), nan);
    s->launches += 1;
    CU_TRY(s, cudaStreamSynchronize(s->stream));
  }
  s->n += (uint32_t)n;
  s->n_ub = s->n;
  s->cnt_dirty = true;
  if (!ids) {
    unsigned long long next = s->last_alloc_agent_id;
    CU_TRY(s, cudaMemcpy(s->d_next_id, &next, sizeof(next), cudaMemcpyHostToDevice));
  }
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
  out.pos[k] = in.pos[i];
  out.vel[k] = in.vel[i];
  out.id[k] = in.id[i];
  out.grp[k] = in.grp[i];
  out.wp[k] = in.wp[i];
  if (in.pv) out.pv[k] = in.pv[i];
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
  s->n_ub = kept;
  s->cnt_dirty = true;
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
  double* const posd = reinterpret_cast<double*>(s->cur.pos);
  double* const veld = reinterpret_cast<double*>(s->cur.vel);
  double* dsts[4] = {posd, posd + 1, veld, veld + 1};  // components of the interleaved arrays: stride 2
  for (int a = 0; a < 4; ++a) {
    if (!srcs[a]) continue;
    CU_TRY(s, cudaMemcpyAsync(d_val, srcs[a], m * sizeof(double), cudaMemcpyHostToDevice, s->stream));
    scatter_by_id_kernel<double><<<blocks_for(m, 256), 256, 0, s->stream>>>(
        (uint32_t)m, s->order_by_id, d_ids, s->slot_of_id, std::max<uint64_t>(s->max_id_plus1, 1), d_val, 1, dsts[a],
        2, s->d_bad);
    s->launches += 1;
  }
  unsigned int bad = 0;
  CU_TRY(s, cudaMemcpyAsync(&bad, s->d_bad, sizeof(bad), cudaMemcpyDeviceToHost, s->stream));
  CU_TRY(s, cudaStreamSynchronize(s->stream));
  s->index_valid = false;
  s->binned_ahead = false;
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
  if (!s->cur.pv) {
    s->err = "no rcs_hl_host planner exists on this handle";
    return RCS_ERR_ARG;
  }
  int rc = RCS_OK;
  if (churn(s)) {  // the agent set may have changed on the device: settle it before addressing agents by id
    rc = do_sync(s);
    if (rc) return rc;
  }
  if (!ids && m != s->n) {
    s->err = "ids == NULL requires n == agent count";
    return RCS_ERR_ARG;
  }
  // The upload runs on its own stream into its own staging buffer, so it overlaps whatever the step stream is still
  // doing (the previous step, a read-back); only the scatter into the agents' rows is ordered after both.
  if (!s->up_stream) {
    CU_TRY(s, cudaStreamCreateWithFlags(&s->up_stream, cudaStreamNonBlocking));
    CU_TRY(s, cudaEventCreateWithFlags(&s->ev_uploaded, cudaEventDisableTiming));
    CU_TRY(s, cudaEventCreateWithFlags(&s->ev_pv_scattered, cudaEventDisableTiming));
  }
  const uint64_t need = m * (sizeof(uint64_t) + 2 * sizeof(double)) + 256;
  if (need > s->pv_stage_bytes) {
    CU_TRY(s, cudaStreamSynchronize(s->up_stream));
    CU_TRY(s, cudaStreamSynchronize(s->stream));
    cudaFree(s->pv_stage);
    s->pv_stage = nullptr;
    s->pv_stage_bytes = 0;
    CU_TRY(s, cudaMalloc(&s->pv_stage, need + need / 4));
    s->pv_stage_bytes = need + need / 4;
    s->pv_inflight = false;
  }
  rc = build_slot_table(s);
  if (rc) return rc;
  double2* d_val = static_cast<double2*>(s->pv_stage);
  uint64_t* d_ids = nullptr;
  if (s->pv_inflight) CU_TRY(s, cudaStreamWaitEvent(s->up_stream, s->ev_pv_scattered, 0));  // buffer consumed
  CU_TRY(s, cudaMemcpyAsync(d_val, vxy, 2 * m * sizeof(double), cudaMemcpyHostToDevice, s->up_stream));
  if (ids) {
    d_ids = reinterpret_cast<uint64_t*>(static_cast<char*>(s->pv_stage) + 2 * m * sizeof(double));
    CU_TRY(s, cudaMemcpyAsync(d_ids, ids, m * sizeof(uint64_t), cudaMemcpyHostToDevice, s->up_stream));
  }
  CU_TRY(s, cudaEventRecord(s->ev_uploaded, s->up_stream));
  CU_TRY(s, cudaStreamWaitEvent(s->stream, s->ev_uploaded, 0));
  CU_TRY(s, cudaMemsetAsync(s->d_bad, 0, sizeof(unsigned int), s->stream));
  const uint64_t L = std::max<uint64_t>(s->max_id_plus1, 1);
  scatter_by_id_kernel<double2><<<blocks_for(m, 256), 256, 0, s->stream>>>((uint32_t)m, s->order_by_id, d_ids,
                                                                           s->slot_of_id, L, d_val, 1, s->cur.pv, 1,
                                                                           s->d_bad);
  s->launches += 1;
  CU_TRY(s, cudaGetLastError());
  CU_TRY(s, cudaEventRecord(s->ev_pv_scattered, s->stream));
  s->pv_inflight = true;
  if (ids) {
    unsigned int bad = 0;
    CU_TRY(s, cudaMemcpyAsync(&bad, s->d_bad, sizeof(bad), cudaMemcpyDeviceToHost, s->stream));
    CU_TRY(s, cudaStreamSynchronize(s->stream));
    if (bad) {
      s->err = "unknown agent id";
      return RCS_ERR_ARG;
    }
  }
  // `vxy` (pinned) may be reused once the step stream has passed the scatter: after any rcs_sync / rcs_step / blocking read
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
  // positions and velocities are stored as interleaved pairs: every requested component is gathered (in the
  // requested order; identity for storage order) into a staging buffer and copied out from there
  rc = ensure_stage(s, (uint64_t)n * 48 + 256);
  if (rc) return rc;
  uint64_t off = 0;
  if (ids) { rc = read_array<uint64_t>(s, s->cur.id, 1, ord, n, ids, off); off += (uint64_t)n * 8; if (rc) return rc; }
  if (x || y || vx || vy) {
    double* st = reinterpret_cast<double*>(static_cast<char*>(s->stage) + off);
    double* host[4] = {x, y, vx, vy};
    double* dev[4];
    for (int k = 0; k < 4; ++k) dev[k] = host[k] ? st + (uint64_t)k * n : nullptr;
    gather_state_kernel<<<blocks_for(n, 256), 256, 0, s->stream>>>(n, ord, s->cur.pos, s->cur.vel, dev[0], dev[1],
                                                                   dev[2], dev[3]);
    s->launches += 1;
    for (int k = 0; k < 4; ++k)
      if (host[k])
        CU_TRY(s, cudaMemcpyAsync(host[k], dev[k], (uint64_t)n * 8, cudaMemcpyDeviceToHost, s->stream));
    off += (uint64_t)n * 32;
  }
  if (next_waypoint) { rc = read_array<uint32_t>(s, s->cur.wp, 1, ord, n, next_waypoint, off); if (rc) return rc; }
  CU_TRY(s, cudaStreamSynchronize(s->stream));
  if (next_waypoint && s->any_route)  // the high half of the word is the route follower's cache entry
    for (uint32_t k = 0; k < n; ++k) next_waypoint[k] &= WP_MASK;
  return RCS_OK;
}

// Read-back that does not stall the step stream: the requested arrays are gathered into a private staging
// buffer on the step stream (device-speed), the device-to-host copies run on a second stream, and the call
// returns at once.  The next rcs_set_preferred_velocity / rcs_step_async therefore overlap with the copies
// (PCIe is full duplex).  The host buffers are valid after rcs_read_wait.
int rcs_read_agents_async(rcs_sim* s, uint32_t order, uint64_t cap, uint64_t* ids, double* x, double* y, double* vx,
                          double* vy, uint64_t* out_n) {
  if (!s) return RCS_ERR_ARG;
  CU_TRY(s, cudaSetDevice(s->device));
  int rc = RCS_OK;
  if (churn(s)) {  // the agent count lives on the device while steps with churn are in flight
    rc = do_sync(s);
    if (rc) return rc;
  }
  const uint32_t n = s->n;
  if (out_n) *out_n = n;
  if (n == 0) return RCS_OK;
  if (cap < n) {
    s->err = "output capacity too small";
    return RCS_ERR_CAPACITY;
  }
  if (!s->copy_stream) {
    CU_TRY(s, cudaStreamCreateWithFlags(&s->copy_stream, cudaStreamNonBlocking));
    CU_TRY(s, cudaEventCreateWithFlags(&s->ev_gathered, cudaEventDisableTiming));
    CU_TRY(s, cudaEventCreateWithFlags(&s->ev_read_done, cudaEventDisableTiming));
    CU_TRY(s, cudaEventCreateWithFlags(&s->ev_half_done[0], cudaEventDisableTiming));
    CU_TRY(s, cudaEventCreateWithFlags(&s->ev_half_done[1], cudaEventDisableTiming));
  }
  // two staging halves: the gather of this read runs while the previous read's copies are still draining the other
  const uint64_t need = (uint64_t)n * 40 + 256;
  if (need > s->stage2_bytes) {
    CU_TRY(s, cudaStreamSynchronize(s->copy_stream));
    CU_TRY(s, cudaStreamSynchronize(s->stream));
    cudaFree(s->stage2);
    s->stage2 = nullptr;
    s->stage2_bytes = 0;
    const uint64_t half = (need + need / 4 + 255) & ~255ull;
    CU_TRY(s, cudaMalloc(&s->stage2, 2 * half));
    s->stage2_bytes = half;
    s->half_used[0] = s->half_used[1] = false;
  }
  const int hb = (int)(s->read_seq++ & 1u);
  const uint32_t* ord = nullptr;
  if (order == RCS_ORDER_ID) {
    rc = build_slot_table(s);
    if (rc) return rc;
    ord = s->order_by_id;
  }
  // the read before the previous one used this half: its copies must have drained it (device-side wait)
  if (s->half_used[hb]) CU_TRY(s, cudaStreamWaitEvent(s->stream, s->ev_half_done[hb], 0));
  void* host[5] = {ids, x, y, vx, vy};
  char* base = static_cast<char*>(s->stage2) + (uint64_t)hb * s->stage2_bytes;
  if (ids) {
    gather_kernel<unsigned long long><<<blocks_for(n, 256), 256, 0, s->stream>>>(
        n, ord, reinterpret_cast<const unsigned long long*>(s->cur.id), 1, reinterpret_cast<unsigned long long*>(base));
    s->launches += 1;
  }
  if (x || y || vx || vy) {
    double* dev[4];
    for (int k = 0; k < 4; ++k)
      dev[k] = host[k + 1] ? reinterpret_cast<double*>(base + (uint64_t)(k + 1) * n * 8) : nullptr;
    gather_state_kernel<<<blocks_for(n, 256), 256, 0, s->stream>>>(n, ord, s->cur.pos, s->cur.vel, dev[0], dev[1],
                                                                   dev[2], dev[3]);
    s->launches += 1;
  }
  CU_TRY(s, cudaGetLastError());
  CU_TRY(s, cudaEventRecord(s->ev_gathered, s->stream));
  CU_TRY(s, cudaStreamWaitEvent(s->copy_stream, s->ev_gathered, 0));
  for (int k = 0; k < 5; ++k)
    if (host[k])
      CU_TRY(s, cudaMemcpyAsync(host[k], base + (uint64_t)k * n * 8, (uint64_t)n * 8, cudaMemcpyDeviceToHost,
                                s->copy_stream));
  CU_TRY(s, cudaEventRecord(s->ev_half_done[hb], s->copy_stream));
  CU_TRY(s, cudaEventRecord(s->ev_read_done, s->copy_stream));
  s->half_used[hb] = true;
  s->read_inflight = true;
  return RCS_OK;
}

int rcs_read_wait(rcs_sim* s) {
  if (!s) return RCS_ERR_ARG;
  CU_TRY(s, cudaSetDevice(s->device));
  if (s->read_inflight) {
    CU_TRY(s, cudaEventSynchronize(s->ev_read_done));
    s->read_inflight = false;
  }
  return RCS_OK;
}

}  // extern "C"
