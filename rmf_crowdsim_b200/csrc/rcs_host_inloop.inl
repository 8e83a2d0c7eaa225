// rcs_host_inloop.inl -- rcs_step_in_loop: one step under the reference's in-loop index semantic (rcs_inloop.cuh).
// Study mode: synchronous, one handle, no strips, no source sinks, no route follower.

namespace rcs_host {

static int inloop_alloc(rcs_sim* s) {
  if (s->inloop) return RCS_OK;
  InLoopMem* m = new InLoopMem;
  s->inloop = m;
  m->cap = s->cap;
  m->cells = s->grid.len;
  CU_TRY(s, dalloc(&m->np_a, m->cap + 16));
  CU_TRY(s, dalloc(&m->np_b, m->cap + 16));
  CU_TRY(s, dalloc(&m->nv, m->cap + 16));
  CU_TRY(s, dalloc(&m->rank, m->cap + 16));
  CU_TRY(s, dalloc(&m->cellid, m->cap + 16));
  CU_TRY(s, dalloc(&m->perm, m->cap + 16));
  CU_TRY(s, dalloc(&m->cell_count, m->cells + 16));
  CU_TRY(s, dalloc(&m->cell_start, m->cells + 16));
  CU_TRY(s, dalloc(&m->cursor, m->cells + 16));
  CU_TRY(s, dalloc(&m->changed, 4));
  return RCS_OK;
}

__global__ void inloop_bump_steps_kernel(unsigned long long* steps_done) { *steps_done += 1ull; }

}  // namespace rcs_host

extern "C" {

int rcs_step_in_loop(rcs_sim* s, uint64_t secs, uint32_t nanos, const uint64_t* order, uint64_t n_order,
                     uint32_t max_sweeps, uint32_t* out_sweeps) {
  if (!s || (n_order && !order)) return RCS_ERR_ARG;
  CU_TRY(s, cudaSetDevice(s->device));
  if (out_sweeps) *out_sweeps = 0;
  if (s->strip.enabled || s->ever_had_sources || s->any_route) {
    s->err = "rcs_step_in_loop: one handle without strips, source sinks or route followers";
    return RCS_ERR_ARG;
  }
  int rc = ensure_index(s);  // syncs; canonical sorted copy of the old state in srt + cell_start
  if (rc) return rc;
  const uint32_t n = s->n;
  if (n == 0) return RCS_OK;
  uint32_t n_sorted = 0;
  CU_TRY(s, cudaMemcpy(&n_sorted, n_sorted_ptr(s), sizeof(uint32_t), cudaMemcpyDeviceToHost));
  if (n_sorted != n) {  // an injected position outside the grid: the reference would have refused it (add_or_update)
    s->err = "Index out of bounds";
    return RCS_ERR_OUT_OF_BOUNDS;
  }
  rc = inloop_alloc(s);
  if (rc) return rc;
  InLoopMem& m = *s->inloop;
  if (max_sweeps == 0) max_sweeps = 1u << 20;

  // iteration order -> rank of every sorted slot
  uint32_t* d_pos_of_id = nullptr;
  const uint64_t L = std::max<uint64_t>(s->max_id_plus1, 1);
  if (n_order) {
    rc = ensure_stage(s, n_order * sizeof(uint64_t) + L * sizeof(uint32_t) + 256);
    if (rc) return rc;
    uint64_t* d_order = static_cast<uint64_t*>(s->stage);
    d_pos_of_id = reinterpret_cast<uint32_t*>(static_cast<char*>(s->stage) + ((n_order * sizeof(uint64_t) + 255) & ~255ull));
    CU_TRY(s, cudaMemcpyAsync(d_order, order, n_order * sizeof(uint64_t), cudaMemcpyHostToDevice, s->stream));
    CU_TRY(s, cudaMemsetAsync(d_pos_of_id, 0xff, L * sizeof(uint32_t), s->stream));
    inloop_order_table_kernel<<<blocks_for(n_order, 256), 256, 0, s->stream>>>((uint32_t)n_order, d_order, L,
                                                                               d_pos_of_id);
    s->launches += 1;
  }
  inloop_rank_kernel<<<blocks_for(n, 256), 256, 0, s->stream>>>(n, s->srt.id, d_pos_of_id, L, m.rank);
  s->launches += 1;

  begin_step_kernel<<<1, 1, 0, s->stream>>>(s->d_status, s->cnt, nullptr, nullptr, nullptr);
  s->launches += 1;
  const double dt = (double)secs + (double)nanos / 1e9;  // Duration::as_secs_f64
  InLoopArgs q{};
  q.a = make_step_args(s, s->srt, s->cur, dt, n, false, 0);
  q.rank = m.rank;
  q.cell_start_new = m.cell_start;
  q.perm_new = m.perm;
  q.newvel = m.nv;
  q.changed = m.changed;
  double2 *prev = m.np_a, *next = m.np_b;
  const uint64_t len = s->grid.len;
  uint32_t sweeps = 0;
  bool converged = false;
  while (sweeps < max_sweeps) {
    q.first = sweeps == 0 ? 1u : 0u;
    if (!q.first) {
      // index of the previous sweep's new positions: counting sort by cell, ascending id inside a cell
      CU_TRY(s, cudaMemsetAsync(m.cell_count, 0, (len + 1) * sizeof(uint32_t), s->stream));
      inloop_bin_kernel<<<blocks_for(n, 256), 256, 0, s->stream>>>(s->grid, n, prev, m.cellid, m.cell_count);
      s->launches += 1;
      rc = exclusive_scan(s, m.cell_count, len, m.cell_start, m.cursor);
      if (rc) return rc;
      CU_TRY(s, cudaMemsetAsync(&s->d_status->big_cells, 0, sizeof(unsigned int), s->stream));
      scatter_perm_kernel<<<blocks_for(n, 256), 256, 0, s->stream>>>(n, s->cnt + CNT_CUR, m.cellid, m.cursor, m.perm,
                                                                      s->d_status);
      sort_cells_by_id_kernel<<<blocks_for(len, 128), 128, 0, s->stream>>>(0, len, m.cell_start, s->srt.id, m.perm,
                                                                           s->big_list, 4096, s->d_status);
      sort_big_cells_kernel<<<148, 1024, 0, s->stream>>>(0, len, m.cell_start, s->srt.id, m.perm, s->slow_list,
                                                         s->wide_list, s->big_list, 4096, s->d_status);
      s->launches += 3;
    }
    // statistics and trace describe the last sweep only
    CU_TRY(s, cudaMemsetAsync(&s->d_status->finite_tti, 0, 3 * sizeof(unsigned long long), s->stream));
    CU_TRY(s, cudaMemsetAsync(m.changed, 0, sizeof(uint32_t), s->stream));
    q.newpos_prev = prev;
    q.newpos_next = next;
    step_inloop_kernel<<<blocks_for(n, 128), 128, 0, s->stream>>>(q);
    s->launches += 1;
    CU_TRY(s, cudaGetLastError());
    uint32_t changed = 0;
    CU_TRY(s, cudaMemcpyAsync(&changed, m.changed, sizeof(uint32_t), cudaMemcpyDeviceToHost, s->stream));
    CU_TRY(s, cudaStreamSynchronize(s->stream));
    sweeps += 1;
    std::swap(prev, next);  // prev = the sweep just computed
    if (!q.first && changed == 0) {
      converged = true;
      break;
    }
  }
  if (out_sweeps) *out_sweeps = sweeps;
  if (!converged) {
    s->err = "rcs_step_in_loop: no fixed point within max_sweeps";
    return RCS_ERR_ARG;
  }
  // the converged sweep saw prev (== next bit for bit) as the new positions of the agents before each one
  if (s->trace) {
    q.newpos_prev = prev;
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
    trace_neighbours_inloop_kernel<<<blocks_for(n, 128), 128, 0, s->stream>>>(q, s->tr_nbo, s->tr_nbids);
    s->launches += 1;
    s->tr_nb_total = total;
    s->tr_n = n;
  }
  // add_or_update's verdict on the final positions (lib.rs:299-302), then the commit (lib.rs:350-359)
  inloop_commit_kernel<<<blocks_for(n, 256), 256, 0, s->stream>>>(n, s->grid, prev, m.nv, s->srt.id, nullptr, nullptr,
                                                                  s->d_status, 0u);
  s->launches += 1;
  CU_TRY(s, cudaMemcpyAsync(s->h_status, s->d_status, sizeof(DevStatus), cudaMemcpyDeviceToHost, s->stream));
  CU_TRY(s, cudaStreamSynchronize(s->stream));
  const bool oob = s->h_status->oob_count != 0;
  if (!oob) {
    inloop_commit_kernel<<<blocks_for(n, 256), 256, 0, s->stream>>>(n, s->grid, prev, m.nv, s->srt.id, s->srt.pos,
                                                                    s->srt.vel, s->d_status, 1u);
    inloop_bump_steps_kernel<<<1, 1, 0, s->stream>>>(s->d_steps_done);
    s->launches += 2;
    std::swap(s->cur, s->srt);  // the sorted arrays now hold the committed state
  }
  CU_TRY(s, cudaGetLastError());
  invalidate(s);
  s->tr_valid = s->trace;
  rc = do_sync(s);  // refreshes the statistics
  if (rc) return rc;
  if (oob) {
    s->err = "Index out of bounds";
    return RCS_ERR_OUT_OF_BOUNDS;
  }
  return RCS_OK;
}

}  // extern "C"
