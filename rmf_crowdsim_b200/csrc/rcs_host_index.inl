// rcs_host_index.inl -- the SpatialIndex trait in batched form, the parity trace, options and the
// measurement helpers.  Part of rcs.cu (single translation unit).

extern "C" {

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
  rc = build_index(s);
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
  query_radius_kernel<<<blocks_for(nq, 128), 128, 0, s->stream>>>(s->grid, s->cell_start, s->srt.pos, s->srt.id,
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
  query_radius_kernel<<<blocks_for(nq, 128), 128, 0, s->stream>>>(s->grid, s->cell_start, s->srt.pos, s->srt.id,
                                                                  (uint32_t)nq, d_q, d_r, d_t, d_c, d_off, d_ids, 1);
  s->launches += 1;
  cudaMemcpyAsync(out_ids, d_ids, total * 8, cudaMemcpyDeviceToHost, s->stream);
  cudaError_t e = cudaStreamSynchronize(s->stream);
  cudaFree(d_off);
  cudaFree(d_ids);
  CU_TRY(s, e);
  return RCS_OK;
}

int rcs_query_knn(rcs_sim* s, uint64_t nq, const double* qxy, uint64_t k, uint64_t* out_ids, uint64_t* out_counts) {
  if (!s || (nq && (!qxy || !out_counts || (k && !out_ids)))) return RCS_ERR_ARG;
  CU_TRY(s, cudaSetDevice(s->device));
  if (nq == 0) return RCS_OK;
  if (nq >= 0x7fffffffull) {
    s->err = "too many queries in one call";
    return RCS_ERR_ARG;
  }
  int rc = ensure_index(s);
  if (rc) return rc;
  // stage: qxy (16 nq) | counts (4 (nq+1)) | offsets (4 (nq+1))
  rc = ensure_stage(s, nq * 24 + 128);
  if (rc) return rc;
  char* base = static_cast<char*>(s->stage);
  double* d_q = reinterpret_cast<double*>(base);
  uint32_t* d_c = reinterpret_cast<uint32_t*>(base + nq * 16);
  uint32_t* d_o = d_c + ((nq + 1 + 3) & ~3ull);  // the scan kernels use 16-byte vector accesses
  CU_TRY(s, cudaMemcpyAsync(d_q, qxy, nq * 16, cudaMemcpyHostToDevice, s->stream));
  knn_count_kernel<<<blocks_for(nq, 128), 128, 0, s->stream>>>(s->grid, s->cell_start, (uint32_t)nq, d_q, k, d_c);
  s->launches += 1;
  rc = exclusive_scan(s, d_c, nq, d_o, nullptr);
  if (rc) return rc;
  uint32_t total = 0;
  CU_TRY(s, cudaMemcpyAsync(&total, d_o + nq, sizeof(uint32_t), cudaMemcpyDeviceToHost, s->stream));
  CU_TRY(s, cudaStreamSynchronize(s->stream));
  uint32_t* d_slot = nullptr;
  double* d_dist = nullptr;
  uint64_t* d_out = nullptr;
  uint64_t* d_cnt = nullptr;
  auto cleanup = [&]() {
    cudaFree(d_slot); cudaFree(d_dist); cudaFree(d_out); cudaFree(d_cnt);
  };
  cudaError_t e = dalloc(&d_slot, total);
  if (e == cudaSuccess) e = dalloc(&d_dist, total);
  if (e == cudaSuccess) e = dalloc(&d_out, nq * std::max<uint64_t>(k, 1));
  if (e == cudaSuccess) e = dalloc(&d_cnt, nq);
  if (e != cudaSuccess) {
    cleanup();
    CU_TRY(s, e);
  }
  knn_fill_kernel<<<blocks_for(nq, 128), 128, 0, s->stream>>>(s->grid, s->cell_start, s->srt.pos, (uint32_t)nq,
                                                              d_q, k, d_o, d_slot, d_dist);
  knn_select_kernel<<<blocks_for(nq * 32, 128), 128, 0, s->stream>>>((uint32_t)nq, k, d_o, d_slot, d_dist, s->srt.id,
                                                                     d_out, d_cnt);
  s->launches += 2;
  if (k) cudaMemcpyAsync(out_ids, d_out, nq * k * sizeof(uint64_t), cudaMemcpyDeviceToHost, s->stream);
  cudaMemcpyAsync(out_counts, d_cnt, nq * sizeof(uint64_t), cudaMemcpyDeviceToHost, s->stream);
  e = cudaStreamSynchronize(s->stream);
  cleanup();
  CU_TRY(s, e);
  return RCS_OK;
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
  // the slots of the n ids only (the table itself stays on the device: one agent at a time must not cost O(max id))
  std::vector<uint32_t> slots(n);
  {
    rc = ensure_stage(s, n * 12);
    if (rc) return rc;
    uint64_t* d_ids = static_cast<uint64_t*>(s->stage);
    uint32_t* d_slots = reinterpret_cast<uint32_t*>(static_cast<char*>(s->stage) + n * 8);
    CU_TRY(s, cudaMemcpyAsync(d_ids, ids, n * sizeof(uint64_t), cudaMemcpyHostToDevice, s->stream));
    lookup_slots_kernel<<<blocks_for(n, 256), 256, 0, s->stream>>>((uint32_t)n, d_ids, s->slot_of_id,
                                                                   std::max<uint64_t>(s->max_id_plus1, 1), d_slots);
    s->launches += 1;
    CU_TRY(s, cudaMemcpyAsync(slots.data(), d_slots, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, s->stream));
    CU_TRY(s, cudaStreamSynchronize(s->stream));
  }
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
    bool known = slots[k] != 0xffffffffu;
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
    CU_TRY(s, dalloc(&s->tr_id, s->cap + 16));
    CU_TRY(s, dalloc(&s->tr_own, s->cap + 16));
  }
  s->trace = on != 0;
  s->tr_valid = false;
  return RCS_OK;
}

// owned agents of the traced step and the total length of their neighbour lists
static int trace_host_copy(rcs_sim* s, std::vector<uint64_t>& sid, std::vector<uint32_t>& own,
                           std::vector<uint32_t>& hoff) {
  const uint32_t n = s->tr_n;
  sid.resize(n);
  own.resize(n);
  hoff.resize(n + 1);
  CU_TRY(s, cudaStreamSynchronize(s->stream));
  CU_TRY(s, cudaMemcpy(sid.data(), s->tr_id, n * sizeof(uint64_t), cudaMemcpyDeviceToHost));
  CU_TRY(s, cudaMemcpy(own.data(), s->tr_own, n * sizeof(uint32_t), cudaMemcpyDeviceToHost));
  CU_TRY(s, cudaMemcpy(hoff.data(), s->tr_nbo, (n + 1) * sizeof(uint32_t), cudaMemcpyDeviceToHost));
  return RCS_OK;
}

int rcs_trace_sizes(rcs_sim* s, uint64_t* n_agents, uint64_t* n_neighbours) {
  if (!s) return RCS_ERR_ARG;
  if (!s->tr_valid) {
    s->err = "no trace recorded (enable with rcs_set_trace, then step)";
    return RCS_ERR_ARG;
  }
  CU_TRY(s, cudaSetDevice(s->device));
  std::vector<uint64_t> sid;
  std::vector<uint32_t> own, hoff;
  int rc = trace_host_copy(s, sid, own, hoff);
  if (rc) return rc;
  uint64_t na = 0, nn = 0;
  for (uint32_t k = 0; k < s->tr_n; ++k)
    if (own[k]) {
      na += 1;
      nn += hoff[k + 1] - hoff[k];
    }
  if (n_agents) *n_agents = na;
  if (n_neighbours) *n_neighbours = nn;
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
  const uint32_t n = s->tr_n;
  // the trace arrays are in the canonical sorted order of the traced step (ghosts of a strip included:
  // only the agents this rank owns are reported)
  std::vector<uint64_t> sid;
  std::vector<uint32_t> own, hoff;
  int rc = trace_host_copy(s, sid, own, hoff);
  if (rc) return rc;
  std::vector<double> hti(n), hfx(n), hfy(n);
  std::vector<uint64_t> hnb(nb_ids ? s->tr_nb_total : 0);  // nb_ids == NULL: lengths only (nb_offsets)
  CU_TRY(s, cudaMemcpy(hti.data(), s->tr_ti, n * sizeof(double), cudaMemcpyDeviceToHost));
  CU_TRY(s, cudaMemcpy(hfx.data(), s->tr_fx, n * sizeof(double), cudaMemcpyDeviceToHost));
  CU_TRY(s, cudaMemcpy(hfy.data(), s->tr_fy, n * sizeof(double), cudaMemcpyDeviceToHost));
  if (nb_ids && s->tr_nb_total)
    CU_TRY(s, cudaMemcpy(hnb.data(), s->tr_nbids, s->tr_nb_total * sizeof(uint64_t), cudaMemcpyDeviceToHost));
  std::vector<uint32_t> order;
  for (uint32_t k = 0; k < n; ++k)
    if (own[k]) order.push_back(k);
  std::sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return sid[a] < sid[b]; });
  uint64_t off = 0;
  for (size_t r = 0; r < order.size(); ++r) {
    uint32_t k = order[r];
    if (ids) ids[r] = sid[k];
    if (t_i) t_i[r] = hti[k];
    if (fx) fx[r] = hfx[k];
    if (fy) fy[r] = hfy[k];
    if (nb_offsets) nb_offsets[r] = off;
    if (nb_ids)
      for (uint32_t j = hoff[k]; j < hoff[k + 1]; ++j) nb_ids[off + (j - hoff[k])] = hnb[j];
    off += hoff[k + 1] - hoff[k];
  }
  if (nb_offsets) nb_offsets[order.size()] = off;
  return RCS_OK;
}

int rcs_set_option(rcs_sim* s, uint32_t option, uint64_t value) {
  if (!s) return RCS_ERR_ARG;
  if (option == RCS_OPT_STEP_KERNEL && value <= 3) {
    s->graph_epoch += 1;
    s->opt_step_kernel = (uint32_t)value;
    return RCS_OK;
  }
  if (option == RCS_OPT_GRAPHS && value <= 1) {
    s->opt_graphs = (uint32_t)value;
    s->graph_epoch += 1;
    return RCS_OK;
  }
  if (option == RCS_OPT_PDL && value <= 1) {
    s->opt_pdl = (uint32_t)value;
    s->graph_epoch += 1;
    return RCS_OK;
  }
  if (option == RCS_OPT_BIN_AHEAD && value <= 1) {
    s->graph_epoch += 1;
    s->opt_bin_ahead = (uint32_t)value;
    s->binned_ahead = false;
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

int rcs_graph_stats(rcs_sim* s, uint64_t* out_graph_launches, uint64_t* out_captures) {
  if (!s) return RCS_ERR_ARG;
  if (out_graph_launches) *out_graph_launches = s->graph_launches;
  if (out_captures) *out_captures = s->graph_captures;
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

}  // extern "C"
