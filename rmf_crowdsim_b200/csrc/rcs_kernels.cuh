// rcs_kernels.cuh -- CUDA kernels of the per-step pipeline (sm_100a).
//
//   bin_count -> scan (3 small kernels, warp-shuffle scans) -> scatter_perm -> sort_cells_by_id
//   -> gather_sorted -> step_kernel (radius query + Zanlungo + explicit Euler, fused)
//
// Data layout: structure-of-arrays f64 positions / velocities, u64 ids, u32 group / waypoint,
// physically re-sorted every step into canonical (cell index, agent id) order, so the reference's
// scan `for x in left..=right { for y in bottom..=top }` (location_hash_2d.rs:245-246) is, for each
// x, one contiguous slice of the sorted arrays (the cell index is x-major, :59).
#pragma once

#include "rcs_math.cuh"

namespace rcs {

struct DevStatus {
  unsigned int oob_count;
  unsigned int nonfinite_count;
  unsigned int big_cells;       // cells with more agents than SORT_LOCAL_MAX (handled by the slow sorter)
  unsigned int failed;          // sticky: a previous async step failed -> later steps are skipped
  unsigned long long first_oob_id;
  unsigned long long finite_tti;
  unsigned long long neighbour_total;
  unsigned long long candidate_total;
};

struct AgentArrays {
  double *x, *y, *vx, *vy;
  uint64_t* id;
  uint32_t *grp, *wp;
  double *pvx, *pvy;  // host-planner preferred velocities; NaN = None.  nullptr if no host planner exists
};

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 16;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;
constexpr int SORT_LOCAL_MAX = 32;
constexpr uint32_t CELL_DEAD = 0xffffffffu;

// ---------------------------------------------------------------------------------------------
// A1/A2: cell of every agent (LocationHash2D::location_to_index) + histogram.
// ---------------------------------------------------------------------------------------------
__global__ void bin_count_kernel(GridDev g, uint32_t n, const double* __restrict__ x, const double* __restrict__ y,
                                 uint32_t* __restrict__ cellid, uint32_t* __restrict__ cell_count,
                                 DevStatus* status) {
  if (status->failed) return;
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint64_t idx;
  if (location_to_index(g, x[i], y[i], idx)) {
    cellid[i] = (uint32_t)idx;
    atomicAdd(&cell_count[idx], 1u);
  } else {
    // only reachable through snapshot injection; steps never commit an out-of-bounds position
    cellid[i] = CELL_DEAD;
    atomicAdd(&status->oob_count, 1u);
  }
}

// ---------------------------------------------------------------------------------------------
// Exclusive prefix sum over the cell histogram: reduce tiles -> scan tile sums -> scan tiles.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t warp_inclusive_scan(uint32_t v) {
  const unsigned lane = threadIdx.x & 31u;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    uint32_t t = __shfl_up_sync(0xffffffffu, v, d);
    if (lane >= (unsigned)d) v += t;
  }
  return v;
}

// exclusive scan of one value per thread across a SCAN_THREADS block; block total in `total`
__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t& total) {
  constexpr int NW = SCAN_THREADS / 32;
  __shared__ uint32_t warp_off[NW];
  __shared__ uint32_t s_total;
  const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  uint32_t inc = warp_inclusive_scan(v);
  if (lane == 31) warp_off[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    uint32_t w = lane < NW ? warp_off[lane] : 0u;
    uint32_t winc = warp_inclusive_scan(w);
    if (lane < NW) warp_off[lane] = winc - w;
    if (lane == NW - 1) s_total = winc;
  }
  __syncthreads();
  total = s_total;
  uint32_t r = warp_off[warp] + inc - v;
  __syncthreads();  // the shared scratch may be reused by the caller's next scan
  return r;
}

__global__ void scan_reduce_kernel(const uint32_t* __restrict__ in, uint64_t len, uint32_t* __restrict__ tile_sums) {
  uint64_t base = (uint64_t)blockIdx.x * SCAN_TILE + (uint64_t)threadIdx.x * SCAN_ITEMS;
  uint32_t s = 0;
  if (base + SCAN_ITEMS <= len) {
    const uint4* p = reinterpret_cast<const uint4*>(in + base);
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS / 4; ++k) {
      uint4 v = p[k];
      s += v.x + v.y + v.z + v.w;
    }
  } else {
    for (int k = 0; k < SCAN_ITEMS; ++k)
      if (base + k < len) s += in[base + k];
  }
  uint32_t total;
  (void)block_exclusive_scan(s, total);
  if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}

// single block: in-place exclusive scan of the tile sums; writes the grand total to *total_out
__global__ void scan_tile_sums_kernel(uint32_t* __restrict__ tile_sums, uint32_t n_tiles, uint32_t* total_out) {
  uint32_t carry = 0;
  for (uint32_t base = 0; base < n_tiles; base += SCAN_THREADS) {
    uint32_t i = base + threadIdx.x;
    uint32_t v = i < n_tiles ? tile_sums[i] : 0u;
    uint32_t total;
    uint32_t ex = block_exclusive_scan(v, total);
    if (i < n_tiles) tile_sums[i] = carry + ex;
    carry += total;
    __syncthreads();
  }
  if (threadIdx.x == 0) *total_out = carry;
}

// writes cell_start[c] (exclusive) and cursor[c] (= cell_start[c], bumped by the scatter); the
// thread that owns the last element also writes cell_start[len] = total.
__global__ void scan_apply_kernel(const uint32_t* __restrict__ in, uint64_t len,
                                  const uint32_t* __restrict__ tile_sums, uint32_t* __restrict__ cell_start,
                                  uint32_t* __restrict__ cursor) {
  uint64_t base = (uint64_t)blockIdx.x * SCAN_TILE + (uint64_t)threadIdx.x * SCAN_ITEMS;
  uint32_t v[SCAN_ITEMS];
  uint32_t s = 0;
  if (base + SCAN_ITEMS <= len) {
    const uint4* p = reinterpret_cast<const uint4*>(in + base);
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS / 4; ++k) {
      uint4 q = p[k];
      v[4 * k + 0] = q.x;
      v[4 * k + 1] = q.y;
      v[4 * k + 2] = q.z;
      v[4 * k + 3] = q.w;
    }
  } else {
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) v[k] = (base + k < len) ? in[base + k] : 0u;
  }
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; ++k) s += v[k];
  uint32_t total;
  uint32_t run = tile_sums[blockIdx.x] + block_exclusive_scan(s, total);
  if (base + SCAN_ITEMS <= len) {
    uint32_t o[SCAN_ITEMS];
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
      o[k] = run;
      run += v[k];
    }
    uint4* ps = reinterpret_cast<uint4*>(cell_start + base);
    uint4* pc = reinterpret_cast<uint4*>(cursor + base);
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS / 4; ++k) {
      uint4 q = make_uint4(o[4 * k], o[4 * k + 1], o[4 * k + 2], o[4 * k + 3]);
      ps[k] = q;
      if (cursor) pc[k] = q;
    }
    if (base + SCAN_ITEMS == len) cell_start[len] = run;
  } else {
    for (int k = 0; k < SCAN_ITEMS; ++k) {
      if (base + k < len) {
        cell_start[base + k] = run;
        if (cursor) cursor[base + k] = run;
        run += v[k];
        if (base + k + 1 == len) cell_start[len] = run;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// A2: counting-sort scatter of a permutation (slot order inside a cell is fixed by the next kernel)
// ---------------------------------------------------------------------------------------------
__global__ void scatter_perm_kernel(uint32_t n, const uint32_t* __restrict__ cellid, uint32_t* __restrict__ cursor,
                                    uint32_t* __restrict__ perm, const DevStatus* status) {
  if (status->failed) return;
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t c = cellid[i];
  if (c == CELL_DEAD) return;
  uint32_t pos = atomicAdd(&cursor[c], 1u);
  perm[pos] = i;
}

// Canonical order inside a cell = ascending agent id (the reference's HashSet order is random per
// process).  One thread per cell; insertion sort of the cell's permutation
// slice keyed by id.  Cells larger than SORT_LOCAL_MAX go to the block-wide rank sorter.
__global__ void sort_cells_by_id_kernel(uint64_t len, const uint32_t* __restrict__ cell_start,
                                        const uint64_t* __restrict__ id, uint32_t* __restrict__ perm,
                                        uint32_t* __restrict__ big_list, uint32_t big_cap, DevStatus* status) {
  if (status->failed) return;
  uint64_t c = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= len) return;
  uint32_t s = cell_start[c], e = cell_start[c + 1];
  uint32_t m = e - s;
  if (m < 2) return;
  if (m > SORT_LOCAL_MAX) {
    uint32_t k = atomicAdd(&status->big_cells, 1u);
    if (k < big_cap) big_list[k] = (uint32_t)c;
    return;
  }
  uint32_t p[SORT_LOCAL_MAX];
  uint64_t key[SORT_LOCAL_MAX];
  bool sorted = true;
  for (uint32_t k = 0; k < m; ++k) {
    p[k] = perm[s + k];
    key[k] = id[p[k]];
    if (k && key[k] < key[k - 1]) sorted = false;
  }
  if (sorted) return;
  for (uint32_t k = 1; k < m; ++k) {
    uint32_t pk = p[k];
    uint64_t kk = key[k];
    int j = (int)k - 1;
    while (j >= 0 && key[j] > kk) {
      key[j + 1] = key[j];
      p[j + 1] = p[j];
      --j;
    }
    key[j + 1] = kk;
    p[j + 1] = pk;
  }
  for (uint32_t k = 0; k < m; ++k) perm[s + k] = p[k];
}

// One block per oversized cell: rank = number of smaller ids (ids are unique), O(m^2 / threads).
__global__ void sort_big_cells_kernel(const uint32_t* __restrict__ cell_start, const uint64_t* __restrict__ id,
                                      uint32_t* __restrict__ perm, uint32_t* __restrict__ scratch,
                                      const uint32_t* __restrict__ big_list, uint32_t big_cap,
                                      const DevStatus* status) {
  if (status->failed) return;
  uint32_t nbig = status->big_cells < big_cap ? status->big_cells : big_cap;
  for (uint32_t b = blockIdx.x; b < nbig; b += gridDim.x) {
    uint32_t c = big_list[b];
    uint32_t s = cell_start[c], e = cell_start[c + 1];
    for (uint32_t k = s + threadIdx.x; k < e; k += blockDim.x) {
      uint64_t kk = id[perm[k]];
      uint32_t rank = 0;
      for (uint32_t j = s; j < e; ++j) rank += (id[perm[j]] < kk) ? 1u : 0u;
      scratch[s + rank] = perm[k];
    }
    __syncthreads();
    for (uint32_t k = s + threadIdx.x; k < e; k += blockDim.x) perm[k] = scratch[k];
    __syncthreads();
  }
}

// Physical reorder into canonical order: sorted[k] = cur[perm[k]].
__global__ void gather_sorted_kernel(uint32_t n, const uint32_t* __restrict__ perm, AgentArrays cur,
                                     AgentArrays srt, const uint32_t* __restrict__ n_sorted,
                                     const DevStatus* status) {
  if (status->failed) return;
  uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n || k >= *n_sorted) return;
  uint32_t i = perm[k];
  srt.x[k] = cur.x[i];
  srt.y[k] = cur.y[i];
  srt.vx[k] = cur.vx[i];
  srt.vy[k] = cur.vy[i];
  srt.id[k] = cur.id[i];
  srt.grp[k] = cur.grp[i];
  srt.wp[k] = cur.wp[i];
  if (cur.pvx) {
    srt.pvx[k] = cur.pvx[i];
    srt.pvy[k] = cur.pvy[i];
  }
}

// ---------------------------------------------------------------------------------------------
// The hot kernel: lib.rs:259-347 for every agent at once (deferred index semantic).
// ---------------------------------------------------------------------------------------------
struct StepArgs {
  GridDev grid;
  uint32_t n;                 // upper bound on live (sorted) agents
  const uint32_t* n_sorted;   // device-side exact count (= cell_start[len])
  AgentArrays in;             // canonical (cell, id) order
  const uint32_t* cell_start; // len + 1
  const GroupDev* groups;
  double dt;                  // Duration::as_secs_f64 (lib.rs:295)
  double *ox, *oy, *ovx, *ovy;  // new state, same order
  double *t_i, *fx, *fy;      // optional trace outputs (nullptr when tracing is off)
  uint32_t* nb_count;         // optional: neighbour count per agent (trace)
  DevStatus* status;
  uint32_t collect_stats;
};

struct Self {
  double px, py, vx, vy, pfx, pfy;
  uint64_t id;
};

// Walks the reference's radius query for one agent and calls f(j) for every neighbour j that
// passes the strict distance filter (location_hash_2d.rs:251) and the self filter (lib.rs:284),
// in canonical order.  Returns the number of candidates distance-tested.
template <class F>
__device__ __forceinline__ uint32_t for_each_neighbour(const GridDev& g, const uint32_t* __restrict__ cell_start,
                                                       const double* __restrict__ xs, const double* __restrict__ ys,
                                                       const uint64_t* __restrict__ ids, double px, double py,
                                                       uint64_t self_id, double radius, double thr2, F&& f) {
  int64_t left, right, bottom, top;
  get_bounds(g, radius, px, py, left, right, bottom, top);
  if (left < 0) left = 0;
  if (right > g.x_max) right = g.x_max;
  uint32_t cand = 0;
  for (int64_t cx = left; cx <= right; ++cx) {
    uint64_t c_lo, c_hi;
    if (!column_cell_range(g, cx, bottom, top, c_lo, c_hi)) continue;
    uint32_t s = cell_start[c_lo], e = cell_start[c_hi + 1];
    cand += e - s;
    for (uint32_t j = s; j < e; ++j) {
      double dx = xs[j] - px;
      double dy = ys[j] - py;
      double d2 = dx * dx + dy * dy;
      if (d2 < thr2) {
        if (ids[j] != self_id) f(j, dx, dy, d2);
      }
    }
  }
  return cand;
}

// Zanlungo::get_desired_velocity (zanlungo.rs:201-218) for ONE agent, sequentially, in canonical
// neighbour order: t_i (compute_tti) then the force sum.  The reference form of the hot path; the
// warp-cooperative kernel must reproduce its results bit for bit and uses it for agents whose
// stencil does not fit its fast path.
__device__ __noinline__ void zanlungo_sequential(const StepArgs& a, uint32_t i, const Self& me, const GroupDev& g,
                                                 double& t_i, double& fx, double& fy, uint32_t& nbc,
                                                 uint32_t& cand) {
  const double* __restrict__ xs = a.in.x;
  const double* __restrict__ ys = a.in.y;
  const double* __restrict__ vxs = a.in.vx;
  const double* __restrict__ vys = a.in.vy;
  const uint64_t* __restrict__ ids = a.in.id;
  const double rr = g.rr;
  t_i = RCS_INF;
  fx = 0.0;
  fy = 0.0;
  // Zanlungo::compute_tti, zanlungo.rs:76-91
  cand = for_each_neighbour(a.grid, a.cell_start, xs, ys, ids, me.px, me.py, me.id, g.eyesight, g.thr2,
                            [&](uint32_t j, double dx, double dy, double d2) {
                              nbc++;
                              double col_time = time_to_collision(vxs[j] - me.vx, vys[j] - me.vy, dx, dy, d2, rr);
                              if (col_time < t_i) t_i = col_time;
                            });
  // zanlungo.rs:210-215
  if (t_i != RCS_INF) {
    const OwnerPre pre = owner_precompute(me.px, me.py, me.vx, me.vy, me.pfx, me.pfy, t_i, g);
    const double ti = t_i;
    for_each_neighbour(a.grid, a.cell_start, xs, ys, ids, me.px, me.py, me.id, g.eyesight, g.thr2,
                       [&](uint32_t j, double, double, double) {
                         double qx, qy;
                         if (pair_force_dispatch(pre, me.px, me.py, me.vx, me.vy, me.pfx, me.pfy, me.id, xs[j], ys[j],
                                                 vxs[j], vys[j], ids[j], ti, g, qx, qy)) {
                           fx = fx + qx;
                           fy = fy + qy;
                         }
                       });
  }
}

// HighLevelPlanner::get_desired_velocity on the device (lib.rs:263-273): returns the recommended
// velocity and sets me.pfx/pfy (the clone's preferred_vel, lib.rs:271).
__device__ __forceinline__ void high_level_velocity(const StepArgs& a, uint32_t i, const GroupDev& g, Self& me,
                                                    double& velx, double& vely) {
  velx = 0.0;
  vely = 0.0;
  me.pfx = 0.0;
  me.pfy = 0.0;
  switch (g.hl_kind) {
    case HL_CONSTANT:
      velx = g.hl_vx;
      vely = g.hl_vy;
      me.pfx = velx;
      me.pfy = vely;
      break;
    case HL_PARITY:
      if ((me.id & 1ull) == 0ull) {
        velx = -g.hl_vx;
        vely = -g.hl_vy;
      } else {
        velx = g.hl_vx;
        vely = g.hl_vy;
      }
      me.pfx = velx;
      me.pfy = vely;
      break;
    case HL_HOST: {
      double hx = a.in.pvx[i], hy = a.in.pvy[i];
      if (hx == hx) {  // NaN in x encodes None
        velx = hx;
        vely = hy;
        me.pfx = hx;
        me.pfy = hy;
      }
    } break;
    default:
      break;
  }
}

// Explicit Euler + commit outputs + error accounting (lib.rs:295-302) for one agent.
__device__ __forceinline__ void integrate_and_store(const StepArgs& a, uint32_t i, const Self& me, double velx,
                                                    double vely, double t_i, double fx, double fy, uint32_t nbc) {
  const double nx = me.px + velx * a.dt;
  const double ny = me.py + vely * a.dt;
  a.ox[i] = nx;
  a.oy[i] = ny;
  a.ovx[i] = velx;
  a.ovy[i] = vely;
  if (a.t_i) {
    a.t_i[i] = t_i;
    a.fx[i] = fx;
    a.fy[i] = fy;
    a.nb_count[i] = nbc;
  }
  uint64_t idx;
  if (!location_to_index(a.grid, nx, ny, idx)) {  // add_or_update(new_pos) error path, lib.rs:299-302
    atomicAdd(&a.status->oob_count, 1u);
    atomicMin(&a.status->first_oob_id, (unsigned long long)me.id);
  }
  if (!(isfinite(nx) && isfinite(ny) && isfinite(velx) && isfinite(vely))) atomicAdd(&a.status->nonfinite_count, 1u);
}

__device__ __forceinline__ void warp_stats(const StepArgs& a, uint32_t cand, uint32_t nbc, uint32_t finite) {
  if (!a.collect_stats) return;
  unsigned long long c = cand, nb = nbc, ft = finite;
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    c += __shfl_down_sync(0xffffffffu, c, d);
    nb += __shfl_down_sync(0xffffffffu, nb, d);
    ft += __shfl_down_sync(0xffffffffu, ft, d);
  }
  if ((threadIdx.x & 31) == 0) {
    if (c) atomicAdd(&a.status->candidate_total, c);
    if (nb) atomicAdd(&a.status->neighbour_total, nb);
    if (ft) atomicAdd(&a.status->finite_tti, ft);
  }
}

// Thread-per-agent form of the hot kernel (also the streaming kernel of NoLocalPlan-only crowds).
__global__ void __launch_bounds__(128) step_kernel(StepArgs a) {
  if (a.status->failed) return;
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t n_live = *a.n_sorted;
  uint32_t cand = 0, nbc = 0, finite = 0;
  if (i < a.n && i < n_live) {
    const GroupDev& g = a.groups[a.in.grp[i]];
    Self me;
    me.px = a.in.x[i];
    me.py = a.in.y[i];
    me.vx = a.in.vx[i];
    me.vy = a.in.vy[i];
    me.id = a.in.id[i];
    double velx, vely;
    high_level_velocity(a, i, g, me, velx, vely);
    double t_i = RCS_INF, fx = 0.0, fy = 0.0;
    if (g.lp_kind == LP_ZANLUNGO) {
      zanlungo_sequential(a, i, me, g, t_i, fx, fy, nbc, cand);
      finite = t_i != RCS_INF ? 1u : 0u;
      // zanlungo.rs:216
      velx = velx + fx * g.inv_mass;
      vely = vely + fy * g.inv_mass;
    }
    integrate_and_store(a, i, me, velx, vely, t_i, fx, fy, nbc);
  }
  warp_stats(a, cand, nbc, finite);
}

// Trace: neighbour ids in list order (lib.rs:281-286) as CSR, offsets from an exclusive scan of nb_count.
__global__ void trace_neighbours_kernel(StepArgs a, const uint32_t* __restrict__ nb_offsets,
                                        uint64_t* __restrict__ nb_ids) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.n || i >= *a.n_sorted) return;
  const GroupDev& g = a.groups[a.in.grp[i]];
  if (g.lp_kind != LP_ZANLUNGO) return;
  uint32_t o = nb_offsets[i];
  for_each_neighbour(a.grid, a.cell_start, a.in.x, a.in.y, a.in.id, a.in.x[i], a.in.y[i], a.in.id[i], g.eyesight,
                     g.thr2, [&](uint32_t j, double, double, double) { nb_ids[o++] = a.in.id[j]; });
}

// SpatialIndex::get_neighbours_in_radius for arbitrary query points (location_hash_2d.rs:240-258).
// mode 0: count into counts[q]; mode 1: write ids at offsets[q].  No self filter.
__global__ void query_radius_kernel(GridDev g, const uint32_t* __restrict__ cell_start, const double* __restrict__ xs,
                                    const double* __restrict__ ys, const uint64_t* __restrict__ ids, uint32_t nq,
                                    const double* __restrict__ qxy, const double* __restrict__ radius,
                                    const double* __restrict__ thr2, uint32_t* __restrict__ counts,
                                    const uint64_t* __restrict__ offsets, uint64_t* __restrict__ out_ids, int mode) {
  uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= nq) return;
  double px = qxy[2 * q], py = qxy[2 * q + 1];
  uint32_t cnt = 0;
  uint64_t o = mode ? offsets[q] : 0;
  // ids are < 2^63 in practice; ~0 never matches an agent so the self filter is a no-op here
  for_each_neighbour(g, cell_start, xs, ys, ids, px, py, ~0ull, radius[q], thr2[q],
                     [&](uint32_t j, double, double, double) {
                       if (mode) out_ids[o++] = ids[j];
                       cnt++;
                     });
  if (!mode) counts[q] = cnt;
}

__global__ void cell_of_kernel(GridDev g, uint32_t n, const double* __restrict__ xy, long long* __restrict__ out) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint64_t idx;
  out[i] = location_to_index(g, xy[2 * i], xy[2 * i + 1], idx) ? (long long)idx : -1ll;
}

// ---------------------------------------------------------------------------------------------
// id-addressed access (the reference's `agents: HashMap<AgentId, Agent>` view, lib.rs:71)
// ---------------------------------------------------------------------------------------------
__global__ void build_slot_of_id_kernel(uint32_t n, const uint64_t* __restrict__ id, uint32_t* __restrict__ slot_of_id,
                                        uint64_t table_len) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint64_t v = id[i];
  if (v < table_len) slot_of_id[v] = i;
}

// presence[v] = 1 if id v is live (input of the rank scan that yields ascending-id order)
__global__ void presence_kernel(uint64_t table_len, const uint32_t* __restrict__ slot_of_id,
                                uint32_t* __restrict__ presence) {
  uint64_t v = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= table_len) return;
  presence[v] = slot_of_id[v] != 0xffffffffu ? 1u : 0u;
}

// order_by_id[rank] = slot, for rank = number of live ids smaller than v
__global__ void order_by_id_kernel(uint64_t table_len, const uint32_t* __restrict__ slot_of_id,
                                   const uint32_t* __restrict__ rank, uint32_t* __restrict__ order_by_id) {
  uint64_t v = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= table_len) return;
  uint32_t s = slot_of_id[v];
  if (s != 0xffffffffu) order_by_id[rank[v]] = s;
}

// out[k] = src[order[k]] (order == nullptr: identity)
template <class T>
__global__ void gather_kernel(uint32_t n, const uint32_t* __restrict__ order, const T* __restrict__ src,
                              T* __restrict__ out) {
  uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  out[k] = src[order ? order[k] : k];
}

// dst[slot(k)] = src[k] where slot(k) = order[k] (ascending-id addressing) or slot_of_id[ids[k]]
template <class T>
__global__ void scatter_by_id_kernel(uint32_t n, const uint32_t* __restrict__ order, const uint64_t* __restrict__ ids,
                                     const uint32_t* __restrict__ slot_of_id, uint64_t table_len,
                                     const T* __restrict__ src, int src_stride, T* __restrict__ dst,
                                     unsigned int* __restrict__ bad) {
  uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  uint32_t s;
  if (ids) {
    uint64_t v = ids[k];
    s = v < table_len ? slot_of_id[v] : 0xffffffffu;
    if (s == 0xffffffffu) {
      atomicAdd(bad, 1u);
      return;
    }
  } else {
    s = order[k];
  }
  dst[s] = src[(size_t)k * src_stride];
}

__global__ void fill_u32_kernel(uint64_t n, uint32_t* p, uint32_t v) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}
__global__ void fill_f64_kernel(uint64_t n, double* p, double v) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

// marks a step as failed on the device so that later async steps become no-ops (sticky)
__global__ void finish_step_kernel(DevStatus* status) {
  if (status->oob_count) status->failed = 1;
}

// FP64 pipe peak: independent DFMA / DADD chains.
__global__ void fp64_peak_kernel(double* out, int iters, int use_fma) {
  double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6,
         a7 = a0 + 7;
  const double m = 1.0000001, c = 1e-7;
  if (use_fma) {
    for (int i = 0; i < iters; ++i) {
      a0 = __fma_rn(a0, m, c); a1 = __fma_rn(a1, m, c); a2 = __fma_rn(a2, m, c); a3 = __fma_rn(a3, m, c);
      a4 = __fma_rn(a4, m, c); a5 = __fma_rn(a5, m, c); a6 = __fma_rn(a6, m, c); a7 = __fma_rn(a7, m, c);
    }
  } else {
    for (int i = 0; i < iters; ++i) {
      a0 = __dadd_rn(a0, c); a1 = __dadd_rn(a1, c); a2 = __dadd_rn(a2, c); a3 = __dadd_rn(a3, c);
      a4 = __dadd_rn(a4, c); a5 = __dadd_rn(a5, c); a6 = __dadd_rn(a6, c); a7 = __dadd_rn(a7, c);
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

__global__ void flush_l2_kernel(uint4* p, uint64_t n16) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (; i < n16; i += stride) p[i] = make_uint4((uint32_t)i, 1u, 2u, 3u);
}

}  // namespace rcs
