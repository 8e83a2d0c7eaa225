// flat_parallel.cpp -- TEST / MEASUREMENT INFRASTRUCTURE ONLY, and NOT the reference: what a good multi-threaded CPU
// implementation of the same step would look like, to put the GPU numbers next to something stronger than the
// reference's single-threaded HashMap loop ("optional extra row").
//
// Same contract as the CUDA path (deferred index: every query sees start-of-step positions; canonical neighbour
// order: cells x-major then y, ascending id inside a cell) and the SAME arithmetic: the radius test is the literal
// norm(...) < radius of location_hash_2d.rs:251 and the planner is the oracle's Zanlungo class (compute_tti,
// compute_agent_force -- the restatement of zanlungo.rs).  Only the data structures differ: flat arrays, a counting
// sort into cells instead of HashMap / HashSet, no per-agent allocation, and the agents are split over threads (they
// are independent under the deferred contract).  The result is bit-identical to the oracle's deferred mode; a test
// checks that.
#include <algorithm>
#include <cstring>
#include <thread>

#include "crowdsim_oracle.hpp"

using namespace orc;

namespace {

struct Grid {
  double res, offx, offy;
  uint64_t nx, len;
  // location_hash_2d.rs:54-66
  bool insert_cell(Vec2 p, uint64_t& idx) const {
    uint64_t x_idx = f64_as_usize((p.x - offx) / res);
    uint64_t y_idx = f64_as_usize((p.y - offy) / res);
    idx = x_idx * nx + y_idx;
    return idx < len;
  }
  // location_hash_2d.rs:68-72
  int64_t qx(double v) const { return f64_as_i64(std::floor((v - offx) / res)); }
  int64_t qy(double v) const { return f64_as_i64(std::floor((v - offy) / res)); }
};

}  // namespace

namespace {
int flat_step_impl(uint64_t n, double* x, double* y, double* vx, double* vy, double width, double height, double cell,
                   double offx, double offy, const double* zan, double eyesight, double hl_vx, double hl_vy,
                   uint64_t secs, uint32_t nanos, int threads, double* t_i_out, uint32_t* nbc_out, double* fx_out,
                   double* fy_out, const uint64_t* ids);
}  // namespace

extern "C" {

// One deferred step of n agents (ids 0..n-1 = array index) with one Zanlungo planner and the parity high-level rule
// (even id -> -v, odd id -> +v; rmf_crowdsim_viz/src/main.rs:26-29).  x, y, vx, vy are updated in place; t_i_out may
// be null.  Returns 0, or 1 for "Index out of bounds" (nothing is written then).
int orc_flat_step(uint64_t n, double* x, double* y, double* vx, double* vy, double width, double height, double cell,
                  double offx, double offy, const double* zan /*6*/, double eyesight, double hl_vx, double hl_vy,
                  uint64_t secs, uint32_t nanos, int threads, double* t_i_out) {
  return flat_step_impl(n, x, y, vx, vy, width, height, cell, offx, offy, zan, eyesight, hl_vx, hl_vy, secs, nanos,
                        threads, t_i_out, nullptr, nullptr, nullptr, nullptr);
}

// The same step with the per-agent trace the benchmark-size parity tests compare: neighbour-list length (after the
// self filter, lib.rs:284), t_i and the summed force (zanlungo.rs:208-215).  Any output may be null.  ids (may be
// null: id = array index) gives the agent ids, strictly ascending, so that a window cut out of a larger crowd keeps
// its right-of-way priorities (zanlungo.rs:94), its parity rule and its in-cell order.
int orc_flat_step_trace(uint64_t n, double* x, double* y, double* vx, double* vy, double width, double height,
                        double cell, double offx, double offy, const double* zan /*6*/, double eyesight, double hl_vx,
                        double hl_vy, uint64_t secs, uint32_t nanos, int threads, double* t_i_out, uint32_t* nbc_out,
                        double* fx_out, double* fy_out, const uint64_t* ids) {
  for (uint64_t i = 1; ids && i < n; ++i)
    if (ids[i] <= ids[i - 1]) return 2;
  return flat_step_impl(n, x, y, vx, vy, width, height, cell, offx, offy, zan, eyesight, hl_vx, hl_vy, secs, nanos,
                        threads, t_i_out, nbc_out, fx_out, fy_out, ids);
}

}  // extern "C"

namespace {

int flat_step_impl(uint64_t n, double* x, double* y, double* vx, double* vy, double width, double height, double cell,
                   double offx, double offy, const double* zan, double eyesight, double hl_vx, double hl_vy,
                   uint64_t secs, uint32_t nanos, int threads, double* t_i_out, uint32_t* nbc_out, double* fx_out,
                   double* fy_out, const uint64_t* ids) {
  Grid g{cell, offx, offy, f64_as_usize(width / cell), 0};
  g.len = g.nx * f64_as_usize(height / cell);
  const double dt = Duration{secs, nanos}.as_secs_f64();
  if (threads < 1) threads = 1;

  // index: counting sort by insert cell, ascending id inside a cell (ids are visited in order, so the scatter is stable)
  std::vector<uint32_t> cell_of(n), start(g.len + 1, 0), order(n);
  for (uint64_t i = 0; i < n; ++i) {
    uint64_t idx;
    if (!g.insert_cell({x[i], y[i]}, idx)) return 1;
    cell_of[i] = static_cast<uint32_t>(idx);
    start[idx + 1] += 1;
  }
  for (uint64_t c = 0; c < g.len; ++c) start[c + 1] += start[c];
  {
    std::vector<uint32_t> cur(start.begin(), start.end() - 1);
    for (uint64_t i = 0; i < n; ++i) order[cur[cell_of[i]]++] = static_cast<uint32_t>(i);
  }

  std::vector<double> nxv(n), nyv(n), nvx(n), nvy(n);
  std::vector<int> oob(threads, 0);
  auto work = [&](int t) {
    const Zanlungo z(zan[0], zan[1], zan[2], zan[3], zan[4], zan[5]);  // per thread: it records its last t_i
    std::vector<Agent> nearby;  // reused: no allocation per agent
    nearby.reserve(64);
    const uint64_t lo = n * t / threads, hi = n * (t + 1) / threads;
    for (uint64_t i = lo; i < hi; ++i) {
      Agent me;
      me.agent_id = ids ? ids[i] : i;
      me.position = {x[i], y[i]};
      me.velocity = {vx[i], vy[i]};
      me.eyesight_range = eyesight;
      const Vec2 pref = (me.agent_id % 2 == 0) ? Vec2{-hl_vx, -hl_vy} : Vec2{hl_vx, hl_vy};
      me.preferred_vel = pref;  // lib.rs:271
      // get_neighbours_in_radius, location_hash_2d.rs:240-258 (+ self filter lib.rs:284); neighbours keep pref = 0
      nearby.clear();
      const int64_t right = g.qx(me.position.x + eyesight), left = g.qx(me.position.x - eyesight);
      const int64_t top = g.qy(me.position.y + eyesight), bottom = g.qy(me.position.y - eyesight);
      for (int64_t cx = left; cx <= right; ++cx) {
        for (int64_t cy = bottom; cy <= top; ++cy) {
          if (cx < 0 || cy < 0) continue;
          const uint64_t c = static_cast<uint64_t>(cx) * g.nx + static_cast<uint64_t>(cy);
          if (c >= g.len) continue;
          for (uint32_t k = start[c]; k < start[c + 1]; ++k) {
            const uint32_t j = order[k];
            const Vec2 pj{x[j], y[j]};
            if (norm(pj - me.position) < eyesight && j != i) {
              Agent o;
              o.agent_id = ids ? ids[j] : j;
              o.position = pj;
              o.velocity = {vx[j], vy[j]};
              nearby.push_back(o);
            }
          }
        }
      }
      const Vec2 vel = z.get_desired_velocity(me, nearby, pref);  // zanlungo.rs:201-218
      if (t_i_out) t_i_out[i] = z.last_t_i;
      if (nbc_out) nbc_out[i] = static_cast<uint32_t>(nearby.size());
      if (fx_out) fx_out[i] = z.last_force.x;
      if (fy_out) fy_out[i] = z.last_force.y;
      const Vec2 np = me.position + vel * dt;  // lib.rs:295-297
      uint64_t idx;
      if (!g.insert_cell(np, idx)) oob[t] = 1;  // lib.rs:299-302
      nxv[i] = np.x;
      nyv[i] = np.y;
      nvx[i] = vel.x;
      nvy[i] = vel.y;
    }
  };
  std::vector<std::thread> pool;
  for (int t = 1; t < threads; ++t) pool.emplace_back(work, t);
  work(0);
  for (auto& th : pool) th.join();
  for (int t = 0; t < threads; ++t)
    if (oob[t]) return 1;
  std::memcpy(x, nxv.data(), n * sizeof(double));
  std::memcpy(y, nyv.data(), n * sizeof(double));
  std::memcpy(vx, nvx.data(), n * sizeof(double));
  std::memcpy(vy, nvy.data(), n * sizeof(double));
  return 0;
}

}  // namespace
