// rcs_inloop.cuh -- the reference's IN-LOOP index semantic, for studying how far it
// is from the deferred contract.  Not on the benchmark path.
//
// The reference updates its spatial index inside the per-agent loop (lib.rs:299): when agent i is processed, every
// agent j that came EARLIER in the iteration order already sits in the index at its NEW position (cell and
// id_to_exact_location, location_hash_2d.rs:126-149), every later agent still at its old one.  The neighbour SET of
// i is decided by those mixed locations; the neighbour STATES handed to the planner are the old ones in every case
// (agents.at(nid), lib.rs:281-286: the agents map is only written at commit).  That is a Gauss-Seidel sweep: new
// position of i <- new positions of earlier agents near i.
//
// Device form: a fixed-point iteration.  Sweep 0 assumes nobody has moved (the deferred step).  Sweep k recomputes every
// agent with the new positions of sweep k-1 for the agents that precede it in the order.  An agent whose predecessors
// within reach are final is final itself, so after sweep k every dependency chain of length <= k is settled; the
// iteration stops when a sweep changes no bit.  The result is exactly what the sequential loop produces for that order.
// Two indices serve a sweep: the canonical sorted arrays of the old positions (cell_start) and a second counting sort
// of the previous sweep's new positions (cell_start_new / perm_new).  A data cell's members are the valid old
// instances (agents not before i) merged with the valid new instances (agents before i) in ascending id, the order
// the oracle uses for the reference's HashSet.
#pragma once

#include "rcs_kernels.cuh"

namespace rcs {

struct InLoopArgs {
  StepArgs a;                        // in = canonical sorted arrays of the OLD state; trace pointers optional
  const unsigned long long* rank;    // position of every sorted slot's agent in the iteration order
  const double2* newpos_prev;        // new positions of the previous sweep (unused in sweep 0)
  const uint32_t* cell_start_new;    // index of newpos_prev: per cell [start, end) into perm_new
  const uint32_t* perm_new;          // sorted slots, ascending id inside a cell
  double2* newpos_next;              // this sweep's results
  double2* newvel;
  uint32_t* changed;                 // agents whose new position differs (bitwise) from the previous sweep's
  uint32_t first;                    // sweep 0: nobody has moved yet
};

// cellid / histogram of the previous sweep's new positions (entries outside the grid are left out: the step fails
// anyway if a FINAL position is out of bounds, lib.rs:299-302)
__global__ void inloop_bin_kernel(GridDev g, uint32_t n, const double2* __restrict__ newpos,
                                  uint32_t* __restrict__ cellid, uint32_t* __restrict__ cell_count) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint64_t idx;
  const double2 p = newpos[i];
  if (location_to_index(g, p.x, p.y, idx)) {
    cellid[i] = (uint32_t)idx;
    atomicAdd(&cell_count[idx], 1u);
  } else {
    cellid[i] = CELL_DEAD;
  }
}

__global__ void inloop_rank_kernel(uint32_t n, const uint64_t* __restrict__ ids,
                                   const uint32_t* __restrict__ pos_of_id, uint64_t table_len,
                                   unsigned long long* __restrict__ rank) {
  uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const uint64_t id = ids[k];
  // no table: ascending id IS the order; ids missing from a caller's order come last, by id
  rank[k] = pos_of_id ? ((id < table_len && pos_of_id[id] != 0xffffffffu) ? pos_of_id[id] : (1ull << 40) + id) : id;
}

__global__ void inloop_order_table_kernel(uint32_t m, const uint64_t* __restrict__ order, uint64_t table_len,
                                          uint32_t* __restrict__ pos_of_id) {
  uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= m) return;
  if (order[k] < table_len) pos_of_id[order[k]] = k;
}

// get_neighbours_in_radius as agent `self` (order position my_rank) sees the index in the middle of the loop.
// f(j) for every neighbour slot j in the reference's order; returns the number of candidates distance-tested.
template <class F>
__device__ __forceinline__ uint32_t for_each_neighbour_inloop(const InLoopArgs& q, double px, double py,
                                                              unsigned long long my_rank, uint32_t self,
                                                              double radius, double thr2, F&& f) {
  const GridDev& g = q.a.grid;
  const double2* __restrict__ pos = q.a.in.pos;
  const uint64_t* __restrict__ ids = q.a.in.id;
  int64_t left, right, bottom, top;
  get_bounds(g, radius, px, py, left, right, bottom, top);
  if (left < 0) left = 0;
  if (right > g.x_max) right = g.x_max;
  uint32_t cand = 0;
  for (int64_t cx = left; cx <= right; ++cx) {
    uint64_t c_lo, c_hi;
    if (!column_cell_range(g, cx, bottom, top, c_lo, c_hi)) continue;
    for (uint64_t c = c_lo; c <= c_hi; ++c) {
      uint32_t po = q.a.cell_start[c];
      const uint32_t eo = q.a.cell_start[c + 1];
      uint32_t pn = 0, en = 0;
      if (!q.first) {
        pn = q.cell_start_new[c];
        en = q.cell_start_new[c + 1];
      }
      for (;;) {
        // next old instance that is still in place (its agent does not precede this one; the agent itself has
        // not been moved yet either) / next new instance of an agent that does
        while (po < eo && !q.first && q.rank[po] < my_rank) ++po;
        while (pn < en && !(q.rank[q.perm_new[pn]] < my_rank)) ++pn;
        const bool ho = po < eo, hn = pn < en;
        if (!ho && !hn) break;
        uint32_t j;
        double2 loc;
        if (ho && (!hn || ids[po] < ids[q.perm_new[pn]])) {
          j = po++;
          loc = pos[j];
        } else {
          j = q.perm_new[pn++];
          loc = q.newpos_prev[j];
        }
        cand++;
        const double dx = loc.x - px;
        const double dy = loc.y - py;
        const double d2 = dx * dx + dy * dy;
        if (d2 < thr2 && j != self) f(j);
      }
    }
  }
  return cand;
}

__global__ void __launch_bounds__(128) step_inloop_kernel(InLoopArgs q) {
  const StepArgs& a = q.a;
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t cand = 0, nbc = 0, finite = 0;
  if (i < a.n) {
    const double2* __restrict__ pos = a.in.pos;
    const double2* __restrict__ vel = a.in.vel;
    const uint64_t* __restrict__ ids = a.in.id;
    const uint32_t grp = a.in.grp[i];
    const GroupDev& g = a.groups[grp];
    Self me;
    const double2 p0 = pos[i], v0 = vel[i];
    me.px = p0.x;
    me.py = p0.y;
    me.vx = v0.x;
    me.vy = v0.y;
    me.id = ids[i];
    me.rwp = 0u;
    double velx, vely;
    high_level_velocity(a, i, g, me, velx, vely);
    double t_i = RCS_INF, fx = 0.0, fy = 0.0;
    if (g.lp_kind == LP_ZANLUNGO) {
      const unsigned long long my_rank = q.rank[i];
      const double rr = g.rr;
      // compute_tti (zanlungo.rs:76-91) on the OLD states of the neighbours found through the mixed index
      cand = for_each_neighbour_inloop(q, me.px, me.py, my_rank, i, g.eyesight, g.thr2, [&](uint32_t j) {
        nbc++;
        const double2 np = pos[j], nv = vel[j];
        const double dx = np.x - me.px;
        const double dy = np.y - me.py;
        const double d2 = dx * dx + dy * dy;
        const double ct = time_to_collision(nv.x - me.vx, nv.y - me.vy, dx, dy, d2, rr);
        if (ct < t_i) t_i = ct;
      });
      if (t_i != RCS_INF) {  // zanlungo.rs:210-215
        finite = 1;
        const OwnerPre pre = owner_precompute(me.px, me.py, me.vx, me.vy, me.pfx, me.pfy, t_i, g);
        const double ti = t_i;
        for_each_neighbour_inloop(q, me.px, me.py, my_rank, i, g.eyesight, g.thr2, [&](uint32_t j) {
          double qx, qy;
          const double2 np = pos[j], nv = vel[j];
          if (pair_force_dispatch(pre, me.px, me.py, me.vx, me.vy, me.pfx, me.pfy, me.id, np.x, np.y, nv.x, nv.y,
                                  ids[j], ti, g, qx, qy)) {
            fx = fx + qx;
            fy = fy + qy;
          }
        });
      }
      velx = velx + fx * g.inv_mass;  // zanlungo.rs:216
      vely = vely + fy * g.inv_mass;
    }
    const double nx = me.px + velx * a.dt;  // lib.rs:295-297
    const double ny = me.py + vely * a.dt;
    const double2 before = q.first ? make_double2(0.0, 0.0) : q.newpos_prev[i];
    const bool same = !q.first && __double_as_longlong(before.x) == __double_as_longlong(nx) &&
                      __double_as_longlong(before.y) == __double_as_longlong(ny);
    if (!same) atomicAdd(q.changed, 1u);
    q.newpos_next[i] = make_double2(nx, ny);
    q.newvel[i] = make_double2(velx, vely);
    if (a.t_i) {
      a.t_i[i] = t_i;
      a.fx[i] = fx;
      a.fy[i] = fy;
      a.nb_count[i] = nbc;
      a.tr_id[i] = me.id;
      a.tr_own[i] = 1u;
    }
  }
  warp_stats(a, cand, nbc, finite);
}

// Trace: the neighbour ids of the converged sweep, CSR (offsets from an exclusive scan of nb_count)
__global__ void trace_neighbours_inloop_kernel(InLoopArgs q, const uint32_t* __restrict__ nb_offsets,
                                               uint64_t* __restrict__ nb_ids) {
  const StepArgs& a = q.a;
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.n) return;
  const GroupDev& g = a.groups[a.in.grp[i]];
  if (g.lp_kind != LP_ZANLUNGO) return;
  uint32_t o = nb_offsets[i];
  const double2 p0 = a.in.pos[i];
  for_each_neighbour_inloop(q, p0.x, p0.y, q.rank[i], i, g.eyesight, g.thr2,
                            [&](uint32_t j) { nb_ids[o++] = a.in.id[j]; });
}

// Commit of the converged sweep (lib.rs:350-359) + add_or_update's bounds check on the final positions
__global__ void inloop_commit_kernel(uint32_t n, GridDev g, const double2* __restrict__ newpos,
                                     const double2* __restrict__ newvel, const uint64_t* __restrict__ ids,
                                     double2* __restrict__ pos, double2* __restrict__ vel, DevStatus* status,
                                     uint32_t write) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double2 p = newpos[i], v = newvel[i];
  if (write) {
    pos[i] = p;
    vel[i] = v;
    return;
  }
  uint64_t idx;
  if (!location_to_index(g, p.x, p.y, idx)) {
    atomicAdd(&status->oob_count, 1u);
    atomicMin(&status->first_oob_id, (unsigned long long)ids[i]);
  }
  if (!(isfinite(p.x) && isfinite(p.y) && isfinite(v.x) && isfinite(v.y))) atomicAdd(&status->nonfinite_count, 1u);
}

}  // namespace rcs
