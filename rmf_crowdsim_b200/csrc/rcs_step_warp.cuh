// rcs_step_warp.cuh -- the hot kernel, warp-cooperative form (sm_100a).
//
// step_kernel (rcs_kernels.cuh) gives each agent one thread that walks its own candidates and runs
// time_to_collision / the pair force inline; ncu shows those bodies executing with 3-5 of 32 lanes
// (profiles/r01a_step_kernel_v1.md).  Here a warp owns 32 consecutive agents of the canonical
// (cell, id) order and splits the work in two kinds of stages:
//
//   filter : every lane walks its own agent's candidates (cheap: 2 loads, 5 DP ops, 1 compare) and the
//            (owner lane, neighbour index) pairs that pass the radius + self filter are compacted with
//            warp ballots into a shared-memory pair list;
//   dense  : the list is processed 32 pairs at a time, one pair per lane, so time_to_collision (phase
//            1) and the pair force (phase 2) run with all lanes busy.  t_i is reduced with a shared
//            64-bit atomicMin on the bit pattern (all collision times are >= +0, so integer order ==
//            floating order and min is order-independent); pair forces go to per-pair slots and every
//            owner adds its own slots by following a per-owner chain in append order = the canonical
//            neighbour order, so the sum is bit-identical to the sequential one.
//
// Same arithmetic, same order, same results as step_kernel -- tests compare the two bit for bit.
#pragma once

#include "rcs_kernels.cuh"

namespace rcs {

constexpr int SW_WARPS = 4;      // warps per block
constexpr int SW_PL = 256;       // phase-1 pair list entries per warp
constexpr int SW_PLA = 128;      // phase-2 list A (pairs that need a real force evaluation)
constexpr int SW_PLB = 128;      // phase-2 list B (weight-0 pairs: only the "is it exactly zero" check)
constexpr uint32_t SW_NONE = 0xffffu;

struct WarpShared {
  double px[32], py[32], vx[32], vy[32], pfx[32], pfy[32], ti[32];
  unsigned long long id[32];
  unsigned long long tbits[32];
  double sfx[SW_PLA], sfy[SW_PLA];
  uint32_t lj[SW_PL];
  uint32_t grp[32];
  unsigned int poison[32];
  uint16_t nxt[SW_PLA];
  uint8_t lo[SW_PL];
};

// One agent's walk over `for x in left..=right { for y in bottom..=top }` (location_hash_2d.rs:245-246)
// as a sequence of contiguous slices of the sorted arrays.
struct CandIter {
  int64_t cx, right, bottom, top;
  uint32_t j, e;
  bool valid;

  __device__ __forceinline__ void next_column(const GridDev& g, const uint32_t* __restrict__ cell_start,
                                              uint32_t& cand) {
    valid = false;
    while (++cx <= right) {
      uint64_t c_lo, c_hi;
      if (!column_cell_range(g, cx, bottom, top, c_lo, c_hi)) continue;
      j = cell_start[c_lo];
      e = cell_start[c_hi + 1];
      if (j < e) {
        cand += e - j;
        valid = true;
        return;
      }
    }
  }

  __device__ __forceinline__ void init(const GridDev& g, const uint32_t* __restrict__ cell_start, double radius,
                                       double px, double py, uint32_t& cand) {
    int64_t left;
    get_bounds(g, radius, px, py, left, right, bottom, top);
    if (left < 0) left = 0;
    if (right > g.x_max) right = g.x_max;
    cx = left - 1;
    next_column(g, cell_start, cand);
  }
};

__global__ void __launch_bounds__(32 * SW_WARPS) step_warp_kernel(StepArgs a) {
  if (a.status->failed) return;
  __shared__ WarpShared sh[SW_WARPS];
  WarpShared& w = sh[threadIdx.x >> 5];
  const unsigned lane = threadIdx.x & 31u;
  const unsigned lt_mask = (1u << lane) - 1u;
  const unsigned FULL = 0xffffffffu;
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t n_live = *a.n_sorted;
  const bool active = (i < a.n) && (i < n_live);

  const double* __restrict__ xs = a.in.x;
  const double* __restrict__ ys = a.in.y;
  const double* __restrict__ vxs = a.in.vx;
  const double* __restrict__ vys = a.in.vy;
  const uint64_t* __restrict__ ids = a.in.id;
  const uint32_t* __restrict__ cell_start = a.cell_start;

  double px = 0.0, py = 0.0, vx = 0.0, vy = 0.0, pfx = 0.0, pfy = 0.0, velx = 0.0, vely = 0.0;
  double eyesight = 0.0, thr2 = 0.0, inv_mass = 0.0;
  uint64_t my_id = 0;
  uint32_t grp = 0;
  bool zan = false;
  if (active) {
    grp = a.in.grp[i];
    const GroupDev& g = a.groups[grp];
    px = xs[i];
    py = ys[i];
    vx = vxs[i];
    vy = vys[i];
    my_id = ids[i];
    // high-level planner, lib.rs:263-273
    switch (g.hl_kind) {
      case HL_CONSTANT:
        velx = g.hl_vx;
        vely = g.hl_vy;
        pfx = velx;
        pfy = vely;
        break;
      case HL_PARITY:
        if ((my_id & 1ull) == 0ull) {
          velx = -g.hl_vx;
          vely = -g.hl_vy;
        } else {
          velx = g.hl_vx;
          vely = g.hl_vy;
        }
        pfx = velx;
        pfy = vely;
        break;
      case HL_HOST: {
        double hx = a.in.pvx[i], hy = a.in.pvy[i];
        if (hx == hx) {
          velx = hx;
          vely = hy;
          pfx = hx;
          pfy = hy;
        }
      } break;
      default:
        break;
    }
    zan = g.lp_kind == LP_ZANLUNGO;
    eyesight = g.eyesight;
    thr2 = g.thr2;
    inv_mass = g.inv_mass;
  }
  w.px[lane] = px;
  w.py[lane] = py;
  w.vx[lane] = vx;
  w.vy[lane] = vy;
  w.pfx[lane] = pfx;
  w.pfy[lane] = pfy;
  w.id[lane] = my_id;
  w.grp[lane] = grp;
  w.tbits[lane] = 0x7ff0000000000000ull;
  w.poison[lane] = 0u;
  __syncwarp();

  uint32_t cand = 0, nbc = 0;
  double t_i = RCS_INF, fx = 0.0, fy = 0.0;

  if (__any_sync(FULL, zan)) {
    // ---------------- phase 1: t_i = min over neighbours of time_to_collision (zanlungo.rs:76-91)
    CandIter it;
    it.valid = false;
    if (zan) it.init(a.grid, cell_start, eyesight, px, py, cand);
    uint32_t cnt = 0;
    auto flush_ttc = [&]() {
      __syncwarp();
      for (uint32_t e = lane; e < cnt; e += 32) {
        const uint32_t o = w.lo[e];
        const uint32_t j = w.lj[e];
        const double opx = w.px[o], opy = w.py[o];
        const double dx = xs[j] - opx;
        const double dy = ys[j] - opy;
        const double d2 = dx * dx + dy * dy;
        const double ct = time_to_collision(vxs[j] - w.vx[o], vys[j] - w.vy[o], dx, dy, d2, a.groups[w.grp[o]].rr);
        if (ct < RCS_INF) atomicMin(&w.tbits[o], (unsigned long long)__double_as_longlong(ct));
      }
      __syncwarp();
      cnt = 0;
    };
    while (__any_sync(FULL, it.valid)) {
      bool pass = false;
      uint32_t jj = 0;
      if (it.valid) {
        jj = it.j;
        const double dx = xs[jj] - px;
        const double dy = ys[jj] - py;
        const double d2 = dx * dx + dy * dy;
        pass = (d2 < thr2) && (jj != i);  // strict radius filter (:251) and self filter (lib.rs:284)
        if (++it.j == it.e) it.next_column(a.grid, cell_start, cand);
      }
      const unsigned m = __ballot_sync(FULL, pass);
      if (pass) {
        const uint32_t pos = cnt + __popc(m & lt_mask);
        w.lj[pos] = jj;
        w.lo[pos] = (uint8_t)lane;
        nbc++;
      }
      cnt += __popc(m);
      if (cnt > SW_PL - 32) flush_ttc();
    }
    if (cnt) flush_ttc();
    __syncwarp();
    t_i = __longlong_as_double((long long)w.tbits[lane]);

    // ---------------- phase 2: force = sum over neighbours of compute_agent_force (zanlungo.rs:210-215)
    const bool fin = zan && (t_i != RCS_INF);
    if (__any_sync(FULL, fin)) {
      w.ti[lane] = t_i;
      __syncwarp();
      uint32_t dummy = 0;
      it.valid = false;
      if (fin) it.init(a.grid, cell_start, eyesight, px, py, dummy);
      uint32_t cntA = 0, cntB = 0;
      uint32_t first = SW_NONE, last = SW_NONE;
      auto flush_A = [&]() {
        __syncwarp();
        for (uint32_t e = lane; e < cntA; e += 32) {
          const uint32_t o = w.lo[e] & 31u;
          const uint32_t literal = w.lo[e] >> 5;
          const uint32_t j = w.lj[e];
          PairIn p;
          p.px = w.px[o]; p.py = w.py[o]; p.vx = w.vx[o]; p.vy = w.vy[o];
          p.pfx = w.pfx[o]; p.pfy = w.pfy[o]; p.id = w.id[o];
          p.ox = xs[j]; p.oy = ys[j]; p.ovx = vxs[j]; p.ovy = vys[j];
          const GroupDev& g = a.groups[w.grp[o]];
          double qx, qy;
          if (!literal) {
            pair_force_yield(p, w.ti[o], g, qx, qy);
          } else {
            p.oid = ids[j];
            pair_force_literal(p, w.ti[o], g, qx, qy);
          }
          w.sfx[e] = qx;
          w.sfy[e] = qy;
        }
        __syncwarp();
        // every owner adds its own pairs in append order = canonical neighbour order
        for (uint32_t e = first; e != SW_NONE; e = w.nxt[e]) {
          fx = fx + w.sfx[e];
          fy = fy + w.sfy[e];
        }
        first = last = SW_NONE;
        __syncwarp();
        cntA = 0;
      };
      auto flush_B = [&]() {
        __syncwarp();
        for (uint32_t e = lane; e < cntB; e += 32) {
          const uint32_t o = w.lo[SW_PLA + e];
          const uint32_t j = w.lj[SW_PLA + e];
          PairIn p;
          p.px = w.px[o]; p.py = w.py[o]; p.vx = w.vx[o]; p.vy = w.vy[o];
          p.pfx = w.pfx[o]; p.pfy = w.pfy[o]; p.id = w.id[o];
          p.ox = xs[j]; p.oy = ys[j]; p.ovx = vxs[j]; p.ovy = vys[j];
          if (!pair_force_w0_is_zero(p, w.ti[o])) {
            // contributes NaN or +-0 per component (rcs_math.cuh): NaN is order-independent
            p.oid = ids[j];
            double qx, qy;
            pair_force_literal(p, w.ti[o], a.groups[w.grp[o]], qx, qy);
            unsigned bits = (qx != qx ? 1u : 0u) | (qy != qy ? 2u : 0u);
            if (bits) atomicOr(&w.poison[o], bits);
          }
        }
        __syncwarp();
        cntB = 0;
      };
      const uint32_t w0_fast = active ? a.groups[grp].w0_fast : 0u;
      while (__any_sync(FULL, it.valid)) {
        bool passA = false, passB = false;
        uint32_t jj = 0, literal = 0;
        if (it.valid) {
          jj = it.j;
          const double dx = xs[jj] - px;
          const double dy = ys[jj] - py;
          const double d2 = dx * dx + dy * dy;
          if ((d2 < thr2) && (jj != i)) {
            const uint64_t oid = ids[jj];
            double row;
            if (((my_id | oid) >> 53) == 0ull) row = my_id < oid ? -1.0 : 1.0;
            else row = right_of_way(my_id, oid);
            if (row < 0.0) {
              passA = true;
            } else if (row > 0.0 && w0_fast) {
              passB = true;
            } else {
              passA = true;
              literal = 1;
            }
          }
          if (++it.j == it.e) it.next_column(a.grid, cell_start, dummy);
        }
        const unsigned mA = __ballot_sync(FULL, passA);
        const unsigned mB = __ballot_sync(FULL, passB);
        if (passA) {
          const uint32_t pos = cntA + __popc(mA & lt_mask);
          w.lj[pos] = jj;
          w.lo[pos] = (uint8_t)(lane | (literal << 5));
          w.nxt[pos] = (uint16_t)SW_NONE;
          if (last != SW_NONE) w.nxt[last] = (uint16_t)pos;
          else first = pos;
          last = pos;
        }
        if (passB) {
          const uint32_t pos = SW_PLA + cntB + __popc(mB & lt_mask);
          w.lj[pos] = jj;
          w.lo[pos] = (uint8_t)lane;
        }
        cntA += __popc(mA);
        cntB += __popc(mB);
        if (cntA > SW_PLA - 32) flush_A();
        if (cntB > SW_PLB - 32) flush_B();
      }
      if (cntA) flush_A();
      if (cntB) flush_B();
      __syncwarp();
      const unsigned pz = w.poison[lane];
      if (pz & 1u) fx = fx + __longlong_as_double(0x7ff8000000000000LL);
      if (pz & 2u) fy = fy + __longlong_as_double(0x7ff8000000000000LL);
    }
  }

  if (active) {
    if (zan) {
      // zanlungo.rs:216
      velx = velx + fx * inv_mass;
      vely = vely + fy * inv_mass;
    }
    // explicit Euler, lib.rs:295-297
    const double nx = px + velx * a.dt;
    const double ny = py + vely * a.dt;
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
      atomicMin(&a.status->first_oob_id, (unsigned long long)my_id);
    }
    if (!(isfinite(nx) && isfinite(ny) && isfinite(velx) && isfinite(vely))) atomicAdd(&a.status->nonfinite_count, 1u);
  }
  if (a.collect_stats) {
    unsigned long long c = cand, nb = nbc, ft = (active && zan && t_i != RCS_INF) ? 1ull : 0ull;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
      c += __shfl_down_sync(FULL, c, d);
      nb += __shfl_down_sync(FULL, nb, d);
      ft += __shfl_down_sync(FULL, ft, d);
    }
    if (lane == 0) {
      if (c) atomicAdd(&a.status->candidate_total, c);
      if (nb) atomicAdd(&a.status->neighbour_total, nb);
      if (ft) atomicAdd(&a.status->finite_tti, ft);
    }
  }
}

}  // namespace rcs
