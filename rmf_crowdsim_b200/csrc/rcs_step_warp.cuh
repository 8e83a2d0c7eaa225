// rcs_step_warp.cuh -- the hot kernel, warp-cooperative form (sm_100a).
//
// step_kernel (rcs_kernels.cuh) gives each agent one thread that walks its own candidates and runs
// time_to_collision / the pair force inline; ncu shows those bodies executing with 3-5 of 32 lanes
// (profiles/r01a_step_kernel_v1.md).  Here a warp owns 32 consecutive agents of the canonical
// (cell, id) order and splits the work in two kinds of stages:
//
//   filter : every lane walks its own agent's candidates -- at most three contiguous slices of the
//            sorted arrays, one per stencil column -- with a flat counter (2 loads, 5 DP ops, 1 compare
//            per candidate).  Pairs that pass the strict radius test + self filter are (a) recorded in a
//            per-lane 128-bit mask for the force pass and (b) compacted with warp ballots into a
//            shared-memory (owner lane, neighbour index) list;
//   dense  : the list is processed one pair per lane, so time_to_collision (phase 1) and the pair force
//            (phase 2) run with all lanes busy.  t_i is reduced with a shared 64-bit atomicMin on the bit
//            pattern (all collision times are >= +0, so integer order == floating order and min is
//            order-independent); pair forces go to per-pair slots and every owner adds its own slots by
//            following a per-owner chain in append order = the canonical neighbour order, so the sum is
//            bit-identical to the sequential one.
//
// Agents whose stencil has more than 3 columns (eyesight > cell size) or more than 128 candidates are put
// on a device-side list and finished by step_slow_kernel with the sequential routine.  Same arithmetic,
// same order, same results as step_kernel -- tests compare the two bit for bit.
#pragma once

#include "rcs_kernels.cuh"

namespace rcs {

constexpr int SW_WARPS = 4;       // warps per block
constexpr int SW_PL = 512;        // phase-1 pair list entries per warp
constexpr int SW_PLA = 128;       // phase-2 list A (pairs that need a real force evaluation)
constexpr int SW_PLB = 128;       // phase-2 list B (weight-0 pairs: only the "is it exactly zero" check)
constexpr uint32_t SW_MAXC = 128; // candidates per agent on the cooperative path (mask width)
constexpr uint32_t SW_NONE = 0xffffu;

struct WarpShared {
  double px[32], py[32], vx[32], vy[32], pfx[32], pfy[32], ti[32], rr[32];
  double futx[32], futy[32], mag[32], mvx[32], mvy[32], f0x[32], f0y[32];  // OwnerPre
  unsigned long long id[32];
  unsigned long long tbits[32];
  double sfx[SW_PLA], sfy[SW_PLA];
  uint32_t lj[SW_PL];
  uint32_t grp[32];
  unsigned int poison[32];
  uint16_t nxt[SW_PLA];
  uint8_t lo[SW_PL];
};

__global__ void __launch_bounds__(32 * SW_WARPS, 4) step_warp_kernel(StepArgs a) {
  if (a.status->failed) return;
  __shared__ WarpShared sh[SW_WARPS];
  WarpShared& w = sh[threadIdx.x >> 5];
  const unsigned lane = threadIdx.x & 31u;
  const unsigned lt_mask = (1u << lane) - 1u;
  const unsigned FULL = 0xffffffffu;
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t n_live = *a.n_sorted;
  bool active = (i < a.n) && (i < n_live);

  const double* __restrict__ xs = a.in.x;
  const double* __restrict__ ys = a.in.y;
  const double* __restrict__ vxs = a.in.vx;
  const double* __restrict__ vys = a.in.vy;
  const uint64_t* __restrict__ ids = a.in.id;
  const uint32_t* __restrict__ cell_start = a.cell_start;

  Self me;
  me.px = me.py = me.vx = me.vy = me.pfx = me.pfy = 0.0;
  me.id = 0;
  double velx = 0.0, vely = 0.0, thr2 = 0.0, rr = 0.0;
  uint32_t grp = 0, role = ROLE_PASSIVE;
  bool zan = false;
  // candidate slices of this lane (cooperative path): flat index t -> j = t + (off0 | off1 | off2)
  uint32_t off0 = 0, off1 = 0, off2 = 0, len0 = 0, len01 = 0, total = 0;
  bool fast = false;
  uint32_t cand = 0, nbc = 0;
  double t_i = RCS_INF, fx = 0.0, fy = 0.0;

  if (active) {
    role = agent_role(a, i);
    if (role == ROLE_PASSIVE) {
      active = false;
      if (a.keep) a.keep[i] = 0u;
    }
  } else if (a.keep && i < a.n) {
    a.keep[i] = 0u;
  }
  if (active) {
    grp = a.in.grp[i];
    const GroupDev& g = a.groups[grp];
    me.px = xs[i];
    me.py = ys[i];
    me.vx = vxs[i];
    me.vy = vys[i];
    me.id = ids[i];
    high_level_velocity(a, i, g, me, velx, vely);
    zan = g.lp_kind == LP_ZANLUNGO;
    thr2 = g.thr2;
    rr = g.rr;
    if (zan) {
      int64_t left, right, bottom, top;
      get_bounds(a.grid, g.eyesight, me.px, me.py, left, right, bottom, top);
      if (left < 0) left = 0;
      if (right > a.grid.x_max) right = a.grid.x_max;
      fast = (right - left) <= 2;
      if (fast) {
        uint32_t s[3] = {0, 0, 0}, l[3] = {0, 0, 0};
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          uint64_t c_lo, c_hi;
          if (left + k <= right && column_cell_range(a.grid, left + k, bottom, top, c_lo, c_hi)) {
            s[k] = cell_start[c_lo];
            l[k] = cell_start[c_hi + 1] - s[k];
          }
        }
        len0 = l[0];
        len01 = l[0] + l[1];
        total = len01 + l[2];
        off0 = s[0];
        off1 = s[1] - len0;
        off2 = s[2] - len01;
        fast = total <= SW_MAXC;
      }
      if (!fast) {
        // wide stencil or crowded cells: this agent is finished by step_slow_kernel (sequential routine)
        total = 0;
        active = false;
        zan = false;
        a.slow_list[atomicAdd(&a.status->slow_count, 1u)] = i;
      } else {
        cand = total;
      }
    }
  }
  w.px[lane] = me.px;
  w.py[lane] = me.py;
  w.vx[lane] = me.vx;
  w.vy[lane] = me.vy;
  w.pfx[lane] = me.pfx;
  w.pfy[lane] = me.pfy;
  w.rr[lane] = rr;
  w.id[lane] = me.id;
  w.grp[lane] = grp;
  w.tbits[lane] = 0x7ff0000000000000ull;
  w.poison[lane] = 0u;
  __syncwarp();

  const uint32_t maxtot = __reduce_max_sync(FULL, total);
  if (maxtot) {
    // ---------------- phase 1: t_i = min over neighbours of time_to_collision (zanlungo.rs:76-91)
    uint32_t nm0 = 0, nm1 = 0, nm2 = 0, nm3 = 0;  // which of my (<= 128) candidates are neighbours
    uint32_t cnt = 0;
    auto flush_ttc = [&]() {
      __syncwarp();
      for (uint32_t e = lane; e < cnt; e += 32) {
        const uint32_t o = w.lo[e];
        const uint32_t j = w.lj[e];
        const double dx = xs[j] - w.px[o];
        const double dy = ys[j] - w.py[o];
        const double d2 = dx * dx + dy * dy;
        const double ct = time_to_collision(vxs[j] - w.vx[o], vys[j] - w.vy[o], dx, dy, d2, w.rr[o]);
        if (ct < RCS_INF) atomicMin(&w.tbits[o], (unsigned long long)__double_as_longlong(ct));
      }
      __syncwarp();
      cnt = 0;
    };
    for (uint32_t tb = 0; tb < maxtot; tb += 32) {  // tb is warp-uniform: one mask word per 32 candidates
      uint32_t word = 0;
      const uint32_t tend = min(maxtot, tb + 32);
      for (uint32_t t = tb; t < tend; ++t) {
        bool pass = false;
        uint32_t j = 0;
        if (t < total) {
          j = t + (t < len0 ? off0 : (t < len01 ? off1 : off2));
          const double dx = xs[j] - me.px;
          const double dy = ys[j] - me.py;
          const double d2 = dx * dx + dy * dy;
          pass = (d2 < thr2) && (j != i);  // strict radius filter (:251) and self filter (lib.rs:284)
        }
        const unsigned m = __ballot_sync(FULL, pass);
        if (pass) {
          word |= 1u << (t - tb);
          const uint32_t pos = cnt + __popc(m & lt_mask);
          w.lj[pos] = j;
          w.lo[pos] = (uint8_t)lane;
        }
        cnt += __popc(m);
        if (cnt > SW_PL - 32) flush_ttc();
      }
      nbc += __popc(word);
      if (tb == 0) nm0 = word;
      else if (tb == 32) nm1 = word;
      else if (tb == 64) nm2 = word;
      else nm3 = word;
    }
    if (cnt) flush_ttc();
    __syncwarp();
    if (fast) t_i = __longlong_as_double((long long)w.tbits[lane]);

    // ---------------- phase 2: force = sum over neighbours of compute_agent_force (zanlungo.rs:210-215)
    const bool fin = fast && (t_i != RCS_INF);
    if (__any_sync(FULL, fin)) {
      if (fin) {
        const GroupDev& g = a.groups[grp];
        const OwnerPre pre = owner_precompute(me.px, me.py, me.vx, me.vy, me.pfx, me.pfy, t_i, g);
        w.ti[lane] = t_i;
        w.futx[lane] = pre.futx;
        w.futy[lane] = pre.futy;
        w.mag[lane] = pre.mag;
        w.mvx[lane] = pre.mvx;
        w.mvy[lane] = pre.mvy;
        w.f0x[lane] = pre.f0x;
        w.f0y[lane] = pre.f0y;
      } else {
        nm0 = nm1 = nm2 = nm3 = 0u;
      }
      __syncwarp();
      uint32_t cntA = 0, cntB = 0;
      uint32_t first = SW_NONE, last = SW_NONE;
      auto load_pre = [&](uint32_t o) {
        OwnerPre p;
        p.futx = w.futx[o]; p.futy = w.futy[o]; p.mag = w.mag[o];
        p.mvx = w.mvx[o]; p.mvy = w.mvy[o]; p.f0x = w.f0x[o]; p.f0y = w.f0y[o];
        return p;
      };
      auto literal = [&](uint32_t o, uint32_t j, double& qx, double& qy) {
        PairIn p;
        p.px = w.px[o]; p.py = w.py[o]; p.vx = w.vx[o]; p.vy = w.vy[o];
        p.pfx = w.pfx[o]; p.pfy = w.pfy[o]; p.id = w.id[o];
        p.ox = xs[j]; p.oy = ys[j]; p.ovx = vxs[j]; p.ovy = vys[j]; p.oid = ids[j];
        pair_force_literal(p, w.ti[o], a.groups[w.grp[o]], qx, qy);
      };
      auto flush_A = [&]() {
        __syncwarp();
        for (uint32_t e = lane; e < cntA; e += 32) {
          const uint32_t o = w.lo[e] & 31u;
          const uint32_t j = w.lj[e];
          double qx, qy;
          if ((w.lo[e] >> 5) == 0u) {
            pair_force_yield(load_pre(o), w.px[o], w.py[o], w.vx[o], w.vy[o], xs[j], ys[j], vxs[j], vys[j], w.ti[o],
                             a.groups[w.grp[o]], qx, qy);
          } else {
            literal(o, j, qx, qy);
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
          if (!pair_force_w0_is_zero(load_pre(o), xs[j], ys[j], vxs[j], vys[j], w.ti[o])) {
            // contributes NaN or +-0 per component (rcs_math.cuh): NaN is order-independent
            double qx, qy;
            literal(o, j, qx, qy);
            const unsigned bits = (qx != qx ? 1u : 0u) | (qy != qy ? 2u : 0u);
            if (bits) atomicOr(&w.poison[o], bits);
          }
        }
        __syncwarp();
        cntB = 0;
      };
      const uint32_t w0_fast = fin ? a.groups[grp].w0_fast : 0u;
      while (__any_sync(FULL, (nm0 | nm1 | nm2 | nm3) != 0u)) {
        bool passA = false, passB = false;
        uint32_t j = 0, lit = 0;
        if ((nm0 | nm1 | nm2 | nm3) != 0u) {
          uint32_t t;
          if (nm0) {
            t = __ffs(nm0) - 1;
            nm0 &= nm0 - 1u;
          } else if (nm1) {
            t = 32 + __ffs(nm1) - 1;
            nm1 &= nm1 - 1u;
          } else if (nm2) {
            t = 64 + __ffs(nm2) - 1;
            nm2 &= nm2 - 1u;
          } else {
            t = 96 + __ffs(nm3) - 1;
            nm3 &= nm3 - 1u;
          }
          j = t + (t < len0 ? off0 : (t < len01 ? off1 : off2));
          const uint64_t oid = ids[j];
          double row;
          if (((me.id | oid) >> 53) == 0ull) row = me.id < oid ? -1.0 : 1.0;
          else row = right_of_way(me.id, oid);
          if (row < 0.0) {
            passA = true;
          } else if (row > 0.0 && w0_fast) {
            passB = true;
          } else {
            passA = true;
            lit = 1;
          }
        }
        const unsigned mA = __ballot_sync(FULL, passA);
        const unsigned mB = __ballot_sync(FULL, passB);
        if (passA) {
          const uint32_t pos = cntA + __popc(mA & lt_mask);
          w.lj[pos] = j;
          w.lo[pos] = (uint8_t)(lane | (lit << 5));
          w.nxt[pos] = (uint16_t)SW_NONE;
          if (last != SW_NONE) w.nxt[last] = (uint16_t)pos;
          else first = pos;
          last = pos;
        }
        if (passB) {
          const uint32_t pos = SW_PLA + cntB + __popc(mB & lt_mask);
          w.lj[pos] = j;
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
    const GroupDev& g = a.groups[grp];
    if (zan) {
      // zanlungo.rs:216
      velx = velx + fx * g.inv_mass;
      vely = vely + fy * g.inv_mass;
    }
    integrate_and_store(a, i, me, g, grp, role, velx, vely, t_i, fx, fy, nbc);
  }
  const bool own = active && role == ROLE_OWN;
  warp_stats(a, own ? cand : 0u, own ? nbc : 0u, (own && zan && t_i != RCS_INF) ? 1u : 0u);
}

}  // namespace rcs
