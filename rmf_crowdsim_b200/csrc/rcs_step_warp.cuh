// rcs_step_warp.cuh -- the hot kernel, warp-cooperative form (sm_100a).
//
// step_kernel (rcs_kernels.cuh) gives each agent one thread that walks its own candidates and runs
// time_to_collision / the pair force inline; ncu shows those bodies executing with 3-5 of 32 lanes
// (profiles/r01a_step_kernel_v1.md).  Here a warp owns 32 consecutive agents of the canonical (cell, id) order.
// The work is split by cost so that cheap, frequent tests stay per lane with no bookkeeping and only the rare,
// expensive bodies are compacted across the warp (profiles/r01b_step_warp_kernel_v3.md explains why):
//
//   1 filter   every lane walks its own candidates -- three contiguous slices of the sorted arrays, one per
//              stencil column -- and records the strict radius test (location_hash_2d.rs:251) + self filter
//              (lib.rs:284) as one 32-bit mask per slice.  2 loads, 5 DP ops, 2 compares per candidate.
//   2 t_i      every lane walks its own neighbour bits and evaluates the division-free half of
//              time_to_collision (zanlungo.rs:49-60: a, b, c, discriminant).  Only pairs that can return a
//              finite time (a > 0, discriminant >= 0, larger root possibly positive: ~15 % of the neighbours)
//              are compacted with a ballot into a shared pair list; the list is processed one pair per lane with
//              the literal routine (sqrt, two divisions, root selection) and reduced per owner with a 64-bit
//              shared atomicMin on the bit pattern (collision times are >= +0, so integer order == float order,
//              and min is order-independent).  The same walk records which neighbours have the higher id.
//   3 force    owners with a finite t_i split their neighbour bits into yield pairs (own id lower: weight 2,
//              real force, zanlungo.rs:93-170) and weight-0 pairs (provably (+-0, +-0) in the common case,
//              rcs_math.cuh).  Both kinds are compacted into their own lists and evaluated one pair per lane.
//              Every owner's pairs take one contiguous list segment (offsets from a warp scan of the per-lane
//              counts) in canonical neighbour order; pair forces go to per-pair slots and every owner adds its own
//              segment front to back, so the sum is bit-identical to the sequential one.
//
// Agents whose stencil has more than 3 columns (eyesight > cell size) or a slice with more than 32 candidates are
// put on a device-side list and finished by step_slow_kernel with the sequential routine.  Same arithmetic, same
// order, same results as step_kernel -- tests compare the two bit for bit.
#pragma once

#include "rcs_kernels.cuh"

namespace rcs {

constexpr int SW_WARPS = 4;            // warps per block
constexpr uint32_t SW_CAP = 128;       // stage-3 pair lists (A: evaluate, B: prove zero); >= 3 * SW_SLICE_MAX
constexpr uint32_t SW_SLICE_MAX = 32;  // candidates per stencil column on the cooperative path (mask width)

struct WarpShared {
  double px[32], py[32], vx[32], vy[32], rr[32];                              // owners (stages 2 and 3)
  double pfx[32], pfy[32], ti[32];
  double futx[32], futy[32], mag[32], mvx[32], mvy[32], f0x[32], f0y[32];     // OwnerPre (stage 3)
  unsigned long long id[32];
  unsigned long long tbits[32];
  double sfx[SW_CAP], sfy[SW_CAP];     // pair forces of list A
  uint32_t lj[2 * SW_CAP];             // neighbour slot: list A (and the stage-2 hit list) | list B
  uint32_t grp[32];
  unsigned int poison[32];
  uint32_t hcnt;                       // entries in the stage-2 hit list
  uint8_t lo[2 * SW_CAP];              // owner lane: list A | list B
};

__global__ void __launch_bounds__(32 * SW_WARPS, 6) step_warp_kernel(StepArgs a) {
  __shared__ WarpShared sh[SW_WARPS];
  WarpShared& w = sh[threadIdx.x >> 5];
  const unsigned lane = threadIdx.x & 31u;
  const unsigned FULL = 0xffffffffu;
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;

  const double2* __restrict__ pos = a.in.pos;
  const double2* __restrict__ vel = a.in.vel;
  const uint64_t* __restrict__ ids = a.in.id;

  // This warp's own rows go out first, before the status words are even looked at: every array holds at least
  // a.n entries, so the loads are safe for i < a.n and their (DRAM) latency overlaps the dependent checks below.
  Self me;
  me.px = me.py = me.vx = me.vy = me.pfx = me.pfy = 0.0;
  me.id = 0;
  me.rwp = 0u;
  uint32_t grp = 0, wp_in = 0;
  uint4 sl = make_uint4(0u, 0u, 0u, 0u);
  const bool inb = i < a.n;
  if (inb) {
    const double2 p0 = pos[i], v0 = vel[i];
    me.px = p0.x;
    me.py = p0.y;
    me.vx = v0.x;
    me.vy = v0.y;
    me.id = ids[i];
    grp = a.in.grp[i];
    wp_in = a.in.wp[i];
    sl = a.slices[i];
  }
  if (a.status->failed) return;
  const uint32_t n_live = *a.n_sorted;
  bool active = inb && (i < n_live);

  double velx = 0.0, vely = 0.0, thr2 = 0.0, rr = 0.0;
  uint32_t role = ROLE_PASSIVE;
  bool zan = false;
  uint32_t s0 = 0, s1 = 0, s2 = 0, l0 = 0, l1 = 0, l2 = 0;  // candidate slices of this lane (cooperative path)
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
    const GroupDev& g = a.groups[grp];
    high_level_velocity(a, i, g, me, velx, vely);
    zan = g.lp_kind == LP_ZANLUNGO;
    thr2 = g.thr2;
    rr = g.rr;
    if (zan) {
      // candidate slices of the radius query, prepared by gather_sorted_kernel.  ids >= 2^53 round when they
      // become priorities (zanlungo.rs:94) and groups whose weight-0 pairs cannot be proven zero need the literal
      // routine for every pair: both are left to the sequential kernel, like wide or crowded stencils
      s0 = sl.x; s1 = sl.y; s2 = sl.z;
      l0 = sl.w & 0xffu; l1 = (sl.w >> 8) & 0xffu; l2 = (sl.w >> 16) & 0xffu;
      fast = (sl.w >> 24) != 0u && g.w0_fast && (me.id >> 53) == 0ull && l0 <= SW_SLICE_MAX &&
             l1 <= SW_SLICE_MAX && l2 <= SW_SLICE_MAX;
      if (!fast) {
        // wide stencil or crowded cells: this agent is finished by step_slow_kernel (sequential routine)
        l0 = l1 = l2 = 0;
        active = false;
        zan = false;
        a.slow_list[atomicAdd(&a.status->slow_count, 1u)] = i;
      } else {
        cand = l0 + l1 + l2;
      }
    }
  }

  // ---------------- stage 1: radius filter, one mask per stencil column
  uint32_t m0 = 0, m1 = 0, m2 = 0;
  {
    const uint32_t x0 = __reduce_max_sync(FULL, l0), x1 = __reduce_max_sync(FULL, l1),
                   x2 = __reduce_max_sync(FULL, l2);
    // Branch-free body: a lane that has run out of candidates re-reads its own slot, which the self filter
    // rejects; the loads of the unrolled iterations are independent and go out together.
    const uint32_t iself = active ? i : 0u;  /* lanes without an agent have L = 0 and thr2 = 0 */
#define RCS_FILTER_SLICE(MX, S, L, M)                                  \
    _Pragma("unroll 4")                                                \
    for (uint32_t t = 0; t < (MX); ++t) {                              \
      const uint32_t j = (t < (L)) ? (S) + t : iself;                  \
      const double2 c = pos[j];                                        \
      const double dx = c.x - me.px;                                   \
      const double dy = c.y - me.py;                                   \
      const double d2 = dx * dx + dy * dy;                             \
      (M) |= ((d2 < thr2) && (j != i)) ? (1u << t) : 0u;               \
    }
    RCS_FILTER_SLICE(x0, s0, l0, m0)
    RCS_FILTER_SLICE(x1, s1, l1, m1)
    RCS_FILTER_SLICE(x2, s2, l2, m2)
#undef RCS_FILTER_SLICE
  }
  nbc = __popc(m0) + __popc(m1) + __popc(m2);

  if (__any_sync(FULL, (m0 | m1 | m2) != 0u)) {
    w.px[lane] = me.px;
    w.py[lane] = me.py;
    w.vx[lane] = me.vx;
    w.vy[lane] = me.vy;
    w.rr[lane] = rr;
    w.tbits[lane] = 0x7ff0000000000000ull;
    if (lane == 0) w.hcnt = 0u;
    __syncwarp();

    // ---------------- stage 2: t_i = min over neighbours of time_to_collision (zanlungo.rs:76-91)
    // neighbours with the higher id: this agent yields to them (right_of_way = -1; exact for own ids < 2^53)
    uint32_t y0 = 0, y1 = 0, y2 = 0;
    {
      uint32_t it = 0;
      auto flush_hits = [&]() {  // entered after a __syncwarp()
        const uint32_t cnt = *(volatile uint32_t*)&w.hcnt;
        for (uint32_t e = lane; e < cnt; e += 32) {
          const uint32_t o = w.lo[e];
          const uint32_t j = w.lj[e];
          const double2 c = pos[j], cv = vel[j];
          const double dx = c.x - w.px[o];
          const double dy = c.y - w.py[o];
          const double d2 = dx * dx + dy * dy;
          const double ct = time_to_collision(cv.x - w.vx[o], cv.y - w.vy[o], dx, dy, d2, w.rr[o]);
          if (ct < RCS_INF) atomicMin(&w.tbits[o], (unsigned long long)__double_as_longlong(ct));
        }
        __syncwarp();
        if (lane == 0) w.hcnt = 0u;
        __syncwarp();
      };
      // (word, slice start) queue of this lane; empty words are popped with predicated moves, so the walk over
      // the three slices stays free of divergent branches.  Two neighbours are taken per iteration: their loads
      // and arithmetic are independent, which hides half of the load and FP64 latency at this occupancy.
      uint32_t bits = m0, base = s0, nb1 = m1, ns1 = s1, nb2 = m2, ns2 = s2, k = 0;
      const uint32_t iself = active ? i : 0u;
      auto take = [&](uint32_t& j, uint32_t& t, uint32_t& kk) -> bool {
        const bool empty = bits == 0u;  // one pop per take: a lane with two empty words in a row idles once
        bits = empty ? nb1 : bits;
        base = empty ? ns1 : base;
        nb1 = empty ? nb2 : nb1;
        ns1 = empty ? ns2 : ns1;
        nb2 = empty ? 0u : nb2;
        k += empty ? 1u : 0u;
        const bool v = bits != 0u;
        t = v ? (uint32_t)(__ffs(bits) - 1) : 0u;
        bits &= bits - 1u;          // 0 stays 0
        j = v ? base + t : iself;   // always a valid slot: no branch around the loads
        kk = k;
        return v;
      };
      // the division-free half of time_to_collision for neighbour j; same operations as rcs_math.cuh
      auto probe = [&](bool v, uint32_t j, uint32_t t, uint32_t kk) -> bool {
        const uint32_t bit = (v && me.id < ids[j]) ? (1u << t) : 0u;
        y0 |= (kk == 0u) ? bit : 0u;
        y1 |= (kk == 1u) ? bit : 0u;
        y2 |= (kk == 2u) ? bit : 0u;
        const double2 c = pos[j], cv = vel[j];
        const double dx = c.x - me.px;
        const double dy = c.y - me.py;
        const double rvx = cv.x - me.vx;
        const double rvy = cv.y - me.vy;
        const double qa = rvx * rvx + rvy * rvy;
        const double d2 = dx * dx + dy * dy;
        const double qb = 2.0 * (rvx * dx + rvy * dy);
        const double qc = d2 - rr;
        const double bb = qb * qb;
        const double disc = bb - (4.0 * qa) * qc;
        // A finite time needs a > 0, disc >= 0 and -b + sqrt(disc) > 0.  For b >= 0, disc <= b*b gives
        // sqrt(disc) <= sqrt(fl(b*b)) = b (correctly rounded sqrt of a square is exact and monotone), so the
        // numerator is <= 0: INF.  disc == b*b is passed on although it cannot be finite either, so that one
        // compare also covers b*b = inf.  Everything else is decided by the literal routine on the list.
        return v && (qa > 0.0) && (disc >= 0.0) && ((qb < 0.0) || !(disc < bb));
      };
      while (__any_sync(FULL, (bits | nb1 | nb2) != 0u)) {
        uint32_t jA, tA, kA, jB, tB, kB;
        const bool vA = take(jA, tA, kA);
        const bool vB = take(jB, tB, kB);
        const bool hitA = probe(vA, jA, tA, kA);
        const bool hitB = probe(vB, jB, tB, kB);
        if (hitA) {  // order inside the list is irrelevant (min): a shared counter hands out the slots
          const uint32_t pos = atomicAdd(&w.hcnt, 1u);
          w.lj[pos] = jA;
          w.lo[pos] = (uint8_t)lane;
        }
        if (hitB) {
          const uint32_t pos = atomicAdd(&w.hcnt, 1u);
          w.lj[pos] = jB;
          w.lo[pos] = (uint8_t)lane;
        }
        if ((++it & 1u) == 0u) {  // at most 2 x 64 new entries since the last look at the counter
          __syncwarp();
          if (*(volatile uint32_t*)&w.hcnt > 2 * SW_CAP - 128) flush_hits();
        }
      }
      __syncwarp();
      if (*(volatile uint32_t*)&w.hcnt) flush_hits();
      __syncwarp();
      if (fast) t_i = __longlong_as_double((long long)w.tbits[lane]);
    }

    // ---------------- stage 3: force = sum over neighbours of compute_agent_force (zanlungo.rs:210-215)
    const bool fin = fast && (t_i != RCS_INF);
    if (__any_sync(FULL, fin)) {
      uint32_t a0 = 0, a1 = 0, a2 = 0, z0 = 0, z1 = 0, z2 = 0;  // list A bits (evaluate) / list B bits (prove zero)
      w.pfx[lane] = me.pfx;
      w.pfy[lane] = me.pfy;
      w.id[lane] = me.id;
      w.grp[lane] = grp;
      w.poison[lane] = 0u;
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
        a0 = m0 & y0; a1 = m1 & y1; a2 = m2 & y2;
        z0 = m0 & ~y0; z1 = m1 & ~y1; z2 = m2 & ~y2;
      }
      __syncwarp();
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
        const double2 c = pos[j], cv = vel[j];
        p.ox = c.x; p.oy = c.y; p.ovx = cv.x; p.ovy = cv.y; p.oid = ids[j];
        pair_force_literal(p, w.ti[o], a.groups[w.grp[o]], qx, qy);
      };
      // Every owner's pairs take one contiguous segment of a list, in canonical neighbour order (slice 0, 1, 2;
      // ascending slot), so the owner later adds its own slots front to back.  Segment offsets come from one warp
      // scan of the per-lane counts (list A in the low half-word, list B in the high one).
      const uint32_t cA = __popc(a0) + __popc(a1) + __popc(a2);
      const uint32_t cB = __popc(z0) + __popc(z1) + __popc(z2);
      const uint32_t own_cnt = cA | (cB << 16);
      uint32_t inc = own_cnt;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const uint32_t t = __shfl_up_sync(FULL, inc, d);
        if (lane >= (unsigned)d) inc += t;
      }
      const uint32_t exc = inc - own_cnt;
      uint32_t lane_begin = 0;
      while (lane_begin < 32) {  // one round unless the warp holds more pairs than a list does
        const uint32_t base = __shfl_sync(FULL, exc, lane_begin);
        const uint32_t rel = exc - base;  // field-wise: prefix sums are monotone, no borrow between the halves
        const uint32_t offA = rel & 0xffffu, offB = rel >> 16;
        const bool fits = lane >= lane_begin && offA + cA <= SW_CAP && offB + cB <= SW_CAP;
        const unsigned okm = __ballot_sync(FULL, fits || lane < lane_begin);
        const uint32_t lane_end = (okm == FULL) ? 32u : (uint32_t)(__ffs(~okm) - 1);  // > lane_begin: cA, cB <= 96
        const bool part = lane >= lane_begin && lane < lane_end;
        if (part) {
          // one walk per slice feeds both lists: yield pairs to A, weight-0 pairs to B
          uint32_t pA = offA, pB = SW_CAP + offB;
#define RCS_EMIT_SLICE(A, Z, S)                                          \
          for (uint32_t bits = (A) | (Z); bits; bits &= bits - 1u) {     \
            const uint32_t t = __ffs(bits) - 1;                          \
            const bool yv = (((A) >> t) & 1u) != 0u;                     \
            const uint32_t pos = yv ? pA : pB;                           \
            w.lj[pos] = (S) + t;                                         \
            w.lo[pos] = (uint8_t)lane;                                   \
            pA += yv ? 1u : 0u;                                          \
            pB += yv ? 0u : 1u;                                          \
          }
          RCS_EMIT_SLICE(a0, z0, s0)
          RCS_EMIT_SLICE(a1, z1, s1)
          RCS_EMIT_SLICE(a2, z2, s2)
#undef RCS_EMIT_SLICE
        }
        const uint32_t tot = __shfl_sync(FULL, inc, lane_end - 1) - base;
        const uint32_t nA = tot & 0xffffu, nB = tot >> 16;
        __syncwarp();
        for (uint32_t e = lane; e < nA; e += 32) {  // yield pairs: one per lane
          const uint32_t o = w.lo[e];
          const uint32_t j = w.lj[e];
          double qx, qy;
          const double2 c = pos[j], cv = vel[j];
          pair_force_yield(load_pre(o), w.px[o], w.py[o], w.vx[o], w.vy[o], c.x, c.y, cv.x, cv.y, w.ti[o],
                           a.groups[w.grp[o]], qx, qy);
          w.sfx[e] = qx;
          w.sfy[e] = qy;
        }
        for (uint32_t e = lane; e < nB; e += 32) {  // weight-0 pairs: prove the contribution is (+-0, +-0)
          const uint32_t o = w.lo[SW_CAP + e];
          const uint32_t j = w.lj[SW_CAP + e];
          const double2 c = pos[j], cv = vel[j];
          if (!pair_force_w0_is_zero(load_pre(o), c.x, c.y, cv.x, cv.y, w.ti[o])) {
            // contributes NaN or +-0 per component (rcs_math.cuh): NaN is order-independent
            double qx, qy;
            literal(o, j, qx, qy);
            const unsigned bits = (qx != qx ? 1u : 0u) | (qy != qy ? 2u : 0u);
            if (bits) atomicOr(&w.poison[o], bits);
          }
        }
        __syncwarp();
        if (part) {
          for (uint32_t r = 0; r < cA; ++r) {
            fx = fx + w.sfx[offA + r];
            fy = fy + w.sfy[offA + r];
          }
        }
        __syncwarp();
        lane_begin = lane_end;
      }
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
    integrate_and_store(a, i, me, g, grp, wp_in, role, velx, vely, t_i, fx, fy, nbc);
  }
  const bool own = active && role == ROLE_OWN;
  warp_stats(a, own ? cand : 0u, own ? nbc : 0u, (own && zan && t_i != RCS_INF) ? 1u : 0u);
}

}  // namespace rcs
