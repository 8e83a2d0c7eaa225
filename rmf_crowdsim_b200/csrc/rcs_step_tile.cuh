// rcs_step_tile.cuh -- the hot kernel with the query stencil staged in shared memory (sm_100a).
//
// The reference scans, for every agent, the cells `for x in left..=right { for y in bottom..=top }`
// (location_hash_2d.rs:245-256) and hands the survivors of the strict radius test to the planner twice
// (zanlungo.rs:76-91 and :210-215).  With the agents in canonical (cell, id) order each stencil column is one
// contiguous slice of the sorted arrays, and the slices of CONSECUTIVE agents overlap almost completely: a block of
// 128 consecutive agents (a run of ~32 cells of one cell column at 4 agents per cell) touches three runs of ~34 cells.
// step_warp_kernel reads those rows through L1 / L2, agent by agent; ncu attributes 36 % of its stall cycles to
// long-scoreboard waits on them (profiles/r01d_final_build.md).  Here the block first copies the union of its
// agents' slices -- three contiguous row ranges of (position, velocity, id), one per stencil column -- into shared
// memory, once, and the three stages of rcs_step_warp.cuh (filter, t_i, force) then run out of shared memory.
//
//   staging   cooperative 16-byte loads (default), or one-dimensional bulk copies issued by one thread and
//             completed on an mbarrier (cp.async.bulk, -DRCS_TILE_TMA=1): both forms move the same bytes.
//   agents    whose slices cannot be staged -- stencil wider than three columns, more than 32 candidates in a column,
//             ids >= 2^53, or a block whose union exceeds the staging capacity -- go to step_aside_kernel's lists,
//             exactly as in step_warp_kernel.
// Same arithmetic in the same canonical order as every other form of the kernel: tests compare them bit for bit.
#pragma once

#include "rcs_step_warp.cuh"

namespace rcs {

#ifndef RCS_TILE_ROWS
#define RCS_TILE_ROWS 576
#endif
#ifndef RCS_TILE_BLOCKS
#define RCS_TILE_BLOCKS 4
#endif
#ifndef RCS_TILE_TMA
#define RCS_TILE_TMA 0
#endif

constexpr uint32_t ST_ROWS = RCS_TILE_ROWS;  // staged rows per block (40 B each)

struct TileShared {
  double2 pos[ST_ROWS + 2];
  double2 vel[ST_ROWS + 2];
  unsigned long long id[ST_ROWS + 4];
  WarpShared w[SW_WARPS];
  uint32_t red[SW_WARPS][6];
  uint32_t st_cand[SW_WARPS], st_nbc[SW_WARPS], st_fin[SW_WARPS];
  unsigned long long mbar;
};

// ---------------- stage 1 out of shared memory: one mask per slice; q = slice start as a staged row index.
// Lanes that have run out of candidates re-read row 0 with a threshold that rejects everything (the mask bit is
// masked by t < l); the agent's own bit is cleared by the caller.
__device__ __forceinline__ void st_radius_masks(const double2* spos, double mpx, double mpy, double thr2, uint32_t q0,
                                                uint32_t l0, uint32_t q1, uint32_t l1, uint32_t q2, uint32_t l2,
                                                uint32_t& m0, uint32_t& m1, uint32_t& m2) {
  const unsigned FULL = 0xffffffffu;
  const uint32_t x0 = __reduce_max_sync(FULL, l0), x1 = __reduce_max_sync(FULL, l1), x2 = __reduce_max_sync(FULL, l2);
#define RCS_TILE_FILTER(MX, Q, L, M)                                   \
  _Pragma("unroll 4")                                                  \
  for (uint32_t t = 0; t < (MX); ++t) {                                \
    const bool in = t < (L);                                           \
    const double2 c = spos[in ? (Q) + t : 0u];                         \
    const double dx = c.x - mpx;                                       \
    const double dy = c.y - mpy;                                       \
    const double d2 = dx * dx + dy * dy;                               \
    (M) |= ((d2 < thr2) && in) ? (1u << t) : 0u;                       \
  }
  RCS_TILE_FILTER(x0, q0, l0, m0)
  RCS_TILE_FILTER(x1, q1, l1, m1)
  RCS_TILE_FILTER(x2, q2, l2, m2)
#undef RCS_TILE_FILTER
}

#if RCS_TILE_TMA
__device__ __forceinline__ uint32_t st_smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
// one-dimensional bulk copy global -> shared, completion counted in bytes on an mbarrier
__device__ __forceinline__ void st_bulk_copy(void* dst, const void* src, uint32_t bytes, unsigned long long* mbar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   st_smem_addr(dst)),
               "l"(src), "r"(bytes), "r"(st_smem_addr(mbar))
               : "memory");
}
#endif

__global__ void __launch_bounds__(32 * SW_WARPS, RCS_TILE_BLOCKS) step_tile_kernel(StepArgs a) {
  extern __shared__ __align__(16) unsigned char st_smem_raw[];
  TileShared& sh = *reinterpret_cast<TileShared*>(st_smem_raw);
  const unsigned warp = threadIdx.x >> 5;
  WarpShared& w = sh.w[warp];
  const unsigned lane = threadIdx.x & 31u;
  const unsigned FULL = 0xffffffffu;
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;

  const double2* __restrict__ pos = a.in.pos;
  const double2* __restrict__ vel = a.in.vel;
  const uint64_t* __restrict__ ids = a.in.id;

  // own rows first (every array holds at least a.n entries): their latency overlaps the checks below
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
#if RCS_TILE_TMA
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(st_smem_addr(&sh.mbar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
#endif
  if (a.status->failed) return;  // block-uniform
  const uint32_t n_live = *a.n_sorted;
  bool active = inb && (i < n_live);

  double velx = 0.0, vely = 0.0, thr2 = 0.0, rr = 0.0;
  uint32_t role = ROLE_PASSIVE;
  bool zan = false;
  uint32_t s0 = 0, s1 = 0, s2 = 0, l0 = 0, l1 = 0, l2 = 0;  // candidate slices of this lane (sorted slots)
  bool fast = false;
  uint32_t cand = 0, nbc = 0;
  uint32_t aside = 0;
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
  bool coop = false;
  if (active) {
    const GroupDev& g = a.groups[grp];
    high_level_velocity(a, i, g, me, velx, vely);
    zan = g.lp_kind == LP_ZANLUNGO;
    thr2 = g.thr2;
    rr = g.rr;
    if (zan) {
      s0 = sl.x; s1 = sl.y; s2 = sl.z;
      l0 = sl.w & 0xffu; l1 = (sl.w >> 8) & 0xffu; l2 = (sl.w >> 16) & 0xffu;
      coop = g.w0_fast && (me.id >> 53) == 0ull;
      fast = coop && (sl.w >> 24) != 0u && l0 <= SW_SLICE_MAX && l1 <= SW_SLICE_MAX && l2 <= SW_SLICE_MAX;
    }
  }

  // ---- union of the block's slices, per stencil column: three contiguous row ranges of the sorted arrays
  {
    const uint32_t NONE = 0xffffffffu;
    const uint32_t a0 = __reduce_min_sync(FULL, (fast && l0) ? s0 : NONE);
    const uint32_t a1 = __reduce_min_sync(FULL, (fast && l1) ? s1 : NONE);
    const uint32_t a2 = __reduce_min_sync(FULL, (fast && l2) ? s2 : NONE);
    const uint32_t b0 = __reduce_max_sync(FULL, (fast && l0) ? s0 + l0 : 0u);
    const uint32_t b1 = __reduce_max_sync(FULL, (fast && l1) ? s1 + l1 : 0u);
    const uint32_t b2 = __reduce_max_sync(FULL, (fast && l2) ? s2 + l2 : 0u);
    if (lane == 0) {
      sh.red[warp][0] = a0; sh.red[warp][1] = a1; sh.red[warp][2] = a2;
      sh.red[warp][3] = b0; sh.red[warp][4] = b1; sh.red[warp][5] = b2;
    }
  }
  __syncthreads();
  uint32_t lo[3], len[3];
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    uint32_t mn = 0xffffffffu, mx = 0u;
#pragma unroll
    for (int ww = 0; ww < SW_WARPS; ++ww) {
      mn = min(mn, sh.red[ww][d]);
      mx = max(mx, sh.red[ww][3 + d]);
    }
    lo[d] = mx > mn ? mn : 0u;
    len[d] = mx > mn ? mx - mn : 0u;
  }
  // row r of range d sits at staged index off[d] + r.  Bulk copies need 16-byte aligned ends on both sides: the id
  // rows are 8 bytes, so a range starts at an even row and its staged offset is even as well
#if RCS_TILE_TMA
  uint32_t lo_al[3], len_al[3], off[3];
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    lo_al[d] = lo[d] & ~1u;
    len_al[d] = len[d] ? ((lo[d] + len[d] + 1u) & ~1u) - lo_al[d] : 0u;
  }
  off[0] = 0u;
  off[1] = len_al[0];
  off[2] = len_al[0] + len_al[1];
  const uint32_t total = off[2] + len_al[2];
#else
  uint32_t off[3];
  off[0] = 0u;
  off[1] = len[0];
  off[2] = len[0] + len[1];
  const uint32_t total = off[2] + len[2];
#endif
  const bool staged = total <= ST_ROWS;  // block-uniform
  if (!staged) fast = false;
  if (zan && !fast) {
    l0 = l1 = l2 = 0;
    active = false;
    zan = false;
    aside = coop ? 1u : 2u;
  }
  if (fast) cand = l0 + l1 + l2;
  // agents for step_aside_kernel keep their neighbours in this warp as neighbours on the list
  {
    const unsigned wm = __ballot_sync(FULL, aside == 1u);
    if (wm) {
      uint32_t at = 0;
      if (lane == (unsigned)(__ffs(wm) - 1)) at = atomicAdd(&a.status->wide_count, (unsigned)__popc(wm));
      at = __shfl_sync(FULL, at, __ffs(wm) - 1);
      if (aside == 1u) a.wide_list[at + __popc(wm & ((1u << lane) - 1u))] = i;
    }
    if (aside == 2u) a.slow_list[atomicAdd(&a.status->slow_count, 1u)] = i;
  }

  // ---- staging
  if (staged && total) {
#if RCS_TILE_TMA
    if (threadIdx.x == 0) {
      const uint32_t bytes = total * 40u;
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(st_smem_addr(&sh.mbar)), "r"(bytes)
                   : "memory");
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        if (!len_al[d]) continue;
        st_bulk_copy(&sh.pos[off[d]], pos + lo_al[d], len_al[d] * 16u, &sh.mbar);
        st_bulk_copy(&sh.vel[off[d]], vel + lo_al[d], len_al[d] * 16u, &sh.mbar);
        st_bulk_copy(&sh.id[off[d]], ids + lo_al[d], len_al[d] * 8u, &sh.mbar);
      }
    }
    // wait for phase 0 of the barrier (all bytes have landed)
    {
      uint32_t done = 0;
      while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(st_smem_addr(&sh.mbar)), "r"(0u)
            : "memory");
      }
    }
#else
    const uint32_t dl1 = lo[1] - off[1], dl2 = lo[2] - off[2];  // sorted slot = staged row + delta (mod 2^32)
    for (uint32_t r = threadIdx.x; r < total; r += blockDim.x) {
      const uint32_t j = r + (r < off[1] ? lo[0] : (r < off[2] ? dl1 : dl2));
      sh.pos[r] = pos[j];
      sh.vel[r] = vel[j];
      sh.id[r] = ids[j];
    }
    __syncthreads();
#endif
  }
#if RCS_TILE_TMA
  const uint32_t q0 = fast ? s0 - lo_al[0] + off[0] : 0u, q1 = fast ? s1 - lo_al[1] + off[1] : 0u,
                 q2 = fast ? s2 - lo_al[2] + off[2] : 0u;
#else
  const uint32_t q0 = fast ? s0 - lo[0] + off[0] : 0u, q1 = fast ? s1 - lo[1] + off[1] : 0u,
                 q2 = fast ? s2 - lo[2] + off[2] : 0u;
#endif
  const double2* spos = sh.pos;
  const double2* svel = sh.vel;
  const uint64_t* sid = reinterpret_cast<const uint64_t*>(sh.id);

  uint32_t m0 = 0, m1 = 0, m2 = 0;
  if (staged && total) {
    st_radius_masks(spos, me.px, me.py, thr2, q0, l0, q1, l1, q2, l2, m0, m1, m2);
    // the self filter (lib.rs:284): the agent's own row is candidate i - s of the slice that holds its cell
    if (fast) {
      const uint32_t d0 = i - s0, d1 = i - s1, d2 = i - s2;
      if (d0 < l0) m0 &= ~(1u << d0);
      if (d1 < l1) m1 &= ~(1u << d1);
      if (d2 < l2) m2 &= ~(1u << d2);
    }
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
    uint32_t y0 = 0, y1 = 0, y2 = 0;
    sw_collision_times(w, lane, me, rr, 0u, spos, svel, sid, m0, m1, m2, q0, q1, q2, y0, y1, y2);
    if (fast) t_i = __longlong_as_double((long long)w.tbits[lane]);

    const bool fin = fast && (t_i != RCS_INF);
    if (__any_sync(FULL, fin)) {
      uint32_t a0 = 0, a1 = 0, a2 = 0, z0 = 0, z1 = 0, z2 = 0;
      sw_store_owner(w, lane, me, grp);
      if (fin) {
        sw_store_owner_pre(w, lane, me, t_i, a.groups[grp]);
        a0 = m0 & y0; a1 = m1 & y1; a2 = m2 & y2;
        z0 = m0 & ~y0; z1 = m1 & ~y1; z2 = m2 & ~y2;
      }
      __syncwarp();
      sw_pair_forces(a, w, lane, spos, svel, sid, a0, a1, a2, z0, z1, z2, q0, q1, q2, fx, fy);
      sw_apply_poison(w, lane, fx, fy);
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
