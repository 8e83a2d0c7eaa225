// rcs_host_dist.inl -- spatial strips over several GPUs.
// Part of rcs.cu (single translation unit).
//
// Rank r owns the cell columns [c0, c1) of the x-major LocationHash2D grid.  Every step it sends the agents
// of its outermost W = h + reach columns to each neighbour (h = ring width, reach = columns a radius query
// can touch beyond the agent's own column) and receives the neighbours' as ghosts.  Ghosts inside the ring
// (h columns beyond the boundary) are advanced redundantly -- same inputs, same canonical order, same
// kernels, hence bit-identical results on both ranks -- so an agent that crosses the boundary is simply
// kept by the rank that owns its new column and dropped by the other: migration needs no second message.
// Transports: NCCL point-to-point (one process per GPU), or peer copies between handles that live in one
// process (tests on a single GPU; no kernel ever waits on another kernel).

#include <dlfcn.h>

namespace rcs_host {

static NcclApi g_nccl;

static bool nccl_load(std::string& err) {
  if (g_nccl.lib) return true;
  void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!h) {
    err = std::string("cannot load NCCL: ") + dlerror();
    return false;
  }
  NcclApi a;
  a.lib = h;
  a.GetUniqueId = reinterpret_cast<decltype(a.GetUniqueId)>(dlsym(h, "ncclGetUniqueId"));
  a.CommInitRank = reinterpret_cast<decltype(a.CommInitRank)>(dlsym(h, "ncclCommInitRank"));
  a.CommDestroy = reinterpret_cast<decltype(a.CommDestroy)>(dlsym(h, "ncclCommDestroy"));
  a.Send = reinterpret_cast<decltype(a.Send)>(dlsym(h, "ncclSend"));
  a.Recv = reinterpret_cast<decltype(a.Recv)>(dlsym(h, "ncclRecv"));
  a.AllReduce = reinterpret_cast<decltype(a.AllReduce)>(dlsym(h, "ncclAllReduce"));
  a.GroupStart = reinterpret_cast<decltype(a.GroupStart)>(dlsym(h, "ncclGroupStart"));
  a.GroupEnd = reinterpret_cast<decltype(a.GroupEnd)>(dlsym(h, "ncclGroupEnd"));
  a.GetErrorString = reinterpret_cast<decltype(a.GetErrorString)>(dlsym(h, "ncclGetErrorString"));
  if (!a.GetUniqueId || !a.CommInitRank || !a.CommDestroy || !a.Send || !a.Recv || !a.AllReduce || !a.GroupStart ||
      !a.GroupEnd) {
    err = "libnccl.so.2 lacks a required entry point";
    return false;
  }
  g_nccl = a;
  return true;
}

#define NCCL_TRY(sim, call)                                                                            \
  do {                                                                                                 \
    int r__ = (call);                                                                                  \
    if (r__ != 0) {                                                                                    \
      (sim)->err = std::string("NCCL error: ") +                                                       \
                   (g_nccl.GetErrorString ? g_nccl.GetErrorString(r__) : "?") + " at " #call;         \
      return RCS_ERR_NCCL;                                                                             \
    }                                                                                                  \
  } while (0)

static int halo_alloc(rcs_sim* s, HaloMem& m, uint32_t cap) {
  const uint64_t bytes = 64 + 8ull * 8ull * cap;
  CU_TRY(s, cudaMalloc(&m.base, bytes));
  CU_TRY(s, cudaMemset(m.base, 0, 64));
  m.bytes = bytes;
  char* p = static_cast<char*>(m.base);
  HaloBuf& b = m.buf;
  b.count = reinterpret_cast<uint32_t*>(p);
  double* arr = reinterpret_cast<double*>(p + 64);
  b.pos = reinterpret_cast<double2*>(arr);
  b.vel = reinterpret_cast<double2*>(arr + 2ull * cap);
  b.id = reinterpret_cast<unsigned long long*>(arr + 4ull * cap);
  b.meta = reinterpret_cast<unsigned long long*>(arr + 5ull * cap);
  b.pv = reinterpret_cast<double2*>(arr + 6ull * cap);
  b.cap = cap;
  return RCS_OK;
}

static void halo_free(HaloMem& m) {
  cudaFree(m.base);
  m = HaloMem{};
}

// columns of the grid that can hold an agent: x_idx in [0, x_max]
static uint64_t strip_columns(const rcs_sim* s) { return s->grid.x_max < 0 ? 0 : (uint64_t)s->grid.x_max + 1; }

static void strip_range(const rcs_sim* s, int rank, int world, uint64_t& c0, uint64_t& c1) {
  if ((int)s->strip_bounds.size() == world + 1) {  // caller-supplied boundaries (balanced by agent count)
    c0 = s->strip_bounds[rank];
    c1 = s->strip_bounds[rank + 1];
    return;
  }
  const uint64_t cols = strip_columns(s);
  c0 = cols * (uint64_t)rank / (uint64_t)world;
  c1 = cols * (uint64_t)(rank + 1) / (uint64_t)world;
}

static int strip_setup(rcs_sim* s, int rank, int world, uint64_t halo_capacity) {
  if (world <= 0 || rank < 0 || rank >= world) {
    s->err = "bad rank / world";
    return RCS_ERR_ARG;
  }
  if (s->strip.enabled) {
    s->err = "strips are already initialised on this handle";
    return RCS_ERR_ARG;
  }
  if (s->n != 0 || s->ever_had_sources) {
    s->err = "rcs_dist_init must come before agents and source sinks are added";
    return RCS_ERR_ARG;
  }
  const uint64_t ny = s->grid.nx ? s->grid.len / s->grid.nx : 0;
  if (s->grid.nx == 0 || ny > s->grid.nx) {
    // with n_y > n_x the reference's width-stride index aliases columns (location_hash_2d.rs:59)
    s->err = "strips need a grid with n_y <= n_x";
    return RCS_ERR_ARG;
  }
  uint64_t c0, c1, t0, t1;
  strip_range(s, rank, world, c0, c1);
  if (c1 <= c0) {
    s->err = "more ranks than cell columns";
    return RCS_ERR_ARG;
  }
  StripDev st{};
  st.enabled = 1;
  st.c0 = (uint32_t)c0;
  st.c1 = (uint32_t)c1;
  st.h = 1;
  strip_range(s, std::max(rank - 1, 0), world, t0, t1);
  st.lc0 = rank > 0 ? (uint32_t)t0 : (uint32_t)c0;
  strip_range(s, std::min(rank + 1, world - 1), world, t0, t1);
  st.rc1 = rank + 1 < world ? (uint32_t)t1 : (uint32_t)c1;
  uint64_t hc = halo_capacity ? halo_capacity : std::max<uint64_t>(s->cap / 8, 1024);
  hc = std::min<uint64_t>(hc, s->cap);
  for (HaloMem* m : {&s->send_l, &s->send_r, &s->recv_l, &s->recv_r}) {
    int rc = halo_alloc(s, *m, (uint32_t)hc);
    if (rc) return rc;
  }
  CU_TRY(s, dalloc(&s->srt_cell, s->cap + 16));
  // (ghosts carry the host-planner preferred velocity only when such a planner exists; the halo buffers always
  // have room for it)
  s->strip = st;
  s->rank = rank;
  s->world = world;
  s->n_ub = (uint32_t)s->cap;
  return RCS_OK;
}

// W = h + reach, from the groups registered so far (reach: see the header comment of this file)
static int strip_halo_width(rcs_sim* s) {
  uint64_t reach = 1;
  for (const GroupDev& g : s->groups) {
    if (g.lp_kind != LP_ZANLUNGO) continue;
    double r = g.eyesight / s->grid.res;
    if (!(r >= 0.0) || r > 1e6) {
      s->err = "eyesight range too large for strips";
      return RCS_ERR_HALO;
    }
    reach = std::max<uint64_t>(reach, (uint64_t)std::floor(r) + 1);
  }
  const uint64_t w = s->strip.h + reach;
  if (s->world > 1 && w > (uint64_t)(s->strip.c1 - s->strip.c0)) {
    s->err = "strip narrower than the halo (fewer ranks or a larger cell size needed)";
    return RCS_ERR_HALO;
  }
  s->halo_width = (uint32_t)w;
  s->strip.reach = (uint32_t)reach;
  // index only the columns [c0 - W, c1 + W] : every query of an owned or ring agent stays inside [c0 - W, c1 + W);
  // the extra column keeps the aliased reads of top-row queries (which the step kernel refuses) defined
  const uint64_t col_lo = s->strip.c0 > w ? s->strip.c0 - w : 0;
  const uint64_t col_hi = std::min<uint64_t>((uint64_t)s->strip.c1 + w + 1, strip_columns(s));
  s->cell_lo = (col_lo * s->grid.nx) & ~3ull;
  s->cell_hi = std::min<uint64_t>(col_hi * s->grid.nx, s->grid.len);
  return RCS_OK;
}

// Strips with source sinks: the ranks' spawn bitmaps are disjoint, their sum is the global spawn set of the step.
static int step_spawn_set_nccl(rcs_sim* s) {
  if (s->world == 1) {
    CU_TRY(s, cudaMemcpyAsync(s->d_ss_bits, s->d_ss_bits_local, s->ss_words * sizeof(uint32_t),
                              cudaMemcpyDeviceToDevice, s->stream));
    return RCS_OK;
  }
  if (!s->nccl_comm) {
    s->err = "strip handle has no communicator";
    return RCS_ERR_NCCL;
  }
  NCCL_TRY(s, g_nccl.AllReduce(s->d_ss_bits_local, s->d_ss_bits, s->ss_words, /*ncclUint32*/ 3, /*ncclSum*/ 0,
                               s->nccl_comm, s->stream));
  return RCS_OK;
}

static int step_exchange_nccl(rcs_sim* s) {
  if (s->world == 1 || s->peer.enabled) return RCS_OK;  // peer stores: halo_unpack_kernel publishes and waits
  if (!s->nccl_comm) {
    s->err = "strip handle has no communicator";
    return RCS_ERR_NCCL;
  }
  const int has_l = s->rank > 0, has_r = s->rank + 1 < s->world;
  NCCL_TRY(s, g_nccl.GroupStart());
  if (has_l) {
    NCCL_TRY(s, g_nccl.Send(s->send_l.base, s->send_l.bytes, /*ncclUint8*/ 1, s->rank - 1, s->nccl_comm, s->stream));
    NCCL_TRY(s, g_nccl.Recv(s->recv_l.base, s->recv_l.bytes, 1, s->rank - 1, s->nccl_comm, s->stream));
  }
  if (has_r) {
    NCCL_TRY(s, g_nccl.Send(s->send_r.base, s->send_r.bytes, 1, s->rank + 1, s->nccl_comm, s->stream));
    NCCL_TRY(s, g_nccl.Recv(s->recv_r.base, s->recv_r.bytes, 1, s->rank + 1, s->nccl_comm, s->stream));
  }
  NCCL_TRY(s, g_nccl.GroupEnd());
  return RCS_OK;
}

// ---- peer-store transport ----------------------------------------------------------------------------------------
constexpr uint32_t PEER_MAGIC = 0x52435348u;  // "RCSH"
constexpr uint64_t PEER_ROW_BYTES = 64;       // pos 16 + vel 16 + id 8 + meta 8 + pv 16

static uint64_t peer_arena_bytes(uint32_t cap) { return 192 + 2ull * 2ull * cap * PEER_ROW_BYTES; }

// the receive buffer of one side (0: from the left neighbour, 1: from the right one) inside an arena
static HaloBuf peer_side(void* arena, uint32_t cap, int side, uint32_t** hdr) {
  char* p = static_cast<char*>(arena);
  *hdr = reinterpret_cast<uint32_t*>(p + 64 + 64 * side);
  char* rows = p + 192 + (uint64_t)side * 2ull * cap * PEER_ROW_BYTES;
  HaloBuf b{};
  b.count = *hdr;
  b.cap = cap;
  const uint64_t n = 2ull * cap;  // two halves
  b.pos = reinterpret_cast<double2*>(rows);
  b.vel = reinterpret_cast<double2*>(rows + 16 * n);
  b.id = reinterpret_cast<unsigned long long*>(rows + 32 * n);
  b.meta = reinterpret_cast<unsigned long long*>(rows + 40 * n);
  b.pv = reinterpret_cast<double2*>(rows + 48 * n);
  return b;
}

static void peer_teardown(rcs_sim* s) {
  for (void*& m : s->peer.nb_arena) {
    if (m) cudaIpcCloseMemHandle(m);
    m = nullptr;
  }
  cudaFree(s->peer.arena);
  cudaFree(s->peer.xseq);
  s->peer = PeerHalo{};
}

}  // namespace rcs_host

static void dist_teardown(rcs_sim* s) {
  using namespace rcs_host;
  peer_teardown(s);
  if (s->nccl_comm && g_nccl.CommDestroy) g_nccl.CommDestroy(s->nccl_comm);
  s->nccl_comm = nullptr;
  halo_free(s->send_l);
  halo_free(s->send_r);
  halo_free(s->recv_l);
  halo_free(s->recv_r);
  if (s->ev_packed) cudaEventDestroy(s->ev_packed);
  if (s->ev_copied) cudaEventDestroy(s->ev_copied);
  if (s->ev_flags) cudaEventDestroy(s->ev_flags);
  s->ev_packed = s->ev_copied = s->ev_flags = nullptr;
  // unlink from a single-process group: the other handles must not touch this one any more
  for (rcs_sim* o : s->local_group)
    if (o && o != s)
      for (rcs_sim*& q : o->local_group)
        if (q == s) q = nullptr;
  s->local_group.clear();
}

extern "C" {

int rcs_nccl_unique_id(uint8_t out_id[128]) {
  if (!out_id) return RCS_ERR_ARG;
  if (!nccl_load(g_create_error)) return RCS_ERR_NCCL;
  Id128 id;
  int r = g_nccl.GetUniqueId(&id);
  if (r != 0) {
    g_create_error = std::string("NCCL error: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?");
    return RCS_ERR_NCCL;
  }
  std::memcpy(out_id, id.internal, 128);
  return RCS_OK;
}

int rcs_dist_init(rcs_sim* s, int32_t rank, int32_t world, const uint8_t nccl_id[128], uint64_t halo_capacity) {
  if (!s || (world > 1 && !nccl_id)) return RCS_ERR_ARG;
  CU_TRY(s, cudaSetDevice(s->device));
  int rc = do_sync(s);
  if (rc) return rc;
  if (world > 1) {
    if (!nccl_load(s->err)) return RCS_ERR_NCCL;
    Id128 id;
    std::memcpy(id.internal, nccl_id, 128);
    NCCL_TRY(s, g_nccl.CommInitRank(&s->nccl_comm, world, id, rank));
  }
  return strip_setup(s, rank, world, halo_capacity);
}

int rcs_dist_init_local(rcs_sim** sims, int32_t world, uint64_t halo_capacity) {
  if (!sims || world <= 0) return RCS_ERR_ARG;
  for (int r = 0; r < world; ++r)
    if (!sims[r]) return RCS_ERR_ARG;
  for (int r = 0; r < world; ++r) {
    rcs_sim* s = sims[r];
    CU_TRY(s, cudaSetDevice(s->device));
    int rc = do_sync(s);
    if (rc) return rc;
    rc = strip_setup(s, r, world, halo_capacity);
    if (rc) return rc;
    CU_TRY(s, cudaEventCreateWithFlags(&s->ev_packed, cudaEventDisableTiming));
    CU_TRY(s, cudaEventCreateWithFlags(&s->ev_copied, cudaEventDisableTiming));
    CU_TRY(s, cudaEventCreateWithFlags(&s->ev_flags, cudaEventDisableTiming));
    s->local_group.assign(sims, sims + world);
  }
  return RCS_OK;
}

int rcs_dist_step_local(rcs_sim** sims, int32_t world, uint64_t secs, uint32_t nanos, uint32_t flags) {
  if (!sims || world <= 0) return RCS_ERR_ARG;
  const double dt = (double)secs + (double)nanos / 1000000000.0;
  for (int r = 0; r < world; ++r) {
    rcs_sim* s = sims[r];
    if (!s || (int)s->local_group.size() != world || s->local_group[r] != s) return RCS_ERR_ARG;
  }
  bool any_sources = false;
  for (int r = 0; r < world; ++r) {
    rcs_sim* s = sims[r];
    CU_TRY(s, cudaSetDevice(s->device));
    int rc = step_phase_a1(s, dt);
    if (rc) return rc;
    any_sources = any_sources || s->n_sources_alive != 0;
    if (s->n_sources_alive) CU_TRY(s, cudaEventRecord(s->ev_flags, s->stream));
  }
  for (int r = 0; r < world; ++r) {
    rcs_sim* s = sims[r];
    CU_TRY(s, cudaSetDevice(s->device));
    if (any_sources) {
      // the spawn set of the step: every rank's bitmap, summed (what ncclAllReduce does between processes)
      if (!s->n_sources_alive || s->ss_words != sims[0]->ss_words) {
        s->err = "every rank of a strip group must hold the same source sinks";
        return RCS_ERR_ARG;
      }
      for (int q = 0; q < world; ++q) {
        if (q != r) CU_TRY(s, cudaStreamWaitEvent(s->stream, sims[q]->ev_flags, 0));
        CU_TRY(s, cudaMemcpyAsync(s->d_ss_bits_parts + (size_t)q * s->ss_words, sims[q]->d_ss_bits_local,
                                  s->ss_words * sizeof(uint32_t), cudaMemcpyDefault, s->stream));
      }
      ss_bits_sum_kernel<<<blocks_for(s->ss_words, 256), 256, 0, s->stream>>>(s->ss_words, (uint32_t)world,
                                                                             s->d_ss_bits_parts, s->d_ss_bits);
      s->launches += 1;
    }
    int rc = step_phase_a2(s, dt);
    if (rc) return rc;
  }
  for (int r = 0; r < world; ++r) {
    rcs_sim* s = sims[r];
    CU_TRY(s, cudaSetDevice(s->device));
    int rc = step_exchange_local(s);
    if (rc) return rc;
    rc = step_phase_b(s, dt, flags);
    if (rc) return rc;
  }
  return RCS_OK;
}

int rcs_dist_peer_export(rcs_sim* s, uint8_t out_handle[64]) {
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  if (!s || !out_handle) return RCS_ERR_ARG;
  CU_TRY(s, cudaSetDevice(s->device));
  if (!s->strip.enabled || s->world < 2 || !s->local_group.empty()) {
    s->err = "the peer-store transport is for strip handles of a multi-process job (one GPU per rank)";
    return RCS_ERR_ARG;
  }
  int rc = do_sync(s);
  if (rc) return rc;
  if (!s->peer.arena) {
    const uint32_t cap = s->recv_l.buf.cap;
    const uint64_t bytes = peer_arena_bytes(cap);
    CU_TRY(s, cudaMalloc(&s->peer.arena, bytes));
    CU_TRY(s, cudaMemset(s->peer.arena, 0, bytes));
    const uint32_t desc[2] = {PEER_MAGIC, cap};
    CU_TRY(s, cudaMemcpy(s->peer.arena, desc, sizeof(desc), cudaMemcpyHostToDevice));
    CU_TRY(s, dalloc(&s->peer.xseq, 1));
    CU_TRY(s, cudaMemset(s->peer.xseq, 0, sizeof(uint32_t)));
    CU_TRY(s, cudaDeviceSynchronize());  // zeroed before any neighbour can see it
    s->peer.arena_bytes = bytes;
    s->peer.cap = cap;
    uint32_t* hdr = nullptr;
    s->peer.local[0] = peer_side(s->peer.arena, cap, 0, &hdr);
    s->peer.local[1] = peer_side(s->peer.arena, cap, 1, &hdr);
  }
  cudaIpcMemHandle_t h;
  CU_TRY(s, cudaIpcGetMemHandle(&h, s->peer.arena));
  std::memcpy(out_handle, &h, 64);
  return RCS_OK;
}

int rcs_dist_peer_connect(rcs_sim* s, const uint8_t* left_handle, const uint8_t* right_handle) {
  if (!s) return RCS_ERR_ARG;
  CU_TRY(s, cudaSetDevice(s->device));
  if (!s->peer.arena || s->peer.enabled) {
    s->err = "rcs_dist_peer_connect follows rcs_dist_peer_export, once";
    return RCS_ERR_ARG;
  }
  const bool has[2] = {s->rank > 0, s->rank + 1 < s->world};
  const uint8_t* hs[2] = {left_handle, right_handle};
  if ((has[0] && !left_handle) || (has[1] && !right_handle)) {
    s->err = "a neighbour's handle is missing";
    return RCS_ERR_ARG;
  }
  int rc = do_sync(s);
  if (rc) return rc;
  for (int side = 0; side < 2; ++side) {
    if (!has[side]) continue;
    cudaIpcMemHandle_t h;
    std::memcpy(&h, hs[side], 64);
    cudaError_t e = cudaIpcOpenMemHandle(&s->peer.nb_arena[side], h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
      s->err = std::string("CUDA error: ") + cudaGetErrorString(e) + " at cudaIpcOpenMemHandle (peer-store transport)";
      peer_teardown(s);
      return RCS_ERR_CUDA;
    }
    uint32_t desc[2] = {0, 0};
    e = cudaMemcpy(desc, s->peer.nb_arena[side], sizeof(desc), cudaMemcpyDeviceToHost);
    if (e != cudaSuccess || desc[0] != PEER_MAGIC || desc[1] == 0) {
      s->err = "the neighbour's receive arena cannot be read (no peer access between the two GPUs?)";
      peer_teardown(s);
      return RCS_ERR_CUDA;
    }
    // this rank's LEFT boundary columns are what the left neighbour receives from its RIGHT side, and vice versa
    s->peer.remote[side] = peer_side(s->peer.nb_arena[side], desc[1], 1 - side, &s->peer.remote_hdr[side]);
  }
  s->peer.enabled = true;
  s->graph_epoch += 1;
  return RCS_OK;
}

int rcs_dist_peer_disable(rcs_sim* s) {
  if (!s) return RCS_ERR_ARG;
  CU_TRY(s, cudaSetDevice(s->device));
  int rc = do_sync(s);
  if (rc) return rc;
  peer_teardown(s);  // back to ncclSend / ncclRecv
  s->graph_epoch += 1;
  return RCS_OK;
}

int rcs_dist_set_boundaries(rcs_sim* s, int32_t world, const uint64_t* bounds) {
  if (!s || world <= 0) return RCS_ERR_ARG;
  if (s->strip.enabled) {
    s->err = "strip boundaries must be set before rcs_dist_init";
    return RCS_ERR_ARG;
  }
  if (!bounds) {
    s->strip_bounds.clear();
    return RCS_OK;
  }
  const uint64_t cols = strip_columns(s);
  bool ok = bounds[0] == 0 && bounds[world] == cols;
  for (int r = 0; r < world && ok; ++r) ok = bounds[r] < bounds[r + 1];
  if (!ok) {
    s->err = "strip boundaries must rise strictly from 0 to the number of cell columns";
    return RCS_ERR_ARG;
  }
  s->strip_bounds.assign(bounds, bounds + world + 1);
  return RCS_OK;
}

int rcs_dist_strip(rcs_sim* s, int32_t rank, int32_t world, uint64_t* c0, uint64_t* c1) {
  if (!s || !c0 || !c1 || world <= 0 || rank < 0 || rank >= world) return RCS_ERR_ARG;
  strip_range(s, rank, world, *c0, *c1);
  return RCS_OK;
}

}  // extern "C"
