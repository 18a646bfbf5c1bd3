// mapf_capi.cu -- the C ABI of libmapf_b200.so (include/mapf_b200.h): handle management, argument
// validation, tile sizing and kernel launches.  No torch types, no exceptions across the boundary.
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <new>

#include "mapf_internal.h"

#include <chrono>
#define MAPF_HOST_CHUNKS 16

struct mapf_handle {
  mapf_cfg cfg;
  MapfDims d;
  MapfTileLayout L;       // full layout
  MapfTileLayout L_lite;  // PRIMAL without mid-sweep outputs: no second occupancy grid (more resident tiles per SM)
  MapfState S;
  int device;
  int fov_fast;        // a specialised tile kernel exists for cfg.fov
  int64_t* partial_state_out;   // mapf_partial_bind_state_out: get_state() written by every observation launch
  int debug_corrupt;            // mapf_debug_corrupt_canary: consumed by the next tile / pipe launch
  int64_t launches;
  int mag_lut_len;
  // mapf_lifelong_bind: goal queues popped by the step kernel itself
  const int16_t* life_queue;
  int32_t* life_head;
  int life_Q;
  int32_t* life_lists;   // device [2][E*N]: re-assigned (env, agent) pairs of the last two step launches
  int32_t* life_cnts;    // device [2][2]: list length, overflow counter of the BFS
  int life_slot;         // the list the next step launch appends to
  int life_last;         // the list of the most recent step launch (-1: none yet)
  // device staging for the *_host entry points (allocated on first use)
  uint8_t* hs_actions;
  double* hs_reward;
  uint8_t* hs_terminated;
  uint8_t* hs_dones;
  uint8_t* hs_avail;
  void* hs_obs;
  size_t hs_obs_bytes;
  double* hs_vec;
  // bit-packed PCIe transport of the host entry point (mapf_host_unpack.cpp)
  int packed_transport;        // 1 (default): ship MAPF_BITS and expand on the host when supported
  uint32_t* hs_bits;           // device: the packed observation
  uint32_t* hp_bits;           // pinned host staging of the same size
  size_t bits_words;
  struct MapfUnpackPool* pool;
  uint32_t* hs_prog;           // device: cumulative word count after each chunk (constant per handle)
  volatile uint32_t* hp_prog;  // pinned host word: the device copies hs_prog[c] here behind chunk c
  char err[512];
};

extern "C" {
struct MapfUnpackPool* mapf_unpack_pool_create(int threads);
void mapf_unpack_pool_destroy(struct MapfUnpackPool* p);
int mapf_host_unpack_range(const uint32_t* bits, uint64_t first_cell, uint64_t n_cells, void* out, int elem_bytes);
int mapf_unpack_pool_expand_streamed(struct MapfUnpackPool* p, const uint32_t* bits, void* out, size_t cells,
                                     int elem_bytes, const volatile uint32_t* ready_words, int (*poll)(void*),
                                     void* poll_arg);
}

static thread_local char g_create_err[512] = "";

static int fail(mapf_handle* h, int code, const char* fmt, ...) {
  char* dst = h ? h->err : g_create_err;
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(dst, 512, fmt, ap);
  va_end(ap);
  return code;
}

static int cuda_fail(mapf_handle* h, cudaError_t e, const char* what) {
  return fail(h, MAPF_ERR_CUDA, "%s: %s (%s)", what, cudaGetErrorName(e), cudaGetErrorString(e));
}

#define CK(call)                                               \
  do {                                                         \
    cudaError_t _e = (call);                                   \
    if (_e != cudaSuccess) return cuda_fail(h, _e, #call);     \
  } while (0)

static int gcd_i(int a, int b) { return b == 0 ? a : gcd_i(b, a % b); }
static int align_up(int x, int a) { return (x + a - 1) / a * a; }

static void compute_layout(const MapfDims& d, int epb, MapfTileLayout* L, bool second_grid = true) {
  const int na = epb * d.N;
  const bool fov = d.F > 0;
  int off = 0;
  auto take = [&](int bytes) {
    int o = off;
    off = align_up(off + bytes, 16);
    return o;
  };
  // ---- alive for the whole kernel: maps, occupancy, positions, goals
  L->obst_off = take((d.shared_map ? 1 : epb) * d.bm_words * 4);
  L->guard_off[0] = take(16);   // canaries (mapf_tile_kernel plants and verifies them: MAPF_FLAG_INTERNAL)
  L->agt_off = take(fov ? epb * d.bm_words * 4 : 16);
  L->grida_off = take(epb * d.grid_bytes);
  // PRIMAL needs the second grid only for the mid-sweep outputs (a copy of the pre-sweep ids)
  L->gridb_off = second_grid ? take(epb * d.grid_bytes) : L->grida_off;
  L->guard_off[1] = take(16);
  L->posnew_off = take(2 * na);
  L->goal_off = take(2 * na);
  L->guard_off[2] = take(16);
  // ---- step-phase scratch; dead once the state has been written back, so the observation's bit strings reuse
  //      the same bytes (the kernel puts a barrier between the two uses)
  const int scratch0 = off;
  L->posold_off = take(2 * na);
  L->mv_off = take(4 * na);
  L->res_off = take(na);
  L->dep_off = take(na);
  L->act_off = take(na);
  L->status_off = take(na);
  L->done_off = take(na);
  L->flag_off = take(na);
  L->avail_off = take(na);
  L->nextmid_off = take(na);
  L->node_off = take(na);
  L->edge_off = take(na);
  L->isint_off = take(na);
  L->atgoal_off = take(na);
  L->rew_off = take(8 * na);
  L->envrew_off = take(8 * epb * ((d.N + 31) / 32));
  L->envterm_off = take(epb);
  L->envcnt_off = take(4 * epb);
  L->envcnt2_off = take(4 * epb);
  L->envstep_off = take(4 * epb);
  if (d.mode != MAPF_MODE_PRIMAL) {   // sum(rewards) of GRID / PARTIAL (py_sum_* in mapf_kernels.cu)
    L->envff_off = take(4 * epb);
    L->pre_off = take(8 * na);
  } else {
    L->envff_off = L->pre_off = scratch0;
  }
  if (d.diag) {   // PRIMAL with DIAGONAL_MOVEMENT only: the common layouts stay as they are
    L->pastold_off = take(2 * na);
    L->pastnew_off = take(2 * na);
    L->mask16_off = take(2 * na);
    L->nextmid16_off = take(2 * na);
  } else {
    L->pastold_off = L->pastnew_off = L->mask16_off = L->nextmid16_off = scratch0;
  }
  const int scratch1 = off;
  const int str_bytes = fov ? ((na + d.G - 1) / d.G) * d.GW * 4 + 16 : 16;
  L->str_off = scratch0;
  off = align_up(scratch0 + (str_bytes > scratch1 - scratch0 ? str_bytes : scratch1 - scratch0), 16);
  L->guard_off[3] = take(16);
  L->total_bytes = off;
}

static const int kMaxSmem = 227 * 1024;

// Environments per tile: up to 128 agents per 128-thread block, a multiple of the alignment the
// 16-byte observation chunks need, and -- for small batches, which are one latency-bound wave -- enough tiles to
// put six on every one of the 148 SMs (c2: 4096 envs x 8 agents run 8.7 us per step with 4 envs per tile, 9.3 with 6,
// 10.2 us with 16).
static int choose_epb(MapfDims& d, MapfTileLayout* L) {
  const int lcm = d.G % 4 == 0 ? d.G : (4 % d.G == 0 ? 4 : d.G * 4 / gcd_i(d.G, 4));
  const int mult = lcm / gcd_i(d.N, lcm);
  int epb = (MAPF_TILE_THREADS / d.N) / mult * mult;
  if (epb < mult) epb = mult;
  while (epb > mult && (d.E + epb - 1) / epb < 6 * 148 && epb * d.N > 32) epb -= mult;
  if (const char* env = getenv("MAPF_B200_EPB")) {   // tuning knob for experiments: environments per tile
    const int want = atoi(env);
    if (want >= mult) epb = want / mult * mult;
  }
  for (;;) {
    compute_layout(d, epb, L);
    if (L->total_bytes <= kMaxSmem) break;
    if (epb <= mult) {
      // large maps with an agent count that is not a multiple of the group size: give up the 16-byte alignment of
      // the tile's observation block (the kernel has a path for it) rather than refuse the configuration
      epb = mult - 1;
      while (epb >= 1) {
        compute_layout(d, epb, L);
        if (L->total_bytes <= kMaxSmem) break;
        --epb;
      }
      if (epb < 1) return -1;
      break;
    }
    epb -= mult;
  }
  d.epb = epb;
  return 0;
}

extern "C" {

const char* mapf_last_error(const mapf_handle* h) { return h ? h->err : g_create_err; }
int mapf_abi_version(void) { return MAPF_ABI_VERSION; }
const char* mapf_build_arch(void) { return "sm_100a"; }
int64_t mapf_launch_count(const mapf_handle* h) { return h ? h->launches : 0; }

void mapf_default_cfg(mapf_cfg* c) {
  memset(c, 0, sizeof(*c));
  c->abi_version = MAPF_ABI_VERSION;
  c->n_envs = 1;
  c->n_agents = 4;          /* GRID:25 */
  c->height = c->width = 10;
  c->mode = MAPF_MODE_GRID;
  c->obs_mode = MAPF_OBS_FULLMAP;
  c->fov = 10;              /* PRIMAL:175 observation_size=10 */
  c->shared_map = 0;
  c->episode_limit = 10000; /* GRID:26 */
  c->goal_dist = 0;
  c->collect_stats = 1;
  c->step_reward = -0.01;   /* GRID:29 */
  c->collide_reward = -10;  /* GRID:30 */
  c->action_cost = -0.3;    /* PRIMAL:25 */
  c->idle_cost = -0.5;
  c->goal_reward = 0.0;
  c->collision_reward = -2.0;
  c->reward_sum_mode = 0;
  c->step_reward_is_int = 0;
  c->collide_reward_is_int = 1;
  c->blocking_reward = 0;
  c->blocking_cost = -1.0;  /* PRIMAL:25 */
}

int mapf_destroy(mapf_handle* h) {
  if (!h) return MAPF_OK;
  cudaFree(h->S.obst_bits);
  cudaFree(h->S.pos);
  cudaFree(h->S.goal);
  cudaFree(h->S.start);
  cudaFree(h->S.done);
  cudaFree(h->S.prev_action);
  cudaFree(h->S.step_count);
  cudaFree(h->S.goal_dist);
  cudaFree((void*)h->S.mag_lut);
  cudaFree((void*)h->S.vec_lut);
  cudaFree(h->S.stats);
  cudaFree(h->S.bfs_list);
  cudaFree(h->life_lists);
  cudaFree(h->life_cnts);
  cudaFree(h->S.err_flags);
  cudaFree(h->S.pos_prev);
  cudaFree(h->S.past);
  cudaFree(h->S.last_status);
  cudaFree(h->S.last_reward);
  cudaFree(h->S.at_goal);
  cudaFree(h->S.goal_cost);
  cudaFree(h->S.agent_steps);
  cudaFree(h->S.pnode);
  cudaFree(h->S.pedge);
  cudaFree(h->S.total_coll);
  cudaFree(h->S.terminated);
  cudaFree((void*)h->S.complete_lut);
  cudaFree(h->hs_actions);
  cudaFree(h->hs_reward);
  cudaFree(h->hs_terminated);
  cudaFree(h->hs_dones);
  cudaFree(h->hs_avail);
  cudaFree(h->hs_obs);
  cudaFree(h->hs_vec);
  cudaFree(h->hs_bits);
  if (h->hp_bits) cudaFreeHost(h->hp_bits);
  cudaFree(h->hs_prog);
  if (h->hp_prog) cudaFreeHost((void*)h->hp_prog);
  if (h->pool) mapf_unpack_pool_destroy(h->pool);
  delete h;
  return MAPF_OK;
}

int mapf_create(const mapf_cfg* c, mapf_handle** out) {
  mapf_handle* h = nullptr;
  if (!c || !out) return fail(h, MAPF_ERR_INVALID_ARG, "mapf_create: NULL argument");
  *out = nullptr;
  if (c->abi_version != MAPF_ABI_VERSION)
    return fail(h, MAPF_ERR_INVALID_ARG, "mapf_create: abi_version %d, library is %d", c->abi_version,
                MAPF_ABI_VERSION);
  if (c->n_envs < 1 || c->n_agents < 1 || c->n_agents > MAPF_MAX_AGENTS || c->height < 1 ||
      c->height > MAPF_MAX_SIDE || c->width < 1 || c->width > MAPF_MAX_SIDE)
    return fail(h, MAPF_ERR_INVALID_ARG, "mapf_create: need E>=1, 1<=N<=255, 1<=H,W<=255 (got E=%d N=%d H=%d W=%d)",
                c->n_envs, c->n_agents, c->height, c->width);
  if (c->mode != MAPF_MODE_GRID && c->mode != MAPF_MODE_PRIMAL && c->mode != MAPF_MODE_PARTIAL)
    return fail(h, MAPF_ERR_INVALID_ARG, "mapf_create: unknown mode %d", c->mode);
  if (c->obs_mode != MAPF_OBS_FULLMAP && c->obs_mode != MAPF_OBS_PRIMAL_FOV && c->obs_mode != MAPF_OBS_PARTIAL_WINDOW)
    return fail(h, MAPF_ERR_INVALID_ARG, "mapf_create: unknown obs_mode %d", c->obs_mode);
  if (c->obs_mode == MAPF_OBS_PARTIAL_WINDOW && c->mode != MAPF_MODE_PARTIAL)
    return fail(h, MAPF_ERR_UNSUPPORTED, "mapf_create: the PARTIAL window observation needs mode PARTIAL");
  if (c->blocking_reward) {
    if (c->mode != MAPF_MODE_PRIMAL || c->obs_mode != MAPF_OBS_PRIMAL_FOV)
      return fail(h, MAPF_ERR_UNSUPPORTED, "mapf_create: blocking_reward needs mode PRIMAL with the FOV observation");
    if (c->height > 64 || c->width > 64)
      return fail(h, MAPF_ERR_UNSUPPORTED, "mapf_create: blocking_reward supports maps up to 64 x 64");
  }
  if (c->diagonal_movement) {
    if (c->mode != MAPF_MODE_PRIMAL)
      return fail(h, MAPF_ERR_UNSUPPORTED, "mapf_create: diagonal_movement exists only in mode PRIMAL");
    if (c->blocking_reward)
      return fail(h, MAPF_ERR_UNSUPPORTED, "mapf_create: blocking_reward with diagonal_movement is not supported");
  }
  if (c->mode == MAPF_MODE_PARTIAL) {
    if (c->obs_window < 1 || c->obs_window > 255 || c->obs_knn_agents < 1 || c->obs_knn_agents > 254)
      return fail(h, MAPF_ERR_INVALID_ARG, "mapf_create: obs_window %d / obs_knn_agents %d out of range",
                  c->obs_window, c->obs_knn_agents);
    if (!c->complete_lut_host || c->complete_lut_len < 1)
      return fail(h, MAPF_ERR_INVALID_ARG, "mapf_create: mode PARTIAL needs complete_lut_host");
    if (c->episode_limit < 1) return fail(h, MAPF_ERR_INVALID_ARG, "mapf_create: episode_limit must be >= 1");
  }
  if (c->obs_mode == MAPF_OBS_PRIMAL_FOV) {
    if (c->mode != MAPF_MODE_PRIMAL)
      return fail(h, MAPF_ERR_UNSUPPORTED,
                  "mapf_create: the PRIMAL field-of-view observation needs one agent per cell (mode PRIMAL)");
    if (c->fov < 1 || c->fov > 127) return fail(h, MAPF_ERR_INVALID_ARG, "mapf_create: fov %d not in [1,127]", c->fov);
  }
  const long long need = (long long)(c->height - 1) * (c->height - 1) + (long long)(c->width - 1) * (c->width - 1);
  if (c->obs_mode == MAPF_OBS_PRIMAL_FOV && (!c->mag_lut_host || c->mag_lut_len <= need))
    return fail(h, MAPF_ERR_INVALID_ARG, "mapf_create: mag_lut_host must cover s in [0, %lld]", need);

  h = new (std::nothrow) mapf_handle();
  if (!h) return fail(nullptr, MAPF_ERR_ALLOC, "mapf_create: out of host memory");
  memset(h, 0, sizeof(*h));
  h->cfg = *c;
  h->cfg.mag_lut_host = nullptr;
  h->packed_transport = 1;
  cudaError_t e = cudaGetDevice(&h->device);
  if (e != cudaSuccess) {
    int rc = cuda_fail(nullptr, e, "cudaGetDevice");
    delete h;
    return rc;
  }
  MapfDims& d = h->d;
  d.E = c->n_envs;
  d.N = c->n_agents;
  d.H = c->height;
  d.W = c->width;
  d.HW = d.H * d.W;
  d.F = c->obs_mode == MAPF_OBS_PRIMAL_FOV ? c->fov : 0;
  d.P = d.F / 2 > 1 ? d.F / 2 : 1;
  // the PARTIAL observation reads its window straight from the padded bit rows: pad by half a window
  if (c->obs_mode == MAPF_OBS_PARTIAL_WINDOW && c->obs_window / 2 > d.P) d.P = c->obs_window / 2;
  d.PR = d.H + 2 * d.P;
  d.RW = ((d.W + 2 * d.P - 1) >> 5) + 2;
  d.RW |= 1;   // odd row stride: the bit rows of 32 agents spread over all shared-memory banks (an even stride
               // folds them onto half / a quarter of the banks: 54 % of c4's wavefronts were bank conflicts)
  d.bm_words = align_up(d.PR * d.RW, 4);
  d.GS = d.W + 2;
  d.grid_bytes = align_up((d.H + 2) * (d.W + 2), 16);
  d.shared_map = c->shared_map ? 1 : 0;
  d.mode = c->mode;
  d.obs_mode = c->obs_mode;
  d.episode_limit = c->episode_limit;
  d.inv_limit = 1.0 / (double)c->episode_limit;   // IEEE division on the host == __ddiv_rn(1.0, limit)
  d.sum_mode = c->reward_sum_mode;
  d.step_is_int = c->step_reward_is_int;
  d.collide_is_int = c->collide_reward_is_int;
  d.collect_stats = c->collect_stats;
  d.rsum_mode = (32 % c->n_agents == 0) ? 1 : ((c->n_agents % 32 == 0) ? 2 : 0);
  d.blocking = c->blocking_reward ? 1 : 0;
  d.blocking_cost = c->blocking_cost;
  d.diag = c->diagonal_movement ? 1 : 0;
  d.nact = d.diag ? 9 : 5;
  d.pW = c->obs_window;
  d.pK = c->obs_knn_agents;
  d.posz = 2 * d.pW * d.pW + 13 * d.pK;
  d.complete_len = c->complete_lut_len;
  d.p_move = c->move_reward;
  d.p_stay = c->stay_reward;
  d.p_stay_goal = c->stay_goal_reward;
  d.p_nc = c->node_collide_reward;
  d.p_ec = c->edge_collide_reward;
  d.p_envc = c->env_collide_reward;
  d.invN = (uint32_t)((0x100000000ULL + d.N - 1) / d.N);
  d.invW = (uint32_t)((0x100000000ULL + d.W - 1) / d.W);
  d.step_reward = c->step_reward;
  d.collide_reward = c->collide_reward;
  d.action_cost = c->action_cost;
  d.idle_cost = c->idle_cost;
  d.goal_reward = c->goal_reward;
  d.collision_reward = c->collision_reward;
  h->fov_fast = d.F > 0 && mapf_tile_has_fov(d.F);
  if (d.F > 0 && h->fov_fast) {
    const int nb = 4 * d.F * d.F;
    d.G = 32 / gcd_i(nb, 32);
    d.GW = d.G * nb / 32;
  } else {
    d.G = 1;
    d.GW = 0;
  }
  MapfDims dt = d;
  if (!h->fov_fast) dt.F = 0;   // the tile kernel then only steps; the generic kernel observes
  if (choose_epb(dt, &h->L) != 0) {
    int rc = fail(nullptr, MAPF_ERR_UNSUPPORTED, "mapf_create: one environment needs more than %d bytes of shared memory",
                  kMaxSmem);
    delete h;
    return rc;
  }
  d.epb = dt.epb;
  dt.epb = d.epb;
  compute_layout(dt, d.epb, &h->L_lite, d.mode != MAPF_MODE_PRIMAL);

#define ALLOC(ptr, bytes)                                                                          \
  do {                                                                                             \
    cudaError_t _e = cudaMalloc((void**)&(ptr), (bytes));                                          \
    if (_e != cudaSuccess) {                                                                       \
      int rc = fail(nullptr, MAPF_ERR_ALLOC, "mapf_create: cudaMalloc(%zu) for " #ptr ": %s",      \
                    (size_t)(bytes), cudaGetErrorString(_e));                                      \
      mapf_destroy(h);                                                                             \
      return rc;                                                                                   \
    }                                                                                              \
  } while (0)
  const size_t EN = (size_t)d.E * d.N;
  ALLOC(h->S.obst_bits, (size_t)(d.shared_map ? 1 : d.E) * d.bm_words * 4);
  ALLOC(h->S.pos, EN * 2);
  ALLOC(h->S.goal, EN * 2);
  ALLOC(h->S.start, EN * 2);
  ALLOC(h->S.done, EN);
  ALLOC(h->S.prev_action, EN);
  ALLOC(h->S.step_count, (size_t)d.E * 4);
  if (c->goal_dist || c->mode == MAPF_MODE_PARTIAL) ALLOC(h->S.goal_dist, EN * d.HW * 2);
  if (d.diag) ALLOC(h->S.past, EN * 2);
  if (c->blocking_reward) {
    ALLOC(h->S.pos_prev, EN * 2);
    ALLOC(h->S.last_status, EN);
    ALLOC(h->S.last_reward, EN * 8);
  }
  if (c->mode == MAPF_MODE_PARTIAL) {
    ALLOC(h->S.at_goal, EN);
    ALLOC(h->S.goal_cost, EN * 4);
    ALLOC(h->S.agent_steps, EN * 4);
    ALLOC(h->S.pnode, EN);
    ALLOC(h->S.pedge, EN);
    ALLOC(h->S.total_coll, (size_t)d.E * 8);
    ALLOC(h->S.terminated, (size_t)d.E);
    double* clut = nullptr;
    ALLOC(clut, (size_t)c->complete_lut_len * 8);
    h->S.complete_lut = clut;
  }
  ALLOC(h->S.stats, MAPF_N_STATS * 8);
  ALLOC(h->S.bfs_list, (2 * (size_t)h->d.E * h->d.N + 2) * 4);
  ALLOC(h->S.err_flags, 4);
  double* lut = nullptr;
  const int lut_len = c->mag_lut_host ? c->mag_lut_len : 1;
  ALLOC(lut, (size_t)lut_len * 8);
  h->S.mag_lut = lut;
  h->mag_lut_len = lut_len;
#undef ALLOC
  // blocking initialisation: the handle is usable on any stream afterwards
  e = cudaMemset(h->S.obst_bits, 0xff, (size_t)(d.shared_map ? 1 : d.E) * d.bm_words * 4);
  if (e == cudaSuccess) e = cudaMemset(h->S.pos, 0, EN * 2);
  if (e == cudaSuccess) e = cudaMemset(h->S.goal, 0, EN * 2);
  if (e == cudaSuccess) e = cudaMemset(h->S.start, 0, EN * 2);
  if (e == cudaSuccess) e = cudaMemset(h->S.done, 0, EN);
  if (e == cudaSuccess) e = cudaMemset(h->S.prev_action, 0, EN);
  if (e == cudaSuccess) e = cudaMemset(h->S.step_count, 0, (size_t)d.E * 4);
  if (e == cudaSuccess && d.diag) e = cudaMemset(h->S.past, 0, EN * 2);
  if (e == cudaSuccess) e = cudaMemset(h->S.stats, 0, MAPF_N_STATS * 8);
  if (e == cudaSuccess) e = cudaMemset(h->S.err_flags, 0, 4);
  if (e == cudaSuccess) {
    if (c->mag_lut_host) e = cudaMemcpy(lut, c->mag_lut_host, (size_t)lut_len * 8, cudaMemcpyHostToDevice);
    else e = cudaMemset(lut, 0, 8);
  }
  if (e == cudaSuccess && c->obs_mode == MAPF_OBS_PRIMAL_FOV) {
    // goal unit vectors for every (|dx|, |dy|): the same correctly rounded IEEE division CPython performs for
    // `dx / mag` (PRIMAL:383-384), done once on the host instead of two DDIVs per agent per step
    const size_t n = (size_t)d.H * d.W;
    double* tab = (double*)malloc(n * 32);
    double* dv = nullptr;
    if (!tab) e = cudaErrorMemoryAllocation;
    if (e == cudaSuccess) e = cudaMalloc((void**)&dv, n * 32);
    if (e == cudaSuccess) {
      const double* ml = (const double*)c->mag_lut_host;
      for (int a = 0; a < d.H; ++a)
        for (int b = 0; b < d.W; ++b) {
          const double mag = ml[a * a + b * b];
          double* t = tab + ((size_t)a * d.W + b) * 4;
          t[0] = mag != 0.0 ? (double)a / mag : (double)a;
          t[1] = mag != 0.0 ? (double)b / mag : (double)b;
          t[2] = mag;
          t[3] = 0.0;
        }
      e = cudaMemcpy(dv, tab, n * 32, cudaMemcpyHostToDevice);
      h->S.vec_lut = dv;
    }
    free(tab);
  }
  if (e == cudaSuccess && c->mode == MAPF_MODE_PARTIAL) {
    e = cudaMemcpy((void*)h->S.complete_lut, c->complete_lut_host, (size_t)c->complete_lut_len * 8,
                   cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemset(h->S.at_goal, 0, EN);
    if (e == cudaSuccess) e = cudaMemset(h->S.goal_cost, 0xff, EN * 4);
    if (e == cudaSuccess) e = cudaMemset(h->S.agent_steps, 0, EN * 4);
    if (e == cudaSuccess) e = cudaMemset(h->S.pnode, 0, EN);
    if (e == cudaSuccess) e = cudaMemset(h->S.pedge, 0, EN);
    if (e == cudaSuccess) e = cudaMemset(h->S.total_coll, 0, (size_t)d.E * 8);
    if (e == cudaSuccess) e = cudaMemset(h->S.terminated, 0, (size_t)d.E);
  }
  if (e == cudaSuccess && h->L.total_bytes > 48 * 1024)
    e = (cudaError_t)mapf_configure_tile(h->fov_fast ? d.F : 0, d.diag ? MAPF_MODE_PRIMAL_DIAG : d.mode, h->L.total_bytes);
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    int rc = cuda_fail(nullptr, e, "mapf_create: initialisation");
    mapf_destroy(h);
    return rc;
  }
  *out = h;
  return MAPF_OK;
}

static int check_aligned(mapf_handle* h, const void* p, const char* name) {
  if (p && (((uintptr_t)p) & 15) != 0)
    return fail(h, MAPF_ERR_INVALID_ARG, "%s must be 16-byte aligned", name);
  return MAPF_OK;
}

int mapf_reset(mapf_handle* h, const int8_t* map_dev, const int16_t* starts_dev, const int16_t* goals_dev,
               const uint8_t* env_mask_dev, void* stream) {
  if (!h) return MAPF_ERR_INVALID_ARG;
  if (map_dev) {
    CK((cudaError_t)mapf_launch_build_obst(h->d, h->S, map_dev, env_mask_dev, stream));
    h->launches++;
  }
  CK((cudaError_t)mapf_launch_reset(h->d, h->S, starts_dev, goals_dev, env_mask_dev, stream));
  h->launches++;
  if (h->d.mode == MAPF_MODE_PARTIAL) {   // __setup_agent_goal_dist runs inside the reference's reset, PARTIAL:130, 928
    int n = 0;
    CK((cudaError_t)mapf_launch_bfs(h->d, h->S, nullptr, env_mask_dev, h->S.goal_dist, 0, stream, &n));
    h->launches += n;
  }
  return MAPF_OK;
}

int mapf_set_goals(mapf_handle* h, const int16_t* goals_dev, const uint8_t* dirty_dev, void* stream) {
  if (!h || !goals_dev) return fail(h, MAPF_ERR_INVALID_ARG, "mapf_set_goals: NULL argument");
  CK((cudaError_t)mapf_launch_set_goals(h->d, h->S, goals_dev, dirty_dev, stream));
  h->launches++;
  return MAPF_OK;
}

int mapf_pop_goals(mapf_handle* h, const int16_t* queue_dev, int32_t* head_dev, int queue_len, uint8_t* dirty_dev,
                   void* stream) {
  if (!h || !queue_dev || !head_dev) return fail(h, MAPF_ERR_INVALID_ARG, "mapf_pop_goals: NULL argument");
  if (queue_len < 1) return fail(h, MAPF_ERR_INVALID_ARG, "mapf_pop_goals: queue_len must be >= 1");
  CK((cudaError_t)mapf_launch_pop_goals(h->d, h->S, queue_dev, head_dev, queue_len, dirty_dev, stream));
  h->launches++;
  return MAPF_OK;
}

int mapf_lifelong_bind(mapf_handle* h, const int16_t* queue_dev, int32_t* head_dev, int queue_len) {
  if (!h) return MAPF_ERR_INVALID_ARG;
  if (!queue_dev) {
    h->life_queue = nullptr;
    h->life_head = nullptr;
    h->life_last = -1;
    return MAPF_OK;
  }
  if (!head_dev) return fail(h, MAPF_ERR_INVALID_ARG, "mapf_lifelong_bind: head_dev is NULL");
  if (queue_len < 1) return fail(h, MAPF_ERR_INVALID_ARG, "mapf_lifelong_bind: queue_len must be >= 1");
  if (h->d.mode != MAPF_MODE_PRIMAL)
    return fail(h, MAPF_ERR_UNSUPPORTED, "mapf_lifelong_bind: PRIMAL mode only (use mapf_pop_goals)");
  const size_t EN = (size_t)h->d.E * h->d.N;
  if (!h->life_lists) {
    CK(cudaMalloc((void**)&h->life_lists, 2 * EN * 4));
    CK(cudaMalloc((void**)&h->life_cnts, 4 * 4));
  }
  CK(cudaMemset(h->life_cnts, 0, 4 * 4));
  h->life_queue = queue_dev;
  h->life_head = head_dev;
  h->life_Q = queue_len;
  h->life_slot = 0;
  h->life_last = -1;
  return MAPF_OK;
}

int mapf_bfs_popped(mapf_handle* h, int16_t* dist_dev, void* stream) {
  if (!h) return MAPF_ERR_INVALID_ARG;
  if (!h->life_queue) return fail(h, MAPF_ERR_INVALID_ARG, "mapf_bfs_popped: no queues bound (mapf_lifelong_bind)");
  int16_t* dist = dist_dev ? dist_dev : h->S.goal_dist;
  if (!dist) return fail(h, MAPF_ERR_INVALID_ARG, "mapf_bfs_popped: no output (dist_dev NULL and cfg.goal_dist == 0)");
  int rc;
  if ((rc = check_aligned(h, dist, "dist_dev")) != MAPF_OK) return rc;
  if (h->life_last < 0) return MAPF_OK;   // no step since the queues were bound
  const size_t EN = (size_t)h->d.E * h->d.N;
  int n = 0;
  CK((cudaError_t)mapf_launch_bfs(h->d, h->S, nullptr, nullptr, dist, 0, stream, &n, h->life_lists + h->life_last * EN,
                                  h->life_cnts + 2 * h->life_last));
  h->launches += n;
  h->life_last = -1;                      // consumed
  return MAPF_OK;
}

// MAPF_BITS output: a specialised field-of-view kernel, and every tile holds whole observation groups (tile strings
// start on word boundaries).
static bool bits_supported(const mapf_handle* h) {
  return h->d.obs_mode == MAPF_OBS_PRIMAL_FOV && h->fov_fast && (h->d.epb * h->d.N) % h->d.G == 0;
}

static int run_tile(mapf_handle* h, const void* actions, int act_dtype, int lo, int hi, const mapf_step_out* out,
                    void* obs, int obs_dtype, double* vec, void* stream, int n_steps = 1) {
  MapfTileArgs A;
  memset(&A, 0, sizeof(A));
  A.T = n_steps;
  A.debug_corrupt = h->debug_corrupt;
  h->debug_corrupt = 0;
  A.actions = actions;
  A.act_dtype = act_dtype;
  A.do_step = actions != nullptr;
  A.agent_lo = lo;
  A.agent_hi = hi;
  if (out) A.out = *out;
  int rc;
  if ((rc = check_aligned(h, actions, "actions_dev")) != MAPF_OK) return rc;
  if ((rc = check_aligned(h, obs, "obs_dev")) != MAPF_OK) return rc;
  if ((rc = check_aligned(h, vec, "vec_dev")) != MAPF_OK) return rc;
  if (actions && act_dtype != MAPF_U8 && act_dtype != MAPF_I64)
    return fail(h, MAPF_ERR_INVALID_ARG, "actions must be MAPF_U8 or MAPF_I64");
  const bool fov = h->d.obs_mode == MAPF_OBS_PRIMAL_FOV;
  const bool pwin = h->d.obs_mode == MAPF_OBS_PARTIAL_WINDOW;
  if (obs) {
    if (fov && obs_dtype == MAPF_BITS && !bits_supported(h))
      return fail(h, MAPF_ERR_UNSUPPORTED, "MAPF_BITS observations are not available for this configuration "
                                           "(mapf_obs_bits_supported)");
    if (fov && obs_dtype != MAPF_U8 && obs_dtype != MAPF_F32 && obs_dtype != MAPF_BITS)
      return fail(h, MAPF_ERR_INVALID_ARG, "FOV observations are MAPF_U8, MAPF_F32 or MAPF_BITS");
    if (pwin && obs_dtype != MAPF_F64 && obs_dtype != MAPF_F32)
      return fail(h, MAPF_ERR_INVALID_ARG, "PARTIAL window observations are MAPF_F64 or MAPF_F32");
    if (!fov && !pwin && obs_dtype != MAPF_I8)
      return fail(h, MAPF_ERR_INVALID_ARG, "full-map observations are MAPF_I8");
  }
  MapfDims d = h->d;
  const bool generic_obs = fov && !h->fov_fast && (obs || vec);
  if (fov && !h->fov_fast) d.F = 0, d.obs_mode = MAPF_OBS_FULLMAP;   // the tile kernel skips the observation
  if (!generic_obs && !pwin) {
    A.obs = obs;
    A.obs_dtype = obs_dtype;
    A.vec = fov ? vec : nullptr;
  }
  const bool any_out = A.do_step || A.obs || A.vec || A.out.avail_dev;
  const bool blocking = A.do_step && h->d.blocking;
  int8_t* user_status = A.out.status_dev;
  double* user_agent_reward = A.out.agent_reward_dev;
  if (blocking) {
    // the blocking-reward kernel that follows the sweep needs the pre-sweep positions, the statuses and the per-agent
    // rewards: route them into the handle's own buffers (the hot kernel stays free of blocking-specific code)
    const size_t EN = (size_t)h->d.E * h->d.N;
    CK(cudaMemcpyAsync(h->S.pos_prev, h->S.pos, EN * 2, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    A.out.status_dev = h->S.last_status;
    A.out.agent_reward_dev = h->S.last_reward;
  }
  const bool life = h->life_queue && A.do_step && n_steps == 1 && lo == 0 && hi == h->d.N;
  if (life) {   // the write-back of this launch pops the goal queues (mapf_lifelong_bind)
    A.life_queue = h->life_queue;
    A.life_head = h->life_head;
    A.life_Q = h->life_Q;
    A.life_list = h->life_lists + (size_t)h->life_slot * h->d.E * h->d.N;
    A.life_cnt = h->life_cnts + 2 * h->life_slot;
  }
  if (any_out) {
    const bool need_mid = A.out.done_mid_dev || A.out.next_mid_dev;
    CK((cudaError_t)mapf_launch_tile(d, need_mid ? h->L : h->L_lite, h->S, A, stream));
    h->launches++;
    if (life) {
      h->life_last = h->life_slot;
      h->life_slot ^= 1;
    }
  }
  if (blocking) {   // PRIMAL:579-585: stay-on-goal rewards get the blocking term
    int n = 0;
    mapf_step_out o2 = A.out;
    o2.agent_reward_dev = user_agent_reward;
    CK((cudaError_t)mapf_launch_blocking(h->d, h->S, lo, hi, o2, stream, &n));
    h->launches += n;
    if (user_status)
      CK(cudaMemcpyAsync(user_status, h->S.last_status, (size_t)h->d.E * h->d.N, cudaMemcpyDeviceToDevice,
                         (cudaStream_t)stream));
  }
  if (pwin && obs) {
    CK((cudaError_t)mapf_launch_partial_obs(h->d, h->S, obs, obs_dtype == MAPF_F32, (long long*)h->partial_state_out,
                                            stream));
    h->launches++;
  }
  if (generic_obs) {
    CK((cudaError_t)mapf_launch_observe_generic(h->d, h->S, obs_dtype == MAPF_U8 ? (uint8_t*)obs : nullptr,
                                                obs_dtype == MAPF_F32 ? (float*)obs : nullptr, vec, stream));
    h->launches++;
  }
  return MAPF_OK;
}

int mapf_step_agents(mapf_handle* h, const void* actions_dev, int act_dtype, int agent_lo, int agent_hi,
                     const mapf_step_out* out, void* stream) {
  if (!h || !actions_dev) return fail(h, MAPF_ERR_INVALID_ARG, "mapf_step: NULL argument");
  if (agent_lo < 0 || agent_hi > h->d.N || agent_lo >= agent_hi)
    return fail(h, MAPF_ERR_INVALID_ARG, "mapf_step_agents: bad range [%d, %d)", agent_lo, agent_hi);
  if (h->d.mode != MAPF_MODE_PRIMAL && (agent_lo != 0 || agent_hi != h->d.N))
    return fail(h, MAPF_ERR_UNSUPPORTED, "mapf_step_agents: partial sweeps exist only in PRIMAL mode");
  return run_tile(h, actions_dev, act_dtype, agent_lo, agent_hi, out, nullptr, 0, nullptr, stream);
}

int mapf_step(mapf_handle* h, const void* actions_dev, int act_dtype, const mapf_step_out* out, void* stream) {
  if (!h) return MAPF_ERR_INVALID_ARG;
  return mapf_step_agents(h, actions_dev, act_dtype, 0, h->d.N, out, stream);
}

int mapf_observe(mapf_handle* h, void* obs_dev, int obs_dtype, double* vec_dev, void* stream) {
  if (!h || (!obs_dev && !vec_dev)) return fail(h, MAPF_ERR_INVALID_ARG, "mapf_observe: NULL argument");
  return run_tile(h, nullptr, 0, 0, h->d.N, nullptr, obs_dev, obs_dtype, vec_dev, stream);
}

int mapf_step_observe(mapf_handle* h, const void* actions_dev, int act_dtype, const mapf_step_out* out,
                      void* obs_dev, int obs_dtype, double* vec_dev, void* stream) {
  if (!h || !actions_dev) return fail(h, MAPF_ERR_INVALID_ARG, "mapf_step_observe: NULL argument");
  return run_tile(h, actions_dev, act_dtype, 0, h->d.N, out, obs_dev, obs_dtype, vec_dev, stream);
}

// Can `n_steps` consecutive steps with these outputs run as ONE launch (the ROLL instantiation of the tile kernel)?
static bool rollout_in_kernel(const mapf_handle* h, const void* obs, int obs_dtype, const double* vec) {
  const MapfDims& d = h->d;
  if (!mapf_tile_has_rollout(d.mode) || d.diag || d.blocking) return false;
  if (h->life_queue) return false;                                          // the queues are popped once per launch
  if (d.epb * d.N > MAPF_TILE_THREADS) return false;                       // one thread per agent of the tile
  const bool fov = d.obs_mode == MAPF_OBS_PRIMAL_FOV;
  if (fov && !h->fov_fast && (obs || vec)) return false;                    // generic-F observation kernel
  if (d.obs_mode == MAPF_OBS_PARTIAL_WINDOW) return false;
  if (obs && obs_dtype == MAPF_BITS && (((size_t)d.E * d.N * 4 * d.F * d.F) & 31) != 0) return false;
  if (obs && obs_dtype != MAPF_BITS && fov &&
      (((size_t)d.E * d.N * 4 * d.F * d.F * (obs_dtype == MAPF_F32 ? 4 : 1)) & 15) != 0)
    return false;                                                          // every step's block starts 16-byte aligned
  return true;
}

int mapf_rollout_in_one_launch(const mapf_handle* h, int obs_dtype) {
  if (!h) return 0;
  int dummy = 0;
  return rollout_in_kernel(h, &dummy, obs_dtype, nullptr) ? 1 : 0;
}

// How mapf_rollout runs: 2 = the pipelined kernel (small PRIMAL batches: every tile resident at once, step t+1 under
// the observation of step t), 1 = the tile kernel's in-kernel loop, 0 = n_steps consecutive launches.
static int rollout_plan(const mapf_handle* h, int n_steps, bool has_obs, int obs_dtype, bool mid_outputs) {
  static const int pipe_env = getenv("MAPF_B200_PIPE") ? atoi(getenv("MAPF_B200_PIPE")) : -1;   // experiments: 0 off, 1 force
  const MapfDims& d = h->d;
  int dummy = 0;
  const bool in_kernel = rollout_in_kernel(h, has_obs ? &dummy : nullptr, obs_dtype, nullptr);
  if (n_steps >= 2 && has_obs && in_kernel && !mid_outputs && pipe_env != 0 && mapf_pipe_supported(d) &&
      (pipe_env == 1 || mapf_pipe_tiles(d) <= 148 * 8))   // all tiles resident at 8 blocks per SM
    return 2;
  return (n_steps == 1 || in_kernel) ? 1 : 0;
}

int mapf_rollout_plan(const mapf_handle* h, int n_steps, int obs_dtype, int mid_outputs) {
  return h ? rollout_plan(h, n_steps, true, obs_dtype, mid_outputs != 0) : 0;
}

int mapf_rollout(mapf_handle* h, const void* actions_dev, int act_dtype, int n_steps, const mapf_step_out* out,
                 void* obs_dev, int obs_dtype, double* vec_dev, void* stream) {
  if (!h || !actions_dev) return fail(h, MAPF_ERR_INVALID_ARG, "mapf_rollout: NULL argument");
  if (n_steps < 1) return fail(h, MAPF_ERR_INVALID_ARG, "mapf_rollout: n_steps must be >= 1");
  const MapfDims& d = h->d;
  const int plan = rollout_plan(h, n_steps, obs_dev != nullptr, obs_dtype,
                                out && (out->done_mid_dev || out->next_mid_dev || out->blocking_dev));
  if (plan == 2) {
    int rc;
    if ((rc = check_aligned(h, actions_dev, "actions_dev")) != MAPF_OK) return rc;
    if ((rc = check_aligned(h, obs_dev, "obs_dev")) != MAPF_OK) return rc;
    if ((rc = check_aligned(h, vec_dev, "vec_dev")) != MAPF_OK) return rc;
    if (act_dtype != MAPF_U8 && act_dtype != MAPF_I64)
      return fail(h, MAPF_ERR_INVALID_ARG, "actions must be MAPF_U8 or MAPF_I64");
    if (obs_dtype != MAPF_U8 && obs_dtype != MAPF_F32 && obs_dtype != MAPF_BITS)
      return fail(h, MAPF_ERR_INVALID_ARG, "FOV observations are MAPF_U8, MAPF_F32 or MAPF_BITS");
    MapfTileArgs A;
    memset(&A, 0, sizeof(A));
    A.T = n_steps;
    A.debug_corrupt = h->debug_corrupt;
    h->debug_corrupt = 0;
    A.actions = actions_dev;
    A.act_dtype = act_dtype;
    A.do_step = 1;
    A.agent_lo = 0;
    A.agent_hi = d.N;
    if (out) A.out = *out;
    A.obs = obs_dev;
    A.obs_dtype = obs_dtype;
    A.vec = vec_dev;
    CK((cudaError_t)mapf_launch_pipe(d, h->S, A, stream));
    h->launches++;
    return MAPF_OK;
  }
  if (plan == 1)
    return run_tile(h, actions_dev, act_dtype, 0, d.N, out, obs_dev, obs_dtype, vec_dev, stream, n_steps);
  // configurations without the in-kernel loop: the same result from n_steps consecutive launches
  const size_t EN = (size_t)d.E * d.N, E = (size_t)d.E;
  size_t obs_step = 0;   // bytes of one step's observation
  if (obs_dev) {
    if (d.obs_mode == MAPF_OBS_PRIMAL_FOV) {
      const size_t cells = EN * 4 * d.F * d.F;
      if (obs_dtype == MAPF_BITS) {
        if (cells & 31) return fail(h, MAPF_ERR_UNSUPPORTED, "mapf_rollout: MAPF_BITS needs E*N*4*F*F to be a multiple of 32");
        obs_step = cells / 8;
      } else {
        obs_step = cells * (obs_dtype == MAPF_F32 ? 4 : 1);
      }
    } else if (d.obs_mode == MAPF_OBS_PARTIAL_WINDOW) {
      obs_step = EN * d.posz * (obs_dtype == MAPF_F32 ? 4 : 8);
    } else {
      obs_step = E * d.HW;
    }
    if (obs_step & 15) return fail(h, MAPF_ERR_UNSUPPORTED, "mapf_rollout: one step's observation must be a multiple of 16 bytes");
  }
  if (act_dtype != MAPF_U8 && act_dtype != MAPF_I64) return fail(h, MAPF_ERR_INVALID_ARG, "actions must be MAPF_U8 or MAPF_I64");
  if ((EN * (act_dtype == MAPF_I64 ? 8 : 1)) & 15)
    return fail(h, MAPF_ERR_UNSUPPORTED, "mapf_rollout: one step's actions must be a multiple of 16 bytes");
  for (int t = 0; t < n_steps; ++t) {
    mapf_step_out o;
    memset(&o, 0, sizeof(o));
    if (out) {
      o = *out;
#define ADV(field, count) if (o.field) o.field += (size_t)t * (count)
      ADV(reward_dev, E); ADV(terminated_dev, E); ADV(agent_reward_dev, EN); ADV(dones_dev, EN); ADV(status_dev, EN);
      ADV(node_dev, EN); ADV(edge_dev, EN); ADV(valid_dev, EN); ADV(done_mid_dev, EN); ADV(next_mid_dev, EN * d.nact);
      ADV(avail_dev, EN * d.nact); ADV(blocking_dev, EN);
#undef ADV
    }
    const char* a = (const char*)actions_dev + (size_t)t * EN * (act_dtype == MAPF_I64 ? 8 : 1);
    void* ob = obs_dev ? (void*)((char*)obs_dev + (size_t)t * obs_step) : nullptr;
    double* vc = vec_dev ? vec_dev + (size_t)t * EN * 3 : nullptr;
    const int rc = run_tile(h, a, act_dtype, 0, d.N, &o, ob, obs_dtype, vc, stream);
    if (rc != MAPF_OK) return rc;
  }
  return MAPF_OK;
}

int mapf_avail(mapf_handle* h, uint8_t* avail_dev, void* stream) {
  if (!h || !avail_dev) return fail(h, MAPF_ERR_INVALID_ARG, "mapf_avail: NULL argument");
  mapf_step_out out;
  memset(&out, 0, sizeof(out));
  out.avail_dev = avail_dev;
  return run_tile(h, nullptr, 0, 0, h->d.N, &out, nullptr, 0, nullptr, stream);
}

int mapf_avail_prev(mapf_handle* h, const uint8_t* prev_dev, uint8_t* avail_dev, void* stream) {
  if (!h || !avail_dev || !prev_dev) return fail(h, MAPF_ERR_INVALID_ARG, "mapf_avail_prev: NULL argument");
  mapf_step_out out;
  memset(&out, 0, sizeof(out));
  out.avail_dev = avail_dev;
  // a launch without a step never writes the state: pointing it at the caller's prev_action array is a pure query
  uint8_t* stored = h->S.prev_action;
  h->S.prev_action = const_cast<uint8_t*>(prev_dev);
  const int rc = run_tile(h, nullptr, 0, 0, h->d.N, &out, nullptr, 0, nullptr, stream);
  h->S.prev_action = stored;
  return rc;
}

int mapf_bfs(mapf_handle* h, const uint8_t* dirty_dev, int16_t* dist_dev, int primal_costs, void* stream) {
  if (!h) return MAPF_ERR_INVALID_ARG;
  int16_t* dist = dist_dev ? dist_dev : h->S.goal_dist;
  if (!dist) return fail(h, MAPF_ERR_INVALID_ARG, "mapf_bfs: no output (dist_dev NULL and cfg.goal_dist == 0)");
  int rc;
  if ((rc = check_aligned(h, dist, "dist_dev")) != MAPF_OK) return rc;
  int n = 0;
  CK((cudaError_t)mapf_launch_bfs(h->d, h->S, dirty_dev, nullptr, dist, primal_costs, stream, &n));
  h->launches += n;
  return MAPF_OK;
}

int mapf_set_prev_actions(mapf_handle* h, const uint8_t* prev_dev, void* stream) {
  if (!h || !prev_dev) return fail(h, MAPF_ERR_INVALID_ARG, "mapf_set_prev_actions: NULL argument");
  CK(cudaMemcpyAsync(h->S.prev_action, prev_dev, (size_t)h->d.E * h->d.N, cudaMemcpyDeviceToDevice,
                     (cudaStream_t)stream));
  return MAPF_OK;
}

int mapf_get_positions(mapf_handle* h, int16_t* pos_dev, void* stream) {
  if (!h || !pos_dev) return fail(h, MAPF_ERR_INVALID_ARG, "mapf_get_positions: NULL argument");
  CK((cudaError_t)mapf_launch_export16(h->d, h->S.pos, pos_dev, stream));
  h->launches++;
  return MAPF_OK;
}

int mapf_get_goals(mapf_handle* h, int16_t* goals_dev, void* stream) {
  if (!h || !goals_dev) return fail(h, MAPF_ERR_INVALID_ARG, "mapf_get_goals: NULL argument");
  CK((cudaError_t)mapf_launch_export16(h->d, h->S.goal, goals_dev, stream));
  h->launches++;
  return MAPF_OK;
}

int mapf_get_dones(mapf_handle* h, uint8_t* dones_dev, void* stream) {
  if (!h || !dones_dev) return fail(h, MAPF_ERR_INVALID_ARG, "mapf_get_dones: NULL argument");
  CK(cudaMemcpyAsync(dones_dev, h->S.done, (size_t)h->d.E * h->d.N, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  return MAPF_OK;
}

int mapf_get_step_count(mapf_handle* h, int32_t* step_count_dev, void* stream) {
  if (!h || !step_count_dev) return fail(h, MAPF_ERR_INVALID_ARG, "mapf_get_step_count: NULL argument");
  CK(cudaMemcpyAsync(step_count_dev, h->S.step_count, (size_t)h->d.E * 4, cudaMemcpyDeviceToDevice,
                     (cudaStream_t)stream));
  return MAPF_OK;
}

int mapf_partial_state(mapf_handle* h, int64_t* state_dev, uint8_t* at_goal_dev, int32_t* goal_cost_dev,
                       int32_t* agent_steps_dev, void* stream) {
  if (!h) return MAPF_ERR_INVALID_ARG;
  if (h->d.mode != MAPF_MODE_PARTIAL) return fail(h, MAPF_ERR_UNSUPPORTED, "mapf_partial_state: mode is not PARTIAL");
  const size_t EN = (size_t)h->d.E * h->d.N;
  cudaStream_t st = (cudaStream_t)stream;
  if (state_dev) {
    CK((cudaError_t)mapf_launch_partial_state(h->d, h->S, (long long*)state_dev, stream));
    h->launches++;
  }
  if (at_goal_dev) CK(cudaMemcpyAsync(at_goal_dev, h->S.at_goal, EN, cudaMemcpyDeviceToDevice, st));
  if (goal_cost_dev) CK(cudaMemcpyAsync(goal_cost_dev, h->S.goal_cost, EN * 4, cudaMemcpyDeviceToDevice, st));
  if (agent_steps_dev) CK(cudaMemcpyAsync(agent_steps_dev, h->S.agent_steps, EN * 4, cudaMemcpyDeviceToDevice, st));
  return MAPF_OK;
}

int mapf_random_actions(mapf_handle* h, const uint8_t* avail_dev, uint32_t seed, uint32_t step, int64_t env_offset,
                        void* actions_dev, int act_dtype, void* stream) {
  if (!h || !actions_dev) return fail(h, MAPF_ERR_INVALID_ARG, "mapf_random_actions: NULL argument");
  if (act_dtype != MAPF_U8 && act_dtype != MAPF_I64) return fail(h, MAPF_ERR_INVALID_ARG, "actions must be MAPF_U8 or MAPF_I64");
  CK((cudaError_t)mapf_launch_random_actions(h->d, avail_dev, seed, step, (long long)env_offset, actions_dev,
                                             act_dtype == MAPF_I64, stream));
  h->launches++;
  return MAPF_OK;
}

int mapf_runner_mask_actions(mapf_handle* h, const void* actions_dev, int act_dtype, const uint8_t* alive_dev,
                             int stay_action, uint8_t* actions_u8_dev, int64_t* actions_i64_dev, void* stream) {
  if (!h || !actions_dev || !alive_dev || !actions_u8_dev)
    return fail(h, MAPF_ERR_INVALID_ARG, "mapf_runner_mask_actions: NULL argument");
  if (act_dtype != MAPF_U8 && act_dtype != MAPF_I64) return fail(h, MAPF_ERR_INVALID_ARG, "actions must be MAPF_U8 or MAPF_I64");
  CK((cudaError_t)mapf_launch_runner_mask_actions(h->d, actions_dev, act_dtype == MAPF_I64, alive_dev, stay_action,
                                                  actions_u8_dev, (long long*)actions_i64_dev, stream));
  h->launches++;
  return MAPF_OK;
}

int mapf_runner_account(mapf_handle* h, const double* reward_dev, const uint8_t* terminated_dev, uint8_t* alive_dev,
                        double* returns_dev, int64_t* lengths_dev, uint8_t* filled_next_dev, void* stream) {
  if (!h || !reward_dev || !terminated_dev || !alive_dev || !returns_dev || !lengths_dev || !filled_next_dev)
    return fail(h, MAPF_ERR_INVALID_ARG, "mapf_runner_account: NULL argument");
  CK((cudaError_t)mapf_launch_runner_account(h->d, reward_dev, terminated_dev, alive_dev, returns_dev,
                                             (long long*)lengths_dev, filled_next_dev, stream));
  h->launches++;
  return MAPF_OK;
}

// Self-test hook (not part of the public header): the next step / rollout launch of this handle overwrites guard word
// `which` of its tile on purpose, so that tests can see MAPF_FLAG_INTERNAL being raised by the canary check.
int mapf_debug_corrupt_canary(mapf_handle* h, int which) {
  if (!h || which < 0) return MAPF_ERR_INVALID_ARG;
  h->debug_corrupt = which + 1;
  return MAPF_OK;
}

int mapf_partial_bind_state_out(mapf_handle* h, int64_t* state_dev) {
  if (!h) return MAPF_ERR_INVALID_ARG;
  if (h->d.mode != MAPF_MODE_PARTIAL) return fail(h, MAPF_ERR_UNSUPPORTED, "mapf_partial_bind_state_out: mode is not PARTIAL");
  h->partial_state_out = state_dev;
  return MAPF_OK;
}

int mapf_stats(mapf_handle* h, int64_t* stats_host, void* stream) {
  if (!h || !stats_host) return fail(h, MAPF_ERR_INVALID_ARG, "mapf_stats: NULL argument");
  CK(cudaMemcpyAsync(stats_host, h->S.stats, MAPF_N_STATS * 8, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  CK(cudaStreamSynchronize((cudaStream_t)stream));
  return MAPF_OK;
}

int mapf_error_flags(mapf_handle* h, uint32_t* flags_host, void* stream) {
  if (!h || !flags_host) return fail(h, MAPF_ERR_INVALID_ARG, "mapf_error_flags: NULL argument");
  CK(cudaMemcpyAsync(flags_host, h->S.err_flags, 4, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  CK(cudaMemsetAsync(h->S.err_flags, 0, 4, (cudaStream_t)stream));
  CK(cudaStreamSynchronize((cudaStream_t)stream));
  return MAPF_OK;
}

int mapf_obs_bits_supported(const mapf_handle* h) { return (h && bits_supported(h)) ? 1 : 0; }

int mapf_host_transport(mapf_handle* h, int packed) {
  if (!h) return 0;
  h->packed_transport = packed ? 1 : 0;
  return (h->packed_transport && bits_supported(h)) ? 1 : 0;
}

int mapf_host_unpack(const void* bits_host, uint64_t first_cell, uint64_t n_cells, void* out_host, int out_dtype) {
  if (out_dtype != MAPF_U8 && out_dtype != MAPF_F32) return MAPF_ERR_INVALID_ARG;
  return mapf_host_unpack_range((const uint32_t*)bits_host, first_cell, n_cells, out_host, out_dtype == MAPF_F32 ? 4 : 1) == 0
             ? MAPF_OK
             : MAPF_ERR_INVALID_ARG;
}

int mapf_host_transport_get(const mapf_handle* h) { return (h && h->packed_transport && bits_supported(h)) ? 1 : 0; }

int mapf_step_observe_host(mapf_handle* h, const mapf_host_io* io, void* stream) {
  if (!h || !io || !io->actions_host) return fail(h, MAPF_ERR_INVALID_ARG, "mapf_step_observe_host: NULL argument");
  // MAPF_HOST_TRACE=1: host-side timestamps (microseconds since entry) of this call on stderr, for profiles/
  static const bool trace = getenv("MAPF_HOST_TRACE") != nullptr;
  const auto t_entry = std::chrono::steady_clock::now();
  auto us = [&]() { return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t_entry).count(); };
  double t_queued = 0, t_unpacked = 0;
  cudaStream_t st = (cudaStream_t)stream;
  const MapfDims& d = h->d;
  const size_t EN = (size_t)d.E * d.N;
  const bool fov = d.obs_mode == MAPF_OBS_PRIMAL_FOV;
  // uint8 field-of-view observations cross PCIe as bits and are expanded by the host pool (byte-identical result)
  const bool packed = io->obs_host && fov && (io->obs_dtype == MAPF_U8 || io->obs_dtype == MAPF_F32) &&
                      h->packed_transport && bits_supported(h);
  // MAPF_BITS host output: the bit stream itself goes to the caller's buffer (ceil(cells / 32) words), no expansion
  const bool bits_out = io->obs_host && fov && io->obs_dtype == MAPF_BITS;
  if (bits_out && !bits_supported(h))
    return fail(h, MAPF_ERR_UNSUPPORTED, "mapf_step_observe_host: MAPF_BITS needs a specialised FOV kernel and whole "
                "observation groups per tile (mapf_obs_bits_supported)");
  const int elem = io->obs_dtype == MAPF_F32 ? 4 : 1;
  size_t obs_bytes = 0;
  if (io->obs_host) {
    if (bits_out) obs_bytes = (EN * 4 * d.F * d.F + 31) / 32 * 4;
    else if (fov) obs_bytes = EN * 4 * d.F * d.F * (io->obs_dtype == MAPF_F32 ? 4 : 1);
    else if (d.obs_mode == MAPF_OBS_PARTIAL_WINDOW) obs_bytes = EN * d.posz * (io->obs_dtype == MAPF_F32 ? 4 : 8);
    else obs_bytes = (size_t)d.E * d.HW;
  }
#define LAZY(ptr, bytes)                                                \
  if (!(ptr)) {                                                         \
    cudaError_t _e = cudaMalloc((void**)&(ptr), (bytes));               \
    if (_e != cudaSuccess) return cuda_fail(h, _e, "cudaMalloc " #ptr); \
  }
  LAZY(h->hs_actions, EN);
  if (io->reward_host) LAZY(h->hs_reward, (size_t)d.E * 8);
  if (io->terminated_host) LAZY(h->hs_terminated, (size_t)d.E);
  if (io->dones_host) LAZY(h->hs_dones, EN);
  if (io->avail_host) LAZY(h->hs_avail, EN * d.nact);
  if (io->vec_host) LAZY(h->hs_vec, EN * 24);
  const size_t tile_words = (packed || bits_out) ? ((size_t)d.epb * d.N * 4 * d.F * d.F) >> 5 : 0;
  const size_t ntiles = ((size_t)d.E + d.epb - 1) / d.epb;
  if (bits_out) {
    h->bits_words = ntiles * tile_words;
    LAZY(h->hs_bits, h->bits_words * 4);
  } else if (packed) {
    h->bits_words = ntiles * tile_words;
    LAZY(h->hs_bits, h->bits_words * 4);
    if (!h->hp_bits) CK(cudaHostAlloc((void**)&h->hp_bits, h->bits_words * 4, cudaHostAllocDefault));
    if (!h->pool) {
      h->pool = mapf_unpack_pool_create(0);
      if (!h->pool) return fail(h, MAPF_ERR_ALLOC, "mapf_step_observe_host: cannot start the host unpack threads");
    }
    if (!h->hp_prog) CK(cudaHostAlloc((void**)&h->hp_prog, 64, cudaHostAllocDefault));
    if (!h->hs_prog) {
      uint32_t prog[MAPF_HOST_CHUNKS];
      const size_t per = (ntiles + MAPF_HOST_CHUNKS - 1) / MAPF_HOST_CHUNKS;
      for (int c = 0; c < MAPF_HOST_CHUNKS; ++c) {
        const size_t t1 = (c + 1) * per < ntiles ? (c + 1) * per : ntiles;
        prog[c] = (uint32_t)(t1 * tile_words);
      }
      LAZY(h->hs_prog, sizeof(prog));
      CK(cudaMemcpyAsync(h->hs_prog, prog, sizeof(prog), cudaMemcpyHostToDevice, st));   // pageable source: staged before return
    }
  } else if (!bits_out && obs_bytes > h->hs_obs_bytes) {
    cudaFree(h->hs_obs);
    h->hs_obs = nullptr;
    h->hs_obs_bytes = 0;
    CK(cudaMalloc(&h->hs_obs, obs_bytes));
    h->hs_obs_bytes = obs_bytes;
  }
#undef LAZY
  CK(cudaMemcpyAsync(h->hs_actions, io->actions_host, EN, cudaMemcpyHostToDevice, st));
  mapf_step_out out;
  memset(&out, 0, sizeof(out));
  out.reward_dev = io->reward_host ? h->hs_reward : nullptr;
  out.terminated_dev = io->terminated_host ? h->hs_terminated : nullptr;
  out.dones_dev = io->dones_host ? h->hs_dones : nullptr;
  out.avail_dev = io->avail_host ? h->hs_avail : nullptr;
  void* obs_dev = !io->obs_host ? nullptr : ((packed || bits_out) ? (void*)h->hs_bits : h->hs_obs);
  int rc = run_tile(h, h->hs_actions, MAPF_U8, 0, d.N, &out, obs_dev, (packed || bits_out) ? (int)MAPF_BITS : io->obs_dtype,
                    io->vec_host ? h->hs_vec : nullptr, stream);
  if (rc != MAPF_OK) return rc;
  auto copy_small_outputs = [&]() -> int {
    if (io->reward_host) CK(cudaMemcpyAsync(io->reward_host, h->hs_reward, (size_t)d.E * 8, cudaMemcpyDeviceToHost, st));
    if (io->terminated_host)
      CK(cudaMemcpyAsync(io->terminated_host, h->hs_terminated, (size_t)d.E, cudaMemcpyDeviceToHost, st));
    if (io->dones_host) CK(cudaMemcpyAsync(io->dones_host, h->hs_dones, EN, cudaMemcpyDeviceToHost, st));
    if (io->avail_host) CK(cudaMemcpyAsync(io->avail_host, h->hs_avail, EN * d.nact, cudaMemcpyDeviceToHost, st));
    if (io->vec_host) CK(cudaMemcpyAsync(io->vec_host, h->hs_vec, EN * 24, cudaMemcpyDeviceToHost, st));
    return MAPF_OK;
  };
  if (packed) {
    // Chunks of whole tiles, each followed on the stream by a 4-byte copy of "words in host memory so far" into a
    // pinned word.  The pool (one job for the whole call, blocks claimed in order, the calling thread works too)
    // expands chunk c while chunk c+1 and then the small outputs, queued last, are still crossing PCIe.
    const size_t per = (ntiles + MAPF_HOST_CHUNKS - 1) / MAPF_HOST_CHUNKS;
    const size_t ncells = obs_bytes / elem;
    *h->hp_prog = 0;                                  // the stream has nothing of this handle's host path in flight
    int c = 0;
    for (size_t t0 = 0; t0 < ntiles; t0 += per, ++c) {
      const size_t t1 = t0 + per < ntiles ? t0 + per : ntiles;
      CK(cudaMemcpyAsync(h->hp_bits + t0 * tile_words, h->hs_bits + t0 * tile_words, (t1 - t0) * tile_words * 4,
                         cudaMemcpyDeviceToHost, st));
      CK(cudaMemcpyAsync((void*)h->hp_prog, h->hs_prog + c, 4, cudaMemcpyDeviceToHost, st));
    }
    if ((rc = copy_small_outputs()) != MAPF_OK) return rc;
    t_queued = us();
    struct PollCtx { cudaStream_t st; cudaError_t err; } pc = {st, cudaSuccess};
    auto poll = [](void* a) -> int {                  // a stream that failed will never publish the words
      PollCtx* p = (PollCtx*)a;
      const cudaError_t e = cudaStreamQuery(p->st);
      if (e == cudaSuccess || e == cudaErrorNotReady) return 0;
      p->err = e;
      return 1;
    };
    const int ok = mapf_unpack_pool_expand_streamed(h->pool, h->hp_bits, io->obs_host, ncells, elem, h->hp_prog, poll, &pc);
    t_unpacked = us();
    if (!ok) return cuda_fail(h, pc.err, "mapf_step_observe_host: device-to-host transfer");
  } else {
    if ((rc = copy_small_outputs()) != MAPF_OK) return rc;
    if (io->obs_host)
      CK(cudaMemcpyAsync(io->obs_host, bits_out ? (void*)h->hs_bits : h->hs_obs, obs_bytes, cudaMemcpyDeviceToHost, st));
  }
  CK(cudaStreamSynchronize(st));
  if (trace)
    fprintf(stderr, "mapf_step_observe_host trace (us): queued %.0f, expanded %.0f, stream idle %.0f\n", t_queued,
            t_unpacked, us());
  return MAPF_OK;
}

}  // extern "C"
