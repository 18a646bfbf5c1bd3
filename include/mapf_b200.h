/*
 * mapf_b200.h -- C ABI of libmapf_b200.so: the batched MAPF grid-world step/observation
 * engine for NVIDIA B200 (sm_100a).
 *
 * The reference (DongmingShen/MAPF-MARL) has NO foreign-function interface for this path:
 * its environments are pure-Python classes.  Each entry point below therefore cites the
 * reference *method* it replaces (paths relative to the reference root; GRID =
 * mapf_gridworld.py, PRIMAL = mapf_primal.py, PARTIAL = MARL-curve-main/src/envs/marl_partial.py).
 * INTEGRATION.md shows the ctypes binding a maintainer of the reference would add.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes, no C++ / torch types, no exceptions.
 *   - Every function returns 0 (MAPF_OK) or a negative mapf_status; the message is
 *     available from mapf_last_error().
 *   - Pointers named *_dev are DEVICE pointers owned by the caller (e.g. torch tensors),
 *     pointers named *_host are HOST pointers.  The handle owns the environment state
 *     (positions, goals, done flags, bitmaps, distance maps) in device memory.
 *   - Calls are stream-ordered on `stream` (a cudaStream_t passed as void*, NULL = the
 *     legacy default stream) and never synchronise, except the *_host entry points and
 *     mapf_error_flags / mapf_stats which return host data and therefore wait for the stream.
 *   - Not thread-safe per handle; one handle per (process, device); the caller has made
 *     the device current (cudaSetDevice / torch.cuda.set_device) before mapf_create.
 *   - Arrays are row-major and dense.  E = n_envs, N = n_agents, H = height, W = width,
 *     F = fov.  A position is (p0, p1): p0 indexes the FIRST axis of the map, exactly as
 *     `_full_obs[pos[0]][pos[1]]` (GRID:299) and `state[x, y]` (PRIMAL:117) do.
 */
#ifndef MAPF_B200_H
#define MAPF_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MAPF_ABI_VERSION 1

typedef enum mapf_status {
  MAPF_OK = 0,
  MAPF_ERR_INVALID_ARG = -1,
  MAPF_ERR_CUDA = -2,
  MAPF_ERR_UNSUPPORTED = -3,
  MAPF_ERR_ALLOC = -4
} mapf_status;

/* Collision / reward semantics (SURVEY.md section 8 preamble). */
typedef enum mapf_mode {
  /* mode A "detect": MAPF_GRID.step, GRID:85-141.  Actions {0:(-1,0) 1:(+1,0) 2:(0,-1)
   * 3:(0,+1) 4:stay} (GRID:319-342); agents may overlap; node/edge collisions are
   * counted and penalised, never resolved (GRID:344-383). */
  MAPF_MODE_GRID = 0,
  /* mode B "sequential claim": State.moveAgent swept over ids 1..N, PRIMAL:103-135 and
   * MAPFEnv._step, PRIMAL:549-637.  Actions {0:stay 1:(0,+1) 2:(+1,0) 3:(0,-1) 4:(-1,0)}
   * (dirDict, PRIMAL:28). */
  MAPF_MODE_PRIMAL = 1,
  /* mode A with the rewards and bookkeeping of MARL_PARTIAL_ENV.step (output == False), PARTIAL:169-310: move /
   * stay / stay-on-goal costs, goal-distance shaping `(dist[old] - dist[new]) / episode_limit` from the BFS maps
   * (:229-234), at_goal recomputed every step, termination at the limit or when all agents are on goal, with the
   * completion bonus (:291-299).  The same GRID action deltas on (row, col). */
  MAPF_MODE_PARTIAL = 2
} mapf_mode;

typedef enum mapf_obs_mode {
  MAPF_OBS_FULLMAP = 0,   /* GRID get_obs/get_state, GRID:143-196: int8[E, H*W], -1 wall else agent count */
  MAPF_OBS_PRIMAL_FOV = 1, /* PRIMAL _observe, PRIMAL:343-386: [E, N, 4, F, F] + double[E, N, 3] */
  /* PARTIAL get_obs_agent, PARTIAL:319-382: double[E, N, 2*W*W + 13*K]: W x W obstacle map, W x W agent-count map,
   * K nearest agents x 13 features (self first, stable sort by L2 distance, missing rows -1). */
  MAPF_OBS_PARTIAL_WINDOW = 2
} mapf_obs_mode;

typedef enum mapf_dtype {
  MAPF_U8 = 0,
  MAPF_I64 = 1,
  MAPF_F32 = 2,
  MAPF_I8 = 3,
  MAPF_F64 = 4,
  /* FOV observations only: the 0/1 cells packed one bit per cell, bit i of the stream == byte i of the MAPF_U8
   * tensor [E,N,4,F,F] (little-endian bit order inside 32-bit words), ceil(E*N*4*F*F / 32) words.  8x fewer bytes
   * for a consumer that takes bits (and the transport format of mapf_step_observe_host).  Needs a specialised
   * field-of-view kernel and whole observation groups per tile: mapf_obs_bits_supported(). */
  MAPF_BITS = 5
} mapf_dtype;

/* Device-side error flag bits (mapf_error_flags). */
#define MAPF_FLAG_BAD_ACTION 1u   /* action outside {0..4}: GRID:92 / PRIMAL:556 assert */
#define MAPF_FLAG_BAD_POSITION 2u /* start/goal out of bounds at reset */
#define MAPF_FLAG_START_ON_WALL 4u
#define MAPF_FLAG_START_OVERLAP 8u /* PRIMAL only: two agents on one cell (State.scanForAgents, PRIMAL:53-66) */
#define MAPF_FLAG_GOAL_OVERLAP 16u /* PRIMAL only: two agents share a goal cell (State.goals holds one id per cell) */
/* Internal consistency check: a guard word between two shared-memory regions of a kernel was overwritten (the kernels
 * plant canaries around their tile regions and verify them before they exit; compute-sanitizer is not available on every
 * pool, this check always runs).  Never expected; report it as a bug. */
#define MAPF_FLAG_INTERNAL 32u

/* Indices into the int64[MAPF_N_STATS] vector returned by mapf_stats. */
enum {
  MAPF_STAT_ENV_STEPS = 0,    /* environment steps executed */
  MAPF_STAT_AGENT_STEPS = 1,  /* agent-steps executed (= env steps * N, or swept agents) */
  MAPF_STAT_ENV_COLLISIONS = 2,   /* GRID: wall/border bumps (GRID:105-107); PRIMAL: status -1/-2 */
  MAPF_STAT_NODE_COLLISIONS = 3,  /* GRID: sum of node flags (GRID:344-362); PRIMAL: status -3 */
  MAPF_STAT_EDGE_COLLISIONS = 4,  /* GRID: sum of edge counts (GRID:364-383) */
  MAPF_STAT_GOAL_ARRIVALS = 5,    /* agents that newly reached their goal */
  MAPF_STAT_EPISODES_DONE = 6,    /* env steps that ended with terminated == 1 */
  MAPF_STAT_RESERVED = 7,
  MAPF_N_STATS = 8
};

/* Constructor arguments.  Mirrors the keyword arguments of MAPF_GRID.__init__ (GRID:21-32)
 * and MAPFEnv.__init__ (PRIMAL:175-176) plus the batch dimensions. */
typedef struct mapf_cfg {
  int32_t abi_version;   /* MAPF_ABI_VERSION */
  int32_t n_envs;        /* E >= 1 */
  int32_t n_agents;      /* N in [1, 255] */
  int32_t height;        /* H in [1, 255] */
  int32_t width;         /* W in [1, 255] */
  int32_t mode;          /* mapf_mode */
  int32_t obs_mode;      /* mapf_obs_mode */
  int32_t fov;           /* F = observation_size (PRIMAL:188); ignored for MAPF_OBS_FULLMAP */
  int32_t shared_map;    /* 1: one map [H, W] shared by all envs; 0: int8[E, H, W] */
  int32_t episode_limit; /* GRID:26, 116 */
  int32_t goal_dist;     /* 1: keep int16[E, N, H, W] BFS distance maps in the handle */
  int32_t collect_stats; /* 1: accumulate the mapf_stats counters */
  double step_reward;    /* GRID:29 */
  double collide_reward; /* GRID:30 */
  /* PRIMAL reward table, PRIMAL:25 (ACTION_COST, IDLE_COST, GOAL_REWARD, COLLISION_REWARD) */
  double action_cost;
  double idle_cost;
  double goal_reward;
  double collision_reward;
  /* mag_lut_host[s] = |(dx,dy)| for s = dx*dx + dy*dy, s in [0, mag_lut_len).  Built by the
   * host with the reference's own expression so that the goal vector is bit-identical:
   * `(dx**2 + dy**2) ** .5` (PRIMAL:382).  Must cover s <= (H-1)^2 + (W-1)^2. */
  const double* mag_lut_host;
  int32_t mag_lut_len;
  /* How `sum(rewards)` (GRID:141) is folded, i.e. which interpreter the drop-in imitates:
   *   0 = plain left fold (CPython <= 3.11 builtin sum),
   *   1 = CPython >= 3.12 builtin sum (exact ints, then Neumaier-compensated float adds).
   * The two *_is_int flags say whether the keyword arguments were Python ints: sum() treats int and
   * float items differently, so the type of every rewards[i] is part of the result. */
  int32_t reward_sum_mode;
  int32_t step_reward_is_int;
  int32_t collide_reward_is_int;
  int32_t reserved;
  /* MAPF_MODE_PARTIAL: constructor keywords of MARL_PARTIAL_ENV, PARTIAL:26-45. */
  int32_t obs_window;         /* W */
  int32_t obs_knn_agents;     /* K */
  double move_reward, stay_reward, stay_goal_reward;
  double node_collide_reward, edge_collide_reward, env_collide_reward;
  /* complete_lut_host[t] = the bonus every agent receives when all agents are on goal at step t:
   * `(complete_reward / (gamma ** (episode_limit - t))) * complete_fac` (PARTIAL:296), evaluated by the host. */
  const double* complete_lut_host;
  int32_t complete_lut_len;
  /* MAPF_MODE_PRIMAL: 1 = add get_blocking_reward (PRIMAL:513-546) to the reward of an agent that stays on its goal:
   * -1 (blocking_cost) for every visible robot (ids 1..N-1, the reference's loop skips the last id) whose
   * single-robot shortest path to its goal is cut, or lengthened by more than 10, by this agent.  The reference
   * gets those path lengths from the un-vendored od_mstar3 (optimal single-robot paths); here they are BFS hop
   * counts.  Maps up to 64 x 64. */
  int32_t blocking_reward;
  double blocking_cost;      /* BLOCKING_COST, PRIMAL:25 (-1.0) */
  /* MAPF_MODE_PRIMAL: DIAGONAL_MOVEMENT=True (PRIMAL:175): 9 actions {5:(1,1) 6:(1,-1) 7:(-1,-1) 8:(-1,1)} (dirDict,
   * PRIMAL:28), State.diagonalCollision (PRIMAL:77-100) against every other agent's last recorded move, 9-wide action
   * masks ([E,N,9] for avail_dev / next_mid_dev) and 8-connected getAstarCosts (primal_costs in mapf_bfs). */
  int32_t diagonal_movement;
  int32_t reserved3;
} mapf_cfg;

/* Outputs of one step.  Every pointer is a device pointer and may be NULL (not written). */
typedef struct mapf_step_out {
  /* [E] team reward.  GRID / PARTIAL: `sum(rewards)` exactly as the interpreter folds it (GRID:141, PARTIAL:310).
   * PRIMAL: the PAIRWISE sum of the per-agent rewards of the swept agents (a convenience with no reference
   * counterpart -- MAPFEnv._step returns per-agent rewards only): leaves r[0..N) (+0.0 outside the swept range), padded
   * with +0.0 to the next power of two, y[i] += y[i + s] for s = 1, 2, 4, ...; a fixed order, so the bits do not
   * depend on tile shapes or GPU counts. */
  double* reward_dev;
  /* [E].  GRID: episode_done(), GRID:267.  PRIMAL: world.done() after the sweep, PRIMAL:159-165. */
  uint8_t* terminated_dev;
  /* [E, N] per-agent reward.  GRID: `rewards[i]` before the sum (GRID:94-130).  PRIMAL: `reward`
   * returned by _step (PRIMAL:579-597), blocking reward excluded. */
  double* agent_reward_dev;
  /* [E, N].  GRID: `_agent_dones` after the step (GRID:112-117).  PRIMAL: `on_goal` (PRIMAL:633). */
  uint8_t* dones_dev;
  /* [E, N].  PRIMAL: moveAgent status {2,1,0,-1,-2,-3} (PRIMAL:138-144).  GRID: env-collision flag. */
  int8_t* status_dev;
  /* [E, N] GRID `_node_collision_agents` (GRID:344-362); PRIMAL: 0. */
  int16_t* node_dev;
  /* [E, N] GRID `_edge_collision_agents` (GRID:364-383); PRIMAL: 0. */
  int16_t* edge_dev;
  /* [E, N] PRIMAL `valid_action` (PRIMAL:571); GRID: 1. */
  uint8_t* valid_dev;
  /* [E, N] PRIMAL `done` as returned by the i-th _step call of the sweep (mid-sweep, PRIMAL:626). */
  uint8_t* done_mid_dev;
  /* [E, N, A] PRIMAL `nextActions` as returned by the i-th _step call (mid-sweep, PRIMAL:630); A = 5, or 9 with
   * cfg.diagonal_movement. */
  uint8_t* next_mid_dev;
  /* [E, N, A] available-action mask after the step (A = 5, or 9 with cfg.diagonal_movement).
   * GRID: get_avail_actions (GRID:198-224).
   * PRIMAL: _listNextValidActions(i, action_i) evaluated after the whole sweep (PRIMAL:639-667). */
  uint8_t* avail_dev;
  /* [E, N] PRIMAL `blocking` as returned by _step (PRIMAL:578-585, 637); needs cfg.blocking_reward. */
  uint8_t* blocking_dev;
} mapf_step_out;

/* Host-buffer mirror of the outputs used by mapf_step_observe_host (pinned memory recommended). */
typedef struct mapf_host_io {
  const uint8_t* actions_host; /* [E, N] */
  double* reward_host;         /* [E] or NULL */
  uint8_t* terminated_host;    /* [E] or NULL */
  uint8_t* dones_host;         /* [E, N] or NULL */
  uint8_t* avail_host;         /* [E, N, 5] or NULL */
  void* obs_host;              /* obs in obs_dtype or NULL */
  double* vec_host;            /* [E, N, 3] or NULL */
  int32_t obs_dtype;           /* mapf_dtype */
  int32_t reserved;
} mapf_host_io;

typedef struct mapf_handle mapf_handle;

/* Error text of the last failing call on `h` (h == NULL: last failing mapf_create of this thread). */
const char* mapf_last_error(const mapf_handle* h);

/* Fills *cfg with the reference's defaults (GRID:25-31, PRIMAL:25, 175-176). */
void mapf_default_cfg(mapf_cfg* cfg);

/* Replaces MAPF_GRID.__init__ (GRID:21-68) / MAPFEnv.__init__ + State.__init__ (PRIMAL:175-203, 44-51):
 * allocates the device-resident state for E environments.  No map is loaded yet. */
int mapf_create(const mapf_cfg* cfg, mapf_handle** out);
int mapf_destroy(mapf_handle* h);

/* Replaces MAPF_GRID.reset (GRID:70-83) / MAPFEnv._reset + _setWorld with world0/goals0
 * (PRIMAL:389-402, 278-309).
 *   map_dev    int8 [E,H,W] (or [H,W] when shared_map): non-zero = obstacle.  NULL keeps the maps.
 *   starts_dev int16[E,N,2], goals_dev int16[E,N,2].  NULL keeps the stored starts / goals
 *              (GRID re-uses the positions sampled in __init__, GRID:79).
 *   env_mask_dev uint8[E] or NULL: only environments with a non-zero mask are reset. */
int mapf_reset(mapf_handle* h, const int8_t* map_dev, const int16_t* starts_dev, const int16_t* goals_dev,
               const uint8_t* env_mask_dev, void* stream);

/* Replaces the goal re-assignment of the lifelong variant (MAPF-490-main/Global.cpp:85-94 informs the
 * rule): overwrite the goals of agents whose dirty flag is set.  goals_dev int16[E,N,2], dirty_dev uint8[E,N]. */
int mapf_set_goals(mapf_handle* h, const int16_t* goals_dev, const uint8_t* dirty_dev, void* stream);

/* Replaces the lifelong task hand-out itself (MAPF-490-main/Global.cpp:85-94: when an agent has arrived, the front
 * of its deque becomes its goal and is popped; the deques are filled by generate_tasks, main.cpp:56-85).
 *   queue_dev int16[E,N,queue_len,2]  per-agent FIFO of (row, col) goals, entry 0 = first re-assignment
 *   head_dev  int32[E,N]              number of goals already popped (read and advanced here)
 *   dirty_dev uint8[E,N] or NULL      out: 1 where a goal was re-assigned (the mask mapf_bfs takes), else 0
 * An agent standing on its goal with head < queue_len receives queue[head]; everything else is untouched. */
int mapf_pop_goals(mapf_handle* h, const int16_t* queue_dev, int32_t* head_dev, int queue_len, uint8_t* dirty_dev,
                   void* stream);

/* The same hand-out fused into the step: binds the queues to the handle, and from then on every full-range PRIMAL step
 * launch (mapf_step, mapf_step_observe, mapf_step_observe_host, the launches of mapf_rollout) pops the queue of every
 * agent that ENDS the step on its goal, in the step kernel's own write-back -- bit for bit the state that
 * mapf_step* followed by mapf_pop_goals leaves (the step's outputs and observation still show the old goal), without
 * the extra launch and without the dirty mask: the re-assigned (env, agent) pairs are collected in a list that
 * mapf_bfs_popped hands to the BFS.  queue_dev / head_dev as in mapf_pop_goals, owned by the caller and read at every
 * step; queue_dev == NULL unbinds.  PRIMAL mode only; multi-step single-launch rollouts fall back to one launch per
 * step while queues are bound. */
int mapf_lifelong_bind(mapf_handle* h, const int16_t* queue_dev, int32_t* head_dev, int queue_len);

/* Goal-distance maps (mapf_bfs) of the agents whose goals the MOST RECENT step launch re-assigned (queues bound with
 * mapf_lifelong_bind).  dist_dev as in mapf_bfs (NULL: the handle's own maps).  May run on another stream than the
 * step, concurrently with the NEXT step: the handle keeps two lists and alternates between them, so the caller only
 * has to order this call behind the step it belongs to and ahead of the step after the next one. */
int mapf_bfs_popped(mapf_handle* h, int16_t* dist_dev, void* stream);

/* Replaces MAPF_GRID.step (GRID:85-141) or one full sweep `for id in 1..N: MAPFEnv._step((id, a[id]))`
 * (PRIMAL:549-637).  actions_dev: [E,N] of act_dtype (MAPF_U8 or MAPF_I64). */
int mapf_step(mapf_handle* h, const void* actions_dev, int act_dtype, const mapf_step_out* out, void* stream);

/* PRIMAL only: sweep agents agent_lo <= i < agent_hi (0-based).  [i, i+1) is exactly one
 * MAPFEnv._step((i+1, a)) call (PRIMAL:549). */
int mapf_step_agents(mapf_handle* h, const void* actions_dev, int act_dtype, int agent_lo, int agent_hi,
                     const mapf_step_out* out, void* stream);

/* Replaces get_obs/get_state (GRID:143-196) or `_observe(id)` for every id (PRIMAL:343-386).
 *   MAPF_OBS_FULLMAP:    obs_dev int8[E, H*W] (MAPF_I8); vec_dev ignored.
 *   MAPF_OBS_PRIMAL_FOV: obs_dev [E,N,4,F,F] of obs_dtype (MAPF_U8 or MAPF_F32; MAPF_BITS: the same cells as a
 *                        bit stream), channel order [poss_map, goal_map, goals_map, obs_map] (PRIMAL:386);
 *                        vec_dev double[E,N,3] = [dx/mag, dy/mag, mag] or NULL.
 *   MAPF_OBS_PARTIAL_WINDOW (MARL_PARTIAL_ENV.get_obs, PARTIAL:312-382): obs_dev [E,N,2*W*W + 13*K] of MAPF_F64 (the
 *                        reference's dtype) or MAPF_F32 (the same values rounded once at the store: what pymarl's
 *                        episode batch keeps, src/run.py:133-140); vec_dev ignored. */
int mapf_observe(mapf_handle* h, void* obs_dev, int obs_dtype, double* vec_dev, void* stream);

/* mapf_step followed by mapf_observe in ONE kernel launch (state is staged in shared memory once). */
int mapf_step_observe(mapf_handle* h, const void* actions_dev, int act_dtype, const mapf_step_out* out,
                      void* obs_dev, int obs_dtype, double* vec_dev, void* stream);

/* n_steps consecutive mapf_step_observe calls with pre-supplied actions in ONE call -- and, for PRIMAL (without
 * diagonal movement / blocking reward) and GRID batches whose tiles hold one thread per agent, in ONE kernel launch:
 * the tile's state stays in shared memory and registers between the steps, only the per-step outputs stream out;
 * small PRIMAL batches go through a pipelined kernel (mapf_rollout_plan)
 * (the reference's equivalent is the env loop of its runner,
 * MARL-curve-main/src/runners/parallel_runner.py:127-171, driven by a fixed action sequence).
 *   actions_dev [n_steps, E, N] of act_dtype.
 *   Every non-NULL pointer of *out, obs_dev and vec_dev is TIME-MAJOR: [n_steps, <the shape mapf_step_observe
 *   documents>]; step t's outputs are exactly what the t-th of n_steps consecutive mapf_step_observe calls writes.
 * One step's actions and one step's observation must be multiples of 16 bytes. */
int mapf_rollout(mapf_handle* h, const void* actions_dev, int act_dtype, int n_steps, const mapf_step_out* out,
                 void* obs_dev, int obs_dtype, double* vec_dev, void* stream);
/* 1 when mapf_rollout runs as a single launch for this handle and observation dtype (otherwise n_steps launches). */
int mapf_rollout_in_one_launch(const mapf_handle* h, int obs_dtype);
/* Which kernel a rollout with an observation of obs_dtype would use: 0 = n_steps launches of the step kernel, 1 = the
 * step kernel's in-kernel loop, 2 = the pipelined kernel for small PRIMAL batches (at most 32 agents per environment,
 * every 32-agent tile resident at once: one warp sweeps step t+1 while three build the observation of step t).
 * mid_outputs != 0: the caller also wants done_mid / next_mid / blocking, which only the step kernel produces. */
int mapf_rollout_plan(const mapf_handle* h, int n_steps, int obs_dtype, int mid_outputs);

/* Host-buffer form of mapf_step_observe: copies io->actions_host to the device, runs the fused
 * kernel, copies the requested outputs back and waits for them.  This is the call the e2e
 * benchmark times.
 * MAPF_U8 / MAPF_F32 field-of-view observations cross PCIe as packed bits (MAPF_BITS, 8x / 32x fewer bytes) in
 * chunks and are expanded to the 0/1 cells of obs_host by the library's host threads while the next chunk is in flight; the
 * result is byte-identical to the dense copy.  mapf_host_transport(h, 0) switches back to the dense copy.
 * io->obs_dtype == MAPF_BITS hands the bit stream itself to the caller (obs_host: ceil(E*N*4*F*F / 32) uint32 words,
 * bit i of the stream = cell i of the [E,N,4,F,F] tensor, little-endian inside a word) with no host expansion.
 * The unpack pool uses MAPF_HOST_THREADS threads if that variable is set, else (CPUs of the process) / LOCAL_WORLD_SIZE. */
int mapf_step_observe_host(mapf_handle* h, const mapf_host_io* io, void* stream);

/* Host-side helper for consumers of the MAPF_BITS host output (io->obs_dtype == MAPF_BITS): expands cells
 * [first_cell, first_cell + n_cells) of the bit stream bits_host into out_host as MAPF_U8 (0/1) or MAPF_F32 (0.0/1.0)
 * on the calling thread -- a lazy view: a CPU consumer pays the 8x / 32x larger bytes only for the environments it
 * actually reads (environment e of an [E,N,4,F,F] observation is cells [e*N*4*F*F, (e+1)*N*4*F*F)).  Pure host code:
 * a transport decode of values the GPU computed, no handle and no CUDA call involved. */
int mapf_host_unpack(const void* bits_host, uint64_t first_cell, uint64_t n_cells, void* out_host, int out_dtype);

/* 1 when observations of this handle can be produced as MAPF_BITS. */
int mapf_obs_bits_supported(const mapf_handle* h);
/* packed != 0 (default): bit-packed PCIe transport in mapf_step_observe_host when supported; 0: dense copies.
 * Returns the mode now in effect (1 packed, 0 dense). */
int mapf_host_transport(mapf_handle* h, int packed);
/* The transport mode in effect (1 packed, 0 dense) without changing it. */
int mapf_host_transport_get(const mapf_handle* h);

/* Replaces get_avail_actions (GRID:198-224) / _listNextValidActions(id, prev_action) (PRIMAL:639-667)
 * for the current state.  avail_dev uint8[E,N,5]. */
int mapf_avail(mapf_handle* h, uint8_t* avail_dev, void* stream);

/* Replaces MARL_PARTIAL_ENV.__setup_agent_goal_dist (PARTIAL:931-955) and MAPFEnv.getAstarCosts
 * (PRIMAL:407-499): 4-connected hop distance from every cell to each agent's goal.
 *   dirty_dev uint8[E,N] or NULL (all).  dist_dev int16[E,N,H,W] or NULL (= the handle's own maps,
 *   requires cfg.goal_dist).  Walls = -1, unreachable free cells = -2.
 *   primal_costs != 0 reproduces getAstarCosts' quirk: unreachable cells keep `state` (0 or the id
 *   of the agent standing there, PRIMAL:496-498). */
int mapf_bfs(mapf_handle* h, const uint8_t* dirty_dev, int16_t* dist_dev, int primal_costs, void* stream);

/* mapf_avail with an explicit prev_action per agent (uint8[E,N], PRIMAL action ids) instead of the stored one: the
 * read-only query `_listNextValidActions(id, prev_action)` (PRIMAL:639-667) -- the handle's state, including the
 * stored previous actions, is left untouched. */
int mapf_avail_prev(mapf_handle* h, const uint8_t* prev_dev, uint8_t* avail_dev, void* stream);

/* Overrides the stored previous action of every agent (uint8[E,N], PRIMAL action ids).  mapf_avail removes
 * the opposite of this action, which is how `_listNextValidActions(id, prev_action)` (PRIMAL:639, 664) takes an
 * explicit prev_action. */
int mapf_set_prev_actions(mapf_handle* h, const uint8_t* prev_dev, void* stream);

/* State read-back (getPositions/getGoals, PRIMAL:236-246; agent_positions, GRID:61). int16[E,N,2]. */
int mapf_get_positions(mapf_handle* h, int16_t* pos_dev, void* stream);
int mapf_get_goals(mapf_handle* h, int16_t* goals_dev, void* stream);
/* uint8[E,N] `_agent_dones` (GRID) / on_goal (PRIMAL); int32[E] `_step_count` (GRID:93). */
int mapf_get_dones(mapf_handle* h, uint8_t* dones_dev, void* stream);
int mapf_get_step_count(mapf_handle* h, int32_t* step_count_dev, void* stream);

/* MAPF_MODE_PARTIAL read-backs.  state_dev int64[E,3] = get_state() (PARTIAL:384-393): [total collisions, step
 * count, sum of per-agent goal costs]; at_goal_dev uint8[E,N] (`_agent_at_goals`), goal_cost_dev int32[E,N]
 * (`_each_goal_cost`), agent_steps_dev int32[E,N] (`_agent_step_count`); any pointer may be NULL. */
int mapf_partial_state(mapf_handle* h, int64_t* state_dev, uint8_t* at_goal_dev, int32_t* goal_cost_dev,
                       int32_t* agent_steps_dev, void* stream);

/* A random policy for rollouts and benchmarks, drawn on the device (the reference's runners sample random actions
 * from the action mask on the host, e.g. MARL-curve-main/src/envs/marl_partial.py's __main__ loop): the action of
 * (env, agent) is a counter hash of (seed, env_offset + env, step, agent), uniform over the n_actions actions, or --
 * with avail_dev uint8[E,N,A] -- uniform over the agent's available actions.  A pure function of the GLOBAL env index
 * env_offset + env, so the shards of a batch draw exactly what the unsharded batch would.  actions_dev [E,N] of
 * act_dtype (MAPF_U8 or MAPF_I64). */
int mapf_random_actions(mapf_handle* h, const uint8_t* avail_dev, uint32_t seed, uint32_t step, int64_t env_offset,
                        void* actions_dev, int act_dtype, void* stream);

/* The bookkeeping of a rollout loop around one vector-env step, replacing what pymarl's ParallelRunner.run does on the
 * host per environment (MARL-curve-main/src/runners/parallel_runner.py:123-175: actions only to running envs;
 * episode_returns / episode_lengths / terminated bookkeeping; the `filled` mask of EpisodeBatch.update,
 * components/episode_buffer.py:100-134).
 * mapf_runner_mask_actions: actions_dev [E,N] (act_dtype MAPF_U8 or MAPF_I64) -> actions_u8_dev uint8[E,N] (the step's
 *   input) and, when not NULL, actions_i64_dev int64[E,N] (the episode batch's storage); environments with
 *   alive_dev[e] == 0 (uint8[E]) get stay_action.
 * mapf_runner_account: for every environment with alive != 0: returns += reward (float64[E], round-to-nearest add),
 *   lengths += 1 (int64[E]); filled_next_dev[e] = alive (uint8[E]: the NEXT time slot holds data); then
 *   alive &= (terminated == 0). */
int mapf_runner_mask_actions(mapf_handle* h, const void* actions_dev, int act_dtype, const uint8_t* alive_dev,
                             int stay_action, uint8_t* actions_u8_dev, int64_t* actions_i64_dev, void* stream);
int mapf_runner_account(mapf_handle* h, const double* reward_dev, const uint8_t* terminated_dev, uint8_t* alive_dev,
                        double* returns_dev, int64_t* lengths_dev, uint8_t* filled_next_dev, void* stream);

/* MAPF_MODE_PARTIAL: from now on every observation launch of this handle (mapf_observe, mapf_step_observe,
 * mapf_rollout) also writes get_state() (PARTIAL:384-393) to state_dev int64[E,3] -- the state rides along with the
 * observation kernel instead of costing a launch of its own per environment step.  NULL unbinds.  The pointer is
 * re-read at every launch; rebinding between launches (e.g. to the next time slice of an episode batch) is free. */
int mapf_partial_bind_state_out(mapf_handle* h, int64_t* state_dev);

/* Copies the int64[MAPF_N_STATS] counters to stats_host (waits for the stream). */
int mapf_stats(mapf_handle* h, int64_t* stats_host, void* stream);
/* Reads and clears the device error-flag word (waits for the stream). */
int mapf_error_flags(mapf_handle* h, uint32_t* flags_host, void* stream);

/* Number of kernels this handle has launched so far (bench.py reports it as gpu_launches). */
int64_t mapf_launch_count(const mapf_handle* h);
/* Library identification: ABI version and the -gencode it was built for ("sm_100a"). */
int mapf_abi_version(void);
const char* mapf_build_arch(void);

#ifdef __cplusplus
}
#endif
#endif /* MAPF_B200_H */
