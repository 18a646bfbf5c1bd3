/*
 * mapf_oracle.c -- CPU restatement of the reference's step/observation path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it.  The product path
 * (mapf_marl_b200/) never calls into this file and has no CPU fallback.
 *
 * Parity status: PINNED.  Every function below is checked bit-for-bit against traces
 * recorded from the live, unmodified reference classes (the .npz files under tests/golden/, produced by
 * tests/golden/gen_golden.py) in tests/test_oracle_golden.py.
 *
 * The reference is pure Python (no compilable sources), so this file restates its
 * algorithms in plain C, one environment at a time, parallelised over environments with
 * OpenMP exactly like the reference's ParallelRunner runs one env per worker process
 * (MARL-curve-main/src/runners/parallel_runner.py:23-31).  Citations: GRID =
 * mapf_gridworld.py, PRIMAL = mapf_primal.py, PARTIAL = MARL-curve-main/src/envs/marl_partial.py.
 *
 * Floating point: compiled with -ffp-contract=off so `r += c * k` is a rounded multiply
 * followed by a rounded add, as in CPython.  The goal-vector magnitude uses libm pow(s, .5),
 * the same call CPython makes for `(dx**2 + dy**2) ** .5` (PRIMAL:382).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct oracle_env {
  int E, N, H, W, F;
  int shared_map;
  int episode_limit;
  double step_reward, collide_reward;                              /* GRID:29-30 */
  int sum_mode;          /* 0: left fold (CPython <= 3.11 sum); 1: CPython >= 3.12 sum (Neumaier) */
  int step_is_int;       /* the step_reward kwarg was a Python int */
  int collide_is_int;    /* the collide_reward kwarg was a Python int (the GRID default -10 is) */
  double action_cost, idle_cost, goal_reward, collision_reward;    /* PRIMAL:25 */
  int8_t* map;       /* [Emap,H,W] non-zero = obstacle */
  int16_t* state;    /* [E,H,W] GRID: _full_obs (-1 wall, else agent count) GRID:57,132-135
                               PRIMAL: State.state (-1 wall, 0 free, id) PRIMAL:32-47 */
  int16_t* goals;    /* [E,H,W] PRIMAL State.goals (id at goal cell) */
  int16_t* pos;      /* [E,N,2] */
  int16_t* goal;     /* [E,N,2] */
  int16_t* start;    /* [E,N,2] */
  uint8_t* done;     /* [E,N]  GRID _agent_dones */
  int32_t* step_count; /* [E] */
  int threads;
  int diagonal;          /* PRIMAL DIAGONAL_MOVEMENT (PRIMAL:175): 9 actions, diagonalCollision, 8-connected costs */
  int16_t* past;         /* [E,N,2] State.agents_past (PRIMAL:49, 109, 128) */
  int blocking;          /* PRIMAL: compute get_blocking_reward (PRIMAL:513-546) with BFS path lengths */
  /* PARTIAL (marl_partial.py) */
  int pW, pK;                                   /* obs_window, obs_knn_agents */
  double p_move, p_stay, p_stay_goal, p_nc, p_ec, p_envc, p_complete, p_fac, p_gamma;
  uint8_t* at_goal;    /* [E,N] _agent_at_goals */
  int32_t* goal_cost;  /* [E,N] _each_goal_cost */
  int32_t* agent_steps;/* [E,N] _agent_step_count */
  int16_t* pnode;      /* [E,N] _node_collision_agents */
  int16_t* pedge;      /* [E,N] _edge_collision_agents */
  int64_t* total_coll; /* [E]   _total_number_collisions */
  uint8_t* terminated; /* [E]   _terminated */
  int16_t* pdist;      /* [E,N,H,W] _goal_dist */
} oracle_env;

static const int8_t* env_map(const oracle_env* o, int e) {
  return o->map + (size_t)(o->shared_map ? 0 : e) * o->H * o->W;
}

int oracle_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

oracle_env* oracle_create(int E, int N, int H, int W, int F, int shared_map, int episode_limit,
                          double step_reward, double collide_reward, double action_cost, double idle_cost,
                          double goal_reward, double collision_reward, int threads, int sum_mode,
                          int step_is_int, int collide_is_int) {
  oracle_env* o = (oracle_env*)calloc(1, sizeof(oracle_env));
  o->E = E; o->N = N; o->H = H; o->W = W; o->F = F;
  o->shared_map = shared_map;
  o->episode_limit = episode_limit;
  o->step_reward = step_reward; o->collide_reward = collide_reward;
  o->action_cost = action_cost; o->idle_cost = idle_cost;
  o->goal_reward = goal_reward; o->collision_reward = collision_reward;
  o->sum_mode = sum_mode; o->step_is_int = step_is_int; o->collide_is_int = collide_is_int;
  size_t cells = (size_t)H * W;
  o->map = (int8_t*)calloc((shared_map ? 1 : (size_t)E) * cells, 1);
  o->state = (int16_t*)calloc((size_t)E * cells, sizeof(int16_t));
  o->goals = (int16_t*)calloc((size_t)E * cells, sizeof(int16_t));
  o->pos = (int16_t*)calloc((size_t)E * N * 2, sizeof(int16_t));
  o->goal = (int16_t*)calloc((size_t)E * N * 2, sizeof(int16_t));
  o->start = (int16_t*)calloc((size_t)E * N * 2, sizeof(int16_t));
  o->done = (uint8_t*)calloc((size_t)E * N, 1);
  o->step_count = (int32_t*)calloc((size_t)E, sizeof(int32_t));
  o->threads = threads > 0 ? threads : oracle_max_threads();
  o->at_goal = (uint8_t*)calloc((size_t)E * N, 1);
  o->goal_cost = (int32_t*)calloc((size_t)E * N, sizeof(int32_t));
  o->agent_steps = (int32_t*)calloc((size_t)E * N, sizeof(int32_t));
  o->pnode = (int16_t*)calloc((size_t)E * N, sizeof(int16_t));
  o->pedge = (int16_t*)calloc((size_t)E * N, sizeof(int16_t));
  o->total_coll = (int64_t*)calloc((size_t)E, sizeof(int64_t));
  o->terminated = (uint8_t*)calloc((size_t)E, 1);
  o->pdist = NULL;
  o->past = (int16_t*)calloc((size_t)E * N * 2, sizeof(int16_t));
  return o;
}

void oracle_destroy(oracle_env* o) {
  if (!o) return;
  free(o->map); free(o->state); free(o->goals); free(o->pos); free(o->goal); free(o->start);
  free(o->done); free(o->step_count);
  free(o->at_goal); free(o->goal_cost); free(o->agent_steps); free(o->pnode); free(o->pedge);
  free(o->total_coll); free(o->terminated); free(o->pdist); free(o->past); free(o);
}

void oracle_get_positions(const oracle_env* o, int16_t* out) { memcpy(out, o->pos, (size_t)o->E * o->N * 4); }
void oracle_get_dones(const oracle_env* o, uint8_t* out) { memcpy(out, o->done, (size_t)o->E * o->N); }
void oracle_get_step_count(const oracle_env* o, int32_t* out) { memcpy(out, o->step_count, (size_t)o->E * 4); }

/* ------------------------------------------------------------------------------------------
 * GRID   (mapf_gridworld.py)
 * ---------------------------------------------------------------------------------------- */

/* __create_grid + __init_full_obs + __update_agent_view, GRID:282-299: grid = -1 on obstacles,
 * else the number of agents standing on the cell. */
static void grid_rebuild_full_obs(oracle_env* o, int e) {
  const int8_t* m = env_map(o, e);
  int16_t* st = o->state + (size_t)e * o->H * o->W;
  for (int c = 0; c < o->H * o->W; ++c) st[c] = m[c] ? -1 : 0;
  const int16_t* p = o->pos + (size_t)e * o->N * 2;
  for (int i = 0; i < o->N; ++i) st[p[2 * i] * o->W + p[2 * i + 1]] += 1;
}

/* reset, GRID:70-83 (positions restored from _agent_init_pos; counters zeroed). */
void oracle_grid_reset(oracle_env* o, const int8_t* map, const int16_t* starts, const int16_t* goals) {
  size_t cells = (size_t)o->H * o->W;
  if (map) memcpy(o->map, map, (o->shared_map ? 1 : (size_t)o->E) * cells);
  if (starts) memcpy(o->start, starts, (size_t)o->E * o->N * 4);
  if (goals) memcpy(o->goal, goals, (size_t)o->E * o->N * 4);
  memcpy(o->pos, o->start, (size_t)o->E * o->N * 4);
  memset(o->done, 0, (size_t)o->E * o->N);
  memset(o->step_count, 0, (size_t)o->E * 4);
#pragma omp parallel for schedule(static) num_threads(o->threads)
  for (int e = 0; e < o->E; ++e) grid_rebuild_full_obs(o, e);
}

/* __is_valid GRID:270 and __is_cell_obstacle GRID:278 / PARTIAL:521: a cell is an obstacle iff `_full_obs` reads -1
 * there.  `_full_obs` is -1 on walls PLUS the number of agents on the cell (GRID:299), so a wall cell that holds an
 * agent (the .scen x/y transposition of GRID:432-443 produces such starts) reads >= 0 and is NOT an obstacle.
 * `st` is the environment's _full_obs as it stands when the reference evaluates the test: the pre-step grid inside
 * step() (it is only rebuilt at GRID:132-135), the rebuilt grid for the avail masks (GRID:140). */
static int grid_free(const oracle_env* o, const int16_t* st, int p0, int p1) {
  if (!(0 <= p0 && p0 < o->H && 0 <= p1 && p1 < o->W)) return 0;
  return st[p0 * o->W + p1] != -1;
}

/* get_avail_agent_actions, GRID:203-224. */
static void grid_avail_agent(const oracle_env* o, const int16_t* m, int p0, int p1, uint8_t* out5) {
  out5[0] = (uint8_t)grid_free(o, m, p0 - 1, p1);
  out5[1] = (uint8_t)grid_free(o, m, p0 + 1, p1);
  out5[2] = (uint8_t)grid_free(o, m, p0, p1 - 1);
  out5[3] = (uint8_t)grid_free(o, m, p0, p1 + 1);
  out5[4] = 1;
}

void oracle_grid_avail(const oracle_env* o, uint8_t* avail) {
#pragma omp parallel for schedule(static) num_threads(o->threads)
  for (int e = 0; e < o->E; ++e) {
    const int16_t* m = o->state + (size_t)e * o->H * o->W;
    for (int i = 0; i < o->N; ++i) {
      const int16_t* p = o->pos + ((size_t)e * o->N + i) * 2;
      grid_avail_agent(o, m, p[0], p[1], avail + ((size_t)e * o->N + i) * 5);
    }
  }
}

/* get_obs / get_state, GRID:143-196: the flattened _full_obs (every agent sees the same row). */
void oracle_grid_state(const oracle_env* o, int8_t* out) {
  size_t n = (size_t)o->E * o->H * o->W;
  for (size_t k = 0; k < n; ++k) out[k] = (int8_t)o->state[k];
}

/* `sum(rewards)`, GRID:141, as the interpreter running the reference computes it.
 *   sum_mode 0: plain left fold (builtin sum of CPython <= 3.11).
 *   sum_mode 1: CPython >= 3.12 builtin sum (Python/bltinmodule.c): exact integer accumulation while the
 *     items are ints; at the first float the partial sum becomes float(i) + x; afterwards floats are
 *     added with Neumaier compensation, ints with a plain add; the compensation is folded in at the end.
 * is_int[i] tells whether rewards[i] is a Python int (no float constant was ever added to it). */
static double py_sum(const double* x, const uint8_t* is_int, int n, int sum_mode) {
  if (sum_mode == 0) {
    double total = 0.0;
    for (int i = 0; i < n; ++i) total += x[i];
    return total;
  }
  double acc = 0.0, c = 0.0;
  int in_float = 0;
  for (int i = 0; i < n; ++i) {
    if (!in_float) {
      acc = acc + x[i];                 /* exact while ints; the first float item: float(i_result) + x */
      if (!is_int[i]) { in_float = 1; c = 0.0; }
    } else if (is_int[i]) {
      acc += x[i];
    } else {
      double t = acc + x[i];
      if (fabs(acc) >= fabs(x[i])) c += (acc - t) + x[i];
      else c += (x[i] - t) + acc;
      acc = t;
    }
  }
  if (in_float && c != 0.0 && isfinite(c)) acc += c;
  return acc;
}

/* step, GRID:85-141 for one environment. Returns the number of invalid actions seen. */
static int grid_step_env(oracle_env* o, int e, const uint8_t* act, double* reward, uint8_t* terminated,
                         double* agent_reward, int8_t* envflag, int16_t* node_out, int16_t* edge_out,
                         uint8_t* avail, int16_t* scratch /* [2N + H*W] */) {
  const int N = o->N, W = o->W;
  const int16_t* m = o->state + (size_t)e * o->H * W;   /* _full_obs: pre-step until GRID:132 rebuilds it */
  int16_t* pos = o->pos + (size_t)e * N * 2;
  const int16_t* goal = o->goal + (size_t)e * N * 2;
  uint8_t* done = o->done + (size_t)e * N;
  int16_t* newp = scratch;            /* new_agent_position, GRID:95 */
  int16_t* cnt = scratch + 2 * N;     /* agent_at_grid, GRID:348 */
  double rew_local[256];
  uint8_t rew_is_int[256];
  int bad = 0;
  o->step_count[e] += 1;                                            /* GRID:93 */
  for (int i = 0; i < N; ++i) {                                     /* GRID:99-118 */
    int n0 = pos[2 * i], n1 = pos[2 * i + 1];
    double r = 0.0;
    int flag = 0;
    rew_is_int[i] = (uint8_t)(done[i] ? o->collide_is_int : (o->collide_is_int && o->step_is_int));
    if (!done[i]) {
      int a = act[i];
      int t0 = n0, t1 = n1;                                         /* __agent_step, GRID:319-342 */
      if (a == 0) t0 -= 1; else if (a == 1) t0 += 1; else if (a == 2) t1 -= 1; else if (a == 3) t1 += 1;
      else if (a != 4) bad++;
      if (a >= 0 && a <= 3) {
        if (grid_free(o, m, t0, t1)) { n0 = t0; n1 = t1; } else flag = 1;
      }
      if (flag) r += o->collide_reward;                             /* GRID:105-106 */
      r += o->step_reward;                                          /* GRID:110 */
    }
    newp[2 * i] = (int16_t)n0; newp[2 * i + 1] = (int16_t)n1;
    if (n0 == goal[2 * i] && n1 == goal[2 * i + 1]) done[i] = 1;    /* GRID:112-113 */
    if (o->step_count[e] >= o->episode_limit) done[i] = 1;          /* GRID:116-117 */
    rew_local[i] = r;
    if (envflag) envflag[i] = (int8_t)flag;
  }
  /* __count_node_collision, GRID:344-362 */
  memset(cnt, 0, sizeof(int16_t) * (size_t)o->H * W);
  for (int i = 0; i < N; ++i) cnt[newp[2 * i] * W + newp[2 * i + 1]] += 1;
  /* __count_edge_collision, GRID:364-383 */
  for (int i = 0; i < N; ++i) {
    int node = cnt[newp[2 * i] * W + newp[2 * i + 1]] > 1 ? 1 : 0;
    int edge = 0;
    int io0 = pos[2 * i], io1 = pos[2 * i + 1], in0 = newp[2 * i], in1 = newp[2 * i + 1];
    if (!(io0 == in0 && io1 == in1)) {
      for (int j = 0; j < N; ++j) {
        if (j == i) continue;
        if (pos[2 * j] == in0 && pos[2 * j + 1] == in1) {           /* j_old == i_new */
          int jn0 = newp[2 * j], jn1 = newp[2 * j + 1];
          if (jn0 == io0 && jn1 == io1 && !(jn0 == in0 && jn1 == in1)) edge++;
        }
      }
    }
    rew_local[i] += o->collide_reward * node;                       /* GRID:128 */
    rew_local[i] += o->collide_reward * edge;                       /* GRID:129 */
    if (node_out) node_out[i] = (int16_t)node;
    if (edge_out) edge_out[i] = (int16_t)edge;
  }
  /* GRID:132-135 */
  for (int i = 0; i < 2 * N; ++i) pos[i] = newp[i];
  grid_rebuild_full_obs(o, e);
  double total = py_sum(rew_local, rew_is_int, N, o->sum_mode);     /* sum(rewards), GRID:141 */
  int all_done = 1;
  for (int i = 0; i < N; ++i) {
    if (agent_reward) agent_reward[i] = rew_local[i];
    all_done &= done[i];
  }
  if (reward) *reward = total;
  if (terminated) *terminated = (uint8_t)all_done;                  /* episode_done, GRID:267 */
  if (avail)                                                        /* GRID:140 */
    for (int i = 0; i < N; ++i) grid_avail_agent(o, m, pos[2 * i], pos[2 * i + 1], avail + 5 * i);
  return bad;
}

int oracle_grid_step(oracle_env* o, const uint8_t* actions, double* reward, uint8_t* terminated,
                     double* agent_reward, int8_t* envflag, int16_t* node, int16_t* edge, uint8_t* avail) {
  int bad = 0;
  const int N = o->N;
#pragma omp parallel num_threads(o->threads) reduction(+ : bad)
  {
    int16_t* scratch = (int16_t*)malloc(sizeof(int16_t) * (2 * (size_t)N + (size_t)o->H * o->W));
#pragma omp for schedule(static)
    for (int e = 0; e < o->E; ++e) {
      size_t b = (size_t)e * N;
      bad += grid_step_env(o, e, actions + b, reward ? reward + e : 0, terminated ? terminated + e : 0,
                           agent_reward ? agent_reward + b : 0, envflag ? envflag + b : 0,
                           node ? node + b : 0, edge ? edge + b : 0, avail ? avail + 5 * b : 0, scratch);
    }
    free(scratch);
  }
  return bad;
}

/* ------------------------------------------------------------------------------------------
 * PRIMAL   (mapf_primal.py)
 * ---------------------------------------------------------------------------------------- */

static const int PRIMAL_DIR[9][2] = {{0, 0}, {0, 1}, {1, 0}, {0, -1}, {-1, 0},
                                     {1, 1}, {1, -1}, {-1, -1}, {-1, 1}};           /* dirDict, PRIMAL:28 */
static const int PRIMAL_OPPOSITE[9] = {-1, 3, 4, 1, 2, 7, 8, 5, 6};                 /* opposite_actions, PRIMAL:26 */

void oracle_primal_set_diagonal(oracle_env* o, int on) { o->diagonal = on; }
int oracle_n_actions(const oracle_env* o) { return o->diagonal ? 9 : 5; }

/* _setWorld with world0/goals0 + State.__init__/scanForAgents, PRIMAL:278-309, 44-66. */
void oracle_primal_reset(oracle_env* o, const int8_t* map, const int16_t* starts, const int16_t* goals) {
  size_t cells = (size_t)o->H * o->W;
  if (map) memcpy(o->map, map, (o->shared_map ? 1 : (size_t)o->E) * cells);
  if (starts) memcpy(o->start, starts, (size_t)o->E * o->N * 4);
  if (goals) memcpy(o->goal, goals, (size_t)o->E * o->N * 4);
  memcpy(o->pos, o->start, (size_t)o->E * o->N * 4);
  memcpy(o->past, o->start, (size_t)o->E * o->N * 4);                /* agents_last == agents, PRIMAL:61-65 */
  memset(o->step_count, 0, (size_t)o->E * 4);
#pragma omp parallel for schedule(static) num_threads(o->threads)
  for (int e = 0; e < o->E; ++e) {
    const int8_t* m = env_map(o, e);
    int16_t* st = o->state + (size_t)e * cells;
    int16_t* gg = o->goals + (size_t)e * cells;
    for (size_t c = 0; c < cells; ++c) { st[c] = m[c] ? -1 : 0; gg[c] = 0; }
    for (int i = 0; i < o->N; ++i) {
      const int16_t* p = o->pos + ((size_t)e * o->N + i) * 2;
      const int16_t* g = o->goal + ((size_t)e * o->N + i) * 2;
      st[p[0] * o->W + p[1]] = (int16_t)(i + 1);
      gg[g[0] * o->W + g[1]] = (int16_t)(i + 1);
      o->done[(size_t)e * o->N + i] = (uint8_t)(p[0] == g[0] && p[1] == g[1]);
    }
  }
}

/* Overwrite goals of flagged agents (lifelong variant; not in PRIMAL itself). */
void oracle_primal_set_goals(oracle_env* o, const int16_t* goals, const uint8_t* dirty) {
  size_t cells = (size_t)o->H * o->W;
  for (int e = 0; e < o->E; ++e) {
    int16_t* gg = o->goals + (size_t)e * cells;
    for (int i = 0; i < o->N; ++i) {
      size_t k = (size_t)e * o->N + i;
      if (dirty && !dirty[k]) continue;
      int16_t* g = o->goal + k * 2;
      if (gg[g[0] * o->W + g[1]] == i + 1) gg[g[0] * o->W + g[1]] = 0;
      g[0] = goals[2 * k]; g[1] = goals[2 * k + 1];
      gg[g[0] * o->W + g[1]] = (int16_t)(i + 1);
    }
  }
}

/* State.diagonalCollision, PRIMAL:77-100: the midpoint of this move equals the midpoint of another agent's last
 * recorded move (past -> present).  np.isclose on half-integers == equality of the integer sums. */
static int primal_diagonal_collision(const oracle_env* o, int e, int id, int nx, int ny) {
  const int16_t* pos = o->pos + (size_t)e * o->N * 2;
  const int16_t* past = o->past + (size_t)e * o->N * 2;
  const int lx = pos[2 * (id - 1)], ly = pos[2 * (id - 1) + 1];
  for (int a = 1; a <= o->N; ++a) {
    if (a == id) continue;
    if (past[2 * (a - 1)] + pos[2 * (a - 1)] == lx + nx && past[2 * (a - 1) + 1] + pos[2 * (a - 1) + 1] == ly + ny)
      return 1;
  }
  return 0;
}

/* State.moveAgent, PRIMAL:103-135. */
static int primal_move_agent(oracle_env* o, int e, int id, int action) {
  const int W = o->W, H = o->H;
  int16_t* st = o->state + (size_t)e * H * W;
  const int16_t* gg = o->goals + (size_t)e * H * W;
  int16_t* p = o->pos + ((size_t)e * o->N + (id - 1)) * 2;
  int16_t* pp = o->past + ((size_t)e * o->N + (id - 1)) * 2;
  int ax = p[0], ay = p[1];
  int dx = PRIMAL_DIR[action][0], dy = PRIMAL_DIR[action][1];
  if (dx == 0 && dy == 0) {
    pp[0] = (int16_t)ax; pp[1] = (int16_t)ay;                        /* :109 */
    return gg[ax * W + ay] == id ? 1 : 0;
  }
  if (ax + dx >= H || ax + dx < 0 || ay + dy >= W || ay + dy < 0) return -1;
  if (st[(ax + dx) * W + ay + dy] < 0) return -2;
  if (st[(ax + dx) * W + ay + dy] > 0) return -3;
  if (o->diagonal && primal_diagonal_collision(o, e, id, ax + dx, ay + dy)) return -3;   /* :122-124 */
  st[ax * W + ay] = 0;
  st[(ax + dx) * W + ay + dy] = (int16_t)id;
  pp[0] = (int16_t)ax; pp[1] = (int16_t)ay;                          /* :128 */
  p[0] = (int16_t)(ax + dx); p[1] = (int16_t)(ay + dy);
  if (gg[(ax + dx) * W + ay + dy] == id) return 1;
  if (gg[ax * W + ay] == id) return 2;
  return 0;
}

void oracle_primal_set_blocking(oracle_env* o, int on) { o->blocking = on; }

/* Single-robot shortest path length on the 4-connected grid with `blocked` cells removed: what
 * od_mstar3.cpp_mstar.find_path(world, [start], [goal], 1, 5) returns for ONE robot is an optimal path, so only its
 * length matters to get_blocking_reward (len(path) = hops + 1).  -1 = NoSolutionError (PRIMAL:505-508). */
static int single_robot_path_hops(const oracle_env* o, const int8_t* m, const uint8_t* blocked, int s, int g,
                                  int32_t* queue, int16_t* dist) {
  const int H = o->H, W = o->W;
  if (m[s] || blocked[s] || m[g] || blocked[g]) return -1;
  for (int c = 0; c < H * W; ++c) dist[c] = -1;
  int head = 0, tail = 0;
  dist[s] = 0;
  queue[tail++] = s;
  while (head < tail) {
    int c = queue[head++];
    if (c == g) return dist[c];
    int r0 = c / W, c0 = c % W;
    static const int D[4][2] = {{0, 1}, {1, 0}, {0, -1}, {-1, 0}};
    for (int q = 0; q < 4; ++q) {
      int r1 = r0 + D[q][0], c1 = c0 + D[q][1];
      if (r1 < 0 || r1 >= H || c1 < 0 || c1 >= W) continue;
      int n = r1 * W + c1;
      if (m[n] || blocked[n] || dist[n] >= 0) continue;
      dist[n] = (int16_t)(dist[c] + 1);
      queue[tail++] = n;
    }
  }
  return -1;
}

/* get_blocking_reward, PRIMAL:513-546: how many visible robots (ids 1..N-1 -- the loop skips the last id, :523)
 * have their path to their goal cut, or lengthened by more than 10, by this robot standing where it stands. */
static double primal_blocking_reward(const oracle_env* o, int e, int id) {
  const int H = o->H, W = o->W, N = o->N, F = o->F;
  const int8_t* m = env_map(o, e);
  const int16_t* pos = o->pos + (size_t)e * N * 2;
  const int16_t* goal = o->goal + (size_t)e * N * 2;
  uint8_t* blocked = (uint8_t*)calloc((size_t)H * W, 1);
  int32_t* queue = (int32_t*)malloc(sizeof(int32_t) * (size_t)H * W);
  int16_t* dist = (int16_t*)malloc(sizeof(int16_t) * (size_t)H * W);
  int others[256], n_others = 0;
  const int tl0 = pos[2 * (id - 1)] - F / 2, tl1 = pos[2 * (id - 1) + 1] - F / 2;
  for (int a = 1; a < N; ++a) {                                       /* range(1, num_agents) */
    if (a == id) continue;
    int x = pos[2 * (a - 1)], y = pos[2 * (a - 1) + 1];
    if (x < tl0 || x >= tl0 + F || y >= tl1 + F || y < tl1) continue;
    others[n_others++] = a;
    blocked[x * W + y] = 1;
  }
  const int me = pos[2 * (id - 1)] * W + pos[2 * (id - 1) + 1];
  int num_blocking = 0;
  for (int k = 0; k < n_others; ++k) {
    const int a = others[k];
    const int s = pos[2 * (a - 1)] * W + pos[2 * (a - 1) + 1];
    const int g = goal[2 * (a - 1)] * W + goal[2 * (a - 1) + 1];
    blocked[s] = 0;                                                   /* other_locations.remove(pos(agent)) */
    const uint8_t me_was = blocked[me];
    blocked[me] = 1;
    const int before = single_robot_path_hops(o, m, blocked, s, g, queue, dist);
    blocked[me] = me_was;
    const int after = single_robot_path_hops(o, m, blocked, s, g, queue, dist);
    blocked[s] = 1;
    if (before < 0 && after < 0) continue;
    if (before >= 0 && after < 0) continue;
    if ((before < 0 && after >= 0) || before > after + 10) num_blocking++;   /* len(path) = hops + 1 on both sides */
  }
  free(blocked); free(queue); free(dist);
  return num_blocking * -1.0;                                         /* BLOCKING_COST, PRIMAL:25 */
}

/* State.done, PRIMAL:159-165. */
static int primal_done(const oracle_env* o, int e) {
  const int16_t* gg = o->goals + (size_t)e * o->H * o->W;
  int complete = 0;
  for (int i = 1; i <= o->N; ++i) {
    const int16_t* p = o->pos + ((size_t)e * o->N + (i - 1)) * 2;
    if (gg[p[0] * o->W + p[1]] == i) complete++;
  }
  return complete == o->N;
}

/* _listNextValidActions, PRIMAL:639-667, as an n_actions-entry mask (5, or 9 with DIAGONAL_MOVEMENT). */
static void primal_valid_actions(const oracle_env* o, int e, int id, int prev_action, uint8_t* out) {
  const int W = o->W, H = o->H, na = o->diagonal ? 9 : 5;
  const int16_t* st = o->state + (size_t)e * H * W;
  const int16_t* p = o->pos + ((size_t)e * o->N + (id - 1)) * 2;
  int ax = p[0], ay = p[1];
  out[0] = 1;
  for (int a = 1; a < na; ++a) {
    int dx = PRIMAL_DIR[a][0], dy = PRIMAL_DIR[a][1];
    out[a] = 0;
    if (ax + dx >= H || ax + dx < 0 || ay + dy >= W || ay + dy < 0) continue;
    if (st[(ax + dx) * W + ay + dy] < 0) continue;
    if (st[(ax + dx) * W + ay + dy] > 0) continue;
    if (o->diagonal && primal_diagonal_collision(o, e, id, ax + dx, ay + dy)) continue;   /* :658-660 */
    out[a] = 1;
  }
  int opp = PRIMAL_OPPOSITE[prev_action];
  if (opp >= 0) out[opp] = 0;                                       /* PRIMAL:664-665 */
}

/* The engine's PRIMAL team reward (include/mapf_b200.h, mapf_step_out.reward_dev; no reference counterpart): the
 * pairwise sum of x[0..n) -- leaves padded with +0.0 to the next power of two, y[i] += y[i + s] for s = 1, 2, 4, ... */
static double pairwise_sum(const double* x, int n) {
  double y[256];
  int m = 1;
  while (m < n) m <<= 1;
  for (int i = 0; i < m; ++i) y[i] = i < n ? x[i] : 0.0;
  for (int s = 1; s < m; s <<= 1)
    for (int i = 0; i + s < m; i += 2 * s) y[i] = y[i] + y[i + s];
  return n > 0 ? y[0] : 0.0;
}

/* One sweep `for id in lo+1..hi: _step((id, a))`, PRIMAL:549-637 without the observation
 * (observe_all is separate) and with the blocking reward fenced off (returns 0, see refload.py). */
int oracle_primal_sweep(oracle_env* o, const uint8_t* actions, int lo, int hi, int8_t* status_out,
                        double* agent_reward, uint8_t* on_goal_out, uint8_t* valid_out, uint8_t* done_mid,
                        uint8_t* next_mid, uint8_t* avail, uint8_t* terminated, double* reward,
                        uint8_t* blocking_out) {
  int bad = 0;
  const int N = o->N;
#pragma omp parallel for schedule(static) num_threads(o->threads) reduction(+ : bad)
  for (int e = 0; e < o->E; ++e) {
    double rloc[256];
    for (int i = 0; i < N; ++i) rloc[i] = 0.0;
    if (lo == 0) o->step_count[e] += 1;
    for (int i = lo; i < hi; ++i) {
      size_t k = (size_t)e * N + i;
      int id = i + 1;
      int action = actions[k];
      if (action > (o->diagonal ? 8 : 4)) { bad++; action = 0; }
      int status = primal_move_agent(o, e, id, action);            /* world.act, PRIMAL:570 */
      double r;
      int is_blocking = 0;
      if (action == 0) {                                            /* PRIMAL:579-587 */
        if (status == 1) {
          double x = o->blocking ? primal_blocking_reward(o, e, id) : 0.0 * -1.0;   /* num_blocking * BLOCKING_COST */
          r = o->goal_reward + x;
          is_blocking = x < 0;
        } else r = o->idle_cost;
      } else {                                                      /* PRIMAL:588-596 */
        if (status == 1) r = o->goal_reward;
        else if (status < 0) r = o->collision_reward;
        else r = o->action_cost;
      }
      rloc[i] = r;
      const int16_t* p = o->pos + k * 2;
      const int16_t* g = o->goal + k * 2;
      int on_goal = (p[0] == g[0] && p[1] == g[1]);                 /* PRIMAL:633 */
      o->done[k] = (uint8_t)on_goal;
      if (status_out) status_out[k] = (int8_t)status;
      if (blocking_out) blocking_out[k] = (uint8_t)is_blocking;
      if (agent_reward) agent_reward[k] = r;
      if (on_goal_out) on_goal_out[k] = (uint8_t)on_goal;
      if (valid_out) valid_out[k] = (uint8_t)(status >= 0);         /* PRIMAL:571 */
      if (done_mid) done_mid[k] = (uint8_t)primal_done(o, e);       /* PRIMAL:626 */
      if (next_mid) primal_valid_actions(o, e, id, action, next_mid + (size_t)(o->diagonal ? 9 : 5) * k); /* PRIMAL:630 */
    }
    if (avail)
      for (int i = 0; i < N; ++i) {
        size_t k = (size_t)e * N + i;
        int prev = (i >= lo && i < hi) ? (actions[k] > (o->diagonal ? 8 : 4) ? 0 : actions[k]) : 0;
        primal_valid_actions(o, e, i + 1, prev, avail + (size_t)(o->diagonal ? 9 : 5) * k);
      }
    if (terminated) terminated[e] = (uint8_t)primal_done(o, e);
    if (reward) reward[e] = pairwise_sum(rloc, N);
  }
  return bad;
}

/* _listNextValidActions(id, prev_action) for every agent on the current state. prev may be NULL (= 0). */
void oracle_primal_avail(const oracle_env* o, const uint8_t* prev_action, uint8_t* avail) {
#pragma omp parallel for schedule(static) num_threads(o->threads)
  for (int e = 0; e < o->E; ++e)
    for (int i = 0; i < o->N; ++i) {
      size_t k = (size_t)e * o->N + i;
      primal_valid_actions(o, e, i + 1, prev_action ? prev_action[k] : 0, avail + (size_t)(o->diagonal ? 9 : 5) * k);
    }
}

/* _observe(agent_id), PRIMAL:343-386.  maps: uint8 [4][F][F] in the returned order
 * [poss_map, goal_map, goals_map, obs_map] (values are exactly 0.0/1.0 in the reference). */
static void primal_observe_agent(const oracle_env* o, int e, int id, uint8_t* maps, double* vec) {
  const int H = o->H, W = o->W, F = o->F, N = o->N;
  const int16_t* st = o->state + (size_t)e * H * W;
  const int16_t* gg = o->goals + (size_t)e * H * W;
  const int16_t* p = o->pos + ((size_t)e * N + (id - 1)) * 2;
  const int16_t* g = o->goal + ((size_t)e * N + (id - 1)) * 2;
  int tl0 = p[0] - F / 2, tl1 = p[1] - F / 2;                       /* PRIMAL:345-346 */
  uint8_t* poss_map = maps;
  uint8_t* goal_map = maps + F * F;
  uint8_t* goals_map = maps + 2 * F * F;
  uint8_t* obs_map = maps + 3 * F * F;
  memset(maps, 0, (size_t)4 * F * F);
  int visible[256];
  int nvis = 0;
  for (int i = tl0; i < tl0 + F; ++i)                               /* PRIMAL:354-372 */
    for (int j = tl1; j < tl1 + F; ++j) {
      int w = (i - tl0) * F + (j - tl1);
      if (i >= H || i < 0 || j >= W || j < 0) { obs_map[w] = 1; continue; }
      int s = st[i * W + j];
      if (s == -1) obs_map[w] = 1;
      if (s == id) poss_map[w] = 1;
      if (gg[i * W + j] == id) goal_map[w] = 1;
      if (s > 0 && s != id) { visible[nvis++] = s; poss_map[w] = 1; }
    }
  for (int v = 0; v < nvis; ++v) {                                  /* PRIMAL:374-378 */
    const int16_t* og = o->goal + ((size_t)e * N + (visible[v] - 1)) * 2;
    int x = og[0], y = og[1];
    int mx = x < tl0 + F - 1 ? x : tl0 + F - 1; if (mx < tl0) mx = tl0;
    int my = y < tl1 + F - 1 ? y : tl1 + F - 1; if (my < tl1) my = tl1;
    goals_map[(mx - tl0) * F + (my - tl1)] = 1;
  }
  if (vec) {                                                        /* PRIMAL:380-385 */
    int dx = g[0] - p[0], dy = g[1] - p[1];
    double mag = pow((double)(dx * dx + dy * dy), 0.5);
    double fx = (double)dx, fy = (double)dy;
    if (mag != 0) { fx = fx / mag; fy = fy / mag; }
    vec[0] = fx; vec[1] = fy; vec[2] = mag;
  }
}

void oracle_primal_observe(const oracle_env* o, uint8_t* obs, double* vec) {
  const int N = o->N, F = o->F;
#pragma omp parallel for schedule(static) num_threads(o->threads)
  for (int e = 0; e < o->E; ++e)
    for (int i = 0; i < N; ++i) {
      size_t k = (size_t)e * N + i;
      primal_observe_agent(o, e, i + 1, obs + k * 4 * F * F, vec ? vec + 3 * k : 0);
    }
}

/* ------------------------------------------------------------------------------------------
 * Per-goal hop-distance maps
 *   PARTIAL __setup_agent_goal_dist / __get_path_dist, PARTIAL:931-955: networkx A* length on the
 *     4-connected unit-weight graph of free cells == BFS hop count (edge attribute 'cost' is
 *     absent, so every edge weighs 1).
 *   PRIMAL getAstarCosts, PRIMAL:407-499: exhaustive A* from the goal (runs until the open set is
 *     empty) == BFS hop count; `costs = state.copy()` first, so walls are -1 and unreachable cells
 *     keep `state` (0 or an agent id) -- selected with primal_costs != 0.
 * Output int16[E,N,H,W]: walls -1, unreachable free cells -2 (PARTIAL raises NetworkXNoPath there).
 * ---------------------------------------------------------------------------------------- */
void oracle_goal_dist(const oracle_env* o, const uint8_t* dirty, int primal_costs, int16_t* dist) {
  const int H = o->H, W = o->W, N = o->N;
  const size_t cells = (size_t)H * W;
#pragma omp parallel num_threads(o->threads)
  {
    int32_t* queue = (int32_t*)malloc(sizeof(int32_t) * cells);
#pragma omp for schedule(static)
    for (int e = 0; e < o->E; ++e) {
      const int8_t* m = env_map(o, e);
      const int16_t* st = o->state + (size_t)e * cells;
      for (int i = 0; i < N; ++i) {
        size_t k = (size_t)e * N + i;
        if (dirty && !dirty[k]) continue;
        int16_t* d = dist + k * cells;
        for (size_t c = 0; c < cells; ++c) d[c] = m[c] ? -1 : -2;
        const int16_t* g = o->goal + k * 2;
        int head = 0, tail = 0;
        int gc = g[0] * W + g[1];
        if (!m[gc]) { d[gc] = 0; queue[tail++] = gc; }
        while (head < tail) {
          int c = queue[head++];
          int r0 = c / W, c0 = c % W;
          static const int D[8][2] = {{-1, 0}, {1, 0}, {0, -1}, {0, 1}, {1, 1}, {1, -1}, {-1, -1}, {-1, 1}};
          const int nd = (primal_costs && o->diagonal) ? 8 : 4;      /* getNeighbors, PRIMAL:421-437 */
          for (int q = 0; q < nd; ++q) {
            int r1 = r0 + D[q][0], c1 = c0 + D[q][1];
            if (r1 < 0 || r1 >= H || c1 < 0 || c1 >= W) continue;
            int n = r1 * W + c1;
            if (d[n] != -2) continue;
            d[n] = (int16_t)(d[c] + 1);
            queue[tail++] = n;
          }
        }
        if (primal_costs)                                          /* PRIMAL:496-498 */
          for (size_t c = 0; c < cells; ++c)
            if (d[c] == -2) d[c] = st[c];
      }
    }
    free(queue);
  }
}


/* ------------------------------------------------------------------------------------------
 * PARTIAL   (MARL-curve-main/src/envs/marl_partial.py) -- the env the reference registers
 * ---------------------------------------------------------------------------------------- */
void oracle_partial_config(oracle_env* o, int obs_window, int obs_knn, double move, double stay, double stay_goal,
                           double nc, double ec, double envc, double complete, double fac, double gamma) {
  o->pW = obs_window; o->pK = obs_knn;
  o->p_move = move; o->p_stay = stay; o->p_stay_goal = stay_goal;
  o->p_nc = nc; o->p_ec = ec; o->p_envc = envc;
  o->p_complete = complete; o->p_fac = fac; o->p_gamma = gamma;
}

/* reset, PARTIAL:125-167 with pinned starts/goals (the reference samples them from a random .scen,
 * :907-927) + __setup_agent_goal_dist :931-945. */
void oracle_partial_reset(oracle_env* o, const int8_t* map, const int16_t* starts, const int16_t* goals) {
  oracle_grid_reset(o, map, starts, goals);
  size_t EN = (size_t)o->E * o->N;
  memset(o->at_goal, 0, EN);
  for (size_t k = 0; k < EN; ++k) { o->goal_cost[k] = -1; o->agent_steps[k] = 0; o->pnode[k] = 0; o->pedge[k] = 0; }
  memset(o->total_coll, 0, sizeof(int64_t) * (size_t)o->E);
  memset(o->terminated, 0, (size_t)o->E);
  if (!o->pdist) o->pdist = (int16_t*)malloc(EN * o->H * o->W * sizeof(int16_t));
  oracle_goal_dist(o, NULL, 0, o->pdist);
}

/* step, PARTIAL:169-310 (output == False) for one environment. */
static int partial_step_env(oracle_env* o, int e, const uint8_t* act, double* reward, uint8_t* terminated_out,
                            double* agent_reward, uint8_t* avail, int16_t* scratch) {
  const int N = o->N, W = o->W, H = o->H;
  const int16_t* m = o->state + (size_t)e * H * W;      /* _full_obs: pre-step until :279 rebuilds it */
  int16_t* pos = o->pos + (size_t)e * N * 2;
  const int16_t* goal = o->goal + (size_t)e * N * 2;
  uint8_t* done = o->done + (size_t)e * N;
  uint8_t* at_goal = o->at_goal + (size_t)e * N;
  int16_t* newp = scratch;
  int16_t* cnt = scratch + 2 * N;
  double rew[256];
  uint8_t isint[256];
  int bad = 0;
  o->step_count[e] += 1;                                              /* :178 */
  const int step = o->step_count[e];
  for (int i = 0; i < N; ++i) {                                       /* :192-235 */
    int n0 = pos[2 * i], n1 = pos[2 * i + 1];
    double r = 0.0;
    if (!done[i]) {
      o->agent_steps[(size_t)e * N + i] += 1;
      int a = act[i];
      int t0 = n0, t1 = n1, flag = 0;                                 /* __agent_step :618-643 */
      if (a == 0) t0 -= 1; else if (a == 1) t0 += 1; else if (a == 2) t1 -= 1; else if (a == 3) t1 += 1;
      else if (a != 4) bad++;
      if (a >= 0 && a <= 3) {
        if (grid_free(o, m, t0, t1)) { n0 = t0; n1 = t1; } else flag = 1;
      }
      if (flag) r += o->p_envc;                                       /* :203-205 */
      if (a >= 0 && a <= 3) r += o->p_move;                           /* :207-208 */
      else r += at_goal[i] ? o->p_stay_goal : o->p_stay;              /* :209-213 */
    }
    newp[2 * i] = (int16_t)n0; newp[2 * i + 1] = (int16_t)n1;
    at_goal[i] = 0;                                                   /* :217-222 */
    if (n0 == goal[2 * i] && n1 == goal[2 * i + 1]) {
      o->goal_cost[(size_t)e * N + i] = step;
      at_goal[i] = 1;
    }
    if (step >= o->episode_limit) { o->terminated[e] = 1; done[i] = 1; }   /* :224-227 */
    const int16_t* dm = o->pdist + ((size_t)e * N + i) * H * W;       /* :229-234 */
    int opd = dm[pos[2 * i] * W + pos[2 * i + 1]], npd = dm[n0 * W + n1];
    double closer = (double)(opd - npd) / (double)o->episode_limit;
    r += closer;
    rew[i] = r;
    isint[i] = 0;                                                     /* closer_rew is always a float */
  }
  memset(cnt, 0, sizeof(int16_t) * (size_t)H * W);                    /* __check_node_collisions :713-737 */
  for (int i = 0; i < N; ++i) cnt[newp[2 * i] * W + newp[2 * i + 1]] += 1;
  long nsum = 0, esum = 0;
  for (int i = 0; i < N; ++i) {                                       /* __check_edge_collisions :822-857 */
    int node = cnt[newp[2 * i] * W + newp[2 * i + 1]] > 1 ? 1 : 0;
    int edge = 0;
    int io0 = pos[2 * i], io1 = pos[2 * i + 1], in0 = newp[2 * i], in1 = newp[2 * i + 1];
    if (!(io0 == in0 && io1 == in1))
      for (int j = 0; j < N; ++j) {
        if (j == i) continue;
        if (pos[2 * j] == in0 && pos[2 * j + 1] == in1 && newp[2 * j] == io0 && newp[2 * j + 1] == io1) edge++;
      }
    o->pnode[(size_t)e * N + i] = (int16_t)node;
    o->pedge[(size_t)e * N + i] = (int16_t)edge;
    nsum += node; esum += edge;
    rew[i] += o->p_nc * node;                                         /* :255-259 */
    rew[i] += o->p_ec * edge;
  }
  o->total_coll[e] += (nsum + esum) / 2;                              /* :250 */
  for (int i = 0; i < 2 * N; ++i) pos[i] = newp[i];                   /* :279-283 */
  grid_rebuild_full_obs(o, e);
  int all_goal = 1;
  for (int i = 0; i < N; ++i) all_goal &= at_goal[i];
  if (all_goal) {                                                     /* :291-299 */
    for (int i = 0; i < N; ++i) done[i] = 1;
    o->terminated[e] = 1;
    double tmp = (o->p_complete / pow(o->p_gamma, (double)(o->episode_limit - step))) * o->p_fac;
    for (int i = 0; i < N; ++i) rew[i] += tmp;
  }
  if (reward) *reward = py_sum(rew, isint, N, o->sum_mode);           /* :310 */
  if (terminated_out) *terminated_out = o->terminated[e];
  if (agent_reward) for (int i = 0; i < N; ++i) agent_reward[i] = rew[i];
  if (avail) for (int i = 0; i < N; ++i) grid_avail_agent(o, m, pos[2 * i], pos[2 * i + 1], avail + 5 * i);
  return bad;
}

int oracle_partial_step(oracle_env* o, const uint8_t* actions, double* reward, uint8_t* terminated,
                        double* agent_reward, uint8_t* avail) {
  int bad = 0;
  const int N = o->N;
#pragma omp parallel num_threads(o->threads) reduction(+ : bad)
  {
    int16_t* scratch = (int16_t*)malloc(sizeof(int16_t) * (2 * (size_t)N + (size_t)o->H * o->W));
#pragma omp for schedule(static)
    for (int e = 0; e < o->E; ++e) {
      size_t b = (size_t)e * N;
      bad += partial_step_env(o, e, actions + b, reward ? reward + e : 0, terminated ? terminated + e : 0,
                              agent_reward ? agent_reward + b : 0, avail ? avail + 5 * b : 0, scratch);
    }
    free(scratch);
  }
  return bad;
}

void oracle_partial_get(const oracle_env* o, uint8_t* at_goal, int32_t* goal_cost, int32_t* agent_steps,
                        int16_t* node, int16_t* edge) {
  size_t EN = (size_t)o->E * o->N;
  if (at_goal) memcpy(at_goal, o->at_goal, EN);
  if (goal_cost) memcpy(goal_cost, o->goal_cost, EN * 4);
  if (agent_steps) memcpy(agent_steps, o->agent_steps, EN * 4);
  if (node) memcpy(node, o->pnode, EN * 2);
  if (edge) memcpy(edge, o->pedge, EN * 2);
}

/* get_state, PARTIAL:384-393: [total collisions, step count, sum of per-agent goal costs]. */
void oracle_partial_state(const oracle_env* o, int64_t* out) {
  for (int e = 0; e < o->E; ++e) {
    int64_t sum = 0;
    for (int i = 0; i < o->N; ++i) sum += o->goal_cost[(size_t)e * o->N + i];
    out[3 * e] = o->total_coll[e];
    out[3 * e + 1] = o->step_count[e];
    out[3 * e + 2] = sum;
  }
}

/* get_obs_agent, PARTIAL:319-382: W x W obstacle map, W x W agent-count map, K x 13 nearest-agent features. */
static void partial_observe_agent(const oracle_env* o, int e, int id, double* out) {
  const int H = o->H, W = o->W, N = o->N, Wn = o->pW, K = o->pK;
  const int16_t* st = o->state + (size_t)e * H * W;
  const int16_t* pos = o->pos + (size_t)e * N * 2;
  const int16_t* goal = o->goal + (size_t)e * N * 2;
  const int16_t* start = o->start + (size_t)e * N * 2;
  int tl0 = pos[2 * id] - Wn / 2, tl1 = pos[2 * id + 1] - Wn / 2;     /* :326-327 */
  double* obstacle_map = out;
  double* agents_map = out + Wn * Wn;
  double* knn = out + 2 * Wn * Wn;
  for (int k = 0; k < 2 * Wn * Wn; ++k) out[k] = 0.0;
  for (int i = tl0; i < tl0 + Wn; ++i)                                 /* :332-342 */
    for (int j = tl1; j < tl1 + Wn; ++j) {
      int w = (i - tl0) * Wn + (j - tl1);
      if (!(0 <= i && i < H && 0 <= j && j < W)) { obstacle_map[w] = 1.0; continue; }
      if (st[i * W + j] == -1) obstacle_map[w] = 1.0;
      else if (st[i * W + j] > 0) agents_map[w] = (double)st[i * W + j];
    }
  for (int k = 0; k < K * 13; ++k) knn[k] = -1.0;                      /* :346 */
  /* agent_distance_matrix row (:548-567) and the stable sort by distance (:351-352) */
  double dist[256];
  int order[256];
  for (int j = 0; j < N; ++j) {
    if (j == id) dist[j] = (double)(H * W);
    else {
      int dx = pos[2 * id] - pos[2 * j], dy = pos[2 * id + 1] - pos[2 * j + 1];
      dist[j] = sqrt((double)(dx * dx + dy * dy));
    }
    order[j] = j;
  }
  for (int a = 1; a < N; ++a) {                                         /* insertion sort == stable */
    int v = order[a], b = a - 1;
    while (b >= 0 && dist[order[b]] > dist[v]) { order[b + 1] = order[b]; --b; }
    order[b + 1] = v;
  }
  int k_m1 = (N < K ? N : K) - 1;
  for (int row = 0; row <= k_m1; ++row) {                               /* :353-371 */
    int na = row == 0 ? id : order[row - 1];
    int dx = goal[2 * na] - pos[2 * na], dy = goal[2 * na + 1] - pos[2 * na + 1];
    double norm = sqrt((double)(dx * dx + dy * dy));                    /* __update_goal_vectors :957-972 */
    double ux = 0.0, uy = 0.0;
    if (norm != 0) { ux = (double)dx / norm; uy = (double)dy / norm; }
    double* f = knn + 13 * row;
    f[0] = pos[2 * na]; f[1] = pos[2 * na + 1];
    f[2] = start[2 * na]; f[3] = start[2 * na + 1];
    f[4] = goal[2 * na]; f[5] = goal[2 * na + 1];
    f[6] = ux; f[7] = uy; f[8] = norm;
    f[9] = o->pnode[(size_t)e * N + na]; f[10] = o->pedge[(size_t)e * N + na];
    f[11] = dist[na];
    f[12] = o->agent_steps[(size_t)e * N + na];
  }
}

void oracle_partial_observe(const oracle_env* o, double* obs) {
  const int N = o->N;
  const size_t osz = (size_t)2 * o->pW * o->pW + (size_t)13 * o->pK;
#pragma omp parallel for schedule(static) num_threads(o->threads)
  for (int e = 0; e < o->E; ++e)
    for (int i = 0; i < N; ++i) partial_observe_agent(o, e, i, obs + ((size_t)e * N + i) * osz);
}
