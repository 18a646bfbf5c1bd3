// mapf_kernels.cu -- sm_100a kernels of the batched MAPF step/observation engine.
//
// Reference semantics reproduced here (paths relative to the reference root):
//   GRID   = mapf_gridworld.py          step :85-141, avail :203-224, collisions :344-383
//   PRIMAL = mapf_primal.py             moveAgent :103-135, _step :549-637, _observe :343-386,
//                                       _listNextValidActions :639-667, getAstarCosts :407-499
//   PARTIAL= MARL-curve-main/src/envs/marl_partial.py   goal distance maps :931-955
//
// Design:
//   * A thread block owns a TILE of `epb` consecutive environments.  Their whole state
//     (padded obstacle bitmap, positions, goals, actions, occupancy grid) is staged in
//     shared memory once; step, collision handling, reward/done, action masks and the
//     observation are all produced from that copy (one launch, state read from HBM once).
//   * Sequential semantics: one warp per environment.  The agent-order dependent part of
//     PRIMAL's sweep (the occupancy claim) is a short serial loop over a shared-memory id
//     grid; everything that does not depend on agent order runs lane-parallel.
//   * Field-of-view observation: maps are kept as bit rows padded by F/2 on every side, so an
//     F-cell window row is one funnel shift.  Each agent assembles its 4*F*F observation bits
//     in registers, the bit strings of G consecutive agents are concatenated in shared memory,
//     and the whole tile is expanded to bytes (or floats) with perfectly coalesced 16-byte
//     streaming stores: thread q writes bytes [16q, 16q+16) of the tile's output.
//   * No tensor cores: nothing here is a contraction.  The path is bound by HBM writes.
#include <cuda_runtime.h>
#include <stdint.h>

#include "mapf_internal.h"

#ifdef MAPF_PHASE_TIMING
__device__ long long g_phase_clk[32];
// (a rollout records its second-to-last step: the last one also flushes state and statistics)
#define PHASE_MARK(i) do { if (blockIdx.x == gridDim.x / 2 && threadIdx.x == 0 && (A.T < 2 || t_roll + 2 == A.T)) g_phase_clk[i] = clock64(); } while (0)
#else
#define PHASE_MARK(i) do { } while (0)
#endif

namespace {

constexpr int kThreads = MAPF_TILE_THREADS;
// PARTIAL observation: agents per unrolled pass of the element-per-thread output loops.  Eight loads in flight per
// thread instead of four (float32) / one (the sector-aligned float64 loops): 181 -> 157 us and 254 -> 233 us for the
// c3-shaped batch; 16 is slower again (registers).
constexpr int kPartialUnroll = 8, kPartialUnroll64 = 8;
static_assert(kThreads % 32 == 0 && kThreads >= 32 && kThreads <= 1024, "tile block size");

// pre-status codes used inside the PRIMAL sweep (outside the reference's {-3..2})
constexpr int8_t PRE_SKIP = 10, PRE_STAY = 11, PRE_MOVE = 12, PRE_MOVED = 13;

struct Smem {
  uint32_t* obst;
  uint32_t* agt;
  uint8_t* grida;
  uint8_t* gridb;
  uchar2 *posold, *posnew, *goal;
  uint32_t* mv;      // [epb*N] padded target cell of the agent's move (0xffffffff: none)
  uint8_t* res;      // [epb*N] outcome of the move: RES_*
  uint8_t* dep;      // [epb*N] lower-id agent whose outcome decides this one
  uint8_t *act, *done, *flag, *avail, *nextmid, *node, *edge, *isint;
  int8_t* status;
  double* rew;
  double* envrew;    // PRIMAL: per-environment partial sums of the pairwise team reward (one per 32 agents)
  uint8_t* envterm;
  uint8_t* atgoal;
  uchar2 *pastold, *pastnew;     // diagonal mode: agents_past before / after the sweep
  uint16_t *mask16, *nextmid16;  // diagonal mode: 9-wide action masks
  uint32_t* str;
};

__device__ __forceinline__ Smem carve(unsigned char* base, const MapfTileLayout& L) {
  Smem s;
  s.obst = (uint32_t*)(base + L.obst_off);
  s.agt = (uint32_t*)(base + L.agt_off);
  s.grida = base + L.grida_off;
  s.gridb = base + L.gridb_off;
  s.posold = (uchar2*)(base + L.posold_off);
  s.posnew = (uchar2*)(base + L.posnew_off);
  s.goal = (uchar2*)(base + L.goal_off);
  s.mv = (uint32_t*)(base + L.mv_off);
  s.res = base + L.res_off;
  s.dep = base + L.dep_off;
  s.act = base + L.act_off;
  s.status = (int8_t*)(base + L.status_off);
  s.done = base + L.done_off;
  s.flag = base + L.flag_off;
  s.avail = base + L.avail_off;
  s.nextmid = base + L.nextmid_off;
  s.node = base + L.node_off;
  s.edge = base + L.edge_off;
  s.isint = base + L.isint_off;
  s.rew = (double*)(base + L.rew_off);
  s.envrew = (double*)(base + L.envrew_off);
  s.envterm = base + L.envterm_off;
  s.atgoal = base + L.atgoal_off;
  s.pastold = (uchar2*)(base + L.pastold_off);
  s.pastnew = (uchar2*)(base + L.pastnew_off);
  s.mask16 = (uint16_t*)(base + L.mask16_off);
  s.nextmid16 = (uint16_t*)(base + L.nextmid16_off);
  s.str = (uint32_t*)(base + L.str_off);
  return s;
}

// j / N for j < 65536 with inv = ceil(2^32 / N) (exact for N <= 255): one IMAD.HI instead of an integer division.
__device__ __forceinline__ int fast_div(int j, uint32_t inv) {
  return inv ? (int)__umulhi((uint32_t)j, inv) : j;   // inv == 0 encodes a divisor of 1
}

// Bit test in a padded bitmap; (r, c) are map coordinates and may lie up to P cells outside.
__device__ __forceinline__ uint32_t bm_test(const uint32_t* bm, int RW, int P, int r, int c) {
  const int pr = r + P, pc = c + P;
  return (bm[pr * RW + (pc >> 5)] >> (pc & 31)) & 1u;
}

// F consecutive bits of padded row `prow` starting at padded column `bit0`.
__device__ __forceinline__ uint32_t row_field(const uint32_t* bm, int RW, int prow, int bit0, uint32_t fmask) {
  const uint32_t* r = bm + prow * RW + (bit0 >> 5);
  return __funnelshift_r(r[0], r[1], bit0) & fmask;
}

__device__ __forceinline__ void byte_inc(uint8_t* grid, int cell) {
  atomicAdd((unsigned int*)(grid + (cell & ~3)), 1u << (8 * (cell & 3)));
}

// 16-byte streaming store (the observation is written once and never re-read by this kernel).
__device__ __forceinline__ void st_stream16(void* p, uint4 v) {
  asm volatile("st.global.cs.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
               : "memory");
}

// 4 bits -> 4 bytes of 0/1 (bit k lands in byte k).
__device__ __forceinline__ uint32_t expand4(uint32_t nib) { return (nib * 0x00204081u) & 0x01010101u; }

// PRIMAL dirDict (PRIMAL:28) for the 9 actions {0:(0,0) 1:(0,1) 2:(1,0) 3:(0,-1) 4:(-1,0) 5:(1,1) 6:(1,-1) 7:(-1,-1)
// 8:(-1,1)}, two bits per action holding delta + 1; and opposite_actions (PRIMAL:26), four bits per action (0 = none).
__device__ __forceinline__ int dir9_dx(int a) { return (int)((0x002865u >> (2 * a)) & 3u) - 1; }
__device__ __forceinline__ int dir9_dy(int a) { return (int)((0x020919u >> (2 * a)) & 3u) - 1; }
__device__ __forceinline__ int opposite9(int a) { return (int)((0x658721430ull >> (4 * a)) & 15ull); }

// Expands per-agent n-bit masks (uint16) to the [na][nact] uint8 layout (diagonal mode: nact = 9).
__device__ __noinline__ void write_mask_n(uint8_t* dst, const uint16_t* mask, int na, int nact, int tid) {
  const int n = nact * na;
  for (int idx = tid; idx < n; idx += kThreads) {
    const int ag = idx / nact;
    dst[idx] = (mask[ag] >> (idx - nact * ag)) & 1u;
  }
}

// Expands per-agent 5-bit masks to the [na][5] uint8 layout of get_avail_actions.
__device__ __noinline__ void write_mask5(uint8_t* dst, const uint8_t* mask, int na, int tid) {
  const int n = 5 * na;
  if ((((uintptr_t)dst) & 3) == 0) {
    const int nw = n >> 2;
    for (int q = tid; q < nw; q += kThreads) {
      uint32_t w = 0;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int idx = 4 * q + k;
        const int ag = idx / 5;
        w |= ((mask[ag] >> (idx - 5 * ag)) & 1u) << (8 * k);
      }
      ((uint32_t*)dst)[q] = w;
    }
    for (int idx = (nw << 2) + tid; idx < n; idx += kThreads) {
      const int ag = idx / 5;
      dst[idx] = (mask[ag] >> (idx - 5 * ag)) & 1u;
    }
  } else {
    for (int idx = tid; idx < n; idx += kThreads) {
      const int ag = idx / 5;
      dst[idx] = (mask[ag] >> (idx - 5 * ag)) & 1u;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// `sum(rewards)` (GRID:141) exactly as the interpreter that runs the reference evaluates it.
//   sum_mode 0: left fold (CPython <= 3.11).
//   sum_mode 1: CPython >= 3.12 builtin sum: exact while the items are ints, float(i) + x at the first
//   float, then Neumaier-compensated adds for floats / plain adds for ints, compensation added last.
// All operations use explicit round-to-nearest intrinsics so that no FMA contraction can occur.
// ------------------------------------------------------------------------------------------------
// The loop is branch-free: one thread folds a whole environment while the rest of its warp waits, so what counts is
// the dependent chain per item.  Every candidate is computed and selected (the compensation term is a chain of its
// own that trails the accumulator), the loads of four items are issued together: one dependent double add per item
// instead of three plus two shared-memory round trips and a divergent branch (GRID phase D: 12 400 -> ~2 000 cycles
// for 32 agents, it was 60 % of the tile's critical path).  Same operations on the same operands in the same order.
__device__ double py_sum(const double* x, const uint8_t* is_int, int n, int sum_mode) {
  double acc = 0.0;
  if (sum_mode == 0) {
#pragma unroll 4
    for (int i = 0; i < n; ++i) acc = __dadd_rn(acc, x[i]);
    return acc;
  }
  double c = 0.0;
  bool in_float = false;
#pragma unroll 4
  for (int i = 0; i < n; ++i) {
    const double xi = x[i];
    const bool flt = is_int[i] == 0;
    const double t = __dadd_rn(acc, xi);
    const bool big = fabs(acc) >= fabs(xi);
    const double hi = big ? acc : xi, lo = big ? xi : acc;
    const double cn = __dadd_rn(c, __dadd_rn(__dsub_rn(hi, t), lo));
    c = (in_float && flt) ? cn : c;          // compensated only between floats; the first float is a plain add
    in_float = in_float || flt;
    acc = t;
  }
  if (in_float && c != 0.0 && isfinite(c)) acc = __dadd_rn(acc, c);
  return acc;
}

// The same sum split over the tile (sum_mode 1), for the thread-per-environment phase D of GRID / PARTIAL, where one
// warp folds the environments of a tile while every other warp of every resident tile competes for the same issue
// slots (under load the fold above was 10 000 of a GRID tile's 18 000 cycles).  The accumulator chain needs one add
// per item; the compensation term of item i is a function of (accumulator before i, item i) alone, so the chain
// thread only records the accumulators (py_sum_head), every agent's thread computes its own term in parallel
// (py_sum_term), and the chain thread folds the terms in order (py_sum_tail).  A term that the serial loop would not
// have added is stored as +0.0: c starts as +0.0 and can never become -0.0, so c + 0.0 leaves c's bits alone.
__device__ __forceinline__ double py_sum_head(const double* __restrict__ x, double* __restrict__ pre, int n) {
  double acc = 0.0;
#pragma unroll 4
  for (int i = 0; i < n; ++i) {
    pre[i] = acc;
    acc = __dadd_rn(acc, x[i]);
  }
  return acc;
}
__device__ __forceinline__ double py_sum_term(double acc, double xi, bool compensated) {
  if (!compensated) return 0.0;
  const double t = __dadd_rn(acc, xi);
  const bool big = fabs(acc) >= fabs(xi);
  return __dadd_rn(__dsub_rn(big ? acc : xi, t), big ? xi : acc);
}
__device__ __forceinline__ double py_sum_tail(double acc, const double* term, int n, bool any_float) {
  double c = 0.0;
#pragma unroll 4
  for (int i = 0; i < n; ++i) c = __dadd_rn(c, term[i]);
  if (any_float && c != 0.0 && isfinite(c)) acc = __dadd_rn(acc, c);
  return acc;
}

// ------------------------------------------------------------------------------------------------
// PRIMAL team reward (mapf_step_out.reward_dev; MAPFEnv._step only returns per-agent rewards, so the team value has no
// reference counterpart and its summation order is this library's to define): the PAIRWISE sum of the per-agent
// rewards -- leaves x[0..N) (agents outside the swept range count as +0.0), padded with +0.0 to the next power of two,
// y[i] += y[i + s] for s = 1, 2, 4, ...  A fixed binary tree: log2(N) dependent double adds instead of the N of a left
// fold (one thread folding the 128 rewards of a c4 environment was 20 % of that tile's critical path), and the same
// bits on any tile shape.  TreeSum evaluates that tree from left to right with a binary-counter stack.
// ------------------------------------------------------------------------------------------------
struct TreeSum {
  double st[9];      // st[k]: sum of a finished subtree of 2^k leaves waiting for its right sibling
  unsigned n = 0;
  __device__ __forceinline__ void push(double v) {
    unsigned idx = n++;
#pragma unroll
    for (int k = 0; k < 9; ++k) {
      if (idx & 1u) {
        v = __dadd_rn(st[k], v);
        idx >>= 1;
      } else {
        st[k] = v;
        break;
      }
    }
  }
  __device__ __forceinline__ double finish() {
    if (n == 0) return 0.0;
    while (n & (n - 1)) push(0.0);
    double r = st[0];
#pragma unroll
    for (int k = 1; k < 9; ++k)   // compile-time indices only: the stack stays in registers
      if (n == (1u << k)) r = st[k];
    return r;
  }
};

// The same tree over the cnt <= 8 block sums p[0..cnt) of one environment (cnt = N / 32), a few live registers.
__device__ __forceinline__ double tree_sum8(const double* p, int cnt) {
  auto ld = [&](int k) { return k < cnt ? p[k] : 0.0; };
  double tot = p[0];
  if (cnt > 1) {
    tot = __dadd_rn(tot, p[1]);
    if (cnt > 2) {
      tot = __dadd_rn(tot, __dadd_rn(p[2], ld(3)));
      if (cnt > 4) tot = __dadd_rn(tot, __dadd_rn(__dadd_rn(p[4], ld(5)), __dadd_rn(ld(6), ld(7))));
    }
  }
  return tot;
}

// Any N: one thread walks the agents of its environment (kept out of line: its stack of partial sums must not cost the
// hot instantiations registers or a stack frame).
__device__ __noinline__ double tree_sum_agents(const double* rew, int N, int lo, int hi) {
  TreeSum ts;
  for (int i = 0; i < N; ++i) ts.push((i >= lo && i < hi) ? rew[i] : 0.0);
  return ts.finish();
}

// ------------------------------------------------------------------------------------------------
// Step phases.  Both modes share the structure
//   A  thread per agent : occupancy of the current positions + everything that does not depend on
//                         the other agents (direction, bounds / wall checks, GRID done latch)
//   B  (PRIMAL only) LANE per environment: the agent-order dependent occupancy claim
//   C  thread per agent : outcome, reward, collisions, statistics
//   D  thread per environment: done flag, team reward, step counter
// ------------------------------------------------------------------------------------------------

// Occupancy grids are padded by one cell on every side (stride GS = W + 2): the four neighbours of any map cell
// can be read without bounds checks, the padding is never occupied.
__device__ __forceinline__ int gcell(const MapfDims& d, int r, int c) { return (r + 1) * d.GS + c + 1; }

constexpr uint8_t RES_UNRESOLVED = 0, RES_MOVED = 1, RES_STAYS = 2;
constexpr uint32_t kCanary = 0xC0FFEE00u;   // guard words between shared-memory regions (MAPF_FLAG_INTERNAL)

// PRIMAL phase A for agent j: State.moveAgent's checks that do not involve other robots (PRIMAL:107-118).
// mv[j] = padded target cell (0xffffffff: no claim to make); status[j] = pre-status.
template <bool DIAG>
__device__ __forceinline__ void primal_phase_a(const MapfDims& d, const Smem& s, const MapfTileArgs& A, int j, int el,
                                               int a) {
  const uint32_t* ob = s.obst + (d.shared_map ? 0 : el * d.bm_words);
  const int act = s.act[j];
  const uchar2 p = s.posold[j];
  uint32_t tc = 0xffffffffu;
  int8_t st;
  if (a < A.agent_lo || a >= A.agent_hi) {
    st = PRE_SKIP;
  } else if (act == 0) {
    st = PRE_STAY;
  } else {
    const int t0 = (int)p.x + (DIAG ? dir9_dx(act) : (act == 2 ? 1 : (act == 4 ? -1 : 0)));  // dirDict, PRIMAL:28
    const int t1 = (int)p.y + (DIAG ? dir9_dy(act) : (act == 1 ? 1 : (act == 3 ? -1 : 0)));
    if (bm_test(ob, d.RW, d.P, t0, t1)) {
      st = (t0 < 0 || t0 >= d.H || t1 < 0 || t1 >= d.W) ? -1 : -2;   // PRIMAL:114-118
    } else {
      st = PRE_MOVE;
      tc = (uint32_t)gcell(d, t0, t1);
    }
  }
  s.status[j] = st;
  s.mv[j] = tc;
}

// Diagonal mode, State.diagonalCollision (PRIMAL:77-100): does the midpoint of the move (lx,ly) -> (nx,ny) of agent a
// equal the midpoint of another agent's last recorded move?  (np.isclose on half-integers == integer sums equal.)
__device__ __forceinline__ bool diagonal_collision(const uchar2* past, const uchar2* pos, int N, int a, int lx, int ly,
                                                   int nx, int ny) {
  for (int k = 0; k < N; ++k) {
    if (k == a) continue;
    const uchar2 q = past[k], r = pos[k];
    if ((int)q.x + r.x == lx + nx && (int)q.y + r.y == ly + ny) return true;
  }
  return false;
}

// Diagonal mode, phase B: the sweep itself, one lane per environment.  The crossing rule looks at other agents' last
// RECORDED move (agents_past is only updated when an agent acts, PRIMAL:109, 128), which makes the outcome depend on
// the running state of the whole sweep; the mode is off by default in the reference, so it simply walks the agents in
// id order against the live grid, live positions (posnew) and live agents_past (pastnew).
__device__ __forceinline__ void primal_phase_b_diag(const MapfDims& d, const Smem& s, const MapfTileArgs& A, int el) {
  const int N = d.N, jb = el * N;
  uint8_t* grid = s.grida + el * d.grid_bytes;
  uchar2* pos = s.posnew + jb;
  uchar2* past = s.pastnew + jb;
  for (int i = A.agent_lo; i < A.agent_hi; ++i) {
    const int j = jb + i;
    const int8_t st = s.status[j];
    const uchar2 p = pos[i];
    if (st == PRE_STAY) {
      past[i] = p;                                                   // PRIMAL:109
    } else if (st == PRE_MOVE) {
      const int tc = (int)s.mv[j];
      const int act = s.act[j];
      const int nx = (int)p.x + dir9_dx(act), ny = (int)p.y + dir9_dy(act);
      if (grid[tc] != 0 || diagonal_collision(past, pos, N, i, p.x, p.y, nx, ny)) {
        s.status[j] = -3;                                            // PRIMAL:119-124
      } else {
        grid[gcell(d, p.x, p.y)] = 0;                                // PRIMAL:126-129
        grid[tc] = (uint8_t)(i + 1);
        past[i] = p;
        pos[i] = make_uchar2((unsigned char)nx, (unsigned char)ny);
        s.status[j] = PRE_MOVED;
      }
    }
  }
}

// PRIMAL phase B: the outcome of the ordered sweep `for id in 1..N: moveAgent(id)` (PRIMAL:119-129) WITHOUT walking
// the agents one by one.  Agent a's move succeeds iff its target cell t is free when a's turn comes.  With k = the
// agent standing on t before the sweep (if any) and "claimants" = agents whose target is t (they sit on t's four
// neighbours), the ordered semantics reduce to
//     k > a                       : k has not moved yet                                        -> blocked
//     a claimant b with k < b < a : b found t free earlier (t was empty, or k had left) and took it -> blocked
//     otherwise, t empty (no k)   : nobody could have entered before a                          -> moves
//     otherwise, k < a            : a moves iff k itself moved away (k's own outcome, a lower id) -> depends on k
// (claimants below k always fail: k is still there at their turn).  Only the last case is order dependent, and only
// on a LOWER id, so the dependencies resolve in rounds; a round is one barrier, and random traffic needs one or two.
// Equivalence to the serial walk is checked bit-for-bit against the reference traces (tests/test_gpu_parity.py).
__device__ __forceinline__ uint8_t primal_classify(const MapfDims& d, const Smem& s, int j, int el, int a) {
  if (s.status[j] != PRE_MOVE) return RES_STAYS;
  const uint8_t* grid = s.grida + el * d.grid_bytes;   // ids before the sweep
  const uint32_t* mv = s.mv + el * d.N;
  const int8_t* st = s.status + el * d.N;
  const int tc = (int)s.mv[j];
  const int occ = grid[tc];
  if (occ > a + 1) return RES_STAYS;
  const int k = occ - 1;
  bool taken = false;
  const int nb[4] = {tc - 1, tc + 1, tc - d.GS, tc + d.GS};
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int b = (int)grid[nb[q]] - 1;
    if (b > k && b < a && st[b] == PRE_MOVE && (int)mv[b] == tc) taken = true;
  }
  if (taken) return RES_STAYS;
  if (k < 0) return RES_MOVED;
  s.dep[j] = (uint8_t)k;
  return RES_UNRESOLVED;
}

// PRIMAL phase C for agent j: final status (PRIMAL:108-110, 130-135), reward table (:579-597), on_goal (:633).
template <bool DIAG>
__device__ __forceinline__ bool primal_phase_c(const MapfDims& d, const Smem& s, int j, int el, int a,
                                               unsigned int& c_env, unsigned int& c_rob, unsigned int& c_arr) {
  const uchar2 po = s.posold[j], g = s.goal[j];
  int st = s.status[j];
  const bool swept = st != PRE_SKIP;
  uchar2 pn = po;
  const int act = s.act[j];
  if (DIAG) {
    pn = s.posnew[j];                                                // the serial sweep already moved the agent
  } else if (st == PRE_MOVE) {
    if (s.res[j] == RES_MOVED) {
      pn = make_uchar2((unsigned char)((int)po.x + (act == 2 ? 1 : (act == 4 ? -1 : 0))),
                       (unsigned char)((int)po.y + (act == 1 ? 1 : (act == 3 ? -1 : 0))));
      s.grida[el * d.grid_bytes + s.mv[j]] = (uint8_t)(a + 1);       // old cells were cleared before this phase
      st = PRE_MOVED;
    } else {
      st = -3;                                                       // collide with robot, PRIMAL:119-120
    }
  }
  const bool on_old = (po.x == g.x && po.y == g.y);
  const bool on_new = (pn.x == g.x && pn.y == g.y);
  if (st == PRE_STAY) st = on_old ? 1 : 0;
  else if (st == PRE_MOVED) st = on_new ? 1 : (on_old ? 2 : 0);
  else if (st == PRE_SKIP) st = 0;
  double r = 0.0;
  if (swept) {
    if (act == 0) r = (st == 1) ? __dadd_rn(d.goal_reward, -0.0) : d.idle_cost;
    else r = (st == 1) ? d.goal_reward : (st < 0 ? d.collision_reward : d.action_cost);
    c_env += (st == -1 || st == -2);
    c_rob += (st == -3);
    c_arr += (on_new && !on_old);
  }
  s.posnew[j] = pn;
  s.status[j] = (int8_t)st;
  s.rew[j] = r;
  s.done[j] = on_new ? 1 : 0;
  s.flag[j] = (uint8_t)((on_new ? 1 : 0) | ((st >= 0) ? 2 : 0));     // valid_action, PRIMAL:571
  return on_new;
}

// `done` and `nextActions` exactly as the i-th _step call of the sweep returned them (PRIMAL:626-630): at that
// moment agents < i stand on their new cells and agents > i still on their old ones.  One warp per environment.
// gridb holds the pre-sweep id grid.  Only used when the caller asks for the mid-sweep outputs.
__device__ void primal_mid_outputs(const MapfDims& d, const Smem& s, const MapfTileArgs& A, int ne, int tid) {
  const int N = d.N, lane = tid & 31, warp = tid >> 5;
  for (int el = warp; el < ne; el += kThreads / 32) {
    const int jb = el * N;
    const uint8_t* grid = s.grida + el * d.grid_bytes;
    const uint8_t* gold = s.gridb + el * d.grid_bytes;
    const uint32_t* ob = s.obst + (d.shared_map ? 0 : el * d.bm_words);
    int tot_old = 0;
    for (int a0 = 0; a0 < N; a0 += 32) {
      const int a = a0 + lane;
      bool on_old = false;
      if (a < N) {
        const uchar2 p = s.posold[jb + a], g = s.goal[jb + a];
        on_old = (p.x == g.x && p.y == g.y);
      }
      tot_old += __popc(__ballot_sync(0xffffffffu, on_old));
    }
    int new_prefix = 0, old_prefix = 0;
    for (int a0 = 0; a0 < N; a0 += 32) {
      const int a = a0 + lane;
      const bool active = a < N;
      const int j = jb + (active ? a : 0);
      const uchar2 po = s.posold[j], pn = s.posnew[j], g = s.goal[j];
      const bool on_old = active && (po.x == g.x && po.y == g.y);
      const bool on_new = active && (pn.x == g.x && pn.y == g.y);
      const unsigned bn = __ballot_sync(0xffffffffu, on_new);
      const unsigned bo = __ballot_sync(0xffffffffu, on_old);
      if (active) {
        const bool swept = (a >= A.agent_lo && a < A.agent_hi);
        const unsigned le = (lane == 31) ? 0xffffffffu : ((2u << lane) - 1u);
        const int cnt = new_prefix + __popc(bn & le) + (tot_old - old_prefix - __popc(bo & le));
        if (cnt == N) s.flag[j] |= 4;
        uint8_t m = 1;
#pragma unroll
        for (int k = 1; k <= 4; ++k) {
          const int n0 = (int)pn.x + (k == 2 ? 1 : (k == 4 ? -1 : 0));
          const int n1 = (int)pn.y + (k == 1 ? 1 : (k == 3 ? -1 : 0));
          if (!bm_test(ob, d.RW, d.P, n0, n1)) {
            const int c = gcell(d, n0, n1);
            const int idn = grid[c], ido = gold[c];
            const bool occ = (ido > a + 1) || (idn != 0 && idn < a + 1);
            if (!occ) m |= (uint8_t)(1u << k);
          }
        }
        const int act = s.act[j];
        const int opp = (act == 0) ? -1 : (((act + 1) & 3) + 1);     // opposite_actions, PRIMAL:26
        if (opp > 0) m &= (uint8_t)~(1u << opp);
        s.nextmid[j] = swept ? m : 0;
      }
      new_prefix += __popc(bn);
      old_prefix += __popc(bo);
    }
  }
}

// Diagonal mode: the 9-wide masks.  `ahead` = index of the sweeping agent at the moment the mask is evaluated: agents
// below it are taken from their post-sweep records (posnew / pastnew), the others from the pre-sweep ones; ahead = N
// gives the state after the whole sweep.  _listNextValidActions, PRIMAL:639-667 with the crossing test :658-660.
__device__ __forceinline__ bool diagonal_collision_at(const Smem& s, int jb, int N, int a, int ahead, int sx, int sy) {
  for (int k = 0; k < N; ++k) {
    if (k == a) continue;
    const uchar2 q = (k < ahead) ? s.pastnew[jb + k] : s.pastold[jb + k];
    const uchar2 r = (k < ahead) ? s.posnew[jb + k] : s.posold[jb + k];
    if ((int)q.x + r.x == sx && (int)q.y + r.y == sy) return true;
  }
  return false;
}

// Mid-sweep `done` and `nextActions` of the diagonal mode (PRIMAL:626-630), one thread per agent.
__device__ void primal_mid_outputs_diag(const MapfDims& d, const Smem& s, const MapfTileArgs& A, int ne, int tid) {
  const int N = d.N, na = ne * N;
  for (int j = tid; j < na; j += kThreads) {
    const int el = fast_div(j, d.invN), a = j - el * N, jb = el * N;
    const uint8_t* grid = s.grida + el * d.grid_bytes;
    const uint8_t* gold = s.gridb + el * d.grid_bytes;
    const uint32_t* ob = s.obst + (d.shared_map ? 0 : el * d.bm_words);
    int cnt = 0;
    for (int k = 0; k < N; ++k) {
      const uchar2 q = (k <= a) ? s.posnew[jb + k] : s.posold[jb + k], g = s.goal[jb + k];
      cnt += (q.x == g.x && q.y == g.y) ? 1 : 0;
    }
    if (cnt == N) s.flag[j] |= 4;
    const uchar2 pn = s.posnew[j];
    uint32_t m = 1;
    for (int k = 1; k <= 8; ++k) {
      const int n0 = (int)pn.x + dir9_dx(k), n1 = (int)pn.y + dir9_dy(k);
      if (bm_test(ob, d.RW, d.P, n0, n1)) continue;
      const int c = gcell(d, n0, n1);
      const int idn = grid[c], ido = gold[c];
      if ((ido > a + 1) || (idn != 0 && idn < a + 1)) continue;
      if (diagonal_collision_at(s, jb, N, a, a + 1, (int)pn.x + n0, (int)pn.y + n1)) continue;
      m |= 1u << k;
    }
    const int opp = opposite9(s.act[j]);
    if (opp > 0) m &= ~(1u << opp);
    const bool swept = (a >= A.agent_lo && a < A.agent_hi);
    s.nextmid16[j] = (uint16_t)(swept ? m : 0);
  }
}

// __is_cell_obstacle of GRID (GRID:278) and PARTIAL (PARTIAL:521): `_full_obs[cell] == -1`.  _full_obs is -1 on walls
// PLUS the number of agents on the cell (GRID:299), so a wall cell that holds an agent -- the reference's .scen x/y
// transposition produces such starts -- is NOT an obstacle: neighbours may enter it and the action masks open up.
// `cnt` is the count grid the reference would be looking at (pre-step inside step(), rebuilt for the masks).
__device__ __forceinline__ bool count_obstacle(const MapfDims& d, const uint32_t* ob, const uint8_t* cnt, int r, int c) {
  return bm_test(ob, d.RW, d.P, r, c) && cnt[gcell(d, r, c)] == 0;
}

// GRID phase A for agent j (GRID:99-118): move unless wall / border, done latch, new-position counts.
__device__ __forceinline__ void grid_phase_a(const MapfDims& d, const Smem& s, int j, int el, int step_now,
                                             unsigned int& c_env, unsigned int& c_arr) {
  const uint32_t* ob = s.obst + (d.shared_map ? 0 : el * d.bm_words);
  const uchar2 p = s.posold[j], g = s.goal[j];
  const bool done_old = s.done[j] != 0;
  uchar2 np = p;
  double r = 0.0;
  int flag = 0;
  if (!done_old) {
    const int act = s.act[j];
    if (act < 4) {                                                   // __agent_step, GRID:319-342
      const int t0 = (int)p.x + (act == 0 ? -1 : (act == 1 ? 1 : 0));
      const int t1 = (int)p.y + (act == 2 ? -1 : (act == 3 ? 1 : 0));
      if (count_obstacle(d, ob, s.grida + el * d.grid_bytes, t0, t1)) flag = 1;
      else np = make_uchar2((unsigned char)t0, (unsigned char)t1);
    }
    if (flag) r = __dadd_rn(r, d.collide_reward);                    // GRID:105-106
    r = __dadd_rn(r, d.step_reward);                                 // GRID:110
  }
  bool dn = done_old;
  const bool reached = (np.x == g.x && np.y == g.y);
  if (reached) dn = true;                                            // GRID:112-113
  if (step_now >= d.episode_limit) dn = true;                        // GRID:116-117
  s.posnew[j] = np;
  s.status[j] = (int8_t)flag;
  s.done[j] = dn ? 1 : 0;
  s.rew[j] = r;
  s.isint[j] = (uint8_t)(done_old ? d.collide_is_int : (d.collide_is_int && d.step_is_int));
  byte_inc(s.gridb + el * d.grid_bytes, gcell(d, np.x, np.y));
  c_env += flag;
  c_arr += (reached && !done_old);
}

// PARTIAL phase A for agent j (PARTIAL:192-235): move unless wall / border, action cost, at_goal, limit latch and the
// goal-distance shaping term from the BFS maps.
__device__ __forceinline__ void partial_phase_a(const MapfDims& d, const Smem& s, const MapfState& S, size_t gj, int j,
                                                int el, int step_now, unsigned int& c_env, unsigned int& c_arr) {
  const uint32_t* ob = s.obst + (d.shared_map ? 0 : el * d.bm_words);
  const uchar2 p = s.posold[j], g = s.goal[j];
  const bool done_old = s.done[j] != 0;
  const bool at_old = s.atgoal[j] != 0;
  // the global loads of this phase are issued before anything waits for one: the agent's step counter and the goal
  // distance of the old cell now, the distance of the new cell as soon as the move is decided (they were three exposed
  // memory latencies in a row: 12 000 cycles of the tile's critical path)
  const int16_t* dm = S.goal_dist + gj * d.HW;                       // :229-234
  const int opd = __ldg(dm + (int)p.x * d.W + p.y);
  int steps_old = 0;
  if (!done_old) steps_old = S.agent_steps[gj];
  uchar2 np = p;
  double r = 0.0;
  int flag = 0;
  const int act = s.act[j];
  if (!done_old && act < 4) {                                        // __agent_step, :618-643
    const int t0 = (int)p.x + (act == 0 ? -1 : (act == 1 ? 1 : 0));
    const int t1 = (int)p.y + (act == 2 ? -1 : (act == 3 ? 1 : 0));
    if (count_obstacle(d, ob, s.grida + el * d.grid_bytes, t0, t1)) flag = 1;
    else np = make_uchar2((unsigned char)t0, (unsigned char)t1);
  }
  const int npd = (np.x == p.x && np.y == p.y) ? opd : (int)__ldg(dm + (int)np.x * d.W + np.y);
  if (!done_old) {
    S.agent_steps[gj] = steps_old + 1;                               // _agent_step_count, :194
    if (flag) r = __dadd_rn(r, d.p_envc);                            // :203-205
    if (act < 4) r = __dadd_rn(r, d.p_move);                         // :207-208
    else r = __dadd_rn(r, at_old ? d.p_stay_goal : d.p_stay);        // :209-213
  }
  const bool at_new = (np.x == g.x && np.y == g.y);                  // :217-222
  if (at_new) S.goal_cost[gj] = step_now;
  const bool dn = done_old || (step_now >= d.episode_limit);         // :224-227
  // (old - new) / episode_limit: a step changes the hop distance by -1, 0 or +1, and IEEE division is sign-symmetric:
  // +-inv_limit or +0.0 without the division routine; anything else (unreachable-cell sentinels) divides
  const int diff = opd - npd;
  double closer = diff == 0 ? 0.0 : (diff == 1 ? d.inv_limit : -d.inv_limit);
  if (diff < -1 || diff > 1) closer = __ddiv_rn((double)diff, (double)d.episode_limit);
  r = __dadd_rn(r, closer);
  s.posnew[j] = np;
  s.status[j] = (int8_t)flag;
  s.done[j] = dn ? 1 : 0;
  s.atgoal[j] = at_new ? 1 : 0;
  s.rew[j] = r;
  s.isint[j] = 0;                                                    // closer_rew is always a float
  byte_inc(s.gridb + el * d.grid_bytes, gcell(d, np.x, np.y));
  c_env += flag;
  c_arr += (at_new && !at_old);
}

// GRID / PARTIAL phase C for agent j: node flag (GRID:344-362, PARTIAL:713-737), edge count (GRID:364-383,
// PARTIAL:822-857), collision penalties (GRID:127-130, PARTIAL:255-259).
__device__ __forceinline__ bool grid_phase_c(const MapfDims& d, const Smem& s, int j, int el, int a, bool partial,
                                             int* envcnt2, unsigned int& c_node, unsigned int& c_edge) {
  const int N = d.N, jb = el * N;
  const uint8_t* cold = s.grida + el * d.grid_bytes;
  const uint8_t* cnew = s.gridb + el * d.grid_bytes;
  const uchar2 p = s.posold[j], np = s.posnew[j];
  const int nc = gcell(d, np.x, np.y);
  const int node = cnew[nc] > 1 ? 1 : 0;
  int edge = 0;
  // a swap partner stood on my new cell and stands on my old one now: both count grids must be non-zero there (I
  // moved, so I am not the one on my old cell).  Without the second test a third of all moves walked the N agents.
  if ((p.x != np.x || p.y != np.y) && cold[nc] != 0 && cnew[gcell(d, p.x, p.y)] != 0) {
    for (int k = 0; k < N; ++k) {
      if (k == a) continue;
      const uchar2 qo = s.posold[jb + k], qn = s.posnew[jb + k];
      edge += (qo.x == np.x && qo.y == np.y && qn.x == p.x && qn.y == p.y) ? 1 : 0;
    }
  }
  double r = s.rew[j];
  r = __dadd_rn(r, __dmul_rn(partial ? d.p_nc : d.collide_reward, (double)node));
  r = __dadd_rn(r, __dmul_rn(partial ? d.p_ec : d.collide_reward, (double)edge));
  s.rew[j] = r;
  s.node[j] = (uint8_t)node;
  s.edge[j] = (uint8_t)edge;
  c_node += node;
  c_edge += edge;
  if (partial) {
    if (node + edge) atomicAdd(&envcnt2[el], node + edge);           // _total_number_collisions, PARTIAL:250
    return s.atgoal[j] != 0;
  }
  return s.done[j] != 0;
}

// ------------------------------------------------------------------------------------------------
// Field-of-view observation (PRIMAL _observe :343-386), F known at compile time.
// Bit index == byte index of the [4][F][F] output: channel*F*F + wi*F + wj, channels in the returned
// order [poss_map, goal_map, goals_map, obs_map] (:386).
// ------------------------------------------------------------------------------------------------
template <int F>
struct Fov {
  static constexpr int FF = F * F;
  static constexpr int NB = 4 * FF;
  static constexpr int NW = (NB + 31) / 32;
  static constexpr int CW = (FF + 31) / 32;
  static constexpr uint32_t FMASK = (1u << F) - 1u;
  // goal_map / goals_map bits never fall into a string word shared with a neighbouring agent iff channel 0
  // and channel 3 each cover a whole word
  static constexpr bool kInterior = FF >= 32;
};

template <int NWORDS>
__device__ __forceinline__ void or_field(uint32_t (&w)[NWORDS], int off, int bits, uint32_t v) {
  const int k = off >> 5, sh = off & 31;
  w[k] |= v << sh;
  if (sh + bits > 32) w[k + 1] |= v >> (32 - sh);
}

// poss_map (channel 0) and obs_map (channel 3): 2 x F window rows, each one funnel shift of a padded bit row.
template <int F>
__device__ __forceinline__ void fov_window_planes(uint32_t (&w)[Fov<F>::NW], const MapfDims& d, const uint32_t* ob,
                                                  const uint32_t* ag, uchar2 p) {
  using T = Fov<F>;
#pragma unroll
  for (int q = 0; q < T::NW; ++q) w[q] = 0u;
#pragma unroll
  for (int wi = 0; wi < F; ++wi) {
    // padded row p.x + wi holds map row p.x - P + wi; padded column p.y holds map column p.y - P
    const uint32_t fo = row_field(ob, d.RW, (int)p.x + wi, (int)p.y, T::FMASK);   // walls + out of bounds, :356-362
    const uint32_t fa = row_field(ag, d.RW, (int)p.x + wi, (int)p.y, T::FMASK);   // any agent, :363-372
    or_field<T::NW>(w, 0 * T::FF + wi * F, F, fa);
    or_field<T::NW>(w, 3 * T::FF + wi * F, F, fo);
  }
}

// goal_map (own goal cell, :366-368) and goals_map (goals of the visible other agents clamped into the window,
// :374-378): single bits OR-ed into the agent's string in shared memory.  `vis` = poss_map without the agent itself.
template <int F, bool ATOMIC>
__device__ __forceinline__ void fov_goal_bits(uint32_t* str, int bit0, uint32_t (&vis)[Fov<F>::CW], int GS,
                                              const uint8_t* idgrid, const uchar2* goals_env, uchar2 p, uchar2 g) {
  using T = Fov<F>;
  constexpr int P = F / 2;
  auto set = [&](int pos) {
    const int b = bit0 + pos;
    if (ATOMIC) atomicOr(&str[b >> 5], 1u << (b & 31));
    else str[b >> 5] |= 1u << (b & 31);
  };
  const int t0 = (int)p.x - P, t1 = (int)p.y - P;
  const int gi = (int)g.x - t0, gj = (int)g.y - t1;
  if ((unsigned)gi < (unsigned)F && (unsigned)gj < (unsigned)F) set(T::FF + gi * F + gj);
  const uint8_t* gbase = idgrid + (t0 + 1) * GS + t1 + 1;
  // One loop over ALL window words: a warp runs max-over-lanes(visible agents) iterations instead of the sum of the
  // per-word maxima (the visible set is sparse: ~3 agents spread over CW words).
  for (;;) {
    uint32_t v = vis[0];
    int q = 0;
#pragma unroll
    for (int k = 1; k < T::CW; ++k) {
      const bool next = (v == 0);
      v = next ? vis[k] : v;
      q = next ? k : q;
    }
    if (v == 0) break;
    const int idx = 32 * q + __ffs(v) - 1;
    v &= v - 1;
#pragma unroll
    for (int k = 0; k < T::CW; ++k)
      if (q == k) vis[k] = v;
    const int wi = (int)((unsigned)idx / (unsigned)F), wj = idx - wi * F;
    const int id = gbase[wi * GS + wj];            // the agent bit map says somebody stands here: id >= 1
    const uchar2 og = goals_env[id - 1];
    const int ci = min(max((int)og.x - t0, 0), F - 1);
    const int cj = min(max((int)og.y - t1, 0), F - 1);
    set(2 * T::FF + ci * F + cj);
  }
}

// goal_map and goals_map for ODD F in a single-pass tile: the windows are symmetric (j sees b iff b sees j), so every
// visible pair is enumerated ONCE -- by the agent that finds the other one in the FORWARD half of its window (cells
// after the centre in row-major order) -- and both agents' bits are set.  Half the cells to scan, half the divergent
// iterations per warp.  All sets are shared-memory atomics on the tile's flat bit string (agent j's bit i is bit
// j * 4F^2 + i), issued after every thread has stored its window planes.
template <int F>
__device__ __forceinline__ void fov_goal_bits_half(uint32_t* str, int j, int jenv0, const uint32_t (&vis)[Fov<F>::CW],
                                                   int GS, const uint8_t* idgrid, const uchar2* goals_tile, uchar2 p,
                                                   uchar2 g) {
  using T = Fov<F>;
  constexpr int P = F / 2;
  constexpr int START = P * F + P + 1;          // first cell after the centre
  constexpr int NFWD = T::FF - START;           // cells of the forward half
  constexpr int Q0 = START >> 5, SH = START & 31, CWF = (NFWD + 31) / 32;
  auto set = [&](int flat) { atomicOr(&str[flat >> 5], 1u << (flat & 31)); };
  const int t0 = (int)p.x - P, t1 = (int)p.y - P;
  const int gi = (int)g.x - t0, gj = (int)g.y - t1;
  if ((unsigned)gi < (unsigned)F && (unsigned)gj < (unsigned)F) set(j * T::NB + T::FF + gi * F + gj);
  uint32_t f[CWF];
#pragma unroll
  for (int k = 0; k < CWF; ++k) {
    const uint32_t lo = vis[Q0 + k];
    const uint32_t hi = (Q0 + k + 1 < T::CW) ? vis[Q0 + k + 1] : 0u;
    f[k] = SH ? __funnelshift_r(lo, hi, SH) : lo;
  }
  const uint8_t* gbase = idgrid + (t0 + 1) * GS + t1 + 1;
  for (;;) {
    uint32_t v = f[0];
    int q = 0;
#pragma unroll
    for (int k = 1; k < CWF; ++k) {
      const bool next = (v == 0);
      v = next ? f[k] : v;
      q = next ? k : q;
    }
    if (v == 0) break;
    const int idx = START + 32 * q + __ffs(v) - 1;
    v &= v - 1;
#pragma unroll
    for (int k = 0; k < CWF; ++k)
      if (q == k) f[k] = v;
    const int wi = (int)((unsigned)idx / (unsigned)F), wj = idx - wi * F;
    const int jb = jenv0 + gbase[wi * GS + wj] - 1;        // the agent standing there (its index in the tile)
    const uchar2 og = goals_tile[jb];
    // its goal, clamped into my window (PRIMAL:374-378)
    const int ci = min(max((int)og.x - t0, 0), F - 1);
    const int cj = min(max((int)og.y - t1, 0), F - 1);
    set(j * T::NB + 2 * T::FF + ci * F + cj);
    // my goal, clamped into its window (origin = its position - P)
    const int u0 = t0 + wi - P, u1 = t1 + wj - P;
    const int di = min(max((int)g.x - u0, 0), F - 1);
    const int dj = min(max((int)g.y - u1, 0), F - 1);
    set(jb * T::NB + 2 * T::FF + di * F + dj);
  }
}

// ------------------------------------------------------------------------------------------------
// The tile kernel: stage -> [step] -> state write-back + small outputs -> [observation].
// ------------------------------------------------------------------------------------------------
// SINGLE: the tile holds at most kThreads agents (the host guarantees it), so every per-agent / per-environment loop is
// one guarded pass -- no loop counters, compares and back edges around each phase (worth 7 % of the c3 step).
// ROLL (mapf_rollout; needs SINGLE): A.T consecutive steps in ONE launch.  The tile stays resident in shared memory and
// in the registers of the threads that own its agents: the obstacle rows are staged once, positions / done flags / step
// counters never make the round trip through global memory between steps, and step t+1's actions are fetched while
// step t is still being computed.  Step t's outputs go to element offset t * (size of one step's output) of every
// output pointer (time-major [T, E, ...] storage); the handle's state and statistics are written after the last step.
// WIDE (rollouts only): 64 registers per thread instead of 48.  A batch small enough to be resident all at once (c2) is
// bound by the dependent instruction chain of a step, not by occupancy: without the register cap the chain has no
// spills and a better schedule (7.1 -> 6.7 us per c2 step); large batches keep the 10-blocks-per-SM variant.
template <int F, int MODE, bool SINGLE, bool ROLL, bool WIDE = false>
__global__ void __launch_bounds__(kThreads, ROLL ? (WIDE ? 8 : 10) : 12)
mapf_tile_kernel(const MapfDims d, const MapfTileLayout L, const MapfState S, const MapfTileArgs A) {
  static_assert(!WIDE || ROLL, "the wide-register build exists for rollouts");
  static_assert(!ROLL || SINGLE, "rollouts keep the agents in the registers of their threads");
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ unsigned int stat[MAPF_N_STATS];
  __shared__ unsigned int bad_flag;
  const Smem s = carve(smem_raw, L);
  int* envcnt = (int*)(smem_raw + L.envcnt_off);
  const int tid = threadIdx.x, lane = tid & 31;
  const int N = d.N;
  const int e0 = blockIdx.x * d.epb;
  const int ne = min(d.epb, d.E - e0);
  const int na = ne * N;
  const size_t a0 = (size_t)e0 * N;
  // the mode is a template parameter: the other modes' code is not even in this kernel's instruction stream
  constexpr bool diag = MODE == MAPF_MODE_PRIMAL_DIAG;   // PRIMAL with DIAGONAL_MOVEMENT (9 actions)
  constexpr bool primal = MODE == MAPF_MODE_PRIMAL || diag;
  constexpr bool partial = MODE == MAPF_MODE_PARTIAL;
  int* envcnt2 = (int*)(smem_raw + L.envcnt2_off);
  int* envff = (int*)(smem_raw + L.envff_off);      // GRID / PARTIAL only (aliases the scratch otherwise: never touched)
  double* sumpre = (double*)(smem_raw + L.pre_off);
  const bool do_step = A.do_step != 0;
  const bool need_mid = primal && do_step && ((A.out.done_mid_dev != nullptr) || (A.out.next_mid_dev != nullptr));
  // The fused step+observation launch of the 5-action PRIMAL mode does not need a pass of its own for the agent bit map
  // and the action masks: phase C sets the agent bit of the cell it just decided, and the masks fall out of the window
  // rows the observation reads anyway (a neighbour cell is free iff neither its wall nor its agent bit is set).  One
  // pass over the agents and one block-wide barrier less per step.
  const bool fused_avail = MODE == MAPF_MODE_PRIMAL && SINGLE && F >= 3 && do_step && A.obs != nullptr &&
                           A.out.avail_dev != nullptr;
  int my_act = 0;   // fused_avail: the agent's action, kept across the barrier behind which the scratch bytes are reused

  // ---- stage the tile (one exposed global-memory latency): obstacle bitmaps, per-agent records, zeroed grids
  if (tid < MAPF_N_STATS) stat[tid] = 0;
  if (tid == 0) bad_flag = 0;
  if (tid < 4) *(uint32_t*)(smem_raw + L.guard_off[tid]) = kCanary + tid;   // verified before the kernel exits
  int* envstep = (int*)(smem_raw + L.envstep_off);
  // Every global load of the tile is issued before anything waits on one, and the zero fill runs while they are in
  // flight: ONE exposed memory latency per tile (load -> store -> load -> store chains cost two or three).
  struct AgentRec {
    uchar2 p, g, past;
    uint8_t dn, pv, atg;
    long long av;
  };
  size_t a0t = a0;   // first agent of the tile in the outputs of the current step (a0 + t * E * N in a rollout)
  size_t e0t = (size_t)e0;
  int t_roll = 0;
  auto load_action = [&](int j, size_t at) -> long long {
    return (A.act_dtype == MAPF_I64) ? ((const long long*)A.actions)[at + j]
                                     : (long long)((const uint8_t*)A.actions)[at + j];
  };
  auto load_rec = [&](int j, AgentRec& r) {
    r.p = ((const uchar2*)S.pos)[a0 + j];
    r.g = ((const uchar2*)S.goal)[a0 + j];
    r.dn = S.done[a0 + j];
    r.pv = S.prev_action[a0 + j];
    r.atg = 0;
    if (partial) r.atg = S.at_goal[a0 + j];
    r.past = make_uchar2(0, 0);
    if (diag) r.past = ((const uchar2*)S.past)[a0 + j];
    r.av = -1;
    if (do_step) {
      const int el = fast_div(j, d.invN), a = j - el * N;
      if (a >= A.agent_lo && a < A.agent_hi) r.av = load_action(j, a0);
    }
  };
  bool bad = false;
  auto store_rec = [&](int j, const AgentRec& r) {
    s.posold[j] = r.p;
    s.posnew[j] = r.p;
    s.goal[j] = r.g;
    s.done[j] = r.dn;
    if (partial) s.atgoal[j] = r.atg;
    if (diag) {
      s.pastold[j] = r.past;
      s.pastnew[j] = r.past;
    }
    int act = r.pv;
    if (do_step) {
      const int el = fast_div(j, d.invN), a = j - el * N;
      if (a >= A.agent_lo && a < A.agent_hi) {
        long long v = r.av;
        if (v < 0 || v > (diag ? 8 : 4)) {                           // GRID:92 / PRIMAL:556 assert
          bad = true;
          v = primal ? 0 : 4;
        }
        act = (int)v;
      }
    }
    s.act[j] = (uint8_t)act;
  };
  const int nvec = ((d.shared_map ? 1 : ne) * d.bm_words) >> 2;
  const uint4* osrc = (const uint4*)(S.obst_bits + (d.shared_map ? (size_t)0 : (size_t)e0 * d.bm_words));
  uint4* odst = (uint4*)s.obst;
  uint4 ob0 = make_uint4(0, 0, 0, 0);
  if (tid < nvec) ob0 = __ldg(osrc + tid);
  AgentRec r0;
  if (tid < na) load_rec(tid, r0);
  int sc0 = 0;
  if (tid < ne) sc0 = S.step_count[e0 + tid];
  const bool incr = ROLL && MODE == MAPF_MODE_PRIMAL && fused_avail;
  unsigned int c0 = 0, c1 = 0, c2 = 0, c3 = 0;   // statistics of this thread's agents (a rollout flushes them once)
  unsigned int ep_done = 0;                      // thread per environment: steps that ended with everybody on goal
  for (;;) {   // one iteration per step of a rollout; a single pass otherwise
  PHASE_MARK(10);
  // Rollouts: `last` = the step after which state and statistics go back to global memory (every step otherwise).
  // `fresh` = the occupancy grid and the agent bit rows are rebuilt from the positions; false in the later steps of a
  // PRIMAL rollout, where both are simply what the previous step left behind: the sweep vacates and enters cells in
  // the id grid anyway, and the agent bits are cleared where an agent left (before the barrier of phase B) and set
  // where it arrived (phase C).  That saves the zero fill, the rebuild and one block-wide barrier per step.
  const bool last = !ROLL || t_roll + 1 >= A.T;
  const bool fresh = !(incr && t_roll > 0);
  bad = false;
  if (fresh) {
    // zero the agent bit rows and the occupancy grid(s): they are adjacent in the tile layout, one loop clears them
    const uint4 z = make_uint4(0, 0, 0, 0);
    const int nz = (L.grida_off - L.agt_off + d.epb * d.grid_bytes * ((!primal && do_step) ? 2 : 1)) >> 4;
    uint4* zp = (uint4*)(smem_raw + L.agt_off);
#pragma unroll 1
    for (int i = tid; i < nz; i += kThreads) zp[i] = z;
  }
  if (!ROLL || t_roll == 0) {
    if (tid < nvec) odst[tid] = ob0;
#pragma unroll 1
    for (int i = tid + kThreads; i < nvec; i += kThreads) odst[i] = __ldg(osrc + i);
  }
  if (tid < ne) {
    envcnt[tid] = 0;
    envcnt2[tid] = 0;
    envstep[tid] = sc0;
    if (!primal) envff[tid] = N;
  }
  if (tid < na) store_rec(tid, r0);
  if (!SINGLE) {
    for (int el = tid + kThreads; el < ne; el += kThreads) {
      envcnt[el] = 0;
      envcnt2[el] = 0;
      envstep[el] = S.step_count[e0 + el];
      if (!primal) envff[el] = N;
    }
    for (int j = tid + kThreads; j < na; j += kThreads) {
      AgentRec r;
      load_rec(j, r);
      store_rec(j, r);
    }
  }
  if (fresh) __syncthreads();   // (not fresh: phase A below only touches the agent's own records)
  if (bad) bad_flag = 1;        // sticky for the whole launch
  long long next_av = -1;   // rollout: the action of the NEXT step, in flight while this step is computed
  if (ROLL && t_roll + 1 < A.T && tid < na) next_av = load_action(tid, a0t + (size_t)d.E * N);

  PHASE_MARK(1);
  // ---- phase A: occupancy of the current positions (PRIMAL State.state ids, PRIMAL:32-47; GRID agent counts,
  //      GRID:299) and the agent-independent part of the step
  for (int j = tid, it_ = 0; j < na && (!SINGLE || it_ == 0); j += kThreads, ++it_) {
    const int el = fast_div(j, d.invN), a = j - el * N;
    const uchar2 p = s.posold[j];
    uint8_t* grid = s.grida + el * d.grid_bytes;
    const int cell = gcell(d, p.x, p.y);
    if (primal) {
      if (fresh) grid[cell] = (uint8_t)(a + 1);
      if (do_step) primal_phase_a<diag>(d, s, A, j, el, a);
    } else {
      byte_inc(grid, cell);
    }
  }
  __syncthreads();
  if (!primal && do_step) {
    // GRID / PARTIAL moves test `_full_obs == -1` (count_obstacle), i.e. they read the finished pre-step count grid
    for (int j = tid, it_ = 0; j < na && (!SINGLE || it_ == 0); j += kThreads, ++it_) {
      const int el = fast_div(j, d.invN);
      const int step_now = envstep[el] + 1;                          // GRID:93, PARTIAL:178
      if (partial) partial_phase_a(d, s, S, a0 + j, j, el, step_now, c0, c3);
      else grid_phase_a(d, s, j, el, step_now, c0, c3);
    }
    __syncthreads();
  }

  if (do_step) {
  PHASE_MARK(2);
    // ---- phase B (PRIMAL): outcome of the ordered sweep, resolved in parallel (see primal_classify)
    if (primal) {
      if (need_mid)   // keep the pre-sweep id grid for the mid-sweep outputs
        for (int i = tid; i < ((ne * d.grid_bytes) >> 4); i += kThreads)
          ((uint4*)s.gridb)[i] = ((const uint4*)s.grida)[i];
      if constexpr (diag) {
        __syncthreads();                       // the copy above reads the grid the sweep is about to change
        for (int el = tid, it_ = 0; el < ne && (!SINGLE || it_ == 0); el += kThreads, ++it_) primal_phase_b_diag(d, s, A, el);
        __syncthreads();
      } else {
      bool pending = false;
      for (int j = tid, it_ = 0; j < na && (!SINGLE || it_ == 0); j += kThreads, ++it_) {
        const int el = fast_div(j, d.invN), a = j - el * N;
        const uint8_t r = primal_classify(d, s, j, el, a);
        s.res[j] = r;
        pending |= (r == RES_UNRESOLVED);
      }
      while (__syncthreads_or(pending)) {      // an agent waits only for a LOWER id: every round makes progress
        pending = false;
        for (int j = tid, it_ = 0; j < na && (!SINGLE || it_ == 0); j += kThreads, ++it_) {
          if (s.res[j] != RES_UNRESOLVED) continue;
          const int el = fast_div(j, d.invN);
          const int k = el * N + s.dep[j];
          const uint8_t rk = s.res[k];
          if (rk == RES_UNRESOLVED) {
            // pointer jumping: k's outcome is its own predecessor's, so wait for that one directly -- a convoy of
            // length L resolves in log2(L) rounds (any value read here is an agent further up the same chain)
            s.dep[j] = s.dep[k];
            pending = true;
          } else {
            s.res[j] = rk;                     // moves iff the agent ahead of it moved away
          }
        }
      }
      for (int j = tid, it_ = 0; j < na && (!SINGLE || it_ == 0); j += kThreads, ++it_)   // vacate the old cells; phase C enters the new ones
        if (s.res[j] == RES_MOVED) {
          const uchar2 p = s.posold[j];
          const int el = fast_div(j, d.invN);
          s.grida[el * d.grid_bytes + gcell(d, p.x, p.y)] = 0;
          if (F > 0 && incr) {   // the agent bit rows live across the steps of the rollout: the cell is left
            const int pc = (int)p.y + d.P;
            atomicAnd(&s.agt[el * d.bm_words + ((int)p.x + d.P) * d.RW + (pc >> 5)], ~(1u << (pc & 31)));
          }
        }
      __syncthreads();
      }
    }
  PHASE_MARK(3);
    // ---- phase C
    for (int j0 = 0; j0 < na && (!SINGLE || j0 == 0); j0 += kThreads) {
      const int j = j0 + tid;
      const bool active = j < na;
      int el = -1;
      bool flag = false;
      if (active) {
        el = fast_div(j, d.invN);
        const int a = j - el * N;
        flag = primal ? primal_phase_c<diag>(d, s, j, el, a, c0, c1, c3)
                      : grid_phase_c(d, s, j, el, a, partial, envcnt2, c1, c2);
        if (F > 0 && fused_avail) {
          const uchar2 pn = s.posnew[j];
          const int pc = (int)pn.y + d.P;
          atomicOr(&s.agt[el * d.bm_words + ((int)pn.x + d.P) * d.RW + (pc >> 5)], 1u << (pc & 31));
        }
      }
      if (primal && d.rsum_mode != 0 && A.out.reward_dev != nullptr) {
        // pairwise team reward, the part of the tree that lives inside a warp: segments of N | 32 agents, or the
        // N / 32 aligned blocks of a larger environment (see TreeSum)
        const int a = active ? j - el * N : 0;
        double v = (active && a >= A.agent_lo && a < A.agent_hi) ? s.rew[j] : 0.0;
        const int span = N < 32 ? N : 32;
        for (int sft = 1; sft < span; sft <<= 1) v = __dadd_rn(v, __shfl_down_sync(0xffffffffu, v, sft));
        if (active && (a & (span - 1)) == 0) s.envrew[d.rsum_mode == 1 ? el : el * (N >> 5) + (a >> 5)] = v;
      }
      // per-environment count of agents on goal (PRIMAL) / done (GRID), one shared-memory atomic per (warp, env)
      const unsigned peers = __match_any_sync(0xffffffffu, el);
      const unsigned bf = __ballot_sync(0xffffffffu, flag);
      if (active && lane == __ffs(peers) - 1) atomicAdd(&envcnt[el], __popc(bf & peers));
      if (!primal && d.sum_mode == 1) {
        // first float item of the environment's sum(rewards): consecutive lanes hold consecutive agents
        const unsigned fl = __ballot_sync(0xffffffffu, active && s.isint[j] == 0) & peers;
        if (active && fl != 0 && lane == __ffs(peers) - 1) atomicMin(&envff[el], j - el * N + __ffs(fl) - 1 - lane);
      }
    }
    if (d.collect_stats && last) {
      c0 = __reduce_add_sync(0xffffffffu, c0);
      c1 = __reduce_add_sync(0xffffffffu, c1);
      c2 = __reduce_add_sync(0xffffffffu, c2);
      c3 = __reduce_add_sync(0xffffffffu, c3);
      if (lane == 0) {
        if (c0) atomicAdd(&stat[MAPF_STAT_ENV_COLLISIONS], c0);
        if (c1) atomicAdd(&stat[MAPF_STAT_NODE_COLLISIONS], c1);
        if (c2) atomicAdd(&stat[MAPF_STAT_EDGE_COLLISIONS], c2);
        if (c3) atomicAdd(&stat[MAPF_STAT_GOAL_ARRIVALS], c3);
      }
    }
    if (primal) __syncthreads();   // phase C entered the new cells into the id grid
    if (need_mid) {
      if constexpr (diag) primal_mid_outputs_diag(d, s, A, ne, tid);
      else primal_mid_outputs(d, s, A, ne, tid);
    }
  }

  PHASE_MARK(4);
  // ---- agent bitmap of the post-step positions + available-action masks
  const uint8_t* gridcur = (!primal && do_step) ? s.gridb : s.grida;
  const bool want_avail = A.out.avail_dev != nullptr;
  for (int j = tid, it_ = 0; !fused_avail && j < na && (!SINGLE || it_ == 0); j += kThreads, ++it_) {
    const int el = fast_div(j, d.invN);
    const uchar2 p = s.posnew[j];
    if (F > 0) {
      uint32_t* ag = s.agt + el * d.bm_words;
      const int pc = (int)p.y + d.P;
      atomicOr(&ag[((int)p.x + d.P) * d.RW + (pc >> 5)], 1u << (pc & 31));
    }
    if (diag && want_avail) {                                        // 9-wide _listNextValidActions
      const uint32_t* ob = s.obst + (d.shared_map ? 0 : el * d.bm_words);
      const uint8_t* grid = gridcur + el * d.grid_bytes;
      const int a = j - el * N;
      uint32_t m = 1;
      for (int k = 1; k <= 8; ++k) {
        const int n0 = (int)p.x + dir9_dx(k), n1 = (int)p.y + dir9_dy(k);
        if (bm_test(ob, d.RW, d.P, n0, n1) || grid[gcell(d, n0, n1)] != 0) continue;
        if (diagonal_collision_at(s, el * N, N, a, N, (int)p.x + n0, (int)p.y + n1)) continue;
        m |= 1u << k;
      }
      const int opp = opposite9(s.act[j]);
      if (opp > 0) m &= ~(1u << opp);
      s.mask16[j] = (uint16_t)m;
    } else if (want_avail) {
      const uint32_t* ob = s.obst + (d.shared_map ? 0 : el * d.bm_words);
      uint8_t m;
      if (primal) {                                                  // _listNextValidActions, PRIMAL:639-667
        const uint8_t* grid = gridcur + el * d.grid_bytes;
        m = 1;
#pragma unroll
        for (int k = 1; k <= 4; ++k) {
          const int n0 = (int)p.x + (k == 2 ? 1 : (k == 4 ? -1 : 0));
          const int n1 = (int)p.y + (k == 1 ? 1 : (k == 3 ? -1 : 0));
          if (!bm_test(ob, d.RW, d.P, n0, n1) && grid[gcell(d, n0, n1)] == 0) m |= (uint8_t)(1u << k);
        }
        const int act = s.act[j];
        const int opp = (act == 0) ? -1 : (((act + 1) & 3) + 1);
        if (opp > 0) m &= (uint8_t)~(1u << opp);
      } else {                                                       // get_avail_agent_actions, GRID:203-224
        const uint8_t* grid = gridcur + el * d.grid_bytes;           // the rebuilt _full_obs counts, GRID:132-140
        m = 16;
        m |= count_obstacle(d, ob, grid, (int)p.x - 1, (int)p.y) ? 0 : 1;
        m |= count_obstacle(d, ob, grid, (int)p.x + 1, (int)p.y) ? 0 : 2;
        m |= count_obstacle(d, ob, grid, (int)p.x, (int)p.y - 1) ? 0 : 4;
        m |= count_obstacle(d, ob, grid, (int)p.x, (int)p.y + 1) ? 0 : 8;
      }
      // [E,N,5] mask straight from the register: 5 byte stores per agent (the L2 merges the partial sectors)
      uint8_t* o = A.out.avail_dev + 5 * (a0t + j);
      o[0] = m & 1;
      o[1] = (m >> 1) & 1;
      o[2] = (m >> 2) & 1;
      o[3] = (m >> 3) & 1;
      o[4] = (m >> 4) & 1;
    }
  }
  if (!fused_avail || need_mid) __syncthreads();   // (fused_avail: the barrier behind phase C already covers phase D)

  PHASE_MARK(5);
  // ---- phase D + state write-back + the small per-agent / per-env outputs (coalesced over the tile)
  if (do_step) {
    for (int el = tid, it_ = 0; el < ne && (!SINGLE || it_ == 0); el += kThreads, ++it_) {
      const bool all = envcnt[el] == N;        // PRIMAL State.done (:159-165) / GRID episode_done (:267)
      if (A.out.terminated_dev) A.out.terminated_dev[e0t + el] = all ? 1 : 0;
      if (primal) {
        if (A.out.reward_dev) {
          double tot;
          if (d.rsum_mode == 1) {
            tot = s.envrew[el];                                      // the whole tree was inside one warp
          } else if (d.rsum_mode == 2) {
            tot = tree_sum8(s.envrew + el * (N >> 5), N >> 5);
          } else {
            tot = tree_sum_agents(s.rew + el * N, N, A.agent_lo, A.agent_hi);
          }
          A.out.reward_dev[e0t + el] = tot;
        }
        if (A.agent_lo == 0 && last) S.step_count[e0 + el] = envstep[el] + 1;
      } else if (partial) {
        const int step_now = envstep[el] + 1;
        bool term = (S.terminated[e0 + el] != 0) || (step_now >= d.episode_limit);     // PARTIAL:224-226
        if (all) {                                                                     // all at goal, PARTIAL:291-299
          term = true;
          const double bonus = S.complete_lut[min(step_now, d.complete_len - 1)];
          for (int i = 0; i < N; ++i) {
            s.done[el * N + i] = 1;
            s.rew[el * N + i] = __dadd_rn(s.rew[el * N + i], bonus);
          }
        }
        if (A.out.reward_dev) {                                                        // sum(rewards), PARTIAL:310
          if (d.sum_mode == 1) s.envrew[el] = py_sum_head(s.rew + el * N, sumpre + el * N, N);
          else A.out.reward_dev[e0t + el] = py_sum(s.rew + el * N, s.isint + el * N, N, d.sum_mode);
        }
        if (A.out.terminated_dev) A.out.terminated_dev[e0t + el] = term ? 1 : 0;
        S.terminated[e0 + el] = term ? 1 : 0;
        S.total_coll[e0 + el] += envcnt2[el] / 2;                                      // PARTIAL:250
        S.step_count[e0 + el] = step_now;
      } else {
        if (A.out.reward_dev) {                                                        // sum(rewards), GRID:141
          if (d.sum_mode == 1) s.envrew[el] = py_sum_head(s.rew + el * N, sumpre + el * N, N);
          else A.out.reward_dev[e0t + el] = py_sum(s.rew + el * N, s.isint + el * N, N, d.sum_mode);
        }
        if (last) S.step_count[e0 + el] = envstep[el] + 1;
      }
      ep_done += all ? 1u : 0u;
      if (d.collect_stats && last && ep_done) atomicAdd(&S.stats[MAPF_STAT_EPISODES_DONE], (unsigned long long)ep_done);
    }
    PHASE_MARK(13);
    if (!primal && d.sum_mode == 1 && A.out.reward_dev != nullptr) {
      // CPython >= 3.12 sum(rewards): compensation terms by the agents' own threads, folded in order (see py_sum_head)
      __syncthreads();
      PHASE_MARK(14);
      for (int j = tid, it_ = 0; j < na && (!SINGLE || it_ == 0); j += kThreads, ++it_) {
        const int el = fast_div(j, d.invN), a = j - el * N;
        sumpre[j] = py_sum_term(sumpre[j], s.rew[j], s.isint[j] == 0 && a > envff[el]);
      }
      __syncthreads();
      PHASE_MARK(15);
      for (int el = tid, it_ = 0; el < ne && (!SINGLE || it_ == 0); el += kThreads, ++it_)
        A.out.reward_dev[e0t + el] = py_sum_tail(s.envrew[el], sumpre + el * N, N, envff[el] < N);
    }
    PHASE_MARK(11);
    if (d.collect_stats && last) {   // one global atomic per counter per tile (a rollout: once for all its steps)
      const unsigned long long nsteps = ROLL ? (unsigned long long)A.T : 1ull;
      if (tid == 0) {
        if (!primal || A.agent_lo == 0) atomicAdd(&S.stats[MAPF_STAT_ENV_STEPS], nsteps * ne);
        atomicAdd(&S.stats[MAPF_STAT_AGENT_STEPS], nsteps * (unsigned long long)(ne * (primal ? A.agent_hi - A.agent_lo : N)));
      } else if (tid < MAPF_N_STATS && stat[tid] != 0) {
        atomicAdd(&S.stats[tid], (unsigned long long)stat[tid]);
      }
    }
    PHASE_MARK(12);
    if (partial) __syncthreads();   // phase D changed done / rewards of whole environments
    // state write-back and the per-agent outputs: thread j stores agent j's records (byte stores of a warp cover
    // whole 32-byte sectors; a shared-memory staging pass costs more instructions than it saves transactions)
    for (int j = tid, it_ = 0; j < na && (!SINGLE || it_ == 0); j += kThreads, ++it_) {
      const size_t gj = a0 + j, gt = a0t + j;   // the handle's state / this step's outputs
      const uint8_t dn = s.done[j];
      const uchar2 pn = s.posnew[j];
      my_act = s.act[j];
      if (last) {   // the handle's state (a rollout keeps it in registers until its last step)
        ((uchar2*)S.pos)[gj] = pn;
        S.done[gj] = dn;
        S.prev_action[gj] = (uint8_t)my_act;
      }
      if (partial) {
        S.at_goal[gj] = s.atgoal[j];
        S.pnode[gj] = s.node[j];
        S.pedge[gj] = s.edge[j];
      }
      if (ROLL) {   // the agent's record for the next step stays in this thread's registers
        r0.p = pn;
        r0.dn = dn;
        if (partial) r0.atg = s.atgoal[j];
      }
      if (diag) ((uchar2*)S.past)[gj] = s.pastnew[j];
      if (A.out.dones_dev) A.out.dones_dev[gt] = dn;
      if (A.out.status_dev) A.out.status_dev[gt] = s.status[j];
      if (A.out.agent_reward_dev) A.out.agent_reward_dev[gt] = s.rew[j];
      if (A.out.node_dev) A.out.node_dev[gt] = primal ? 0 : (int16_t)s.node[j];
      if (A.out.edge_dev) A.out.edge_dev[gt] = primal ? 0 : (int16_t)s.edge[j];
      if (A.out.valid_dev) A.out.valid_dev[gt] = primal ? ((s.flag[j] >> 1) & 1) : 1;
      if (primal && A.out.done_mid_dev) A.out.done_mid_dev[gt] = (s.flag[j] >> 2) & 1;
    }
    if (primal && A.life_queue != nullptr) {
      // Lifelong hand-out (MAPF-490-main/Global.cpp:85-94), exactly mapf_pop_goals_kernel behind this step: an agent that
      // ends the step on its goal takes the front of its queue.  This step's outputs and observation keep the old
      // goal (they read the tile's copy); the handle's goal, its on-goal flag and the list of re-assigned
      // (env, agent) pairs for the BFS that follows change here.  A loop of its own, behind a uniform branch: inside
      // the write-back loop it cost the plain step 3 % (registers kept live across the stores).
      for (int j = tid; j < na; j += kThreads) {
        const size_t gj = a0 + j;
        const uchar2 pn = s.posnew[j], g = s.goal[j];
        if (pn.x != g.x || pn.y != g.y) continue;
        const int hd = A.life_head[gj];
        if (hd >= A.life_Q) continue;
        const short2 q = ((const short2*)A.life_queue)[gj * (size_t)A.life_Q + hd];
        int g0 = q.x, g1 = q.y;
        if (g0 < 0 || g0 >= d.H || g1 < 0 || g1 >= d.W) {
          atomicOr(S.err_flags, MAPF_FLAG_BAD_POSITION);
          g0 = min(max(g0, 0), d.H - 1);
          g1 = min(max(g1, 0), d.W - 1);
        }
        A.life_head[gj] = hd + 1;
        ((uchar2*)S.goal)[gj] = make_uchar2((unsigned char)g0, (unsigned char)g1);
        S.done[gj] = (uint8_t)(pn.x == g0 && pn.y == g1);
        A.life_list[atomicAdd(A.life_cnt, 1)] = (int32_t)gj;
      }
    }
    if constexpr (diag) {
      if (A.out.next_mid_dev) write_mask_n(A.out.next_mid_dev + 9 * a0t, s.nextmid16, na, 9, tid);
    } else {
      if (primal && A.out.next_mid_dev) write_mask5(A.out.next_mid_dev + 5 * a0t, s.nextmid, na, tid);
    }
  }
  if constexpr (diag) {
    if (want_avail) write_mask_n(A.out.avail_dev + 9 * a0t, s.mask16, na, 9, tid);
  }
  if (tid == 0 && bad_flag) atomicOr(S.err_flags, MAPF_FLAG_BAD_ACTION);

  PHASE_MARK(6);
  // ---- observation (a `break` leaves the observation of this step)
  do {
  if (A.obs == nullptr && A.vec == nullptr) break;

  if (d.obs_mode == MAPF_OBS_FULLMAP) {
    // get_obs / get_state, GRID:143-196: -1 on walls, else the number of agents on the cell.
    if (A.obs == nullptr) break;
    int8_t* out = (int8_t*)A.obs + e0t * d.HW;
    if ((d.W & 15) == 0) {
      // sixteen cells of a row per thread: one funnel shift for the wall bits, five aligned words of the count grid
      // shifted into place, one 16-byte store (a quarter of the instructions of the four-cell path below: the
      // full-map observation was 17 % of the GRID step's instructions)
      const int q16 = d.HW >> 4;
      int el = 0, q = tid;                                             // chunk q of environment el (no divisions)
      while (q >= q16) {
        q -= q16;
        ++el;
      }
      for (; el < ne;) {
        const int r = fast_div(q << 4, d.invW), c = (q << 4) - r * d.W;
        const uint32_t* ob = s.obst + (d.shared_map ? 0 : el * d.bm_words);
        const int g0 = el * d.grid_bytes + gcell(d, r, c);
        const uint32_t* gw = (const uint32_t*)(gridcur + (g0 & ~3));
        const int sh = (g0 & 3) << 3;
        const uint32_t w0 = gw[0], w1 = gw[1], w2 = gw[2], w3 = gw[3], w4 = gw[4];
        const uint32_t wall = row_field(ob, d.RW, r + d.P, c + d.P, 0xFFFFu);
        uint4 v;                                                       // `+= 1` on a -1 cell per agent, GRID:299
        v.x = __vsub4(__funnelshift_r(w0, w1, sh), expand4(wall & 15u));
        v.y = __vsub4(__funnelshift_r(w1, w2, sh), expand4((wall >> 4) & 15u));
        v.z = __vsub4(__funnelshift_r(w2, w3, sh), expand4((wall >> 8) & 15u));
        v.w = __vsub4(__funnelshift_r(w3, w4, sh), expand4(wall >> 12));
        *(uint4*)(out + (size_t)el * d.HW + ((size_t)q << 4)) = v;
        q += kThreads;
        while (q >= q16) {
          q -= q16;
          ++el;
        }
      }
      break;
    }
    if ((d.W & 3) == 0) {
      // four cells of a row per thread: one funnel shift for the wall bits, four count bytes, one packed 32-bit store
      const int q4 = d.HW >> 2;
      for (int el = 0; el < ne; ++el) {
        const uint32_t* ob = s.obst + (d.shared_map ? 0 : el * d.bm_words);
        const uint8_t* grid = gridcur + el * d.grid_bytes;
        uint32_t* o32 = (uint32_t*)(out + (size_t)el * d.HW);
        for (int q = tid; q < q4; q += kThreads) {
          const int cell = q << 2;
          const int r = fast_div(cell, d.invW), c = cell - r * d.W;
          const uint8_t* gp = grid + gcell(d, r, c);
          const uint32_t cnt = (uint32_t)gp[0] | ((uint32_t)gp[1] << 8) | ((uint32_t)gp[2] << 16) | ((uint32_t)gp[3] << 24);
          const uint32_t wall = expand4(row_field(ob, d.RW, r + d.P, c + d.P, 0xFu));
          o32[q] = __vsub4(cnt, wall);                               // `+= 1` on a -1 cell per agent, GRID:299
        }
      }
      break;
    }
    for (int i = tid; i < ne * d.HW; i += kThreads) {
      const int el = i / d.HW, cell = i - el * d.HW;
      const int r = cell / d.W, c = cell - r * d.W;
      const uint32_t* ob = s.obst + (d.shared_map ? 0 : el * d.bm_words);
      const int cnt = gridcur[el * d.grid_bytes + gcell(d, r, c)];
      out[i] = bm_test(ob, d.RW, d.P, r, c) ? (int8_t)(cnt - 1) : (int8_t)cnt;   // `+= 1` on a -1 cell, GRID:299
    }
    break;
  }

  if constexpr (F > 0) {
    using T = Fov<F>;
    PHASE_MARK(7);
    __syncthreads();   // the bit strings reuse the step-phase scratch: everybody has finished the write-back
    // phase 1: one thread per agent builds its 4*F*F bits and the goal vector; G agents share a
    // word-aligned group string.
    constexpr bool kHalf = SINGLE && (F & 1) != 0;   // symmetric windows + one pass: each visible pair is walked once
    for (int base = 0; base < na && (!SINGLE || base == 0); base += kThreads) {
      const int j = base + tid;
      const bool valid = (j < na) && (A.obs != nullptr);
      uint32_t first = 0;
      uint32_t vis[T::CW];
      int w0 = 0, sh = 0, el = 0;
      uchar2 p = make_uchar2(0, 0), g = make_uchar2(0, 0);
      const double2* vt = nullptr;
      if (j < na) {
        el = fast_div(j, d.invN);
        p = s.posnew[j];
        g = s.goal[j];
        if (A.vec != nullptr) {   // the goal-vector table entry is needed at the END of this pass: start fetching it now
          vt = (const double2*)S.vec_lut + 2 * (abs((int)g.x - (int)p.x) * d.W + abs((int)g.y - (int)p.y));
          asm volatile("prefetch.global.L1 [%0];" ::"l"(vt));
        }
      }
      if (valid) {
        uint32_t w[T::NW];
        fov_window_planes<F>(w, d, s.obst + (d.shared_map ? 0 : el * d.bm_words), s.agt + el * d.bm_words, p);
        if constexpr (F >= 3 && MODE == MAPF_MODE_PRIMAL) {
          if (fused_avail) {
            // _listNextValidActions (PRIMAL:639-667) from the window planes: the agent sits at window cell (P, P); a
            // direction is open iff the neighbour cell shows neither a wall / border (channel 3) nor an agent (channel 0)
            constexpr int Pw = F / 2;
            auto open_cell = [&](int wi, int wj) -> uint32_t {
              const int i0 = wi * F + wj, i3 = 3 * T::FF + wi * F + wj;          // compile-time bit positions
              return (((w[i0 >> 5] >> (i0 & 31)) | (w[i3 >> 5] >> (i3 & 31))) & 1u) ^ 1u;
            };
            uint32_t m = 1u | (open_cell(Pw, Pw + 1) << 1) | (open_cell(Pw + 1, Pw) << 2) |
                         (open_cell(Pw, Pw - 1) << 3) | (open_cell(Pw - 1, Pw) << 4);      // dirDict, PRIMAL:28
            const int opp = (my_act == 0) ? -1 : (((my_act + 1) & 3) + 1);                // opposite_actions, :26
            if (opp > 0) m &= ~(1u << opp);
            uint8_t* o = A.out.avail_dev + 5 * (a0t + j);
            o[0] = m & 1;
            o[1] = (m >> 1) & 1;
            o[2] = (m >> 2) & 1;
            o[3] = (m >> 3) & 1;
            o[4] = (m >> 4) & 1;
          }
        }
#pragma unroll
        for (int q = 0; q < T::CW; ++q) vis[q] = w[q];
        if ((T::FF & 31) != 0) vis[T::CW - 1] &= (1u << (T::FF & 31)) - 1u;
        vis[((F / 2) * F + F / 2) >> 5] &= ~(1u << (((F / 2) * F + F / 2) & 31));   // not the agent itself
        const int grp = j / d.G, k = j - grp * d.G;
        const int boff = k * T::NB;
        w0 = grp * d.GW + (boff >> 5);
        sh = boff & 31;
        const int nwords = (sh + T::NB + 31) >> 5;
        uint32_t prev = 0;
#pragma unroll
        for (int q = 0; q <= T::NW; ++q) {
          const uint32_t cur = (q < T::NW) ? w[q] : 0u;
          const uint32_t o = __funnelshift_l(prev, cur, sh);     // (cur:prev << sh) >> 32
          prev = cur;
          if (q == 0) first = o;                                 // word shared with the previous agent: OR-ed below
          if (q < nwords && !(q == 0 && sh > 0)) s.str[w0 + q] = o;
        }
        if (!kHalf && T::kInterior)
          fov_goal_bits<F, false>(s.str + w0, sh, vis, d.GS, gridcur + el * d.grid_bytes, s.goal + el * N, p, g);
      }
      if (j < na && A.vec != nullptr) {                              // PRIMAL:380-385
        const int dx = (int)g.x - (int)p.x, dy = (int)g.y - (int)p.y;
        const double2 u = __ldg(vt);                                 // |dx| / mag, |dy| / mag (IEEE division is
        const double2 m = __ldg(vt + 1);                             // sign-symmetric; 0 / mag = +0.0 either way)
        double* v = A.vec + 3 * (a0t + j);
        v[0] = dx < 0 ? -u.x : u.x;
        v[1] = dy < 0 ? -u.y : u.y;
        v[2] = m.x;
      }
      __syncthreads();
      if (valid) {
        if (kHalf) {
          if (sh > 0) atomicOr(&s.str[w0], first);
          fov_goal_bits_half<F>(s.str, j, el * N, vis, d.GS, gridcur + el * d.grid_bytes, s.goal, p, g);
        } else if (T::kInterior) {
          if (sh > 0) s.str[w0] |= first;
        } else {
          if (sh > 0) atomicOr(&s.str[w0], first);
          fov_goal_bits<F, true>(s.str + w0, sh, vis, d.GS, gridcur + el * d.grid_bytes, s.goal + el * N, p, g);
        }
      }
    }
    __syncthreads();
    if (A.obs == nullptr) break;

  PHASE_MARK(8);
    // phase 2: expand the tile's bit string; thread q writes output bytes [16q, 16q+16)
    const size_t nbits = (size_t)na * T::NB;
    if (A.obs_dtype == MAPF_BITS) {
      // the strings as they are: tile t starts at word a0 * NB / 32 (the host only selects this output when every
      // tile holds whole groups, so the tile's first bit sits on a word boundary)
      uint32_t* out = (uint32_t*)A.obs + ((a0t * T::NB) >> 5);
      const int nw = (int)((nbits + 31) >> 5);
      for (int q = tid; q < nw; q += kThreads) out[q] = s.str[q];
    } else if (A.obs_dtype == MAPF_U8) {
      uint8_t* out = (uint8_t*)A.obs + a0t * T::NB;
      const int lead = (int)((16 - ((uintptr_t)out & 15)) & 15);   // 0 unless the tile holds an odd number of groups
      if (lead == 0) {
        const int nchunk = (int)(nbits >> 4);
        const uint16_t* s16 = (const uint16_t*)s.str;
#pragma unroll 2
        for (int q = tid; q < nchunk; q += kThreads) {
          const uint32_t h = s16[q];
          uint4 v;
          v.x = expand4(h & 15u);
          v.y = expand4((h >> 4) & 15u);
          v.z = expand4((h >> 8) & 15u);
          v.w = expand4(h >> 12);
          st_stream16(out + ((size_t)q << 4), v);
        }
        for (int b = (nchunk << 4) + tid; b < (int)nbits; b += kThreads) out[b] = (s.str[b >> 5] >> (b & 31)) & 1u;
      } else {
        // tile start not 16-byte aligned (environments-per-tile had to drop below the aligned multiple to fit shared
        // memory): bytes up to the first boundary one by one, then aligned chunks taken at a 4-bit-aligned string offset
        const int first = min(lead, (int)nbits);
        for (int b = tid; b < first; b += kThreads) out[b] = (s.str[b >> 5] >> (b & 31)) & 1u;
        const int nchunk = ((int)nbits - first) >> 4;
        for (int q = tid; q < nchunk; q += kThreads) {
          const int bit = first + (q << 4);
          const uint32_t h = __funnelshift_r(s.str[bit >> 5], s.str[(bit >> 5) + 1], bit) & 0xffffu;
          uint4 v;
          v.x = expand4(h & 15u);
          v.y = expand4((h >> 4) & 15u);
          v.z = expand4((h >> 8) & 15u);
          v.w = expand4(h >> 12);
          st_stream16(out + first + ((size_t)q << 4), v);
        }
        for (int b = first + (nchunk << 4) + tid; b < (int)nbits; b += kThreads)
          out[b] = (s.str[b >> 5] >> (b & 31)) & 1u;
      }
    } else {
      float* out = (float*)A.obs + a0t * T::NB;
      const int nchunk = (int)(nbits >> 2);
      for (int q = tid; q < nchunk; q += kThreads) {
        const uint32_t nib = (s.str[q >> 3] >> ((q & 7) << 2)) & 15u;
        uint4 v;
        v.x = (nib & 1u) ? 0x3f800000u : 0u;
        v.y = (nib & 2u) ? 0x3f800000u : 0u;
        v.z = (nib & 4u) ? 0x3f800000u : 0u;
        v.w = (nib & 8u) ? 0x3f800000u : 0u;
        st_stream16(out + ((size_t)q << 2), v);
      }
    }
  }
  } while (0);
  PHASE_MARK(9);
  if (!ROLL || ++t_roll >= A.T) break;
  // ---- next step of the rollout: outputs move on by one step, the agents' records are already in registers
  a0t += (size_t)d.E * N;
  e0t += (size_t)d.E;
  r0.av = next_av;
  sc0 += 1;
  __syncthreads();   // the bit strings alias the step scratch the next staging pass is about to write
  }
  // nobody wrote outside its region of the tile: the canaries between the regions are intact
  if (A.debug_corrupt && tid == 0) smem_raw[L.guard_off[(A.debug_corrupt - 1) & 3]] ^= 0xff;   // self-test hook
  __syncthreads();
  if (tid < 4 && *(volatile uint32_t*)(smem_raw + L.guard_off[tid]) != kCanary + tid) atomicOr(S.err_flags, MAPF_FLAG_INTERNAL);
}

// ------------------------------------------------------------------------------------------------
// mapf_pipe_kernel<F>: mapf_rollout for SMALL PRIMAL batches -- a two-stage pipeline inside the block.
//
// A batch like c2 (4096 envs x 8 agents) is resident on the GPU all at once, and in mapf_tile_kernel<.., ROLL> every
// tile then walks ONE dependent chain of ~11 000 cycles per step (step phases ~5 900, observation phases ~4 700): there
// is no second tile to hide it behind, issue slots are half empty, HBM a third used.  Here a tile is at most 32 agents
// (one thread per agent) and its block of four warps splits into two ROLES that run concurrently:
//   warp 0      STEP role: the sweep of step t+1 (phases A-D of the tile kernel, barriers are __syncwarp), the small
//               per-step outputs, then a SNAPSHOT of the post-step tile state (positions, actions, id grid, agent bit
//               rows: < 4 KB) into buffer t & 1;
//   warps 1-3   OBSERVATION role: window planes, action masks, goal bits and goal vectors of step t from snapshot
//               t & 1 (one thread per agent), then the bit -> byte expansion with all 96 threads.
// Named barriers (bar.sync / bar.arrive, ids 1-4) hand the two snapshot buffers back and forth: "ready[b]" (step role
// arrives, observation role waits) and "free[b]" (the other way round); barrier 5 is the observation role's own.  The
// chain of a step becomes max(step role, observation role) instead of their sum.  Same helper functions, same
// results: the parity tests compare mapf_rollout with consecutive mapf_step_observe calls bit for bit.
// Restrictions (the launcher falls back to the ROLL kernel otherwise): 5-action PRIMAL, odd specialised F, at most 32
// agents per environment, full sweeps, no mid-sweep outputs, observation requested.
// ------------------------------------------------------------------------------------------------
struct MapfPipeLayout {
  int obst_off;                 // [shared ? 1 : epb][bm_words] u32
  int grid_off, agt_off;        // live id grid [epb][grid_bytes] and agent bit rows [epb][bm_words] (adjacent)
  int snap_off[2];              // snapshots of the two, same shape, adjacent
  int snappos_off[2];           // uchar2 [32]
  int snapact_off[2];           // u8 [32]
  int goal_off, posold_off, posnew_off;   // uchar2 [32]
  int mv_off;                   // u32 [32]
  int res_off, dep_off, act_off, status_off, done_off, flag_off;   // u8 [32]
  int rew_off;                  // double [32]
  int envrew_off;               // double [32]
  int envcnt_off;               // int [32]
  int str_off;                  // observation bit strings
  int live_bytes;               // bytes of (id grid + agent bit rows) of the tile: what a snapshot copies
  int guard_off[8];             // 16-byte canaries between the regions (MAPF_FLAG_INTERNAL)
  int total_bytes;
};

// Barrier ids are IMMEDIATES: with a register id ptxas reserves all 16 barriers of the CTA, and the SM then holds 4 such
// CTAs instead of 8 (measured: the pipelined kernel ran in two waves).
template <int ID>
__device__ __forceinline__ void named_sync(int count) {
  asm volatile("bar.sync %0, %1;" ::"n"(ID), "r"(count) : "memory");
}
template <int ID>
__device__ __forceinline__ void named_arrive(int count) {
  asm volatile("bar.arrive %0, %1;" ::"n"(ID), "r"(count) : "memory");
}

#ifdef MAPF_PHASE_TIMING
#define PIPE_MARK(i, leader) do { if (blockIdx.x == gridDim.x / 2 && (leader) && t + 2 == nsteps) g_phase_clk[i] = clock64(); } while (0)
#else
#define PIPE_MARK(i, leader) do { } while (0)
#endif

template <int F>
__global__ void __launch_bounds__(kThreads, 8) mapf_pipe_kernel(const MapfDims d, const MapfPipeLayout L, const MapfState S,
                                                                const MapfTileArgs A, const int epb) {
  static_assert(kThreads == 128 && (F & 1) == 1 && F >= 3, "one step warp + three observation warps; symmetric windows");
  extern __shared__ __align__(16) unsigned char smem_raw[];
  using T = Fov<F>;
  constexpr unsigned full = 0xffffffffu;
  constexpr int kObsThreads = kThreads - 32;
  const int tid = threadIdx.x, lane = tid & 31;
  const int N = d.N;
  const int e0 = blockIdx.x * epb;
  const int ne = min(epb, d.E - e0);
  const int na = ne * N;                                   // <= 32: one lane of the step warp per agent
  const size_t a0 = (size_t)e0 * N, EN = (size_t)d.E * N;
  uint32_t* obst = (uint32_t*)(smem_raw + L.obst_off);
  uchar2* goal = (uchar2*)(smem_raw + L.goal_off);
  const int nsteps = A.T;

  // ---- both roles: the obstacle rows of the tile, once
  {
    const int nvec = ((d.shared_map ? 1 : ne) * d.bm_words) >> 2;
    const uint4* osrc = (const uint4*)(S.obst_bits + (d.shared_map ? (size_t)0 : (size_t)e0 * d.bm_words));
    for (int i = tid; i < nvec; i += kThreads) ((uint4*)obst)[i] = __ldg(osrc + i);
  }
  if (tid < na) goal[tid] = ((const uchar2*)S.goal)[a0 + tid];
  if (tid >= 64 && tid < 70) *(uint32_t*)(smem_raw + L.guard_off[tid - 64]) = kCanary + (tid - 64);
  __syncthreads();

  if (tid < 32) {
    // =========================================================== STEP role (warp 0) ================================
    Smem s;
    s.obst = obst;
    s.agt = (uint32_t*)(smem_raw + L.agt_off);
    s.grida = smem_raw + L.grid_off;
    s.gridb = s.grida;
    s.posold = (uchar2*)(smem_raw + L.posold_off);
    s.posnew = (uchar2*)(smem_raw + L.posnew_off);
    s.goal = goal;
    s.mv = (uint32_t*)(smem_raw + L.mv_off);
    s.res = smem_raw + L.res_off;
    s.dep = smem_raw + L.dep_off;
    s.act = smem_raw + L.act_off;
    s.status = (int8_t*)(smem_raw + L.status_off);
    s.done = smem_raw + L.done_off;
    s.flag = smem_raw + L.flag_off;
    s.rew = (double*)(smem_raw + L.rew_off);
    s.envrew = (double*)(smem_raw + L.envrew_off);
    int* envcnt = (int*)(smem_raw + L.envcnt_off);
    const int j = lane;
    const bool active = j < na;
    const int el = active ? fast_div(j, d.invN) : 0, a = active ? j - el * N : 0;
    uchar2 p = make_uchar2(0, 0);
    int sc = 0;
    if (active) p = ((const uchar2*)S.pos)[a0 + j];
    if (lane < ne) sc = S.step_count[e0 + lane];
    auto load_action = [&](size_t at) -> long long {
      return (A.act_dtype == MAPF_I64) ? ((const long long*)A.actions)[at + j]
                                       : (long long)((const uint8_t*)A.actions)[at + j];
    };
    long long av = active ? load_action(a0) : 0;
    // the live id grid (PRIMAL State.state, PRIMAL:32-47) and agent bit rows: built once, then kept up to date
    for (int i = lane; i < (L.live_bytes >> 4); i += 32) ((uint4*)(smem_raw + L.grid_off))[i] = make_uint4(0, 0, 0, 0);
    __syncwarp();
    if (active) {
      s.grida[el * d.grid_bytes + gcell(d, p.x, p.y)] = (uint8_t)(a + 1);
      const int pc = (int)p.y + d.P;
      atomicOr(&s.agt[el * d.bm_words + ((int)p.x + d.P) * d.RW + (pc >> 5)], 1u << (pc & 31));
    }
    __syncwarp();
    unsigned int c0 = 0, c1 = 0, c3 = 0, ep_done = 0;
    bool bad = false;
    int act = 0;
    uint8_t dn = 0;
    size_t a0t = a0, e0t = (size_t)e0;
    for (int t = 0; t < nsteps; ++t, a0t += EN, e0t += (size_t)d.E) {
      const int b = t & 1;
      PIPE_MARK(16, lane == 0);
      long long next_av = 0;
      if (t + 1 < nsteps && active) next_av = load_action(a0t + EN);
      if (active) {
        long long v = av;
        if (v < 0 || v > 4) {                                          // PRIMAL:556 assert
          bad = true;
          v = 0;
        }
        act = (int)v;
        s.posold[j] = p;
        s.posnew[j] = p;
        s.act[j] = (uint8_t)act;
        primal_phase_a<false>(d, s, A, j, el, a);                      // reads only this agent's records + the walls
      }
      if (lane < ne) envcnt[lane] = 0;
      __syncwarp();
      // ---- phase B: the ordered sweep, resolved in parallel (primal_classify)
      bool pending = false;
      if (active) {
        const uint8_t r = primal_classify(d, s, j, el, a);
        s.res[j] = r;
        pending = (r == RES_UNRESOLVED);
      }
      __syncwarp();
      while (__any_sync(full, pending)) {
        pending = false;
        if (active && s.res[j] == RES_UNRESOLVED) {
          const int k = el * N + s.dep[j];
          const uint8_t rk = s.res[k];
          if (rk == RES_UNRESOLVED) {
            s.dep[j] = s.dep[k];                                       // pointer jumping
            pending = true;
          } else {
            s.res[j] = rk;
          }
        }
        __syncwarp();
      }
      if (active && s.res[j] == RES_MOVED) {                           // vacate: id grid and agent bit
        s.grida[el * d.grid_bytes + gcell(d, p.x, p.y)] = 0;
        const int pc = (int)p.y + d.P;
        atomicAnd(&s.agt[el * d.bm_words + ((int)p.x + d.P) * d.RW + (pc >> 5)], ~(1u << (pc & 31)));
      }
      __syncwarp();
      // ---- phase C: outcome, reward, arrival; the new cell enters the id grid and the agent bit rows
      bool on_goal = false;
      uchar2 pn = p;
      if (active) {
        on_goal = primal_phase_c<false>(d, s, j, el, a, c0, c1, c3);
        pn = s.posnew[j];
        dn = s.done[j];
        const int pc = (int)pn.y + d.P;
        atomicOr(&s.agt[el * d.bm_words + ((int)pn.x + d.P) * d.RW + (pc >> 5)], 1u << (pc & 31));
      }
      if (d.rsum_mode == 1 && A.out.reward_dev != nullptr) {           // pairwise team reward inside the warp (N | 32)
        double v = active ? s.rew[j] : 0.0;
        for (int sft = 1; sft < N; sft <<= 1) v = __dadd_rn(v, __shfl_down_sync(full, v, sft));
        if (active && a == 0) s.envrew[el] = v;
      }
      {
        const unsigned peers = __match_any_sync(full, active ? el : -1);
        const unsigned bf = __ballot_sync(full, on_goal);
        if (active && lane == __ffs(peers) - 1) envcnt[el] = __popc(bf & peers);
      }
      __syncwarp();
      // ---- phase D (lane per environment) + the small per-agent outputs of this step
      if (lane < ne) {
        const bool all = envcnt[lane] == N;                            // State.done, PRIMAL:159-165
        if (A.out.terminated_dev) A.out.terminated_dev[e0t + lane] = all ? 1 : 0;
        if (A.out.reward_dev)
          A.out.reward_dev[e0t + lane] = d.rsum_mode == 1 ? s.envrew[lane]
                                                           : tree_sum_agents(s.rew + lane * N, N, 0, N);
        ep_done += all ? 1u : 0u;
      }
      if (active) {
        const size_t gt = a0t + j;
        if (A.out.dones_dev) A.out.dones_dev[gt] = dn;
        if (A.out.status_dev) A.out.status_dev[gt] = s.status[j];
        if (A.out.agent_reward_dev) A.out.agent_reward_dev[gt] = s.rew[j];
        if (A.out.node_dev) A.out.node_dev[gt] = 0;
        if (A.out.edge_dev) A.out.edge_dev[gt] = 0;
        if (A.out.valid_dev) A.out.valid_dev[gt] = (s.flag[j] >> 1) & 1;
      }
      // ---- hand the post-step state to the observation role: wait until it has let go of buffer b (step t - 2)
      PIPE_MARK(17, lane == 0);
      if (t >= 2) {
        if (b) named_sync<4>(kThreads);
        else named_sync<3>(kThreads);
      }
      PIPE_MARK(18, lane == 0);
      if (active) {
        ((uchar2*)(smem_raw + (b ? L.snappos_off[1] : L.snappos_off[0])))[j] = pn;
        (smem_raw + (b ? L.snapact_off[1] : L.snapact_off[0]))[j] = (uint8_t)act;
      }
      {
        const uint4* src = (const uint4*)(smem_raw + L.grid_off);
        uint4* dst = (uint4*)(smem_raw + (b ? L.snap_off[1] : L.snap_off[0]));
        for (int i = lane; i < (L.live_bytes >> 4); i += 32) dst[i] = src[i];
      }
      __threadfence_block();
      if (b) named_arrive<2>(kThreads);
      else named_arrive<1>(kThreads);
      PIPE_MARK(19, lane == 0);
      p = pn;
      av = next_av;
    }
    // ---- the handle's state and statistics, once
    if (active) {
      ((uchar2*)S.pos)[a0 + j] = p;
      S.done[a0 + j] = dn;
      S.prev_action[a0 + j] = (uint8_t)act;
    }
    if (lane < ne) S.step_count[e0 + lane] = sc + nsteps;
    if (__any_sync(full, bad) && lane == 0) atomicOr(S.err_flags, MAPF_FLAG_BAD_ACTION);
    if (d.collect_stats) {
      c0 = __reduce_add_sync(full, c0);
      c1 = __reduce_add_sync(full, c1);
      c3 = __reduce_add_sync(full, c3);
      ep_done = __reduce_add_sync(full, ep_done);
      if (lane == 0) {
        atomicAdd(&S.stats[MAPF_STAT_ENV_STEPS], (unsigned long long)ne * nsteps);
        atomicAdd(&S.stats[MAPF_STAT_AGENT_STEPS], (unsigned long long)na * nsteps);
        if (c0) atomicAdd(&S.stats[MAPF_STAT_ENV_COLLISIONS], (unsigned long long)c0);
        if (c1) atomicAdd(&S.stats[MAPF_STAT_NODE_COLLISIONS], (unsigned long long)c1);
        if (c3) atomicAdd(&S.stats[MAPF_STAT_GOAL_ARRIVALS], (unsigned long long)c3);
        if (ep_done) atomicAdd(&S.stats[MAPF_STAT_EPISODES_DONE], (unsigned long long)ep_done);
      }
    }
  } else {
    // =========================================================== OBSERVATION role (warps 1-3) =======================
    const int ot = tid - 32;                               // 0 .. 95; agent ot of the tile when ot < na
    const int j = ot;
    const bool valid = j < na;
    const int el = valid ? fast_div(j, d.invN) : 0;
    uint32_t* str = (uint32_t*)(smem_raw + L.str_off);
    const uchar2 g = valid ? goal[j] : make_uchar2(0, 0);
    size_t a0t = a0;
    for (int t = 0; t < nsteps; ++t, a0t += EN) {
      const int b = t & 1;
      PIPE_MARK(20, ot == 0);
      if (b) named_sync<2>(kThreads);                      // snapshot b holds the state after step t
      else named_sync<1>(kThreads);
      PIPE_MARK(21, ot == 0);
      const uchar2* pos = (const uchar2*)(smem_raw + (b ? L.snappos_off[1] : L.snappos_off[0]));
      const uint8_t* idgrid = smem_raw + (b ? L.snap_off[1] : L.snap_off[0]);
      const uint32_t* agt = (const uint32_t*)(smem_raw + (b ? L.snap_off[1] : L.snap_off[0]) + (L.agt_off - L.grid_off));
      // ---- phase 1: one thread per agent
      uint32_t first = 0;
      uint32_t vis[T::CW];
      int w0 = 0, sh = 0;
      uchar2 p = make_uchar2(0, 0);
      const double2* vt = nullptr;
      if (valid) {
        p = pos[j];
        if (A.vec != nullptr) {
          vt = (const double2*)S.vec_lut + 2 * (abs((int)g.x - (int)p.x) * d.W + abs((int)g.y - (int)p.y));
          asm volatile("prefetch.global.L1 [%0];" ::"l"(vt));
        }
        uint32_t w[T::NW];
        fov_window_planes<F>(w, d, obst + (d.shared_map ? 0 : el * d.bm_words), agt + el * d.bm_words, p);
        if (A.out.avail_dev != nullptr) {                  // _listNextValidActions from the window planes (see the tile kernel)
          constexpr int Pw = F / 2;
          auto open_cell = [&](int wi, int wj) -> uint32_t {
            const int i0 = wi * F + wj, i3 = 3 * T::FF + wi * F + wj;
            return (((w[i0 >> 5] >> (i0 & 31)) | (w[i3 >> 5] >> (i3 & 31))) & 1u) ^ 1u;
          };
          uint32_t m = 1u | (open_cell(Pw, Pw + 1) << 1) | (open_cell(Pw + 1, Pw) << 2) | (open_cell(Pw, Pw - 1) << 3) |
                       (open_cell(Pw - 1, Pw) << 4);
          const int my_act = (smem_raw + (b ? L.snapact_off[1] : L.snapact_off[0]))[j];
          const int opp = (my_act == 0) ? -1 : (((my_act + 1) & 3) + 1);
          if (opp > 0) m &= ~(1u << opp);
          uint8_t* o = A.out.avail_dev + 5 * (a0t + j);
          o[0] = m & 1;
          o[1] = (m >> 1) & 1;
          o[2] = (m >> 2) & 1;
          o[3] = (m >> 3) & 1;
          o[4] = (m >> 4) & 1;
        }
#pragma unroll
        for (int q = 0; q < T::CW; ++q) vis[q] = w[q];
        if ((T::FF & 31) != 0) vis[T::CW - 1] &= (1u << (T::FF & 31)) - 1u;
        vis[((F / 2) * F + F / 2) >> 5] &= ~(1u << (((F / 2) * F + F / 2) & 31));   // not the agent itself
        const int grp = j / d.G, k = j - grp * d.G;
        const int boff = k * T::NB;
        w0 = grp * d.GW + (boff >> 5);
        sh = boff & 31;
        const int nwords = (sh + T::NB + 31) >> 5;
        uint32_t prev = 0;
#pragma unroll
        for (int q = 0; q <= T::NW; ++q) {
          const uint32_t cur = (q < T::NW) ? w[q] : 0u;
          const uint32_t o = __funnelshift_l(prev, cur, sh);
          prev = cur;
          if (q == 0) first = o;
          if (q < nwords && !(q == 0 && sh > 0)) str[w0 + q] = o;
        }
        if (A.vec != nullptr) {                            // PRIMAL:380-385
          const int dx = (int)g.x - (int)p.x, dy = (int)g.y - (int)p.y;
          const double2 u = __ldg(vt);
          const double2 m = __ldg(vt + 1);
          double* v = A.vec + 3 * (a0t + j);
          v[0] = dx < 0 ? -u.x : u.x;
          v[1] = dy < 0 ? -u.y : u.y;
          v[2] = m.x;
        }
      }
      PIPE_MARK(22, ot == 0);
      named_sync<5>(kObsThreads);
      if (valid) {
        if (sh > 0) atomicOr(&str[w0], first);
        fov_goal_bits_half<F>(str, j, el * N, vis, d.GS, idgrid + el * d.grid_bytes, goal, p, g);
      }
      // this thread is done with snapshot b; the step role may refill it once all 96 have said so
      if (t + 2 < nsteps) {
        if (b) named_arrive<4>(kThreads);
        else named_arrive<3>(kThreads);
      }
      named_sync<5>(kObsThreads);
      PIPE_MARK(23, ot == 0);
      // ---- phase 2: expand the tile's bit string with all observation threads
      const size_t nbits = (size_t)na * T::NB;
      if (A.obs_dtype == MAPF_BITS) {
        uint32_t* out = (uint32_t*)A.obs + ((a0t * T::NB) >> 5);
        const int nw = (int)((nbits + 31) >> 5);
        for (int q = ot; q < nw; q += kObsThreads) out[q] = str[q];
      } else if (A.obs_dtype == MAPF_U8) {
        uint8_t* out = (uint8_t*)A.obs + a0t * T::NB;
        const int nchunk = (int)(nbits >> 4);
        const uint16_t* s16 = (const uint16_t*)str;
#pragma unroll 2
        for (int q = ot; q < nchunk; q += kObsThreads) {
          const uint32_t h = s16[q];
          uint4 v;
          v.x = expand4(h & 15u);
          v.y = expand4((h >> 4) & 15u);
          v.z = expand4((h >> 8) & 15u);
          v.w = expand4(h >> 12);
          st_stream16(out + ((size_t)q << 4), v);
        }
        for (int bb = (nchunk << 4) + ot; bb < (int)nbits; bb += kObsThreads) out[bb] = (str[bb >> 5] >> (bb & 31)) & 1u;
      } else {
        float* out = (float*)A.obs + a0t * T::NB;
        const int nchunk = (int)(nbits >> 2);
        for (int q = ot; q < nchunk; q += kObsThreads) {
          const uint32_t nib = (str[q >> 3] >> ((q & 7) << 2)) & 15u;
          uint4 v;
          v.x = (nib & 1u) ? 0x3f800000u : 0u;
          v.y = (nib & 2u) ? 0x3f800000u : 0u;
          v.z = (nib & 4u) ? 0x3f800000u : 0u;
          v.w = (nib & 8u) ? 0x3f800000u : 0u;
          st_stream16(out + ((size_t)q << 2), v);
        }
      }
      PIPE_MARK(24, ot == 0);
      // (the next step's phase 1 overwrites the strings only after its "ready" barrier, which every observation
      // thread reaches after finishing this expansion)
    }
  }
  // both roles are done: nobody wrote outside its region of the tile (the canaries between the regions are intact)
  if (A.debug_corrupt && tid == 0) smem_raw[L.guard_off[(A.debug_corrupt - 1) % 6]] ^= 0xff;   // self-test hook
  __syncthreads();
  if (tid < 6 && *(volatile uint32_t*)(smem_raw + L.guard_off[tid]) != kCanary + tid) atomicOr(S.err_flags, MAPF_FLAG_INTERNAL);
}

// ------------------------------------------------------------------------------------------------
// Generic-F observation (any F in [1, 255]); byte-wise gather, used when no specialised tile kernel
// exists for F and as an independent cross-check of the bit-string path in the tests.
// One thread per (agent, window cell).
// ------------------------------------------------------------------------------------------------
__global__ void mapf_observe_generic_kernel(const MapfDims d, const MapfState S, uint8_t* obs_u8, float* obs_f32,
                                            double* vec) {
  const int F = d.F, FF = F * F, Pw = F / 2;
  const long long total = (long long)d.E * d.N * FF;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const long long ja = idx / FF;
    const int cell = (int)(idx - ja * FF);
    const int e = (int)(ja / d.N), a = (int)(ja - (long long)e * d.N);
    const int wi = cell / F, wj = cell - wi * F;
    const uchar2* pos = (const uchar2*)S.pos + (size_t)e * d.N;
    const uchar2* goal = (const uchar2*)S.goal + (size_t)e * d.N;
    const uint32_t* ob = S.obst_bits + (d.shared_map ? (size_t)0 : (size_t)e * d.bm_words);
    const uchar2 p = pos[a], g = goal[a];
    const int t0 = (int)p.x - Pw, t1 = (int)p.y - Pw;
    const int i = t0 + wi, jx = t1 + wj;
    const bool oob = (i < 0 || i >= d.H || jx < 0 || jx >= d.W);
    uint8_t c_obs = oob ? 1 : (uint8_t)bm_test(ob, d.RW, d.P, i, jx);
    uint8_t c_poss = 0, c_goal = 0, c_goals = 0;
    if (!oob && g.x == i && g.y == jx) c_goal = 1;
    for (int b = 0; b < d.N; ++b) {
      const uchar2 q = pos[b];
      if (q.x == i && q.y == jx) c_poss = 1;
      if (b != a) {
        const int bi = (int)q.x - t0, bj = (int)q.y - t1;
        if ((unsigned)bi < (unsigned)F && (unsigned)bj < (unsigned)F) {   // b is visible to a
          const uchar2 og = goal[b];
          const int ci = min(max((int)og.x - t0, 0), F - 1), cj = min(max((int)og.y - t1, 0), F - 1);
          if (ci == wi && cj == wj) c_goals = 1;
        }
      }
    }
    const size_t o = (size_t)ja * 4 * FF + cell;
    if (obs_u8) {
      obs_u8[o] = c_poss;
      obs_u8[o + FF] = c_goal;
      obs_u8[o + 2 * FF] = c_goals;
      obs_u8[o + 3 * FF] = c_obs;
    }
    if (obs_f32) {
      obs_f32[o] = (float)c_poss;
      obs_f32[o + FF] = (float)c_goal;
      obs_f32[o + 2 * FF] = (float)c_goals;
      obs_f32[o + 3 * FF] = (float)c_obs;
    }
    if (vec && cell == 0) {
      const int dx = (int)g.x - (int)p.x, dy = (int)g.y - (int)p.y;
      const double mag = S.mag_lut[dx * dx + dy * dy];
      double fx = (double)dx, fy = (double)dy;
      if (mag != 0.0) {
        fx = __ddiv_rn(fx, mag);
        fy = __ddiv_rn(fy, mag);
      }
      vec[3 * ja] = fx;
      vec[3 * ja + 1] = fy;
      vec[3 * ja + 2] = mag;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Reset-time kernels
// ------------------------------------------------------------------------------------------------
// Padded obstacle bitmap: bit = 1 on walls and everywhere outside the map (out of bounds is treated as
// an obstacle by every consumer: GRID:336-340, PRIMAL:114-118, :356-359).
__global__ void mapf_build_obst_kernel(const MapfDims d, uint32_t* obst_bits, const int8_t* map,
                                       const uint8_t* env_mask) {
  const int emap = d.shared_map ? 1 : d.E;
  const long long total = (long long)emap * d.bm_words;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int e = (int)(idx / d.bm_words);
    const int wdx = (int)(idx - (long long)e * d.bm_words);
    if (env_mask && !d.shared_map && !env_mask[e]) continue;
    uint32_t bits = 0xffffffffu;
    const int prow = wdx / d.RW, k = wdx - prow * d.RW;
    const int r = prow - d.P;
    if (prow < d.PR && r >= 0 && r < d.H) {
      const int8_t* mrow = map + ((size_t)e * d.H + r) * d.W;
      bits = 0;
      for (int b = 0; b < 32; ++b) {
        const int c = 32 * k + b - d.P;
        const bool wall = (c < 0 || c >= d.W) ? true : (mrow[c] != 0);
        bits |= (wall ? 1u : 0u) << b;
      }
    }
    obst_bits[idx] = bits;
  }
}

__global__ void mapf_reset_kernel(const MapfDims d, const MapfState S, const int16_t* starts, const int16_t* goals,
                                  const uint8_t* env_mask) {
  const long long total = (long long)d.E * d.N;
  for (long long j = blockIdx.x * (long long)blockDim.x + threadIdx.x; j < total;
       j += (long long)gridDim.x * blockDim.x) {
    const int e = (int)(j / d.N), a = (int)(j - (long long)e * d.N);
    if (env_mask && !env_mask[e]) continue;
    uint32_t flags = 0;
    uchar2 st = ((const uchar2*)S.start)[j], g = ((const uchar2*)S.goal)[j];
    if (starts) {
      int s0 = starts[2 * j], s1 = starts[2 * j + 1];
      if (s0 < 0 || s0 >= d.H || s1 < 0 || s1 >= d.W) {
        flags |= MAPF_FLAG_BAD_POSITION;
        s0 = min(max(s0, 0), d.H - 1);
        s1 = min(max(s1, 0), d.W - 1);
      }
      st = make_uchar2((unsigned char)s0, (unsigned char)s1);
      ((uchar2*)S.start)[j] = st;
    }
    if (goals) {
      int g0 = goals[2 * j], g1 = goals[2 * j + 1];
      if (g0 < 0 || g0 >= d.H || g1 < 0 || g1 >= d.W) {
        flags |= MAPF_FLAG_BAD_POSITION;
        g0 = min(max(g0, 0), d.H - 1);
        g1 = min(max(g1, 0), d.W - 1);
      }
      g = make_uchar2((unsigned char)g0, (unsigned char)g1);
      ((uchar2*)S.goal)[j] = g;
    }
    const uint32_t* ob = S.obst_bits + (d.shared_map ? (size_t)0 : (size_t)e * d.bm_words);
    if (bm_test(ob, d.RW, d.P, st.x, st.y)) flags |= MAPF_FLAG_START_ON_WALL;
    ((uchar2*)S.pos)[j] = st;                                        // GRID:79
    if (d.diag) ((uchar2*)S.past)[j] = st;                           // agents_past == agents after scanForAgents, PRIMAL:61-65
    S.done[j] = (d.mode == MAPF_MODE_PRIMAL) ? (uint8_t)(st.x == g.x && st.y == g.y) : 0;   // GRID:75
    S.prev_action[j] = (d.mode == MAPF_MODE_PRIMAL) ? 0 : 4;
    if (a == 0) S.step_count[e] = 0;                                 // GRID:73
    if (d.mode == MAPF_MODE_PARTIAL) {                               // PARTIAL:135-150
      S.at_goal[j] = 0;
      S.goal_cost[j] = -1;
      S.agent_steps[j] = 0;
      S.pnode[j] = 0;
      S.pedge[j] = 0;
      if (a == 0) {
        S.total_coll[e] = 0;
        S.terminated[e] = 0;
      }
    }
    if (d.mode == MAPF_MODE_PRIMAL && starts) {                     // one agent per cell, PRIMAL:53-66
      for (int b = 0; b < a; ++b)
        if (starts[2 * ((long long)e * d.N + b)] == starts[2 * j] &&
            starts[2 * ((long long)e * d.N + b) + 1] == starts[2 * j + 1])
          flags |= MAPF_FLAG_START_OVERLAP;
    }
    if (d.mode == MAPF_MODE_PRIMAL && goals) {
      for (int b = 0; b < a; ++b)
        if (goals[2 * ((long long)e * d.N + b)] == goals[2 * j] &&
            goals[2 * ((long long)e * d.N + b) + 1] == goals[2 * j + 1])
          flags |= MAPF_FLAG_GOAL_OVERLAP;
    }
    if (flags) atomicOr(S.err_flags, flags);
  }
}

__global__ void mapf_set_goals_kernel(const MapfDims d, const MapfState S, const int16_t* goals,
                                      const uint8_t* dirty) {
  const long long total = (long long)d.E * d.N;
  for (long long j = blockIdx.x * (long long)blockDim.x + threadIdx.x; j < total;
       j += (long long)gridDim.x * blockDim.x) {
    if (dirty && !dirty[j]) continue;
    int g0 = goals[2 * j], g1 = goals[2 * j + 1];
    if (g0 < 0 || g0 >= d.H || g1 < 0 || g1 >= d.W) {
      atomicOr(S.err_flags, MAPF_FLAG_BAD_POSITION);
      g0 = min(max(g0, 0), d.H - 1);
      g1 = min(max(g1, 0), d.W - 1);
    }
    ((uchar2*)S.goal)[j] = make_uchar2((unsigned char)g0, (unsigned char)g1);
    if (d.mode == MAPF_MODE_PRIMAL) {
      const uchar2 p = ((const uchar2*)S.pos)[j];
      S.done[j] = (uint8_t)(p.x == g0 && p.y == g1);
    }
  }
}

// Lifelong task hand-out (MAPF-490-main/Global.cpp:85-94): an agent standing on its goal pops the front of its queue.
__global__ void mapf_pop_goals_kernel(const MapfDims d, const MapfState S, const int16_t* __restrict__ queue,
                                      int32_t* __restrict__ head, int Q, uint8_t* __restrict__ dirty) {
  const long long total = (long long)d.E * d.N;
  for (long long j = blockIdx.x * (long long)blockDim.x + threadIdx.x; j < total;
       j += (long long)gridDim.x * blockDim.x) {
    const uchar2 p = ((const uchar2*)S.pos)[j];
    const uchar2 g = ((const uchar2*)S.goal)[j];
    const int hd = head[j];
    const bool pop = p.x == g.x && p.y == g.y && hd < Q;
    if (dirty) dirty[j] = (uint8_t)pop;
    if (!pop) continue;
    const short2 q = ((const short2*)queue)[j * Q + hd];
    int g0 = q.x, g1 = q.y;
    if (g0 < 0 || g0 >= d.H || g1 < 0 || g1 >= d.W) {
      atomicOr(S.err_flags, MAPF_FLAG_BAD_POSITION);
      g0 = min(max(g0, 0), d.H - 1);
      g1 = min(max(g1, 0), d.W - 1);
    }
    head[j] = hd + 1;
    ((uchar2*)S.goal)[j] = make_uchar2((unsigned char)g0, (unsigned char)g1);
    if (d.mode == MAPF_MODE_PRIMAL) S.done[j] = (uint8_t)(p.x == g0 && p.y == g1);
  }
}

// Counter-based random policy (mapf_random_actions): action of (env, agent) at `step` = a 32-bit mix of
// (seed, GLOBAL env index, step, agent) -- the same function as mapf_marl_b200/workloads.py hash_actions_np -- taken
// uniformly over {0..nact-1}, or over the set bits of the agent's action mask (the r-th available action,
// r = hash mod popcount).  A pure function of the global env index: shards of a batch draw what the whole batch would.
__device__ __forceinline__ uint32_t mix32(uint32_t x) {
  x ^= x >> 15;
  x *= 0x2C1B3C6Du;
  x ^= x >> 12;
  x *= 0x297A2D39u;
  x ^= x >> 15;
  return x;
}

__global__ void mapf_random_actions_kernel(const MapfDims d, const uint8_t* __restrict__ avail, uint32_t seed,
                                           uint32_t step, long long env_offset, uint8_t* out8, long long* out64) {
  const long long total = (long long)d.E * d.N;
  for (long long j = blockIdx.x * (long long)blockDim.x + threadIdx.x; j < total;
       j += (long long)gridDim.x * blockDim.x) {
    const long long e = j / d.N;
    const uint32_t a = (uint32_t)(j - e * d.N);
    const uint32_t x = (uint32_t)(e + env_offset) * 0x9E3779B1u + step * 0x85EBCA77u + a * 0xC2B2AE3Du +
                       seed * 0x27D4EB2Fu;
    const uint32_t h = mix32(mix32(x) + 0x165667B1u) >> 8;
    int act;
    if (avail == nullptr) {
      act = (int)((h * (uint32_t)d.nact) >> 24);
    } else {
      const uint8_t* m = avail + j * d.nact;
      int cnt = 0;
      for (int k = 0; k < d.nact; ++k) cnt += m[k] != 0;
      act = 0;
      if (cnt > 0) {
        int r = (int)(h % (uint32_t)cnt);
        for (int k = 0; k < d.nact; ++k)
          if (m[k] != 0 && r-- == 0) {
            act = k;
            break;
          }
      }
    }
    if (out8) out8[j] = (uint8_t)act;
    if (out64) out64[j] = act;
  }
}

__global__ void mapf_export16_kernel(long long n, const uint8_t* src, int16_t* dst) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    dst[i] = (int16_t)src[i];
}

// ------------------------------------------------------------------------------------------------
// Per-goal BFS distance maps: one warp per (env, agent).  Rows are bit masks; a wavefront step is
// new = (left | right | up | down neighbours of the frontier) & free & ~visited.
// (PARTIAL:931-955 == hop distance; PRIMAL getAstarCosts :407-499 == hop distance from the goal.)
// smem per warp: 4 bitmaps [H][RWB] (free, visited, frontier A/B) and, when it fits, the int16 map.
// ------------------------------------------------------------------------------------------------
// Masked launches (goal re-assignment, masked reset) first compact the flagged (env, agent) indices into S.bfs_list
// (two counters, then the list, then the overflow list of the register kernel) and then run a grid of resident warps over that list: a launch of E*N mostly idle warps
// costs a block dispatch per 8 maps (150 us at c4 for a few dozen dirty maps), the list costs two small launches.
__global__ void mapf_bfs_compact_kernel(const MapfDims d, const uint8_t* __restrict__ dirty,
                                        const uint8_t* __restrict__ env_mask, int32_t* __restrict__ list) {
  const long long total = (long long)d.E * d.N;
  int32_t* count = list;             // [0] flagged maps, [1] overflow (mapf_launch_bfs zeroes both)
  list += 2;
  const int lane = threadIdx.x & 31;
  for (long long base = (blockIdx.x * (long long)blockDim.x + threadIdx.x) - lane; base < total;
       base += (long long)gridDim.x * blockDim.x) {
    const long long j = base + lane;
    bool on = j < total;
    if (on && dirty) on = dirty[j] != 0;
    if (on && env_mask) on = env_mask[j / d.N] != 0;
    const unsigned ballot = __ballot_sync(0xffffffffu, on);
    if (!ballot) continue;
    int off = 0;
    if (lane == 0) off = atomicAdd(count, __popc(ballot));
    off = __shfl_sync(0xffffffffu, off, 0);
    if (on) list[off + __popc(ballot & ((1u << lane) - 1))] = (int32_t)j;
  }
}

__device__ __forceinline__ void bfs_smem_one(const MapfDims& d, const MapfState& S, const long long m, int16_t* dist,
                                             int RWB, int stage_dist, int conn8, unsigned char* smem_raw) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int e = (int)(m / d.N);
  const int H = d.H, W = d.W, items = H * RWB;
  const size_t per_warp = (size_t)4 * items * 4 + (stage_dist ? (((size_t)d.HW * 2 + 15) & ~(size_t)15) : 0);
  unsigned char* base = smem_raw + per_warp * warp;
  uint32_t* freeb = (uint32_t*)base;
  uint32_t* vis = freeb + items;
  uint32_t* fa = vis + items;
  uint32_t* fb = fa + items;
  int16_t* sd = (int16_t*)(fb + items);
  int16_t* gd = dist + (size_t)m * d.HW;
  int16_t* dd = stage_dist ? sd : gd;
  const uint32_t* ob = S.obst_bits + (d.shared_map ? (size_t)0 : (size_t)e * d.bm_words);
  const uchar2 g = ((const uchar2*)S.goal)[m];

  for (int it = lane; it < items; it += 32) {
    const int r = it / RWB, k = it - r * RWB;
    const uint32_t* prow = ob + (r + d.P) * d.RW;
    const int pb = 32 * k + d.P;
    uint32_t wall = __funnelshift_r(prow[pb >> 5], prow[(pb >> 5) + 1], pb);
    const int rem = W - 32 * k;
    const uint32_t valid = rem >= 32 ? 0xffffffffu : ((1u << rem) - 1u);
    const uint32_t fr = ~wall & valid;
    freeb[it] = fr;
    uint32_t start = 0;
    if (r == g.x && (g.y >> 5) == k) start = (1u << (g.y & 31)) & fr;
    vis[it] = start;
    fa[it] = start;
    for (int b = 0; b < 32 && 32 * k + b < W; ++b) {
      const int c = 32 * k + b;
      dd[r * W + c] = ((fr >> b) & 1u) ? (((start >> b) & 1u) ? 0 : -2) : -1;
    }
  }
  __syncwarp();
  uint32_t* cur = fa;
  uint32_t* nxt = fb;
  for (int level = 1; level < 32767; ++level) {
    uint32_t any = 0;
    for (int it = lane; it < items; it += 32) {
      const int r = it / RWB, k = it - r * RWB;
      // a row of the frontier spread one cell to the left and right (bits carried across the 32-cell words)
      auto spread = [&](int idx) {
        const uint32_t f = cur[idx];
        uint32_t sp = f | (f << 1) | (f >> 1);
        if (k > 0) sp |= cur[idx - 1] >> 31;
        if (k < RWB - 1) sp |= cur[idx + 1] << 31;
        return sp;
      };
      uint32_t nb = spread(it);
      if (r > 0) nb |= conn8 ? spread(it - RWB) : cur[it - RWB];     // 8 neighbours: getNeighbors with diagonals, PRIMAL:421-437
      if (r < H - 1) nb |= conn8 ? spread(it + RWB) : cur[it + RWB];
      uint32_t nw = nb & freeb[it] & ~vis[it];
      vis[it] |= nw;
      nxt[it] = nw;
      any |= nw;
      while (nw) {
        const int b = __ffs(nw) - 1;
        nw &= nw - 1;
        dd[r * W + 32 * k + b] = (int16_t)level;
      }
    }
    __syncwarp();
    if (!__any_sync(0xffffffffu, any != 0)) break;
    uint32_t* t = cur;
    cur = nxt;
    nxt = t;
  }
  __syncwarp();
  if (stage_dist) {
    if ((d.HW & 7) == 0) {
      const int nv = d.HW >> 3;
      for (int i = lane; i < nv; i += 32) ((uint4*)gd)[i] = ((const uint4*)sd)[i];
    } else {
      for (int i = lane; i < d.HW; i += 32) gd[i] = sd[i];
    }
  }
}


__global__ void mapf_bfs_kernel(const MapfDims d, const MapfState S, const int32_t* __restrict__ list,
                                const int32_t* __restrict__ count, int16_t* dist, int RWB, int warps_per_block,
                                int stage_dist, int conn8) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5;
  const long long all = (long long)d.E * d.N;
  const long long total = list ? (long long)*count : all;
  for (long long i = (long long)blockIdx.x * warps_per_block + warp; i < total;
       i += (long long)gridDim.x * warps_per_block) {
    bfs_smem_one(d, S, list ? (long long)list[i] : i, dist, RWB, stage_dist, conn8, smem_raw);
    __syncwarp();
  }
}

// ------------------------------------------------------------------------------------------------
// Warp-synchronous BFS for maps up to 64 x 64: the whole map lives in registers, lane l holds rows
// [l*RPL, (l+1)*RPL) as W-bit masks (Row = uint32_t or uint64_t).  One wavefront step is two shuffles (the rows
// above and below) and a handful of logic ops per row.  The DISTANCES live in registers too, bit-sliced: plane p of a
// row has bit c set when bit p of cell c's distance is set.  Levels are processed in batches of eight: inside a batch
// the low three bits of the level are compile-time constants (a cell reached at sub-level i is OR-ed into the planes
// of i's set bits), the upper five planes take the union of the batch when the batch's base has that bit.  No
// per-cell work, no divergence and no shared memory until the end, when every lane turns the planes of its own rows
// into int16 cells (4 cells per multiply-and-mask, walls -1, unreachable -2) and stores them.  Eight planes cover
// 255 levels; a map whose wavefront is still moving after that is put on the overflow list and redone by the
// shared-memory kernel (mapf_bfs_kernel), which has no such limit.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t spread4(uint32_t nib) {      // bits 0..3 -> bit 0 of bytes 0..3
  return (nib * 0x00204081u) & 0x01010101u;
}

template <typename Row, int RPL, bool CONN8>
__device__ __forceinline__ bool bfs_warp_one(const MapfDims& d, const MapfState& S, const long long m, int16_t* dist) {
  const int lane = threadIdx.x & 31;
  const int e = (m >> 31) == 0 ? (int)((unsigned)m / (unsigned)d.N) : (int)(m / d.N);   // (the 64-bit division is a call)
  const int H = d.H, W = d.W;
  constexpr int RB = sizeof(Row) * 8;
  constexpr int NP = 8;
  const unsigned full = 0xffffffffu;
  int16_t* gd = dist + (size_t)m * d.HW;
  const uint32_t* ob = S.obst_bits + (d.shared_map ? (size_t)0 : (size_t)e * d.bm_words);
  const uchar2 g = ((const uchar2*)S.goal)[m];
  const Row valid = (W >= RB) ? ~(Row)0 : (((Row)1 << W) - 1);
  Row freeR[RPL], fnv[RPL], f[RPL], D[RPL][NP];      // fnv: free and not yet visited
#pragma unroll
  for (int k = 0; k < RPL; ++k) {
    const int r = lane * RPL + k;
    Row fr = 0;
    if (r < H) {
      const uint32_t* prow = ob + (r + d.P) * d.RW + (d.P >> 5);
      Row wall = __funnelshift_r(prow[0], prow[1], d.P);
      if (RB == 64) wall |= (Row)__funnelshift_r(prow[1], prow[2], d.P) << (RB / 2);
      fr = ~wall & valid;
    }
    freeR[k] = fr;
    const Row start = (r == g.x) ? (((Row)1 << g.y) & fr) : 0;     // level 0: all planes zero
    f[k] = start;
    fnv[k] = fr & ~start;
#pragma unroll
    for (int p = 0; p < NP; ++p) D[k][p] = 0;
  }
  bool open = true;
  // Levels run in GROUPS of 32 (runtime loop) made of four unrolled BATCHES of 8: inside a batch bits 0-2 of the level
  // are compile-time constants, inside a group bits 3-4 are, and bits 5-7 are constant for the whole group -- the three
  // upper planes take the union of the group ONCE, when the group ends or the wavefront dies (per batch that was a
  // mask, an AND and an add for each of them: 12 of a batch's 83 instructions).
  for (int G0 = 0; G0 < (1 << NP) && open; G0 += 32) {
    Row fnvG[RPL];
#pragma unroll
    for (int k = 0; k < RPL; ++k) fnvG[k] = fnv[k];
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      // The kernel is bound by the integer ALU pipe (LOP3 / SHF: one warp instruction per two cycles per scheduler),
      // the FMA pipe (IMAD) idles.  Every update below whose operands are DISJOINT bit sets is therefore written as an
      // add / subtract instead of an or / and-not -- a cell is reached exactly once, so the frontier never overlaps
      // the distance planes or leaves the free-and-not-visited set -- which lets ptxas place it on the FMA pipe.
      Row fnv0[RPL];
#pragma unroll
      for (int k = 0; k < RPL; ++k) fnv0[k] = fnv[k];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (i > 0 || b > 0 || G0 > 0) {
          // rows beyond the first / last lane come back as the lane's own row, which is adjacent anyway: harmless
          const Row from_above = __shfl_up_sync(full, f[RPL - 1], 1);     // last row of the lane above
          const Row from_below = __shfl_down_sync(full, f[0], 1);         // first row of the lane below
          Row nw[RPL];
#pragma unroll
          for (int k = 0; k < RPL; ++k) {
            Row up = (k > 0) ? f[k - 1] : from_above;
            Row dn = (k < RPL - 1) ? f[k + 1] : from_below;
            if (CONN8) {                                                    // diagonal neighbours, PRIMAL:421-437
              up |= (up << 1) | (up >> 1);
              dn |= (dn << 1) | (dn >> 1);
            }
            // (the right shift stays a SHF: as IMAD.HI -- __umulhi(f, 1u << 31) -- it measured 5 % slower)
            nw[k] = ((f[k] + f[k]) | (f[k] >> 1) | up | dn) & fnv[k];
          }
#pragma unroll
          for (int k = 0; k < RPL; ++k) {
            f[k] = nw[k];
            fnv[k] -= nw[k];                 // nw is a subset of fnv
          }
        }
#pragma unroll
        for (int k = 0; k < RPL; ++k) {
          if (i & 1) D[k][0] += f[k];        // disjoint: this cell's distance bits are written once
          if (i & 2) D[k][1] += f[k];
          if (i & 4) D[k][2] += f[k];
        }
      }
      bool any = false;
#pragma unroll
      for (int k = 0; k < RPL; ++k) {
        const Row U = fnv0[k] - fnv[k];      // everything reached in this batch (level 0 has all planes zero anyway)
        if (b & 1) D[k][3] += U;
        if (b & 2) D[k][4] += U;
        any |= f[k] != 0;
      }
      if (!__any_sync(full, any)) {
        open = false;
        break;
      }
    }
#pragma unroll
    for (int k = 0; k < RPL; ++k) {
      const Row UG = fnvG[k] - fnv[k];       // everything reached in this group
#pragma unroll
      for (int p = 5; p < NP; ++p)
        if ((G0 >> p) & 1) D[k][p] += UG;
    }
  }
  if (open) return false;                  // deeper than 255 levels: the caller hands the map to the generic kernel
  // ---- planes -> int16 cells.  Not-visited cells get low byte 0xff (wall) / 0xfe (free) and high byte 0xff.
#pragma unroll
  for (int k = 0; k < RPL; ++k) {
    const int r = lane * RPL + k;
    const Row wall = ~freeR[k];
    const Row nv = wall | fnv[k];
#pragma unroll
    for (int p = 0; p < NP; ++p) D[k][p] |= (p == 0) ? wall : nv;
    if (r >= H) continue;
    int16_t* grow = gd + (size_t)r * W;
    if ((W & 3) == 0) {
      // 32 cells at a time: an 8 x 8 bit-matrix transpose per byte lane (three rounds of masked delta swaps between
      // the plane words) leaves, in byte g of word r, the low byte of cell 8g + r; PRMTs gather them into cell order
      // and interleave the high bytes (0xff on not-visited cells)
#pragma unroll
      for (int hh = 0; hh < RB / 32; ++hh) {
        if (32 * hh >= W) break;
        uint32_t a[NP];
#pragma unroll
        for (int p = 0; p < NP; ++p) a[p] = (uint32_t)(D[k][p] >> (32 * hh));
#define BFS_SWAP(x, y, sft, m)                         \
  do {                                                 \
    const uint32_t t_ = (((x) >> (sft)) ^ (y)) & (m);  \
    (y) ^= t_;                                         \
    (x) ^= t_ << (sft);                                \
  } while (0)
        BFS_SWAP(a[0], a[1], 1, 0x55555555u); BFS_SWAP(a[2], a[3], 1, 0x55555555u);
        BFS_SWAP(a[4], a[5], 1, 0x55555555u); BFS_SWAP(a[6], a[7], 1, 0x55555555u);
        BFS_SWAP(a[0], a[2], 2, 0x33333333u); BFS_SWAP(a[1], a[3], 2, 0x33333333u);
        BFS_SWAP(a[4], a[6], 2, 0x33333333u); BFS_SWAP(a[5], a[7], 2, 0x33333333u);
        BFS_SWAP(a[0], a[4], 4, 0x0f0f0f0fu); BFS_SWAP(a[1], a[5], 4, 0x0f0f0f0fu);
        BFS_SWAP(a[2], a[6], 4, 0x0f0f0f0fu); BFS_SWAP(a[3], a[7], 4, 0x0f0f0f0fu);
#undef BFS_SWAP
        const uint32_t nvh = (uint32_t)(nv >> (32 * hh));
#pragma unroll
        for (int q = 0; q < 8; q += 2) {                       // groups of four cells, two groups per 16-byte store
          if (32 * hh + 4 * q >= W) break;
          uint32_t o[4];
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            constexpr uint32_t kPick[4] = {0x0040u, 0x0051u, 0x0062u, 0x0073u};   // byte g of x, byte g of y
            const int g = (q + h) >> 1, r0 = 4 * ((q + h) & 1);
            const uint32_t t01 = __byte_perm(a[r0], a[r0 + 1], kPick[g]);
            const uint32_t t23 = __byte_perm(a[r0 + 2], a[r0 + 3], kPick[g]);
            const uint32_t lo = __byte_perm(t01, t23, 0x5410);
            const uint32_t hi = spread4((nvh >> (4 * (q + h))) & 0xfu) * 0xffu;
            o[2 * h] = __byte_perm(lo, hi, 0x5140);         // cells 0,1 of the group: (lo0, hi0, lo1, hi1)
            o[2 * h + 1] = __byte_perm(lo, hi, 0x7362);     // cells 2,3
          }
          int16_t* dst = grow + 32 * hh + 4 * q;
          if ((W & 7) == 0) {
            *(uint4*)dst = make_uint4(o[0], o[1], o[2], o[3]);
          } else {
            *(uint2*)dst = make_uint2(o[0], o[1]);
            if (32 * hh + 4 * (q + 1) < W) *(uint2*)(dst + 4) = make_uint2(o[2], o[3]);
          }
        }
      }
    } else {
      for (int c = 0; c < W; ++c) {
        uint32_t v = ((uint32_t)(nv >> c) & 1u) * 0xff00u;
#pragma unroll
        for (int p = 0; p < NP; ++p) v |= ((uint32_t)(D[k][p] >> c) & 1u) << p;
        grow[c] = (int16_t)v;
      }
    }
  }
  return true;
}

// cnt[0]: length of the list of flagged maps (masked launch), cnt[1]: length of the overflow list
template <typename Row, int RPL, bool CONN8>
__global__ void mapf_bfs_warp_kernel(const MapfDims d, const MapfState S, const int32_t* __restrict__ list,
                                     int32_t* __restrict__ cnt, int32_t* __restrict__ overflow, int16_t* dist,
                                     int warps_per_block) {
  const int warp = threadIdx.x >> 5;
  const long long all = (long long)d.E * d.N;
  const long long total = list ? (long long)cnt[0] : all;
  for (long long i = (long long)blockIdx.x * warps_per_block + warp; i < total;
       i += (long long)gridDim.x * warps_per_block) {
    const long long m = list ? (long long)list[i] : i;
    if (!bfs_warp_one<Row, RPL, CONN8>(d, S, m, dist) && (threadIdx.x & 31) == 0)
      overflow[atomicAdd(&cnt[1], 1)] = (int32_t)m;
  }
}

// ------------------------------------------------------------------------------------------------
// PRIMAL blocking reward (get_blocking_reward, PRIMAL:513-546), evaluated after the sweep for every agent that
// stayed on its goal (action 0, status 1; PRIMAL:579-585).  One warp per (env, agent).  The reference evaluates it in
// the middle of the sweep, so positions are taken "at time i": agents below i stand on their new cells, agents above
// i still on their old ones.  For every visible robot b (ids 1..N-1: the reference's loop skips the last id, :523)
// two single-robot shortest paths from b to its goal are compared -- the other visible robots are obstacles, with and
// without this agent (:535-539).  Paths are warp-synchronous BFS runs over register-resident bit rows (64 x 64 max).
// ------------------------------------------------------------------------------------------------
template <typename Row, int RPL>
__device__ __forceinline__ int warp_path_hops(const Row (&freeR)[RPL], int s0, int s1, int g0, int g1, int lane) {
  // -1: start or goal blocked, or the goal cannot be reached (NoSolutionError, PRIMAL:505-508)
  Row f[RPL], vis[RPL];
  bool ok_s = false, ok_g = false;
#pragma unroll
  for (int k = 0; k < RPL; ++k) {
    const int r = lane * RPL + k;
    f[k] = (r == s0) ? (((Row)1 << s1) & freeR[k]) : 0;
    vis[k] = f[k];
    ok_s |= f[k] != 0;
    ok_g |= (r == g0) && ((freeR[k] >> g1) & 1);
  }
  if (!__any_sync(0xffffffffu, ok_s) || !__any_sync(0xffffffffu, ok_g)) return -1;
  if (s0 == g0 && s1 == g1) return 0;
  for (int level = 1; level < 8192; ++level) {
    const Row from_above = __shfl_up_sync(0xffffffffu, f[RPL - 1], 1);
    const Row from_below = __shfl_down_sync(0xffffffffu, f[0], 1);
    bool any = false, hit = false;
    Row nw[RPL];
#pragma unroll
    for (int k = 0; k < RPL; ++k) {
      const Row up = (k > 0) ? f[k - 1] : (lane > 0 ? from_above : 0);
      const Row dn = (k < RPL - 1) ? f[k + 1] : (lane < 31 ? from_below : 0);
      nw[k] = ((f[k] << 1) | (f[k] >> 1) | up | dn) & freeR[k] & ~vis[k];
      any |= nw[k] != 0;
      hit |= (lane * RPL + k == g0) && ((nw[k] >> g1) & 1);
    }
#pragma unroll
    for (int k = 0; k < RPL; ++k) {
      vis[k] |= nw[k];
      f[k] = nw[k];
    }
    if (__any_sync(0xffffffffu, hit)) return level;
    if (!__any_sync(0xffffffffu, any)) return -1;
  }
  return -1;
}

template <typename Row, int RPL>
__global__ void mapf_blocking_kernel(const MapfDims d, const MapfState S, int agent_lo, int agent_hi,
                                     uint8_t* blocking_out, int warps_per_block) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long m = (long long)blockIdx.x * warps_per_block + warp;
  if (m >= (long long)d.E * d.N) return;
  const int e = (int)(m / d.N), i = (int)(m - (long long)e * d.N);
  if (lane == 0 && blocking_out) blocking_out[m] = 0;
  if (i < agent_lo || i >= agent_hi) return;
  if (S.prev_action[m] != 0 || S.last_status[m] != 1) return;          // only "stayed on goal", PRIMAL:579-580
  const int H = d.H, W = d.W, N = d.N, F = d.F;
  constexpr int RB = sizeof(Row) * 8;
  Row* rows = (Row*)(smem_raw + (size_t)warp * 64 * sizeof(Row));      // robots bit rows of this warp
  const uchar2* pnew = (const uchar2*)S.pos + (size_t)e * N;
  const uchar2* pold = (const uchar2*)S.pos_prev + (size_t)e * N;
  const uchar2* goal = (const uchar2*)S.goal + (size_t)e * N;
  const uint32_t* ob = S.obst_bits + (d.shared_map ? (size_t)0 : (size_t)e * d.bm_words);
  const uchar2 me = pnew[i];
  const int tl0 = (int)me.x - F / 2, tl1 = (int)me.y - F / 2;
  const Row valid = (W >= RB) ? ~(Row)0 : (((Row)1 << W) - 1);
  Row freeW[RPL];
#pragma unroll
  for (int k = 0; k < RPL; ++k) {
    const int r = lane * RPL + k;
    Row fr = 0;
    if (r < H) {
      const uint32_t* prow = ob + (r + d.P) * d.RW + (d.P >> 5);
      Row wall = __funnelshift_r(prow[0], prow[1], d.P);
      if (RB == 64) wall |= (Row)__funnelshift_r(prow[1], prow[2], d.P) << (RB / 2);
      fr = ~wall & valid;
    }
    freeW[k] = fr;
  }
  for (int r = lane; r < 64; r += 32) rows[r] = 0;
  __syncwarp();
  // visible robots at time i (ids 1..N-1 => 0-based 0..N-2), PRIMAL:523-529
  auto pos_at = [&](int b) { return b < i ? pnew[b] : pold[b]; };
  for (int b0 = 0; b0 < N - 1; b0 += 32) {
    const int b = b0 + lane;
    if (b < N - 1 && b != i) {
      const uchar2 p = pos_at(b);
      if ((int)p.x >= tl0 && (int)p.x < tl0 + F && (int)p.y >= tl1 && (int)p.y < tl1 + F)
        atomicOr((unsigned int*)&rows[p.x] + (sizeof(Row) == 8 ? (p.y >> 5) : 0), 1u << (p.y & 31));
    }
  }
  __syncwarp();
  Row robots[RPL];
#pragma unroll
  for (int k = 0; k < RPL; ++k) robots[k] = (lane * RPL + k < 64) ? rows[lane * RPL + k] : 0;
  int num_blocking = 0;
  for (int b0 = 0; b0 < N - 1; b0 += 32) {
    const int bl = b0 + lane;
    bool vis = false;
    if (bl < N - 1 && bl != i) {
      const uchar2 p = pos_at(bl);
      vis = (int)p.x >= tl0 && (int)p.x < tl0 + F && (int)p.y >= tl1 && (int)p.y < tl1 + F;
    }
    unsigned todo = __ballot_sync(0xffffffffu, vis);
    while (todo) {
      const int b = b0 + __ffs(todo) - 1;
      todo &= todo - 1;
      const uchar2 pb = pos_at(b), gb = goal[b];
      Row fa[RPL], fbef[RPL];
#pragma unroll
      for (int k = 0; k < RPL; ++k) {
        const int r = lane * RPL + k;
        Row rb = robots[k];
        if (r == pb.x) rb &= ~((Row)1 << pb.y);                      // other_locations.remove(pos(agent)), :533
        fa[k] = freeW[k] & ~rb;                                      // robots = other_locations, :538-539
        fbef[k] = fa[k];
        if (r == me.x) fbef[k] &= ~((Row)1 << me.y);                 // ... + [pos(agent_id)], :535-536
      }
      const int before = warp_path_hops<Row, RPL>(fbef, pb.x, pb.y, gb.x, gb.y, lane);
      const int after = warp_path_hops<Row, RPL>(fa, pb.x, pb.y, gb.x, gb.y, lane);
      if (before < 0 && after < 0) continue;                         // :541
      if (before >= 0 && after < 0) continue;                        // :542
      if ((before < 0 && after >= 0) || before > after + 10) ++num_blocking;   // :543-545 (len(path) = hops + 1)
    }
  }
  if (lane == 0) {
    const double x = __dmul_rn((double)num_blocking, d.blocking_cost);   // num_blocking * BLOCKING_COST, :546
    S.last_reward[m] = __dadd_rn(d.goal_reward, x);                      // reward = GOAL_REWARD; reward += x, :581-583
    if (blocking_out) blocking_out[m] = x < 0 ? 1 : 0;                    // :584-585
  }
}

// After the blocking kernel: publish the per-agent rewards and refold the team reward.
__global__ void mapf_blocking_finish_kernel(const MapfDims d, const MapfState S, int agent_lo, int agent_hi,
                                            double* agent_reward, double* reward) {
  const long long total = (long long)d.E * d.N;
  for (long long j = blockIdx.x * (long long)blockDim.x + threadIdx.x; j < total;
       j += (long long)gridDim.x * blockDim.x) {
    if (agent_reward) agent_reward[j] = S.last_reward[j];
    if (reward && (j % d.N) == 0) {
      TreeSum ts;                                                    // the same pairwise order as the tile kernel
      for (int a = 0; a < d.N; ++a) ts.push((a >= agent_lo && a < agent_hi) ? S.last_reward[j + a] : 0.0);
      reward[j / d.N] = ts.finish();
    }
  }
}

// getAstarCosts quirk (PRIMAL:496-498): `costs = state.copy()`, so cells the search never reached keep
// `state`: 0 when free, the agent id when an agent stands there.
__global__ void mapf_primal_costs_agents_kernel(const MapfDims d, const MapfState S, const uint8_t* dirty,
                                                int16_t* dist) {
  const long long total = (long long)d.E * d.N * d.N;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const long long m = idx / d.N;   // (env, map owner)
    const int b = (int)(idx - m * d.N);
    if (dirty && !dirty[m]) continue;
    const int e = (int)(m / d.N);
    const uchar2 p = ((const uchar2*)S.pos)[(size_t)e * d.N + b];
    int16_t* c = dist + (size_t)m * d.HW + (int)p.x * d.W + p.y;
    if (*c == -2) *c = (int16_t)(b + 1);
  }
}
__global__ void mapf_primal_costs_free_kernel(const MapfDims d, const uint8_t* dirty, int16_t* dist) {
  const long long total = (long long)d.E * d.N * d.HW;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    if (dirty && !dirty[idx / d.HW]) continue;
    if (dist[idx] == -2) dist[idx] = 0;
  }
}

// ------------------------------------------------------------------------------------------------
// PARTIAL observation (get_obs_agent, PARTIAL:319-382): T[E][N][2*W*W + 13*K], T = double (the reference's dtype) or
// float (what pymarl's episode batch stores, src/run.py:133-140: the double is rounded once, at the store).
// One block of 128 threads per environment.
//   * walls come straight from the obstacle bit rows (copied to shared memory; they are padded by half a window of
//     wall cells, so a window cell is one unchecked bit test at `agent base + thread offset`), agent counts from a
//     zero-filled byte map padded the same way;
//   * the K-1 nearest agents of every agent (stable order by L2 distance == order by squared integer distance, ties
//     by index; the agent itself has distance H*W, :560-567): one warp per agent, one lane per candidate, every key
//     (distance^2 << 8 | index) computed once, K-1 warp-wide minima (REDUX); the square roots of the selected
//     distances and the goal-vector features are computed afterwards with all lanes busy;
//   * thread q decodes element q of an agent's observation ONCE (window cell or feature slot) and walks the N
//     agents: for one agent consecutive lanes write consecutive elements (coalesced streaming stores).
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void mapf_partial_obs_kernel(const MapfDims d, const MapfState S, T* obs, long long* state_out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int e = blockIdx.x, N = d.N, Wn = d.pW, K = d.pK;
  const int half = Wn / 2, Wp = d.W + Wn, Hp = d.H + Wn;     // count map: `half` empty cells on every side (+1 spare)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  size_t off = 0;
  uint32_t* obw = (uint32_t*)(smem_raw + off);               // [bm_words] obstacle bit rows (padding = wall)
  off += (size_t)d.bm_words * 4;
  uint8_t* cnt = smem_raw + off;                             // [Hp*Wp] agents per cell (0 on walls)
  off += ((size_t)Hp * Wp + 15) & ~(size_t)15;
  double* kdist = (double*)(smem_raw + off);                 // [N][K] distances to the selected agents
  off += (size_t)N * K * 8;
  double* feat = (double*)(smem_raw + off);                  // [N][13]
  off += (size_t)N * 13 * 8;
  int* cbase = (int*)(smem_raw + off);                       // [N] count-map index of the window's top-left cell
  off += ((size_t)N * 4 + 15) & ~(size_t)15;
  int* bbase = (int*)(smem_raw + off);                       // [N] bit index of the same cell in the obstacle rows
  off += ((size_t)N * 4 + 15) & ~(size_t)15;
  uchar2* spos = (uchar2*)(smem_raw + off);                  // [N]
  off += ((size_t)N * 2 + 15) & ~(size_t)15;
  unsigned* kkey = (unsigned*)(smem_raw + off);              // [N][K] selected keys (distance^2 << 8 | index)
  off += ((size_t)N * K * 4 + 15) & ~(size_t)15;
  uint8_t* knn = smem_raw + off;                             // [N][K] ids, 255 = empty row
  off += ((size_t)N * K + 15) & ~(size_t)15;
  unsigned char* dec_raw = smem_raw + off;                  // [posz] u32 decode table of the sector-aligned path
  const uchar2* goal = (const uchar2*)S.goal + (size_t)e * N;
  const uchar2* start = (const uchar2*)S.start + (size_t)e * N;
  const uint32_t* ob = S.obst_bits + (d.shared_map ? (size_t)0 : (size_t)e * d.bm_words);
  const int rowbits = d.RW * 32;
  if (state_out != nullptr && threadIdx.x == blockDim.x - 1) {
    // get_state (PARTIAL:384-393) rides along with the observation when the caller bound an output for it
    // (mapf_partial_bind_state_out): the last thread has the least to do in the phases below
    long long sum = 0;
    for (int i = 0; i < N; ++i) sum += S.goal_cost[(size_t)e * N + i];
    state_out[3 * e] = S.total_coll[e];
    state_out[3 * e + 1] = S.step_count[e];
    state_out[3 * e + 2] = sum;
  }
  for (int i = threadIdx.x; i < (d.bm_words >> 2); i += blockDim.x) ((uint4*)obw)[i] = ((const uint4*)ob)[i];
  for (int i = threadIdx.x; i < ((Hp * Wp + 15) >> 4); i += blockDim.x) ((uint4*)cnt)[i] = make_uint4(0, 0, 0, 0);
  for (int a = threadIdx.x; a < N; a += blockDim.x) {
    const uchar2 p = ((const uchar2*)S.pos)[(size_t)e * N + a];
    spos[a] = p;
    cbase[a] = (int)p.x * Wp + p.y;                          // map cell (i, j) sits at count-map (i + half, j + half)
    bbase[a] = ((int)p.x + d.P - half) * rowbits + (int)p.y + d.P - half;   // d.P >= half (mapf_create)
  }
  __syncthreads();
  for (int a = threadIdx.x; a < N; a += blockDim.x) {
    const uchar2 p = spos[a];
    // `_full_obs` is -1 + k on a wall cell holding k agents (PARTIAL:592): the first agent only lifts the cell to 0
    // (no longer an obstacle, :339; nobody counted, :341), the others count
    const int pr = (int)p.x + d.P, pc = (int)p.y + d.P;
    const uint32_t bit = 1u << (pc & 31);
    uint32_t* wp = &obw[pr * d.RW + (pc >> 5)];
    if ((*wp & bit) == 0 || (atomicAnd(wp, ~bit) & bit) == 0) byte_inc(cnt, cbase[a] + half * Wp + half);
  }
  // K - 1 nearest agents
  const int k_m1 = min(N, K) - 1;
  constexpr unsigned kSelf = 1u << 22, kNone = 0x7fffffffu;  // kSelf > 2 * 254^2
  constexpr int kMaxPerLane = 8;                             // N <= 255
  if (N <= 32 && K <= 32) {
    // at most one candidate per lane and all rows in lanes: no inner loops (the general version below spends half of
    // its instructions on loop control; the selection was 30 % of this kernel's instructions at N = 32)
    for (int a = warp; a < N; a += nwarps) {
      const uchar2 p = spos[a];
      unsigned key = kNone;
      if (lane < N) {
        const uchar2 q = spos[lane];
        const int dx = (int)p.x - (int)q.x, dy = (int)p.y - (int)q.y;
        key = (((lane == a) ? kSelf : (unsigned)(dx * dx + dy * dy)) << 8) | (unsigned)lane;
      }
      unsigned sel = lane == 0 ? (unsigned)a : kNone;        // row 0: knn_agents.insert(0, agent_id), :352
      for (int r = 0; r < k_m1; ++r) {
        const unsigned best = __reduce_min_sync(0xffffffffu, key);
        if (key == best) key = kNone;                        // keys are unique (the index is part of them)
        if (lane == r + 1) sel = best;                       // round r's winner stays in lane r + 1
      }
      if (lane < K) {
        knn[a * K + lane] = sel == kNone ? (uint8_t)255 : (uint8_t)(sel & 255u);
        kkey[a * K + lane] = sel;
      }
    }
  } else
  for (int a = warp; a < N; a += nwarps) {
    const uchar2 p = spos[a];
    unsigned key[kMaxPerLane];
#pragma unroll
    for (int i = 0; i < kMaxPerLane; ++i) {
      const int b = lane + 32 * i;
      key[i] = kNone;
      if (b < N) {
        const uchar2 q = spos[b];
        const int dx = (int)p.x - (int)q.x, dy = (int)p.y - (int)q.y;
        key[i] = (((b == a) ? kSelf : (unsigned)(dx * dx + dy * dy)) << 8) | (unsigned)b;
      }
      if (32 * (i + 1) >= N) break;
    }
    // round r's winner stays in lane r + 1; one parallel store after the rounds
    unsigned sel = lane == 0 ? (unsigned)a : kNone;          // row 0: knn_agents.insert(0, agent_id), :352
    for (int r = 0; r < k_m1; ++r) {
      unsigned mine = kNone;
#pragma unroll
      for (int i = 0; i < kMaxPerLane; ++i) {
        mine = min(mine, key[i]);
        if (32 * (i + 1) >= N) break;
      }
      const unsigned best = __reduce_min_sync(0xffffffffu, mine);
#pragma unroll
      for (int i = 0; i < kMaxPerLane; ++i) {
        if (key[i] == best) key[i] = kNone;                  // keys are unique (the index is part of them)
        if (32 * (i + 1) >= N) break;
      }
      if (r < 31) {
        if (lane == r + 1) sel = best;
      } else if (lane == 0) {                                // more than 32 rows (K > 32): stored as they come
        knn[a * K + r + 1] = best == kNone ? (uint8_t)255 : (uint8_t)(best & 255u);
        kkey[a * K + r + 1] = best;
      }
    }
    for (int r = lane; r < K; r += 32) {
      if (r >= 32 && r <= k_m1) continue;                    // written by its round
      const unsigned v = r < 32 ? sel : kNone;               // rows past the last round are empty
      knn[a * K + r] = v == kNone ? (uint8_t)255 : (uint8_t)(v & 255u);
      kkey[a * K + r] = v;
    }
  }
  // positions, goal vectors and counters of every agent: one thread per agent (:353-371; column 11 = kdist)
  for (int na = threadIdx.x; na < N; na += blockDim.x) {
    const uchar2 q0 = spos[na], g = goal[na], st = start[na];
    const int dx = (int)g.x - (int)q0.x, dy = (int)g.y - (int)q0.y;
    const double norm = __dsqrt_rn((double)(dx * dx + dy * dy));   // __update_goal_vectors, :957-972
    double* f = feat + na * 13;
    f[0] = q0.x;
    f[1] = q0.y;
    f[2] = st.x;
    f[3] = st.y;
    f[4] = g.x;
    f[5] = g.y;
    f[6] = norm != 0.0 ? __ddiv_rn((double)dx, norm) : 0.0;
    f[7] = norm != 0.0 ? __ddiv_rn((double)dy, norm) : 0.0;
    f[8] = norm;
    f[9] = S.pnode[(size_t)e * N + na];
    f[10] = S.pedge[(size_t)e * N + na];
    f[11] = 0.0;
    f[12] = S.agent_steps[(size_t)e * N + na];
  }
  __syncthreads();
  // the window rows of the wall map, once per (agent, row): one funnel shift here instead of an index computation, a
  // word load and a variable shift per ELEMENT in the output loops below (windows of up to 16 cells).  float64 only
  // (233 -> 224 us at c3 shape); the float32 instantiation measured 3 % slower with it and keeps the direct bit test.
  const bool rows_ok = sizeof(T) == 8 && Wn <= 16;
  uint16_t* rows16 = (uint16_t*)(dec_raw + (((size_t)d.posz * 4 + 15) & ~(size_t)15));   // [N][Wn]
  if (rows_ok)
    for (int it = threadIdx.x; it < N * Wn; it += blockDim.x) {
      const int a = it / Wn, wi = it - a * Wn;
      const int bit = bbase[a] + wi * rowbits;
      rows16[it] = (uint16_t)(__funnelshift_r(obw[bit >> 5], obw[(bit >> 5) + 1], bit) & ((1u << Wn) - 1u));
    }
  for (int q = threadIdx.x; q < N * K; q += blockDim.x) {    // keys -> distances, all lanes busy
    const unsigned v = kkey[q];
    // row 0 is the agent itself: distance H*W (:560-567); empty rows: -1
    kdist[q] = (q % K == 0) ? (double)d.HW : (v == kNone ? -1.0 : __dsqrt_rn((double)(v >> 8)));
  }
  __syncthreads();
  const int osz = d.posz, ww = Wn * Wn;
  T* out = obs + (size_t)e * N * osz;
  if (sizeof(T) == 8 && (osz & 3) != 0) {
    // float64 with agent blocks that do not start on 32-byte sectors (c3-shaped: 307 elements): a warp's 256-byte
    // store would straddle sectors at both ends, and those partial-sector writes hold the kernel at 4.3 TB/s (blocks
    // of 4k elements reach 5.6-6.3 TB/s, profiles/partial_align_probe.py).  Here the element -> source decode goes
    // through a table in shared memory, and for every agent the threads are shifted so that each warp's store starts
    // on a sector of the output.
    uint32_t* dec = (uint32_t*)dec_raw;                     // [osz]: kind << 28 | payload
    for (int idx = threadIdx.x; idx < osz; idx += blockDim.x) {
      uint32_t v;
      if (idx < 2 * ww) {
        const bool agents_map = idx >= ww;
        const int c = agents_map ? idx - ww : idx;
        const int wi = c / Wn, wj = c - wi * Wn;
        v = agents_map ? ((1u << 28) | (uint32_t)(wi * Wp + wj))
                       : (rows_ok ? (uint32_t)((wi << 8) | wj) : (uint32_t)(wi * rowbits + wj));
      } else {
        const int f0 = idx - 2 * ww;
        const int row = f0 / 13, f = f0 - 13 * row;
        v = ((f == 11 ? 3u : 2u) << 28) | (uint32_t)(f == 11 ? row : row * 16 + f);
      }
      dec[idx] = v;
    }
    __syncthreads();
    // agents a = 4j + k share their shift (4j * osz elements are a whole number of sectors): for each k the thread
    // decodes its element once and walks the agents k, k + 4, ... exactly like the unshifted loops below
    const size_t step4 = (size_t)4 * osz;
    for (int base = 0; base < osz + 3; base += blockDim.x) {
#pragma unroll 1
      for (int k = 0; k < 4 && k < N; ++k) {
        double* o = (double*)out + (size_t)k * osz;
        const int idx = base + (int)threadIdx.x - (int)(((uintptr_t)o >> 3) & 3);   // elements past the last sector start
        if (idx < 0 || idx >= osz) continue;
        o += idx;
        const uint32_t dv = dec[idx];
        const int kind = (int)(dv >> 28), pay = (int)(dv & 0x0fffffffu);
        if (kind == 0 && rows_ok) {                          // 1 on walls and outside the map
          const uint16_t* rp = rows16 + (pay >> 8);
          const int wj = pay & 255;
#pragma unroll kPartialUnroll64
          for (int a = k; a < N; a += 4, o += step4) __stcs(o, (double)(((uint32_t)rp[a * Wn] >> wj) & 1u));
        } else if (kind == 0) {
#pragma unroll kPartialUnroll64
          for (int a = k; a < N; a += 4, o += step4) {
            const int bit = bbase[a] + pay;
            __stcs(o, (double)((obw[bit >> 5] >> (bit & 31)) & 1u));
          }
        } else if (kind == 1) {                              // agents standing on the cell
          const uint8_t* cp = cnt + pay;
#pragma unroll kPartialUnroll64
          for (int a = k; a < N; a += 4, o += step4) __stcs(o, (double)(int)cp[cbase[a]]);
        } else if (kind == 3) {                              // distance to the row's agent
#pragma unroll kPartialUnroll64
          for (int a = k; a < N; a += 4, o += step4)
            __stcs(o, (knn[a * K + pay] == 255) ? -1.0 : kdist[a * K + pay]);
        } else {                                             // the row's agent's own features
          const int row = pay >> 4, f = pay & 15;
#pragma unroll kPartialUnroll64
          for (int a = k; a < N; a += 4, o += step4) {
            const int na = knn[a * K + row];
            __stcs(o, (na == 255) ? -1.0 : feat[na * 13 + f]);
          }
        }
      }
    }
    return;
  }
  for (int idx = threadIdx.x; idx < osz; idx += blockDim.x) {
    T* o = out + idx;
    if (idx < 2 * ww) {                                      // the two W x W maps, :326-342
      const bool agents_map = idx >= ww;
      const int c = agents_map ? idx - ww : idx;
      const int wi = c / Wn, wj = c - wi * Wn;
      if (!agents_map && rows_ok) {                          // 1 on walls and outside the map
        const uint16_t* rp = rows16 + wi;
#pragma unroll kPartialUnroll
        for (int a = 0; a < N; ++a, o += osz, rp += Wn) __stcs(o, (T)(((uint32_t)*rp >> wj) & 1u));
      } else if (!agents_map) {
        const int boff = wi * rowbits + wj;
#pragma unroll kPartialUnroll
        for (int a = 0; a < N; ++a, o += osz) {
          const int bit = bbase[a] + boff;
          __stcs(o, (T)((obw[bit >> 5] >> (bit & 31)) & 1u));
        }
      } else {                                               // agents standing on the cell
        const uint8_t* cp = cnt + wi * Wp + wj;
#pragma unroll kPartialUnroll
        for (int a = 0; a < N; ++a, o += osz) __stcs(o, (T)(int)cp[cbase[a]]);
      }
    } else {                                                 // K x 13 features, :344-371
      const int f0 = idx - 2 * ww;
      const int row = f0 / 13, f = f0 - 13 * row;
      const double* src = (f == 11) ? kdist + row : feat + f;
      const int stride = (f == 11) ? K : 0;
#pragma unroll kPartialUnroll
      for (int a = 0; a < N; ++a, o += osz) {
        const int na = knn[a * K + row];
        // column 11 is the observer's distance to the row's agent, the others are that agent's own features
        const double v = (na == 255) ? -1.0 : ((f == 11) ? src[a * stride] : src[na * 13]);
        __stcs(o, (T)v);
      }
    }
  }
}

// get_state, PARTIAL:384-393: [total collisions, step count, sum of per-agent goal costs].
__global__ void mapf_partial_state_kernel(const MapfDims d, const MapfState S, long long* state) {
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < d.E; e += gridDim.x * blockDim.x) {
    long long sum = 0;
    for (int i = 0; i < d.N; ++i) sum += S.goal_cost[(size_t)e * d.N + i];
    state[3 * e] = S.total_coll[e];
    state[3 * e + 1] = S.step_count[e];
    state[3 * e + 2] = sum;
  }
}

// In-kernel rollouts (ROLL) exist for the modes whose whole step, observation included, is this one kernel: PRIMAL
// (without diagonal movement) and GRID.  PARTIAL observes in its own kernel, the diagonal mode is off the hot path.
template <int MODE>
constexpr bool kHasRoll = (MODE == MAPF_MODE_PRIMAL || MODE == MAPF_MODE_GRID);

template <int F, int MODE>
cudaError_t launch_tile_f(const MapfDims& d, const MapfTileLayout& L, const MapfState& S, const MapfTileArgs& A,
                          cudaStream_t st) {
  const int grid = (d.E + d.epb - 1) / d.epb;
  constexpr bool kSingle = MODE != MAPF_MODE_PRIMAL_DIAG;   // the diagonal mode only has the looped instantiation
  if (A.T > 1) {
    if constexpr (kHasRoll<MODE>) {
      if (d.epb * d.N > kThreads) return cudaErrorInvalidValue;
      if (grid <= 148 * 8)   // every tile of the batch is resident at 8 blocks per SM: the wide-register build
        mapf_tile_kernel<F, MODE, true, true, true><<<grid, kThreads, L.total_bytes, st>>>(d, L, S, A);
      else
        mapf_tile_kernel<F, MODE, true, true, false><<<grid, kThreads, L.total_bytes, st>>>(d, L, S, A);
      return cudaGetLastError();
    } else {
      return cudaErrorInvalidValue;
    }
  }
  if (kSingle && d.epb * d.N <= kThreads)
    mapf_tile_kernel<F, MODE, kSingle, false, false><<<grid, kThreads, L.total_bytes, st>>>(d, L, S, A);
  else
    mapf_tile_kernel<F, MODE, false, false, false><<<grid, kThreads, L.total_bytes, st>>>(d, L, S, A);
  return cudaGetLastError();
}

template <int F, int MODE>
cudaError_t configure_tile_f(int smem_bytes) {
  const auto attr = cudaFuncAttributeMaxDynamicSharedMemorySize;
  cudaError_t e = cudaFuncSetAttribute(mapf_tile_kernel<F, MODE, false, false>, attr, smem_bytes);
  if (e == cudaSuccess && MODE != MAPF_MODE_PRIMAL_DIAG)
    e = cudaFuncSetAttribute(mapf_tile_kernel<F, MODE, (MODE != MAPF_MODE_PRIMAL_DIAG), false>, attr, smem_bytes);
  if constexpr (kHasRoll<MODE>) {
    if (e == cudaSuccess) e = cudaFuncSetAttribute(mapf_tile_kernel<F, MODE, true, true, false>, attr, smem_bytes);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(mapf_tile_kernel<F, MODE, true, true, true>, attr, smem_bytes);
  }
  return e;
}

int grid_for(long long total, int block) {
  long long g = (total + block - 1) / block;
  if (g > 148LL * 32) g = 148LL * 32;   // grid-stride loops; a multiple of the 148 SMs
  if (g < 1) g = 1;
  return (int)g;
}

}  // namespace

#ifdef MAPF_PHASE_TIMING
// debug builds only (-DMAPF_PHASE_TIMING): clock64() at the phase boundaries of the middle block, see
// profiles/phase_probe.py
extern "C" int mapf_debug_phase_clocks(long long* out32) {
  return (int)cudaMemcpyFromSymbol(out32, g_phase_clk, sizeof(long long) * 32);
}
#endif

// Specialised tile kernels: PRIMAL with a field of view F (and F = 0: step only, the observation of an unlisted F
// comes from the generic kernel), GRID and PARTIAL without a window.
#define MAPF_FOR_EACH_FOV(X) X(3) X(5) X(7) X(9) X(10) X(11)

// 1 when mapf_launch_tile accepts A.T > 1 for this mode (the caller also needs epb * N <= MAPF_TILE_THREADS).
extern "C" int mapf_tile_has_rollout(int mode) {
  return (mode == MAPF_MODE_PRIMAL || mode == MAPF_MODE_GRID) ? 1 : 0;
}

extern "C" int mapf_tile_has_fov(int F) {
  switch (F) {
#define X(f) case f:
    MAPF_FOR_EACH_FOV(X)
#undef X
    return 1;
    default:
      return 0;
  }
}

extern "C" int mapf_configure_tile(int F, int mode, int smem_bytes) {
  cudaError_t err = cudaErrorInvalidValue;
  if (mode == MAPF_MODE_GRID) err = configure_tile_f<0, MAPF_MODE_GRID>(smem_bytes);
  else if (mode == MAPF_MODE_PARTIAL) err = configure_tile_f<0, MAPF_MODE_PARTIAL>(smem_bytes);
  else if (mode == MAPF_MODE_PRIMAL_DIAG) {
    switch (F) {
      case 0: err = configure_tile_f<0, MAPF_MODE_PRIMAL_DIAG>(smem_bytes); break;
#define X(f) case f: err = configure_tile_f<f, MAPF_MODE_PRIMAL_DIAG>(smem_bytes); break;
      MAPF_FOR_EACH_FOV(X)
#undef X
      default: break;
    }
  } else {
    switch (F) {
      case 0: err = configure_tile_f<0, MAPF_MODE_PRIMAL>(smem_bytes); break;
#define X(f) case f: err = configure_tile_f<f, MAPF_MODE_PRIMAL>(smem_bytes); break;
      MAPF_FOR_EACH_FOV(X)
#undef X
      default: break;
    }
  }
  return (int)err;
}

extern "C" int mapf_launch_tile(const MapfDims& d, const MapfTileLayout& L, const MapfState& S, const MapfTileArgs& A,
                                void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (d.mode == MAPF_MODE_GRID) return (int)launch_tile_f<0, MAPF_MODE_GRID>(d, L, S, A, st);
  if (d.mode == MAPF_MODE_PARTIAL) return (int)launch_tile_f<0, MAPF_MODE_PARTIAL>(d, L, S, A, st);
  const int F = (d.obs_mode == MAPF_OBS_PRIMAL_FOV) ? d.F : 0;
  if (d.diag) {
    switch (F) {
      case 0: return (int)launch_tile_f<0, MAPF_MODE_PRIMAL_DIAG>(d, L, S, A, st);
#define X(f) case f: return (int)launch_tile_f<f, MAPF_MODE_PRIMAL_DIAG>(d, L, S, A, st);
      MAPF_FOR_EACH_FOV(X)
#undef X
      default:
        return (int)cudaErrorInvalidValue;
    }
  }
  switch (F) {
    case 0: return (int)launch_tile_f<0, MAPF_MODE_PRIMAL>(d, L, S, A, st);
#define X(f) case f: return (int)launch_tile_f<f, MAPF_MODE_PRIMAL>(d, L, S, A, st);
    MAPF_FOR_EACH_FOV(X)
#undef X
    default:
      return (int)cudaErrorInvalidValue;
  }
}

// ---- the pipelined rollout kernel (mapf_pipe_kernel): which handles qualify, and its launch
static int pipe_epb(const MapfDims& d) { return d.N <= 32 ? 32 / d.N : 0; }

static bool pipe_layout(const MapfDims& d, int epb, MapfPipeLayout* L) {
  int off = 0;
  auto take = [&](int bytes) {
    const int o = off;
    off = (off + bytes + 15) / 16 * 16;
    return o;
  };
  L->obst_off = take((d.shared_map ? 1 : epb) * d.bm_words * 4);
  L->guard_off[0] = take(16);
  L->grid_off = take(epb * d.grid_bytes);
  L->agt_off = take(epb * d.bm_words * 4);
  L->live_bytes = off - L->grid_off;
  L->guard_off[1] = take(16);
  L->snap_off[0] = take(L->live_bytes);
  L->guard_off[2] = take(16);
  L->snap_off[1] = take(L->live_bytes);
  L->guard_off[3] = take(16);
  for (int b = 0; b < 2; ++b) L->snappos_off[b] = take(64);
  for (int b = 0; b < 2; ++b) L->snapact_off[b] = take(32);
  L->goal_off = take(64);
  L->posold_off = take(64);
  L->posnew_off = take(64);
  L->mv_off = take(128);
  L->res_off = take(32);
  L->dep_off = take(32);
  L->act_off = take(32);
  L->status_off = take(32);
  L->done_off = take(32);
  L->flag_off = take(32);
  L->rew_off = take(256);
  L->envrew_off = take(256);
  L->envcnt_off = take(128);
  L->guard_off[4] = take(16);
  L->str_off = take(((32 + d.G - 1) / d.G) * d.GW * 4 + 16);
  L->guard_off[5] = take(16);
  L->guard_off[6] = L->guard_off[5];
  L->guard_off[7] = L->guard_off[5];
  L->total_bytes = off;
  return off <= 48 * 1024;
}

// 1 when mapf_rollout can run this handle's steps through the pipelined kernel (5-action PRIMAL with an odd specialised
// field of view, at most 32 agents per environment, tiles of whole 8-agent string groups, tile state within 48 KB).
extern "C" int mapf_pipe_supported(const MapfDims& d) {
  if (d.mode != MAPF_MODE_PRIMAL || d.diag || d.blocking || d.obs_mode != MAPF_OBS_PRIMAL_FOV) return 0;
  if (!mapf_tile_has_fov(d.F) || (d.F & 1) == 0 || d.F < 3) return 0;
  const int epb = pipe_epb(d);
  if (epb < 1 || (epb * d.N) % 8 != 0 || d.G != 8) return 0;
  MapfPipeLayout L;
  return pipe_layout(d, epb, &L) ? 1 : 0;
}

extern "C" int mapf_pipe_tiles(const MapfDims& d) { return (d.E + pipe_epb(d) - 1) / pipe_epb(d); }

extern "C" int mapf_launch_pipe(const MapfDims& d, const MapfState& S, const MapfTileArgs& A, void* stream) {
  const int epb = pipe_epb(d);
  MapfPipeLayout L;
  if (epb < 1 || !pipe_layout(d, epb, &L)) return (int)cudaErrorInvalidValue;
  const int grid = (d.E + epb - 1) / epb;
  cudaStream_t st = (cudaStream_t)stream;
  switch (d.F) {
#define X(f)                                                                              \
  case f:                                                                                 \
    if constexpr ((f & 1) == 1) {                                                         \
      mapf_pipe_kernel<f><<<grid, kThreads, L.total_bytes, st>>>(d, L, S, A, epb);        \
      return (int)cudaGetLastError();                                                     \
    }                                                                                     \
    return (int)cudaErrorInvalidValue;
    MAPF_FOR_EACH_FOV(X)
#undef X
    default:
      return (int)cudaErrorInvalidValue;
  }
}

extern "C" int mapf_launch_observe_generic(const MapfDims& d, const MapfState& S, uint8_t* obs_u8, float* obs_f32,
                                           double* vec, void* stream) {
  const long long total = (long long)d.E * d.N * d.F * d.F;
  mapf_observe_generic_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(d, S, obs_u8, obs_f32, vec);
  return (int)cudaGetLastError();
}

extern "C" int mapf_launch_build_obst(const MapfDims& d, const MapfState& S, const int8_t* map,
                                      const uint8_t* env_mask, void* stream) {
  const long long total = (long long)(d.shared_map ? 1 : d.E) * d.bm_words;
  mapf_build_obst_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(d, S.obst_bits, map, env_mask);
  return (int)cudaGetLastError();
}

extern "C" int mapf_launch_reset(const MapfDims& d, const MapfState& S, const int16_t* starts, const int16_t* goals,
                                 const uint8_t* env_mask, void* stream) {
  mapf_reset_kernel<<<grid_for((long long)d.E * d.N, 256), 256, 0, (cudaStream_t)stream>>>(d, S, starts, goals,
                                                                                           env_mask);
  return (int)cudaGetLastError();
}

extern "C" int mapf_launch_set_goals(const MapfDims& d, const MapfState& S, const int16_t* goals,
                                     const uint8_t* dirty, void* stream) {
  mapf_set_goals_kernel<<<grid_for((long long)d.E * d.N, 256), 256, 0, (cudaStream_t)stream>>>(d, S, goals, dirty);
  return (int)cudaGetLastError();
}

// Bookkeeping of a rollout loop around one vector-env step (pymarl ParallelRunner.run, parallel_runner.py:123-175), two
// kernels instead of a dozen element-wise launches.  (1) actions of finished environments become the STAY action;
// the result goes to the engine's uint8 input and, optionally, to the episode batch's int64 storage.
__global__ void mapf_runner_mask_actions_kernel(const MapfDims d, const void* __restrict__ actions, int i64,
                                                const uint8_t* __restrict__ alive, int stay, uint8_t* __restrict__ out8,
                                                long long* __restrict__ out64) {
  const long long total = (long long)d.E * d.N;
  for (long long j = blockIdx.x * (long long)blockDim.x + threadIdx.x; j < total;
       j += (long long)gridDim.x * blockDim.x) {
    const int e = (int)((unsigned long long)j / (unsigned)d.N);
    long long a = i64 ? ((const long long*)actions)[j] : (long long)((const uint8_t*)actions)[j];
    if (!alive[e]) a = stay;
    out8[j] = (uint8_t)a;          // out-of-range values still trip the step kernel's action check (low byte kept)
    if (i64 && (a < 0 || a > 255)) out8[j] = 255;
    if (out64) out64[j] = a;
  }
}
// (2) returns / lengths / filled / alive after the step: an environment that was running accumulates the reward and a
// step, its next time slot is marked filled, and it stops running when the step terminated it.
__global__ void mapf_runner_account_kernel(int E, const double* __restrict__ reward, const uint8_t* __restrict__ term,
                                           uint8_t* __restrict__ alive, double* __restrict__ returns,
                                           long long* __restrict__ lengths, uint8_t* __restrict__ filled_next) {
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < E; e += gridDim.x * blockDim.x) {
    const uint8_t al = alive[e];
    if (al) {
      returns[e] = __dadd_rn(returns[e], reward[e]);
      lengths[e] += 1;
    }
    filled_next[e] = al;
    alive[e] = (uint8_t)(al && term[e] == 0);
  }
}

extern "C" int mapf_launch_runner_mask_actions(const MapfDims& d, const void* actions, int i64, const uint8_t* alive,
                                               int stay, uint8_t* out8, long long* out64, void* stream) {
  mapf_runner_mask_actions_kernel<<<grid_for((long long)d.E * d.N, 256), 256, 0, (cudaStream_t)stream>>>(
      d, actions, i64, alive, stay, out8, out64);
  return (int)cudaGetLastError();
}

extern "C" int mapf_launch_runner_account(const MapfDims& d, const double* reward, const uint8_t* term, uint8_t* alive,
                                          double* returns, long long* lengths, uint8_t* filled_next, void* stream) {
  mapf_runner_account_kernel<<<grid_for(d.E, 256), 256, 0, (cudaStream_t)stream>>>(d.E, reward, term, alive, returns,
                                                                                    lengths, filled_next);
  return (int)cudaGetLastError();
}

extern "C" int mapf_launch_pop_goals(const MapfDims& d, const MapfState& S, const int16_t* queue, int32_t* head,
                                     int queue_len, uint8_t* dirty, void* stream) {
  mapf_pop_goals_kernel<<<grid_for((long long)d.E * d.N, 256), 256, 0, (cudaStream_t)stream>>>(d, S, queue, head,
                                                                                                queue_len, dirty);
  return (int)cudaGetLastError();
}

extern "C" int mapf_launch_blocking(const MapfDims& d, const MapfState& S, int agent_lo, int agent_hi,
                                    const mapf_step_out& out, void* stream, int* n_launches) {
  cudaStream_t st = (cudaStream_t)stream;
  const long long maps = (long long)d.E * d.N;
  const int w2 = 8;
  const long long g2 = (maps + w2 - 1) / w2;
  const bool wide = d.W > 32, tall = d.H > 32;
#define BLK_LAUNCH(ROW, RPL)                                                                              \
  mapf_blocking_kernel<ROW, RPL><<<(unsigned)g2, w2 * 32, (size_t)w2 * 64 * sizeof(ROW), st>>>(           \
      d, S, agent_lo, agent_hi, out.blocking_dev, w2)
  if (!wide && !tall) BLK_LAUNCH(uint32_t, 1);
  else if (!wide && tall) BLK_LAUNCH(uint32_t, 2);
  else if (wide && !tall) BLK_LAUNCH(unsigned long long, 1);
  else BLK_LAUNCH(unsigned long long, 2);
#undef BLK_LAUNCH
  cudaError_t err = cudaGetLastError();
  *n_launches = 1;
  if (err == cudaSuccess && (out.agent_reward_dev || out.reward_dev)) {
    mapf_blocking_finish_kernel<<<grid_for(maps, 256), 256, 0, st>>>(d, S, agent_lo, agent_hi, out.agent_reward_dev,
                                                                     out.reward_dev);
    err = cudaGetLastError();
    *n_launches = 2;
  }
  return (int)err;
}

extern "C" int mapf_launch_partial_obs(const MapfDims& d, const MapfState& S, void* obs, int f32, long long* state_out,
                                       void* stream) {
  const size_t smem = (size_t)d.bm_words * 4 + (((size_t)(d.H + d.pW) * (d.W + d.pW) + 15) & ~(size_t)15) +
                      (size_t)d.N * d.pK * 8 + (size_t)d.N * 13 * 8 + 2 * (((size_t)d.N * 4 + 15) & ~(size_t)15) +
                      (((size_t)d.N * 2 + 15) & ~(size_t)15) + (((size_t)d.N * d.pK * 4 + 15) & ~(size_t)15) +
                      (((size_t)d.N * d.pK + 15) & ~(size_t)15) + (((size_t)d.posz * 4 + 15) & ~(size_t)15) +
                      (((size_t)d.N * d.pW * 2 + 15) & ~(size_t)15) + 16;   // + the decode table, the window rows
  const int threads = 128;   // more threads per block were measured slower (fewer blocks overlap their serial phases)
  if (smem > 48 * 1024) {
    cudaError_t e = f32 ? cudaFuncSetAttribute(mapf_partial_obs_kernel<float>,
                                               cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)
                        : cudaFuncSetAttribute(mapf_partial_obs_kernel<double>,
                                               cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
  }
  if (f32) mapf_partial_obs_kernel<float><<<d.E, threads, smem, (cudaStream_t)stream>>>(d, S, (float*)obs, state_out);
  else mapf_partial_obs_kernel<double><<<d.E, threads, smem, (cudaStream_t)stream>>>(d, S, (double*)obs, state_out);
  return (int)cudaGetLastError();
}

extern "C" int mapf_launch_partial_state(const MapfDims& d, const MapfState& S, long long* state, void* stream) {
  mapf_partial_state_kernel<<<grid_for(d.E, 256), 256, 0, (cudaStream_t)stream>>>(d, S, state);
  return (int)cudaGetLastError();
}

extern "C" int mapf_launch_random_actions(const MapfDims& d, const uint8_t* avail, uint32_t seed, uint32_t step,
                                          long long env_offset, void* out, int i64, void* stream) {
  mapf_random_actions_kernel<<<grid_for((long long)d.E * d.N, 256), 256, 0, (cudaStream_t)stream>>>(
      d, avail, seed, step, env_offset, i64 ? nullptr : (uint8_t*)out, i64 ? (long long*)out : nullptr);
  return (int)cudaGetLastError();
}

extern "C" int mapf_launch_export16(const MapfDims& d, const uint8_t* src, int16_t* dst, void* stream) {
  const long long n = (long long)d.E * d.N * 2;
  mapf_export16_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(n, src, dst);
  return (int)cudaGetLastError();
}

extern "C" int mapf_launch_bfs(const MapfDims& d, const MapfState& S, const uint8_t* dirty, const uint8_t* env_mask,
                               int16_t* dist, int primal_costs, void* stream, int* n_launches, int32_t* ext_list,
                               int32_t* ext_cnt) {
  cudaStream_t st = (cudaStream_t)stream;
  const int RWB = (d.W + 31) / 32;
  const size_t bm_bytes = (size_t)4 * d.H * RWB * 4;
  const size_t dist_bytes = ((size_t)d.HW * 2 + 15) & ~(size_t)15;
  int stage = 1, warps = 8;
  while (warps > 1 && (bm_bytes + dist_bytes) * warps > 200 * 1024) warps >>= 1;
  if ((bm_bytes + dist_bytes) * warps > 200 * 1024) {
    stage = 0;
    warps = 8;
    while (warps > 1 && bm_bytes * warps > 200 * 1024) warps >>= 1;
  }
  const size_t smem = (bm_bytes + (stage ? dist_bytes : 0)) * warps;
  if (smem > 48 * 1024) {   // per device and cheap: set it on every launch that needs it
    cudaError_t e = cudaFuncSetAttribute(mapf_bfs_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
  }
  const long long maps = (long long)d.E * d.N;
  const int conn8 = (primal_costs && d.diag) ? 1 : 0;   // getAstarCosts with DIAGONAL_MOVEMENT, PRIMAL:421-437
  // S.bfs_list: [0] number of flagged maps, [1] number of overflowed maps, then the two lists (maps entries each)
  int32_t* cnt = ext_cnt ? ext_cnt : S.bfs_list;
  int32_t* flagged = S.bfs_list + 2;
  int32_t* overflow = flagged + maps;
  if (!ext_cnt) {
    cudaError_t e = cudaMemsetAsync(cnt, 0, 8, st);
    if (e != cudaSuccess) return (int)e;
  }
  // masked launch: compact the flagged maps, then a resident grid walks the list
  const int32_t* list = nullptr;
  int extra = 0;
  if (ext_list) {   // the list was written by the step kernel (lifelong hand-out): nothing to compact
    list = ext_list;
  } else if (dirty || env_mask) {
    mapf_bfs_compact_kernel<<<grid_for(maps, 256), 256, 0, st>>>(d, dirty, env_mask, S.bfs_list);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return (int)e;
    list = flagged;
    extra = 1;
  }
  const long long resident = 148 * 8;
  if (d.W <= 64 && d.H <= 64) {
    // register-resident warp-synchronous kernel (no shared memory); maps deeper than 255 levels go to the generic
    // kernel through the overflow list (normally empty: that launch ends at once)
    const int w2 = 8;
    long long g2 = (maps + w2 - 1) / w2;
    // (a resident grid for the unmasked launch too was measured: 0.554 ms against 0.529 ms at c3 -- one map per warp
    // balances better than 442 maps per resident warp)
    if (list) g2 = g2 < resident ? g2 : resident;
    const bool wide = d.W > 32, tall = d.H > 32;
#define BFS_LAUNCH(ROW, RPL)                                                                                          \
  do {                                                                                                                \
    if (conn8)                                                                                                        \
      mapf_bfs_warp_kernel<ROW, RPL, true><<<(unsigned)g2, w2 * 32, 0, st>>>(d, S, list, cnt, overflow, dist, w2);    \
    else                                                                                                              \
      mapf_bfs_warp_kernel<ROW, RPL, false><<<(unsigned)g2, w2 * 32, 0, st>>>(d, S, list, cnt, overflow, dist, w2);   \
  } while (0)
    if (!wide && !tall) BFS_LAUNCH(uint32_t, 1);
    else if (!wide && tall) BFS_LAUNCH(uint32_t, 2);
    else if (wide && !tall) BFS_LAUNCH(unsigned long long, 1);
    else BFS_LAUNCH(unsigned long long, 2);
#undef BFS_LAUNCH
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return (int)e;
    long long go = (maps + warps - 1) / warps;
    go = go < 148 * 2 ? go : 148 * 2;
    mapf_bfs_kernel<<<(unsigned)go, warps * 32, smem, st>>>(d, S, overflow, cnt + 1, dist, RWB, warps, stage, conn8);
    extra += 1;
  } else {
    long long grid = (maps + warps - 1) / warps;
    if (list) grid = grid < resident ? grid : resident;
    mapf_bfs_kernel<<<(unsigned)grid, warps * 32, smem, st>>>(d, S, list, cnt, dist, RWB, warps, stage, conn8);
  }
  cudaError_t err = cudaGetLastError();
  *n_launches = 1 + extra;
  if (err == cudaSuccess && ext_cnt) err = cudaMemsetAsync(ext_cnt, 0, 8, st);   // the list is consumed
  if (err == cudaSuccess && primal_costs) {
    mapf_primal_costs_agents_kernel<<<grid_for(maps * d.N, 256), 256, 0, st>>>(d, S, dirty, dist);
    mapf_primal_costs_free_kernel<<<grid_for(maps * d.HW, 256), 256, 0, st>>>(d, dirty, dist);
    err = cudaGetLastError();
    *n_launches = 3 + extra;
  }
  return (int)err;
}
