// mapf_internal.h -- structures shared by the host side (mapf_capi.cu) and the kernels.
// Not part of the public ABI (include/mapf_b200.h is).
#pragma once
#include <stdint.h>

#include "../../include/mapf_b200.h"

#ifndef MAPF_TILE_THREADS
#define MAPF_TILE_THREADS 128
#endif
#define MAPF_MAX_AGENTS 255
#define MAPF_MAX_SIDE 255
#define MAPF_MODE_PRIMAL_DIAG 3   // internal: MAPF_MODE_PRIMAL with cfg.diagonal_movement (its own kernel instantiation)

// Problem dimensions and constants, passed by value to every kernel.
struct MapfDims {
  int E, N, H, W, HW;
  int F;         // FOV side (0 when the obs mode has no window)
  int P;         // padding of the bitmaps on every side: max(F/2, 1)
  int PR;        // padded rows  = H + 2P
  int RW;        // 32-bit words per padded row, including one guard word
  int bm_words;  // words per padded bitmap, rounded up to a multiple of 4 (16-byte rows for bulk copies)
  int GS;          // row stride of the occupancy grids = W + 2 (one cell of padding on every side)
  int grid_bytes;  // (H+2)*(W+2) rounded up to a multiple of 16
  int shared_map;
  int mode, obs_mode;
  int episode_limit;
  int epb;       // environments per thread block (tile)
  int G;         // agents per bit-string group (group length is a whole number of 32-bit words)
  int GW;        // words per group string = G * 4*F*F / 32
  int sum_mode, step_is_int, collide_is_int;
  int collect_stats;
  int rsum_mode;         // PRIMAL team reward tree: 1 = N divides 32 (whole tree inside a warp), 2 = 32 divides N
                         // (32-agent blocks inside warps, the rest by one thread), 0 = one thread per environment
  int blocking;          // PRIMAL blocking reward enabled
  int diag;              // PRIMAL DIAGONAL_MOVEMENT
  int nact;              // 5, or 9 with diagonal movement
  double blocking_cost;
  uint32_t invN, invW;  // ceil(2^32 / N), ceil(2^32 / W): exact division of values < 65536 by IMAD.HI
  double step_reward, collide_reward;
  double action_cost, idle_cost, goal_reward, collision_reward;
  // MAPF_MODE_PARTIAL
  int pW, pK, posz;      // obs window, K nearest agents, obs size = 2*W*W + 13*K
  int complete_len;
  double p_move, p_stay, p_stay_goal, p_nc, p_ec, p_envc;
  double inv_limit;      // 1.0 / episode_limit, correctly rounded (the closer-reward of a unit step, PARTIAL:229-234)
};

// Byte offsets into the dynamic shared memory of a tile kernel.
struct MapfTileLayout {
  int obst_off;     // padded obstacle bitmaps: [shared_map ? 1 : epb][bm_words] u32
  int agt_off;      // padded agent bitmaps:    [epb][bm_words] u32 (FOV only)
  int grida_off;    // [epb][grid_bytes] u8: PRIMAL live id grid / GRID occupancy counts of the current positions
  int gridb_off;    // [epb][grid_bytes] u8: PRIMAL pre-sweep id grid / GRID occupancy counts of the new positions
  int posold_off, posnew_off, goal_off;  // uchar2 [epb*N]
  int mv_off;       // u32 [epb*N]
  int res_off, dep_off;  // u8 [epb*N]
  int act_off, status_off, done_off, flag_off, avail_off, nextmid_off, node_off, edge_off, isint_off;  // u8 [epb*N]
  int rew_off;      // double [epb*N]
  int envrew_off;   // double [epb * max(1, N / 32)]: partial sums of the pairwise team reward
  int envterm_off;  // u8 [epb] (padded)
  int envcnt_off;   // int [epb]: per-environment counters (agents on goal / done)
  int envcnt2_off;  // int [epb]: PARTIAL: sum of node flags + edge counts
  int envstep_off;  // int [epb]: step counter before this step
  int envff_off;    // int [epb]: GRID / PARTIAL: index of the first float item of sum(rewards) (N: none)
  int pre_off;      // double [epb*N]: GRID / PARTIAL: accumulator before item i, then the compensation term of item i
  int atgoal_off;   // u8 [epb*N]: PARTIAL _agent_at_goals
  int pastold_off, pastnew_off;   // uchar2 [epb*N]: PRIMAL diagonal mode, State.agents_past before / after the sweep
  int mask16_off, nextmid16_off;  // u16 [epb*N]: 9-wide action masks (diagonal mode)
  int str_off;      // bit strings: ceil(epb*N / G) * GW u32
  int guard_off[4]; // 16-byte canaries: behind the obstacle rows, behind the occupancy grids, behind the goals, at the end
  int total_bytes;
};

// Device pointers to the state owned by the handle.
struct MapfState {
  uint32_t* obst_bits;     // [Emap][bm_words]
  uint8_t* pos;            // [E][N][2]
  uint8_t* goal;           // [E][N][2]
  uint8_t* start;          // [E][N][2]
  uint8_t* done;           // [E][N]
  uint8_t* prev_action;    // [E][N]
  int32_t* step_count;     // [E]
  int16_t* goal_dist;      // [E][N][H][W] or NULL
  const double* mag_lut;   // [mag_lut_len]
  const double* vec_lut;   // [H][W][4]: {a / mag, b / mag, mag, 0} for |dx| = a, |dy| = b (PRIMAL:380-385), host-built
  unsigned long long* stats;  // [MAPF_N_STATS]
  // PRIMAL blocking reward (cfg.blocking_reward): what the follow-up kernel needs from the sweep
  uint8_t* pos_prev;       // [E][N][2] positions before the last sweep
  uint8_t* past;           // [E][N][2] State.agents_past (diagonal mode)
  int8_t* last_status;     // [E][N]
  double* last_reward;     // [E][N]
  // MAPF_MODE_PARTIAL
  uint8_t* at_goal;        // [E][N]
  int32_t* goal_cost;      // [E][N]
  int32_t* agent_steps;    // [E][N]
  uint8_t* pnode;          // [E][N]
  uint8_t* pedge;          // [E][N]
  long long* total_coll;   // [E]
  uint8_t* terminated;     // [E]
  const double* complete_lut;
  uint32_t* err_flags;     // [1]
  int32_t* bfs_list;       // [2 + 2*E*N]: two counters, the compacted (env, agent) indices of a masked BFS, the
                           // overflow list of the register BFS kernel
};

// Per-launch arguments of the tile kernel.
struct MapfTileArgs {
  const void* actions;   // [E][N] u8 or i64 (NULL when do_step == 0)
  int act_dtype;
  int do_step;
  int agent_lo, agent_hi;  // PRIMAL sweep range
  mapf_step_out out;       // any pointer may be NULL
  void* obs;               // FOV: [E][N][4][F][F] u8/f32 ; FULLMAP: [E][H*W] i8 ; NULL: no observation
  int obs_dtype;
  double* vec;             // [E][N][3] or NULL
  int T;                   // steps in this launch (mapf_rollout): actions / outputs / obs / vec are [T][...], T >= 1
  int debug_corrupt;       // self-test of the canary check: 1 + index of a guard word the kernel overwrites on purpose
  // lifelong goal queues bound to the handle (mapf_lifelong_bind): the step's write-back pops the queue of every agent
  // that ends the step on its goal -- what mapf_pop_goals does as a launch of its own behind the step
  const int16_t* life_queue;   // [E][N][Q][2] or NULL
  int32_t* life_head;          // [E][N]
  int life_Q;
  int32_t* life_list;          // (env, agent) indices popped by this launch ...
  int32_t* life_cnt;           // ... and their number ([0]; [1] is the overflow counter of the BFS that follows)
};

#ifdef __cplusplus
extern "C" {
#endif
// launchers implemented in mapf_kernels.cu; all return a cudaError_t cast to int
int mapf_launch_tile(const MapfDims& d, const MapfTileLayout& L, const MapfState& S, const MapfTileArgs& A,
                     void* stream);
int mapf_launch_observe_generic(const MapfDims& d, const MapfState& S, uint8_t* obs_u8, float* obs_f32, double* vec,
                                void* stream);
int mapf_launch_build_obst(const MapfDims& d, const MapfState& S, const int8_t* map, const uint8_t* env_mask,
                           void* stream);
int mapf_launch_reset(const MapfDims& d, const MapfState& S, const int16_t* starts, const int16_t* goals,
                      const uint8_t* env_mask, void* stream);
int mapf_launch_set_goals(const MapfDims& d, const MapfState& S, const int16_t* goals, const uint8_t* dirty,
                          void* stream);
int mapf_launch_pop_goals(const MapfDims& d, const MapfState& S, const int16_t* queue, int32_t* head, int queue_len,
                          uint8_t* dirty, void* stream);
int mapf_launch_bfs(const MapfDims& d, const MapfState& S, const uint8_t* dirty, const uint8_t* env_mask, int16_t* dist,
                    int primal_costs, void* stream, int* n_launches,   // 8-connected when primal_costs && d.diag
                    int32_t* ext_list = nullptr, int32_t* ext_cnt = nullptr);   // a list some kernel already compacted
int mapf_launch_blocking(const MapfDims& d, const MapfState& S, int agent_lo, int agent_hi, const mapf_step_out& out,
                         void* stream, int* n_launches);
int mapf_launch_partial_obs(const MapfDims& d, const MapfState& S, void* obs, int f32, long long* state_out,
                            void* stream);
int mapf_launch_partial_state(const MapfDims& d, const MapfState& S, long long* state, void* stream);
int mapf_launch_export16(const MapfDims& d, const uint8_t* src_u8x2, int16_t* dst, void* stream);
int mapf_launch_random_actions(const MapfDims& d, const uint8_t* avail, uint32_t seed, uint32_t step,
                               long long env_offset, void* out, int i64, void* stream);
int mapf_launch_runner_mask_actions(const MapfDims& d, const void* actions, int i64, const uint8_t* alive, int stay,
                                    uint8_t* out8, long long* out64, void* stream);
int mapf_launch_runner_account(const MapfDims& d, const double* reward, const uint8_t* term, uint8_t* alive,
                               double* returns, long long* lengths, uint8_t* filled_next, void* stream);
int mapf_tile_has_fov(int F);
int mapf_tile_has_rollout(int mode);
int mapf_pipe_supported(const MapfDims& d);
int mapf_pipe_tiles(const MapfDims& d);
int mapf_launch_pipe(const MapfDims& d, const MapfState& S, const MapfTileArgs& A, void* stream);
int mapf_configure_tile(int F, int mode, int smem_bytes);
#ifdef __cplusplus
}
#endif
