"""Lifelong goal reassignment on top of the engine (BASELINE config c4).

Every agent owns a FIFO queue of goals; when it stands on its current goal after a step, the next goal is popped
and becomes its goal (the rule of the reference's lifelong system: `Global::run` pops the agent's deque on arrival,
MAPF-490-main/Global.cpp:85-94; task queues are built in main.cpp:56-85).  The queue lives on the device; popping
and the BFS of the re-assigned goals are the engine's kernels (mapf_pop_goals, mapf_bfs with a dirty mask): two
launches per step, no host round trip.
"""
import torch


class LifelongGoals:
    def __init__(self, engine, goal_queue):
        """goal_queue: int16 [E, N, Q, 2] (row, col) goals; entry 0 is the first REassignment (the initial goals
        are the ones given to engine.reset)."""
        self.engine = engine
        self.queue = torch.as_tensor(goal_queue).to(device=engine.device, dtype=torch.int16).contiguous()
        E, N, Q, _ = self.queue.shape
        assert (E, N) == (engine.E, engine.N)
        self.Q = Q
        self.head = torch.zeros((E, N), dtype=torch.int32, device=engine.device)

    def reassign(self, on_goal=None):
        """Pops the next goal of every agent that stands on its goal (the engine's PRIMAL `dones` output is exactly
        that flag; it is accepted for symmetry with the reference loop and not needed).  Returns the dirty mask that
        was applied, uint8 [E, N]."""
        d8 = self.engine.pop_goals(self.queue, self.head)
        if self.engine.has_goal_dist:
            self.engine.refresh_goal_dist(d8)
        return d8
