"""Lifelong goal reassignment on top of the engine (BASELINE config c4).

Every agent owns a FIFO queue of goals; when it stands on its current goal after a step, the next goal is popped
and becomes its goal (the rule of the reference's lifelong system: `Global::run` pops the agent's deque on arrival,
MAPF-490-main/Global.cpp:85-94; task queues are built in main.cpp:56-85).  The queue lives on the device; popping
is a gather on device tensors, the goal write and the BFS of the reassigned goals are the engine's kernels
(mapf_set_goals, mapf_bfs with a dirty mask).
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
        self.head = torch.zeros((E, N), dtype=torch.int64, device=engine.device)

    def reassign(self, on_goal):
        """on_goal: uint8/bool [E, N] (the engine's PRIMAL `dones` output).  Returns the dirty mask that was applied."""
        dirty = on_goal.bool() & (self.head < self.Q)
        idx = self.head.clamp(max=self.Q - 1)[..., None, None].expand(-1, -1, 1, 2)
        nxt = torch.gather(self.queue, 2, idx)[:, :, 0, :]
        self.head += dirty.long()
        d8 = dirty.to(torch.uint8)
        self.engine.set_goals(nxt, d8)
        if self.engine.has_goal_dist:
            self.engine.refresh_goal_dist(d8)
        return d8
