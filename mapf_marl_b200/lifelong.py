"""Lifelong goal reassignment on top of the engine (BASELINE config c4).

Every agent owns a FIFO queue of goals; when it stands on its current goal after a step, the next goal is popped
and becomes its goal (the rule of the reference's lifelong system: `Global::run` pops the agent's deque on arrival,
MAPF-490-main/Global.cpp:85-94; task queues are built in main.cpp:56-85).  The queue lives on the device; popping
and the BFS of the re-assigned goals are the engine's kernels (mapf_pop_goals, mapf_bfs with a dirty mask): two
launches per step, no host round trip -- or, with fused=True, no pop launch at all: the queues are bound to the handle
(mapf_lifelong_bind), the step kernel pops them in its own write-back and mapf_bfs_popped takes the list it collected.
"""
import torch


class LifelongGoals:
    def __init__(self, engine, goal_queue, dist_out=None, overlap=False, fused=False):
        """goal_queue: int16 [E, N, Q, 2] (row, col) goals; entry 0 is the first REassignment (the initial goals
        are the ones given to engine.reset).
        dist_out: optional caller-owned int16 [E,N,H,W] tensor kept up to date (otherwise the handle's own maps are
        refreshed when the engine keeps them).
        overlap: run the BFS of the re-assigned goals on a high-priority side stream, concurrently with the NEXT
        step's fused launch (a handful of single-warp BFS runs is pure latency, ~70 us at 64x64; the next pop waits
        for it, so goals never change under a running BFS).  Call sync() before reading the distance maps.
        fused: bind the queues to the engine (mapf_lifelong_bind): the step kernel pops them in its own write-back and
        collects the re-assigned agents in a list, reassign() only starts the BFS of that list -- no pop launch, no
        dirty mask, no list compaction (c4: 115 -> 102 us per step, the cost of the plain step)."""
        self.engine = engine
        self.queue = torch.as_tensor(goal_queue).to(device=engine.device, dtype=torch.int16).contiguous()
        E, N, Q, _ = self.queue.shape
        assert (E, N) == (engine.E, engine.N)
        self.Q = Q
        self.head = torch.zeros((E, N), dtype=torch.int32, device=engine.device)
        self.dist_out = dist_out
        self.fused = bool(fused)
        if self.fused:
            engine.lifelong_bind(self.queue, self.head)
        self._side = None
        if overlap:
            with torch.cuda.device(engine.device):
                self._side = torch.cuda.Stream(priority=-1)
                self._popped = torch.cuda.Event()
                self._bfs_done = torch.cuda.Event()
                self._slot_done = [torch.cuda.Event(), torch.cuda.Event()]
            self._pending = False
            self._slot_pending = [False, False]
            self._slot = 0

    def _bfs(self, d8):
        if self.fused:
            if self.dist_out is not None or self.engine.has_goal_dist:
                self.engine.bfs_popped(self.dist_out)
            return
        if self.dist_out is not None:
            self.engine.goal_dist(dirty=d8, out=self.dist_out)
        elif self.engine.has_goal_dist:
            self.engine.refresh_goal_dist(d8)

    def reassign(self, on_goal=None):
        """Pops the next goal of every agent that stands on its goal (the engine's PRIMAL `dones` output is exactly
        that flag; it is accepted for symmetry with the reference loop and not needed).  Returns the dirty mask that
        was applied, uint8 [E, N] (valid until the next call)."""
        if self.fused:
            return self._reassign_fused()
        if self._side is None:
            d8 = self.engine.pop_goals(self.queue, self.head)
            self._bfs(d8)
            return d8
        main = torch.cuda.current_stream(self.engine.device)
        if self._pending:
            main.wait_event(self._bfs_done)          # the previous BFS still reads the goals and the dirty mask
        d8 = self.engine.pop_goals(self.queue, self.head)
        self._popped.record(main)
        with torch.cuda.stream(self._side):
            self._side.wait_event(self._popped)
            self._bfs(d8)
            self._bfs_done.record(self._side)
        self._pending = True
        return d8

    def _reassign_fused(self):
        """The step that just ran popped the queues itself; start the BFS of its list.  With overlap it runs on the
        side stream under the NEXT step (which appends to the engine's other list); the step after that waits for it."""
        if self._side is None:
            self._bfs(None)
            return None
        main = torch.cuda.current_stream(self.engine.device)
        s = self._slot
        if self._slot_pending[1 - s]:                # the next step appends to the other list: its last BFS must be done
            main.wait_event(self._slot_done[1 - s])
            self._slot_pending[1 - s] = False
        self._popped.record(main)
        with torch.cuda.stream(self._side):
            self._side.wait_event(self._popped)
            self._bfs(None)
            self._slot_done[s].record(self._side)
        self._slot_pending[s] = True
        self._slot = 1 - s
        return None

    def sync(self):
        """Makes the current stream wait for the BFS of the last reassign() (overlap mode)."""
        if self._side is not None and self._pending:
            torch.cuda.current_stream(self.engine.device).wait_event(self._bfs_done)
            self._pending = False
        if self._side is not None and self.fused:
            for s in (0, 1):
                if self._slot_pending[s]:
                    torch.cuda.current_stream(self.engine.device).wait_event(self._slot_done[s])
                    self._slot_pending[s] = False
