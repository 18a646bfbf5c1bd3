"""Map / scenario ingestion and synthetic map generators (host side, reset-time only).

MovingAI formats as the reference reads them:
  .map  : 4 header lines, then one text row per map row; '.' is free, anything else is an
          obstacle (mapf_gridworld.py:421-428, :287).
  .scen : 1 header line, then tab-separated records whose fields 4..7 are start x, start y,
          goal x, goal y (mapf_gridworld.py:442-443).
"""
import numpy as np


def read_movingai_map(path):
    """-> uint8 [H, W], 1 = obstacle."""
    with open(path, "r") as f:
        rows = [row.rstrip() for row in f.readlines()][4:]
    rows = [r for r in rows if len(r) > 0]
    assert len(rows) > 0 and len(rows[0]) > 0
    return np.array([[0 if ch == "." else 1 for ch in row] for row in rows], dtype=np.uint8)


def scen_fields(line):
    parts = line.replace("\t", ",").split(",")
    return int(parts[4]), int(parts[5]), int(parts[6]), int(parts[7])


def read_scen_lines(path):
    """The record lines of a .scen file (header dropped), as the reference slices them."""
    with open(path, "r") as f:
        return [row.rstrip() for row in f.readlines()][1:]


def random_obstacles(rs, height, width, density):
    """`world = -(rand(H, W) < p)` as PRIMAL draws it (mapf_primal.py:315); returns uint8, 1 = obstacle."""
    return (rs.rand(height, width) < density).astype(np.uint8)


def label_components(obst):
    """4-connected component label of every free cell (-1 on obstacles)."""
    try:
        from scipy import ndimage
        lab, _ = ndimage.label(np.asarray(obst) == 0)      # default structure: 4-connectivity
        return lab.astype(np.int32) - 1
    except ImportError:
        pass
    H, W = obst.shape
    lab = np.full((H, W), -1, np.int32)
    cur = 0
    for i in range(H):
        for j in range(W):
            if obst[i, j] or lab[i, j] >= 0:
                continue
            stack = [(i, j)]
            lab[i, j] = cur
            while stack:
                a, b = stack.pop()
                for da, db in ((1, 0), (-1, 0), (0, 1), (0, -1)):
                    x, y = a + da, b + db
                    if 0 <= x < H and 0 <= y < W and not obst[x, y] and lab[x, y] < 0:
                        lab[x, y] = cur
                        stack.append((x, y))
            cur += 1
    return lab


def place_agents_and_goals(rs, obst, n_agents):
    """Distinct start cells and distinct goal cells, every goal inside its agent's connected region
    (the constraint PRIMAL's _setWorld enforces, mapf_primal.py:317-337)."""
    lab = label_components(obst)
    W = obst.shape[1]
    flat = np.flatnonzero(lab.ravel() >= 0)                  # free cells in row-major order
    if len(flat) < 2 * n_agents:
        raise ValueError("map too dense for %d agents" % n_agents)
    order = rs.permutation(len(flat))
    sflat = flat[order[:n_agents]]
    labf = lab.ravel()
    region = {}                                              # label -> its not-yet-taken cells, row-major
    goals = np.zeros((n_agents, 2), np.int64)
    for k in range(n_agents):
        lk = int(labf[sflat[k]])
        cells = region.get(lk)
        if cells is None:
            cells = region[lk] = np.flatnonzero(labf == lk).tolist()
        if len(cells) == 0:
            raise ValueError("no free goal cell left in agent %d's region" % k)
        g = cells.pop(int(rs.randint(len(cells))))
        goals[k] = (g // W, g % W)
    starts = np.stack([sflat // W, sflat % W], axis=1)
    return starts.astype(np.int16), goals.astype(np.int16)


def synthetic_batch(seed, n_envs, height, width, density, n_agents, shared_map=False, env_offset=0,
                    distinct=64):
    """Seeded synthetic worlds for benchmarks and tests.

    Environment e (global index env_offset + e) is world number (env_offset + e) % distinct drawn from
    RandomState(seed + world number): content depends only on the GLOBAL environment index, never on how
    environments are sharded over GPUs.  `distinct` bounds the host-side generation cost for very large
    batches (a million Python-generated worlds would take minutes); the device work is identical."""
    E = n_envs
    n_worlds = min(distinct, E) if distinct else E
    worlds = {}

    def world(k):
        if k not in worlds:
            rs = np.random.RandomState(seed + k)
            if shared_map:
                m = shared
            else:
                m = random_obstacles(rs, height, width, density)
            s, g = place_agents_and_goals(rs, m, n_agents)
            worlds[k] = (m, s, g)
        return worlds[k]

    shared = random_obstacles(np.random.RandomState(seed), height, width, density) if shared_map else None
    obst = np.zeros((1 if shared_map else E, height, width), np.uint8)
    starts = np.zeros((E, n_agents, 2), np.int16)
    goals = np.zeros((E, n_agents, 2), np.int16)
    for e in range(E):
        k = (env_offset + e) % n_worlds if distinct else env_offset + e
        m, s, g = world(k)
        if not shared_map:
            obst[e] = m
        starts[e], goals[e] = s, g
    if shared_map:
        obst[0] = shared
    return (obst[0] if shared_map else obst), starts, goals


def warehouse_layout(height=64, width=64, lane_every=8):
    """A highway/warehouse-style layout: shelf blocks separated by two-cell aisles every `lane_every`
    cells, in the spirit of highway_layout_v19.py's evenly spaced lanes (ideal_strips, :79-91).  Only the
    walls matter to the step/observation path.  Returns uint8 [H, W], 1 = obstacle."""
    obst = np.ones((height, width), np.uint8)
    for r in range(height):
        for c in range(width):
            if r % lane_every in (0, 1) or c % lane_every in (0, 1):
                obst[r, c] = 0
    obst[0, :] = obst[-1, :] = 0
    obst[:, 0] = obst[:, -1] = 0
    return obst


# --------------------------------------------------------------------------- highway_layout_v19.py outputs
HIGHWAY_ALPHABET = "@.`LnsewIX"


def parse_highway_rows(rows):
    """Rows written by highway_layout_v19.py (get_output_hw_label :1158-1180 / get_output_global_label :1183-1205):
    '@' obstacle (in highways.txt town cells are '@' too), '.' or '`' free town cell, 'L' lock, 'n' 's' 'e' 'w'
    one-way highway cell, 'I' intersection, 'X' highway end.  Rows are top row first (the script writes y from
    nrows-1 down to 0, :1454-1465).  Returns dict(obst u8[H,W], lane i8[H,W] (0 none, 1 n, 2 s, 3 e, 4 w),
    lock u8[H,W], junction u8[H,W]); only `obst` matters to the step/observation path."""
    import numpy as np
    H, W = len(rows), len(rows[0])
    obst = np.zeros((H, W), np.uint8)
    lane = np.zeros((H, W), np.int8)
    lock = np.zeros((H, W), np.uint8)
    junction = np.zeros((H, W), np.uint8)
    code = {"n": 1, "s": 2, "e": 3, "w": 4}
    for r, row in enumerate(rows):
        assert len(row) == W, "ragged highway map"
        for c, ch in enumerate(row):
            if ch not in HIGHWAY_ALPHABET:
                raise ValueError("unknown highway map symbol %r at (%d, %d)" % (ch, r, c))
            obst[r, c] = ch == "@"
            lane[r, c] = code.get(ch, 0)
            lock[r, c] = ch == "L"
            junction[r, c] = ch in "IX"
    return dict(obst=obst, lane=lane, lock=lock, junction=junction)


def read_highway_map(path):
    """highways.txt as written by highway_layout_v19.py to_file (:1536-1546): three header lines
    (`height (n_rows): H`, `width (n_cols): W`, `Highway Map:`) then H rows."""
    with open(path, "r") as f:
        lines = [ln.rstrip("\n") for ln in f.readlines()]
    h = int(lines[0].split(":")[1])
    w = int(lines[1].split(":")[1])
    rows = [ln for ln in lines[3:3 + h]]
    assert len(rows) == h and all(len(r) == w for r in rows)
    return parse_highway_rows(rows)
