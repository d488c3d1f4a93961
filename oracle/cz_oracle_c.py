"""ctypes front end of oracle/cz_oracle.c (the compiled CPU oracle).  TEST INFRASTRUCTURE ONLY.

`COracleEnv` has the same surface as oracle/cz_oracle.OracleEnv (step / observe / export_state / error) so the
replay helpers treat both alike; `build_c_oracle()` compiles the shared object with gcc into oracle/_build/.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "cz_oracle.c")
LIB = os.path.join(HERE, "_build", "libcz_oracle.so")
TYPES = ["Floor", "Counter", "Deliversquare", "Switch", "Block", "Cutboard", "Blender", "Plate", "Onion", "Tomato",
         "Lettuce", "Carrot", "Cucumber", "Banana", "Apple", "Watermelon", "Bread", "Agent"]
CODE = {n: i for i, n in enumerate(TYPES)}
MAX_NODES = 8
_lib = None


def build_c_oracle(force=False):
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(SRC):
        return LIB
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    subprocess.run(["gcc", "-O2", "-fPIC", "-shared", "-std=c11", "-ffp-contract=off", "-o", LIB, SRC], check=True)
    return LIB


def load():
    global _lib
    if _lib is None:
        lib = C.CDLL(build_c_oracle())
        P = C.c_void_p
        lib.czo_create.restype = P
        lib.czo_create.argtypes = [P, P, P, C.c_int, P, C.c_int, P, P, P]
        lib.czo_destroy.argtypes = [P]
        lib.czo_reset.argtypes = [P]
        lib.czo_set_stream.argtypes = [P, C.c_uint64, C.c_uint64, C.c_uint64]
        lib.czo_step.argtypes = [P, P, P, P, P, P]
        lib.czo_obs_len.argtypes = [P]
        lib.czo_obs_len.restype = C.c_int
        lib.czo_observe.argtypes = [P, C.c_int, P]
        lib.czo_export.argtypes = [P, P, P, P, P, P]
        lib.czo_error.argtypes = [P]
        lib.czo_error.restype = C.c_uint32
        lib.czo_teleport.argtypes = [P, C.c_int, C.c_int, C.c_int]
        lib.czo_batch_step.argtypes = [P, C.c_int, P, P, P, P, P]
        _lib = lib
    return _lib


_RECIPE_CACHE = {}


def _recipe_rows(names):
    """node_list of each recipe as (type code, condition, children mask) — from the Python oracle's book."""
    key = tuple(names)
    if key not in _RECIPE_CACHE:
        _RECIPE_CACHE[key] = _recipe_rows_uncached(names)
    return _RECIPE_CACHE[key]


def _recipe_rows_uncached(names):
    from .cz_oracle import make_recipe
    out = np.zeros((len(names), 1 + 3 * MAX_NODES), np.int32)
    for r, name in enumerate(names):
        nodes = make_recipe(name)
        pos = {id(n): k for k, n in enumerate(nodes)}
        out[r, 0] = len(nodes)
        for k, n in enumerate(nodes):
            kids = 0
            for c in n.kids:
                kids |= 1 << pos[id(c)]
            out[r, 1 + 3 * k:4 + 3 * k] = (CODE[n.type], {None: 0, "chopped": 1, "mashed": 2}[n.cond], kids)
    return out


class COracleEnv:
    DEFAULT_REWARD = {"recipe_reward": 20, "max_time_penalty": -5, "recipe_penalty": -40, "recipe_node_reward": 0}

    def __init__(self, layout, recipes, max_steps, reward_scheme=None, end_condition_all_dishes=False,
                 agent_respawn_rate=0.0, grace_period=20, agent_despawn_rate=0.0, spawn_stream=None,
                 action_scheme="scheme3"):
        lib = load()
        self.lib = lib
        rs = reward_scheme or self.DEFAULT_REWARD
        A = len(layout["agents"])
        self.A, self.R = A, len(recipes)
        cfg = np.array([layout["width"], layout["height"], A, len(recipes), max_steps, int(end_condition_all_dishes),
                        1 if action_scheme == "scheme1" else 3, grace_period], np.int32)
        rw = np.array([rs["recipe_node_reward"], rs["recipe_reward"], rs["recipe_penalty"],
                       rs["max_time_penalty"] / max_steps, agent_respawn_rate, agent_despawn_rate], np.float64)
        meta = np.array([[0 if CODE[k] <= CODE["Blender"] else (2 if k == "Agent" else 1), CODE[k], v]
                         for k, v in layout["meta"]], np.int32)
        world = np.array([[CODE[t], x, y] for t, locs in layout["objects"] for x, y in locs], np.int32)
        agents = np.array(layout["agents"], np.int32)
        spawn = np.zeros((A, 18), np.int32)
        for i, (xs, ys) in enumerate(layout.get("agent_spawn", [])[:A]):
            spawn[i, 0], spawn[i, 1:1 + len(xs)] = len(xs), xs
            spawn[i, 9], spawn[i, 10:10 + len(ys)] = len(ys), ys
        rec = _recipe_rows(recipes)
        self._keep = (cfg, rw, meta, world, agents, spawn, rec)
        self.h = lib.czo_create(cfg.ctypes.data, rw.ctypes.data, meta.ctypes.data, len(meta), world.ctypes.data,
                                len(world), agents.ctypes.data, spawn.ctypes.data, rec.ctypes.data)
        if spawn_stream is not None:
            lib.czo_set_stream(self.h, spawn_stream.seed, spawn_stream.env, spawn_stream.episode)
        self.L = lib.czo_obs_len(self.h)
        self.nd = int(sum(v for k, v in layout["meta"] if CODE["Plate"] <= CODE[k] <= CODE["Bread"]))
        self.ns = int(sum(v for k, v in layout["meta"] if CODE[k] <= CODE["Blender"]))

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.czo_destroy(self.h)
            self.h = None

    @property
    def error(self):
        return int(self.lib.czo_error(self.h))

    def step(self, actions):
        a = np.ascontiguousarray(np.asarray(actions, dtype=np.uint8))
        rew = np.zeros(self.A, np.float64)
        te, tu, rel = (np.zeros(self.A, np.uint8) for _ in range(3))
        self.lib.czo_step(self.h, a.ctypes.data, rew.ctypes.data, te.ctypes.data, tu.ctypes.data, rel.ctypes.data)
        return rew, te, tu, rel

    def observe(self, i):
        out = np.zeros(self.L, np.float64)
        self.lib.czo_observe(self.h, i, out.ctypes.data)
        return out

    def export_state(self):
        agents = np.zeros((self.A, 6), np.int16)
        objs = np.zeros((self.nd, 9), np.int16)
        statics = np.zeros((self.ns, 4), np.int16)
        marks = np.zeros(self.R, np.int32)
        t = np.zeros(1, np.int32)
        self.lib.czo_export(self.h, agents.ctypes.data, objs.ctypes.data, statics.ctypes.data, marks.ctypes.data,
                            t.ctypes.data)
        return {"agents": agents, "objs": objs, "statics": statics, "marks": marks, "t": np.int32(t[0])}

    def teleport(self, i, x, y):
        self.lib.czo_teleport(self.h, int(i), int(x), int(y))


class CBatch:
    """n independent compiled-oracle environments stepped together (slices run on a thread pool: ctypes drops
    the GIL inside czo_batch_step).  The unit of work equals the GPU path's: step + every agent's observation."""

    def __init__(self, layouts, recipes_per_env, max_steps, threads=None, **kw):
        from concurrent.futures import ThreadPoolExecutor
        self.envs = [COracleEnv(lay, rec, max_steps, **kw) for lay, rec in zip(layouts, recipes_per_env)]
        self.n, self.A, self.L = len(self.envs), self.envs[0].A, self.envs[0].L
        self.handles = (C.c_void_p * self.n)(*[e.h for e in self.envs])
        self.threads = threads or min(32, os.cpu_count() or 1)
        self.pool = ThreadPoolExecutor(self.threads)
        self.obs = np.zeros((self.n, self.A, self.L), np.float64)
        self.reward = np.zeros((self.n, self.A), np.float64)
        self.term = np.zeros((self.n, self.A), np.uint8)
        self.trunc = np.zeros((self.n, self.A), np.uint8)
        self.lib = load()

    def observe(self):
        for i, e in enumerate(self.envs):
            for a in range(self.A):
                self.obs[i, a] = e.observe(a)
        return self.obs

    def step(self, actions):
        act = np.ascontiguousarray(actions, dtype=np.uint8)
        bounds = np.linspace(0, self.n, self.threads + 1).astype(int)
        hsz = C.sizeof(C.c_void_p)

        def run(k):
            lo, hi = int(bounds[k]), int(bounds[k + 1])
            if hi > lo:
                self.lib.czo_batch_step(C.addressof(self.handles) + lo * hsz, hi - lo, act[lo:].ctypes.data,
                                        self.obs[lo:].ctypes.data, self.reward[lo:].ctypes.data,
                                        self.term[lo:].ctypes.data, self.trunc[lo:].ctypes.data)
        list(self.pool.map(run, range(self.threads)))
        return self.obs, self.reward, self.term, self.trunc
