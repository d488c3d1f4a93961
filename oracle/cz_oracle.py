"""CPU oracle: a plain-Python restatement of CookingZoo's per-step hot path.

TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this file; the product package
(cooking_zoo_b200/) never does and fails loudly when its CUDA library is missing.

Parity status: the reference ships NO tests or golden vectors, so this oracle
is pinned against outputs of the reference itself, run in the build container behind
oracle/refshim (tests/golden/make_golden.py -> tests/golden/*.npz, plus the live
cross-check tests/test_oracle_vs_reference.py that runs whenever /root/reference exists).

All `file:line` citations are relative to /root/reference/cooking_zoo/.  The restatement is
organised around plain records and lists, not the reference's class hierarchy:

    step()                      environment/cooking_env.py:243-269 (accumulated_step)
      _world_step()             cooking_world/cooking_world.py:104-112
        _agent_actions()        cooking_world/cooking_action_util/action_scheme3.py:4-43
        _progress_world()       cooking_world/cooking_world.py:77-88
        _linked()               cooking_world/cooking_world.py:90-92
        _agent_spawn()          cooking_world/cooking_world.py:267-290
      _rewards()                environment/cooking_env.py:290-350
    observe()                   environment/cooking_env.py:352-373
    BOOK / _update_recipe()     cooking_book/recipe_drawer.py:40-118, cooking_book/recipe.py:29-104
"""
import numpy as np

# ---------------------------------------------------------------------------------------
# Entity traits (cooking_world/world_objects.py, cooking_world/abstract_classes.py)
# ---------------------------------------------------------------------------------------
# kind: "static" | "dynamic" | "agent"; fv = feature_vector_length() of the class.
TRAITS = {
    # statics: (walkable at construction)                                world_objects.py
    "Floor":         dict(kind="static", walkable=True, fv=0),            # :17-54
    "Counter":       dict(kind="static", walkable=False, fv=3),           # :57-98
    "Deliversquare": dict(kind="static", walkable=False, fv=3),           # :101-141
    "Switch":        dict(kind="static", walkable=True, fv=4),            # :144-192
    "Block":         dict(kind="static", walkable=False, fv=4),           # :195-239
    "Cutboard":      dict(kind="static", walkable=False, fv=3),           # :242-311
    "Blender":       dict(kind="static", walkable=False, fv=3),           # :314-383
    # dynamics
    "Plate":         dict(kind="dynamic", fv=3, plate=True),              # :386-432
    "Onion":         dict(kind="dynamic", fv=5, chop=True),               # :435-468
    "Tomato":        dict(kind="dynamic", fv=5, chop=True),               # :471-504
    "Lettuce":       dict(kind="dynamic", fv=5, chop=True),               # :507-540
    "Carrot":        dict(kind="dynamic", fv=6, chop=True, blend=True),   # :543-580
    "Cucumber":      dict(kind="dynamic", fv=5, chop=True),               # :583-613
    "Banana":        dict(kind="dynamic", fv=6, chop=True, blend=True),   # :616-653
    "Apple":         dict(kind="dynamic", fv=5, chop=True),               # :656-689
    "Watermelon":    dict(kind="dynamic", fv=5, chop=True),               # :692-725
    "Bread":         dict(kind="dynamic", fv=5, chop=True, spawn=True),   # :728-771
    "Agent":         dict(kind="agent", fv=7),                            # :774-826
}

# ---------------------------------------------------------------------------------------
# Recipe book (cooking_book/recipe_drawer.py:40-118).  Node ids are allocated in
# definition order by get_next_default_id (:28-31): 10 leaves, 7 plates, 7 deliveries,
# floor, no_recipe  ->  DEFAULT_NUM_GOALS = 26.
# A node is (id, type name, condition or None, [child node keys]).
# ---------------------------------------------------------------------------------------
_LEAVES = [  # recipe_drawer.py:40-59
    ("ChoppedLettuce", "Lettuce", "chopped"), ("ChoppedOnion", "Onion", "chopped"),
    ("ChoppedTomato", "Tomato", "chopped"), ("ChoppedApple", "Apple", "chopped"),
    ("ChoppedCucumber", "Cucumber", "chopped"), ("ChoppedWatermelon", "Watermelon", "chopped"),
    ("ChoppedBanana", "Banana", "chopped"), ("MashedBanana", "Banana", "mashed"),
    ("ChoppedCarrot", "Carrot", "chopped"), ("MashedCarrot", "Carrot", "mashed"),
]
_PLATES = [  # recipe_drawer.py:62-79
    ("TomatoSaladPlate", ["ChoppedTomato"]),
    ("TomatoLettucePlate", ["ChoppedTomato", "ChoppedLettuce"]),
    ("TomatoLettuceOnionPlate", ["ChoppedTomato", "ChoppedLettuce", "ChoppedOnion"]),
    ("CarrotBananaPlate", ["ChoppedCarrot", "ChoppedBanana"]),
    ("MashedCarrotBananaPlate", ["MashedCarrot", "MashedBanana"]),
    ("CucumberOnionPlate", ["ChoppedCucumber", "ChoppedOnion"]),
    ("AppleWatermelonPlate", ["ChoppedApple", "ChoppedWatermelon"]),
]
_DELIVERIES = [  # recipe_drawer.py:84-100
    ("TomatoSalad", "TomatoSaladPlate"), ("TomatoLettuceSalad", "TomatoLettucePlate"),
    ("TomatoLettuceOnionSalad", "TomatoLettuceOnionPlate"), ("CarrotBanana", "CarrotBananaPlate"),
    ("MashedCarrotBanana", "MashedCarrotBananaPlate"), ("CucumberOnion", "CucumberOnionPlate"),
    ("AppleWatermelon", "AppleWatermelonPlate"),
]


def _build_book():
    nodes = {}
    nid = 0
    for key, typ, cond in _LEAVES:
        nodes[key] = (nid, typ, cond, [])
        nid += 1
    for key, kids in _PLATES:
        nodes[key] = (nid, "Plate", None, list(kids))
        nid += 1
    for key, kid in _DELIVERIES:
        nodes[key] = (nid, "Deliversquare", None, [kid])
        nid += 1
    nodes["floor"] = (nid, "Floor", None, [])            # recipe_drawer.py:104
    nid += 1
    nodes["no_recipe"] = (nid, "Deliversquare", None, ["floor"])   # :106-107
    nid += 1
    return nodes, nid


BOOK_NODES, NUM_GOALS = _build_book()
# RECIPES dict order (recipe_drawer.py:109-118); value = root node key
RECIPES = {
    "TomatoSalad": "TomatoSalad", "TomatoLettuceSalad": "TomatoLettuceSalad",
    "CarrotBanana": "CarrotBanana", "MashedCarrotBanana": "MashedCarrotBanana",
    "CucumberOnion": "CucumberOnion", "AppleWatermelon": "AppleWatermelon",
    "TomatoLettuceOnionSalad": "TomatoLettuceOnionSalad", "no_recipe": "no_recipe",
}


CUSTOM_RECIPES = {}      # name -> root node key; kept apart from the book (RECIPES / NUM_GOALS stay the reference's defaults)


def register_recipe(name, tree):
    """recipe_drawer.register_recipe (:34-35) for the oracle: `tree` = (type name, condition or None, [child trees]).
    Node ids continue after the book's, as get_next_id would allot them."""
    if name in CUSTOM_RECIPES:
        return

    def add(t, path):
        typ, cond, kids = t
        key = f"{name}/{path}"
        kid_keys = [add(k, f"{path}.{j}") for j, k in enumerate(kids)]
        BOOK_NODES[key] = (NUM_GOALS + sum(1 for k in BOOK_NODES if "/" in k), typ, cond, kid_keys)
        return key

    CUSTOM_RECIPES[name] = add(tree, "0")


class _Node:
    __slots__ = ("id", "type", "cond", "kids", "marked", "hits")

    def __init__(self, key):
        nid, typ, cond, kids = BOOK_NODES[key]
        self.id, self.type, self.cond = nid, typ, cond
        self.kids = [_Node(k) for k in kids]
        self.marked = False
        self.hits = []


def _expand(node):
    """Recipe.expand_child_nodes (recipe.py:89-93): children, then each child's expansion."""
    out = list(node.kids)
    for k in node.kids:
        out.extend(_expand(k))
    return out


def make_recipe(name):
    root = _Node(RECIPES[name] if name in RECIPES else CUSTOM_RECIPES[name])
    return [root] + _expand(root)          # node_list, recipe.py:31-33


# ---------------------------------------------------------------------------------------
# Records
# ---------------------------------------------------------------------------------------
class _Obj:
    """One world object (static or dynamic)."""
    __slots__ = ("type", "x", "y", "walkable", "content", "ready", "toggle", "switch_active",
                 "pressed", "chopped", "blend", "progress", "free", "tr")

    def __init__(self, typ, x, y):
        tr = TRAITS[typ]
        self.type, self.x, self.y, self.tr = typ, x, y, tr
        self.walkable = tr.get("walkable", False)
        self.content = []
        self.ready = False          # ActionObject.status == READY (abstract_classes.py:99)
        self.toggle = False         # ToggleObject (abstract_classes.py:108)
        self.switch_active = False  # world_objects.py:150
        self.pressed = False        # world_objects.py:151
        self.chopped = False        # ChopFood.chop_state (abstract_classes.py:246-248)
        self.blend = 0              # 0 FRESH, 1 IN_PROGRESS, 2 MASHED (abstract_classes.py:259-264)
        self.progress = 1           # BlenderFood.current_progress
        self.free = True            # DynamicObject.free (abstract_classes.py:228)

    def done(self):
        """Food.done(): chopped, or mashed for Carrot/Banana (world_objects.py:441,549,622...)."""
        return self.chopped or (self.tr.get("blend", False) and self.blend == 2)

    def is_food(self):
        return self.tr.get("chop", False) or self.tr.get("blend", False)


class _Agent:
    __slots__ = ("x", "y", "orientation", "holding")

    def __init__(self, x, y):
        self.x, self.y = x, y
        self.orientation = 1        # world_objects.py:782
        self.holding = None


def _refresh_free(content):
    """every item not free, the last one free (world_objects.py:71-75 and siblings)."""
    for c in content:
        c.free = False
    content[-1].free = True


_DELTA = {1: (-1, 0), 2: (1, 0), 3: (0, 1), 4: (0, -1)}     # cooking_world.py:172-184


_M64 = (1 << 64) - 1


def spawn_uniform(seed, env, episode, t, c):
    """The shared counter-based stream that stands in for the reference's global RNG draws in
    handle_agent_spawn (include/cz_b200.h: cz_spawn_uniform): splitmix64 finaliser -> [0, 1)."""
    z = (seed + 0x9E3779B97F4A7C15 * (env + 1) + 0xD1B54A32D192ED03 * (episode + 1)
         + 0x8CB92BA72F3D8DD7 * (t + 1) + 0xF1357AEA2E62A9C5 * (c + 1)) & _M64
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & _M64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & _M64
    z ^= z >> 31
    return (z >> 11) * (1.0 / 9007199254740992.0)


class SpawnStream:
    """Draw c = 0, 1, 2 ... of (seed, env, episode, t); `begin_step(t)` rewinds c."""

    def __init__(self, seed, env, episode=1):
        self.seed, self.env, self.episode = seed, env, episode
        self.t = self.c = 0

    def begin_step(self, t):
        self.t, self.c = t, 0

    def uniform(self):
        u = spawn_uniform(self.seed, self.env, self.episode, self.t, self.c)
        self.c += 1
        return u

    def choice(self, seq):
        return seq[min(len(seq) - 1, int(self.uniform() * len(seq)))]


class OracleEnv:
    """One environment.  `layout` is the plain-data dict produced by
    oracle/ref_dump.describe_layout or cooking_zoo_b200.layout.Layout.to_dict()."""

    DEFAULT_REWARD = {"recipe_reward": 20, "max_time_penalty": -5, "recipe_penalty": -40,
                      "recipe_node_reward": 0}   # cooking_env.py:79-80

    def __init__(self, layout, recipes, max_steps, reward_scheme=None,
                 end_condition_all_dishes=False, agent_respawn_rate=0.0, grace_period=20,
                 agent_despawn_rate=0.0, spawn_stream=None, action_scheme="scheme3"):
        self.layout = layout
        self.recipe_names = list(recipes)
        self.max_steps = max_steps
        self.reward_scheme = reward_scheme or dict(self.DEFAULT_REWARD)
        self.end_all = end_condition_all_dishes
        self.respawn_rate = agent_respawn_rate
        self.despawn_rate = agent_despawn_rate
        self.grace_period = grace_period
        self.spawn_stream = spawn_stream   # SpawnStream: stands in for np.random.random / random.sample
        if action_scheme not in ("scheme1", "scheme3"):
            raise ValueError("scheme2 raises AttributeError in the reference itself (action_scheme2.py:15)")
        self.scheme = action_scheme
        self.error = 0                  # bit flags: "the reference would have raised here"
        self.events = {}                # branch-coverage counters for the test-suite
        self.reset()

    # ----------------------------------------------------------------- reset
    def reset(self):
        """cooking_env.py:178-210 with the layout already sampled (parsing.py:5-151)."""
        L = self.layout
        self.W, self.H = L["width"], L["height"]
        self.meta = [(k, v) for k, v in L["meta"]]
        self.by_type = {}                       # insertion-ordered like world.world_objects
        self.static_at = {}
        for typ, locs in L["objects"]:
            lst = self.by_type.setdefault(typ, [])
            for x, y in locs:
                o = _Obj(typ, x, y)
                lst.append(o)
                if o.tr["kind"] == "static":
                    self.static_at[(x, y)] = o
        # dynamic objects start as the content of the Counter under them (parsing.py:107-108)
        for typ, lst in self.by_type.items():
            if TRAITS[typ]["kind"] == "dynamic":
                for o in lst:
                    holder = self.static_at[(o.x, o.y)]
                    holder.content.append(o)
                    _refresh_free(holder.content)
        self.agents = [_Agent(x, y) for x, y in L["agents"]]
        self.spawn_ranges = L.get("agent_spawn", [])
        A = len(self.agents)
        self.active = [True] * A                 # load_level.py:67
        self.status_changed = [False] * A
        self.grace = [self.grace_period] * A     # parsing.py:141
        self.t = 0
        self.recipes = [make_recipe(n) for n in self.recipe_names]
        for rec in self.recipes:                 # cooking_env.py:197-198
            self._update_recipe(rec)
        self.rewards = [0.0] * A
        self.terminated = [False] * A
        self.truncated = [False] * A
        self.relevant = [True] * A
        self.n_live = A        # len(env.agents) in the reference: relevant agents of the last step

    def _ev(self, name):
        self.events[name] = self.events.get(name, 0) + 1

    # ----------------------------------------------------------------- queries
    def _scan(self, x, y):
        """Dynamic objects at a cell in the reference's scan order: type names in
        world_objects insertion order, list order within a type (cooking_world.py:232-241)."""
        out = []
        for typ, lst in self.by_type.items():
            if TRAITS[typ]["kind"] != "dynamic":
                continue
            for o in lst:
                if o.x == x and o.y == y:
                    out.append(o)
        return out

    def _target(self, ag, action):
        d = _DELTA.get(action)
        return (ag.x + d[0], ag.y + d[1]) if d else (ag.x, ag.y)

    def _walkable(self, cell):
        return self.static_at[cell].walkable     # cooking_world.py:223-227

    def _move_obj(self, o, x, y):
        """move_to; a Plate drags its content along (world_objects.py:393-396)."""
        o.x, o.y = x, y
        for c in o.content:
            c.x, c.y = x, y

    # ----------------------------------------------------------------- step
    def step(self, actions):
        """actions: one entry per agent slot (entries of inactive agents are ignored)."""
        self.t += 1                                               # cooking_env.py:244
        if self.spawn_stream is not None:
            self.spawn_stream.begin_step(self.t)
        active_start = list(self.active)
        self._world_step([int(a) for a in actions])
        self._rewards(active_start)
        return self.rewards, self.terminated, self.truncated, self.relevant

    def _world_step(self, actions):
        idx = [i for i in range(len(self.agents)) if self.active[i]]   # cooking_world.py:295
        self.status_changed = [False] * len(self.agents)
        if self.scheme == "scheme1":
            self._agent_actions_scheme1(idx, [actions[i] for i in idx])
        else:
            self._agent_actions(idx, [actions[i] for i in idx])
        self._progress_world()
        self._linked()
        self._agent_spawn()
        self.relevant = [self.active[i] or self.status_changed[i] for i in range(len(self.agents))]

    def _agent_actions(self, idx, acts):
        """action_scheme3.perform_agent_actions (action_scheme3.py:4-16)."""
        ags = [self.agents[i] for i in idx]
        faced = []
        for ag, a in zip(ags, acts):
            if a in _DELTA:                                      # WALK_ACTIONS, :8-10
                faced.append(self._target(ag, a))
                ag.orientation = a
            else:
                faced.append((ag.x, ag.y))
        # check_inbounds (cooking_world.py:192-204)
        acts = list(acts)
        for k, (ag, a) in enumerate(zip(ags, acts)):
            if a == 0 or a == 5:
                continue
            tx, ty = self._target(ag, a)
            if tx > self.W - 1 or tx < 0 or ty > self.H - 1 or ty < 0:
                acts[k] = 0
        # check_collisions (cooking_world.py:206-221): one pass on pre-move positions
        ends, walk = [], []
        for ag, a in zip(ags, acts):
            tgt = self._target(ag, a)
            w = self._walkable(tgt)
            ends.append(tgt if w else (ag.x, ag.y))
            walk.append(w)
        final = []
        for k, a in enumerate(acts):
            others = ends[:k] + ends[k + 1:]
            final.append(0 if (ends[k] in others and walk[k]) else a)
            if final[-1] != a:
                self._ev("collision_cancel")
        # sequential resolution in agent order (action_scheme3.py:15-23)
        for ag, a, cell in zip(ags, final, faced):
            tgt = self._target(ag, a)
            moved = False
            if self._walkable(tgt):                              # resolve_walking_action :26-34
                ag.x, ag.y = tgt
                if ag.holding is not None:
                    self._move_obj(ag.holding, tgt[0], tgt[1])   # world_objects.py:793-796
                st = self.static_at[tgt]
                if st.type == "Switch":                          # Switch.add_content :159-163
                    st.switch_active = not st.switch_active
                    st.pressed = True
                moved = True
            if not moved and a != 0:                             # `orig_location is agent.location` :22
                self._interact(ag, cell)

    def _checked_actions(self, ags, acts):
        """check_inbounds + check_collisions (cooking_world.py:192-221), shared by the schemes."""
        acts = list(acts)
        for k, (ag, a) in enumerate(zip(ags, acts)):
            if a == 0 or a == 5:
                continue
            tx, ty = self._target(ag, a)
            if tx > self.W - 1 or tx < 0 or ty > self.H - 1 or ty < 0:
                acts[k] = 0
        ends, walk = [], []
        for ag, a in zip(ags, acts):
            tgt = self._target(ag, a)
            w = self._walkable(tgt)
            ends.append(tgt if w else (ag.x, ag.y))
            walk.append(w)
        final = []
        for k, a in enumerate(acts):
            others = ends[:k] + ends[k + 1:]
            final.append(0 if (ends[k] in others and walk[k]) else a)
            if final[-1] != a:
                self._ev("collision_cancel")
        return final

    def _agent_actions_scheme1(self, idx, acts):
        """action_scheme1.perform_agent_actions (action_scheme1.py:4-40): eight actions; walking never
        interacts, 5 = primary interaction, 6 = pick up from a plate, 7 = execute, all on the faced cell."""
        ags = [self.agents[i] for i in idx]
        for ag, a in zip(ags, acts):
            if a in _DELTA:
                ag.orientation = a                               # :7-8
        final = self._checked_actions(ags, acts)
        for ag, a in zip(ags, final):
            if a in _DELTA:                                      # resolve_walking_action :22-30
                tgt = self._target(ag, a)
                if self._walkable(tgt):
                    ag.x, ag.y = tgt
                    if ag.holding is not None:
                        self._move_obj(ag.holding, tgt[0], tgt[1])
                    st = self.static_at[tgt]
                    if st.type == "Switch":
                        st.switch_active = not st.switch_active
                        st.pressed = True
            elif a in (5, 6, 7):                                 # resolve_interaction :33-40
                cell = self._target(ag, ag.orientation)
                if cell not in self.static_at:
                    # get_objects_at(...)[0] -> IndexError off the grid (cooking_world.py:119, :160); the pick-up-special
                    # path never looks the static object up (:138-154): nothing is there, nothing happens
                    if a != 6:
                        self.error |= 64
                    continue
                st = self.static_at[cell]
                if a == 5:
                    self._primary(ag, cell, st, self._scan(*cell))
                elif a == 6:
                    self._pickup_special(ag, cell)
                else:
                    self._execute(ag, cell, st)

    def _pickup_special(self, ag, cell):
        """resolve_interaction_pick_up_special (cooking_world.py:138-154): take the last item off the
        one plate at the faced cell."""
        if self._agent_on(cell):
            return
        dyn = self._scan(*cell)
        if ag.holding is None and dyn:
            plates = [d for d in dyn if d.tr.get("plate", False)]
            if len(plates) == 1 and plates[0].content:
                obj = plates[0].content.pop(-1)
                ag.holding = obj
                self._move_obj(obj, ag.x, ag.y)
                self._ev("pickup_special")

    def _interact(self, ag, cell):
        """resolve_interaction (action_scheme3.py:37-43)."""
        st = self.static_at[cell]
        dyn = self._scan(*cell)
        if st.type in ("Cutboard", "Blender") and any(not d.done() for d in dyn):
            self._execute(ag, cell, st)
        else:
            self._primary(ag, cell, st, dyn)

    def _agent_on(self, cell):
        return any((o.x, o.y) == cell for o in self.agents)     # every agent, active or not

    def _execute(self, ag, cell, st):
        """resolve_execute_action (cooking_world.py:156-170)."""
        if self._agent_on(cell):
            return
        if st.type == "Cutboard":                                # Cutboard.action :250-269
            if not st.ready:
                return
            for o in list(st.content):
                if not o.tr.get("chop", False):
                    return
                if o.tr.get("spawn", False):                     # Bread.chop :738-745
                    if o.chopped:
                        continue
                    o.chopped = True
                    self._ev("chop_bread")
                    twin = _Obj("Bread", o.x, o.y)
                    twin.chopped = True
                    st.content.append(twin)                      # :263-264 (no free refresh here)
                    st.ready = False
                    self.by_type["Bread"].append(twin)           # cooking_world.py:168-170
                    return
                if o.chopped:                                    # ChopFood.chop :250-254
                    continue
                o.chopped = True
                st.ready = False
                self._ev("chop")
                return
            self.error |= 1          # quirk C-12: action() would return None -> TypeError
        elif st.type == "Blender":                               # Blender.action :356-360
            if st.ready:
                st.toggle = not st.toggle
                self._ev("blender_toggle")

    def _releases(self, st):
        """StaticObject.releases() including its side effects."""
        if st.type == "Deliversquare":                           # :117-118
            return False
        if st.type == "Cutboard":                                # :275-278
            if len(st.content) == 1:
                st.ready = False
            return True
        if st.type == "Blender":                                 # :340-346
            if st.toggle:
                return False
            if len(st.content) - 1 == 0:
                st.ready = False
            return True
        return True                                              # Floor/Counter/Switch/Block

    def _accepts(self, holder, o):
        """accepts() of every ContentObject."""
        t = holder.type
        if t in ("Counter", "Deliversquare"):                    # :64-66, :107-108
            return len(holder.content) < 1
        if t == "Cutboard":                                      # :271-273
            return o.tr.get("chop", False) and len(holder.content) < 1 and not o.chopped
        if t == "Blender":                                       # :337-338
            return (o.tr.get("blend", False) and not holder.toggle and len(holder.content) + 1 <= 1
                    and o.blend == 0)
        if t == "Plate":                                         # :408-409
            return o.is_food() and o.done() and len(holder.content) < 64
        return False                                             # Floor, Switch, Block

    def _add_content(self, holder, o):
        if holder.type in ("Cutboard", "Blender"):               # :280-288, :348-354
            holder.ready = True
        holder.content.append(o)
        _refresh_free(holder.content)

    def _primary(self, ag, cell, st, dyn):
        """resolve_primary_interaction (cooking_world.py:114-136)."""
        if self._agent_on(cell):
            return
        h = ag.holding
        if h is None:
            if not dyn:
                return
            if self._releases(st):
                grab = dyn[-1]
                for o in dyn:
                    if o.free:
                        grab = o
                        break
                if any(grab is c for c in st.content):
                    self._ev("grab_plate" if grab.tr.get("plate", False) else "grab")
                    ag.holding = grab                            # Agent.grab :786-788
                    self._move_obj(grab, ag.x, ag.y)
                    st.content.remove(grab)
            return
        # attempt_merge (cooking_world.py:243-261)
        plates = [d for d in dyn if d.tr.get("plate", False)]
        if len(plates) == 1:
            if self._accepts(plates[0], h):
                self._ev("merge_onto_plate")
                self._add_content(plates[0], h)
                self._move_obj(h, cell[0], cell[1])              # put_down :790-792
                ag.holding = None
        elif h.tr.get("plate", False) and dyn:
            p = dyn[-1]
            if self._accepts(h, p):
                self._ev("merge_scoop")
                self._add_content(h, p)
                self._move_obj(p, ag.x, ag.y)
                if any(p is c for c in st.content):
                    st.content.remove(p)
                else:
                    self.error |= 2      # list.remove would raise ValueError
        else:
            if self._accepts(st, h):
                self._ev("put_" + st.type)
                self._add_content(st, h)
                self._move_obj(h, cell[0], cell[1])
                ag.holding = None

    def _containers(self):
        for typ, lst in self.by_type.items():
            for o in lst:
                yield o

    def _progress_world(self):
        """progress_world (cooking_world.py:77-88); Blender.process (world_objects.py:321-335)."""
        for b in self.by_type.get("Blender", []):
            if b.content and b.toggle:
                for c in b.content:                              # BlenderFood.blend :266-273
                    if c.done():
                        continue
                    if c.blend in (0, 1):
                        self._ev("blend")
                        c.progress -= 1
                        c.blend = 1 if c.progress > 0 else 2
                if all(c.blend == 2 for c in b.content):
                    b.toggle = False
                    b.ready = False
                    for c in b.content:
                        c.progress = 1
        for o in self._containers():
            if o.content:
                _refresh_free(o.content)

    def _linked(self):
        """resolve_linked_interactions (cooking_world.py:90-92).  Level ATTRIBUTES are never
        applied (parsing.py:50-51 quirk), so every Switch and Block shares group None and a
        pressed Switch flips every Block; a second Switch in the group would make the
        reference raise AttributeError (Switch has no switch_state)."""
        switches = self.by_type.get("Switch", [])
        for s in switches:
            if s.pressed:
                if len(switches) > 1:
                    self.error |= 4
                for b in self.by_type.get("Block", []):
                    b.walkable = not b.walkable                  # Block.switch_state :215-216
            s.pressed = False

    def _agent_spawn(self):
        """handle_agent_spawn / despawn_agent / respawn_agent (cooking_world.py:267-290)."""
        for i in range(len(self.agents)):
            if self.grace[i] > 0:
                self.grace[i] -= 1
                continue
            if self.active.count(True) > 1 and self.active[i] and self._u() < self.despawn_rate:
                if self.agents[i].holding is None:               # :279-284
                    self.active[i] = False
                    self.status_changed[i] = True
            elif not self.active[i] and self._u() < self.respawn_rate:
                self.active[i] = True
                self.status_changed[i] = True
                self.grace[i] = self.grace_period
                self.agents[i].x, self.agents[i].y = self._spawn_location(i)

    def _u(self):
        return self.spawn_stream.uniform() if self.spawn_stream is not None else 1.0

    def _spawn_location(self, i):
        """parsing.generate_location (parsing.py:154-167)."""
        xs, ys = self.spawn_ranges[i]
        for _ in range(1002):
            x = self.spawn_stream.choice(xs)
            y = self.spawn_stream.choice(ys)
            st = self.static_at.get((x, y))
            if st is not None and st.type == "Floor" and not self._agent_on((x, y)):
                return int(x), int(y)
        self.error |= 8
        return self.agents[i].x, self.agents[i].y

    # ----------------------------------------------------------------- recipes / rewards
    def _update_recipe(self, node_list):
        """Recipe.update_recipe_state + check_conditions (recipe.py:77-104)."""
        for node in reversed(node_list):
            node.marked = False
            node.hits = []
            if not all(k.marked for k in node.kids):
                continue
            for o in self.by_type.get(node.type, []):
                if node.cond == "chopped" and not o.chopped:
                    continue
                if node.cond == "mashed" and o.blend != 2:
                    continue
                if all(any((h.x, h.y) == (o.x, o.y) for h in k.hits) for k in node.kids):
                    node.hits.append(o)
                    node.marked = True

    def _rewards(self, active_start):
        """compute_rewards / compute_truncated + the agent mapping of accumulated_step
        (cooking_env.py:250-262, 290-350)."""
        A = len(self.agents)
        n_rel = sum(self.relevant)
        # compute_truncated (:333-350)
        if self.t >= self.max_steps:
            if self.n_live < A:
                # quirk C-9: `[False] * self.num_agents` (:337) is sized by the *live* agent
                # list, so the reference raises IndexError at :348 / :252.  Defined
                # behaviour here: flag it and truncate every relevant agent.
                self.error |= 16
            trunc = [True] * n_rel
            self.active = [False] * A
            self.status_changed = list(self.relevant)
        else:
            trunc = [False] * n_rel
        k = 0
        for i in range(A):
            if not self.relevant[i]:
                continue
            if self.status_changed[i] and not self.active[i]:
                trunc[k] = True
            k += 1
        rs = self.reward_scheme
        rewards = [0] * max(len(self.recipes), A, len(RECIPES))   # :291-293 sizes by the book
        for r, rec in enumerate(self.recipes):
            before = sum(1 for n in rec if not n.marked)
            was = rec[0].marked
            self._update_recipe(rec)
            after = sum(1 for n in rec if not n.marked)
            now = rec[0].marked
            v = 0
            v += (before - after) * rs["recipe_node_reward"]
            if now and not was:
                self._ev("recipe_done")
            if was and not now:
                self._ev("recipe_undone")
            v += (now and not was) * rs["recipe_reward"]
            v += ((not now) and was) * rs["recipe_penalty"]
            v += rs["max_time_penalty"] / self.max_steps
            rewards[r] = v
        comp = [rec[0].marked for rec in self.recipes]
        done = all(comp) if self.end_all else any(comp)
        # accumulated_step (:250-262): the k-th *relevant* agent receives entry k
        self.rewards = [0.0] * A
        self.terminated = [False] * A
        self.truncated = [False] * A
        k = 0
        for i in range(A):
            if not (self.active[i] or self.status_changed[i]):
                continue
            self.rewards[i] = float(rewards[k])
            self.terminated[i] = bool(done)
            self.truncated[i] = bool(trunc[k])
            k += 1
        self.relevant = [self.active[i] or self.status_changed[i] for i in range(A)]
        self.n_live = sum(self.relevant)

    # ----------------------------------------------------------------- observation
    def observe(self, i):
        """get_feature_vector (cooking_env.py:352-373) -> float64 [L]."""
        me = self.agents[i]
        out = []
        for typ, num in self.meta:
            tr = TRAITS[typ]
            n = 0
            if typ == "Agent":
                for ag in self.agents:
                    f = [ag.x, ag.y] + [int(ag.orientation == k) for k in (1, 2, 3, 4)] + [1]
                    if ag is me:
                        f[0] = f[0] / self.W
                        f[1] = f[1] / self.H
                    else:
                        f[0] = (f[0] - me.x) / self.W
                        f[1] = (f[1] - me.y) / self.H
                    out.extend(f)
                    n += 1
            else:
                for o in self.by_type.get(typ, []):
                    f = self._features(o)
                    if f:
                        f[0] = (f[0] - me.x) / self.W
                        f[1] = (f[1] - me.y) / self.H
                    out.extend(f)
                    n += 1
            if n > num:
                self.error |= 32             # vector would silently grow (:371)
            out.extend([0] * (num - n) * tr["fv"])
        return np.array(out, dtype=np.float64)

    def _features(self, o):
        """feature_vector_representation of every class (world_objects.py:80,123,174,221,
        293,369,414,447,483,519,555,595,628,668,704,754)."""
        t = o.type
        if t == "Floor":
            return []
        if t == "Switch":
            return [o.x, o.y, int(o.switch_active), 1]
        if t == "Block":
            return [o.x, o.y, int(o.walkable), 1]
        tr = o.tr
        if tr["kind"] == "static" or tr.get("plate", False):
            return [o.x, o.y, 1]
        if tr.get("blend", False):
            return [o.x, o.y, int(not o.done()), int(o.chopped), int(o.blend == 2), 1]
        return [o.x, o.y, int(not o.done()), int(o.chopped), 1]

    def feature_length(self):
        return sum(TRAITS[t]["fv"] * n for t, n in self.meta)    # cooking_env.py:114-117

    # ----------------------------------------------------------------- canonical export
    def export_state(self):
        """Same arrays as oracle/ref_dump.dump_state."""
        dyn, sta = {}, {}
        nd = ns = 0
        for typ, num in self.meta:
            k = TRAITS[typ]["kind"]
            if k == "dynamic":
                dyn[typ] = nd
                nd += num
            elif k == "static":
                sta[typ] = ns
                ns += num
        slot = {}
        for typ, base in dyn.items():
            for k, o in enumerate(self.by_type.get(typ, [])):
                slot[id(o)] = base + k
        A = len(self.agents)
        agents = np.zeros((A, 6), np.int16)
        objs = np.zeros((nd, 9), np.int16)
        statics = np.zeros((ns, 4), np.int16)
        for typ, base in dyn.items():
            for k, o in enumerate(self.by_type.get(typ, [])):
                objs[base + k, :6] = (1, o.x, o.y, int(o.chopped), o.blend, int(o.free))
        for i, ag in enumerate(self.agents):
            hs = slot[id(ag.holding)] if ag.holding is not None else -1
            agents[i] = (ag.x, ag.y, ag.orientation, hs, int(self.active[i]), self.grace[i])
            if hs >= 0:
                objs[hs, 6:9] = (0, i, 0)
        for typ, lst in self.by_type.items():
            for o in lst:
                for pos, c in enumerate(o.content):
                    if o.tr["kind"] == "static":
                        objs[slot[id(c)], 6:9] = (1, o.y * self.W + o.x, pos)
                    else:
                        objs[slot[id(c)], 6:9] = (2, slot[id(o)], pos)
        for typ, base in sta.items():
            for k, o in enumerate(self.by_type.get(typ, [])):
                bits = (1 if o.ready else 0) | (2 if o.toggle else 0) | (4 if o.switch_active else 0) \
                    | (8 if o.walkable else 0) | (16 if o.pressed else 0)
                statics[base + k] = (1, o.x, o.y, bits)
        marks = np.zeros(len(self.recipes), np.int32)
        for r, rec in enumerate(self.recipes):
            for k, n in enumerate(rec):
                if n.marked:
                    marks[r] |= 1 << k
        return {"agents": agents, "objs": objs, "statics": statics, "marks": marks,
                "t": np.int32(self.t)}

    # ----------------------------------------------------------------- test helpers
    def teleport(self, i, x, y):
        """Agent.move_to (world_objects.py:794-797) — used by directed scenarios only."""
        ag = self.agents[i]
        ag.x, ag.y = x, y
        if ag.holding is not None:
            self._move_obj(ag.holding, x, y)
