"""Dump a live reference `CookingEnvironment` into the canonical arrays parity tests compare.

Test infrastructure (build container only; needs /root/reference through ref_loader).
Nothing here re-implements game logic: it only *reads* attributes of reference objects.

Canonical numbering (shared convention with oracle/cz_oracle.py and
cooking_zoo_b200.BatchedCookingEnv.export_state):

* dynamic slot  = for each meta-file entry whose class is a DynamicObject, in meta order,
                  `count` consecutive slots; object k of `world.world_objects[type]` (list
                  order) sits in slot base[type]+k.
* static slot   = same for the static classes named in the meta file (Cutboard, Counter,
                  Blender, Deliversquare, Block, Switch ...; Floor is not in the meta file).
* agents[A,6]   = x, y, orientation, holding dynamic slot (-1 none), active, grace
* objs[D,9]     = present, x, y, chopped, blend (0 fresh,1 in progress,2 mashed), free,
                  cont_kind (0 held by agent / 1 static content / 2 plate content),
                  cont_id (agent index / cell y*W+x / plate dynamic slot), pos in content
* statics[S,4]  = present, x, y, bits (1 status READY | 2 toggle | 4 switch_active |
                  8 walkable | 16 button_pressed)
* marks[R]      = bit k set iff recipe_graphs[r].node_list[k].marked
"""
import numpy as np

BLEND_CODE = {"Fresh": 0, "InProgress": 1, "Mashed": 2}


def _classes():
    from cooking_zoo.cooking_world import world_objects as wo
    from cooking_zoo.cooking_world import abstract_classes as ac
    return wo, ac


def slot_bases(env):
    """(dyn_base, dyn_total, static_base, static_total) from the meta file, in meta order."""
    wo, ac = _classes()
    dyn, sta = {}, {}
    nd = ns = 0
    for name, num in env.world.meta_object_information.items():
        cls = wo.StringToClass[name]
        if issubclass(cls, ac.DynamicObject):
            dyn[name] = (nd, num)
            nd += num
        elif issubclass(cls, ac.StaticObject):
            sta[name] = (ns, num)
            ns += num
    return dyn, nd, sta, ns


def describe_layout(env):
    """Plain-data description of the world as loaded (everything a restatement needs)."""
    wo, ac = _classes()
    w = env.world
    objects = []
    for name, lst in w.world_objects.items():
        if lst:
            objects.append([name, [[int(o.location[0]), int(o.location[1])] for o in lst]])
    return {
        "width": int(w.width), "height": int(w.height),
        "meta": [[k, int(v)] for k, v in w.meta_object_information.items()],
        "objects": objects,
        "agents": [[int(a.location[0]), int(a.location[1])] for a in w.agents],
        "agent_spawn": [[list(map(int, xs)), list(map(int, ys))] for xs, ys in w.agent_spawn_locations],
    }


def dump_state(env):
    wo, ac = _classes()
    w = env.world
    W = w.width
    dyn, nd, sta, ns = slot_bases(env)
    slot_of = {}
    for name, (base, num) in dyn.items():
        for k, o in enumerate(w.world_objects.get(name, [])):
            assert k < num, f"more {name} objects than meta slots"
            slot_of[id(o)] = base + k
    A = len(w.agents)
    agents = np.zeros((A, 6), np.int16)
    objs = np.zeros((nd, 9), np.int16)
    statics = np.zeros((ns, 4), np.int16)
    seen = set()
    for name, (base, num) in dyn.items():
        for k, o in enumerate(w.world_objects.get(name, [])):
            s = base + k
            objs[s, 0] = 1
            objs[s, 1], objs[s, 2] = o.location
            objs[s, 3] = int(getattr(o, "chop_state", None) == wo.ChopFoodStates.CHOPPED)
            objs[s, 4] = BLEND_CODE[o.blend_state.value] if hasattr(o, "blend_state") else 0
            objs[s, 5] = int(o.free)
    for i, a in enumerate(w.agents):
        agents[i, 0], agents[i, 1] = a.location
        agents[i, 2] = a.orientation
        agents[i, 3] = slot_of[id(a.holding)] if a.holding is not None else -1
        agents[i, 4] = int(w.active_agents[i])
        agents[i, 5] = int(w.agent_grace_period[i])
        if a.holding is not None:
            s = slot_of[id(a.holding)]
            objs[s, 6:9] = (0, i, 0)
            seen.add(s)
    for name, lst in w.world_objects.items():
        cls = wo.StringToClass[name]
        if not issubclass(cls, ac.ContentObject):
            continue
        for o in lst:
            for pos, c in enumerate(o.content):
                if isinstance(c, wo.Agent):
                    continue
                s = slot_of[id(c)]
                assert s not in seen, "object in two containers"
                seen.add(s)
                if issubclass(cls, ac.StaticObject):
                    objs[s, 6:9] = (1, o.location[1] * W + o.location[0], pos)
                else:
                    objs[s, 6:9] = (2, slot_of[id(o)], pos)
    assert len(seen) == int(objs[:, 0].sum()), "dynamic object neither held nor contained"
    for name, (base, num) in sta.items():
        for k, o in enumerate(w.world_objects.get(name, [])):
            assert k < num
            bits = 0
            if getattr(o, "status", None) == wo.ActionObjectState.READY:
                bits |= 1
            if getattr(o, "toggle", False):
                bits |= 2
            if getattr(o, "switch_active", False):
                bits |= 4
            if o.walkable:
                bits |= 8
            if getattr(o, "button_pressed", False):
                bits |= 16
            statics[base + k] = (1, o.location[0], o.location[1], bits)
    marks = np.zeros(len(env.recipe_graphs), np.int32)
    for r, rec in enumerate(env.recipe_graphs):
        for k, node in enumerate(rec.node_list):
            if node.marked:
                marks[r] |= 1 << k
    return {"agents": agents, "objs": objs, "statics": statics, "marks": marks,
            "t": np.int32(env.t)}


def observe_all(env):
    """float64 [A, L] feature vectors for every agent slot (cooking_env.py:352-373)."""
    return np.stack([np.asarray(env.get_feature_vector(a), dtype=np.float64) for a in env.possible_agents])


def step_outputs(env):
    """rewards f64[A], terminated u8[A], truncated u8[A], relevant u8[A] after accumulated_step.

    Agents absent from the reference's dicts (not relevant this step) get 0 / 0 / 0.
    """
    A = len(env.possible_agents)
    rew = np.zeros(A, np.float64)
    term = np.zeros(A, np.uint8)
    trunc = np.zeros(A, np.uint8)
    rel = np.zeros(A, np.uint8)
    for i, name in enumerate(env.possible_agents):
        if name in env.rewards:
            rel[i] = 1
            rew[i] = float(env.rewards[name])
            term[i] = int(bool(env.terminations[name]))
            trunc[i] = int(bool(env.truncations[name]))
    return rew, term, trunc, rel
