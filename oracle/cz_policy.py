"""CPU oracle of the reference's scripted cook.

TEST INFRASTRUCTURE ONLY (same rule as oracle/cz_oracle.py): the product's device policy
(cooking_zoo_b200/csrc/cz_policy.cu) is checked against this file, never routed through it.

Restates `CookingAgent.step` (cooking_agents/cooking_agent.py:9-122) and the helpers of
`BaseAgent` (cooking_agents/base_agent.py:49-199) as a pure function of an OracleEnv's world:

    heuristic_action(env, i, recipe)      cooking_agent.py:9-24  (step / compute_optimal_action)
      _condition_action()                 cooking_agent.py:42-55 + base_agent.py:133-146
      _appliance_sequence()               base_agent.py:148-190  (generic_sequence)
      _contains_action()                  cooking_agent.py:26-40, 57-122
      reachable() / walk()                base_agent.py:62-92, 94-126

The reference agent keeps a `reachable` cache (base_agent.py:95-96); Floor tiles never move, so
the cache only memoises a pure function and a fresh agent per call is equivalent.  Where the
reference raises (no object of the node's type: `sorted([])[0]`, cooking_agent.py:49; no main
object but children to look for: `None.location`, :66; `closest` found nothing:
`tuple(None)`, base_agent.py:63) PolicyCrash is raised; the device policy reports action 0 and
sets CZ_POLICY_ERR for that agent.

Parity pins: tests/golden/policy_*.npz hold the raw CookingAgent actions recorded from the
unmodified reference next to the states they were computed on (tests/golden/make_golden.py).
"""
from collections import deque

from .cz_oracle import TRAITS, make_recipe

_MOVES = ((1, (-1, 0)), (2, (1, 0)), (3, (0, 1)), (4, (0, -1)))     # base_agent.py:27-33


class PolicyCrash(Exception):
    """the reference's CookingAgent would raise here"""


def _floor(env):
    return {(o.x, o.y) for o in env.by_type.get("Floor", [])}


def reachable(env, a, b, floor=None):
    """base_agent.py:94-126: b can be entered from a through Floor tiles only"""
    if a == b:
        return True
    floor = _floor(env) if floor is None else floor
    seen = {a}
    todo = deque([a])
    while todo:
        cx, cy = todo.popleft()
        for _, (dx, dy) in _MOVES:
            n = (cx + dx, cy + dy)
            if n == b:
                return True
            if n in floor and n not in seen:
                seen.add(n)
                todo.append(n)
    return False


def walk(env, start, goal, floor=None):
    """base_agent.py:62-92, including its queue discipline: tiles are marked when POPPED, so a
    tile can sit in the queue several times and is expanded every time it comes out."""
    if start == goal:
        return 0
    floor = _floor(env) if floor is None else floor
    visited = set()
    queue = deque([(start, 0)])          # (tile, first action of the path that reached it)
    pops = 0
    while queue:
        cur, first = queue.popleft()
        if cur == goal:
            return first
        visited.add(cur)
        pops += 1
        if pops > 200000:
            raise RuntimeError("walk(): queue blow-up")
        for a, (dx, dy) in _MOVES:
            n = (cur[0] + dx, cur[1] + dy)
            if n not in visited and (n in floor or n == goal):
                queue.append((n, first or a))
    return 0


def _d2(a, b):
    # BaseAgent.distance is sqrt of this integer; sqrt is monotone and exact enough to keep
    # distinct small integers distinct, so comparisons are done on the squares
    return (a[0] - b[0]) ** 2 + (a[1] - b[1]) ** 2


def _loc(o):
    return (o.x, o.y)


def _unmet(node, o):
    """check_node_conditions (base_agent.py:192-198): number of unmet conditions (0 or 1)"""
    if node.cond == "chopped":
        if not TRAITS[o.type].get("chop"):
            raise PolicyCrash("no chop_state")
        return 0 if o.chopped else 1
    if node.cond == "mashed":
        if not TRAITS[o.type].get("blend"):
            raise PolicyCrash("no blend_state")
        return 0 if o.blend == 2 else 1
    return 0


def _closest(env, origin, locs, floor):
    """BaseAgent.closest (base_agent.py:128-138): first strictly nearest reachable location"""
    best, best_d = None, None
    for loc in locs:
        if not reachable(env, origin, loc, floor):
            continue
        d = _d2(origin, loc)
        if best_d is None or d < best_d:
            best, best_d = loc, d
    if best is None:
        raise PolicyCrash("closest() found nothing")       # walk_to_location(None)
    return best


def _appliance_sequence(env, ag, me, kind, obj, floor):
    """generic_sequence (base_agent.py:148-190)"""
    apps = env.by_type.get(kind, [])
    near = [a for a in apps if reachable(env, me, _loc(a), floor)]
    for a in near:
        if any(c is obj for c in a.content):
            return walk(env, me, _loc(obj), floor)
    empty = [_loc(a) for a in near if not a.content]
    if ag.holding is obj:
        if empty:
            return walk(env, me, _closest(env, _loc(obj), empty, floor), floor)
        counters = [_loc(c) for c in env.by_type.get("Counter", []) if reachable(env, me, _loc(c), floor)]
        return walk(env, me, _closest(env, _loc(obj), counters, floor), floor)
    if empty:
        if ag.holding is not None:
            counters = [_loc(c) for c in env.by_type.get("Counter", [])
                        if reachable(env, me, _loc(c), floor) and not c.content]
            return walk(env, me, _closest(env, me, counters, floor), floor)
        return walk(env, me, _loc(obj), floor)
    return walk(env, me, _closest(env, me, [_loc(a) for a in near], floor), floor)


def _condition_action(env, ag, me, node, floor):
    """compute_condition_action (cooking_agent.py:42-55)"""
    objs = env.by_type.get(node.type, [])
    if not objs:
        raise PolicyCrash("no object of the node's type")
    best = min(objs, key=lambda o: (_unmet(node, o), _d2(me, _loc(o))))     # sorted(...)[0]: first minimum
    if node.cond and _unmet(node, best):
        kind = "Cutboard" if node.cond == "chopped" else "Blender"            # base_agent.py:141-147
        return _appliance_sequence(env, ag, me, kind, best, floor)
    return 0


def _near_objects(env, me, typ, floor):
    """convert_node_to_world_objects (cooking_agent.py:78-81)"""
    return [o for o in env.by_type.get(typ, []) if reachable(env, _loc(o), me, floor)]


def _contains_action(env, me, node, floor):
    """compute_contains_action (cooking_agent.py:26-40)"""
    # get_location_with_most_objects (:99-122)
    main, main_count = None, -1
    for m in _near_objects(env, me, node.type, floor):
        count = 0
        for kid in node.kids:
            for o in _near_objects(env, me, kid.type, floor):
                if _loc(o) == _loc(m) and _unmet(kid, o) == 0:
                    count += 1
        if count > main_count or (count == main_count and _d2(me, _loc(m)) < _d2(me, _loc(main))):
            main, main_count = m, count
    # get_best_contains_obj (:57-76)
    target, target_d = None, None
    for kid in node.kids:
        for o in _near_objects(env, me, kid.type, floor):
            if main is None:
                raise PolicyCrash("children to fetch but nowhere to bring them")
            if _loc(o) == _loc(main):
                continue
            d = _d2(me, _loc(o))
            if target_d is None or d < target_d:
                target, target_d = o, d
    if target is None:
        return 0
    if _loc(target) == me:
        return walk(env, me, _loc(main), floor)
    return walk(env, me, _loc(target), floor)


def heuristic_action(env, i, recipe):
    """The action CookingAgent(recipe, name-of-agent-i).step(symbolic observation) returns for the
    current world of `env` (an OracleEnv)."""
    ag = env.agents[i]
    me = (ag.x, ag.y)
    nodes = make_recipe(recipe)
    env._update_recipe(nodes)                               # cooking_agent.py:12
    node = next((n for n in reversed(nodes) if not n.marked), None)     # find_node, base_agent.py:49-53
    if node is None:
        return 0
    floor = _floor(env)
    act = _condition_action(env, ag, me, node, floor)
    if act:
        return act
    return _contains_action(env, me, node, floor)


def heuristic_actions(env, recipes):
    """one action per agent; -1 where the reference would raise"""
    out = []
    for i, r in enumerate(recipes):
        try:
            out.append(heuristic_action(env, i, r))
        except PolicyCrash:
            out.append(-1)
    return out
