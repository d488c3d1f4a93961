"""Host-built tables of the device policy: the scripted cook's graph searches.

The reference cook (cooking_zoo/cooking_agents/base_agent.py) runs two breadth-first searches over
the Floor tiles on every decision: `reachable` (:94-126) and `walk_to_location` (:62-92).  Floor
tiles never move, so both are compiled once per static variant:

    reach[v][a]          bit b set  <=>  reachable(a, b): a == b, or b can be entered from a along
                         Floor tiles (the goal itself may be any cell)
    first_step[v][a][b]  the action walk_to_location returns for start a, goal b: neighbours are
                         expanded left, right, down, up (:27-33), a tile's path is the one of its first
                         discoverer, 0 when a == b or b cannot be entered
    lists[v][kind]       cells of every static kind in world_objects list order (the order
                         `observation["Counter"]` etc. are iterated in)

cz_policy.cuh reads them; tests/test_policy_oracle.py checks them against the queue-level
restatement in oracle/cz_policy.py for every pair of cells.
"""
import numpy as np

from . import entities as E
from .tables import ROW_VARIANT

MOVES = ((1, -1, 0), (2, 1, 0), (3, 0, 1), (4, 0, -1))     # action, dx, dy (base_agent.py:27-33)


def _search_from(start, floor, W, H):
    """one BFS over the Floor tiles from `start`: (cells that can be entered, first action towards each)"""
    first = {start: 0}
    order = [start]
    k = 0
    enter = {}                 # non-Floor cell -> first action of the earliest expanded neighbour
    while k < len(order):
        cur = order[k]
        k += 1
        for a, dx, dy in MOVES:
            n = (cur[0] + dx, cur[1] + dy)
            if not (0 <= n[0] < W and 0 <= n[1] < H):
                continue
            step = first[cur] or a
            if n in floor:
                if n not in first:
                    first[n] = step
                    order.append(n)
            elif n != start and n not in enter:
                enter[n] = step
    first.update(enter)
    first[start] = 0
    return first


def compile_policy_tables(t):
    """CompiledTables -> dict of numpy arrays for cz_policy_desc (include/cz_b200.h)"""
    V, W, H = t.num_variants, t.width, t.height
    lists = np.full((V, 8, 64), 0xFF, np.uint8)
    list_len = np.zeros((V, 8), np.uint8)
    reach = np.zeros((V, 64), np.uint64)
    first_step = np.zeros((V, 64, 64), np.uint8)
    col = t.num_dyn_slots + t.num_agents + ROW_VARIANT
    seen = set()
    for li, lay in enumerate(t.layouts):
        v = int(t.pool[li, col])
        if v in seen:
            continue
        seen.add(v)
        floor = set()
        for name, locs in lay["objects"]:
            et = E.entity(name)
            if et.kind != "static":
                continue
            cells = [y * 8 + x for x, y in locs]
            lists[v, et.static_code, :len(cells)] = cells
            list_len[v, et.static_code] = len(cells)
            if et.static_code == E.ST_FLOOR:
                floor = {(x, y) for x, y in locs}
        for y in range(H):
            for x in range(W):
                a = y * 8 + x
                bits = 0
                for (bx, by), act in _search_from((x, y), floor, W, H).items():
                    b = by * 8 + bx
                    bits |= 1 << b
                    first_step[v, a, b] = act
                reach[v, a] = bits
    assert len(seen) == V
    # the cook reads attributes named by the node's condition (base_agent.py:195): a condition on a type
    # without that attribute raises in the reference and cannot be compiled here
    for b, name in enumerate(t.recipe_names):
        for k in range(int(t.recipe_len[b])):
            node = int(t.recipe_nodes[b, k])
            cond = (node >> 9) & 3
            if not cond:
                continue
            ty = node & 255
            if node & 256 or (ty != 255 and not int(t.type_flags[ty]) & (E.TF_CHOP if cond == 1 else E.TF_BLEND)):
                raise ValueError(f"recipe {name}: node {k} tests a state its object type does not have")
    return {"lists": lists, "list_len": list_len, "reach": reach, "first_step": first_step}
