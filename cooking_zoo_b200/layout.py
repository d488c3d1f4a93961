"""Initial-layout sampling: the reference's level parser with the RNG made explicit.

Reference: cooking_zoo/cooking_world/engine/parsing.py:5-151 draws placements from the
*global* `random` module.  `sample_layout(level, meta, num_agents, rng)` consumes
`rng.random()` / `rng.sample(seq, 1)` in exactly the same order, so
`random.Random(s)` here reproduces the world the reference builds after `random.seed(s)`
(SURVEY.md §7 stage 2, Appendix C-11).  The result is plain data:

    {"width", "height", "meta": [[type, count]...],
     "objects": [[type, [[x, y]...]]...]   # world_objects insertion order, list order
     "agents": [[x, y]...], "agent_spawn": [[xs, ys]...]}
"""
import math
from fractions import Fraction

from .entities import entity


class LayoutError(ValueError):
    pass


class TooManyLayouts(Exception):
    """enumerate_layouts gave up: the level has more distinct initial layouts than the caller wants to pool"""


def sample_layout(level_object, meta, num_agents, rng):
    meta_count = dict(meta)
    loaded = {}
    by_type = {}            # insertion ordered: type -> [(x, y)]
    static_at = {}          # (x, y) -> type name
    dynamic_at = set()

    def add(typ, loc):
        by_type.setdefault(typ, []).append(loc)

    def count_one(name):
        if name not in meta_count:
            raise KeyError(name)                               # reference: KeyError on meta lookup
        if meta_count[name] <= loaded.get(name, 0):
            raise LayoutError(f"Too many {name} objects loaded")
        loaded[name] = loaded.get(name, 0) + 1

    # parse_level_layout (parsing.py:5-18)
    x = y = 0
    for y, line in enumerate(level_object["LEVEL_LAYOUT"].splitlines()):
        for x, ch in enumerate(line):
            typ = "Counter" if ch == "-" else "Floor"
            add(typ, (x, y))
            static_at[(x, y)] = typ
    width, height = x + 1, y + 1

    def draw(spec, what):
        px = rng.sample(spec["X_POSITION"], 1)[0]
        py = rng.sample(spec["Y_POSITION"], 1)[0]
        if px < 0 or py < 0 or px > width or py > height:
            raise LayoutError(f"Position {px} {py} of {what} is out of bounds set by the level layout!")
        return px, py

    # parse_static_objects (parsing.py:21-76)
    for entry in level_object["STATIC_OBJECTS"]:
        name = next(iter(entry))
        spec = entry[name]
        for _ in range(spec["COUNT"]):
            tries = 0
            while True:
                if "OPTIONAL" in spec and spec["OPTIONAL"] <= rng.random():
                    break
                px, py = draw(spec, f"object {name}")
                under = static_at.get((px, py))
                if under in ("Counter", "Floor"):
                    count_one(name)
                    by_type[under].remove((px, py))
                    if entity(name).kind != "static":
                        raise LayoutError(f"{name} is not a static object")
                    add(name, (px, py))
                    static_at[(px, py)] = name
                    break
                tries += 1
                if tries > 10000:
                    raise LayoutError(f"Can't find valid position for object: {entry}")

    # parse_dynamic_objects (parsing.py:79-115)
    excluded = [list(p) for p in level_object["DYNAMIC_EXCLUDED_POSITIONS"]]
    for entry in level_object["DYNAMIC_OBJECTS"]:
        name = next(iter(entry))
        spec = entry[name]
        for _ in range(spec["COUNT"]):
            tries = 0
            while True:
                if "OPTIONAL" in spec and spec["OPTIONAL"] <= rng.random():
                    break
                px, py = draw(spec, f"object {name}")
                if static_at.get((px, py)) == "Counter" and (px, py) not in dynamic_at \
                        and [px, py] not in excluded:
                    count_one(name)
                    if entity(name).kind != "dynamic":
                        raise LayoutError(f"{name} is not a dynamic object")
                    add(name, (px, py))
                    dynamic_at.add((px, py))
                    break
                tries += 1
                if tries > 10000:
                    raise LayoutError(f"Can't find valid position for object: {entry}")

    # parse_agents (parsing.py:118-151)
    agents, spawn = [], []
    placed = 0
    done = False
    for spec in level_object["AGENTS"]:
        if done:
            break
        for _ in range(spec["MAX_COUNT"]):
            placed += 1
            if placed > num_agents:
                done = True
                break
            tries = 0
            while True:
                px, py = draw(spec, "agent")
                if (px, py) not in agents and static_at.get((px, py)) == "Floor":
                    count_one("Agent")
                    agents.append((int(px), int(py)))
                    spawn.append([list(spec["X_POSITION"]), list(spec["Y_POSITION"])])
                    break
                tries += 1
                if tries > 1000:
                    raise LayoutError(f"Can't find valid position for agent: {spec}")

    return {
        "width": width, "height": height,
        "meta": [[k, int(v)] for k, v in meta],
        "objects": [[t, [list(p) for p in locs]] for t, locs in by_type.items() if locs],
        "agents": [list(p) for p in agents],
        "agent_spawn": spawn,
    }


def enumerate_layouts(level_object, meta, num_agents, max_layouts=4096):
    """The EXACT distribution of the reference's level parser: every layout `sample_layout` can return, with its
    probability as a Fraction -> [(layout, probability)], probabilities summing to 1.

    Each placement of parsing.py is a rejection loop whose iterations are independent given the world built so far
    (:25-76, :86-115, :127-151): with probability 1 - OPTIONAL the object is skipped, otherwise a cell is drawn as
    (uniform pick from X_POSITION, uniform pick from Y_POSITION) and rejected cells repeat the whole iteration.  The
    accepted outcome is therefore distributed as one iteration conditioned on not being rejected, which makes the
    parser a finite tree of independent choices (the time-outs after 10001 / 1001 straight rejections are unreachable
    unless no cell is valid at all, which raises here like there).  Layouts reached along several paths are merged.
    Raises TooManyLayouts when the tree has more than `max_layouts` leaves."""
    meta_count = dict(meta)
    lines = level_object["LEVEL_LAYOUT"].splitlines()
    width, height = len(lines[-1]), len(lines)      # the reference takes the width from the last row (parsing.py:17)
    base_static = {}
    base_types = {}
    for y, line in enumerate(lines):
        for x, ch in enumerate(line):
            typ = "Counter" if ch == "-" else "Floor"
            base_types.setdefault(typ, []).append((x, y))
            base_static[(x, y)] = typ

    tasks = []
    for entry in level_object["STATIC_OBJECTS"]:
        name = next(iter(entry))
        tasks += [("static", name, entry[name])] * entry[name]["COUNT"]
    for entry in level_object["DYNAMIC_OBJECTS"]:
        name = next(iter(entry))
        tasks += [("dynamic", name, entry[name])] * entry[name]["COUNT"]
    placed = 0
    for spec in level_object["AGENTS"]:
        for _ in range(spec["MAX_COUNT"]):
            placed += 1
            if placed <= num_agents:
                tasks.append(("agent", "Agent", spec))
    excluded = {tuple(p) for p in level_object["DYNAMIC_EXCLUDED_POSITIONS"]}

    def cells_with_weight(spec, what):
        xs, ys = spec["X_POSITION"], spec["Y_POSITION"]
        out = {}
        for px in xs:
            for py in ys:
                if px < 0 or py < 0 or px > width or py > height:
                    raise LayoutError(f"Position {px} {py} of {what} is out of bounds set by the level layout!")
                out[(px, py)] = out.get((px, py), 0) + Fraction(1, len(xs) * len(ys))
        return out

    leaves = {}
    n_leaves = [0]

    def finish(world, prob):
        by_type, agents, spawn = world["by_type"], world["agents"], world["spawn"]
        layout = {"width": width, "height": height, "meta": [[k, int(v)] for k, v in meta],
                  "objects": [[t, [list(p) for p in locs]] for t, locs in by_type if locs],
                  "agents": [list(p) for p in agents], "agent_spawn": spawn}
        key = repr(layout)
        if key in leaves:
            leaves[key][1] += prob
        else:
            n_leaves[0] += 1
            if n_leaves[0] > max_layouts:
                raise TooManyLayouts(f"more than {max_layouts} distinct initial layouts")
            leaves[key] = [layout, prob]

    def walk(k, world, prob):
        if k == len(tasks):
            return finish(world, prob)
        kind, name, spec = tasks[k]
        p_opt = Fraction(1)
        if "OPTIONAL" in spec and kind != "agent":
            # `OPTIONAL <= random.random()` skips, and random() is k / 2**53: P(place) = ceil(OPTIONAL * 2**53) / 2**53
            p_opt = Fraction(min(max(math.ceil(Fraction(spec["OPTIONAL"]) * (1 << 53)), 0), 1 << 53), 1 << 53)
        static_at, dynamic_at, agents = world["static_at"], world["dynamic_at"], world["agents"]
        valid = {}
        for cell, w in cells_with_weight(spec, f"object {name}" if kind != "agent" else "agent").items():
            under = static_at.get(cell)
            if kind == "static":
                ok = under in ("Counter", "Floor")
            elif kind == "dynamic":
                ok = under == "Counter" and cell not in dynamic_at and cell not in excluded
            else:
                ok = cell not in agents and under == "Floor"
            if ok:
                valid[cell] = w
        p_valid = p_opt * sum(valid.values(), Fraction(0))
        p_skip = 1 - p_opt
        norm = p_valid + p_skip
        if norm == 0:
            raise LayoutError(f"Can't find valid position for {kind}: {spec}")
        if valid:          # the count check fires only when a cell is accepted (parsing.py:43, :57, :102, :137)
            if name not in meta_count:
                raise KeyError(name)
            if meta_count[name] <= world["loaded"].get(name, 0):
                raise LayoutError(f"Too many {name} objects loaded")
            if kind != "agent" and entity(name).kind != kind:
                raise LayoutError(f"{name} is not a {kind} object")
        if p_skip:
            walk(k + 1, world, prob * p_skip / norm)
        for cell, w in valid.items():
            nw = {"static_at": static_at, "dynamic_at": dynamic_at, "agents": agents, "spawn": world["spawn"],
                  "loaded": dict(world["loaded"]), "by_type": [(t, list(l)) for t, l in world["by_type"]]}
            nw["loaded"][name] = nw["loaded"].get(name, 0) + 1
            types = dict(nw["by_type"])

            def add(typ, loc):
                if typ in types:
                    types[typ].append(loc)
                else:
                    lst = [loc]
                    types[typ] = lst
                    nw["by_type"].append((typ, lst))
            if kind == "static":
                under = static_at[cell]
                types[under].remove(cell)
                add(name, cell)
                nw["static_at"] = dict(static_at)
                nw["static_at"][cell] = name
            elif kind == "dynamic":
                add(name, cell)
                nw["dynamic_at"] = dynamic_at | {cell}
            else:
                nw["agents"] = agents + [(int(cell[0]), int(cell[1]))]
                nw["spawn"] = world["spawn"] + [[list(spec["X_POSITION"]), list(spec["Y_POSITION"])]]
            walk(k + 1, nw, prob * p_opt * w / norm)

    world0 = {"static_at": base_static, "dynamic_at": frozenset(), "agents": [], "spawn": [], "loaded": {},
              "by_type": [(t, list(l)) for t, l in base_types.items()]}
    walk(0, world0, Fraction(1))
    out = [(lay, p) for lay, p in leaves.values()]
    assert sum(p for _, p in out) == 1
    return out
