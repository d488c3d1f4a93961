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
from .entities import entity


class LayoutError(ValueError):
    pass


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
