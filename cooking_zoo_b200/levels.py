"""Level and meta-file definitions.

The reference reads JSON files named by `level` / `meta_file`, or any path ending in
".json" (cooking_zoo/cooking_world/engine/load_level.py:11-20, 39-52).  The three shipped
levels and the one shipped meta file are restated here as Python data (same keys and values
as cooking_zoo/utils/level/*.json and utils/meta_files/example.json) so that
`level="coop_test", meta_file="example"` keeps working without the reference installed;
".json" paths are loaded with the reference's schema.
"""
import json

_BORDERED_SPLIT = "-------\n-  -  -\n-  -  -\n-  -  -\n-  -  -\n-  -  -\n-------"
_BORDERED_OPEN = "-------\n-     -\n-     -\n-     -\n-     -\n-     -\n-------"
_CORNERS = [[0, 0], [3, 0], [6, 0], [0, 6], [3, 6], [6, 6]]
_TWO_AGENTS = [{"MAX_COUNT": 1, "X_POSITION": [1, 2], "Y_POSITION": [1, 2, 3, 4, 5]},
               {"MAX_COUNT": 1, "X_POSITION": [4, 5], "Y_POSITION": [1, 2, 3, 4, 5]}]


def _one(name, xs, ys, count=1, optional=None, attributes=None):
    body = {"COUNT": count, "X_POSITION": list(xs), "Y_POSITION": list(ys)}
    if optional is not None:
        body["OPTIONAL"] = optional
    if attributes is not None:
        body["ATTRIBUTES"] = attributes
    return {name: body}


def _kitchen(optional):
    """statics/dynamics shared by coop_test (no OPTIONAL) and coexistence_test."""
    so = (lambda p: p) if optional else (lambda p: None)
    statics = [_one("Cutboard", [6], [4], optional=so(0.7)), _one("Cutboard", [0], [1], optional=so(0.7)),
               _one("Cutboard", [3], [3], optional=so(0.7)), _one("Blender", [4], [6]),
               _one("Deliversquare", [2, 4], [0], count=2)]
    dynamics = [_one("Plate", [0], [4], optional=so(1.0)), _one("Plate", [6], [3], optional=so(1.0)),
                _one("Lettuce", [0], [5], optional=so(0.9)), _one("Tomato", [0], [2], optional=so(0.9)),
                _one("Banana", [6], [1], optional=so(0.9)), _one("Apple", [5], [6], optional=so(0.9)),
                _one("Watermelon", [6], [2], optional=so(0.9)), _one("Bread", [0], [3], optional=so(0.9)),
                _one("Bread", [5], [0], optional=so(0.9))]
    return statics, dynamics


def _coop_test():
    statics, dynamics = _kitchen(False)
    dynamics.append(_one("Carrot", [0, 1, 2], [1, 2, 3, 4, 5, 6]))
    return {"LEVEL_LAYOUT": _BORDERED_SPLIT, "STATIC_OBJECTS": statics, "DYNAMIC_OBJECTS": dynamics,
            "AGENTS": _TWO_AGENTS, "DYNAMIC_EXCLUDED_POSITIONS": _CORNERS}


def _coexistence_test():
    statics, dynamics = _kitchen(True)
    dynamics.append(_one("Carrot", [0, 1, 2, 3], [1, 2, 3, 4, 5, 6], optional=1.0))
    return {"LEVEL_LAYOUT": _BORDERED_SPLIT, "STATIC_OBJECTS": statics, "DYNAMIC_OBJECTS": dynamics,
            "AGENTS": _TWO_AGENTS, "DYNAMIC_EXCLUDED_POSITIONS": _CORNERS}


def _switch_test():
    statics = [_one("Cutboard", [6], [4]), _one("Cutboard", [0], [1]), _one("Blender", [4], [6]),
               _one("Deliversquare", [2, 4], [0], count=2),
               _one("Block", [2], [3], attributes={"walkable": False, "linked_group_id": 1}),
               _one("Switch", [4], [3], attributes={"linked_group_id": 1})]
    dynamics = [_one("Plate", [0], [4]), _one("Plate", [6], [3]), _one("Lettuce", [0], [5]),
                _one("Tomato", [0], [2]), _one("Banana", [6], [1]),
                _one("Carrot", [3, 4, 5, 6], [0, 1, 2, 3, 4, 5, 6])]
    return {"LEVEL_LAYOUT": _BORDERED_OPEN, "STATIC_OBJECTS": statics, "DYNAMIC_OBJECTS": dynamics,
            "AGENTS": _TWO_AGENTS, "DYNAMIC_EXCLUDED_POSITIONS": _CORNERS}


LEVELS = {"coop_test": _coop_test, "coexistence_test": _coexistence_test, "switch_test": _switch_test}

META_FILES = {
    "example": [("Cutboard", 3), ("Counter", 29), ("Blender", 2), ("Deliversquare", 2), ("Plate", 3),
                ("Tomato", 3), ("Onion", 3), ("Lettuce", 3), ("Carrot", 3), ("Banana", 3), ("Apple", 3),
                ("Watermelon", 3), ("Bread", 4), ("Agent", 2), ("Block", 2), ("Switch", 2)],
}


def load_level_object(level):
    """Level dict in the reference's JSON schema (load_level.py:11-20)."""
    if isinstance(level, dict):
        return level
    if level.endswith(".json"):
        with open(level) as f:
            return json.load(f)
    try:
        return LEVELS[level]()
    except KeyError:
        raise FileNotFoundError(f"unknown level {level!r}; pass a path ending in .json") from None


def load_meta(meta_file):
    """Ordered [(type name, max count)] (load_level.py:39-52)."""
    if isinstance(meta_file, (list, tuple)):
        return [(str(k), int(v)) for k, v in meta_file]
    if meta_file.endswith(".json"):
        with open(meta_file) as f:
            raw = json.load(f)
        return [(list(d.keys())[0], int(list(d.values())[0])) for d in raw]
    try:
        return list(META_FILES[meta_file])
    except KeyError:
        raise FileNotFoundError(f"unknown meta file {meta_file!r}; pass a path ending in .json") from None
