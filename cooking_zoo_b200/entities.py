"""Entity-type registry: the reference's class hierarchy flattened into constant rows.

Reference: cooking_zoo/cooking_world/world_objects.py (concrete classes) and
cooking_zoo/cooking_world/abstract_classes.py (mixins).  Every behavioural difference
between two classes that the hot path can observe is a column here; the CUDA kernels
dispatch on these columns instead of on a class.
"""
from dataclasses import dataclass

# static kind codes stored in the per-cell grid table (low nibble of a grid byte)
ST_NONE, ST_FLOOR, ST_COUNTER, ST_CUTBOARD, ST_BLENDER, ST_DELIVER, ST_SWITCH, ST_BLOCK = range(8)

# dynamic type flag bits (type_flags table)
TF_PLATE, TF_CHOP, TF_BLEND, TF_SPAWN = 1, 2, 4, 8

# observation feature layouts (what follows x, y in feature_vector_representation)
FV_NONE = 0        # not observed (Floor)                                   world_objects.py:38
FV_ONE = 1         # [x, y, 1]                        Counter/Cutboard/Blender/Deliversquare/Plate
FV_CHOP = 2        # [x, y, !done, chopped, 1]        pure ChopFood (:447,483,519,595,668,704,754)
FV_CHOPBLEND = 3   # [x, y, !done, chopped, mashed, 1] Carrot/Banana (:555,628)
FV_AGENT = 4       # [x, y, o==1, o==2, o==3, o==4, 1] Agent (:806)
FV_SWITCH = 5      # [x, y, switch_active, 1]         Switch (:174)
FV_BLOCK = 6       # [x, y, walkable, 1]              Block (:221)
FV_LEN = {FV_NONE: 0, FV_ONE: 3, FV_CHOP: 5, FV_CHOPBLEND: 6, FV_AGENT: 7, FV_SWITCH: 4, FV_BLOCK: 4}


@dataclass(frozen=True)
class EntityType:
    name: str
    kind: str            # "static" | "dynamic" | "agent"
    fv: int              # FV_* layout
    static_code: int = ST_NONE
    walkable: bool = False   # value at construction (Block can flip later)
    flags: int = 0       # TF_* for dynamic types


def _s(name, code, fv, walkable=False):
    return EntityType(name, "static", fv, static_code=code, walkable=walkable)


def _d(name, fv, flags):
    return EntityType(name, "dynamic", fv, flags=flags)


ENTITY_TYPES = {t.name: t for t in [
    _s("Floor", ST_FLOOR, FV_NONE, walkable=True),         # world_objects.py:17
    _s("Counter", ST_COUNTER, FV_ONE),                     # :57
    _s("Deliversquare", ST_DELIVER, FV_ONE),               # :101
    _s("Switch", ST_SWITCH, FV_SWITCH, walkable=True),     # :144
    _s("Block", ST_BLOCK, FV_BLOCK),                       # :195
    _s("Cutboard", ST_CUTBOARD, FV_ONE),                   # :242
    _s("Blender", ST_BLENDER, FV_ONE),                     # :314
    _d("Plate", FV_ONE, TF_PLATE),                         # :386
    _d("Onion", FV_CHOP, TF_CHOP),                         # :435
    _d("Tomato", FV_CHOP, TF_CHOP),                        # :471
    _d("Lettuce", FV_CHOP, TF_CHOP),                       # :507
    _d("Carrot", FV_CHOPBLEND, TF_CHOP | TF_BLEND),        # :543
    _d("Cucumber", FV_CHOP, TF_CHOP),                      # :583
    _d("Banana", FV_CHOPBLEND, TF_CHOP | TF_BLEND),        # :616
    _d("Apple", FV_CHOP, TF_CHOP),                         # :656
    _d("Watermelon", FV_CHOP, TF_CHOP),                    # :692
    _d("Bread", FV_CHOP, TF_CHOP | TF_SPAWN),              # :728 (chop() spawns a twin, :738-745)
    EntityType("Agent", "agent", FV_AGENT),                # :774
]}


def entity(name):
    try:
        return ENTITY_TYPES[name]
    except KeyError:
        raise KeyError(f"unknown object type {name!r} (not a class of the reference's world_objects)") from None
