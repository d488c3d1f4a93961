"""The recipe book as data, plus the reference's registration hook.

Reference: cooking_zoo/cooking_book/recipe_drawer.py:40-118 (node definitions, id allocation,
RECIPES), cooking_zoo/cooking_book/recipe.py:12-33 (RecipeNode, Recipe.node_list).
A recipe compiles to its `node_list` — root first, then `expand_child_nodes` order — where
every node is (type name, condition, positions of its children in the list); the kernels
evaluate the list back to front as cell bitmasks.
"""
from dataclasses import dataclass, field
from typing import List, Optional

COND_NONE, COND_CHOPPED, COND_MASHED = 0, 1, 2
_COND_BY_ATTR = {("chop_state", "Chopped"): COND_CHOPPED, ("blend_state", "Mashed"): COND_MASHED}


@dataclass
class RecipeNode:
    """Mirror of recipe.RecipeNode's constructor surface (recipe.py:14-24).

    `conditions` accepts the reference's [(attribute, value)] pairs where value is the enum
    or its string ("Chopped" / "Mashed"); only chop_state / blend_state can be compiled.
    """
    name: str
    id_num: int = -1
    conditions: Optional[list] = None
    contains: List["RecipeNode"] = field(default_factory=list)
    root_type: object = None
    objects_to_seek: Optional[list] = None

    def condition_code(self):
        conds = self.conditions or []
        if not conds:
            return COND_NONE
        if len(conds) != 1:
            raise ValueError("only single-condition recipe nodes can be compiled")
        attr, value = conds[0]
        value = getattr(value, "value", value)
        try:
            return _COND_BY_ATTR[(attr, value)]
        except KeyError:
            raise ValueError(f"recipe condition {attr}=={value!r} cannot be compiled") from None


@dataclass
class Recipe:
    root_node: RecipeNode
    num_goals: int = 0

    def node_list(self):
        def expand(n):
            out = list(n.contains)
            for c in n.contains:
                out.extend(expand(c))
            return out
        return [self.root_node] + expand(self.root_node)


def _default_book():
    ids = iter(range(1000))
    book = {}

    def leaf(key, typ, attr, val):
        book[key] = RecipeNode(name=typ, id_num=next(ids), conditions=[(attr, val)])

    for key, typ in [("ChoppedLettuce", "Lettuce"), ("ChoppedOnion", "Onion"), ("ChoppedTomato", "Tomato"),
                     ("ChoppedApple", "Apple"), ("ChoppedCucumber", "Cucumber"),
                     ("ChoppedWatermelon", "Watermelon"), ("ChoppedBanana", "Banana")]:
        leaf(key, typ, "chop_state", "Chopped")
    leaf("MashedBanana", "Banana", "blend_state", "Mashed")
    leaf("ChoppedCarrot", "Carrot", "chop_state", "Chopped")
    leaf("MashedCarrot", "Carrot", "blend_state", "Mashed")
    plates = {
        "TomatoSalad": ["ChoppedTomato"],
        "TomatoLettuceSalad": ["ChoppedTomato", "ChoppedLettuce"],
        "TomatoLettuceOnionSalad": ["ChoppedTomato", "ChoppedLettuce", "ChoppedOnion"],
        "CarrotBanana": ["ChoppedCarrot", "ChoppedBanana"],
        "MashedCarrotBanana": ["MashedCarrot", "MashedBanana"],
        "CucumberOnion": ["ChoppedCucumber", "ChoppedOnion"],
        "AppleWatermelon": ["ChoppedApple", "ChoppedWatermelon"],
    }
    for key, kids in plates.items():
        book[key + "Plate"] = RecipeNode(name="Plate", id_num=next(ids), contains=[book[k] for k in kids])
    for key in plates:
        book[key] = RecipeNode(name="Deliversquare", id_num=next(ids), contains=[book[key + "Plate"]])
    book["floor"] = RecipeNode(name="Floor", id_num=next(ids))
    book["no_recipe"] = RecipeNode(name="Deliversquare", id_num=next(ids), contains=[book["floor"]])
    return book, next(ids)


_BOOK, DEFAULT_NUM_GOALS = _default_book()
# same key order as the reference's RECIPES dict (recipe_drawer.py:109-118)
RECIPES = {name: Recipe(_BOOK[name], DEFAULT_NUM_GOALS) for name in
           ["TomatoSalad", "TomatoLettuceSalad", "CarrotBanana", "MashedCarrotBanana", "CucumberOnion",
            "AppleWatermelon", "TomatoLettuceOnionSalad", "no_recipe"]}
RECIPE_STORE = {}


def register_recipe(recipe, name):
    """recipe_drawer.register_recipe (:34-35): once anything is registered the store replaces
    the default book (cooking_env.py:100-105)."""
    RECIPE_STORE[name] = recipe


def active_book():
    return RECIPE_STORE if RECIPE_STORE else RECIPES
