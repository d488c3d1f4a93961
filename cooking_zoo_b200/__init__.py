"""cooking_zoo_b200 — B200-native batched simulator for CookingZoo's per-step hot path."""
from .batched import BatchedCookingEnv  # noqa: F401
from .mixed import MixedAgentCookingEnv  # noqa: F401
from .recipes import RECIPES, RecipeNode, Recipe, register_recipe  # noqa: F401
from .tables import compile_tables  # noqa: F401
