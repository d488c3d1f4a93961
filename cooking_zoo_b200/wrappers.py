"""The reference's env wrappers re-exposed over the batched CUDA backend.

Same constructor arguments and return shapes as
  cooking_env.parallel_env            (environment/cooking_env.py:26-46; dicts keyed "player_i"),
  environment.GymCookingEnvironment   (environment/environment.py:10-31; id cookingEnv-v1),
  multi_agent_gym.GymCookingEnvironment (environment/multi_agent_gym.py:10-34; id cookingEnvMA-v1),
backed by a BatchedCookingEnv with num_envs=1 (each call copies one small result to the host; the
batched entry point is the fast path).  gymnasium / pettingzoo are optional: when they are
importable the spaces are real gymnasium spaces and the ids get registered, otherwise light
stand-ins with the same attributes are used.
"""
from dataclasses import dataclass

import numpy as np

from .batched import BatchedCookingEnv
from .recipes import active_book

try:  # optional
    import gymnasium as _gym
except Exception:  # pragma: no cover - not installed in the build image
    _gym = None


@dataclass
class _Box:
    low: float
    high: float
    shape: tuple
    dtype: type = np.float32


@dataclass
class _Discrete:
    n: int


def _box(low, high, shape):
    return _gym.spaces.Box(low=low, high=high, shape=shape) if _gym else _Box(low, high, shape)


def _discrete(n):
    return _gym.spaces.Discrete(n) if _gym else _Discrete(n)


class ParallelCookingEnv:
    """PettingZoo-parallel-style env: reset() -> (obs, infos); step({agent: action}) ->
    (obs, rewards, terminations, truncations, infos), all dicts keyed by "player_i"."""

    metadata = {"render_modes": [], "name": "cookingzoo_v1", "is_parallelizable": True}

    def __init__(self, level, meta_file, num_agents, max_steps, recipes, agent_visualization=None, obs_spaces=None,
                 end_condition_all_dishes=False, action_scheme="scheme1", render=False, reward_scheme=None,
                 agent_respawn_rate=0.0, grace_period=20, agent_despawn_rate=0.0, **backend_kwargs):
        self._b = BatchedCookingEnv(1, level, meta_file, num_agents, max_steps, recipes, agent_visualization,
                                    obs_spaces, end_condition_all_dishes, action_scheme, render, reward_scheme,
                                    agent_respawn_rate, grace_period, agent_despawn_rate, **backend_kwargs)
        self.possible_agents = list(self._b.possible_agents)
        self.agents = self.possible_agents[:]
        self.recipe_names = list(recipes)
        self.max_steps = max_steps
        self._episode = 0
        self._done = True
        book = list(active_book().keys())
        # quirk C-7 (cooking_env.py:155-161): agent i's goal vector is one-hot(i) over the book
        self.goal_vectors = {a: np.eye(len(book))[i] for i, a in enumerate(self.possible_agents)}
        L = self._b.obs_len
        self.observation_spaces = {a: _box(-1, 1, (L,)) for a in self.possible_agents}
        self.action_spaces = {a: _discrete(self._b.tables.num_actions) for a in self.possible_agents}  # cooking_env.py:131

    def observation_space(self, agent):
        return self.observation_spaces[agent]

    def action_space(self, agent):
        return self.action_spaces[agent]

    @property
    def unwrapped(self):
        return self

    @property
    def backend(self):
        return self._b

    def reset(self, seed=None, options=None):
        """The reference re-samples the layout from the global `random` stream and ignores `seed`
        (cooking_env.py:178); here episode k takes pool layout cz_layout_draw(seed, 0, k) % P, or
        options["layout_id"]."""
        options = options or {}
        if seed is not None:
            self._b.seed = int(seed)
        lid = options.get("layout_id")
        if lid is None:
            lid = self._b.lib.cz_layout_draw(self._b.seed, self._b.env_offset, self._episode) % self._b.tables.num_layouts
        self._episode += 1
        obs = self._b.reset(layout_ids=np.array([lid], np.int32)).cpu().numpy()[0]
        self.agents = self.possible_agents[:]
        self._done = False
        self._termination_info = ""
        self._was_active = np.ones(len(self.possible_agents), bool)
        return ({a: obs[i].copy() for i, a in enumerate(self.possible_agents)},
                {a: {} for a in self.possible_agents})

    def step(self, actions):
        """accumulated_step through the dict surface (cooking_env.py:243-269).  Only agents that are active or
        whose status changed this step appear in the returned dicts, and `self.agents` follows them: a despawned
        agent leaves (truncated=True on its despawn step, :333-350) and comes back when it respawns, while the
        episode goes on for the others; the episode ends when the recipes are complete or max_steps is reached."""
        if self._done:
            raise RuntimeError("step() called on a finished episode: call reset() first")
        A = len(self.possible_agents)
        act = np.zeros((1, A), np.uint8)
        for i, a in enumerate(self.possible_agents):
            act[0, i] = int(actions.get(a, 0))
        obs, rew, term, trunc, info = self._b.step(act)
        obs, rew = obs.cpu().numpy()[0], rew.cpu().numpy()[0]
        term, trunc = term.cpu().numpy()[0].astype(bool), trunc.cpu().numpy()[0].astype(bool)
        full = self._b.info()
        t = int(full["t"][0])
        over = bool(full["done"][0])
        active = full["active"][0].cpu().numpy().astype(bool)
        done = full["recipe_done"][0].cpu().numpy().astype(bool)
        # compute_truncated sets the message only when the clock runs out (cooking_env.py:334-335); it then stays set
        if t >= self.max_steps:
            self._termination_info = f"Terminating because {self.max_steps} timesteps passed"
        relevant = active | trunc          # active, or status changed this step (despawned now / truncated by the clock)
        out_obs, out_r, out_te, out_tr, out_i = {}, {}, {}, {}, {}
        for i, a in enumerate(self.possible_agents):
            if not relevant[i]:
                continue
            out_obs[a] = obs[i].copy()
            out_r[a] = np.float64(rew[i])
            out_te[a] = bool(term[i])
            out_tr[a] = bool(trunc[i])
            out_i[a] = {"goal_vector": self.goal_vectors[a], "t": t, "termination_info": self._termination_info,
                        "recipe_done": bool(done[i]), "action": int(act[0, i]) if self._was_active[i] else 0,
                        "task": self.recipe_names[i]}
        self._was_active = active.copy()
        self.agents = [a for i, a in enumerate(self.possible_agents) if relevant[i]]     # cooking_env.py:264-266
        if over:
            self._done = True
            self.agents = []
        return out_obs, out_r, out_te, out_tr, out_i

    def render(self, **kwargs):
        raise NotImplementedError("rendering is out of scope")

    def close(self):
        self._b.close()


def parallel_env(**kwargs):
    """cooking_env.parallel_env (cooking_env.py:46)."""
    return ParallelCookingEnv(**kwargs)


class GymCookingEnvironment:
    """environment.GymCookingEnvironment (environment/environment.py:5-31): single agent."""

    metadata = {"render.modes": [], "name": "cooking_zoo"}

    def __init__(self, level, meta_file, max_steps, recipes, agent_visualization=None, obs_spaces=None,
                 end_condition_all_dishes=False, action_scheme="scheme1", render=False, reward_scheme=None,
                 **backend_kwargs):
        self.zoo_env = parallel_env(level=level, meta_file=meta_file, num_agents=1, max_steps=max_steps,
                                    recipes=recipes, agent_visualization=agent_visualization, obs_spaces=obs_spaces,
                                    end_condition_all_dishes=end_condition_all_dishes, action_scheme=action_scheme,
                                    render=render, reward_scheme=reward_scheme, **backend_kwargs)
        self.observation_space = self.zoo_env.observation_space("player_0")
        self.action_space = self.zoo_env.action_space("player_0")

    def step(self, action):
        obs, reward, termination, truncation, info = self.zoo_env.step({"player_0": action})
        return obs["player_0"], reward["player_0"], termination["player_0"], truncation["player_0"], info["player_0"]

    def reset(self, **kwargs):
        obs, info = self.zoo_env.reset(**{k: v for k, v in kwargs.items() if k in ("seed", "options")})
        return obs["player_0"], info["player_0"]

    def close(self):
        self.zoo_env.close()


class GymCookingEnvironmentMA:
    """multi_agent_gym.GymCookingEnvironment (environment/multi_agent_gym.py:5-34): lists indexed by agent."""

    metadata = {"render.modes": [], "name": "multi_agent_cooking_zoo"}

    def __init__(self, level, meta_file, num_agents, max_steps, recipes, agent_visualization=None, obs_spaces=None,
                 end_condition_all_dishes=False, action_scheme="scheme1", render=False, reward_scheme=None,
                 **backend_kwargs):
        self.zoo_env = parallel_env(level=level, meta_file=meta_file, num_agents=num_agents, max_steps=max_steps,
                                    recipes=recipes, agent_visualization=agent_visualization, obs_spaces=obs_spaces,
                                    end_condition_all_dishes=end_condition_all_dishes, action_scheme=action_scheme,
                                    render=render, reward_scheme=reward_scheme, **backend_kwargs)
        self.num_agents = num_agents
        self.observation_space = self.zoo_env.observation_space("player_0")
        self.action_space = self.zoo_env.action_space("player_0")

    def step(self, actions):
        n = self.num_agents
        obs, reward, termination, truncation, info = self.zoo_env.step({f"player_{i}": actions[i] for i in range(n)})
        return ([obs[f"player_{i}"] for i in range(n)], [reward[f"player_{i}"] for i in range(n)],
                [termination[f"player_{i}"] for i in range(n)], [truncation[f"player_{i}"] for i in range(n)],
                [info[f"player_{i}"] for i in range(n)])

    def reset(self, **kwargs):
        obs, info = self.zoo_env.reset(**{k: v for k, v in kwargs.items() if k in ("seed", "options")})
        n = self.num_agents
        return [obs[f"player_{i}"] for i in range(n)], [info[f"player_{i}"] for i in range(n)]

    def close(self):
        self.zoo_env.close()


if _gym is not None:  # pragma: no cover - ids of cooking_zoo/__init__.py:3-8
    for _id, _ep in (("cookingEnv-v1", "cooking_zoo_b200.wrappers:GymCookingEnvironment"),
                     ("cookingEnvMA-v1", "cooking_zoo_b200.wrappers:GymCookingEnvironmentMA")):
        try:
            _gym.envs.registration.register(id=_id, entry_point=_ep)
        except Exception:
            pass
