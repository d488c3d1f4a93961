"""Stub of pettingzoo: AECEnv base + the few utils cooking_env.py imports.

`last`, `agent_iter`, `_clear_rewards`, `_accumulate_rewards` and `_was_dead_step` restate the published behaviour of
pettingzoo.utils.env.AECEnv (1.24): bookkeeping over the dicts the environment itself fills, no game arithmetic.  They
exist so that tests/golden/make_golden.py can record an AEC-surface trace (agent_selection / last() / per-agent step)
from the unmodified reference."""
from . import utils  # noqa: F401


class AECEnv:
    def __init__(self, *a, **k):
        pass

    @property
    def num_agents(self):
        return len(self.agents)

    @property
    def max_num_agents(self):
        return len(self.possible_agents)

    @property
    def unwrapped(self):
        return self

    def last(self, observe=True):
        agent = self.agent_selection
        observation = self.observe(agent) if observe else None
        return (observation, self._cumulative_rewards[agent], self.terminations[agent], self.truncations[agent],
                self.infos[agent])

    def agent_iter(self, max_iter=2 ** 63):
        it = 0
        while self.agents and it < max_iter:
            it += 1
            yield self.agent_selection

    def _clear_rewards(self):
        for agent in self.rewards:
            self.rewards[agent] = 0

    def _accumulate_rewards(self):
        for agent, reward in self.rewards.items():
            self._cumulative_rewards[agent] += reward

    def _deads_step_first(self):
        _deads_order = [agent for agent in self.agents if (self.terminations[agent] or self.truncations[agent])]
        if _deads_order:
            self._skip_agent_selection = self.agent_selection
            self.agent_selection = _deads_order[0]
        return self.agent_selection

    def _was_dead_step(self, action):
        if action is not None:
            raise ValueError("when an agent is dead, the only valid action is None")
        agent = self.agent_selection
        assert self.terminations[agent] or self.truncations[agent], "an agent that was not dead as attempted to be removed"
        del self.terminations[agent]
        del self.truncations[agent]
        del self.rewards[agent]
        del self._cumulative_rewards[agent]
        del self.infos[agent]
        self.agents.remove(agent)
        _deads_order = [agent for agent in self.agents if (self.terminations[agent] or self.truncations[agent])]
        if _deads_order:
            if getattr(self, "_skip_agent_selection", None) is None:
                self._skip_agent_selection = self.agent_selection
            self.agent_selection = _deads_order[0]
        else:
            if getattr(self, "_skip_agent_selection", None) is not None:
                self.agent_selection = self._skip_agent_selection
            self._skip_agent_selection = None
        self._clear_rewards()
