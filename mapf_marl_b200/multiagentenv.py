"""The pymarl environment contract, stated as data.

pymarl's runners talk to an environment through a fixed set of method names (the SMAC `MultiAgentEnv`
convention that the reference's MARL-curve-main/src/envs/multiagentenv.py spells out as stub methods).
Here the contract is a table: CONTRACT maps every required method to what it must return, and
`MultiAgentEnv.__init_subclass__` refuses a subclass that forgets one, so a drop-in env cannot silently
miss a call the rollout loop makes (episode_runner.py:70-115, parallel_runner.py:219-258).
"""

CONTRACT = {
    # name: (arguments, what the runner expects back)
    "reset": ((), "initial observations; the runner ignores them and calls the getters"),
    "step": (("actions",), "(reward, terminated, info)"),
    "get_obs": ((), "one observation per agent, [n_agents, obs_size]"),
    "get_obs_agent": (("agent_id",), "the observation of one agent"),
    "get_obs_size": ((), "int"),
    "get_state": ((), "global state vector, [state_size]"),
    "get_state_size": ((), "int"),
    "get_avail_actions": ((), "[n_agents, n_actions] of 0/1"),
    "get_avail_agent_actions": (("agent_id",), "[n_actions] of 0/1"),
    "get_total_actions": ((), "int, size of the discrete action space of an agent"),
    "close": ((), "None"),
}

# optional hooks the runners call when present; defaults are harmless no-ops
OPTIONAL = ("render", "seed", "save_replay", "get_stats")


class MultiAgentEnv:
    """Base class of the drop-in environments.  Subclasses must implement every method of CONTRACT and expose
    `n_agents` and `episode_limit` (read by get_env_info, which pymarl uses to build its replay scheme, run.py:125-148)."""

    def __init_subclass__(cls, **kwargs):
        super().__init_subclass__(**kwargs)
        missing = [name for name in CONTRACT if getattr(cls, name, None) is getattr(MultiAgentEnv, name, None)]
        if missing:
            raise TypeError("%s does not implement the MultiAgentEnv contract: %s" % (cls.__name__, ", ".join(missing)))

    def get_env_info(self):
        return dict(state_shape=self.get_state_size(), obs_shape=self.get_obs_size(),
                    n_actions=self.get_total_actions(), n_agents=self.n_agents, episode_limit=self.episode_limit)

    def render(self):
        return None

    def seed(self):
        return None

    def save_replay(self):
        return None

    def get_stats(self):
        return {}


def _unimplemented(name):
    def method(self, *args, **kwargs):
        raise NotImplementedError("%s.%s" % (type(self).__name__, name))
    method.__name__ = name
    method.__doc__ = "-> " + CONTRACT[name][1]
    return method


for _name in CONTRACT:
    setattr(MultiAgentEnv, _name, _unimplemented(_name))
del _name
