"""Host side of a device brain: hyper-parameters with the reference's names, the epsilon schedule descriptor,
state_dict round trips under the reference's attribute names (`eval_net`, `target_net`, `agent`, `target`, `model`),
and the binding to a vectorised Environment (device network + per-world replay rings)."""
import torch

from .. import _lib
from . import packing
from .utils import BasicBrain


class _NetHandle:
    """What `brain.eval_net` / `.agent` / `.model` is here: something with state_dict()/load_state_dict(), so that
    Agent.save_brain (World/entities.py:224-242) and Saver keep working.  Before the brain is bound to an
    Environment it holds the host state_dict; afterwards it reads/writes the device buffer."""

    def __init__(self, brain, target=False):
        self._brain, self._target = brain, target

    def state_dict(self):
        b = self._brain
        if b._dev is not None:
            return b._dev.state_dict(target=self._target)
        return b._host_sd_target if self._target else b._host_sd

    def load_state_dict(self, sd):
        b = self._brain
        sd = {k: torch.as_tensor(v).detach().clone().float() for k, v in sd.items()}
        if b._dev is not None:
            b._dev.load_state_dict(sd, target=self._target)
        elif self._target:
            b._host_sd_target = sd
        else:
            b._host_sd = sd

    def eval(self):
        return self

    def parameters(self):
        return list(self.state_dict().values())


def _copy_tensors(dst, src):
    """dst.<tensor attributes> <- src's, recursively through helper objects (ReplayRings / SumTrees / PpoData)."""
    if dst is None or src is None:
        return
    for name, val in src.__dict__.items():
        if isinstance(val, torch.Tensor):
            getattr(dst, name).copy_(val)
        elif hasattr(val, "__dict__") and not isinstance(val, (type, torch.device)) and hasattr(getattr(dst, name, None), "__dict__") \
                and type(val).__module__.startswith("reinlife_b200"):
            _copy_tensors(getattr(dst, name), val)


class DeviceBrainBase(BasicBrain):
    KIND = None          # packing.DUELING / DQN / PPO
    RULE = None          # _lib.ACT_*
    PRIORITIZED = False
    HAS_TARGET = True
    DEVICE_LEARN = False  # True once the brain's learn step exists as CUDA kernels

    def _init_common(self):
        self._dev = None            # brains.DeviceBrain once bound
        self._replay = None
        self._env = None
        self._plugin_host = None    # plugin.PluginHost: private one-world context of the per-agent learn() calls

    # ---- binding ---------------------------------------------------------------------------------
    def _bind(self, env, gene):
        """Called by Environment: move the network to env.device, allocate the per-world replay rings."""
        from ..brains import DeviceBrain, ReplayRings
        if self._env is not None and self._env is not env:
            raise RuntimeError("a brain object can be bound to one Environment only")
        if self._dev is None:
            self._dev = DeviceBrain(self.KIND, self._host_sd, env.device, lr=self._lr(), gamma=self._gamma(),
                                    batch=self._batch(), has_target=self.HAS_TARGET)
            if self.HAS_TARGET and self._host_sd_target is not None:
                self._dev.load_state_dict(self._host_sd_target, target=True)
        self._env, self._gene = env, gene
        if env.training and self._trains():
            if self._replay is None:
                # the dueling brains' rings hold float16 rows when their events run on fp16 tensor-core operands
                fp16 = bool(getattr(env, "_learn_fp16", False)) and self.KIND == packing.DUELING and not getattr(env, "_learn_single", False)
                need = ReplayRings.bytes_needed(env.n_worlds, self._capacity(), fp16=fp16)
                free, _ = torch.cuda.mem_get_info(env.device)
                if need > 0.9 * free:
                    raise MemoryError(f"replay rings for {env.n_worlds} worlds x capacity {self._capacity()} need "
                                      f"{need / 2**30:.1f} GiB, {free / 2**30:.1f} GiB free; lower `capacity`")
                self._replay = ReplayRings(env.n_worlds, self._capacity(), env.device, prioritized=self.PRIORITIZED, fp16=fp16)
            self._dev.alloc_learn(env.rows.row_cap)

    def _trains(self):
        return bool(getattr(self, "training", True)) and self.DEVICE_LEARN

    def _sched(self):
        return _lib.BrainSched(self.RULE, int(bool(getattr(self, "training", False))), 0.0, 1.0, 0)

    def _sync_host_scalars(self, eps, seen):
        """Device -> host copy of the schedule state the kernels own once the brain is bound (rl_brain_epsilon_update,
        rl_perdqn_epsilon_step), so that `brain.epsilon` / `brain.n_epi`, the per-agent plugin calls and the
        parameters_*.json written by the Saver (Helpers/saver.py:170-194) show the decayed values like the reference's."""
        if hasattr(self, "epsilon") and getattr(self, "training", True):
            self.epsilon = float(eps)
        if hasattr(self, "n_epi"):
            self.n_epi = int(seen)

    # ---- copy.deepcopy(brain) of the reference (World/environment.py:149,544) --------------------------------
    def clone(self):
        """Deep copy of this brain the way `copy.deepcopy` copies a reference brain: networks, optimizer state, replay
        memory, schedule scalars -- device-to-device tensor copies.  The copy is an independent plugin-mode brain."""
        import copy
        from ..brains import DeviceBrain
        new = copy.copy(self)                                   # hyper-parameters and schedule scalars (epsilon, n_epi, ...)
        new._dev = new._replay = new._env = new._plugin_host = None
        if hasattr(new, "memory"):
            new.memory = None
        src_sd = self._dev.state_dict() if self._dev is not None else self._host_sd
        new._host_sd = {k: torch.as_tensor(v).clone() for k, v in src_sd.items()}
        if self.HAS_TARGET:
            tgt = self._dev.state_dict(target=True) if self._dev is not None else self._host_sd_target
            new._host_sd_target = {k: torch.as_tensor(v).clone() for k, v in tgt.items()}
        for name, val in list(self.__dict__.items()):
            if isinstance(val, _NetHandle):
                setattr(new, name, _NetHandle(new, target=val._target))
        if self._dev is not None:
            d = self._dev
            new._dev = DeviceBrain(self.KIND, self._host_sd, d.device, lr=self._lr(), gamma=self._gamma(), batch=self._batch(),
                                   has_target=self.HAS_TARGET)
            for name in ("params", "target", "adam_m", "adam_v", "adam_step"):
                if getattr(d, name) is not None:
                    getattr(new._dev, name).copy_(getattr(d, name))
            new._dev.wimg_stale = True
        if self._plugin_host is not None:                       # the replay memory / data list / SumTree and the device schedule state
            from ..plugin import PluginHost
            new._plugin_host = PluginHost(new, device=self._plugin_host.env.device)
            _copy_tensors(new._replay, self._replay)
            if getattr(self, "memory", None) is not None:
                _copy_tensors(new.memory, self.memory)
            new._plugin_host.env._eps.copy_(self._plugin_host.env._eps)
            new._plugin_host.env._seen.copy_(self._plugin_host.env._seen)
            new._plugin_host.calls = self._plugin_host.calls
        return new

    # ---- plugin surface for single observations (the reference's per-agent calls) ------------------
    def _plugin_learn(self, **kw):
        """brain.learn(...) with host buffers (World/entities.py:194-208): see reinlife_b200/plugin.py."""
        if self._plugin_host is None:
            if self._env is not None:
                raise RuntimeError("this brain is bound to a vectorised Environment: its agents learn through "
                                   "Environment.learn(n_epi); per-agent learn() is for a brain used on its own")
            if not self._trains():
                return                      # a non-training brain: the reference would still fill its buffer; nothing reads it
            from ..plugin import PluginHost
            self._plugin_host = PluginHost(self)
        self._plugin_host.learn(**kw)

    def _q_single(self, state):
        """Network output for ONE observation through the same act kernel (1 row)."""
        from ..single import forward_single
        return forward_single(self, state)
