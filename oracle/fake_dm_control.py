"""ORACLE (test infrastructure, not product code) — stand-ins for the python packages the reference imports
but this container lacks (`gym`, `dm_control`), so that the reference's UNMODIFIED RobotEnv / Actuator /
Reward / RGBDSensor code can be executed here on top of oracle/engine.c.

install() registers fake `gym`, `gym.spaces`, `gym.utils.seeding`, `dm_control`, `dm_control.mujoco`,
`dm_control.mujoco.wrapper.mjbindings` in sys.modules.  `Physics` exposes exactly the 16 uses listed in
SURVEY.md §8b-lower (robot_env.py / actuator.py / sensor.py call sites), with numpy VIEWS into the engine's
memory so the reference's aliasing behaviour (robot_env.py:98 `current_qpos` is a live view) is preserved.

Used by tools/gen_golden.py (run in the build container, where /root/reference exists) to produce the
fixtures under tests/golden/.  Nothing on the GPU box imports this.
"""
import sys
import types

import numpy as np

from . import engine, mjcf


class _Box:
    def __init__(self, low, high, shape=None, dtype=np.float32):
        self.low, self.high, self.shape, self.dtype = low, high, tuple(shape), np.dtype(dtype)

    def sample(self):
        if self.dtype == np.uint8:
            return np.random.randint(0, 256, self.shape).astype(np.uint8)
        return np.random.uniform(-1, 1, self.shape).astype(self.dtype)


class _Dict(dict):
    def __init__(self, spaces):
        super().__init__(spaces)
        self.spaces = spaces


class _Named:
    """physics.named.data.xpos["ee"] / xfrc_applied["ee", 2] = v / named.model.geom_bodyid["object"]."""

    def __init__(self, arr, names):
        self._a, self._n = arr, names

    def _idx(self, key):
        if isinstance(key, tuple):
            return (self._n.index(key[0]),) + tuple(key[1:])
        return self._n.index(key)

    def __getitem__(self, key):
        return self._a[self._idx(key)]

    def __setitem__(self, key, v):
        self._a[self._idx(key)] = v


class _Contact:
    def __init__(self, c):
        self.geom1, self.geom2, self.dist = c.geom1, c.geom2, c.dist


class _ContactList:
    def __init__(self, data):
        self._d = data

    def __getitem__(self, i):
        return _Contact(self._d.p.contents.contact[i])


class Physics:
    render_fn = None  # optional callable(physics, camera_id, width, height, depth) -> image
    step_count = 0

    def __init__(self, md):
        self._md = md
        self._m = engine.Model(md)
        self._d = engine.Data(self._m)
        d, m, md_ = self._d, self._m, md
        outer = self

        class _DataNS(types.SimpleNamespace):
            @property
            def ncon(self_inner):
                return outer._d.ncon
        self.data = _DataNS(qpos=d.qpos, qvel=d.qvel, ctrl=d.ctrl, ptr=d, contact=_ContactList(d))
        self.model = types.SimpleNamespace(
            nv=m.nv, ptr=m, geom_bodyid=np.array(md_["geom_bodyid"]),
            opt=types.SimpleNamespace(gravity=np.array(md_["gravity"])),
            name2id=lambda name, kind: {"body": md_["body_names"], "geom": md_["geom_names"]}[kind].index(name))
        self.named = types.SimpleNamespace(
            data=types.SimpleNamespace(xpos=_Named(d.xpos, md_["body_names"]), xquat=_Named(d.xquat, md_["body_names"]),
                                       xfrc_applied=_XfrcNamed(d.xfrc_applied, md_["body_names"])),
            model=types.SimpleNamespace(geom_bodyid=_Named(np.array(md_["geom_bodyid"]), md_["geom_names"])))
        self.reset()

    @classmethod
    def from_xml_path(cls, path):
        return cls(mjcf.compile_mjcf(path))

    def reset(self):
        self._d.reset()

    def step(self):
        self._d.step()
        Physics.step_count += 1

    def render(self, camera_id=None, width=64, height=64, depth=False):
        if Physics.render_fn is not None:
            return Physics.render_fn(self, camera_id, width, height, depth)
        # deterministic state-independent stand-in images (the renderer has its own tests)
        yy, xx = np.mgrid[0:height, 0:width]
        if depth:
            return (0.2 + 2.5 * (yy + 1) / height + 0.01 * xx).astype(np.float32)
        return np.stack([(xx * 4) % 256, (yy * 4) % 256, (xx + yy) % 256], -1).astype(np.uint8)


class _XfrcNamed(_Named):
    """mjData.xfrc_applied rows are [force(3), torque(3)]; engine.c uses the same layout."""


def _mj_jacBody(model_ptr, data_ptr, jacp, jacr, body):
    jp, jr = data_ptr.jac_body(body)
    if jacp is not None:
        jacp[...] = jp
    if jacr is not None:
        jacr[...] = jr


def install():
    gym = types.ModuleType("gym")

    class GoalEnv:
        pass
    gym.GoalEnv = GoalEnv
    gym.Env = GoalEnv
    gym.register = lambda *a, **k: None
    spaces = types.ModuleType("gym.spaces")
    spaces.Box, spaces.Dict = _Box, _Dict
    gym.spaces = spaces
    utils = types.ModuleType("gym.utils")
    seeding = types.ModuleType("gym.utils.seeding")
    seeding.np_random = lambda seed=None: (np.random.RandomState(seed), seed)
    utils.seeding = seeding
    gym.utils = utils
    envs = types.ModuleType("gym.envs")
    registration = types.ModuleType("gym.envs.registration")
    registration.register = lambda *a, **k: None
    envs.registration = registration
    gym.envs = envs
    gym.__path__ = []
    dmc = types.ModuleType("dm_control")
    mj = types.ModuleType("dm_control.mujoco")
    mj.Physics = Physics
    wrapper = types.ModuleType("dm_control.mujoco.wrapper")
    mjb = types.ModuleType("dm_control.mujoco.wrapper.mjbindings")
    mjb.mjlib = types.SimpleNamespace(mj_jacBody=_mj_jacBody)
    dmc.mujoco, mj.wrapper, wrapper.mjbindings = mj, wrapper, mjb
    for name, mod in {"gym": gym, "gym.envs": envs, "gym.envs.registration": registration, "gym.spaces": spaces, "gym.utils": utils, "gym.utils.seeding": seeding,
                      "dm_control": dmc, "dm_control.mujoco": mj, "dm_control.mujoco.wrapper": wrapper,
                      "dm_control.mujoco.wrapper.mjbindings": mjb}.items():
        sys.modules[name] = mod
