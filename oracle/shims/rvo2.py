"""ORACLE / TEST INFRASTRUCTURE: stand-in for the un-vendored `rvo2` Cython module.

Exposes `PyRVOSimulator` with the call surface the reference uses
(crowd_nav/policy/orca.py:80-114), backed by oracle/rvo2_ref.cpp through ctypes.
Python doubles are narrowed to C float at this boundary exactly as the Cython wrapper
does.  Put `oracle/shims` on sys.path to let the unmodified reference import it.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "..", "_build", "librvo2_ref.so")


def _load():
    if not os.path.exists(_LIB_PATH):
        raise ImportError(
            "oracle/_build/librvo2_ref.so missing: run `make -C oracle` (or __graft_entry__.build())")
    lib = ctypes.CDLL(_LIB_PATH)
    f, i, p = ctypes.c_float, ctypes.c_int, ctypes.c_void_p
    lib.rvo2ref_create.restype = p
    lib.rvo2ref_create.argtypes = [f, f, i, f, f, f, f]
    lib.rvo2ref_destroy.argtypes = [p]
    lib.rvo2ref_add_agent.restype = i
    lib.rvo2ref_add_agent.argtypes = [p, f, f, f, i, f, f, f, f, f, f]
    lib.rvo2ref_num_agents.restype = i
    lib.rvo2ref_num_agents.argtypes = [p]
    for name in ("set_position", "set_velocity", "set_pref_velocity"):
        getattr(lib, "rvo2ref_" + name).argtypes = [p, i, f, f]
    for name in ("get_velocity", "get_position"):
        getattr(lib, "rvo2ref_" + name).argtypes = [p, i, ctypes.POINTER(f)]
    lib.rvo2ref_do_step.argtypes = [p]
    lib.rvo2ref_do_step_only.argtypes = [p, i]
    lib.rvo2ref_num_lines.restype = i
    lib.rvo2ref_num_lines.argtypes = [p, i]
    lib.rvo2ref_line_fail.restype = i
    lib.rvo2ref_line_fail.argtypes = [p, i]
    lib.rvo2ref_neighbor_ids.restype = i
    lib.rvo2ref_neighbor_ids.argtypes = [p, i, ctypes.POINTER(i), i]
    lib.rvo2ref_get_line.argtypes = [p, i, i, ctypes.POINTER(f)]
    lib.rvo2ref_solve_one.argtypes = [ctypes.POINTER(f), ctypes.POINTER(f), i, f, f, f, f,
                                      ctypes.POINTER(f), ctypes.POINTER(i)]
    return lib


_lib = _load()

# When True, doStep() only solves agent 0 (the only velocity the reference reads back,
# crowd_nav/policy/orca.py:114).  Results for agent 0 are identical; used to keep tests fast.
ONLY_AGENT0 = False


class PyRVOSimulator(object):
    def __init__(self, timeStep, neighborDist, maxNeighbors, timeHorizon, timeHorizonObst,
                 radius, maxSpeed, velocity=(0.0, 0.0)):
        self._h = _lib.rvo2ref_create(timeStep, neighborDist, int(maxNeighbors), timeHorizon,
                                      timeHorizonObst, radius, maxSpeed)
        self._buf = (ctypes.c_float * 4)()

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h and _lib is not None:   # _lib is None during interpreter shutdown
            _lib.rvo2ref_destroy(h)

    def addAgent(self, pos, neighborDist, maxNeighbors, timeHorizon, timeHorizonObst, radius,
                 maxSpeed, velocity=(0.0, 0.0)):
        return _lib.rvo2ref_add_agent(self._h, pos[0], pos[1], neighborDist, int(maxNeighbors),
                                      timeHorizon, timeHorizonObst, radius, maxSpeed,
                                      velocity[0], velocity[1])

    def getNumAgents(self):
        return _lib.rvo2ref_num_agents(self._h)

    def setAgentPosition(self, i, pos):
        _lib.rvo2ref_set_position(self._h, i, pos[0], pos[1])

    def setAgentVelocity(self, i, vel):
        _lib.rvo2ref_set_velocity(self._h, i, vel[0], vel[1])

    def setAgentPrefVelocity(self, i, vel):
        _lib.rvo2ref_set_pref_velocity(self._h, i, vel[0], vel[1])

    def getAgentVelocity(self, i):
        _lib.rvo2ref_get_velocity(self._h, i, self._buf)
        return (self._buf[0], self._buf[1])

    def getAgentPosition(self, i):
        _lib.rvo2ref_get_position(self._h, i, self._buf)
        return (self._buf[0], self._buf[1])

    def doStep(self):
        if ONLY_AGENT0:
            _lib.rvo2ref_do_step_only(self._h, 0)
        else:
            _lib.rvo2ref_do_step(self._h)

    # diagnostics (not part of the Python-RVO2 surface)
    def _numLines(self, i):
        return _lib.rvo2ref_num_lines(self._h, i)

    def _lineFail(self, i):
        return _lib.rvo2ref_line_fail(self._h, i)

    def _neighborIds(self, i, cap=256):
        arr = (ctypes.c_int * cap)()
        n = _lib.rvo2ref_neighbor_ids(self._h, i, arr, cap)
        return [arr[k] for k in range(min(n, cap))]

    def _line(self, i, k):
        _lib.rvo2ref_get_line(self._h, i, k, self._buf)
        return tuple(self._buf[j] for j in range(4))


def solve_one(ego, others, neighbor_dist, time_horizon, time_step, other_max_speed=1.0):
    """One ego agent vs others with a fresh simulator; float32 numpy arrays in, (vx, vy, nlines, fail) out."""
    import numpy as np
    ego = np.ascontiguousarray(ego, dtype=np.float32)
    others = np.ascontiguousarray(others, dtype=np.float32).reshape(-1, 5)
    out = (ctypes.c_float * 2)()
    diag = (ctypes.c_int * 2)()
    fp = ctypes.POINTER(ctypes.c_float)
    _lib.rvo2ref_solve_one(ego.ctypes.data_as(fp), others.ctypes.data_as(fp), others.shape[0],
                           neighbor_dist, time_horizon, time_step, other_max_speed, out, diag)
    return out[0], out[1], diag[0], diag[1]
