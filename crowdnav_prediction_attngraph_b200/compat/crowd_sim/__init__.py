"""crowd_sim/__init__.py:3-31 -> registers the reference's gym ids against the engine's single-env classes when a
`gym` module is installed (the vec-env path, make_vec_envs, does not need gym at all)."""
try:
    from gym.envs.registration import register
except Exception:            # gym is optional for the engine
    register = None

if register is not None:
    for _id, _cls in (("CrowdSimPred-v0", "CrowdSimPred"), ("CrowdSimVarNum-v0", "CrowdSimVarNum"),
                      ("CrowdSimPredRealGST-v0", "CrowdSimPredRealGST")):
        try:
            register(id=_id, entry_point="crowd_sim.envs:" + _cls)
        except Exception:    # already registered
            pass
