"""ORACLE / TEST INFRASTRUCTURE: minimal stand-in for OpenAI `baselines` (absent in this image).
Pure plumbing (no arithmetic).  Only what `rl.networks.envs`/`rl.networks.model` import."""
from . import bench, logger  # noqa: F401
