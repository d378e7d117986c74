"""`rl` of the reference with the hot path shadowed by the CUDA engine (see ../__init__.py)."""
from crowdnav_prediction_attngraph_b200.compat import extend_with_reference

extend_with_reference(__path__, "rl")
