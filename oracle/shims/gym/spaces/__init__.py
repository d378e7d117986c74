from .box import Box
from .dict import Dict
