"""harmonic_b200: B200-native (sm_100a) drop-in for the hot path of astro-informatics/harmonic --
flow ``predict`` fused with the learned-harmonic-mean accumulation.  Mirrors the reference
namespace (harmonic/__init__.py:1-10)."""
from . import logs
from . import chains
from . import model
from . import evidence
from . import utils
from .chains import Chains
from .evidence import Evidence, Shifting

__version__ = "0.1.0"
