"""Re-exports the reference's VENDORED Attention (architecture/attention_processor.py) so that the reference's own
container class is the one exercised. Its lazy `from .normalization import ...` is satisfied by registering the
stand-in normalization module as `architecture.normalization` (the reference tree ships no such file)."""
import sys

from . import normalization as _norm

sys.modules.setdefault("architecture.normalization", _norm)
from architecture.attention_processor import *  # noqa: F401,F403,E402
from architecture.attention_processor import Attention  # noqa: F401,E402
