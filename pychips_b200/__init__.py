"""pychips_b200 — B200-native (sm_100a) implementation of the pyCHIPS coronal-hole
detection hot path behind the reference's own entry points (`Chips`, `SynopticChips`)."""
from .chips import Chips  # noqa: F401
from .syn_chips import SynopticChips  # noqa: F401

__version__ = "0.1.0"
