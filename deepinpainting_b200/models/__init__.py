"""Host-side mirror of the reference's ``models`` package for the shift-layer path:
IPSRFunction, IPSR_model, InnerCos, InnerCos2 (same names, constructor arguments and methods)."""
from .IPSRFunction import IPSRFunction      # noqa: F401
from .IPSR_model import IPSR_model          # noqa: F401
from .InnerCos import InnerCos              # noqa: F401
from .InnerCos2 import InnerCos2            # noqa: F401
