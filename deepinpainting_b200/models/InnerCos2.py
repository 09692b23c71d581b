"""``InnerCos2`` -- as InnerCos, on the first 512 channels of the skip-concatenated activation
(models/InnerCos2.py:5-57: ``in_data.narrow(1, 0, 512)``); extra constructor argument ``infe``."""
from .InnerCos import InnerCos


class InnerCos2(InnerCos):
    def __init__(self, crit='MSE', strength=1, skip=0, infe=None):
        super(InnerCos2, self).__init__(crit=crit, strength=strength, skip=skip)
        self.inin = None
        self.infe = infe
        self._c_limit = 512
