from collections import OrderedDict

from .. import logging  # noqa: F401


class BaseOutput(OrderedDict):
    def __init__(self, *a, **k):
        super().__init__(*a, **k)
        for kk, v in self.items():
            object.__setattr__(self, kk, v)
