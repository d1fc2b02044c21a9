import dataclasses as _dc


def dataclass(cls):
  return _dc.dataclass(cls)
