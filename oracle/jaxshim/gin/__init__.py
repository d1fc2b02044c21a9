"""Stand-in for gin-config: decorators are identities (configs.py:27-35)."""
REQUIRED = object()


def configurable(*args, **kwargs):
  if len(args) == 1 and callable(args[0]) and not kwargs:
    return args[0]
  return lambda f: f


class _Config:
  @staticmethod
  def external_configurable(fn, module=None, name=None):
    return fn


config = _Config()
