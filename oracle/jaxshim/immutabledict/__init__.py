"""Stand-in for the `immutabledict` package (hashable read-only dict)."""


class immutabledict(dict):
  def __hash__(self):
    return hash(tuple(sorted((k, repr(v)) for k, v in self.items())))

  def _ro(self, *a, **k):
    raise TypeError('immutabledict is read-only')

  __setitem__ = __delitem__ = clear = pop = popitem = setdefault = update = _ro
