"""flax.training.checkpoints stand-in: the fixture generators never save or restore."""


def save_checkpoint(*a, **k):
  raise NotImplementedError('jaxshim: checkpoints.save_checkpoint')


def restore_checkpoint(*a, **k):
  raise NotImplementedError('jaxshim: checkpoints.restore_checkpoint')
