"""Stand-in for jax.lax (forward-only)."""


def stop_gradient(x):
  return x
