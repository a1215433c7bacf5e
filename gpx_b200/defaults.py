"""Default arguments (mirror of gpx/defaults.py:20-36)."""
from collections import namedtuple

from .random import PRNGKey

GPX_DEFAULTS = {
    "num_evals": 5,
    "num_lanczos": 8,
    "key_lanczos": PRNGKey(2023),
    "key_precond": PRNGKey(2023),
}

gpxargs = namedtuple("GPX_DEFAULTS_ARGUMENTS", GPX_DEFAULTS.keys(), defaults=GPX_DEFAULTS.values())()

# jax_threefry_partitionable of the JAX the results are compared with (True since JAX 0.5.0)
THREEFRY_PARTITIONABLE = True
