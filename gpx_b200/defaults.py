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

# jax_threefry_partitionable of the JAX whose random streams are reproduced: True is JAX's default
# since 0.5.0; False is the layout of older JAX (the one the reference's notebooks were run with --
# tests/golden/notebook_anchors.json reproduces their printed numbers in this mode).
import os as _os

THREEFRY_PARTITIONABLE = _os.environ.get("GPXB_THREEFRY_PARTITIONABLE", "1") not in ("0", "false", "False")


def threefry_partitionable() -> bool:
    return bool(THREEFRY_PARTITIONABLE)


def set_threefry_partitionable(flag: bool) -> None:
    """Mirror of jax.config.update("jax_threefry_partitionable", flag)."""
    global THREEFRY_PARTITIONABLE
    THREEFRY_PARTITIONABLE = bool(flag)
