"""Mean functions (mirror of gpx/mean_functions.py:25-30).  The device code knows these two
by identity (GPXB_MEAN_ZERO / GPXB_MEAN_DATA in include/gpxb200.h)."""
import numpy as np


def zero_mean(y):
    return np.float64(0.0)


def data_mean(y):
    return np.mean(np.asarray(y, dtype=np.float64))


def mean_mode(fn) -> int:
    if fn is zero_mean:
        return 0
    if fn is data_mean:
        return 1
    raise NotImplementedError(
        "gpx_b200 evaluates the mean on the device and supports zero_mean and data_mean only"
    )
