"""gpx_b200 -- B200-native (sm_100a) drop-in for the Gaussian-process hot path of
Molecolab-Pisa/GPX: the gpx.models GPR / SGPR / SGPR_RPChol fit / predict /
log_marginal_likelihood API over hand-written FP64 CUDA kernels behind a C ABI
(include/gpxb200.h, gpx_b200/csrc).  There is no CPU fallback."""
from . import bijectors, defaults, kernels, mean_functions, parameters, priors, random  # noqa: F401

__version__ = "0.1.0"
