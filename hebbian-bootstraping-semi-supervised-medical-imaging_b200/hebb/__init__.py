"""B200 (sm_100a) drop-in for the reference `hebb` package.

Put this package's parent directory ahead of the reference on sys.path and the reference's
scripts (`from hebb.makehebbian import makehebbian`, `from hebb import HebbianConv2d`, ...) pick
up the CUDA implementation unchanged.
"""
from .hebb import *      # noqa: F401,F403
from .hebb3d import *    # noqa: F401,F403
from ._native import set_default_precision as set_precision, get_default_precision, prec_name  # noqa: F401
