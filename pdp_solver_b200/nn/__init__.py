"""Mirror of the reference package `pdp.nn` for the accelerated path."""
from . import util, pdp_propagate, pdp_decimate, pdp_predict, solver  # noqa: F401
