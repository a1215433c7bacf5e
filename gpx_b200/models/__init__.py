from .gpr import GPR  # noqa: F401
from .sgpr import SGPR  # noqa: F401
from .sgpr_rpchol import SGPR_RPChol  # noqa: F401
from .gpr_td import GPR_TD  # noqa: F401
