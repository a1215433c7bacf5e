from .gpr import GPR  # noqa: F401
