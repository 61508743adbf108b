from .outputs import BaseOutput  # noqa: F401
