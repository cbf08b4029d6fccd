from . import segment_reduction  # noqa: F401
