from . import p3d_gn  # noqa: F401
