from .chamfer import chamfer_distance  # noqa: F401
