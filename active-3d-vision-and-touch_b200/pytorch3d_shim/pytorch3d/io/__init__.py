from .obj_io import load_obj, save_obj  # noqa: F401
