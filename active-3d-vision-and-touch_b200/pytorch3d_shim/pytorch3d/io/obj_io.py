"""pytorch3d.io.obj_io.{load_obj, save_obj} (utils.py:23)."""
from ptk_b200.obj_io import load_obj, save_obj  # noqa: F401
