"""Stand-in for the five PyTorch3D names the reference imports (pterotactyl/utility/utils.py:20-23),
backed by libptk_b200.so.  Put `active-3d-vision-and-touch_b200/pytorch3d_shim` on sys.path (or call
ptk_b200.install_pytorch3d_shim()) and the reference's utils.py imports unedited."""
__version__ = "0.5.0+ptk_b200"
