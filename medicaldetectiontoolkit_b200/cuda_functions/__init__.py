"""Same dotted import paths as the reference's `cuda_functions` package (mrcnn.py:24-27, retina_unet.py:26-27), backed by libmdt_b200.so."""
