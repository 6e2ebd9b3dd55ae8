"""pyrenderer_b200 -- B200-native path-tracing core behind pyrenderer's Python API.

Host side (this package) mirrors the reference's objects: ``io_utils.read_tungsten.read_file``,
``core.scene.Scene``, ``core.camera.Camera``, ``core.tracing`` render entry points.
Everything underneath is hand-written CUDA for sm_100a in ``csrc/`` behind the
C ABI of ``include/prt.h`` (``libprt.so``, bound in ``_abi.py``).
"""
__version__ = "0.1.0"
