"""slr_b200 -- B200-native (sm_100a) implementation of SLR's path-tracing hot path.

The product is two native libraries built in-tree under ``slr_b200/lib``:

* ``libslrgpu.so``  -- hand-written CUDA kernels behind the C ABI of ``include/slrgpu.h``
* ``libslrhost.so`` -- the host C++ side (scene graph, flattening, SBVH->QBVH build, scene language,
  renderer front end), C entry points in ``include/slrhost.h``

This Python package is only the FFI harness used by the tests and ``bench.py`` (ctypes bindings in
:mod:`slr_b200.capi`) plus synthetic scene/asset generators (:mod:`slr_b200.synth`). There is no
Python or CPU implementation of the hot path here: if the native libraries are missing, importing
:mod:`slr_b200.capi` raises.
"""

__version__ = "0.1.0"
