"""``cupy`` shim: lets the reference's example scripts (``import cupy; xp = cupy``)
run unchanged on the B200 backend.  Put ``melvin.py_b200/shims`` on PYTHONPATH only
when CuPy itself is not wanted; everything is re-exported from ``melvin.b200``."""
from melvin.b200 import *  # noqa: F401,F403
from melvin.b200 import (DeviceArray as ndarray, asarray, asnumpy, fft, max, mean, min,  # noqa: F401
                         sum, synchronize)
