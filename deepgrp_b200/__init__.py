"""deepgrp_b200 -- B200-native implementation of DeepGRP's prediction hot path.

Modules mirror the reference package on that path: ``sequence``, ``mss``, ``prediction``,
``model``, ``preprocessing`` (``Data`` only) and the ``deepgrp`` command line (``__main__``).
All computation runs in hand-written sm_100a CUDA kernels behind ``libdeepgrp_b200.so``
(C ABI in ``include/deepgrp_b200.h``); importing the package needs neither the library nor a GPU,
calling into it needs both.
"""
__version__ = "0.1.0"
