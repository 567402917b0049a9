"""deepgrp_b200.preprocessing -- only the ``Data`` tuple of the reference's module
(``deepgrp/preprocessing.py:72``), which is the argument type of ``predict_complete``.
Label preparation for training is out of scope."""
from typing import NamedTuple

import numpy as np

# Collection of forward one hot encoded sequence and true annotations
Data = NamedTuple("Data", [("fwd", np.ndarray), ("truelbl", np.ndarray)])
