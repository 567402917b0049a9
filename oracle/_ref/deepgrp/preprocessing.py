"""Empty stub: sequence.pyx imports deepgrp.preprocessing but never uses it."""
