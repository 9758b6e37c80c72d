"""TEST INFRASTRUCTURE ONLY -- CPU oracle of the reference sort path.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package.  The product (radix-sort_b200/, librsb200.so) never does.
"""
from .oracle import (  # noqa: F401
    DATASETS, DATASET_NAMES, KEY_TYPES, Oracle, Reference, build, have_reference, numpy_dtype,
)
