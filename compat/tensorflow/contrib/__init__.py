"""tf.contrib: only `distributions` (examples/*/main.py:5-6 of the reference)."""
from . import distributions  # noqa: F401
