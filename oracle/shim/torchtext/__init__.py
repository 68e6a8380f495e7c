"""TEST INFRASTRUCTURE ONLY - import-time stub of torchtext (start_end_dataset.py:10)."""
from . import vocab  # noqa: F401
