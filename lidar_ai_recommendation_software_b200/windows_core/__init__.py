"""Desktop shell (windows_implementation/core): the analysis socket of project_manager.py:274-377
and the loader of data_loader.py:15-447 (§8 f1)."""
from .analysis import run_analysis, convert_numpy  # noqa: F401
from .data_loader import DataLoader, Dataset  # noqa: F401
