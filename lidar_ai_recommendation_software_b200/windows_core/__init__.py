"""Socket for the desktop shell: windows_implementation/core/project_manager.py:274-377."""
from .analysis import run_analysis, convert_numpy  # noqa: F401
