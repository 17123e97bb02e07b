"""Stand-in for the reference package (tests/test_capture_tool.py only); see ../README.md."""
from . import model  # noqa: F401
