"""Case tables shared by the golden generator and the tests."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
from make_golden import BLEND_CASES, GIF_NAMES, JPEG_CASES, ORIENT_SRC, PNG_NAMES, RESIZE_CASES  # noqa: E402,F401
