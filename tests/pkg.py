"""Imports the product package (its directory name contains hyphens, so importlib is needed)."""
import importlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

b200lc = importlib.import_module("gpu-lossless-compression_b200")
