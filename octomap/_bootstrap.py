"""Locates the B200 package for the drop-in scripts (its name starts with a digit, so it is imported by string)."""
import importlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def package(sub=None):
    name = "3d_reconstruction_system_b200" + ("." + sub if sub else "")
    return importlib.import_module(name)
