"""Loads tests/golden/multibox_golden.npz (made from the reference itself by tests/golden/make_golden.py) and
re-creates the seeded inputs of the digest cases."""
import importlib.util
import json
import os

import numpy as np

_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load():
    z = np.load(os.path.join(_DIR, "multibox_golden.npz"))
    meta = json.loads(bytes(z["meta"]).decode())
    return z, meta


def generator():
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(_DIR, "make_golden.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod
