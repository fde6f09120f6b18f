"""Loader for tests/golden/*.json.gz (written by tests/golden/make_golden.py from the real reference)."""
import gzip
import json
import os

import numpy as np

_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    with gzip.open(os.path.join(_DIR, name + ".json.gz"), "rb") as fh:
        return json.loads(fh.read().decode())


def unhex(s, dtype):
    return np.frombuffer(bytes.fromhex(s), dtype=dtype).copy()
