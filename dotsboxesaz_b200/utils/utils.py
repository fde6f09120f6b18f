"""Config containers and Elo arithmetic (reference: utils/utils.py:8-58,103-132)."""
import math


class DotDict(dict):
    """Nested dict with attribute access; missing attributes read as None (utils/utils.py:8-48)."""

    def __init__(self, mapping=()):
        super().__init__()
        for k, v in dict(mapping).items():
            self[k] = v

    def __setitem__(self, key, value):
        if isinstance(value, dict) and not isinstance(value, DotDict):
            value = DotDict(value)
        super().__setitem__(key, value)

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return self.get(name, None)

    def __setattr__(self, name, value):
        self[name] = value

    def __delattr__(self, name):
        del self[name]

    def merge(self, other):
        for k, v in other.items():
            if k in self and isinstance(self[k], dict) and isinstance(v, dict):
                self[k].merge(v)
            else:
                self[k] = v

    def rewrite_str(self, tag, replacement):
        for k, v in self.items():
            if isinstance(v, str):
                self[k] = v.replace(tag, replacement)
            elif isinstance(v, DotDict):
                v.rewrite_str(tag, replacement)


class DictWithDefault(dict):
    """dict whose missing keys are built by `factory(key)` and stored (utils/utils.py:51-58)."""

    def __init__(self, factory):
        super().__init__()
        self.factory = factory

    def __missing__(self, key):
        value = self[key] = self.factory(key)
        return value


def _expected(r_a, r_b):
    return 1.0 / (1 + 1.0 * math.pow(10, (r_a - r_b) / 400))


def elo_rating2(elo0, elo1, n0, n1, K=30):
    """Batch Elo update after n0 wins of player 0 and n1 wins of player 1 (utils/utils.py:120-132)."""
    p1 = _expected(elo0, elo1)
    p0 = 1 - p1
    return elo0 + K * (n0 * p1 - n1 * p0), elo1 + K * (n1 * p0 - n0 * p1)


def write_to_hdf(hdf_file, key, dataframe):
    """utils/utils.py:94-96: append the sample rows to table `key` of the HDF store (needs pytables)."""
    import pandas as pd
    with pd.HDFStore(hdf_file, mode="a") as store:
        store.append(key, dataframe, format="table")
