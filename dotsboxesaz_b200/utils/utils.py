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


# ---------------------------------------------------------------- replay store
class ReplayStore:
    """The reference keeps its replay buffer in an HDF5 file with tables `fresh` (new self-play rows) and `data`
    (rows already split into training / validation) -- coach.py:57-67, utils/utils.py:94-96.  pytables is not always
    installed, so the same two tables can also live as parquet files next to each other: the backend is picked from
    the file name (.hdf/.h5 -> pandas HDFStore, anything else -> a directory of parquet parts).  Rows and columns are
    exactly the DataFrame of self_play.get_datasets()."""

    def __init__(self, path):
        self.path = path
        self.hdf = str(path).endswith((".hdf", ".h5", ".hdf5"))

    def _part_dir(self, key):
        import os
        d = os.path.join(self.path, key)
        os.makedirs(d, exist_ok=True)
        return d

    def append(self, key, df, part=None):
        """part: a writer-specific tag (e.g. the rank) so that several processes can append to the same table without
        agreeing on a part number; parts are read back in name order."""
        import os
        import time
        import pandas as pd
        if self.hdf:
            return write_to_hdf(self.path, key, df)
        d = self._part_dir(key)
        n = len([f for f in os.listdir(d) if f.endswith(".parquet")])
        name = "part%05d.parquet" % n if part is None else "part%016d_%s.parquet" % (time.time_ns() // 1000, part)
        df.reset_index().to_parquet(os.path.join(d, name))

    def has(self, key):
        import os
        import pandas as pd
        if self.hdf:
            with pd.HDFStore(self.path, mode="a") as store:
                return "/" + key in store
        d = os.path.join(self.path, key)
        return os.path.isdir(d) and any(f.endswith(".parquet") for f in os.listdir(d))

    def read(self, key, min_generation=None):
        import os
        import pandas as pd
        if self.hdf:
            where = None if min_generation is None else "generation>=%d" % min_generation
            with pd.HDFStore(self.path, mode="r") as store:
                return store.select(key, where)
        d = os.path.join(self.path, key)
        parts = [pd.read_parquet(os.path.join(d, f)) for f in sorted(os.listdir(d)) if f.endswith(".parquet")]
        df = pd.concat(parts, ignore_index=True)
        if min_generation is not None:
            df = df[df.generation >= min_generation]
        return df.set_index(["generation", "game_idx", "move_idx"])

    def delete(self, key):
        import os
        import shutil
        import pandas as pd
        if self.hdf:
            with pd.HDFStore(self.path, mode="a") as store:
                if "/" + key in store:
                    del store["/" + key]
            return
        shutil.rmtree(os.path.join(self.path, key), ignore_errors=True)


class ReplayDataset:
    """utils/utils.py:61-91 (HDFStoreDataset): rows of one split, optional `pos_average` (rows with identical
    features are merged by averaging pi and z), tensors as float32."""

    def __init__(self, df, train, features_shape=None, n_samples=int(1e12), pos_average=False, seed=None):
        import numpy as np
        df = df[df.training == (1 if train else -1)]
        df = df.sample(min(n_samples, df.shape[0]), random_state=seed)
        cols = list(df.columns)
        fcols = [c for c in cols if c.startswith("x_")]
        if pos_average and len(df):
            df = df.groupby(fcols).mean().reset_index()
        self.features = df[fcols].values.astype(np.float32)
        if features_shape:
            self.features = self.features.reshape(-1, *features_shape)
        self.policy = df[[c for c in cols if c.startswith("pi_")]].values.astype(np.float32)
        self.value = df.z.values.astype(np.float32)

    def __len__(self):
        return self.features.shape[0]

    def __getitem__(self, i):
        import numpy as np
        return self.features[i], self.policy[i], np.asarray([self.value[i]])
