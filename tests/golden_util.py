"""Loading of tests/golden/*.npz (fixtures written by oracle/gen_golden.py from the reference)."""
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
NAMES = ["tiger", "ftiger", "ftiger_mu", "gridworld3", "ca", "sysadmin", "sysadmin3", "gridworld3_fba",
         "gridworld5"]
TABULAR = ["tiger", "gridworld3", "gridworld5"]
MUTATE_KIND = {"ftiger_mu": 0, "ca": 1, "sysadmin3": 2}


class Golden:
    def __init__(self, name):
        self.name = name
        self.z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
        self.desc = {k[len("model/"):]: self.z[k] for k in self.z.files if k.startswith("model/")}
        self.t_par = self.z["structs/t_par"]
        self.o_par = self.z["structs/o_par"]
        self.discount = float(self.z["meta/discount"])
        self.horizon = int(self.z["meta/horizon"])
        self.a = self.z["script/a"]
        self.o = self.z["script/o"]
        self.flags = self.z["script/flags"]

    def __getitem__(self, k):
        return self.z[k]

    def has(self, k):
        return k in self.z.files

    def steps(self, prefix, last=None):
        """Script steps that carry an update for the given section ('is', 'rs', 'reinv')."""
        if prefix == "is":  # the IS section replays the whole script (incl. trailing resets)
            return list(range(len(self.a)))
        last = int(self.z[prefix + "/last_step"]) if last is None else last
        return [t for t in range(last + 1)]


_cache = {}


def load(name):
    if name not in _cache:
        _cache[name] = Golden(name)
    return _cache[name]
