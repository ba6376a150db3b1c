"""Loader for the committed reference fixtures (tests/golden/*.npz)."""
import glob
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
# (c2_full_enc_len.npz is a different kind of fixture: lengths / CRCs of the full C2 corpus, see make_c2_full.py)
GOLDEN = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(HERE, "golden", "*.npz"))
                if not os.path.basename(p).startswith("c2_full"))


def _split(data, off):
    b = data.tobytes()
    return [b[off[i]:off[i + 1]] for i in range(len(off) - 1)]


def load(name):
    z = np.load(os.path.join(HERE, "golden", name + ".npz"))
    return dict(keys=_split(z["keys"], z["key_off"]), vals=_split(z["vals"], z["val_off"]),
                enc=_split(z["enc"], z["enc_off"]), rc=z["rc"], pools=z["pools"], pool_used=z["pool_used"],
                packed=(z["keys"], z["key_off"], z["vals"], z["val_off"]))
