"""Disk cache for slow ORACLE computations used by the GPU parity tests (test infrastructure only).

The oracle's Krylov path needs ~10 s per 48^3 timestep on one host core; GPU-box minutes are the scarce resource, so the tests
store the oracle's states under tests/_oracle_cache/ (git-ignored, travels with the snapshot like the built .so).  The key hashes
the oracle's source files and every parameter of the computation, so a stale entry can never be returned; without an entry the
test simply runs the oracle."""
import hashlib
import json
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CACHE = os.path.join(ROOT, "tests", "_oracle_cache")


def _oracle_digest():
    h = hashlib.sha256()
    d = os.path.join(ROOT, "oracle")
    for name in sorted(os.listdir(d)):
        if name.endswith(".py"):
            with open(os.path.join(d, name), "rb") as f:
                h.update(name.encode())
                h.update(f.read())
    return h.hexdigest()


def cached(tag, params, compute):
    """compute() -> dict of numpy arrays.  `params`: JSON-serialisable description of the computation."""
    key = hashlib.sha256((_oracle_digest() + json.dumps(params, sort_keys=True)).encode()).hexdigest()[:20]
    path = os.path.join(CACHE, "%s_%s.npz" % (tag, key))
    if os.path.exists(path):
        with np.load(path) as z:
            return {k: z[k] for k in z.files}
    out = compute()
    try:
        os.makedirs(CACHE, exist_ok=True)
        np.savez_compressed(path, **out)
    except OSError:
        pass
    return out
