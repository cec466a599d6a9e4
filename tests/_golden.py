"""Loader for the committed golden vectors (tests/golden/*.npz, made by make_golden.py)."""
import glob
import json
import os

import numpy as np
import torch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def names(prefix=''):
    out = [os.path.basename(p)[:-4] for p in sorted(glob.glob(os.path.join(GOLDEN_DIR, prefix + '*.npz')))]
    return out


def load(name):
    """-> (meta dict, arrays dict of torch tensors, state dict of torch tensors)."""
    with np.load(os.path.join(GOLDEN_DIR, name + '.npz')) as f:
        meta = json.loads(bytes(f['meta']).decode())
        arrays, sd = {}, {}
        for k in f.files:
            if k == 'meta':
                continue
            t = torch.from_numpy(np.array(f[k]))
            if k.startswith('sd/'):
                sd[k[3:]] = t
            else:
                arrays[k] = t
    return meta, arrays, sd
