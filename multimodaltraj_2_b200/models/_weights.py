"""Shared helpers of the model mirrors: seed-0 N(0,1) initialisation (the reference uses
``tf.initializers.random_normal(mean=0, stddev=1, seed=0)``) and the ``sess_g`` weight lookup."""
from __future__ import annotations

import numpy as np
import torch


def init_normal(shape, seed=0, device="cuda"):
    rng = np.random.Generator(np.random.Philox(seed))
    return torch.from_numpy(rng.standard_normal(tuple(int(s) for s in shape)).astype(np.float32)).to(device)


def size0(in_features):
    """``int(in_features.shape[0])`` of the reference (the tensor is only used for its leading size)."""
    if isinstance(in_features, int):
        return in_features
    return int(in_features.shape[0])


def lookup(sess_g, name, shape, seed, device):
    """The reference pulls trained tensors out of the passed graph by name
    (``sess_g.get_tensor_by_name('krnl_weights_21/cost:0')``, models/g2k_lstm_mcr.py:38-76).  Here
    ``sess_g`` is ``None`` (seed-0 init) or a ``{name: array}`` dict, e.g. a decoded TF checkpoint;
    keys are matched on their last path component with and without the ``:0`` suffix."""
    if isinstance(sess_g, dict):
        for key, val in sess_g.items():
            base = key.split("/")[-1].split(":")[0]
            if base == name and tuple(np.shape(val)) == tuple(shape):
                return torch.as_tensor(np.asarray(val), dtype=torch.float32).contiguous().to(device)
    return init_normal(shape, seed, device)
