"""Host-side parameter holders with the reference's module names and state_dict keys, for the notebook's
`model.acm = BasicAcM(model.ob_dim * 2, model.ac_dim, False)` (notebooks/load_and_test.ipynb cell 24) and for building ACMs to
hand to EvalsWrapperACM-style code.  Assigning one to `model.acm` moves its parameters to the device path (the kind of ACM is
recognised from the state_dict keys); the arithmetic of the drop-in classes never runs through these modules.

Shapes: AcM = rltoolkit/basic_model.py:108-116 (2ob -> 64 -> 32 -> ac, tanh, output scaled by ac_lim);
BasicAcM = rltoolkit/acm/models/basic_acm.py:11-22 (100 / 50 hidden units, skip path fc21 gated by `t`, output gains `t1`)."""
import torch
from torch import nn


class AcM(nn.Module):
    def __init__(self, in_dim, ac_dim, ac_lim, discrete=False):
        super().__init__()
        if discrete:
            raise NotImplementedError("discrete-action ACMs are outside the device path")
        self.ac_lim, self.discrete = ac_lim, discrete
        self.fc1, self.fc2, self.fc3 = nn.Linear(in_dim, 64), nn.Linear(64, 32), nn.Linear(32, ac_dim)

    def forward(self, x):
        for layer in (self.fc1, self.fc2, self.fc3):
            x = torch.tanh(layer(x))
        return x * self.ac_lim

    act = forward


class BasicAcM(nn.Module):
    def __init__(self, in_dim, ac_dim, discrete=False):
        super().__init__()
        if discrete:
            raise NotImplementedError("discrete-action ACMs are outside the device path")
        self.discrete = discrete
        self.fc1, self.fc2 = nn.Linear(in_dim, 100), nn.Linear(100, 50)
        self.fc21, self.fc3 = nn.Linear(in_dim, 50), nn.Linear(50, ac_dim)
        self.t = nn.Parameter(torch.ones(1))
        self.t1 = nn.Parameter(torch.ones(ac_dim))

    def forward(self, x):
        skip = self.t * self.fc21(x)
        hidden = torch.tanh(self.fc2(torch.tanh(self.fc1(x))) + skip)
        return self.t1 * torch.tanh(self.fc3(hidden))

    act = forward


def acm_kind_of(state_dict, ob_dim, ac_dim):
    """'acm' | 'basic' from the keys and shapes of an ACM state_dict; raises on anything the device path has no kernel for."""
    keys = set(state_dict.keys())
    shape = lambda k: tuple(state_dict[k].shape)
    if {"t", "t1", "fc21.weight"} <= keys:
        kind, h1, h2 = "basic", 100, 50
    elif keys == {"fc1.weight", "fc1.bias", "fc2.weight", "fc2.bias", "fc3.weight", "fc3.bias"}:
        kind, h1, h2 = "acm", 64, 32
    else:
        raise TypeError("unsupported ACM module: state_dict keys %s match neither AcM nor BasicAcM" % sorted(keys))
    if shape("fc1.weight") != (h1, 2 * ob_dim) or shape("fc2.weight") != (h2, h1) or shape("fc3.weight") != (ac_dim, h2):
        raise TypeError("ACM shapes %s / %s / %s do not fit ob_dim %d, ac_dim %d with %d / %d hidden units"
                        % (shape("fc1.weight"), shape("fc2.weight"), shape("fc3.weight"), ob_dim, ac_dim, h1, h2))
    return kind
