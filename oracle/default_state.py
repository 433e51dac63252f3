"""Constructor-time state_dict of the reference models, rebuilt from a (name, shape, dtype) table.  TEST
INFRASTRUCTURE ONLY (see oracle/__init__.py).

``hesic_b200.synth.synth_state_dict`` overwrites every learnable tensor with seeded values but keeps what the
reference's constructors fix: bounds, pedestals, the bottleneck's ``target`` and ``_matrices`` initial values, the
``MaskedConv2d`` masks and the empty CDF buffers.  Those are restated here so that the CPU arm of ``bench.py``
(``--impl reference`` / ``cpu_baseline``) can build the benchmark's weights WITHOUT instantiating a model class --
instantiating ``newnet1.HSIC`` from this repository would load libhesic_b200.so into a process that must not run it.
``tests/test_oracle.py`` checks every rule-governed tensor against the SHA-1 the unmodified reference produced
(``tests/golden/*.json: state_dict_init``).

Rules and where the reference states them:
  * ``*.lower_bound.bound`` / ``*.pedestal`` of a GDN reparametrisation: compressai/ops/parametrizers.py:30-39
    (pedestal = 2^-36, bound = sqrt(minimum + pedestal), minimum = 1e-6 for beta, 0 for gamma; layers/gdn.py:46-53)
  * ``likelihood_lower_bound.bound`` = 1e-9, ``lower_bound_scale.bound`` = ``scale_bound`` = 0.11:
    compressai/entropy_models/entropy_models.py:63-66,462-468
  * ``target`` = (-t, 0, t), t = log(2 / tail_mass - 1), tail_mass = 1e-9: entropy_models.py:295-296
  * ``_matrices.i`` = log(expm1(1 / scale / filters[i+1])), scale = init_scale^(1/(len(filters)+1)), init_scale = 10,
    filters = (1, 3, 3, 3, 3, 1): entropy_models.py:276-284
  * ``mask`` of MaskedConv2d type 'A': ones; zero at [kh//2, kw//2:] and [kh//2+1:, :]: compressai/layers/layers.py:37-40
"""
import json
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
_DT = {"float32": torch.float32, "int32": torch.int32, "int64": torch.int64, "float64": torch.float64, "uint8": torch.uint8,
       "bool": torch.bool}
PEDESTAL = 2.0 ** -36


def spec(name):
    """(key -> {'shape', 'dtype'}) table of one reference model, as dumped from the reference by make_golden.py."""
    return json.load(open(os.path.join(GOLDEN, name + ".json")))["state_dict_init"]


def rule_governed(key):
    """True when the reference's constructor fixes this tensor (it is then reproduced exactly below)."""
    leaf = key.rsplit(".", 1)[-1]
    return leaf in ("bound", "pedestal", "target", "mask", "scale_bound") or "_matrices" in key


def initial_state_dict(table):
    out = {}
    for key, ent in table.items():
        shape, dtype = tuple(ent["shape"]), _DT[ent["dtype"]]
        leaf = key.rsplit(".", 1)[-1]
        t = torch.zeros(shape, dtype=dtype)
        if leaf == "pedestal":
            t.fill_(PEDESTAL)
        elif leaf == "bound":
            if "beta_reparam" in key:
                t.fill_((1e-6 + PEDESTAL) ** 0.5)
            elif "gamma_reparam" in key:
                t.fill_(PEDESTAL ** 0.5)
            elif "likelihood_lower_bound" in key:
                t.fill_(1e-9)
            elif "lower_bound_scale" in key:
                t.fill_(0.11)
            else:
                raise KeyError(f"no rule for {key}")
        elif leaf == "scale_bound":
            t.fill_(0.11)
        elif leaf == "target":
            tm = float(np.log(2 / 1e-9 - 1))
            t.copy_(torch.tensor([-tm, 0.0, tm]))
        elif "_matrices" in key:
            filters = (1, 3, 3, 3, 3, 1)
            i = int(leaf)
            scale = 10.0 ** (1 / 5)
            t.fill_(float(np.log(np.expm1(1 / scale / filters[i + 1]))))
        elif leaf == "mask":
            t.fill_(1.0)
            kh, kw = shape[2], shape[3]
            t[:, :, kh // 2, kw // 2:] = 0
            t[:, :, kh // 2 + 1:] = 0
        out[key] = t
    return out
