"""hesic_b200 -- Blackwell-native (sm_100a) forward path of the HESIC stereo image codecs.

    import hesic_b200
    hesic_b200.install()          # make `compressai`, `newnet1`, ... resolve to this package
    from newnet1 import HSIC

The compute lives in hesic_b200/lib/libhesic_b200.so (C ABI: include/hesic_b200.h); importing
``hesic_b200._capi`` fails loudly when that library has not been built.
"""
from .compat import install  # noqa: F401

__version__ = "0.1.0"


def invalidate(model):
    """Force every packed-operand cache under ``model`` to re-pack at its next use.

    The caches key on ``(data_ptr, tensor version, device)`` of the parameters; mutation through ``.data``
    (``w.data.copy_()``, ``w.data.mul_()``, EMA weight swaps) does not bump a parameter's version counter and would
    otherwise keep running with the old packed weights.  ``load_state_dict`` / ``.to(device)`` / in-place ops on the
    parameter itself are detected without this call."""
    for mod in model.modules():
        d = mod.__dict__
        for name in ("_hesic_plan", "_hesic_en_plan", "_plan"):
            plan = d.get(name)
            if plan is not None and hasattr(plan, "invalidate"):
                plan.invalidate()
        if "_packed_key" in d:
            d["_packed_key"] = None
        for ent in (d.get("_plans") or {}).values() if isinstance(d.get("_plans"), dict) else ():
            if ent and hasattr(ent[0], "invalidate"):
                ent[0].invalidate()
        if isinstance(d.get("_plans"), dict):
            d["_plans"].clear()
        eng = d.get("_engine")
        if eng is not None and hasattr(eng, "_perm_plans"):
            eng._perm_plans.clear()
    return model
