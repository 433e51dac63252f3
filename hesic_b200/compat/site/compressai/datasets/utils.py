"""``compressai.datasets`` of the drop-in tree: DELEGATES to the reference's own stereo ``ImageFolder``.

Data loading (left/right folders, same-crop, SURF + RANSAC ``get_H``, the 128x128 gray patches and corners for the
homography net; compressai/datasets/utils.py:30-214) is outside the forward hot path and
is not rewritten here.  But because this tree's ``compressai`` shadows the reference's, the drivers'
``from compressai.datasets import ImageFolder`` (ywz/mywork/test3real.py:31,328-335) lands in this file -- so it
loads the reference's file from a reference checkout and re-exports its names unchanged:

  * ``HESIC_REFERENCE_ROOT=/path/to/HESIC`` names the checkout, else
  * the first ``sys.path`` entry (or parent of one: the drivers run from ``ywz/mywork``) holding
    ``compressai/datasets/utils.py`` that is not this file.

Without a checkout ``ImageFolder(...)`` raises with that instruction (benchmarks and tests use ``hesic_b200.synth``).
"""
import importlib.util
import os
import sys

_HERE = os.path.abspath(__file__)
_REL = os.path.join("compressai", "datasets", "utils.py")


def _candidates():
    env = os.environ.get("HESIC_REFERENCE_ROOT")
    if env:
        yield env
    seen = set()
    for p in list(sys.path) + [os.getcwd()]:
        p = os.path.abspath(p or ".")
        for _ in range(4):                      # the entry itself and up to three parents (ywz/mywork/codec-test -> root)
            if p not in seen:
                seen.add(p)
                yield p
            p = os.path.dirname(p)


def reference_file():
    for root in _candidates():
        f = os.path.join(root, _REL)
        if os.path.isfile(f) and os.path.abspath(f) != _HERE:
            return f
    return None


_ref = None
_ref_path = reference_file()
if _ref_path is not None:
    _spec = importlib.util.spec_from_file_location("compressai.datasets._reference_utils", _ref_path)
    _ref = importlib.util.module_from_spec(_spec)
    sys.modules[_spec.name] = _ref
    _spec.loader.exec_module(_ref)          # imports cv2, kornia (real or compat/shims), torchvision
    for _k, _v in vars(_ref).items():
        if not _k.startswith("_"):
            globals()[_k] = _v
else:
    from torch.utils.data import Dataset

    class ImageFolder(Dataset):
        def __init__(self, root, transform=None, patch_size=(256, 256), split="train", need_file_name=False):
            raise RuntimeError(
                "compressai.datasets.ImageFolder is the reference's own data loader (compressai/datasets/utils.py:68-214); "
                "hesic_b200 delegates to it and found no reference checkout.  Set HESIC_REFERENCE_ROOT to the HESIC "
                "repository (or run the driver from inside it); synthetic pairs: hesic_b200.synth.stereo_pairs")
