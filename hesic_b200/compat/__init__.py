"""Drop-in module tree for the reference's callers.

``install()`` puts ``compat/site`` at the FRONT of ``sys.path`` so that the
unmodified drivers (``ywz/mywork/test3real.py``, ``codec-test/test2_codec.py``)
resolve ``compressai``, ``newnet1``, ``newnet1_joint``, ``newnet9`` and ``model``
to this package, and ``compat/shims`` at the BACK so that third-party modules
the reference imports but does not vendor (``kornia``, ``range_coder``,
``pytorch_msssim``, ``imageio``) resolve to minimal stand-ins only when the real
package is not installed.  See INTEGRATION.md.
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SITE = os.path.join(HERE, "site")
SHIMS = os.path.join(HERE, "shims")


def install():
    if SITE not in sys.path:
        sys.path.insert(0, SITE)
    if SHIMS not in sys.path:
        sys.path.append(SHIMS)
    stale = [m for m in sys.modules if m == "compressai" or m.startswith("compressai.")]
    for m in stale:
        f = getattr(sys.modules[m], "__file__", "") or ""
        if not f.startswith(SITE):
            del sys.modules[m]
    return SITE
