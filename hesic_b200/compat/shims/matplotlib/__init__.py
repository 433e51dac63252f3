"""Stand-in for ``matplotlib`` used only when the real package is not installed (``compat/shims`` sits at the END of
``sys.path``).  The reference's drivers import ``matplotlib`` / ``matplotlib.pyplot`` at module level
(ywz/mywork/test3real.py:36-37, codec-test/test2_codec.py:34-35) but never plot on the evaluated path, so the names
resolve and every plotting call is a recorded no-op."""
_HESIC_STUB = True
__version__ = "0.0+hesic_b200.shim"
_backend = "agg"


def use(backend, *args, **kwargs):
    global _backend
    _backend = str(backend)


def get_backend():
    return _backend


rcParams = {}
