"""``matplotlib.pyplot`` stand-in: figures and axes that accept any call and draw nothing (see the package docstring)."""
_HESIC_STUB = True
calls = []          # (name, args, kwargs) of every plotting call, for anyone who wants to inspect them


class _Sink:
    def __init__(self, name="figure"):
        self._name = name

    def __getattr__(self, k):
        def call(*a, **kw):
            calls.append((f"{self._name}.{k}", a, kw))
            return _Sink(k)
        return call

    def __iter__(self):
        return iter(())


def figure(*a, **kw):
    calls.append(("figure", a, kw))
    return _Sink("figure")


def subplots(nrows=1, ncols=1, *a, **kw):
    calls.append(("subplots", (nrows, ncols) + a, kw))
    n = nrows * ncols
    return _Sink("figure"), (_Sink("axes") if n == 1 else [_Sink("axes") for _ in range(n)])


def __getattr__(name):
    if name.startswith("__"):
        raise AttributeError(name)

    def call(*a, **kw):
        calls.append((name, a, kw))
        return None
    return call
