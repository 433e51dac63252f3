"""Stand-in for the un-vendored PyPI ``range_coder`` (used only when the real package is not installed): the calling
surface the reference's file codec needs (newnet1.py:33,912,983,1040,1123,1179,1252) -- ``RangeEncoder(path)``,
``.encode(data, cumFreq)``, ``.close()``, ``RangeDecoder(path)``, ``.decode(n, cumFreq)`` -- on the host range coder
of libhesic_b200.so.  The byte stream is this library's own format, NOT the PyPI package's (which the reference
neither vendors nor pins: SURVEY.md 8c), so files written with one cannot be read with the other."""
import numpy as np

from hesic_b200.functional import RangeDecoderHandle, RangeEncoderHandle

_HESIC_STUB = True


def _row(cum_freq, n):
    r = np.asarray(cum_freq, dtype=np.int32).reshape(1, -1)
    return np.repeat(r, n, axis=0) if n != 1 else r


class RangeEncoder:
    def __init__(self, filepath):
        self._path = filepath
        self._enc = RangeEncoderHandle()
        self._closed = False

    def encode(self, data, cumFreq):
        data = np.asarray(data, dtype=np.int32).reshape(-1)
        if data.size:
            self._enc.push(data, _row(cumFreq, data.size))

    def close(self):
        if not self._closed:
            with open(self._path, "wb") as f:
                f.write(self._enc.finish())
            self._closed = True

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class RangeDecoder:
    def __init__(self, filepath):
        with open(filepath, "rb") as f:
            self._dec = RangeDecoderHandle(f.read())

    def decode(self, size, cumFreq):
        if size <= 0:
            return []
        return [int(v) for v in self._dec.decode(_row(cumFreq, int(size)))]

    def close(self):
        pass


def prob_to_cum_freq(prob, resolution=1024):
    """Probabilities -> integer cumulative frequencies with every symbol kept codable (total = resolution)."""
    p = np.asarray(prob, dtype=np.float64).reshape(-1)
    freq = np.maximum(np.round(p / p.sum() * resolution), 1).astype(np.int64)
    freq[np.argmax(freq)] += resolution - int(freq.sum())
    return [0] + [int(v) for v in np.cumsum(freq)]
