"""``compressai.ans`` (pybind surface of compressai/cpp_exts/rans/rans_interface.cpp:352-372) on the
host-side coder of libhesic_b200.so.  List arguments are accepted as in the reference; numpy / torch
int32 arrays are accepted too and avoid the Python-list marshalling."""
import numpy as np

from hesic_b200.functional import RansDecoderHandle, RansEncoderHandle


def _tables(cdfs, cdfs_sizes, offsets):
    if isinstance(cdfs, np.ndarray):
        table = np.ascontiguousarray(cdfs, dtype=np.int32)
    else:
        pitch = max(len(c) for c in cdfs)
        table = np.zeros((len(cdfs), pitch), dtype=np.int32)
        for i, c in enumerate(cdfs):
            table[i, :len(c)] = c
    return table, np.asarray(cdfs_sizes, dtype=np.int32), np.asarray(offsets, dtype=np.int32)


class BufferedRansEncoder:
    def __init__(self):
        self._h = RansEncoderHandle()

    def encode_with_indexes(self, symbols, indexes, cdfs, cdfs_sizes, offsets):
        self._h.push(np.asarray(symbols, dtype=np.int32), np.asarray(indexes, dtype=np.int32), *_tables(cdfs, cdfs_sizes, offsets))

    def flush(self):
        return self._h.flush()


class RansEncoder:
    def encode_with_indexes(self, symbols, indexes, cdfs, cdfs_sizes, offsets):
        enc = BufferedRansEncoder()
        enc.encode_with_indexes(symbols, indexes, cdfs, cdfs_sizes, offsets)
        return enc.flush()


class RansDecoder:
    def __init__(self):
        self._h = RansDecoderHandle()

    def set_stream(self, encoded):
        self._h.set_stream(encoded)

    def decode_stream(self, indexes, cdfs, cdfs_sizes, offsets):
        return self._h.decode(np.asarray(indexes, dtype=np.int32), *_tables(cdfs, cdfs_sizes, offsets)).tolist()

    def decode_with_indexes(self, encoded, indexes, cdfs, cdfs_sizes, offsets):
        self.set_stream(encoded)
        return self.decode_stream(indexes, cdfs, cdfs_sizes, offsets)
