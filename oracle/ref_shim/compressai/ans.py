"""compressai.ans stand-in -> oracle/lbic_oracle.c (test infrastructure only).

Same list-in / bytes-out convention as the pybind11 module the reference binds at
graphs/models/BlockBasedImgCompLossy_net.py:9,328,359-360,409-410,439.
"""
from oracle import native as _native


class BufferedRansEncoder:
    def __init__(self):
        self._symbols, self._indexes, self._tables = [], [], None

    def encode_with_indexes(self, symbols, indexes, cdfs, cdfs_sizes, offsets):
        self._symbols.extend(symbols)
        self._indexes.extend(indexes)
        self._tables = _native.Tables(cdfs, cdfs_sizes, offsets)

    def flush(self):
        out = _native.rans_encode(self._symbols, self._indexes, self._tables)
        self._symbols, self._indexes = [], []
        return out


class RansEncoder:
    def encode_with_indexes(self, symbols, indexes, cdfs, cdfs_sizes, offsets):
        return _native.rans_encode(symbols, indexes, _native.Tables(cdfs, cdfs_sizes, offsets))


class RansDecoder:
    def __init__(self):
        self._dec = _native.RansDecoder()
        self._tables_key, self._tables = None, None

    def set_stream(self, stream):
        self._dec.set_stream(stream)

    def _get_tables(self, cdfs, cdfs_sizes, offsets):
        # the reference re-marshals the full 64x3133 list-of-lists on every per-block call
        # (BlockBasedImgCompLossy_net.py:439); cache the conversion by identity
        key = id(cdfs)
        if key != self._tables_key:
            self._tables_key, self._tables = key, _native.Tables(cdfs, cdfs_sizes, offsets)
        return self._tables

    def decode_stream(self, indexes, cdfs, cdfs_sizes, offsets):
        return [int(v) for v in self._dec.decode_stream(indexes, self._get_tables(cdfs, cdfs_sizes, offsets))]

    def decode_with_indexes(self, stream, indexes, cdfs, cdfs_sizes, offsets):
        return [int(v) for v in _native.rans_decode(stream, indexes, _native.Tables(cdfs, cdfs_sizes, offsets))]
