"""Qt-free reader/writer for gpusimilarity's ``.fsim`` database files.

Format (big-endian QDataStream, stream version Qt_5_2; only ints, C strings and
QByteArrays are used so the stream version does not change the layout):

    i32  version (== 3)                      reference gpusim.cpp:184-189
    cstr dbkey   (u32 len incl. NUL + bytes) reference gpusim.cpp:191-194
    i32  fp_bitcount                         reference gpusim.cpp:195
    i32  fp_count                            reference gpusim.cpp:196
    i32  n_fp_chunks;  n x QByteArray(qCompress(raw fingerprint bytes))   gpusim.cpp:198-209
    i32  n_smi_chunks; n x QByteArray(qCompress(sequence of cstr))        gpusim.cpp:211-221
    i32  n_id_chunks;  n x QByteArray(qCompress(sequence of cstr))        gpusim.cpp:223-233

``qCompress`` = 4-byte big-endian uncompressed length + a zlib stream.  The writer
follows reference python/gpusim_createdb.py:86-98,135-143 (chunks split at 1 GiB;
``chunk_bytes`` lets tests force several chunks from a small database).

This is the host-side test/tooling twin of the native reader in
``csrc/fsim_reader.cpp``; both are checked against reference test/small.fsim.
"""
from __future__ import annotations

import struct
import zlib
from dataclasses import dataclass, field
from typing import List, Sequence

import numpy as np

DATABASE_VERSION = 3  # reference gpusim.cpp:43
GIGABYTE_SIZE = 2 ** 30  # reference python/gpusim_createdb.py:14
_NULL_LEN = 0xFFFFFFFF  # Qt encodes a null QByteArray / char* like this


class FsimError(RuntimeError):
    pass


@dataclass
class FsimData:
    dbkey: str
    fp_bitcount: int
    fp_count: int
    fp_chunks: List[bytes]  # raw packed fingerprints, fp_bitcount/8 bytes per row
    smiles: List[bytes] = field(default_factory=list)
    ids: List[bytes] = field(default_factory=list)

    def fingerprints(self) -> np.ndarray:
        """All rows as an (N, words) little-endian int32 matrix (reference
        fingerprintdb_cuda.cu:123-125 reinterprets the bytes the same way)."""
        words = self.fp_bitcount // 32
        raw = b"".join(self.fp_chunks)
        return np.frombuffer(raw, dtype="<i4").reshape(-1, words)


class _Reader:
    def __init__(self, buf: bytes):
        self.buf = buf
        self.off = 0

    def _take(self, n: int) -> bytes:
        if self.off + n > len(self.buf):
            raise FsimError("truncated .fsim stream")
        out = self.buf[self.off:self.off + n]
        self.off += n
        return out

    def i32(self) -> int:
        return struct.unpack(">i", self._take(4))[0]

    def u32(self) -> int:
        return struct.unpack(">I", self._take(4))[0]

    def bytearray_(self) -> bytes:
        n = self.u32()
        if n == _NULL_LEN:
            return b""
        return self._take(n)

    def cstr(self) -> bytes:
        raw = self.bytearray_()
        return raw[:-1] if raw.endswith(b"\0") else raw

    def at_end(self) -> bool:
        return self.off >= len(self.buf)


def q_uncompress(blob: bytes) -> bytes:
    """Qt's qUncompress: u32 BE expected length, then a zlib stream."""
    if len(blob) < 4:
        return b""
    expected = struct.unpack(">I", blob[:4])[0]
    out = zlib.decompress(blob[4:])
    if len(out) != expected:
        raise FsimError("qUncompress length mismatch")
    return out


def q_compress(raw: bytes, level: int = -1) -> bytes:
    return struct.pack(">I", len(raw)) + zlib.compress(raw, level)


def _read_strings(blob: bytes) -> List[bytes]:
    rd = _Reader(q_uncompress(blob))
    out = []
    while not rd.at_end():
        out.append(rd.cstr())
    return out


def read_fsim(path: str) -> FsimData:
    with open(path, "rb") as fh:
        rd = _Reader(fh.read())
    version = rd.i32()
    if version != DATABASE_VERSION:
        # same message as the reference (gpusim.cpp:186-189)
        raise FsimError("Database version incompatible with this GPUSim version")
    dbkey = rd.cstr().decode()
    fp_bitcount = rd.i32()
    fp_count = rd.i32()
    fp_chunks = [q_uncompress(rd.bytearray_()) for _ in range(rd.i32())]
    smiles: List[bytes] = []
    for _ in range(rd.i32()):
        smiles.extend(_read_strings(rd.bytearray_()))
    ids: List[bytes] = []
    for _ in range(rd.i32()):
        ids.extend(_read_strings(rd.bytearray_()))
    return FsimData(dbkey, fp_bitcount, fp_count, fp_chunks, smiles, ids)


def _cstr(s: bytes) -> bytes:
    return struct.pack(">I", len(s) + 1) + s + b"\0"


def _qba(b: bytes) -> bytes:
    return struct.pack(">I", len(b)) + b


def _split_rows(items: Sequence[bytes], limit: int) -> List[bytes]:
    """Chunking rule of reference createdb.py:56-75: start a new chunk once the
    current one has reached ``limit`` bytes; chunks hold whole items."""
    chunks, cur, size = [], [], 0
    for it in items:
        if size >= limit and cur:
            chunks.append(b"".join(cur))
            cur, size = [], 0
        cur.append(it)
        size += len(it)
    chunks.append(b"".join(cur))
    return chunks


def write_fsim(path: str, fingerprints: np.ndarray, smiles: Sequence[bytes],
               ids: Sequence[bytes], dbkey: str = "pass", fp_bitcount: int | None = None,
               chunk_bytes: int = GIGABYTE_SIZE, version: int = DATABASE_VERSION) -> None:
    """Write a synthetic .fsim (tests / tooling).  ``fingerprints`` is (N, words) int32."""
    fps = np.ascontiguousarray(fingerprints, dtype="<i4")
    n, words = fps.shape
    bits = fp_bitcount if fp_bitcount is not None else words * 32
    rows = [fps[i].tobytes() for i in range(n)]
    out = [struct.pack(">i", version), _cstr(dbkey.encode()), struct.pack(">ii", bits, n)]
    for items in (rows, [_cstr(s) for s in smiles], [_cstr(s) for s in ids]):
        chunks = _split_rows(items, chunk_bytes)
        out.append(struct.pack(">i", len(chunks)))
        out.extend(_qba(q_compress(c)) for c in chunks)
    with open(path, "wb") as fh:
        fh.write(b"".join(out))
