"""Host-side mirror of the reference's ``GPUSimServer`` (gpusim.h:23-95) over the C ABI, plus the
client half of its socket protocol (reference python/gpusim_search.py:36-71,
python/gpusim_server.py:77-151) written with ``struct`` instead of PyQt."""
from __future__ import annotations

import ctypes as C
import socket
import struct
import time
from typing import Dict, List, Sequence, Tuple

import numpy as np

from ._lib import GsbError, lib

SOCKET_PATH = "/tmp/gpusimilarity"  # QLocalServer name "gpusimilarity" (gpusim.cpp:257-260)


def _cstr(s: bytes) -> bytes:
    return struct.pack(">I", len(s) + 1) + s + b"\0"


def encode_request(dbname_to_key: Dict[str, str], request_num: int, results_requested: int,
                   similarity_cutoff: float, fingerprint: Sequence[int]) -> bytes:
    """What the reference clients write (gpusim_search.py:36-47): note the cutoff is a QDataStream
    ``float`` and therefore travels as an 8-byte double."""
    out = [struct.pack(">i", len(dbname_to_key))]
    for name, key in dbname_to_key.items():
        out += [_cstr(name.encode()), _cstr(key.encode())]
    fp = np.ascontiguousarray(fingerprint, dtype="<i4").tobytes()
    out += [struct.pack(">iid", request_num, results_requested, similarity_cutoff), struct.pack(">I", len(fp)), fp]
    return b"".join(out)


def decode_response(buf: bytes) -> Tuple[int, int, List[bytes], List[bytes], List[float]]:
    """(request_num, approximate_count, smiles, ids, scores), gpusim_server.py:137-151."""
    request_num, n, approx = struct.unpack_from(">iiQ", buf, 0)
    off = 16

    def strings():
        nonlocal off
        out = []
        for _ in range(n):
            (ln,) = struct.unpack_from(">I", buf, off)
            out.append(buf[off + 4:off + 4 + ln - 1])
            off += 4 + ln
        return out

    smiles, ids = strings(), strings()
    scores = list(struct.unpack_from(">%dd" % n, buf, off))
    return request_num, approx, smiles, ids, scores


def _check(rc: int) -> None:
    if rc != 0:
        raise GsbError(rc, lib().gsb_server_last_error().decode())


class GPUSimServer:
    """``GPUSimServer(database_fnames, gpu_bitcount=0)`` (gpusim.cpp:87-166)."""

    def __init__(self, database_fnames: Sequence[str], gpu_bitcount: int = 0, use_gpu: bool = True):
        self._h = C.c_void_p()
        arr = (C.c_char_p * max(len(database_fnames), 1))(*[f.encode() for f in database_fnames])
        _check(lib().gsb_server_create(arr, len(database_fnames), gpu_bitcount, 1 if use_gpu else 0, C.byref(self._h)))

    def close(self):
        if self._h:
            lib().gsb_server_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def setUseGPU(self, use_gpu: bool) -> None:
        lib().gsb_server_set_use_gpu(self._h, 1 if use_gpu else 0)

    def usingGPU(self) -> bool:
        return bool(lib().gsb_server_using_gpu(self._h))

    def foldFactor(self) -> int:
        return lib().gsb_server_fold_factor(self._h)

    def getFingerprint(self, index: int, dbname: str, words: int = 32) -> np.ndarray:
        out = np.empty(words, dtype=np.int32)
        _check(lib().gsb_server_get_fingerprint(self._h, dbname.encode(), index, out.ctypes.data))
        return out

    def handleRequest(self, request: bytes) -> bytes:
        resp, n = C.c_void_p(), C.c_uint64(0)
        _check(lib().gsb_server_handle_request(self._h, request, len(request), C.byref(resp), C.byref(n)))
        try:
            return C.string_at(resp, n.value)
        finally:
            lib().gsb_server_free(resp)

    def handleBatch(self, requests: Sequence[bytes]) -> List[bytes]:
        """Requests of the same shape answered from one pass over each database."""
        n = len(requests)
        bufs = [C.create_string_buffer(r, len(r)) for r in requests]
        ptrs = (C.c_void_p * n)(*[C.cast(b, C.c_void_p) for b in bufs])
        sizes = (C.c_uint64 * n)(*[len(r) for r in requests])
        out, out_n = (C.c_void_p * n)(), (C.c_uint64 * n)()
        _check(lib().gsb_server_handle_batch(self._h, ptrs, sizes, n, out, out_n))
        res = []
        for i in range(n):
            res.append(C.string_at(out[i], out_n[i]))
            lib().gsb_server_free(out[i])
        return res

    def searchDatabases(self, reference, results_requested: int, similarity_cutoff: float,
                        dbname_to_key: Dict[str, str]):
        """reference searchDatabases (gpusim.cpp:306-374): (smiles, ids, scores, approximate count)."""
        _, approx, smiles, ids, scores = decode_response(
            self.handleRequest(encode_request(dbname_to_key, 0, results_requested, similarity_cutoff, reference)))
        return smiles, ids, scores, approx

    def listen(self, socket_path: str = SOCKET_PATH) -> None:
        _check(lib().gsb_server_listen(self._h, socket_path.encode()))

    def serve(self, max_requests: int = 0) -> None:
        _check(lib().gsb_server_serve(self._h, max_requests))

    def stop(self) -> None:
        lib().gsb_server_stop(self._h)


def search_over_socket(request: bytes, socket_path: str = SOCKET_PATH, timeout: float = 30.0) -> bytes:
    """Client side: one request, one response (gpusim_search.py:49-52)."""
    with socket.socket(socket.AF_UNIX, socket.SOCK_STREAM) as s:
        s.settimeout(timeout)
        deadline = time.monotonic() + timeout
        while True:
            try:
                s.connect(socket_path)
                break
            except BlockingIOError:   # the daemon's accept queue is full: a unix socket says EAGAIN at once
                if time.monotonic() > deadline:
                    raise
                time.sleep(0.0005)
        s.sendall(request)
        buf = b""
        while True:
            chunk = s.recv(1 << 20)
            if not chunk:
                break
            buf += chunk
            if len(buf) >= 16:
                try:
                    decode_response(buf)
                    return buf
                except (struct.error, IndexError):
                    continue
        return buf


def search_socket(socket_path: str, dbname_to_key: Dict[str, str], fingerprint: Sequence[int], results_requested: int,
                  similarity_cutoff: float, request_num: int = 1, timeout: float = 30.0):
    """One search against a running daemon (this repo's or the reference's): returns
    (approximate_count, smiles, ids, scores) with the strings decoded."""
    buf = search_over_socket(encode_request(dbname_to_key, request_num, results_requested, similarity_cutoff,
                                            fingerprint), socket_path, timeout)
    got_num, approx, smiles, ids, scores = decode_response(buf)
    if got_num != request_num:
        raise GsbError(6, f"response for request {got_num}, expected {request_num}")
    return approx, [x.decode() for x in smiles], [x.decode() for x in ids], scores
