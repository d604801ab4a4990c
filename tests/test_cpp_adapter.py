"""Runs the C++ parity test of the drop-in gpusim::FingerprintDB adapter (tests/cpp/test_adapter.cpp,
a mirror of reference test/test_gpusim.cpp) and checks the native .fsim reader against the
Python one."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import GOLDEN, ROOT
from gpusimilarity_b200 import _lib
from gpusimilarity_b200.fsim import read_fsim, write_fsim
from oracle import oracle as O

BIN = os.path.join(ROOT, "tests", "cpp", "test_adapter")


def _run(env_extra):
    if not os.path.exists(BIN):
        subprocess.run(["make", "-C", ROOT, "adapter"], check=True, capture_output=True)
    env = dict(os.environ, **env_extra)
    return subprocess.run([BIN, os.path.join(GOLDEN, "small.fsim")], capture_output=True, text=True, env=env,
                          timeout=300)


def test_adapter_cpu_cases():
    """reference CI mode (.travis.yml:21): SKIP_CUDA=1 — CPUSort, FoldFingerprint, search_cpu."""
    res = _run({"SKIP_CUDA": "1"})
    assert res.returncode == 0, res.stdout + res.stderr
    assert "OK:" in res.stdout and "GPU cases skipped" in res.stdout


@pytest.mark.gpu
def test_adapter_gpu_cases():
    res = _run({})
    assert res.returncode == 0, res.stdout + res.stderr
    assert "GPU cases ran" in res.stdout and "OK:" in res.stdout


def _native_read(path):
    lib = _lib.lib()
    h = C.c_void_p()
    rc = lib.gsb_fsim_open(path.encode(), C.byref(h))
    if rc != 0:
        raise RuntimeError(lib.gsb_fsim_last_error().decode())
    try:
        chunks = [C.string_at(lib.gsb_fsim_chunk_data(h, i), lib.gsb_fsim_chunk_bytes(h, i))
                  for i in range(lib.gsb_fsim_chunk_count(h))]
        smiles = [lib.gsb_fsim_string(h, 0, i) for i in range(lib.gsb_fsim_string_count(h, 0))]
        ids = [lib.gsb_fsim_string(h, 1, i) for i in range(lib.gsb_fsim_string_count(h, 1))]
        return (lib.gsb_fsim_dbkey(h).decode(), lib.gsb_fsim_fp_bits(h), lib.gsb_fsim_fp_count(h), chunks,
                smiles, ids)
    finally:
        lib.gsb_fsim_close(h)


def test_native_fsim_reader_matches_python_reader(tmp_path, small_fsim):
    key, bits, count, chunks, smiles, ids = _native_read(os.path.join(GOLDEN, "small.fsim"))
    assert (key, bits, count) == ("pass", 1024, 100)
    assert chunks == small_fsim.fp_chunks and smiles == small_fsim.smiles and ids == small_fsim.ids
    # several chunks per section, odd sizes
    rows = O.synth_db(3, 777, 32, 0)
    sm = [b"S%d" % i * (i % 5 + 1) for i in range(777)]
    idl = [b"ID%05d" % i for i in range(777)]
    path = str(tmp_path / "multi.fsim")
    write_fsim(path, rows, sm, idl, dbkey="k2", chunk_bytes=9000)
    key, bits, count, chunks, smiles, ids = _native_read(path)
    ref = read_fsim(path)
    assert len(chunks) == len(ref.fp_chunks) > 3 and chunks == ref.fp_chunks
    assert smiles == sm and ids == idl and key == "k2" and count == 777
    write_fsim(path, rows, sm, idl, version=2)
    with pytest.raises(RuntimeError, match="version incompatible"):
        _native_read(path)
    with pytest.raises(RuntimeError):
        _native_read(str(tmp_path / "missing.fsim"))
