"""CPU checks of the C-ABI boundary: the library builds, loads and exports every symbol include/rvc_b200.h declares."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from rvc_maker_b200 import _lib, build

    build.build()
    return _lib.load()


def header_symbols():
    src = open(os.path.join(ROOT, "include", "rvc_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(rvc_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported(lib):
    from rvc_maker_b200 import _lib

    syms = header_symbols()
    assert len(syms) >= 10
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/rvc_b200.h but not exported"
        assert s in _lib.SIGNATURES, f"{s} has no ctypes signature"
    assert sorted(_lib.SIGNATURES) == syms


def test_abi_version_and_error_string(lib):
    assert lib.rvc_abi_version() == 30
    # argument validation happens before any CUDA call, so it is checkable without a GPU
    rc = lib.rvc_conv1d_f32(None, None)
    assert rc == -1
    assert b"null descriptor" in lib.rvc_last_error()


def test_desc_struct_layout_matches_header(tmp_path):
    """ctypes struct sizes must equal the C structs: sizeof() as gcc sees include/rvc_b200.h (plain C, LP64) against the ctypes mirrors."""
    import subprocess

    from rvc_maker_b200._lib import Conv1dDesc, Conv1dTcDesc, ConvTr1dDesc, PairJob, ResblockPairDesc, TcJob

    pairs = [("rvc_conv1d_desc", Conv1dDesc), ("rvc_convtr1d_desc", ConvTr1dDesc), ("rvc_tc_job", TcJob), ("rvc_conv1d_tc_desc", Conv1dTcDesc),
             ("rvc_pair_job", PairJob), ("rvc_resblock_pair_desc", ResblockPairDesc)]
    src = tmp_path / "sizes.c"
    src.write_text('#include <stdio.h>\n#include "rvc_b200.h"\nint main(void) {\n' +
                   "".join(f'  printf("%zu\\n", sizeof({c}));\n' for c, _ in pairs) + "  return 0;\n}\n")
    exe = tmp_path / "sizes"
    subprocess.run(["gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    sizes = [int(v) for v in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()]
    for (cname, ct), sz in zip(pairs, sizes):
        assert ctypes.sizeof(ct) == sz, (cname, ctypes.sizeof(ct), sz)
    assert ctypes.sizeof(TcJob) == 240 and ctypes.sizeof(PairJob) == 96
    assert lib_rows(1000) == 64 + 1024 + 64


def lib_rows(L):
    from rvc_maker_b200 import _lib

    return _lib.load().rvc_packed_rows(L)


def test_no_oracle_import_in_product():
    """oracle/ is test infrastructure: only tests/, __graft_entry__.smoke() and the baseline legs of bench.py may import it."""
    for sub in ("rvc_maker_b200", "tools"):
        pkg = os.path.join(ROOT, sub)
        for fn in os.listdir(pkg):
            if fn.endswith(".py"):
                src = open(os.path.join(pkg, fn)).read()
                assert "import oracle" not in src and "from oracle" not in src, (sub, fn)
    bench = open(os.path.join(ROOT, "bench.py")).read()
    # in bench.py the oracle may only appear inside cpu_baseline() / gpu_eager_baseline() (the reported baselines, `--impl reference` included)
    import ast
    for node in ast.parse(bench).body:
        uses = "oracle" in (ast.get_source_segment(bench, node) or "") and any(
            isinstance(n, (ast.Import, ast.ImportFrom)) and "oracle" in ast.dump(n) for n in ast.walk(node))
        if uses:
            assert isinstance(node, ast.FunctionDef) and node.name in ("cpu_baseline", "gpu_eager_baseline"), getattr(node, "name", node)


def test_f0_tail_argument_errors(lib):
    """rvc_rmvpe_decode_f32 / rvc_f0_coarse_f64 / rvc_rmvpe_pitch_f32 validate before any CUDA call."""
    assert lib.rvc_rmvpe_decode_f32(None, 360, 10, 0.03, 0.0, 0.0, None, None) == -1
    assert lib.rvc_rmvpe_decode_f32(8, 359, 10, 0.03, 0.0, 0.0, 8, None) == -2 and b"360" in lib.rvc_last_error()
    assert lib.rvc_f0_coarse_f64(8, 10, None, 0, 1.0, 0.0, 1.0, 2.0, 8, 8, None, None) == -1 and b"shift" in lib.rvc_last_error()
    assert lib.rvc_f0_coarse_f64(8, 10, 8, 0, 1.0, 1.0, 1.0, 2.0, 8, 8, None, None) == -1 and b"n_notes" in lib.rvc_last_error()
    assert lib.rvc_rmvpe_pitch_f32(8, 360, 10, 0.03, 0.0, 0.0, None, 0, 1.0, 1.0, 2.0, 1.0, 8, 8, None, None) == -1 and b"mel_max" in lib.rvc_last_error()
    assert lib.rvc_rmvpe_pitch_f32(8, 360, 0, 0.03, 0.0, 0.0, None, 0, 1.0, 1.0, 1.0, 2.0, 8, 8, None, None) == 0      # T == 0: nothing to do


def test_pitch_module_refuses_cpu_tensors():
    import numpy as np
    import torch

    from rvc_maker_b200 import pitch

    with pytest.raises(RuntimeError, match="CUDA"):
        pitch.decode(torch.zeros(4, 360))
    with pytest.raises(RuntimeError, match="CUDA"):
        pitch.rmvpe_pitch(np.zeros((4, 360), np.float32))
    with pytest.raises(RuntimeError, match="CUDA"):
        pitch.get_f0_tail(torch.zeros(4, dtype=torch.float64))


def test_split_points_argument_errors(lib):
    assert lib.rvc_split_points_workspace(100 * 16000, 38 * 16000, 6 * 16000) == 2 * 94 * 16
    assert lib.rvc_split_points(8, 1, 100 * 16000, 161, 38 * 16000, 6 * 16000, 8, 8, None) == -2 and b"even" in lib.rvc_last_error()
    assert lib.rvc_split_points(8, 1, 50, 160, 38 * 16000, 6 * 16000, 8, 8, None) == -2
    assert lib.rvc_split_points(8, 1, 100 * 16000, 160, 16000, 6 * 16000, 8, 8, None) == -2 and b"t_query" in lib.rvc_last_error()
    assert lib.rvc_split_points(None, 1, 100 * 16000, 160, 38 * 16000, 6 * 16000, 8, 8, None) == -1
    assert lib.rvc_split_points(None, 1, 30 * 16000, 160, 38 * 16000, 6 * 16000, None, None, None) == 0      # no candidate: nothing to do


def test_filtfilt_argument_errors(lib):
    import numpy as np

    one = np.ones(6)
    p = one.ctypes.data
    assert lib.rvc_filtfilt_block(9_600_000) == 4096 and lib.rvc_filtfilt_block(160000) == 4096
    assert lib.rvc_filtfilt_workspace(160000, 6) == (160036 + 2 * 40 * 8) * 8
    assert lib.rvc_filtfilt(8, 0, 1000, p, p, 12, p, p, 8, 8, None) == -2
    assert lib.rvc_filtfilt(8, 0, 18, p, p, 6, p, p, 8, 8, None) == -2 and b"padlen" in lib.rvc_last_error()
    two = 2 * one
    assert lib.rvc_filtfilt(8, 0, 1000, p, two.ctypes.data, 6, p, p, 8, 8, None) == -1 and b"normalise" in lib.rvc_last_error()
    assert lib.rvc_filtfilt(None, 0, 1000, p, p, 6, p, p, None, None, None) == -1


def test_retrieval_blend_argument_errors(lib):
    assert lib.rvc_retrieval_blend_f32(8, 8, 8, 8, 0.5, 0.5, 8, 10, 768, 4, None) == -2 and b"k must be 8" in lib.rvc_last_error()
    assert lib.rvc_retrieval_blend_f32(None, 8, 8, 8, 0.5, 0.5, 8, 10, 768, 8, None) == -1
    assert lib.rvc_retrieval_blend_f32(8, 8, 8, 8, 0.5, 0.5, 8, 0, 768, 8, None) == 0
