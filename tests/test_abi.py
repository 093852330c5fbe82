"""The C-ABI library loads and exports exactly what include/b200_decode.h declares; the ctypes prototypes agree with
the header; without a GPU every compute entry point fails loudly (no CPU fallback) — CPU only, no compute calls."""
import ctypes as C
import re
import subprocess
from pathlib import Path

import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent
HEADER = (ROOT / "include" / "b200_decode.h").read_text()


def declared_functions():
    body = re.sub(r"/\*.*?\*/", "", HEADER, flags=re.S)
    return sorted(set(re.findall(r"\b(b200_[a-z0-9_]+)\s*\(", body)))


def test_header_declares_every_boundary_b_op():
    names = declared_functions()
    for op in ("gemv", "rmsnorm", "rope", "attn", "silu_mul", "add", "embedding", "argmax"):
        assert f"b200_{op}_bf16" in names
    for fn in ("create", "destroy", "reset", "seek", "forward", "decode", "last_token", "create_tp"):
        assert f"b200_engine_{fn}" in names


def test_library_exports_every_declared_symbol(built_lib):
    out = subprocess.run(["nm", "-D", "--defined-only", str(built_lib)], capture_output=True, text=True, check=True).stdout
    exported = {l.split()[-1] for l in out.splitlines() if " T " in l}
    declared = set(declared_functions())
    assert declared <= exported, f"declared but not exported: {sorted(declared - exported)}"
    assert {e for e in exported if e.startswith("b200_")} <= declared, "exported b200_* symbol missing from the header"


def test_ctypes_prototypes_cover_the_header(built_lib):
    from tinygpt_b200 import _lib
    assert set(_lib.PROTOTYPES) == set(declared_functions())
    h = _lib.lib()
    assert h.b200_abi_version() == int(re.search(r"#define B200_ABI_VERSION (\d+)", HEADER).group(1))
    # argument counts agree with the header
    body = re.sub(r"/\*.*?\*/", "", HEADER, flags=re.S)
    for name, (_, args) in _lib.PROTOTYPES.items():
        m = re.search(rf"\b{name}\s*\(([^;]*?)\)\s*;", body, flags=re.S)
        assert m, name
        params = m.group(1).strip()
        n = 0 if params in ("", "void") else params.count(",") + 1
        assert n == len(args), f"{name}: header has {n} parameters, ctypes prototype {len(args)}"


def test_struct_layouts_match_header(built_lib):
    from tinygpt_b200 import _lib
    assert C.sizeof(_lib.ModelDesc) == 14 * 4
    assert C.sizeof(_lib.LayerWeights) == 9 * 8
    assert C.sizeof(_lib.WeightTable) == 5 * 8
    assert _lib.IPC_HANDLE_BYTES == int(re.search(r"#define B200_IPC_HANDLE_BYTES (\d+)", HEADER).group(1))


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_no_gpu_means_loud_failure_not_fallback(built_lib):
    from tinygpt_b200 import _lib, engine, models, ops
    h = _lib.lib()
    assert h.b200_device_check() == -4  # B200_ERR_NO_DEVICE
    assert b"CUDA" in h.b200_last_error() or b"device" in h.b200_last_error()
    with pytest.raises(_lib.B200Error):
        _lib.require_device()
    x = torch.zeros(8, dtype=torch.bfloat16)
    with pytest.raises(_lib.B200Error):
        ops.add(x, x)
    with pytest.raises(_lib.B200Error):
        engine.DecodeEngine(models.TINY_QWEN2, models.synth_weights(models.TINY_QWEN2))


def test_product_does_not_import_the_oracle():
    """A product path that routes through oracle/ voids every parity claim: no file of the package mentions it."""
    for p in (ROOT / "tinygpt_b200").rglob("*.py"):
        src = p.read_text()
        assert "import oracle" not in src and "from oracle" not in src and "decode_oracle" not in src, p
    for p in (ROOT / "tinygpt_b200" / "csrc").glob("*"):
        assert "#include \"../../oracle" not in p.read_text() and "oracle/" not in p.read_text(), p


def test_model_byte_counts_match_survey():
    from tinygpt_b200 import models
    want = {"Qwen2.5-0.5B": 987_922_432, "Llama-3.2-3B": 6_425_149_440, "Qwen3-1.7B": 3_440_902_144,
            "Mistral-7B-v0.3": 14_227_079_168}  # SURVEY.md §8d weights-only bytes per token
    for name, b in want.items():
        assert 2 * models.SPECS[name].weight_params == b


def test_rope_table_host_equals_oracle():
    from helpers import orc, to_oracle_cfg
    from tinygpt_b200 import models
    for spec in (models.TINY_QWEN2, models.TINY_LLAMA, models.QWEN25_05B.with_ctx(300), models.LLAMA32_3B.with_ctx(300)):
        cfg = to_oracle_cfg(spec)
        assert torch.equal(models.rope_table(spec), orc.rope_table(spec.head_dim, spec.max_ctx, spec.rope_theta,
                                                                   cfg.rope_scaling))


def test_header_is_plain_c_and_cxx():
    """The boundary is a C ABI: the header must compile as C99 (pedantic) and as C++17 on its own."""
    hdr = str(ROOT / "include" / "b200_decode.h")
    for cmd in (["gcc", "-x", "c", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only", hdr],
                ["g++", "-x", "c++", "-std=c++17", "-Wall", "-Werror", "-fsyntax-only", hdr]):
        r = subprocess.run(cmd, capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
