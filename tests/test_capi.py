"""The C-ABI library loads on a CPU-only box, exports every symbol include/frameino_b200.h declares, and refuses to
compute without a GPU (no CPU fallback)."""
import os
import re

import pytest
import torch

from frameino_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "frameino_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(fino_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_hot_path_entry_points():
    syms = _declared_symbols()
    for needed in ("fino_gemm_bf16", "fino_attention_fwd", "fino_ln_modulate", "fino_gate_residual", "fino_qk_norm_rope",
                   "fino_patchify", "fino_unpatchify", "fino_timestep_embedding", "fino_linear_small_m",
                   "fino_build_mod_table", "fino_swap01"):
        assert needed in syms


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    for name in _declared_symbols():
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
        assert name in _lib.SIGNATURES, f"{name} has no ctypes signature in frameino_b200/_lib.py"
    assert lib.fino_abi_version() == 1


def test_ctypes_signatures_match_header_arity():
    text = open(os.path.join(ROOT, "include", "frameino_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    for name, (_, argtypes) in _lib.SIGNATURES.items():
        m = re.search(r"\b" + name + r"\s*\(([^;]*?)\)\s*;", text, flags=re.S)
        assert m, name
        args = m.group(1).strip()
        n = 0 if args in ("", "void") else len(args.split(","))
        assert n == len(argtypes), f"{name}: header has {n} parameters, ctypes has {len(argtypes)}"


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_compute_entry_points_fail_loudly_without_gpu():
    lib = _lib.load()
    status = lib.fino_timestep_embedding(None, None, 1, 4, 1, 0.0, 1.0, 10000.0, None)
    assert status != 0
    assert b"no CPU fallback" in lib.fino_last_error() or b"CUDA" in lib.fino_last_error()


def test_ops_reject_cpu_tensors():
    from frameino_b200 import ops

    with pytest.raises(RuntimeError, match="no CPU"):
        ops.ln_modulate(torch.zeros(4, 64, dtype=torch.bfloat16), 1e-6)


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "frameino_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), fn


def test_header_is_plain_c_and_links_against_the_library(tmp_path):
    """include/frameino_b200.h is the contract for bindings in any language: it must compile as C99 (no C++ leaking in),
    and a C program linked against the library must resolve the entry points and get the no-GPU status from a compute
    call (or run it, on a GPU box)."""
    import shutil
    import subprocess

    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no gcc")
    src = tmp_path / "abi_check.c"
    src.write_text(
        '#include "frameino_b200.h"\n#include <stdio.h>\n'
        "int main(void) {\n"
        "  int n_full = -1, splits = -1;\n"
        "  if (fino_abi_version() != FINO_ABI_VERSION) return 10;\n"
        "  if (fino_attention_plan(28160, 28160, 3, 1, 148, -1, &n_full, &splits) != 0) return 11;\n"
        "  if (n_full != 296 || splits < 2) return 12;   /* 330 tiles on 148 SMs: 2 full waves + split last wave */\n"
        "  if (fino_gemm_set_mode(99) == 0) return 13;    /* bad argument -> status 1 + message */\n"
        "  if (fino_last_error() == 0 || fino_last_error()[0] == 0) return 14;\n"
        '  printf("abi %d n_full %d splits %d\\n", fino_abi_version(), n_full, splits);\n'
        "  return 0;\n}\n")
    exe = tmp_path / "abi_check"
    libdir = os.path.dirname(_lib.LIB_PATH)
    cmd = [gcc, "-std=c99", "-pedantic", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe),
           "-L", libdir, "-l:libframeino_b200.so", f"-Wl,-rpath,{libdir}"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    run = subprocess.run([str(exe)], capture_output=True, text=True)
    assert run.returncode == 0, (run.returncode, run.stdout, run.stderr)
    assert run.stdout.startswith("abi 1 n_full 296")
