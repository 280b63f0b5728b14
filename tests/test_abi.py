"""CPU: the C-ABI library builds, loads and exports every symbol include/ptt_b200.h declares; the Python
binding table mirrors the header; the `pointnet2_ops._ext` drop-in exposes the names the reference calls.
No compute calls (there is no GPU here)."""
import os
import re

import pytest
import torch

from conftest import REPO

HEADER = os.path.join(REPO, "include", "ptt_b200.h")
TUNING_HEADER = os.path.join(REPO, "include", "ptt_b200_tuning.h")


def declared_symbols(header=HEADER):
    text = open(header).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return re.findall(r"PTT_API\s+[\w\s\*]+?\b(ptt_\w+)\s*\(", text)


@pytest.fixture(scope="module")
def built_lib():
    from ptt_b200 import build
    return build.build()


def test_header_declares_the_path():
    names = declared_symbols()
    assert len(names) == len(set(names)) >= 28
    for must in ("ptt_furthest_point_sampling", "ptt_ball_query", "ptt_group_points", "ptt_gather_points",
                 "ptt_three_nn", "ptt_sa_mlp_fwd", "ptt_transformer_block_fwd", "ptt_knn"):
        assert must in names


def test_library_exports_every_declared_symbol(built_lib):
    import ctypes
    handle = ctypes.CDLL(built_lib)
    for name in declared_symbols():
        assert hasattr(handle, name), "libptt_b200.so does not export %s" % name


def test_binding_table_matches_header(built_lib):
    from ptt_b200 import _lib
    assert set(_lib.SIGNATURES) == set(declared_symbols())
    # argument counts agree with the C declarations
    text = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    for name, (_, argtypes) in _lib.SIGNATURES.items():
        m = re.search(r"\b%s\s*\(([^;]*?)\)\s*;" % name, text, flags=re.S)
        assert m, name
        args = m.group(1).strip()
        n = 0 if args in ("", "void") else len(args.split(","))
        assert n == len(argtypes), "%s: header has %d arguments, binding %d" % (name, n, len(argtypes))
    assert _lib.version().startswith("ptt_b200 ") and _lib.version().endswith("sm_100a")
    assert _lib.lib().ptt_error_string(-3).decode().startswith("workspace")


def test_tuning_surface_is_declared_bound_and_unused_by_the_product(built_lib):
    """Every exported symbol is declared in one of the two headers (no undeclared hooks), the process-wide tuning
    switches live in ptt_b200_tuning.h, and nothing under ptt_b200/ calls them."""
    import glob
    import subprocess
    from ptt_b200 import _lib
    tuning = declared_symbols(TUNING_HEADER)
    assert set(tuning) == set(_lib.TUNING_SIGNATURES) and not set(tuning) & set(declared_symbols())
    out = subprocess.run(["nm", "-D", "--defined-only", built_lib], capture_output=True, text=True, check=True).stdout
    exported = {ln.split()[-1] for ln in out.splitlines() if " T " in ln and ln.split()[-1].startswith("ptt_")}
    assert exported == set(declared_symbols()) | set(tuning), exported ^ (set(declared_symbols()) | set(tuning))
    for path in glob.glob(os.path.join(REPO, "ptt_b200", "**", "*.py"), recursive=True):
        if path.endswith("_lib.py"):
            continue
        text = open(path).read()
        for name in tuning:
            assert name not in text, "%s uses the tuning hook %s" % (path, name)


def test_headers_are_plain_c_and_the_pack_descriptor_layout_matches_the_binding(tmp_path):
    """include/*.h must compile as C99 (what a cgo / JNI / ctypes binding sees), and the device descriptor table that
    ptt_b200/train_ops.py builds with numpy for ptt_linear_pack_batch must have the C struct's size and field offsets."""
    import shutil
    import subprocess

    from ptt_b200 import train_ops
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    src = tmp_path / "hdr.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "ptt_b200.h"\n#include "ptt_b200_tuning.h"\n'
                   'int main(void) { printf("%zu %zu %zu %zu %zu %zu %zu %zu\\n", sizeof(PttPackDesc), offsetof(PttPackDesc, weight), '
                   'offsetof(PttPackDesc, ld_c), offsetof(PttPackDesc, ld_k), offsetof(PttPackDesc, bias), offsetof(PttPackDesc, params), '
                   'offsetof(PttPackDesc, K), offsetof(PttPackDesc, Cout)); return 0; }\n')
    exe = tmp_path / "hdr"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(REPO, "include"), str(src), "-o", str(exe)], check=True)
    got = [int(v) for v in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()]
    d = train_ops._DESC
    assert got == [d.itemsize] + [d.fields[n][1] for n in ("w", "ld_c", "ld_k", "b", "p", "K", "C")]


def test_fault_word_is_quiet_without_a_device(built_lib):
    from ptt_b200 import _lib
    assert _lib.fault_status() == 0          # never allocates, never touches CUDA
    _lib.fault_clear()
    assert _lib.lib().ptt_error_string(-4).decode().startswith("a kernel of this library gave up")


def test_workspace_queries_are_host_only(built_lib):
    import ctypes
    from ptt_b200 import _lib
    L = _lib.lib()
    assert L.ptt_furthest_point_sampling_workspace_bytes(48, 1024, 512) == 0
    assert L.ptt_furthest_point_sampling_workspace_bytes(2, 100000, 512) == 2 * 100000 * 4
    dims = (ctypes.c_int * 4)(131, 128, 128, 256)
    assert L.ptt_sa_params_floats(128, 3, dims) > 131 * 128 + 128 * 128 + 128 * 256
    assert L.ptt_sa_params_floats(127, 3, dims) == 0           # dims[0] must be C + 3
    assert L.ptt_sa_mlp_workspace_bytes(2, 512, 256, 32, 128, 3, dims) > 0
    assert L.ptt_transformer_params_floats(256, 512) >= 1839360
    assert L.ptt_transformer_block_workspace_bytes(2, 128, 16, 256, 512) > 0


def test_dropin_exposes_the_ext_surface(built_lib):
    import ptt_b200
    ext = ptt_b200.install_dropin()
    # every `_ext.*` the reference calls (pointnet2_utils.py:48,78,112,118,145,182,204,237,257,287)
    for name in ("furthest_point_sampling", "furthest_point_sampling_with_dist", "gather_points", "gather_points_grad",
                 "three_nn", "three_interpolate", "three_interpolate_grad", "group_points", "group_points_grad",
                 "ball_query"):
        assert callable(getattr(ext, name))
    # CPU tensors are refused loudly (upstream: "CPU not supported"), never silently computed
    with pytest.raises(RuntimeError):
        ext.furthest_point_sampling(torch.zeros(1, 8, 3), 4)
    with pytest.raises(RuntimeError):
        ext.ball_query(torch.zeros(1, 2, 3), torch.zeros(1, 8, 3), 0.3, 4)


def test_modules_keep_the_reference_state_dict_layout():
    from ptt_b200 import modules
    from test_oracle_golden import sa_state_dict, transformer_state_dict
    m = modules.PointnetSAModuleVotes(mlp=[128, 128, 128, 256], radius=0.5, nsample=32, normalize_xyz=True,
                                      sample_method="sequence")
    want = sa_state_dict([128, 128, 128, 256])
    got = m.state_dict()
    assert set(got) == set(want)
    for k in want:
        assert tuple(got[k].shape) == tuple(want[k].shape), k
    t = modules.TransformerBlock(256, 512, 16, heads=1, layers=1)
    want = transformer_state_dict("TransformerBlock", 256, 512)
    got = t.state_dict()
    assert set(got) == set(want)
    for k in want:
        assert tuple(got[k].shape) == tuple(want[k].shape), k
    with pytest.raises(RuntimeError):
        t(torch.zeros(1, 16, 3), torch.zeros(1, 16, 256))


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(REPO, "ptt_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(root, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, os.path.join(root, f)
