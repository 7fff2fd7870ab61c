"""The C-ABI library loads and exports every symbol include/gandanet.h declares (no GPU, no compute calls)."""
import os
import re

from conftest import ROOT


def _header_prototypes():
    src = open(os.path.join(ROOT, "include", "gandanet.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    src = re.sub(r"typedef\s+struct\s*\{.*?\}\s*\w+\s*;", "", src, flags=re.S)
    src = re.sub(r"typedef\s+enum\s*\{.*?\}\s*\w+\s*;", "", src, flags=re.S)
    src = re.sub(r"enum\s*\{.*?\}\s*;", "", src, flags=re.S)
    protos = {}
    for m in re.finditer(r"(?:int|size_t|const char\*)\s+(gdn_\w+)\s*\(([^;{]*?)\)\s*;", src, flags=re.S):
        args = m.group(2).strip()
        n = 0 if args in ("", "void") else len([a for a in args.split(",") if a.strip()])
        protos[m.group(1)] = n
    return protos


def test_library_is_built_and_loads():
    from gan_danet_b200 import _lib
    assert os.path.exists(_lib.LIB_PATH), "run python -m gan_danet_b200.build"
    lib = _lib.load()
    assert lib.gdn_version() >= 100


def test_every_declared_symbol_is_exported_and_bound():
    from gan_danet_b200 import _lib
    lib = _lib.load()
    protos = _header_prototypes()
    assert len(protos) >= 45
    for name, nargs in protos.items():
        assert hasattr(lib, name), f"{name} declared in gandanet.h but not exported"
        assert name in _lib.SIGNATURES, f"{name} has no ctypes signature"
        assert len(_lib.SIGNATURES[name][1]) == nargs, f"{name}: header has {nargs} args, binding {len(_lib.SIGNATURES[name][1])}"
    for name in _lib.SIGNATURES:
        assert name in protos, f"{name} bound but not declared in gandanet.h"


def test_sass_contains_blackwell_instructions():
    """tcgen05.mma / TMA / tcgen05.ld of the PAM kernel must be in the shipped binary (UTCHMMA, UTMALDG, LDTM)."""
    import shutil
    import subprocess
    from gan_danet_b200 import _lib
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        import pytest
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([cuobjdump, "-sass", _lib.LIB_PATH], capture_output=True, text=True).stdout
    for mnemonic in ("UTCHMMA", "UTMALDG", "LDTM"):
        assert mnemonic in sass, mnemonic


def test_no_cpu_fallback():
    import pytest
    import torch
    import gan_danet_b200 as g
    G = g.FlexibleUpsamplingModule(4, growth_rate=4, num_blocks=1, num_layers_per_block=1)
    with pytest.raises(g._lib.GdnError if hasattr(g, "_lib") else Exception):
        G(torch.zeros(1, 4, 4, 4))


def test_custom_op_layer_registered():
    """torch.ops.gandanet.* exist, infer shapes on meta tensors (register_fake) and have NO CPU kernel."""
    import pytest
    import torch
    import gan_danet_b200.ops as O
    for name in O.OPS:
        assert hasattr(torch.ops.gandanet, name), name
    x = torch.empty(2, 8, 16, 160, device="meta")
    q = torch.empty(2, 8, 16, 20, device="meta")
    g = torch.empty(1, device="meta")
    y, o, lse = torch.ops.gandanet.pam_fwd(x, q, q, x, g, "fp16")
    assert y.shape == x.shape and o.shape == (2, 128, 160) and lse.shape == (2, 128)
    y, attn = torch.ops.gandanet.cam_fwd(x, g, True)
    assert attn.shape == (2, 160, 160)
    w = torch.empty(24, 160, 3, 3, device="meta")
    assert torch.ops.gandanet.conv2d(x, w, None, 1, 1, 1, 0.0).shape == (2, 8, 16, 24)
    assert torch.ops.gandanet.conv2d(x, w, None, 2, 1, 0, 0.0).shape == (2, 4, 8, 24)
    assert torch.ops.gandanet.upsample_bicubic2x(x).shape == (2, 16, 32, 160)
    c = torch.empty(160, device="meta")
    assert torch.ops.gandanet.bn_stats_finalize(x, c, c, c, c, 1e-5, 0.1).shape == (4, 160)
    assert torch.ops.gandanet.bilinear_resize_add_fwd(x, torch.empty(2, 32, 64, 160, device="meta")).shape == (2, 32, 64, 160)
    hr = torch.empty(2, 1, 32, 64, device="meta")
    lo, dhr, dz = torch.ops.gandanet.multi_loss_fwd_bwd(hr, hr, torch.empty(2, 1, device="meta"), 0.02, 1e-5)
    assert lo.shape == (4,) and dhr.shape == hr.shape and dz.shape == (2, 1)
    with pytest.raises(NotImplementedError):
        torch.ops.gandanet.upsample_bicubic2x(torch.zeros(1, 2, 2, 4))
    with pytest.raises(NotImplementedError):
        torch.ops.gandanet.pam_fwd(torch.zeros(1, 8, 16, 8), torch.zeros(1, 8, 16, 1), torch.zeros(1, 8, 16, 1), torch.zeros(1, 8, 16, 8), torch.zeros(1), "fp32")


def test_python_free_harness_builds():
    """tools/cabi_bench.cpp (plain C++ over include/gandanet.h, no torch) compiles and links against the library: the header is usable from
    C++ as it stands and the harness stays in step with the argument structs."""
    import os
    import subprocess
    from gan_danet_b200.build import HERE, build_harness, build_library
    build_library()
    exe = build_harness()
    assert os.path.exists(exe)
    out = subprocess.run(["ldd", exe], capture_output=True, text=True).stdout
    assert "libgandanet_sm100.so" in out and os.path.join(HERE, "libgandanet_sm100.so") in out      # resolved through the $ORIGIN rpath


def test_product_never_touches_the_oracle_and_fails_loudly_without_the_library(tmp_path):
    """The oracle is test infrastructure: no module of the product package imports, names a path into, or executes anything under oracle/
    (only tests/, __graft_entry__.smoke() and bench.py's CPU arm may), and a missing library is an error, not a fallback."""
    import ast
    import pytest
    from gan_danet_b200 import _lib
    pkg = os.path.dirname(os.path.abspath(_lib.__file__))
    offenders = []
    for root, _, files in os.walk(pkg):
        for f in files:
            if not f.endswith(".py"):
                continue
            tree = ast.parse(open(os.path.join(root, f)).read())
            for node in ast.walk(tree):
                names = []
                if isinstance(node, ast.Import):
                    names = [a.name for a in node.names]
                elif isinstance(node, ast.ImportFrom):
                    names = [node.module or ""]
                for n in names:
                    if n.split(".")[0] in ("oracle", "gan_danet_oracle", "quantised_oracle", "notebook_step", "postprocess_oracle", "build_ref", "make_golden"):
                        offenders.append((f, n))
                if isinstance(node, ast.Constant) and isinstance(node.value, str) and "oracle" in node.value and ("/" in node.value or node.value == "oracle"):
                    # string constants naming an oracle PATH (docstrings and comments-as-strings mention oracle files by name: removed below)
                    offenders.append((f, node.value[:60]))
    docstrings = set()
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                for node in ast.walk(ast.parse(open(os.path.join(root, f)).read())):
                    if isinstance(node, ast.Expr) and isinstance(node.value, ast.Constant) and isinstance(node.value.value, str):
                        docstrings.add(node.value.value[:60])
    offenders = [o for o in offenders if o[1] not in docstrings]
    assert not offenders, offenders
    with pytest.raises(_lib.GdnError, match="no CPU fallback"):
        saved, _lib._lib = _lib._lib, None
        try:
            _lib.load(str(tmp_path / "libgandanet_sm100.so"))
        finally:
            _lib._lib = saved
