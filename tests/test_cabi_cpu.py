"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol that
include/ddp_b200.h declares (no compute calls without a GPU)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "ddp_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ddp_[a-z_]+)\s*\(", text)))


def test_header_declares_expected_entry_points():
    syms = _declared_symbols()
    for s in ("ddp_create", "ddp_destroy", "ddp_set_weight", "ddp_commit_weights", "ddp_plan", "ddp_sample",
              "ddp_sample_host", "ddp_set_schedule", "ddp_last_error"):
        assert s in syms


def test_header_is_plain_c(tmp_path):
    """The boundary is a C ABI: include/ddp_b200.h must compile as C99 (what a cgo / FFI binding would include)."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("needs gcc")
    src = tmp_path / "hdr.c"
    src.write_text('#include "ddp_b200.h"\nint main(void) { ddp_config c; ddp_neck_config n; ddp_bev_config b; '
                   '(void)c; (void)n; (void)b; return DDP_ABI_VERSION - 1; }\n')
    res = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only", "-I",
                          os.path.join(ROOT, "include"), str(src)], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr


def test_library_exports_every_declared_symbol():
    from ddp_b200 import build, _lib
    build.build()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for s in _declared_symbols():
        assert hasattr(lib, s), f"{s} declared in include/ddp_b200.h but not exported"
    assert set(_lib.EXPORTS) == set(_declared_symbols())
    assert lib.ddp_abi_version() == _lib.ABI_VERSION


def test_binding_struct_matches_header():
    from ddp_b200 import _lib
    text = open(os.path.join(ROOT, "include", "ddp_b200.h")).read()
    body = re.search(r"typedef struct ddp_config \{(.*?)\} ddp_config;", text, flags=re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    fields = re.findall(r"(int32_t|float)\s+([a-z_]+);", body)
    assert [f for _, f in fields] == [n for n, _ in _lib.DDPConfig._fields_]
    for (ty, _), (_, cty) in zip(fields, _lib.DDPConfig._fields_):
        assert cty is (ctypes.c_int32 if ty == "int32_t" else ctypes.c_float)


def test_neck_binding_struct_matches_header():
    from ddp_b200 import _lib
    text = open(os.path.join(ROOT, "include", "ddp_b200.h")).read()
    body = re.search(r"typedef struct ddp_neck_config \{(.*?)\} ddp_neck_config;", text, flags=re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    fields = re.findall(r"(int32_t|float)\s+([a-z_]+)(\[\d+\])?;", body)
    assert [f for _, f, _ in fields] == [n for n, _ in _lib.DDPNeckConfig._fields_]
    for (ty, _, arr), (_, cty) in zip(fields, _lib.DDPNeckConfig._fields_):
        base = ctypes.c_int32 if ty == "int32_t" else ctypes.c_float
        assert cty is base if not arr else (cty._type_ is base and cty._length_ == int(arr[1:-1]))


def test_bev_binding_struct_matches_header():
    from ddp_b200 import _lib
    text = open(os.path.join(ROOT, "include", "ddp_b200.h")).read()
    body = re.search(r"typedef struct ddp_bev_config \{(.*?)\} ddp_bev_config;", text, flags=re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    fields = re.findall(r"(int32_t|float)\s+([a-z_]+);", body)
    assert [f for _, f in fields] == [n for n, _ in _lib.DDPBevConfig._fields_]
    for (ty, _), (_, cty) in zip(fields, _lib.DDPBevConfig._fields_):
        assert cty is (ctypes.c_int32 if ty == "int32_t" else ctypes.c_float)


def test_no_gpu_means_loud_failure():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from ddp_b200 import DecodeEngine
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        DecodeEngine()


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "ddp_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle|oracle/", src, flags=re.M), \
                    f"{f} reaches into oracle/"


def test_host_schedule_matches_oracle_schedule():
    """ddp_b200/schedule.py (host scalars handed to the library) == oracle == reference formulas."""
    import torch
    from ddp_b200 import schedule as S
    from oracle import ddp_oracle as O
    for T in (1, 3, 10):
        l, a, s, an, sn = S.seg_schedule(T, 1, (0, 0.999), "cosine")
        cfg = O.OracleConfig(timesteps=T)
        for k, (t_now, t_next) in enumerate(O.time_pairs_seg(cfg)):
            tt = torch.tensor([t_now, t_next])
            ln = O.log_snr_cosine(tt[0:1]); lx = O.log_snr_cosine(tt[1:2])
            assert float(ln) == l[k]
            assert float(O.alpha_sigma(ln)[0]) == a[k] and float(O.alpha_sigma(ln)[1]) == s[k]
            assert float(O.alpha_sigma(lx)[0]) == an[k] and float(O.alpha_sigma(lx)[1]) == sn[k]
    t, g, gn = S.depth_schedule(20, 1)
    pairs = O.time_pairs_depth(O.OracleConfig(task="depth", timesteps=20))
    for k, (t_now, t_next) in enumerate(pairs):
        tt = torch.tensor([t_now, t_next])
        assert float(O.gamma_depth(tt[0:1])) == g[k] and float(O.gamma_depth(tt[1:2])) == gn[k]
    with pytest.raises(ValueError):
        S.seg_schedule(3, 1, (0, 0.999), "quadratic")


def test_synthetic_recipe_equals_oracle_recipe():
    import torch
    from ddp_b200 import synthetic as S
    from oracle import ddp_oracle as O
    for task in ("seg", "depth"):
        cfg = O.OracleConfig(task=task, num_classes=19, randsteps=2)
        a = O.make_weights(cfg, seed=7)
        b = S.make_weights(task=task, num_classes=19, seed=7)
        assert a.keys() == b.keys() and all(torch.equal(a[k], b[k]) for k in a)
        xa, na = O.make_inputs(cfg, 2, 5, 6, seed=3)
        xb, nb = S.make_inputs(task, 2, 2, 5, 6, seed=3)
        assert torch.equal(xa, xb) and torch.equal(na, nb)


def test_synthetic_neck_and_bev_recipes_equal_oracle_recipes():
    import torch
    from ddp_b200 import synthetic as S
    from oracle import bev_oracle as BO, neck_oracle as NO
    a, b = NO.make_weights([96, 192, 384, 768], seed=5), S.make_neck_weights([96, 192, 384, 768], seed=5)
    assert a.keys() == b.keys() and all(torch.equal(a[k], b[k]) for k in a)
    for feat in (256, 512):
        a = BO.make_weights(BO.BevConfig(feat_channels=feat), seed=6)
        b = S.make_bev_weights(feat_channels=feat, num_layers=5, seed=6)
        assert a.keys() == b.keys() and all(torch.equal(a[k], b[k]) for k in a)


def test_ddpm_host_schedule_matches_oracle_formulas():
    import torch
    from ddp_b200 import schedule as S
    from oracle import ddp_oracle as O
    omc, c, std, on = S.seg_ddpm_schedule(4, 1, (0, 0.999), "cosine")
    cfg = O.OracleConfig(timesteps=4)
    for k, (t_now, t_next) in enumerate(O.time_pairs_seg(cfg)):
        tt = torch.tensor([t_now, t_next])
        ln, lx = O.log_snr_cosine(tt[0:1]), O.log_snr_cosine(tt[1:2])
        cc = -torch.special.expm1(ln - lx)
        _, sn = O.alpha_sigma(lx)
        assert float(cc) == c[k] and float(1 - cc) == omc[k]
        assert float((0.5 * torch.log(((sn ** 2) * cc).clamp(min=1e-20))).exp()) == std[k]
        assert on[k] == int(t_next > 0)
