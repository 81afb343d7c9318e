"""CPU-side checks of the drop-in boundary: the shared library loads without a GPU, exports every symbol that
include/dcvgan_b200.h declares, the ctypes binding lists exactly those symbols, argument validation fails loudly
before any CUDA call, and the product package never imports the oracle."""
import ctypes
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


def _declared():
    hdr = (ROOT / "include" / "dcvgan_b200.h").read_text()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(dcv_[a-z0-9_]+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol():
    from dcvgan_b200 import _lib
    names = _declared()
    assert len(names) >= 35
    assert sorted(_lib.SIGNATURES) == names, (set(names) ^ set(_lib.SIGNATURES))
    handle = ctypes.CDLL(str(_lib.LIB_PATH))
    for n in names:
        assert getattr(handle, n) is not None
    assert _lib.lib().dcv_abi_version() == _lib.ABI_VERSION == 5


def test_geom_struct_matches_header():
    from dcvgan_b200._lib import Geom
    hdr = (ROOT / "include" / "dcvgan_b200.h").read_text()
    body = re.search(r"typedef struct dcv_geom \{(.*?)\} dcv_geom;", hdr, flags=re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    fields = [f.strip() for decl in re.findall(r"int32_t ([^;]+);", body) for f in decl.split(",")]
    assert fields == [n for n, _ in Geom._fields_]
    assert ctypes.sizeof(Geom) == 4 * len(fields)


def test_validation_errors_are_reported_without_a_gpu():
    from dcvgan_b200 import _lib
    L = _lib.lib()
    bad = _lib.Geom(2, 1, 16, 16, 8, 1, 9, 8, 8, 1, 4, 4, 1, 2, 2, 0, 1, 1, 0, 0)      # Hs inconsistent with Hl
    buf = ctypes.create_string_buffer(64)
    rc = L.dcv_conv(ctypes.byref(bad), 0, 0, 0, ctypes.addressof(buf), 8, ctypes.addressof(buf), ctypes.addressof(buf), 8, 0, 0.0, None)
    assert rc < 0 and b"inconsistent" in L.dcv_last_error()
    rc = L.dcv_loss_fwd_bwd(0, ctypes.addressof(buf), 1, 0, 0, ctypes.addressof(buf), 0, None, 1, 1.0, None)
    assert rc < 0 and b"empty" in L.dcv_last_error()
    with pytest.raises(_lib.DcvError):
        _lib.check(rc)


def test_no_cpu_fallback_and_no_oracle_in_the_product():
    import torch
    import dcvgan_b200
    from dcvgan_b200 import generator
    for f in (ROOT / "dcvgan_b200").glob("*.py"):
        src = f.read_text()
        assert "import oracle" not in src and "from oracle" not in src, f
    if not torch.cuda.is_available():
        g = generator.GeometricVideoGenerator(40, 10, 1, "depth", 8, 16)
        with pytest.raises(dcvgan_b200.DcvError):
            g.sample_videos(2)


def test_normaliser_accepts_every_reference_yaml():
    """The five BASELINE configs (and the other seven YAMLs) normalise to the schema train.py reads at HEAD."""
    import yaml
    from oracle import dcvgan_oracle as orc
    cfg_dir = Path("/root/reference/config")
    if not cfg_dir.exists():
        pytest.skip("reference tree not mounted on this box")
    seen = 0
    for f in sorted(cfg_dir.glob("*.yml")):
        cfg = orc.normalise_config(yaml.safe_load(f.read_text()))
        for k in ("ggen", "cgen", "idis", "vdis", "gdis", "loss", "num_gen_update", "num_dis_update", "evaluation"):
            assert k in cfg, (f.name, k)
        assert cfg["geometric_info"]["channel"] in (1, 2, 25)
        assert cfg["ggen"]["ngf"] in (64, 96) and "dim_z_color" in cfg["cgen"]
        seen += 1
    assert seen == 12
