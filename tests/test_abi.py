"""CPU-side checks of the boundary: libnic.so loads, exports every symbol include/nic.h declares, refuses to run
without an sm_100 device, and the host mirror keeps the reference's names and configuration semantics.
No compute calls (there is no GPU in the build container)."""
import ctypes
import os
import re

import pytest
import torch

import neural_image_compression_v2_b200 as nic
from neural_image_compression_v2_b200 import _lib as L
from neural_image_compression_v2_b200 import fp_def, image_compression, models, utils, var2

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "nic.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(nic_[a-z0-9_]+)\s*\(", src)))


def test_header_and_binding_agree():
    syms = declared_symbols()
    assert len(syms) >= 20
    assert sorted(L.SYMBOLS) == syms          # the ctypes table binds exactly what the header declares


def test_exchange_mode_constants_match_the_header():
    src = open(os.path.join(ROOT, "include", "nic.h")).read()
    got = {k: int(v) for k, v in re.findall(r"#define\s+(NIC_EXCHANGE_[A-Z_]+)\s+(\d+)", src)}
    assert got == {"NIC_EXCHANGE_ONE_SHOT": L.EXCHANGE_ONE_SHOT, "NIC_EXCHANGE_SLICED": L.EXCHANGE_SLICED}
    assert int(re.search(r"#define\s+NIC_MAX_PEERS\s+(\d+)", src).group(1)) == L.MAX_PEERS


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(L.LIB_PATH)
    for s in declared_symbols():
        assert hasattr(lib, s), s
    assert nic.load_library().nic_abi_version() == 1


def test_header_is_plain_c_and_struct_layouts_match_the_binding(tmp_path):
    """include/nic.h compiles as C99 on its own (the boundary has no C++ or torch types) and every struct it declares has
    the size and field offsets the ctypes binding assumes."""
    import shutil
    import subprocess
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no gcc")
    structs = {"NicGeom": L.NicGeom, "NicMlp": L.NicMlp, "NicMlpGrad": L.NicMlpGrad, "NicAdamTensor": L.NicAdamTensor,
               "NicExchange": L.NicExchange}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "nic.h"', 'int main(void) {']
    for name, st in structs.items():
        lines.append(f'  printf("{name} size %zu\\n", sizeof({name}));')
        for field, _ in st._fields_:
            lines.append(f'  printf("{name} {field} %zu\\n", offsetof({name}, {field}));')
    lines += ['  return 0;', '}']
    src = tmp_path / "abi.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "abi"
    subprocess.run([gcc, "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split("\n")
    got = {tuple(l.split()[:2]): int(l.split()[2]) for l in out if l.strip()}
    for name, st in structs.items():
        assert got[(name, "size")] == ctypes.sizeof(st), name
        for field, _ in st._fields_:
            assert got[(name, field)] == getattr(st, field).offset, (name, field)


def test_no_cpu_fallback():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(nic.NicError) as e:
        L.handle("cuda:0")
    assert e.value.status == -3
    with pytest.raises(nic.NicError):
        models.quantize4fp(torch.zeros(4), 8)              # CPU tensor: refused, not computed on the host
    dec = image_compression.ColorDecoder(in_channels=73, hidden=64, out_channels=3)
    with pytest.raises(nic.NicError):
        dec(torch.zeros(2, 73))


def test_status_strings():
    lib = nic.load_library()
    assert lib.nic_status_string(0) == b"ok"
    assert b"sm_100" in lib.nic_status_string(-3)
    assert lib.nic_cin(None) == -1


def test_geometry_descriptor():
    g0, g1 = torch.zeros(12, 129, 129), torch.zeros(12, 65, 65)
    geom = L.make_geom(L.METHOD_2D, g0, g1, 256, 8, -2, 0, 6, L.PE_TRIANGULAR)
    assert L.cin_of(geom) == 73 and list(geom.g0_nodes)[:2] == [129, 129] and geom.num_blocks == 8
    g0, g1 = torch.zeros(12, 5, 9, 17), torch.zeros(12, 3, 5, 9)          # [C, z, y, x]
    geom = L.make_geom(L.METHOD_3D, g0, g1, (4, 2, 1), 1, -2, 0, 6, L.PE_TRIANGULAR)
    assert list(geom.g0_nodes) == [17, 9, 5] and list(geom.block) == [4, 2, 1] and L.cin_of(geom) == 127
    geom = L.make_geom(L.METHOD_3D_V2, g0, g1, 4, 1, -2, 0, 6, L.PE_SINUSOIDAL)
    assert L.cin_of(geom) == 79 and abs(geom.pe_div[1] - 0.0464158877) < 1e-7
    with pytest.raises(ValueError):
        L.make_geom(L.METHOD_2D, g0, g1, 4, 1, -2, 0, 6, 0)


def test_config_mirror_defaults_and_overrides():
    var2.update()
    assert (var2.IMAGE_SIZE, var2.FEATURE_PYRAMID_SIZE, var2.DECODER_INPUT_CHANNELS, var2.MAX_MIP_LEVEL) == (512, 128, 73, 0)
    assert var2.CROP_SIZE == 256 and var2.MLP_DTYPE == torch.float32 and var2.FP_DIMENSION == 2
    var2.update("IMAGE_DIMENSION=3", "COMPRESSION_METHOD=3", "TF_NO_MIP=0", "IMAGE_SIZE=64", "CROP_MIP_LEVEL=5")
    assert (var2.DECODER_INPUT_CHANNELS, var2.MAX_MIP_LEVEL, var2.CROP_SIZE) == (127, 9, 32)
    var2.update(IMAGE_DIMENSION=3, COMPRESSION_METHOD=4)
    assert var2.DECODER_INPUT_CHANNELS == 79
    var2.update(COMPRESSION_METHOD=2, IMAGE_DIMENSION=3)
    assert var2.FP_DIMENSION == 2 and var2.DECODER_INPUT_CHANNELS == 73
    with pytest.raises(ValueError):
        var2.update("TF_NO_MIP=maybe")
    with pytest.raises(KeyError):
        var2.update(NOT_A_KEY=1)
    var2.update()


def test_level_tables_match_golden():
    import json
    t = json.load(open(os.path.join(ROOT, "tests", "golden", "tables.json")))
    for s in (16, 64, 512, 4096):
        assert fp_def.return_pyramid_levels(s // 4) == t[str(s)]["levels"]
        assert {str(k): v for k, v in fp_def.create_pyramid_mip_levels(s, s // 4).items()} == t[str(s)]["table"]


def test_reference_names_present():
    for name in ("create_decoder_input_2d", "create_decoder_input_3d", "create_decoder_input_3d_v2",
                 "finally_decode_input_2d", "finally_decode_input_3d", "finally_decode_input_3d_v2", "ColorDecoder",
                 "decode_image", "random_crop_dataset"):
        assert hasattr(image_compression, name), name
    for name in ("create_pyramid", "create_pyramid_3d", "create_pyramid_mip_levels", "fp_quantize_clamp",
                 "fp_all_quantize", "fp_savable", "fp_load", "fp_freeze", "return_pyramid_levels"):
        assert hasattr(fp_def, name), name
    for name in ("quantize4fp", "save4fp", "load4fp", "quantize_to_bit", "quantize_clamp"):
        assert hasattr(models, name), name
    for name in ("triangular_positional_encoding", "positional_encoding", "tri", "calculate_psnr", "bits2dtype_torch"):
        assert hasattr(utils, name), name
    dec = image_compression.ColorDecoder(in_channels=73, hidden=64, out_channels=3)
    assert sorted(dec.state_dict()) == ["decoder.0.bias", "decoder.0.weight", "decoder.2.bias", "decoder.2.weight",
                                        "decoder.4.bias", "decoder.4.weight"]
    assert tuple(dec.state_dict()["decoder.0.weight"].shape) == (64, 73)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "neural_image_compression_v2_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in txt.replace("nic_oracle", "oracle") or "import oracle" not in txt, f
                assert not re.search(r"^\s*(from|import)\s+oracle", txt, flags=re.M), f
