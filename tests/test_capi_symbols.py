"""The C-ABI library loads without a GPU, exports every symbol include/sla_b200.h declares, and has
no CPU path: without a CUDA device handle creation fails loudly instead of falling back."""
import ctypes as C
import os
import re
import subprocess

import pytest

from conftest import PRODUCT_SO, ROOT, _has_gpu
from sla_b200 import capi


def declared_functions():
    text = open(os.path.join(ROOT, "include", "sla_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = re.findall(r"\b((?:SLA|SLAB200_)\w+)\s*\(", text)
    return sorted({n for n in names if not n.startswith("SLA_Calculate")})


@pytest.fixture(scope="module")
def built():
    subprocess.run(["make", "-s", "-C", ROOT, "product"], check=True)
    assert os.path.exists(PRODUCT_SO)
    return C.CDLL(PRODUCT_SO)


def test_every_declared_symbol_is_exported(built):
    names = declared_functions()
    assert len(names) >= 30
    missing = [n for n in names if not hasattr(built, n)]
    assert missing == []


def test_reference_api_symbols_present(built):
    # what the unmodified reference CLI links against (src/main.c)
    for n in ["SLAEncoder_Create", "SLAEncoder_Destroy", "SLAEncoder_SetWaveFormat", "SLAEncoder_SetEncodeParameter",
              "SLAEncoder_EncodeHeader", "SLAEncoder_EncodeBlock", "SLAEncoder_EncodeWhole", "SLADecoder_DecodeHeader",
              "SLADecoder_Create", "SLADecoder_Destroy", "SLADecoder_SetWaveFormat", "SLADecoder_SetEncodeParameter",
              "SLADecoder_DecodeWhole", "SLAStreamingDecoder_Create", "SLAStreamingDecoder_Decode"]:
        assert hasattr(built, n), n


def test_shim_headers_compile_reference_style_program(tmp_path):
    src = tmp_path / "t.c"
    src.write_text('#include "SLAEncoder.h"\n#include "SLADecoder.h"\n'
                   'int main(void){struct SLAEncoderConfig c; struct SLAHeaderInfo h; (void)c; (void)h;'
                   ' return SLA_HEADER_SIZE == 43 && SLA_APIRESULT_PARAMETER_NOT_SET == 15 ? 0 : 1;}\n')
    exe = tmp_path / "t"
    subprocess.run(["gcc", "-std=c89", "-pedantic", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    assert subprocess.run([str(exe)]).returncode == 0


def test_struct_layout_matches_reference_abi():
    assert C.sizeof(capi.WaveFormat) == 16 and C.sizeof(capi.EncodeParameter) == 24
    assert C.sizeof(capi.HeaderInfo) == 56 and C.sizeof(capi.EncoderConfig) == 24 and C.sizeof(capi.DecoderConfig) == 24


@pytest.mark.skipif(_has_gpu(), reason="only meaningful on a machine without a CUDA device")
def test_no_cpu_fallback_without_gpu(built):
    built.SLAEncoder_Create.restype = C.c_void_p
    built.SLADecoder_Create.restype = C.c_void_p
    built.SLAB200_LastError.restype = C.c_char_p
    cfg = capi.EncoderConfig(**capi.CLI_CAPACITY, verpose_flag=0)
    assert built.SLAEncoder_Create(C.byref(cfg)) is None
    assert b"no CUDA device" in built.SLAB200_LastError()
    dcfg = capi.DecoderConfig(**capi.CLI_CAPACITY, enable_crc_check=1, verpose_flag=0)
    assert built.SLADecoder_Create(C.byref(dcfg)) is None
