"""CPU-only checks of the drop-in boundary: the C-ABI library loads and exports
exactly what include/phaserot_cuda.h declares, argument validation works without
a GPU, the product refuses to run without a device (no CPU fallback), and the
host programs keep the reference's command line behaviour."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

import oracle_lib as O
from phaserotate.lv2_b200 import build, capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.fixture(scope="module")
def lib():
    build.build_library()
    return capi.load()


def test_library_exports_every_declared_symbol(lib):
    hdr = open(os.path.join(ROOT, "include", "phaserot_cuda.h")).read()
    declared = set(re.findall(r"PHASEROT_API\s+[\w\s\*]+?\b(phaserot_\w+)\s*\(", hdr))
    assert len(declared) >= 25
    missing = [s for s in sorted(declared) if not hasattr(lib, s)]
    assert not missing, missing
    assert declared == set(capi.SYMBOLS), declared ^ set(capi.SYMBOLS)
    # the header cites the reference interface each entry point replaces
    assert hdr.count("Replaces:") >= 8 and "cli:" in hdr and "src:" in hdr


def test_abi_version_and_error_strings(lib):
    assert lib.phaserot_abi_version() == capi.ABI_VERSION
    for code in (0, -1, -2, -3, -4, -5, -6):
        assert lib.phaserot_strerror(code)
    assert b"no CPU fallback" in lib.phaserot_strerror(capi.E_NO_DEVICE)


def test_create_validates_arguments(lib):
    h = C.c_void_p()
    bad = [
        capi.Cfg(99, 0, 2, 8192, 48000.0, 2, -1, 0),        # ABI version
        capi.Cfg(1, 7, 2, 8192, 48000.0, 2, -1, 0),         # mode
        capi.Cfg(1, 0, 0, 8192, 48000.0, 2, -1, 0),         # channels
        capi.Cfg(1, 0, 2, 1000, 48000.0, 2, -1, 0),         # blksiz not a power of two (cli:749-755)
        capi.Cfg(1, 0, 2, 65536, 48000.0, 2, -1, 0),        # blksiz > 32768
        capi.Cfg(1, 1, 3, 0, 48000.0, 2, -1, 0),            # plugin has at most 2 channels (src/phaserotate.h:97)
        capi.Cfg(1, 1, 1, 0, 0.0, 2, -1, 0),                # sample rate
    ]
    for cfg in bad:
        assert lib.phaserot_create(C.byref(h), C.byref(cfg)) == capi.E_INVAL
        assert not h.value
    assert lib.phaserot_create(None, None) == capi.E_INVAL
    lib.phaserot_destroy(None)  # harmless
    assert lib.phaserot_reset(None) == capi.E_INVAL
    assert lib.phaserot_latency(None) == 0


@pytest.mark.skipif(_has_gpu(), reason="only meaningful on a machine without a GPU")
def test_no_cpu_fallback(lib):
    with pytest.raises(capi.PhaserotError) as e:
        capi.Phaserot(n_channels=2, blksiz=8192)
    assert e.value.code == capi.E_NO_DEVICE
    with pytest.raises(capi.PhaserotError):
        capi.Phaserot(mode=capi.MODE_PLUGIN, n_channels=1, sample_rate=48000)


def _cli(*argv):
    build.build_host()
    return subprocess.run([os.path.join(build.BIN_DIR, "phase-rotate")] + list(argv), capture_output=True, text=True)


def test_cli_argument_errors_match_reference(tmp_path):
    wav = str(tmp_path / "x.wav")
    O.write_wav_f32(wav, O.two_sine(48000, 0.1, 2), 48000)
    cases = [(["-s", "7", wav], "Error: 180 deg is not evenly dividable by given stride.\n"),
             (["-s", "0", wav], "Error: 180 deg is not evenly dividable by given stride.\n"),
             (["-f", "100", wav], "Error: fft-len is out of bounds; valid range 1024..32768\n"),
             (["-a", "10", wav], "Error: -a, --angle option requires an output file to be given.\n"),
             ([], "Error: Missing parameter. See --help for usage information.\n"),
             (["-x", wav], None)]
    ref = os.path.join(O.REF_DIR, "phase-rotate")
    for argv, msg in cases:
        r = _cli(*argv)
        assert r.returncode == 1, argv
        if msg is not None:
            assert r.stderr == msg, (argv, r.stderr)
        if os.path.exists(ref):  # same text and exit code as the reference binary
            rr = subprocess.run([ref] + argv, capture_output=True, text=True)
            assert rr.returncode == r.returncode
            assert rr.stderr.replace(ref, "phase-rotate") == r.stderr.replace(os.path.join(build.BIN_DIR, "phase-rotate"), "phase-rotate"), argv
    r = _cli("--help")
    assert r.returncode == 0 and r.stdout.startswith("phase-rotate - Audio File Phase Rotation Util.")
    if os.path.exists(ref):
        assert subprocess.run([ref, "--help"], capture_output=True, text=True).stdout == r.stdout
    r = _cli(str(tmp_path / "missing.wav"))
    assert r.returncode == 1 and r.stderr.startswith("Cannot open '")
    assert _cli("-V").stdout.startswith("phase-rotate version ")


@pytest.mark.skipif(_has_gpu(), reason="only meaningful on a machine without a GPU")
def test_cli_fails_loudly_without_gpu(tmp_path):
    wav = str(tmp_path / "x.wav")
    O.write_wav_f32(wav, O.two_sine(48000, 0.1, 2), 48000)
    r = _cli(wav)
    assert r.returncode == 1 and "CUDA backend" in r.stderr and r.stdout == ""


def test_lv2_binary_exports_descriptor():
    build.build_host()
    so = C.CDLL(os.path.join(build.BIN_DIR, "phaserotate_cuda.so"))

    class Desc(C.Structure):
        _fields_ = [("URI", C.c_char_p)] + [(n, C.c_void_p) for n in ("instantiate", "connect_port", "activate", "run", "deactivate", "cleanup", "extension_data")]

    so.lv2_descriptor.restype = C.POINTER(Desc)
    so.lv2_descriptor.argtypes = [C.c_uint32]
    d0 = so.lv2_descriptor(0).contents
    assert d0.URI == b"http://gareus.org/oss/lv2/phaserotate"        # src/phaserotate.h:21
    assert all(getattr(d0, n) for n in ("instantiate", "connect_port", "activate", "run", "cleanup", "extension_data"))
    assert d0.deactivate is None                                       # src/phaserotate.c:866
    assert so.lv2_descriptor(1).contents.URI == b"http://gareus.org/oss/lv2/phaserotate#stereo"
    assert not so.lv2_descriptor(2)
    if not _has_gpu():
        # instantiate must return NULL when the backend cannot be created (no CPU fallback)
        lat = C.c_float(-1)
        dt = O.lv2_harness().lv2h_render(os.path.join(build.BIN_DIR, "phaserotate_cuda.so").encode(), 0, 48000.0, 1,
                                         np.zeros(64, np.float32), np.zeros(64, np.float32), 64, 64, np.zeros(1, np.float32), 0, C.byref(lat))
        assert dt == -4
