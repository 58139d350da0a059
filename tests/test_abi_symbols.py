"""The C-ABI library loads here (no GPU), exports every symbol include/csdr_b200.h declares, and refuses to create
handles without a CUDA device (no CPU fallback).  No compute calls.  CPU only."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def _declared():
    src = open(os.path.join(ROOT, "include", "csdr_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(csdr_[a-z0-9_]+)\s*\(", src)))


def test_header_and_binding_agree(cs):
    from composable_sdr_b200 import _lib
    assert set(_declared()) == set(_lib.SIGNATURES)


def test_library_exports_every_declared_symbol(cs):
    from composable_sdr_b200 import _lib
    L = C.CDLL(_lib.LIB_PATH)
    missing = [s for s in _declared() if not hasattr(L, s)]
    assert not missing, missing


def test_liquid_compat_aliases(cs):
    from composable_sdr_b200 import build
    L = C.CDLL(build.COMPAT)
    liquid = ["nco_crcf_create", "nco_crcf_set_frequency", "nco_crcf_mix_block_down", "nco_crcf_mix_block_up",
              "nco_crcf_print", "nco_crcf_destroy",
              # the whole nco_crcf family the reference can reach on such a handle (Liquid.chs:755-770, fmsPll)
              "nco_crcf_pll_set_bandwidth", "nco_crcf_pll_step", "nco_crcf_step", "nco_crcf_cexpf", "nco_crcf_get_phase",
              "nco_crcf_set_phase", "nco_crcf_adjust_frequency", "nco_crcf_adjust_phase", "nco_crcf_get_frequency",
              "nco_crcf_reset", "msresamp_crcf_create", "msresamp_crcf_print",
              "msresamp_crcf_get_rate", "msresamp_crcf_execute", "msresamp_crcf_destroy",
              "iirfilt_crcf_create_dc_blocker", "iirfilt_crcf_print", "iirfilt_crcf_execute_block",
              "iirfilt_crcf_destroy", "firpfbch_crcf_create_kaiser", "firpfbch_crcf_print",
              "firpfbch_crcf_analyzer_execute", "firpfbch_crcf_destroy", "agc_crcf_create", "agc_crcf_print",
              "agc_crcf_set_bandwidth", "agc_crcf_set_signal_level", "agc_crcf_squelch_enable",
              "agc_crcf_squelch_set_threshold", "agc_crcf_squelch_set_timeout", "agc_crcf_execute_block",
              "agc_crcf_get_rssi", "agc_crcf_squelch_get_status", "agc_crcf_destroy", "freqdem_create",
              "freqdem_print", "freqdem_demodulate_block", "freqdem_destroy", "ampmodem_create", "ampmodem_print",
              "ampmodem_demodulate_block", "ampmodem_destroy",          # the hot-path imports of Liquid.chs
              "iirfilt_rrrf_create_prototype", "iirfilt_rrrf_print", "iirfilt_rrrf_execute_block", "iirfilt_rrrf_destroy",
              "firdecim_rrrf_create_kaiser", "firdecim_rrrf_print", "firdecim_rrrf_execute_block",
              "firdecim_rrrf_destroy",                                  # + the wbFMDemodulator tail
              "firpfbch2_crcf_create_kaiser", "firpfbch2_crcf_print", "firpfbch2_crcf_execute",
              "firpfbch2_crcf_destroy"]                                 # + the oversampled analyzer
    missing = [s for s in liquid if not hasattr(L, s)]
    assert not missing, missing


def test_foreign_and_null_handles_are_rejected_not_dereferenced(cs, capfd):
    """create() returns NULL without a CUDA device: every entry point must survive that NULL (and a pointer that is
    not one of its handles) and say so through csdr_last_error().  The alias library passes foreign iirfilt_crcf
    objects (iirfilt_crcf_create_prototype stays liquid's, Liquid.chs:553-571) on to the next library in link order."""
    from composable_sdr_b200 import _lib, build
    L = _lib.load()
    foreign = C.create_string_buffer(256)          # stands in for an object made by the real libliquid
    for h in (None, C.addressof(foreign)):
        assert L.csdr_handle_kind(h) == 0
        L.csdr_agc_crcf_set_bandwidth(h, 0.1)
        assert "csdr_agc_crcf_set_bandwidth" in _lib.last_error()
        L.csdr_nco_crcf_set_frequency(h, 0.1)
        assert L.csdr_nco_crcf_get_phase(h) == 0.0
        assert L.csdr_agc_crcf_squelch_get_status(h) == -1
        assert L.csdr_chain_max_output(h, 1000) == 0
        n = C.c_size_t(7)
        assert L.csdr_chain_process(h, None, 0, 0, None, 0, C.byref(n)) == -1 and n.value == 0
        assert "handle" in _lib.last_error()
    A = C.CDLL(build.COMPAT)
    A.iirfilt_crcf_print.argtypes = [C.c_void_p]
    A.iirfilt_crcf_print(C.addressof(foreign))      # no libliquid here: reported on stderr, not executed as ours
    assert "foreign handle" in capfd.readouterr().err


def test_no_cpu_fallback(cs):
    """without a CUDA device every create() fails loudly"""
    from composable_sdr_b200 import _lib
    L = _lib.load()
    if L.csdr_device_count() > 0:
        pytest.skip("a CUDA device is present")
    assert not L.csdr_nco_crcf_create(1)
    assert "no CUDA device" in _lib.last_error()
    with pytest.raises(cs.CsdrError):
        cs.Chain(2.56e6, 1e5, 200e3)


def test_package_never_imports_the_oracle():
    """the product path must not link, import or call anything under oracle/"""
    pkg = os.path.join(ROOT, "composable-sdr_b200")
    banned = ("liquid_oracle", "liboracle", "orc_", "import oracle", "from oracle", "oracle.oracle", "cuda_emu", "libemu")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".inl", ".c", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                if f == "platform.cuh":
                    continue      # names the emulator header, but only under #ifdef CSDR_EMU (test build)
                hits = [b for b in banned if b in text]
                assert not hits, (f, hits)
