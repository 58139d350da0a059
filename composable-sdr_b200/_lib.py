"""ctypes binding of libcsdr_b200.so (include/csdr_b200.h).  No fallback: a missing library or a missing CUDA
device is an error."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libcsdr_b200.so")


class ChainCfg(C.Structure):
    _fields_ = [("samplerate", C.c_double), ("offset_hz", C.c_double), ("bandwidth_hz", C.c_double),
                ("demod", C.c_int), ("kf", C.c_float), ("agc_thresh_db", C.c_float),
                ("channels", C.c_uint), ("mix", C.c_int), ("nstreams", C.c_uint), ("device", C.c_int),
                ("decim", C.c_uint), ("channelizer", C.c_int)]


# every symbol include/csdr_b200.h declares: name -> (restype, argtypes)
_vp, _u, _f, _i, _sz = C.c_void_p, C.c_uint, C.c_float, C.c_int, C.c_size_t
SIGNATURES = {
    "csdr_version": (C.c_char_p, []), "csdr_last_error": (C.c_char_p, []),
    "csdr_device_count": (_i, []), "csdr_set_device": (_i, [_i]),
    "csdr_host_alloc": (_vp, [_sz]), "csdr_host_free": (None, [_vp]),
    "csdr_kernel_launches": (C.c_uint64, []), "csdr_synchronize": (_i, []),
    "csdr_set_option": (_i, [_i, _i]), "csdr_get_option": (_i, [_i]), "csdr_handle_kind": (_i, [_vp]),
    "csdr_nco_crcf_adjust_frequency": (None, [_vp, _f]), "csdr_nco_crcf_adjust_phase": (None, [_vp, _f]),
    "csdr_nco_crcf_step": (None, [_vp]), "csdr_nco_crcf_reset": (None, [_vp]), "csdr_nco_crcf_get_phase": (_f, [_vp]),
    "csdr_nco_crcf_get_frequency": (_f, [_vp]), "csdr_nco_crcf_cexpf": (None, [_vp, _vp]),
    "csdr_nco_crcf_pll_set_bandwidth": (None, [_vp, _f]), "csdr_nco_crcf_pll_step": (None, [_vp, _f]),
    "csdr_nco_crcf_create": (_vp, [_i]), "csdr_nco_crcf_destroy": (None, [_vp]), "csdr_nco_crcf_print": (None, [_vp]),
    "csdr_nco_crcf_set_frequency": (None, [_vp, _f]), "csdr_nco_crcf_set_phase": (None, [_vp, _f]),
    "csdr_nco_crcf_get_phase_word": (C.c_uint32, [_vp]), "csdr_nco_crcf_get_freq_word": (C.c_uint32, [_vp]),
    "csdr_nco_crcf_mix_block_down": (None, [_vp, _vp, _vp, _u]), "csdr_nco_crcf_mix_block_up": (None, [_vp, _vp, _vp, _u]),
    "csdr_msresamp_crcf_create": (_vp, [_f, _f]), "csdr_msresamp_crcf_destroy": (None, [_vp]),
    "csdr_msresamp_crcf_print": (None, [_vp]), "csdr_msresamp_crcf_get_rate": (_f, [_vp]),
    "csdr_msresamp_crcf_execute": (None, [_vp, _vp, _u, _vp, C.POINTER(_u)]),
    "csdr_msresamp_num_stages": (_u, [_vp]), "csdr_msresamp_stage_m": (_u, [_vp, _u]),
    "csdr_msresamp_stage_taps": (_i, [_vp, _u, _vp]), "csdr_msresamp_resamp_step": (C.c_uint32, [_vp]),
    "csdr_msresamp_resamp_bank": (_i, [_vp, _vp, C.POINTER(_u)]),
    "csdr_iirfilt_crcf_create_dc_blocker": (_vp, [_f]), "csdr_iirfilt_crcf_destroy": (None, [_vp]),
    "csdr_iirfilt_crcf_print": (None, [_vp]), "csdr_iirfilt_crcf_execute_block": (None, [_vp, _vp, _u, _vp]),
    "csdr_firpfbch_crcf_create_kaiser": (_vp, [_i, _u, _u, _f]), "csdr_firpfbch_crcf_destroy": (None, [_vp]),
    "csdr_firpfbch_crcf_print": (None, [_vp]), "csdr_firpfbch_crcf_analyzer_execute": (None, [_vp, _vp, _vp]),
    "csdr_firpfbch_taps": (_i, [_vp, _vp]), "csdr_firpfbch_execute_block": (_i, [_vp, _vp, _vp, _u, _vp]),
    "csdr_agc_crcf_create": (_vp, []), "csdr_agc_crcf_destroy": (None, [_vp]), "csdr_agc_crcf_print": (None, [_vp]),
    "csdr_agc_crcf_set_bandwidth": (None, [_vp, _f]), "csdr_agc_crcf_set_signal_level": (None, [_vp, _f]),
    "csdr_agc_crcf_squelch_enable": (None, [_vp]), "csdr_agc_crcf_squelch_set_threshold": (None, [_vp, _f]),
    "csdr_agc_crcf_squelch_set_timeout": (None, [_vp, _u]), "csdr_agc_crcf_execute_block": (None, [_vp, _vp, _u, _vp]),
    "csdr_agc_crcf_get_rssi": (_f, [_vp]), "csdr_agc_crcf_squelch_get_status": (_i, [_vp]),
    "csdr_agc_squelch_execute_block": (_i, [_vp, _vp, _u, _vp]),
    "csdr_freqdem_create": (_vp, [_f]), "csdr_freqdem_destroy": (None, [_vp]), "csdr_freqdem_print": (None, [_vp]),
    "csdr_freqdem_demodulate_block": (None, [_vp, _vp, _u, _vp]),
    "csdr_ampmodem_create": (_vp, [_f, _i, _i]), "csdr_ampmodem_destroy": (None, [_vp]),
    "csdr_ampmodem_print": (None, [_vp]), "csdr_ampmodem_demodulate_block": (None, [_vp, _vp, _u, _vp]),
    "csdr_firpfbch2_crcf_create_kaiser": (_vp, [_i, _u, _u, _f]), "csdr_firpfbch2_crcf_destroy": (None, [_vp]),
    "csdr_firpfbch2_crcf_print": (None, [_vp]), "csdr_firpfbch2_crcf_execute": (None, [_vp, _vp, _vp]),
    "csdr_firpfbch2_taps": (_i, [_vp, _vp]), "csdr_firpfbch2_execute_block": (_i, [_vp, _vp, _u, _vp]),
    "csdr_iirfilt_rrrf_create_prototype": (_vp, [_i, _i, _i, _u, _f, _f, _f, _f]), "csdr_iirfilt_rrrf_destroy": (None, [_vp]),
    "csdr_iirfilt_rrrf_print": (None, [_vp]), "csdr_iirfilt_rrrf_execute_block": (None, [_vp, _vp, _u, _vp]),
    "csdr_iirfilt_rrrf_coefficients": (_u, [_vp, _vp, _vp]),
    "csdr_firdecim_rrrf_create_kaiser": (_vp, [_u, _u, _f]), "csdr_firdecim_rrrf_destroy": (None, [_vp]),
    "csdr_firdecim_rrrf_print": (None, [_vp]), "csdr_firdecim_rrrf_execute_block": (None, [_vp, _vp, _u, _vp]),
    "csdr_chain_create": (_vp, [C.POINTER(ChainCfg)]), "csdr_chain_destroy": (_i, [_vp]),
    "csdr_chain_print": (None, [_vp]), "csdr_chain_num_outputs": (_u, [_vp]),
    "csdr_chain_out_elem_size": (_sz, [_vp]), "csdr_chain_max_output": (_sz, [_vp, _sz]),
    "csdr_chain_process": (_i, [_vp, _vp, _sz, _sz, C.POINTER(_vp), _sz, C.POINTER(_sz)]),
    "csdr_chain_run_file": (_i, [_vp, C.c_char_p, C.c_char_p, C.c_uint64, _sz, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "csdr_chain_seek": (_i, [_vp, C.c_uint64]), "csdr_chain_warmup_len": (_sz, [_vp]),
    "csdr_chain_cuda_stream": (_vp, [_vp]), "csdr_chain_profile": (_i, [_vp, _i]),
    "csdr_chain_frontend_ms": (C.c_double, [_vp, C.POINTER(C.c_uint64)]), "csdr_chain_agc_fixups": (C.c_uint64, [_vp]),
    "csdr_chain_agc_counters": (_i, [_vp, C.POINTER(C.c_uint64)]),
    "csdr_chain_agc_plan": (_i, [_vp, C.POINTER(C.c_int)]),
}

OPT_VCO_DIRECT, OPT_AMPMODEM_PLL, OPT_RESAMP_FC_OLD, OPT_AGC_SEGMENT, OPT_AGC_WARMUP = 0, 1, 2, 3, 4

_lib = None


def load():
    """Load libcsdr_b200.so and declare every prototype.  Raises if the library has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -m composable_sdr_b200.build` (needs nvcc). "
            "There is no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(L, name)
        fn.restype, fn.argtypes = res, args
    _lib = L
    return L


def last_error():
    return load().csdr_last_error().decode()


class CsdrError(RuntimeError):
    pass


def check_handle(h, what):
    if not h:
        raise CsdrError(f"{what} failed: {last_error() or 'unknown error'}")
    return h


def check_call(what):
    e = last_error()
    if e:
        raise CsdrError(f"{what}: {e}")
