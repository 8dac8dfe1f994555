"""ctypes loader for libblaze_b200.so (the C ABI in include/blaze_b200.h).

The library is built in-tree by `__graft_entry__.build()` / `make -C blaze_b200/csrc`.
There is no fallback: if the shared library is missing this module raises, and if there is no
CUDA device every constructor raises DriverClientError (BZ_ERR_NO_DEVICE).
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("BLAZE_B200_LIB") or os.path.join(_HERE, "libblaze_b200.so")   # override: A/B builds only
_lib = None

u8p = ctypes.POINTER(ctypes.c_uint8)
u32p = ctypes.POINTER(ctypes.c_uint32)
u64p = ctypes.POINTER(ctypes.c_uint64)
vp = ctypes.c_void_p
i32, u32, u64, sz = ctypes.c_int32, ctypes.c_uint32, ctypes.c_uint64, ctypes.c_size_t

# name -> argtypes (restype is int32 unless listed in _RESTYPES)
SIGNATURES = {
    "bz_last_error": [],
    "bz_version": [],
    "bz_kernel_launch_count": [],
    "bz_dclient_new": [ctypes.c_char_p, i32, ctypes.POINTER(vp)],
    "bz_dclient_free": [vp],
    "bz_dclient_device_count": [vp, u32p],
    "bz_comm_unique_id": [vp],
    "bz_dclient_comm_init": [vp, i32, i32, vp],
    "bz_dclient_comm_info": [vp, ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(ctypes.c_int32)],
    "bz_dclient_reset": [vp],
    "bz_dclient_dma_write": [vp, u64, u64, vp, sz],
    "bz_dclient_dma_read": [vp, u64, u64, vp, sz],
    "bz_dclient_firewalls_status": [vp, u32p],
    "bz_dclient_unblock_firewalls": [vp],
    "bz_dclient_initialize_cms": [vp],
    "bz_dclient_reset_sensor_data": [vp],
    "bz_dclient_setup_before_load_binary": [vp],
    "bz_dclient_load_binary": [vp, vp, sz],
    "bz_dclient_device_info": [vp, ctypes.c_char_p, sz, u64p, u64p],
    "bz_host_alloc": [sz, ctypes.POINTER(vp)],
    "bz_host_free": [vp],
    "bz_msm_new": [vp, i32, i32, i32, ctypes.POINTER(vp)],
    "bz_msm_free": [vp],
    "bz_msm_loaded_binary_parameters": [vp, u32p],
    "bz_msm_initialize": [vp, u32, i32, u64, u64],
    "bz_msm_start_process": [vp],
    "bz_msm_set_data": [vp, vp, sz, vp, sz, u32, i32, u64, u64],
    "bz_msm_wait_result": [vp],
    "bz_msm_result": [vp, vp, sz, u32p],
    "bz_msm_task_label": [vp, u32p],
    "bz_msm_nof_elements": [vp, u32p],
    "bz_msm_is_msm_engine_ready": [vp, u32p],
    "bz_msm_load_data_to_hbm": [vp, vp, sz, u64, u64],
    "bz_msm_get_data_from_hbm": [vp, vp, sz, u64, u64],
    "bz_msm_sizes": [vp, u32p, u32p, u32p, u32p],
    "bz_msm_phase_times": [vp, ctypes.POINTER(ctypes.c_float)],
    "bz_msm_set_window_bits": [vp, i32],
    "bz_msm_set_accumulate_mode": [vp, i32, i32],
    "bz_msm_get_api": [vp, u32p, sz],
    "bz_msm_table_build_ms": [vp, ctypes.POINTER(ctypes.c_float)],
    "bz_msm_plan_info": [vp, u32p],
    "bz_msm_plan_info_ex": [vp, u32p],
    "bz_msm_set_precompute": [vp, i32],
    "bz_msm_set_raw_result": [vp, i32],
    "bz_msm_set_scalars_device": [vp, u64, u32, i32, u64, u64],
    "bz_msm_combine_results": [vp, vp, i32, vp, sz],
    "bz_msm_generate_chain_points": [vp, vp, sz, u64, u64, u64, u64],
    "bz_msm_field_selftest": [vp, vp, vp, vp, i32, i32],
    "bz_msm_expand_precompute": [vp, u64, u64, u64],
    "bz_ntt_new": [vp, i32, ctypes.POINTER(vp)],
    "bz_ntt_new_ex": [vp, i32, i32, i32, ctypes.POINTER(vp)],
    "bz_ntt_free": [vp],
    "bz_ntt_loaded_binary_parameters": [vp, u32p],
    "bz_ntt_initialize": [vp],
    "bz_ntt_set_data": [vp, sz, vp, sz],
    "bz_ntt_start_process": [vp, sz],
    "bz_ntt_wait_result": [vp],
    "bz_ntt_result": [vp, sz, vp, sz],
    "bz_ntt_phase_times": [vp, ctypes.POINTER(ctypes.c_float), u32p],
    "bz_ntt_slot_device_ptr": [vp, sz, u64p],
    "bz_ntt_dist_new": [vp, i32, i32, i32, i32, i32, ctypes.POINTER(vp)],
    "bz_ntt_dist_free": [vp],
    "bz_ntt_dist_ipc_handle": [vp, vp],
    "bz_ntt_dist_open_peers": [vp, vp],
    "bz_ntt_dist_set_input": [vp, vp, sz],
    "bz_ntt_dist_get_output": [vp, vp, sz],
    "bz_ntt_dist_buffers": [vp, u64p, u64p, u64p],
    "bz_ntt_dist_run": [vp],
    "bz_ntt_dist_step1": [vp],
    "bz_ntt_dist_sync": [vp],
    "bz_ntt_dist_step3": [vp],
    "bz_ntt_dist_times": [vp, ctypes.POINTER(ctypes.c_float)],
    "bz_ntt_dist_plan": [vp, ctypes.POINTER(ctypes.c_int32)],
    "bz_poseidon_new": [vp, i32, ctypes.POINTER(vp)],
    "bz_poseidon_free": [vp],
    "bz_poseidon_loaded_binary_parameters": [vp, u32p],
    "bz_poseidon_initialize": [vp, u32, i32, ctypes.c_char_p],
    "bz_poseidon_set_data": [vp, vp, sz],
    "bz_poseidon_start_process": [vp],
    "bz_poseidon_wait_result": [vp],
    "bz_poseidon_result": [vp, sz, vp, sz, ctypes.POINTER(sz)],
    "bz_poseidon_get_last_element_sent_to_ring": [vp, u32p],
    "bz_poseidon_get_num_of_pending_results": [vp, u32p],
    "bz_poseidon_get_raw_results": [vp, u32, vp],
    "bz_poseidon_get_last_hash_sent_to_host": [vp, u32p],
    "bz_poseidon_device_ms": [vp, ctypes.POINTER(ctypes.c_float)],
    "bz_poseidon_permute": [vp, i32, i32, vp, sz, vp],
    "bz_poseidon_optimized_constants": [i32, i32, vp, sz, ctypes.POINTER(sz)],
}
_RESTYPES = {"bz_last_error": ctypes.c_char_p, "bz_version": ctypes.c_char_p, "bz_kernel_launch_count": ctypes.c_uint64}


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                "blaze_b200: %s is missing -- build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU fallback)" % LIB_PATH)
        L = ctypes.CDLL(LIB_PATH)
        for name, args in SIGNATURES.items():
            fn = getattr(L, name)
            fn.argtypes = args
            fn.restype = _RESTYPES.get(name, i32)
        _lib = L
    return _lib


def buf_ptr(b):
    """(address, length, keepalive) of a bytes / bytearray / numpy array / (addr, len) tuple."""
    if b is None:
        return None, 0, None
    if isinstance(b, tuple):
        return ctypes.c_void_p(b[0]), b[1], None
    if isinstance(b, (bytes, bytearray)):
        n = len(b)
        if isinstance(b, bytes):
            keep = ctypes.c_char_p(b)
            return ctypes.cast(keep, ctypes.c_void_p), n, (keep, b)
        arr = (ctypes.c_uint8 * n).from_buffer(b)
        return ctypes.cast(arr, ctypes.c_void_p), n, arr
    # numpy-like
    if hasattr(b, "ctypes") and hasattr(b, "nbytes"):
        return ctypes.c_void_p(b.ctypes.data), int(b.nbytes), b
    mv = memoryview(b)
    raise TypeError("unsupported buffer type %r" % type(b))
