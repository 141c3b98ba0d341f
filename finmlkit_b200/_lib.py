"""ctypes binding of libfmk.so (include/fmk.h).  There is no CPU fallback: if the library or a CUDA device is missing,
every compute call raises."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(_HERE, "libfmk.so")

_lib = None

P = C.c_void_p
I64 = C.c_int64
F64 = C.c_double
INT = C.c_int

# name -> (restype, argtypes); every symbol declared in include/fmk.h
SIGNATURES = {
    "fmk_version": (C.c_char_p, []),
    "fmk_device_count": (INT, []),
    "fmk_ctx_create": (INT, [INT, C.POINTER(P)]),
    "fmk_ctx_create_on_stream": (INT, [INT, P, C.POINTER(P)]),
    "fmk_ctx_destroy": (None, [P]),
    "fmk_prof_enable": (INT, [P, INT]),
    "fmk_prof_report": (INT, [P, P, P, P, INT]),
    "fmk_result_cols": (INT, [P, C.POINTER(P), C.POINTER(I64), C.POINTER(I64)]),
    "fmk_last_error": (C.c_char_p, [P]),
    "fmk_ctx_sync": (INT, [P]),
    "fmk_ctx_trim": (INT, [P]),
    "fmk_timer_start": (INT, [P]),
    "fmk_timer_stop": (INT, [P, C.POINTER(C.c_float)]),
    "fmk_launch_count": (I64, [P]),
    "fmk_flush_l2": (INT, [P]),
    "fmk_host_alloc": (INT, [C.POINTER(P), I64]),
    "fmk_host_free": (None, [P]),
    "fmk_trades_upload": (INT, [P, P, P, P, P, I64, C.POINTER(P)]),
    "fmk_trades_upload_f32amt": (INT, [P, P, P, P, P, I64, C.POINTER(P)]),
    "fmk_trades_alloc": (INT, [P, I64, INT, INT, C.POINTER(P)]),
    "fmk_trades_write": (INT, [P, P, I64, I64, P, P, P, INT, P]),
    "fmk_trades_add_column": (INT, [P, P, INT, P]),
    "fmk_trades_synth": (INT, [P, I64, C.c_uint64, C.POINTER(P)]),
    "fmk_trades_refill": (INT, [P, P, P, P, P, P, I64]),
    "fmk_trades_download": (INT, [P, P, P, P, P, P]),
    "fmk_trades_size": (I64, [P]),
    "fmk_trades_free": (None, [P, P]),
    "fmk_buf_upload": (INT, [P, P, I64, C.POINTER(P)]),
    "fmk_buf_alloc": (INT, [P, I64, C.POINTER(P)]),
    "fmk_buf_download": (INT, [P, P, P, I64]),
    "fmk_buf_bytes": (I64, [P]),
    "fmk_buf_devptr": (P, [P]),
    "fmk_buf_free": (None, [P, P]),
    "fmk_buf_gather8": (INT, [P, P, P, I64, P]),
    "fmk_time_bar_index": (INT, [P, P, F64, C.POINTER(P)]),
    "fmk_tick_bar_index": (INT, [P, P, I64, C.POINTER(P)]),
    "fmk_volume_bar_index": (INT, [P, P, F64, C.POINTER(P)]),
    "fmk_dollar_bar_index": (INT, [P, P, F64, C.POINTER(P)]),
    "fmk_cusum_bar_index": (INT, [P, P, P, F64, F64, C.POINTER(P)]),
    "fmk_imbalance_bar_index": (INT, [P, P, F64, INT, INT, C.POINTER(P)]),
    "fmk_cusum_filled_count": (I64, [P]),
    "fmk_index_from_host": (INT, [P, P, P, I64, C.POINTER(P)]),
    "fmk_index_size": (I64, [P]),
    "fmk_index_download": (INT, [P, P, P, P]),
    "fmk_index_free": (None, [P, P]),
    "fmk_index_stats": (INT, [P, P]),
    "fmk_bar_ohlcv": (INT, [P, P, P] + [P] * 8),
    "fmk_bar_ohlcv_device": (INT, [P, P, P, INT]),
    "fmk_bar_directional": (INT, [P, P, P] + [P] * 14),
    "fmk_bar_trade_size": (INT, [P, P, P, P, I64, F64, P, P, P, P]),
    "fmk_bar_footprints": (INT, [P, P, P, F64, P, P, F64, C.POINTER(P)]),
    "fmk_footprint_levels": (I64, [P]),
    "fmk_footprint_download": (INT, [P, P] + [P] * 14),
    "fmk_footprint_free": (None, [P, P]),
    "fmk_lagged_returns": (INT, [P, P, P, I64, F64, INT, P]),
    "fmk_ewmst": (INT, [P, P, P, I64, F64, F64, P]),
    "fmk_lagged_returns_dev": (INT, [P, P, F64, INT, C.POINTER(P)]),
    "fmk_ewmst_dev": (INT, [P, P, P, F64, F64, C.POINTER(P)]),
    "fmk_realized_vol": (INT, [P, P, I64, I64, INT, P]),
    "fmk_ewms": (INT, [P, P, I64, I64, P]),
    "fmk_vpin": (INT, [P, P, P, I64, I64, P]),
    "fmk_flow_acceleration": (INT, [P, P, I64, I64, I64, P]),
    "fmk_average_uniqueness": (INT, [P, I64, P, P, I64, I64, P, P]),
    "fmk_return_attribution": (INT, [P, P, P, I64, P, P, I64, INT, P]),
    "fmk_sample_weights": (INT, [P, P, P, P, I64, INT, P, P, P]),
    "fmk_cusum_filter": (INT, [P, P, I64, P, I64, C.POINTER(P), C.POINTER(I64)]),
    "fmk_trade_side_vector": (INT, [P, P, I64, P]),
    "fmk_merge_split_trades": (INT, [P, P, P, P, P, I64, P, P, P, P, C.POINTER(I64)]),
    "fmk_volume_profile_rolling": (INT, [P, P, P, P, P, I64, P, P, P, F64, I64, F64, F64, P, P, P, P]),
    "fmk_volume_profile_rolling_fp": (INT, [P, P, P, P, P, F64, I64, F64, F64, P, P, P, P]),
    "fmk_bar_features_device": (INT, [P, P, P, INT, P, I64, F64, F64, F64, C.POINTER(P)]),
    "fmk_frame_info": (INT, [P, C.POINTER(I64), C.POINTER(I64), C.POINTER(I64), C.POINTER(I64), P]),
    "fmk_frame_devptrs": (INT, [P, C.POINTER(P), C.POINTER(P)]),
    "fmk_frame_download": (INT, [P, P, P, P]),
    "fmk_frame_free": (None, [P, P]),
    "fmk_comm_unique_id": (INT, [P]),
    "fmk_comm_init": (INT, [P, P, INT, INT, INT, C.POINTER(P)]),
    "fmk_comm_destroy": (None, [P]),
    "fmk_comm_rank": (INT, [P]),
    "fmk_comm_world": (INT, [P]),
    "fmk_comm_nccl_version": (INT, []),
    "fmk_comm_p2p_active": (INT, [P]),
    "fmk_comm_barrier": (INT, [P]),
    "fmk_comm_allreduce_f64": (INT, [P, P, INT, INT]),
    "fmk_comm_gather_submit": (INT, [P, P, P, INT, INT]),
    "fmk_comm_gather_finish": (INT, [P]),
    "fmk_comm_gather_reset": (INT, [P]),
    "fmk_comm_gather_result": (INT, [P, INT, C.POINTER(P), C.POINTER(I64)]),
    "fmk_comm_gather_download": (INT, [P, INT, P, I64]),
    "fmk_triple_barrier": (INT, [P, P, P, P, I64, I64, F64, F64, F64, F64, P, I64, F64, P, P, P, P]),
}


def lib():
    """Load libfmk.so (no build here: __graft_entry__.build() / finmlkit_b200.build.build() produce it)."""
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH):
            raise RuntimeError(f"{SO_PATH} is missing: run `python -m finmlkit_b200.build` (nvcc, sm_100a). "
                               "finmlkit_b200 has no CPU fallback.")
        L = C.CDLL(SO_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib
