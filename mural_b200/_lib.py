"""ctypes binding of libmural_b200.so (include/mural_b200.h).

There is no CPU fallback: if the library is missing the import of anything that needs it raises.
`python -m mural_b200.build` (or __graft_entry__.build()) produces the library in-tree.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmural_b200.so")

MODEL_SNV, MODEL_INDEL = 0, 1
MODE_FP32, MODE_BF16, MODE_AUTO = 0, 1, 2
MODES = {"fp32": MODE_FP32, "bf16": MODE_BF16, "auto": MODE_AUTO}


class SnvConfig(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("local_radius", "local_order", "distal_radius", "hidden1", "hidden2",
                                          "channels", "kernel_size", "n_class", "n_cont")]


class IndelConfig(C.Structure):
    _fields_ = [("distal_radius", C.c_int32), ("channels", C.c_int32), ("kernel_size", C.c_int32), ("n_class", C.c_int32),
                ("downsize", C.c_int32 * 6), ("use_reverse", C.c_int32)]


_vp, _i32, _i64 = C.c_void_p, C.c_int32, C.c_int64
# name -> (restype, argtypes).  tests/test_abi.py checks this table against include/mural_b200.h.
PROTOTYPES = {
    "mural_last_error": (C.c_char_p, []),
    "mural_abi_version": (C.c_int, []),
    "mural_launch_count": (_i64, []),
    "mural_reset_launch_count": (None, []),
    "mural_profile_begin": (None, []),
    "mural_profile_end": (_i64, [C.c_char_p, _i64]),
    "mural_snv_tc_available": (C.c_int, [_vp]),
    "mural_snv_last_auto_sites": (_i64, [_vp]),
    "mural_snv_set_cont": (C.c_int, [_vp, _vp]),
    "mural_bigwig_open": (C.c_int, [C.c_char_p, C.POINTER(_vp)]),
    "mural_bigwig_n_chrom": (_i32, [_vp]),
    "mural_bigwig_chrom_name": (C.c_char_p, [_vp, _i32]),
    "mural_bigwig_chrom_len": (_i64, [_vp, _i32]),
    "mural_bigwig_window_means": (C.c_int, [_vp, C.c_char_p, _i64, _vp, _vp, _vp]),
    "mural_bigwig_close": (None, [_vp]),
    "mural_conv32_layer": (C.c_int, [_vp, _vp, _vp, _vp, _i64, _i32, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _vp]),
    "mural_conv32_wgrad": (C.c_int, [_vp, _vp, _i64, _i32, _i32, _vp, _vp, _vp, _vp, _i32, _vp]),
    "mural_genome_create": (C.c_int, [_i32, C.POINTER(C.c_char_p), C.POINTER(_i64), C.c_int, C.POINTER(_vp)]),
    "mural_genome_destroy": (None, [_vp]),
    "mural_genome_n_chrom": (_i32, [_vp]),
    "mural_genome_chrom_len": (_i64, [_vp, _i32]),
    "mural_genome_device_bytes": (_i64, [_vp]),
    "mural_genome_n_exception_runs": (_i64, [_vp]),
    "mural_encode_local": (C.c_int, [_vp, _vp, _vp, _i64, _i32, _i32, _i32, _vp, _vp]),
    "mural_encode_onehot": (C.c_int, [_vp, _vp, _vp, _i64, _i32, _i32, _vp, _vp]),
    "mural_encode_local_host": (C.c_int, [_vp, _vp, _vp, _i64, _i32, _i32, _i32, _vp]),
    "mural_encode_onehot_host": (C.c_int, [_vp, _vp, _vp, _i64, _i32, _i32, _vp]),
    "mural_onehot_to_symbols": (C.c_int, [_vp, _i64, _i32, _vp, _vp]),
    "mural_snv_model_create": (C.c_int, [C.POINTER(SnvConfig), C.c_int, C.POINTER(_vp)]),
    "mural_snv_model_destroy": (None, [_vp]),
    "mural_snv_model_n_tensors": (_i32, [_vp]),
    "mural_snv_model_tensor": (C.c_int, [_vp, _i32, C.POINTER(C.c_char_p), C.POINTER(_i64), C.POINTER(_i64), C.POINTER(_i32)]),
    "mural_snv_model_n_params": (_i64, [_vp]),
    "mural_snv_model_n_trainable": (_i64, [_vp]),
    "mural_snv_model_load": (C.c_int, [_vp, _vp, _i64]),
    "mural_snv_forward": (C.c_int, [_vp, _vp, _vp, _vp, _i64, _i32, _vp, _vp]),
    "mural_snv_forward_tensors": (C.c_int, [_vp, _vp, _vp, _i64, _i32, _i32, _vp, _vp]),
    "mural_snv_predict_host": (C.c_int, [_vp, _vp, _vp, _vp, _i64, _i32, _vp, _vp]),
    "mural_ce_sum": (C.c_int, [_vp, _vp, _i64, _i32, _vp, _vp]),
    "mural_snv_set_chunk": (C.c_int, [_vp, _i64]),
    "mural_snv_set_debug": (C.c_int, [_vp, _i32]),
    "mural_snv_debug_tap": (C.c_int, [_vp, C.c_char_p, _vp, _i64, C.POINTER(_i64)]),
    "mural_indel_model_create": (C.c_int, [C.POINTER(IndelConfig), C.c_int, C.POINTER(_vp)]),
    "mural_indel_model_destroy": (None, [_vp]),
    "mural_indel_model_n_tensors": (_i32, [_vp]),
    "mural_indel_model_tensor": (C.c_int, [_vp, _i32, C.POINTER(C.c_char_p), C.POINTER(_i64), C.POINTER(_i64), C.POINTER(_i32)]),
    "mural_indel_model_n_params": (_i64, [_vp]),
    "mural_indel_model_load": (C.c_int, [_vp, _vp, _i64]),
    "mural_indel_forward": (C.c_int, [_vp, _vp, _vp, _vp, _i64, _vp, _vp]),
    "mural_indel_forward_tensors": (C.c_int, [_vp, _vp, _i64, _i32, _vp, _vp]),
    "mural_indel_set_mode": (C.c_int, [_vp, _i32]),
    "mural_indel_tc_available": (C.c_int, [_vp]),
    "mural_snv_train_create": (C.c_int, [_vp, C.POINTER(_vp)]),
    "mural_snv_train_destroy": (None, [_vp]),
    "mural_snv_train_set_dropout": (C.c_int, [_vp, C.c_float, C.c_float, C.c_float, C.c_uint64]),
    "mural_snv_train_forward": (C.c_int, [_vp, _vp, _vp, _vp, _i64, _vp, _vp, _vp]),
    "mural_snv_train_backward": (C.c_int, [_vp, _vp, _vp, _vp, _vp]),
    "mural_ce_sum_grad": (C.c_int, [_vp, _vp, _i64, _i32, _vp, _vp, _vp]),
    "mural_optimizer_step": (C.c_int, [_i32, _vp, _vp, _vp, _vp, _vp, _i64, C.c_float, C.c_float, _i64, C.c_float, C.c_float, _vp, _vp]),
    "mural_indel_model_config": (C.c_int, [_vp, _vp]),
    "mural_indel_train_create": (C.c_int, [_vp, C.POINTER(_vp)]),
    "mural_indel_train_destroy": (None, [_vp]),
    "mural_indel_train_set_dropout": (C.c_int, [_vp, C.c_float, C.c_uint64]),
    "mural_indel_train_forward": (C.c_int, [_vp, _vp, _vp, _vp, _i64, _vp, _vp, _vp]),
    "mural_indel_train_forward_tensors": (C.c_int, [_vp, _vp, _i64, _i32, _vp, _vp, _vp]),
    "mural_indel_train_backward": (C.c_int, [_vp, _vp, _vp, _vp, _vp]),
    "mural_optimizer_step_dev": (C.c_int, [_i32, _vp, _vp, _vp, _vp, _vp, _i64, _vp, C.c_float, _vp, C.c_float, C.c_float, _vp, _vp]),
    "mural_calibrate": (C.c_int, [_vp, _i64, _i32, _vp, _i32, _vp, _vp]),
    "mural_kmer_group_stats": (C.c_int, [_vp, _i64, _i32, _i32, _vp, _vp, _i32, _i64, _vp, _vp]),
    "mural_window_runs": (C.c_int, [_vp, _vp, _vp, _vp, _i64, _i32, _i32, _vp, _vp, _i64, _vp]),
    "mural_genome_exception_runs": (C.c_int, [_vp, _vp, _vp, _vp]),
    "mural_bed_read": (C.c_int, [C.c_char_p, _vp]),
    "mural_bed_n": (_i64, [_vp]),
    "mural_bed_n_chrom": (_i32, [_vp]),
    "mural_bed_chrom_name": (C.c_char_p, [_vp, _i32]),
    "mural_bed_columns": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp]),
    "mural_bed_destroy": (None, [_vp]),
    "mural_segment_order": (C.c_int, [_vp, _vp, _vp, _i64, _i64, _vp, _vp, _vp]),
    "mural_pack_sites": (C.c_int, [_vp, _i64, _vp, _vp, _vp, _vp, _vp, _i32, _vp, _vp, _vp, _vp, _vp]),
    "mural_fasta_read": (C.c_int, [C.c_char_p, _vp]),
    "mural_fasta_n": (_i32, [_vp]),
    "mural_fasta_name": (C.c_char_p, [_vp, _i32]),
    "mural_fasta_seq": (_vp, [_vp, _i32]),
    "mural_fasta_len": (_i64, [_vp, _i32]),
    "mural_fasta_destroy": (None, [_vp]),
    "mural_write_tsv": (C.c_int, [C.c_char_p, _i64, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32]),
    "mural_format_g4": (C.c_int, [C.c_double, C.c_char_p]),
}

_lib = None


def lib():
    """The loaded library; raises (loudly) if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("mural_b200: %s is missing — build it with `python -m mural_b200.build` "
                               "(there is no CPU/PyTorch fallback)" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(L, name)          # AttributeError if the .so does not export a declared symbol
            fn.restype, fn.argtypes = res, args
        _lib = L
    return _lib


def check(rc):
    if rc != 0:
        msg = lib().mural_last_error().decode("utf-8", "replace")
        if "KeyError" in msg:
            raise KeyError(msg)
        if "ValueError" in msg:
            raise ValueError(msg)
        if "IndexError" in msg:
            raise IndexError(msg)
        raise RuntimeError("mural_b200: " + msg)


def current_stream():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    """data pointer of a torch tensor / numpy array (None -> NULL)."""
    if t is None:
        return C.c_void_p(0)
    if hasattr(t, "data_ptr"):
        return C.c_void_p(t.data_ptr())
    return C.c_void_p(t.ctypes.data)
