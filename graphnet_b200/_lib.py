"""ctypes binding of the C-ABI library `csrc/libgraphnet_b200.so` (see include/graphnet_b200.h).

There is no CPU fallback: if the CUDA library is missing or a launch fails the
operators raise `RuntimeError`.
"""

from __future__ import annotations

import ctypes
import os
from typing import Dict, Optional

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libgraphnet_b200.so")

_p = ctypes.c_void_p
_i32 = ctypes.c_int32
_i64 = ctypes.c_int64
_f32 = ctypes.c_float

# name -> argument types (return type is always int status); mirrors include/graphnet_b200.h
SIGNATURES: Dict[str, list] = {
    "gnb_batch_to_ptr": [_p, _i64, _i64, _p, _p],
    "gnb_knn_table": [_p, _i64, _p, _i32, _p, _i64, _i64, _i32, _p, _p, _p],
    "gnb_table_to_edge_index": [_p, _p, _p, _i64, _i32, _i64, _p, _p],
    "gnb_global_vars": [_p, _i64, _i32, _p, _p, _i32, _p, _i64, _p, _p, _p, _i64, _p],
    "gnb_edge_hidden_fwd": [_p, _i64, _i32, _p, _p, _i32, _i64, _i32, _p, _i64, _p],
    "gnb_edge_hidden_bwd": [_p, _i64, _p, _i64, _i32, _p, _p, _i32, _i64, _i32, _p, _i64, _p],
    "gnb_edgeconv_fused_fwd_tf32": [_p, _i64, _i32, _p, _p, _i32, _i64, _p, _i64, _p, _i32, _i32, _i32, _p, _i64, _p],
    "gnb_edgeconv_set_variant": [_i32],
    "gnb_linear_set_debug": [_i32],
    "gnb_linear_set_variant": [_i32],
    "gnb_linear_set_pair_resident": [_i32],
    "gnb_knn_set_variant": [_i32],
    "gnb_standardize": [_p, _i64, _i64, _i32, _p, _p, _p, _p, _i64, _p],
    "gnb_ptr_to_batch": [_p, _i64, _i64, _p, _p],
    "gnb_adam_flat": [_p, _p, _p, _p, _i64, _f32, _f32, _f32, _f32, _f32, _f32, _f32, _i32, _p],
    "gnb_task_heads_fwd": [_p, _i64, _i32, _p, _p, _p, _p, _p, _p, _i64, _p, _p, _p, _p, _p],
    "gnb_task_heads_bwd": [_p, _i64, _i32, _p, _p, _p, _p, _i64, _p, _i64, _p, _p, _p, _p, _p],
    "gnb_linear_set_profile_buffer": [_p],
    "gnb_edgeconv_set_profile_buffer": [_p],
    "gnb_edge_cat_fwd": [_p, _i64, _i32, _p, _p, _i32, _i64, _p, _i64, _p],
    "gnb_edge_cat_bwd": [_p, _i64, _i32, _p, _p, _i32, _i64, _p, _i64, _p],
    "gnb_edge_aggregate_fwd": [_p, _i64, _i32, _p, _i32, _i64, _i32, _p, _i64, _p, _p],
    "gnb_edge_aggregate_bwd": [_p, _i64, _i32, _p, _i32, _i64, _i32, _p, _p, _i64, _p],
    "gnb_segment_pool_fwd": [_p, _i64, _i32, _p, _i64, _p, _i32, _p, _p, _p],
    "gnb_segment_pool_bwd": [_p, _i64, _p, _i32, _p, _i64, _i64, _p, _i32, _p, _i64, _p],
    "gnb_relu_bwd": [_p, _i64, _p, _i64, _i64, _i32, _p, _i64, _i32, _p],
    "gnb_linear_fwd_tf32": [_p, _p, _p, _i32, _p, _i64, _p, _p, _i64, _i64, _i32, _i32, _i32, _p],
    "gnb_linear_fwd_tf32x3": [_p, _p, _p, _i32, _p, _p, _i64, _p, _p, _i64, _i64, _i32, _i32, _p],
    "gnb_split_pad_tf32": [_p, _i64, _i64, _i32, _p, _p, _i64, _i32, _p],
    "gnb_edge_linear_agg_fwd_tf32x3": [_p, _i64, _i32, _p, _p, _i64, _p, _p, _i64, _i32, _p, _i64, _p, _p],
    "gnb_edge_linear_aggmax_fwd_tf32": [_p, _i64, _i32, _p, _i64, _p, _p, _i64, _i32, _i32, _i32, _p, _i64, _p, _i64, _p],
    "gnb_edge_linear_aggmax_fwd_tf32x3": [_p, _i64, _i32, _p, _p, _i64, _p, _p, _i64, _i32, _i32, _p, _i64, _p, _i64, _p],
    "gnb_edge_argmax_bwd": [_p, _i64, _p, _i64, _i32, _i32, _i64, _i32, _i32, _p, _i64, _p, _p],
    "gnb_linear_bwd_weight_tf32": [_p, _i64, _p, _i64, _p, _i64, _i64, _i32, _i32, _i32, _p],
    "gnb_edge_linear_agg_fwd_tf32": [_p, _i64, _i32, _p, _i64, _p, _p, _i64, _i32, _i32, _p, _i64, _p, _p],
    "gnb_edge_hidden_dgrad_scatter_tf32": [_p, _i64, _i32, _p, _i64, _p, _i32, _i32, _p, _i64, _p, _i64, _p],
    "gnb_edge_hidden_dgrad_scatter_split_tf32": [_p, _i64, _i32, _p, _i64, _p, _i32, _i32, _p, _i64, _p, _i64, _p, _i64, _p,
                                                 _i32, _p],
    "gnb_edge_hidden_fwd_mask": [_p, _i64, _i32, _p, _p, _i32, _i64, _i32, _p, _i64, _p, _i32, _p],
    "gnb_edge_mask_bwd_colsum": [_p, _i64, _p, _i64, _i32, _p, _p, _i64, _p, _i32, _p],
    "gnb_round_pad_tf32": [_p, _i64, _i64, _i32, _p, _i64, _i32, _p],
    "gnb_act_bwd_colsum": [_p, _i64, _p, _i64, _i64, _i32, _p, _i64, _p, _i32, _p, _i32, _i32, _p],
    "gnb_colsum": [_p, _i64, _i64, _i32, _p, _p],
    "gnb_linear_fwd_f32": [_p, _i64, _p, _i64, _p, _p, _i64, _i64, _i64, _i64, _i32, _i32, _p],
    "gnb_linear_bwd_data_f32": [_p, _i64, _p, _i64, _p, _i64, _i64, _i64, _i64, _i32, _p],
    "gnb_linear_bwd_weight_f32": [_p, _i64, _p, _i64, _p, _i64, _i64, _i64, _i64, _p],
    "gnb_edge_hidden_fwd_bf16": [_p, _i64, _i32, _p, _p, _i32, _i64, _p, _p, _i64, _p, _i32, _p],
    "gnb_edge_linear_agg_fwd_bf16": [_p, _p, _i64, _i32, _p, _p, _i64, _p, _p, _i64, _i32, _i32, _p, _i64, _p, _p],
    "gnb_edge_mask_bwd_colsum_bf16": [_p, _i64, _p, _i64, _i32, _p, _p, _i64, _p, _p],
    "gnb_linear_bwd_weight_bf16": [_p, _p, _i64, _p, _p, _i64, _p, _i64, _i64, _i32, _i32, _i32, _p],
    "gnb_edge_hidden_dgrad_scatter_bf16": [_p, _p, _i64, _i32, _p, _p, _i64, _p, _i32, _i32, _p, _i64, _p, _i64, _p, _i64, _p,
                                           _i32, _p],
    "gnb_linear_fwd_bf16": [_p, _p, _i64, _i32, _p, _p, _i64, _p, _p, _i64, _i64, _i32, _i32, _i32, _p],
    "gnb_to_bf16_planes": [_p, _i64, _i64, _i32, _p, _p, _i64, _i32, _i32, _p],
    "gnb_absmax_bits": [_p, _i64, _i64, _i32, _i32, _p, _p],
    "gnb_edge_hidden_fwd_f16": [_p, _i64, _i32, _p, _p, _i32, _i64, _p, _p, _i64, _p, _i32, _p, _p],
    "gnb_edge_linear_agg_fwd_f16": [_p, _p, _i64, _i32, _p, _p, _i64, _p, _p, _i64, _i32, _i32, _p, _i64, _p, _p, _p],
    "gnb_edge_mask_bwd_colsum_f16": [_p, _i64, _p, _i64, _i32, _p, _i64, _p, _p, _p],
    "gnb_linear_bwd_weight_f16": [_p, _i64, _p, _p, _i64, _p, _i64, _i64, _i32, _i32, _p, _p, _p],
    "gnb_edge_hidden_dgrad_scatter_f16": [_p, _i64, _i32, _p, _i64, _p, _i32, _i32, _p, _i64, _p, _i64, _p, _i64, _p, _i32,
                                          _p, _p],
    "gnb_zero_block": [_p, _i64, _i64, _i32, _p],
    "gnb_wgrad_set_debug": [_i32],
    "gnb_edgeconv_fused_fwd_f16": [_p, _i64, _i32, _p, _p, _i64, _p, _p, _i64, _p, _i32, _i32, _p, _i64, _p, _p, _i64, _p, _i64,
                                   _p, _i32, _p],
    "gnb_edge_dz_prep": [_p, _i64, _p, _i64, _i32, _p, _p, _p, _p, _p],
    "gnb_linear_bwd_weight_f16_masked": [_p, _p, _p, _i64, _p, _i64, _i64, _i32, _i32, _p, _p, _p],
    "gnb_edge_hidden_dgrad_scatter_f16_masked": [_p, _p, _i32, _p, _i64, _p, _i32, _i32, _p, _i64, _p, _i64, _p, _i64, _p,
                                                 _i32, _p, _p],
    "gnb_edge_slot_flag": [_p, _i64, _i32, _p, _p],
    "gnb_edge_slot_flag_or": [_p, _i64, _i32, _p, _p],
    "gnb_edgeconv_fused_fwd_f16_w": [_p, _i64, _i32, _p, _p, _i64, _p, _p, _i64, _p, _i32, _i32, _p, _i64, _p, _p, _i64, _p, _i64,
                                     _p, _i32, _p, _p],
    "gnb_edge_dz_prep_w": [_p, _i64, _p, _i64, _i32, _p, _p, _p, _p, _p, _p],
    "gnb_edge_dz_prep_wz": [_p, _i64, _p, _i64, _i32, _p, _p, _p, _p, _p, _p, _i64, _i64, _i32, _p, _i32, _p],
    "gnb_linear_bwd_weight_f16_masked_w": [_p, _p, _p, _i64, _p, _i64, _i64, _i32, _i32, _p, _p, _p, _p],
    "gnb_edge_hidden_dgrad_scatter_f16_masked_w": [_p, _p, _i32, _p, _i64, _p, _i32, _i32, _p, _i64, _p, _i64, _p, _i64, _p,
                                                   _i32, _p, _p, _p],
    "gnb_linear_next_absmax": [_p, _i32],
    "gnb_to_f16_planes": [_p, _i64, _i64, _i32, _p, _p, _i64, _i32, _i32, _p],
}



class DynEdgeConfig(ctypes.Structure):
    """Mirror of `gnb_dynedge_config` (include/graphnet_b200.h)."""
    _fields_ = [("nb_inputs", _i32), ("k", _i32), ("precision", _i32),
                ("n_conv", _i32), ("conv_hidden", _i32 * 8), ("conv_out", _i32 * 8),
                ("n_post", _i32), ("post_out", _i32 * 8),
                ("n_readout", _i32), ("readout_out", _i32 * 8),
                ("n_pool", _i32), ("pool", _i32 * 4),
                ("globals_after_pooling", _i32), ("skip_readout", _i32),
                ("n_knn_cols", _i32), ("knn_cols", _i32 * 16), ("flags", _i32)]


SIGNATURES.update({
    "gnb_launch_count": [],
    "gnb_dynedge_workspace_bytes": [_p, _i64, _i64, _i32, _i32],
    "gnb_dynedge_layout": [_p, _i64, _i64, _i32, _i32, _p],
    "gnb_dynedge_forward": [_p, _p, _p, _i64, _p, _p, _p, _p, _i32, _p, _i64, _i64, _p, _i64, _p, _i32, _p],
    "gnb_dynedge_backward": [_p, _p, _p, _p, _p, _i32, _i64, _i64, _p, _i64, _p, _p],
    "gnb_dynedge_set_backward_event": [_p, _i32],
})
RESTYPES = {"gnb_dynedge_workspace_bytes": ctypes.c_int64, "gnb_launch_count": ctypes.c_int64}

_lib: Optional[ctypes.CDLL] = None


def library_path() -> str:
    return LIB_PATH


def load() -> ctypes.CDLL:
    """Load the CUDA library; raise loudly when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"graphnet_b200: CUDA library {LIB_PATH} is missing. Build it with "
            "`python -c 'import __graft_entry__ as g; g.build()'` (nvcc, sm_100a). "
            "There is no CPU fallback for the DynEdge hot path.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)   # AttributeError here = header / library mismatch
        fn.argtypes = argtypes
        fn.restype = RESTYPES.get(name, ctypes.c_int)
    _lib = lib
    return lib


def check(status: int, what: str) -> None:
    if status == 0:
        return
    if status > 0:
        try:
            import torch
            msg = torch.cuda.cudart().cudaGetErrorString(status)  # type: ignore[attr-defined]
        except Exception:  # pragma: no cover
            msg = "cudaError"
        raise RuntimeError(f"graphnet_b200::{what}: CUDA error {status} ({msg})")
    reason = {-1: "invalid argument (shape/alignment/size)", -2: "unsupported configuration"}.get(status, "error")
    raise RuntimeError(f"graphnet_b200::{what}: {reason} [{status}]")
