"""ctypes binding of libmosaic_b200.so (declared in include/mosaic_b200.h).

There is deliberately NO fallback here: if the shared library is missing or no CUDA device
is present, every compute entry point raises.  PyTorch is used by the callers only to own
device memory and streams; nothing in this module touches torch.
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
import os as _os
LIB_PATH = Path(_os.environ.get("MOSAIC_B200_LIB", PKG_DIR / "libmosaic_b200.so"))     # the override is for A/B builds of the same sources
CSRC = PKG_DIR / "csrc"

MK_OK = 0
NORM_HAMMING, NORM_L2_U8, NORM_L2_F32, NORM_HAMMING_TC = 0, 1, 2, 3
STREAM_NONE = C.c_void_p(-1).value          # MK_STREAM_NONE
INTER_LINEAR, INTER_CUBIC = 1, 2
BLEND_ALPHA, BLEND_FEATHER = 0, 1

EXPORTS = (
    "mk_abi_version", "mk_version_string", "mk_last_error", "mk_launch_count",
    "mk_knn2_workspace_bytes", "mk_knn2_uses_tensor_cores", "mk_knn2", "mk_knn2_batched_workspace_bytes", "mk_knn2_batched",
    "mk_warp_perspective_u8", "mk_mapper_update", "mk_mapper_update_batch", "mk_mapper_footprint", "mk_upload_2d", "mk_download_2d",
    "mk_event_create", "mk_event_destroy", "mk_stream_follow", "mk_sm_partition_create", "mk_sm_partition_destroy",
    "mk_gather_matches", "mk_ransac_workspace_bytes", "mk_ransac_transform", "mk_chain_homographies",
    "mk_ipc_export", "mk_ipc_open", "mk_ipc_close", "mk_copy_peer", "mk_cvt_color_u8", "mk_remap_u8",
)
MODEL_AFFINE, MODEL_PERSPECTIVE = 0, 1


class MosaicB200Error(RuntimeError):
    """Raised when libmosaic_b200 reports an error (the message comes from mk_last_error)."""


def build(verbose: bool = False) -> Path:
    """Compile csrc/*.cu for sm_100a into libmosaic_b200.so (nvcc cross-compiles without a GPU)."""
    out = subprocess.run(["make", "-C", str(CSRC)], capture_output=True, text=True)
    if out.returncode != 0:
        raise MosaicB200Error(f"nvcc build failed:\n{out.stdout}\n{out.stderr}")
    if verbose:
        print(out.stdout)
    return LIB_PATH


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise MosaicB200Error(
            f"{LIB_PATH} not found: run `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU fallback for the CUDA path)")
    L = C.CDLL(str(LIB_PATH))
    vp, i32, sz, f64 = C.c_void_p, C.c_int, C.c_size_t, C.c_double
    L.mk_abi_version.restype = i32
    L.mk_version_string.restype = C.c_char_p
    L.mk_last_error.restype = C.c_char_p
    L.mk_launch_count.restype = C.c_uint64
    L.mk_knn2_workspace_bytes.restype = sz
    L.mk_knn2_workspace_bytes.argtypes = [i32, i32, i32, i32]
    L.mk_knn2_uses_tensor_cores.restype = i32
    L.mk_knn2_uses_tensor_cores.argtypes = [i32, i32, i32, i32]
    L.mk_knn2_batched_workspace_bytes.restype = sz
    L.mk_knn2_batched_workspace_bytes.argtypes = [i32, i32, i32, i32, i32, i32]
    L.mk_knn2.restype = i32
    L.mk_knn2.argtypes = [i32, vp, i32, vp, i32, i32, sz, vp, vp, f64, vp, vp, vp, sz, vp]
    L.mk_knn2_batched.restype = i32
    L.mk_knn2_batched.argtypes = [i32, vp, i32, i32, sz, vp, vp, i32, vp, i32, i32, vp, vp, f64, vp, vp, vp, sz, vp]
    L.mk_warp_perspective_u8.restype = i32
    L.mk_warp_perspective_u8.argtypes = [vp, i32, i32, i32, sz, vp, vp, i32, i32, sz, i32, vp]
    L.mk_mapper_update.restype = i32
    L.mk_mapper_update.argtypes = [vp, sz, vp, sz, vp, sz, i32, i32, vp, i32, i32, sz, vp, sz, vp, f64, i32, i32,
                                   C.c_float, vp, vp]
    L.mk_mapper_update_batch.restype = i32
    L.mk_mapper_update_batch.argtypes = [vp, sz, vp, sz, vp, sz, i32, i32, i32, vp, vp, vp, vp, vp, f64, i32, i32, C.c_float, vp]
    L.mk_upload_2d.restype = i32
    L.mk_upload_2d.argtypes = [vp, sz, vp, sz, sz, sz, vp]
    L.mk_download_2d.restype = i32
    L.mk_download_2d.argtypes = [vp, sz, vp, sz, sz, sz, vp]
    L.mk_event_create.restype = vp
    L.mk_event_destroy.restype = None
    L.mk_event_destroy.argtypes = [vp]
    L.mk_stream_follow.restype = i32
    L.mk_stream_follow.argtypes = [vp, vp, vp]
    L.mk_sm_partition_create.restype = i32
    L.mk_sm_partition_create.argtypes = [i32, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp), C.POINTER(i32), C.POINTER(i32)]
    L.mk_sm_partition_destroy.restype = None
    L.mk_sm_partition_destroy.argtypes = [vp]
    L.mk_cvt_color_u8.restype = i32
    L.mk_cvt_color_u8.argtypes = [vp, i32, i32, sz, i32, vp, sz, vp]
    L.mk_remap_u8.restype = i32
    L.mk_remap_u8.argtypes = [vp, i32, i32, i32, sz, vp, vp, sz, vp, i32, i32, sz, i32, vp]
    L.mk_mapper_footprint.restype = i32
    L.mk_mapper_footprint.argtypes = [i32, i32, i32, i32, vp, i32, vp]
    L.mk_gather_matches.restype = i32
    L.mk_gather_matches.argtypes = [vp, vp, vp, vp, i32, vp, vp, vp, vp, vp]
    L.mk_ransac_workspace_bytes.restype = sz
    L.mk_ransac_transform.restype = i32
    L.mk_ransac_transform.argtypes = [i32, vp, vp, i32, vp, f64, i32, C.c_uint64, vp, vp, vp, vp, vp]
    L.mk_chain_homographies.restype = i32
    L.mk_chain_homographies.argtypes = [vp, i32, vp, vp]
    L.mk_ipc_export.restype = i32
    L.mk_ipc_export.argtypes = [vp, vp, C.POINTER(C.c_uint64)]
    L.mk_ipc_open.restype = i32
    L.mk_ipc_open.argtypes = [vp, C.POINTER(vp)]
    L.mk_ipc_close.restype = i32
    L.mk_ipc_close.argtypes = [vp]
    L.mk_copy_peer.restype = i32
    L.mk_copy_peer.argtypes = [vp, vp, sz, vp]
    if L.mk_abi_version() != 1:
        raise MosaicB200Error("libmosaic_b200 ABI version mismatch")
    _lib = L
    return L


def check(rc: int) -> None:
    if rc != MK_OK:
        raise MosaicB200Error(f"libmosaic_b200 error {rc}: {lib().mk_last_error().decode()}")


def launch_count() -> int:
    return int(lib().mk_launch_count())
