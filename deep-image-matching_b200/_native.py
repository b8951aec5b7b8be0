"""ctypes binding of libdimb200.so (the C ABI in include/dimb200.h).

There is no CPU fallback: importing this module without the built library, or
creating a context without a CUDA device, raises.  Build with
``python -c "import __graft_entry__ as g; g.build()"`` (or ``make -C csrc``).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("DIMB_LIB") or os.path.join(_HERE, "libdimb200.so")  # DIMB_LIB: A/B builds of tools/ only

OK, ERR_CUDA, ERR_OOM, ERR_ARG, ERR_UNSUPPORTED, ERR_CAPACITY = 0, -1, -2, -3, -4, -5
PRECISION_EXACT, PRECISION_FAST = 0, 1
NN_MODES = {"nn": 0, "mnn": 1, "snn": 2, "smnn": 3}

EXPORTS = [
    "dimb_version", "dimb_ctx_create", "dimb_ctx_destroy", "dimb_last_error", "dimb_ctx_set_precision",
    "dimb_ctx_set_tensor_path", "dimb_ctx_launch_count", "dimb_read_dev",
    "dimb_sp_create", "dimb_sp_destroy", "dimb_sp_extract", "dimb_sp_extract_dev", "dimb_sp_debug_read",
    "dimb_lg_create", "dimb_lg_destroy", "dimb_lg_match", "dimb_lg_match_dev", "dimb_lg_debug_read",
    "dimb_nn_match", "dimb_nn_match_dev", "dimb_ctx_profile", "dimb_ctx_profile_read", "dimb_pipe_create", "dimb_pipe_destroy",
    "dimb_pipe_match_image_pairs", "dimb_pipe_match_image_pairs_u8", "dimb_pipe_match_image_pairs_dev", "dimb_pipe_outputs_dev", "dimb_pipe_features_dev", "dimb_sp_ctx",
    "dimb_sg_weight_count", "dimb_sg_create", "dimb_sg_destroy", "dimb_sg_match",
    "dimb_aliked_create", "dimb_aliked_destroy", "dimb_aliked_extract", "dimb_aliked_extract_dev", "dimb_aliked_debug_read",
    "dimb_fstore_create", "dimb_fstore_destroy", "dimb_fstore_put_dev", "dimb_fstore_put", "dimb_fstore_count", "dimb_fstore_get",
    "dimb_fstore_feats_dev", "dimb_fstore_block_dev", "dimb_gv_fundamental", "dimb_gv_fundamental_batch_dev",
]


class DimbError(RuntimeError):
    pass


class SpConf(C.Structure):
    _fields_ = [("nms_radius", C.c_int), ("keypoint_threshold", C.c_float), ("max_keypoints", C.c_int),
                ("remove_borders", C.c_int), ("fix_sampling", C.c_int), ("max_batch", C.c_int),
                ("max_height", C.c_int), ("max_width", C.c_int)]


class AlikedConf(C.Structure):
    _fields_ = [("max_num_keypoints", C.c_int), ("detection_threshold", C.c_float), ("nms_radius", C.c_int),
                ("max_height", C.c_int), ("max_width", C.c_int)]


class SgConf(C.Structure):
    _fields_ = [("n_layers", C.c_int), ("cross_mask", C.c_ulonglong), ("sinkhorn_iterations", C.c_int), ("match_threshold", C.c_float),
                ("max_kpts", C.c_int)]


class SgFeats(C.Structure):
    _fields_ = [("keypoints", C.c_void_p), ("descriptors", C.c_void_p), ("scores", C.c_void_p), ("n", C.c_int), ("desc_ld", C.c_int),
                ("height", C.c_int), ("width", C.c_int)]


class LgConf(C.Structure):
    _fields_ = [("input_dim", C.c_int), ("descriptor_dim", C.c_int), ("n_layers", C.c_int), ("num_heads", C.c_int),
                ("depth_confidence", C.c_double), ("width_confidence", C.c_double), ("filter_threshold", C.c_double),
                ("prune_min_kpts", C.c_int), ("max_pairs", C.c_int), ("max_kpts", C.c_int)]


class Feats(C.Structure):
    _fields_ = [("keypoints", C.c_void_p), ("descriptors", C.c_void_p), ("n", C.c_int), ("desc_layout", C.c_int),
                ("desc_ld", C.c_int), ("has_size", C.c_int), ("size0", C.c_float), ("size1", C.c_float)]


class FeatsDev(C.Structure):
    _fields_ = [("keypoints", C.c_void_p), ("descriptors", C.c_void_p), ("n", C.c_void_p), ("n_cap", C.c_int),
                ("desc_layout", C.c_int), ("desc_ld", C.c_int), ("size0", C.c_float), ("size1", C.c_float),
                ("round_fp16", C.c_int), ("f16", C.c_int), ("size_dev", C.c_void_p)]


_lib = None


def load_library():
    """Load libdimb200.so and declare prototypes. Raises DimbError if the library is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise DimbError(f"{LIB_PATH} is missing: build the CUDA extension first (__graft_entry__.build()); "
                        "this package has no CPU fallback")
    lib = C.CDLL(LIB_PATH)
    vp, ip, fp = C.c_void_p, C.c_int, C.c_float
    lib.dimb_version.restype = C.c_char_p
    lib.dimb_ctx_create.argtypes = [ip, C.POINTER(vp)]
    lib.dimb_ctx_destroy.argtypes = [vp]
    lib.dimb_ctx_destroy.restype = None
    lib.dimb_last_error.argtypes = [vp]
    lib.dimb_last_error.restype = C.c_char_p
    lib.dimb_ctx_set_precision.argtypes = [vp, ip]
    lib.dimb_ctx_set_tensor_path.argtypes = [vp, ip]
    lib.dimb_ctx_launch_count.argtypes = [vp]
    lib.dimb_ctx_launch_count.restype = C.c_ulonglong
    lib.dimb_read_dev.argtypes = [vp, vp, vp, C.c_size_t]
    lib.dimb_sp_create.argtypes = [vp, vp, C.c_size_t, C.POINTER(SpConf), C.POINTER(vp)]
    lib.dimb_sp_destroy.argtypes = [vp]
    lib.dimb_sp_destroy.restype = None
    lib.dimb_sp_extract.argtypes = [vp, vp, ip, ip, ip, vp, vp, vp, vp, ip]
    lib.dimb_sp_extract_dev.argtypes = [vp, vp, ip, ip, ip, vp, vp, vp, vp, ip, vp]
    lib.dimb_sp_debug_read.argtypes = [vp, ip, vp, C.c_size_t]
    lib.dimb_lg_create.argtypes = [vp, vp, C.c_size_t, C.POINTER(LgConf), C.POINTER(vp)]
    lib.dimb_lg_destroy.argtypes = [vp]
    lib.dimb_lg_destroy.restype = None
    lib.dimb_lg_match.argtypes = [vp, ip, C.POINTER(Feats), C.POINTER(Feats), vp, vp, vp, vp, ip]
    lib.dimb_lg_match_dev.argtypes = [vp, ip, C.POINTER(FeatsDev), C.POINTER(FeatsDev), vp, vp, vp, vp, ip, vp]
    lib.dimb_lg_debug_read.argtypes = [vp, ip, ip, vp, C.c_size_t]
    lib.dimb_nn_match.argtypes = [vp, vp, ip, vp, ip, ip, ip, fp, vp, vp, C.POINTER(ip), ip]
    lib.dimb_nn_match_dev.argtypes = [vp, vp, ip, ip, vp, ip, ip, ip, ip, ip, fp, vp, vp, vp, ip, vp]
    lib.dimb_ctx_profile.argtypes = [vp, ip]
    lib.dimb_ctx_profile_read.argtypes = [vp, C.c_char_p, C.c_size_t]
    lib.dimb_pipe_create.argtypes = [vp, vp, ip, ip, ip, ip, C.POINTER(vp)]
    lib.dimb_pipe_destroy.argtypes = [vp]
    lib.dimb_pipe_destroy.restype = None
    lib.dimb_pipe_match_image_pairs.argtypes = [vp, vp, ip, vp, vp, vp, vp, vp, vp]
    lib.dimb_pipe_match_image_pairs_u8.argtypes = [vp, vp, ip, vp, vp, vp, vp, vp, vp]
    lib.dimb_pipe_match_image_pairs_dev.argtypes = [vp, vp, ip, vp]
    lib.dimb_pipe_outputs_dev.argtypes = [vp] + [C.POINTER(vp)] * 6
    lib.dimb_pipe_features_dev.argtypes = [vp] + [C.POINTER(vp)] * 4
    lib.dimb_sp_ctx.argtypes = [vp]
    lib.dimb_sg_weight_count.argtypes = [ip]
    lib.dimb_sg_weight_count.restype = C.c_size_t
    lib.dimb_sg_create.argtypes = [vp, vp, C.c_size_t, C.POINTER(SgConf), C.POINTER(vp)]
    lib.dimb_sg_destroy.argtypes = [vp]
    lib.dimb_sg_destroy.restype = None
    lib.dimb_sg_match.argtypes = [vp, C.POINTER(SgFeats), C.POINTER(SgFeats), vp, vp, C.POINTER(ip), ip]
    lib.dimb_aliked_create.argtypes = [vp, vp, C.c_size_t, C.POINTER(AlikedConf), C.POINTER(vp)]
    lib.dimb_aliked_destroy.argtypes = [vp]
    lib.dimb_aliked_destroy.restype = None
    lib.dimb_aliked_extract.argtypes = [vp, vp, ip, ip, ip, vp, vp, vp, vp, ip]
    lib.dimb_aliked_extract_dev.argtypes = [vp, vp, ip, ip, ip, vp, vp, vp, vp, ip, vp]
    lib.dimb_aliked_debug_read.argtypes = [vp, ip, vp, C.c_size_t]
    lib.dimb_sp_ctx.restype = vp
    lib.dimb_fstore_create.argtypes = [vp, ip, ip, ip, C.POINTER(vp)]
    lib.dimb_fstore_destroy.argtypes = [vp]
    lib.dimb_fstore_destroy.restype = None
    lib.dimb_fstore_put_dev.argtypes = [vp, ip, vp, vp, vp, vp, ip, vp, ip, ip, vp]
    lib.dimb_fstore_put.argtypes = [vp, ip, vp, vp, vp, vp, ip, ip, ip]
    lib.dimb_fstore_count.argtypes = [vp, ip, C.POINTER(ip), vp]
    lib.dimb_fstore_get.argtypes = [vp, ip, vp, vp, vp, vp, C.POINTER(ip), vp, ip]
    lib.dimb_fstore_feats_dev.argtypes = [vp, ip, C.POINTER(FeatsDev)]
    lib.dimb_gv_fundamental.argtypes = [vp, vp, vp, ip, fp, ip, C.c_uint, vp, vp, C.POINTER(ip)]
    lib.dimb_gv_fundamental_batch_dev.argtypes = [vp, ip, vp, vp, vp, vp, ip, fp, ip, C.c_uint, vp, vp, vp, vp]
    lib.dimb_fstore_block_dev.argtypes = [vp, C.POINTER(vp), C.POINTER(C.c_size_t), C.POINTER(ip), C.POINTER(ip)]
    _lib = lib
    return lib


_selftest = None


def load_selftest_library():
    """libdimb200_selftest.so: the production GEMM template behind a C = A B^T entry plus the UMMA descriptor probes.
    Test / tool infrastructure - the product library exports none of it.  Its context is its own (dimb_ctx_create of
    THIS library); never mix handles of the two libraries."""
    global _selftest
    if _selftest is None:
        path = os.path.join(_HERE, "libdimb200_selftest.so")
        if not os.path.exists(path):
            raise DimbError(f"{path} is missing: build the CUDA extension first (__graft_entry__.build())")
        lib = C.CDLL(path)
        vp, ip = C.c_void_p, C.c_int
        lib.dimb_ctx_create.argtypes = [ip, C.POINTER(vp)]
        lib.dimb_ctx_destroy.argtypes = [vp]
        lib.dimb_ctx_destroy.restype = None
        lib.dimb_last_error.argtypes = [vp]
        lib.dimb_last_error.restype = C.c_char_p
        lib.dimb_ctx_set_precision.argtypes = [vp, ip]
        lib.dimb_selftest_gemm.argtypes = [vp, vp, vp, vp, ip, ip, ip, ip]
        lib.dimb_probe_rowshift.argtypes = [vp, vp, vp, vp, ip, ip, ip, ip]
        lib.dimb_probe_rowshift64.argtypes = [vp, vp, vp, vp, ip, ip, ip]
        lib.dimb_probe_tmem_a.argtypes = [vp, vp, vp, vp]
        lib.dimb_gv_host.argtypes = [vp, vp, ip, C.c_float, ip, C.c_uint, vp, vp]
        _selftest = lib
    return _selftest


class SelfTest:
    """Context of the self-test library (tests/test_gpu_parity.py::test_tensor_core_gemm, tools/probe_umma_rowshift.py)."""

    def __init__(self, device: int = 0):
        self.lib = load_selftest_library()
        h = C.c_void_p()
        if self.lib.dimb_ctx_create(device, C.byref(h)) != OK:
            raise DimbError("selftest: dimb_ctx_create failed (a B200 is required)")
        self.h = h

    def check(self, rc, what):
        if rc != OK:
            raise DimbError(f"{what} failed (code {rc}): {self.lib.dimb_last_error(self.h).decode()}")

    def set_precision(self, precision: str):
        self.check(self.lib.dimb_ctx_set_precision(self.h, {"exact": 0, "fast": 1}[precision]), "set_precision")

    def gemm(self, A: np.ndarray, B: np.ndarray, bn: int = 128) -> np.ndarray:
        A = np.ascontiguousarray(A, np.float32)
        B = np.ascontiguousarray(B, np.float32)
        M, K = A.shape
        N = B.shape[0]
        Cm = np.zeros((M, N), np.float32)
        self.check(self.lib.dimb_selftest_gemm(self.h, _ptr(A), _ptr(B), _ptr(Cm), M, N, K, bn), "selftest_gemm")
        return Cm

    def __del__(self):
        try:
            self.lib.dimb_ctx_destroy(self.h)
        except Exception:
            pass


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


class Context:
    """One per device (dimb_ctx). precision: "exact" (fp16 hi/lo split, fp32-class) or "fast" (plain fp16)."""

    _per_device: dict = {}

    def __init__(self, device: int = 0, precision: str | None = None, tensor_path: bool | None = None):
        self.lib = load_library()
        h = C.c_void_p()
        rc = self.lib.dimb_ctx_create(device, C.byref(h))
        if rc != OK:
            raise DimbError(f"dimb_ctx_create(device={device}) failed with code {rc}: a CUDA device of compute "
                            "capability 10.x (B200, sm_100a) is required; there is no CPU fallback")
        self.h = h
        self.device = device
        if precision is not None:
            self.set_precision(precision)
        if tensor_path is not None:
            self.set_tensor_path(tensor_path)

    @classmethod
    def get(cls, device: int = 0) -> "Context":
        if device not in cls._per_device:
            cls._per_device[device] = cls(device)
        return cls._per_device[device]

    def check(self, rc: int, what: str):
        if rc != OK:
            msg = self.lib.dimb_last_error(self.h).decode()
            raise DimbError(f"{what} failed (code {rc}): {msg}")

    def set_precision(self, precision: str):
        self.check(self.lib.dimb_ctx_set_precision(self.h, {"exact": 0, "fast": 1}[precision]), "set_precision")

    def set_tensor_path(self, use_tc: bool):
        self.check(self.lib.dimb_ctx_set_tensor_path(self.h, int(bool(use_tc))), "set_tensor_path")

    @property
    def launches(self) -> int:
        return int(self.lib.dimb_ctx_launch_count(self.h))

    def profile(self, enable: bool):
        self.check(self.lib.dimb_ctx_profile(self.h, int(bool(enable))), "dimb_ctx_profile")

    def profile_read(self) -> dict:
        """{"group": [total_ms, launches]} of the kernel groups recorded since profile(True)."""
        import json
        buf = C.create_string_buffer(1 << 16)
        self.check(self.lib.dimb_ctx_profile_read(self.h, buf, len(buf)), "dimb_ctx_profile_read")
        return json.loads(buf.value.decode())

    def nn_match(self, desc0: np.ndarray, desc1: np.ndarray, mode: str = "smnn", th: float = 0.8):
        """desc0 (D,n0), desc1 (D,n1) float32 -> (int64 (S,2), float32 (S,))."""
        d0 = np.ascontiguousarray(desc0, np.float32)
        d1 = np.ascontiguousarray(desc1, np.float32)
        D, n0 = d0.shape
        n1 = d1.shape[1]
        cap = max(n0, n1, 1)
        idx = np.zeros((cap, 2), np.int64)
        dist = np.zeros(cap, np.float32)
        n = C.c_int(0)
        self.check(self.lib.dimb_nn_match(self.h, _ptr(d0), n0, _ptr(d1), n1, D, NN_MODES[mode], float(th), _ptr(idx),
                                          _ptr(dist), C.byref(n), cap), "dimb_nn_match")
        return idx[: n.value].copy(), dist[: n.value].copy()

    def gv_fundamental(self, kpts0: np.ndarray, kpts1: np.ndarray, threshold: float = 1.0, max_iters: int = 10000, seed: int = 0):
        """Matched keypoints (n,2) each -> (F (3,3) float32 or None, inlier mask bool (n,))."""
        k0 = np.ascontiguousarray(kpts0, np.float32)
        k1 = np.ascontiguousarray(kpts1, np.float32)
        n = k0.shape[0]
        F = np.zeros(9, np.float32)
        mask = np.ones(max(n, 1), np.uint8)
        cnt = C.c_int(0)
        self.check(self.lib.dimb_gv_fundamental(self.h, _ptr(k0), _ptr(k1), n, float(threshold), int(max_iters), int(seed) & 0xffffffff, _ptr(F),
                                                _ptr(mask), C.byref(cnt)), "dimb_gv_fundamental")
        return (F.reshape(3, 3) if np.any(F) else None), mask[:n].astype(bool)

    def nn_match_dev(self, d_desc0: int, n0: int, d_desc1: int, n1: int, D: int, mode: str, th: float, d_idx: int, d_dist: int,
                     d_n: int, cap: int, f16: bool = False, ld0: int = 0, ld1: int = 0, stream: int = 0):
        """Device-pointer variant (ints are device addresses); asynchronous on `stream`."""
        self.check(self.lib.dimb_nn_match_dev(self.h, d_desc0, n0, ld0, d_desc1, n1, ld1, D, int(f16), NN_MODES[mode], float(th), d_idx,
                                              d_dist, d_n, cap, stream), "dimb_nn_match_dev")


SP_ORDER = ["conv1a", "conv1b", "conv2a", "conv2b", "conv3a", "conv3b", "conv4a", "conv4b", "convPa", "convPb",
            "convDa", "convDb"]


def pack_superpoint_weights(w: dict) -> np.ndarray:
    parts = []
    for name in SP_ORDER:
        parts += [np.asarray(w[name + ".weight"], np.float32).ravel(), np.asarray(w[name + ".bias"], np.float32).ravel()]
    return np.ascontiguousarray(np.concatenate(parts))


def lightglue_weight_names(input_dim: int, descriptor_dim: int, n_layers: int) -> list:
    names = ["posenc.Wr.weight"]
    if input_dim != descriptor_dim:
        names += ["input_proj.weight", "input_proj.bias"]
    for i in range(n_layers):
        p = f"transformers.{i}."
        for m in ("self_attn.Wqkv", "self_attn.out_proj", "self_attn.ffn.0", "self_attn.ffn.1", "self_attn.ffn.3",
                  "cross_attn.to_qk", "cross_attn.to_v", "cross_attn.to_out", "cross_attn.ffn.0", "cross_attn.ffn.1",
                  "cross_attn.ffn.3"):
            names += [p + m + ".weight", p + m + ".bias"]
    for i in range(n_layers):
        for m in ("matchability", "final_proj"):
            names += [f"log_assignment.{i}.{m}.weight", f"log_assignment.{i}.{m}.bias"]
    for i in range(n_layers - 1):
        names += [f"token_confidence.{i}.token.0.weight", f"token_confidence.{i}.token.0.bias"]
    return names


def pack_lightglue_weights(w: dict, input_dim: int, descriptor_dim: int, n_layers: int) -> np.ndarray:
    w = dict(w)
    for i in range(n_layers):  # old checkpoints: self_attn.i. / cross_attn.i. prefixes (lightglue.py:391-396)
        for blk in ("self_attn", "cross_attn"):
            for k in [k for k in w if k.startswith(f"{blk}.{i}.")]:
                w[k.replace(f"{blk}.{i}", f"transformers.{i}.{blk}", 1)] = w.pop(k)
    return np.ascontiguousarray(np.concatenate(
        [np.asarray(w[n], np.float32).ravel() for n in lightglue_weight_names(input_dim, descriptor_dim, n_layers)]))


class SuperPointNet:
    """Handle on dimb_sp: SuperPoint extraction of batches of equally sized gray images."""

    def __init__(self, ctx: Context, weights: dict, nms_radius=4, keypoint_threshold=0.005, max_keypoints=-1,
                 remove_borders=4, fix_sampling=False, max_batch=1, max_height=1024, max_width=1024):
        self.ctx = ctx
        if max_keypoints == 0 or max_keypoints < -1:
            raise ValueError('"max_keypoints" must be positive or "-1"')  # superpoint.py:152-154
        self.conf = SpConf(int(nms_radius), float(keypoint_threshold), int(max_keypoints), int(remove_borders),
                           int(bool(fix_sampling)), int(max_batch), int(max_height), int(max_width))
        blob = pack_superpoint_weights(weights)
        h = C.c_void_p()
        ctx.check(ctx.lib.dimb_sp_create(ctx.h, _ptr(blob), blob.size, C.byref(self.conf), C.byref(h)), "dimb_sp_create")
        self.h = h

    def default_cap(self, H, W):
        k = self.conf.max_keypoints
        return k if k > 0 else (H // 8 * 8) * (W // 8 * 8) // 4

    def extract(self, images: np.ndarray, cap: int | None = None):
        """images float32 (B,H,W) 0..255 -> list of dicts(keypoints (N,2), scores (N,), descriptors (256,N))."""
        images = np.ascontiguousarray(images, np.float32)
        B, H, W = images.shape
        cap = cap or self.default_cap(H, W)
        while True:
            kp = np.zeros((B, cap, 2), np.float32)
            sc = np.zeros((B, cap), np.float32)
            de = np.zeros((B, 256, cap), np.float32)
            cnt = np.zeros(B, np.int32)
            rc = self.ctx.lib.dimb_sp_extract(self.h, _ptr(images), B, H, W, _ptr(kp), _ptr(sc), _ptr(de), _ptr(cnt), cap)
            if rc == ERR_CAPACITY:
                cap = int(cnt.max())
                continue
            self.ctx.check(rc, "dimb_sp_extract")
            break
        return [{"keypoints": kp[b, : cnt[b]].copy(), "scores": sc[b, : cnt[b]].copy(),
                 "descriptors": de[b, :, : cnt[b]].copy()} for b in range(B)]

    def extract_dev(self, d_images, B, H, W, d_kpts, d_scores, d_desc, d_counts, cap, stream=0):
        """Raw device-pointer variant (ints are device addresses, e.g. torch.Tensor.data_ptr())."""
        self.ctx.check(self.ctx.lib.dimb_sp_extract_dev(self.h, d_images, B, H, W, d_kpts, d_scores, d_desc, d_counts, cap,
                                                        stream), "dimb_sp_extract_dev")

    def debug_read(self, which: int, shape) -> np.ndarray:
        out = np.zeros(shape, np.float32)
        self.ctx.check(self.ctx.lib.dimb_sp_debug_read(self.h, which, _ptr(out), out.size), "dimb_sp_debug_read")
        return out

    def __del__(self):
        try:
            self.ctx.lib.dimb_sp_destroy(self.h)
        except Exception:
            pass


def aliked_weight_names() -> list:
    """state_dict order of aliked-n16 / aliked-n16rot (thirdparty/LightGlue/lightglue/aliked.py:596-640), floats only."""
    names = []
    bn = lambda p: [p + s for s in (".weight", ".bias", ".running_mean", ".running_var")]
    names += ["block1.conv1.weight"] + bn("block1.bn1") + ["block1.conv2.weight"] + bn("block1.bn2")
    names += ["block2.conv1.weight"] + bn("block2.bn1") + ["block2.conv2.weight"] + bn("block2.bn2")
    names += ["block2.downsample.weight", "block2.downsample.bias"]
    for b in ("block3", "block4"):
        for c, n in (("conv1", "bn1"), ("conv2", "bn2")):
            names += [f"{b}.{c}.offset_conv.weight", f"{b}.{c}.offset_conv.bias", f"{b}.{c}.regular_conv.weight"] + bn(f"{b}.{n}")
        names += [f"{b}.downsample.weight", f"{b}.downsample.bias"]
    names += ["conv1.weight", "conv2.weight", "conv3.weight", "conv4.weight"]
    names += [f"score_head.{i}.weight" for i in (0, 2, 4, 6)]
    names += ["desc_head.agg_weights", "desc_head.offset_conv.0.weight", "desc_head.offset_conv.0.bias",
              "desc_head.offset_conv.2.weight", "desc_head.offset_conv.2.bias", "desc_head.sf_conv.weight"]
    return names


def pack_aliked_weights(w: dict) -> np.ndarray:
    return np.ascontiguousarray(np.concatenate([np.asarray(w[n], np.float32).ravel() for n in aliked_weight_names()]))


class AlikedNet:
    """Handle on dimb_aliked: ALIKED-n16(rot) extraction of one image per call (the reference path is batch-1)."""

    def __init__(self, ctx: Context, weights: dict, max_num_keypoints=4000, detection_threshold=0.2, nms_radius=2,
                 max_height=1024, max_width=1024):
        self.ctx = ctx
        self.conf = AlikedConf(int(max_num_keypoints), float(detection_threshold), int(nms_radius), int(max_height), int(max_width))
        blob = pack_aliked_weights(weights)
        h = C.c_void_p()
        ctx.check(ctx.lib.dimb_aliked_create(ctx.h, _ptr(blob), blob.size, C.byref(self.conf), C.byref(h)), "dimb_aliked_create")
        self.h = h

    def extract(self, image: np.ndarray, cap: int | None = None) -> dict:
        """image float32 (H,W,3) RGB or (H,W) gray, 0..255 -> keypoints (N,2), scores (N,), descriptors (128,N)."""
        image = np.ascontiguousarray(image, np.float32)
        H, W = image.shape[:2]
        ch = 1 if image.ndim == 2 else image.shape[2]
        k = self.conf.max_num_keypoints
        cap = cap or (k if k > 0 else 16384)
        while True:
            kp = np.zeros((cap, 2), np.float32)
            sc = np.zeros(cap, np.float32)
            de = np.zeros((128, cap), np.float32)
            cnt = np.zeros(1, np.int32)
            rc = self.ctx.lib.dimb_aliked_extract(self.h, _ptr(image), H, W, ch, _ptr(kp), _ptr(sc), _ptr(de), _ptr(cnt), cap)
            if rc == ERR_CAPACITY:
                cap = int(cnt[0])
                continue
            self.ctx.check(rc, "dimb_aliked_extract")
            break
        n = int(cnt[0])
        return {"keypoints": kp[:n].copy(), "scores": sc[:n].copy(), "descriptors": de[:, :n].copy()}

    def extract_dev(self, d_image, H, W, channels, d_kpts, d_scores, d_desc, d_count, cap, stream=0):
        """Raw device-pointer variant (ints are device addresses, e.g. torch.Tensor.data_ptr())."""
        self.ctx.check(self.ctx.lib.dimb_aliked_extract_dev(self.h, d_image, H, W, channels, d_kpts, d_scores, d_desc, d_count, cap,
                                                            stream), "dimb_aliked_extract_dev")

    def debug_read(self, which: int, shape) -> np.ndarray:
        out = np.zeros(shape, np.float32)
        self.ctx.check(self.ctx.lib.dimb_aliked_debug_read(self.h, which, _ptr(out), out.size), "dimb_aliked_debug_read")
        return out

    def __del__(self):
        try:
            self.ctx.lib.dimb_aliked_destroy(self.h)
        except Exception:
            pass


def superglue_weight_names(n_layers: int = 18) -> list:
    bn = lambda p: [p + s for s in (".weight", ".bias", ".running_mean", ".running_var")]
    names = []
    for i in range(5):
        names += [f"kenc.encoder.{3 * i}.weight", f"kenc.encoder.{3 * i}.bias"]
        if i < 4:
            names += bn(f"kenc.encoder.{3 * i + 1}")
    for i in range(n_layers):
        p = f"gnn.layers.{i}."
        names += [p + "attn.merge.weight", p + "attn.merge.bias"]
        for j in range(3):
            names += [p + f"attn.proj.{j}.weight", p + f"attn.proj.{j}.bias"]
        names += [p + "mlp.0.weight", p + "mlp.0.bias"] + bn(p + "mlp.1") + [p + "mlp.3.weight", p + "mlp.3.bias"]
    return names + ["final_proj.weight", "final_proj.bias", "bin_score"]


class SuperGlueNet:
    """Handle on dimb_sg: SuperGlue matching of one pair per call."""

    def __init__(self, ctx: Context, weights: dict, gnn_layers=("self", "cross") * 9, sinkhorn_iterations=100, match_threshold=0.2,
                 max_kpts=2048):
        self.ctx = ctx
        mask = sum(1 << i for i, n in enumerate(gnn_layers) if n == "cross")
        self.conf = SgConf(len(gnn_layers), mask, int(sinkhorn_iterations), float(match_threshold), int(max_kpts))
        blob = np.ascontiguousarray(np.concatenate([np.asarray(weights[n], np.float32).ravel() for n in superglue_weight_names(len(gnn_layers))]))
        h = C.c_void_p()
        ctx.check(ctx.lib.dimb_sg_create(ctx.h, _ptr(blob), blob.size, C.byref(self.conf), C.byref(h)), "dimb_sg_create")
        self.h = h

    def match(self, feats0: dict, feats1: dict) -> dict:
        """feats: keypoints (N,2), descriptors (256,N), scores (N,), image_size [H,W] -> matches int64 (S,2), scores (S,)."""
        fs, keep = [], []
        for f in (feats0, feats1):
            k = np.ascontiguousarray(f["keypoints"], np.float32)
            d = np.ascontiguousarray(f["descriptors"], np.float32)
            s = np.ascontiguousarray(f["scores"], np.float32)
            if d.shape != (256, k.shape[0]):
                raise ValueError(f"SuperGlue expects (256,N) descriptors, got {d.shape} for {k.shape[0]} keypoints")
            hw = np.asarray(f["image_size"]).astype(int).ravel()
            keep += [k, d, s]
            fs.append(SgFeats(k.ctypes.data, d.ctypes.data, s.ctypes.data, k.shape[0], 0, int(hw[0]), int(hw[1])))
        cap = max(1, min(fs[0].n, fs[1].n))
        m = np.zeros((cap, 2), np.int64)
        sc = np.zeros(cap, np.float32)
        n = C.c_int(0)
        self.ctx.check(self.ctx.lib.dimb_sg_match(self.h, C.byref(fs[0]), C.byref(fs[1]), _ptr(m), _ptr(sc), C.byref(n), cap), "dimb_sg_match")
        return {"matches": m[: n.value].copy(), "scores": sc[: n.value].copy()}

    def __del__(self):
        try:
            self.ctx.lib.dimb_sg_destroy(self.h)
        except Exception:
            pass


class LightGlueNet:
    """Handle on dimb_lg: LightGlue matching of batches of pairs."""

    def __init__(self, ctx: Context, weights: dict, input_dim=256, descriptor_dim=256, n_layers=9, num_heads=4,
                 depth_confidence=0.95, width_confidence=0.99, filter_threshold=0.1, prune_min_kpts=1536, max_pairs=1,
                 max_kpts=2048):
        self.ctx = ctx
        self.conf = LgConf(int(input_dim), int(descriptor_dim), int(n_layers), int(num_heads), float(depth_confidence),
                           float(width_confidence), float(filter_threshold), int(prune_min_kpts), int(max_pairs),
                           int(max_kpts))
        blob = pack_lightglue_weights(weights, input_dim, descriptor_dim, n_layers)
        h = C.c_void_p()
        ctx.check(ctx.lib.dimb_lg_create(ctx.h, _ptr(blob), blob.size, C.byref(self.conf), C.byref(h)), "dimb_lg_create")
        self.h = h
        self.NP = (int(max_kpts) + 127) // 128 * 128

    def match(self, pairs):
        """pairs: list of (feats0, feats1) with keypoints (N,2), descriptors (N,D) [layout 1] or (D,N) [layout 0]
        already decided by the caller via key "_layout"; image_size optional.
        Returns list of dict(matches int64 (S,2), scores (S,), stop int)."""
        P = len(pairs)
        f0 = (Feats * P)()
        f1 = (Feats * P)()
        keep = []
        cap = 1
        for p, (a, b) in enumerate(pairs):
            for arr, f in ((f0, a), (f1, b)):
                k = np.ascontiguousarray(f["keypoints"], np.float32)
                d = np.ascontiguousarray(f["descriptors"], np.float32)
                keep += [k, d]
                e = arr[p]
                e.keypoints, e.descriptors = k.ctypes.data, d.ctypes.data
                e.n = k.shape[0]
                e.desc_layout = int(f.get("_layout", 1))
                e.desc_ld = 0
                size = f.get("image_size")
                e.has_size = int(size is not None)
                if size is not None:
                    size = np.asarray(size, np.float32).ravel()
                    e.size0, e.size1 = float(size[0]), float(size[1])
            cap = max(cap, min(a["keypoints"].shape[0], b["keypoints"].shape[0]))
        m = np.zeros((P, cap, 2), np.int64)
        s = np.zeros((P, cap), np.float32)
        nm = np.zeros(P, np.int32)
        sl = np.zeros(P, np.int32)
        self.ctx.check(self.ctx.lib.dimb_lg_match(self.h, P, f0, f1, _ptr(m), _ptr(s), _ptr(nm), _ptr(sl), cap),
                       "dimb_lg_match")
        return [{"matches": m[p, : nm[p]].copy(), "scores": s[p, : nm[p]].copy(), "stop": int(sl[p])} for p in range(P)]

    def match_dev(self, f0: list, f1: list, d_matches, d_mscores, d_n_matches, d_stop, cap, stream=0):
        """f0/f1: lists of FeatsDev (device pointers)."""
        P = len(f0)
        a0 = (FeatsDev * P)(*f0)
        a1 = (FeatsDev * P)(*f1)
        self.ctx.check(self.ctx.lib.dimb_lg_match_dev(self.h, P, a0, a1, d_matches, d_mscores, d_n_matches, d_stop, cap,
                                                      stream), "dimb_lg_match_dev")

    def debug_read(self, which: int, side: int, shape) -> np.ndarray:
        out = np.zeros(shape, np.float32)
        self.ctx.check(self.ctx.lib.dimb_lg_debug_read(self.h, which, side, _ptr(out), out.size), "dimb_lg_debug_read")
        return out

    def __del__(self):
        try:
            self.ctx.lib.dimb_lg_destroy(self.h)
        except Exception:
            pass


class FeatureStoreDev:
    """Handle on dimb_fstore: the content of features.h5 (float16 arrays + int image_size, one block per image) kept in HBM."""

    def __init__(self, ctx: Context, n_slots: int, cap: int, desc_dim: int):
        self.ctx, self.n_slots, self.desc_dim = ctx, int(n_slots), int(desc_dim)
        h = C.c_void_p()
        ctx.check(ctx.lib.dimb_fstore_create(ctx.h, n_slots, cap, desc_dim, C.byref(h)), "dimb_fstore_create")
        self.h = h
        base, sb, ns, cp = C.c_void_p(), C.c_size_t(), C.c_int(), C.c_int()
        ctx.check(ctx.lib.dimb_fstore_block_dev(h, C.byref(base), C.byref(sb), C.byref(ns), C.byref(cp)), "dimb_fstore_block_dev")
        self.base, self.slot_bytes, self.cap = base.value, sb.value, cp.value

    def put_dev(self, slot, d_kpts, d_scores, d_desc, desc_ld, d_count, height, width, d_tile_idx=None, stream=0):
        self.ctx.check(self.ctx.lib.dimb_fstore_put_dev(self.h, slot, d_kpts, d_scores, d_tile_idx, d_desc, desc_ld, d_count, int(height),
                                                        int(width), stream), "dimb_fstore_put_dev")

    def put(self, slot: int, feats: dict):
        """feats: FeaturesDict (keypoints (N,2), descriptors (D,N), optional scores / tile_idx, image_size [H,W])."""
        k = np.ascontiguousarray(feats["keypoints"], np.float32)
        d = np.ascontiguousarray(feats["descriptors"], np.float32)
        n = k.shape[0]
        if d.shape != (self.desc_dim, n):
            raise ValueError(f"descriptors must be ({self.desc_dim},{n}), got {d.shape}")
        s = np.ascontiguousarray(feats["scores"], np.float32) if feats.get("scores") is not None else None
        t = np.ascontiguousarray(feats["tile_idx"], np.float32) if feats.get("tile_idx") is not None else None
        hw = np.asarray(feats.get("image_size", (0, 0))).astype(int).ravel()
        self.ctx.check(self.ctx.lib.dimb_fstore_put(self.h, slot, _ptr(k), _ptr(s) if s is not None else None,
                                                    _ptr(t) if t is not None else None, _ptr(d), n, int(hw[0]), int(hw[1])), "dimb_fstore_put")

    def count(self, slot: int):
        n, size = C.c_int(), np.zeros(2, np.int32)
        self.ctx.check(self.ctx.lib.dimb_fstore_count(self.h, slot, C.byref(n), _ptr(size)), "dimb_fstore_count")
        return n.value, size

    def get(self, slot: int) -> dict:
        """The FeaturesDict get_features (io/h5.py:45-89) would return for this image."""
        n, size = self.count(slot)
        if n < 0:
            raise ValueError(f"slot {slot} of the feature store is empty")
        k, s, t = np.zeros((n, 2), np.float32), np.zeros(n, np.float32), np.zeros(n, np.float32)
        d = np.zeros((self.desc_dim, n), np.float32)
        cnt = C.c_int()
        self.ctx.check(self.ctx.lib.dimb_fstore_get(self.h, slot, _ptr(k), _ptr(s), _ptr(t), _ptr(d), C.byref(cnt), _ptr(size), max(n, 1)),
                       "dimb_fstore_get")
        return {"keypoints": k, "descriptors": d, "scores": s, "tile_idx": t, "image_size": size.astype(np.int32)}

    def feats_dev(self, slot: int) -> FeatsDev:
        f = FeatsDev()
        self.ctx.check(self.ctx.lib.dimb_fstore_feats_dev(self.h, slot, C.byref(f)), "dimb_fstore_feats_dev")
        return f

    def desc_ptr(self, slot: int) -> int:
        """Device address of the slot's float16 (D, cap) descriptor block (for dimb_nn_match_dev, ld = cap)."""
        return self.feats_dev(slot).descriptors

    def __del__(self):
        try:
            self.ctx.lib.dimb_fstore_destroy(self.h)
        except Exception:
            pass


class Pipe:
    """Fused per-pair path (dimb_pipe): images -> SuperPoint x2 -> LightGlue, features stay in HBM."""

    def __init__(self, sp: SuperPointNet, lg: LightGlueNet, max_pairs: int, H: int, W: int, cap: int):
        self.ctx, self.sp, self.lg = sp.ctx, sp, lg
        self.max_pairs, self.H, self.W, self.cap = max_pairs, H, W, cap
        h = C.c_void_p()
        self.ctx.check(self.ctx.lib.dimb_pipe_create(sp.h, lg.h, max_pairs, H, W, cap, C.byref(h)), "dimb_pipe_create")
        self.h = h

    def match_image_pairs(self, images: np.ndarray, out: dict | None = None, want_kpts: bool = False) -> dict:
        """images: float32 or uint8 (2P,H,W) gray host array (pinned for async copies). Returns host arrays."""
        B = images.shape[0]
        P = B // 2
        if out is None:
            out = self.alloc_outputs(P, want_kpts)
        kp = out.get("kpts")
        fn = self.ctx.lib.dimb_pipe_match_image_pairs_u8 if images.dtype == np.uint8 else self.ctx.lib.dimb_pipe_match_image_pairs
        assert images.dtype in (np.uint8, np.float32) and images.flags.c_contiguous
        self.ctx.check(fn(
            self.h, images.ctypes.data, P, out["matches"].ctypes.data, out["mscores"].ctypes.data,
            out["n_matches"].ctypes.data, out["stop"].ctypes.data, out["n_kpts"].ctypes.data,
            kp.ctypes.data if kp is not None else None), "dimb_pipe_match_image_pairs")
        return out

    def alloc_outputs(self, P: int, want_kpts: bool = False) -> dict:
        out = {"matches": np.zeros((P, self.cap, 2), np.int64), "mscores": np.zeros((P, self.cap), np.float32),
               "n_matches": np.zeros(P, np.int32), "stop": np.zeros(P, np.int32), "n_kpts": np.zeros(2 * P, np.int32)}
        if want_kpts:
            out["kpts"] = np.zeros((2 * P, self.cap, 2), np.float32)
        return out

    def match_image_pairs_dev(self, d_images: int, P: int, stream: int = 0):
        self.ctx.check(self.ctx.lib.dimb_pipe_match_image_pairs_dev(self.h, d_images, P, stream),
                       "dimb_pipe_match_image_pairs_dev")

    def outputs_dev(self) -> dict:
        ptrs = [C.c_void_p() for _ in range(6)]
        self.ctx.check(self.ctx.lib.dimb_pipe_outputs_dev(self.h, *[C.byref(p) for p in ptrs]), "dimb_pipe_outputs_dev")
        return dict(zip(["matches", "mscores", "n_matches", "stop", "n_kpts", "kpts"], [p.value for p in ptrs]))

    def features_dev(self) -> dict:
        ptrs = [C.c_void_p() for _ in range(4)]
        self.ctx.check(self.ctx.lib.dimb_pipe_features_dev(self.h, *[C.byref(p) for p in ptrs]), "dimb_pipe_features_dev")
        return dict(zip(["kpts", "scores", "desc", "counts"], [p.value for p in ptrs]))

    def read_features(self, P: int) -> list:
        """Host copies of the last call's SuperPoint features (2P FeaturesDicts, float32, before the fp16 cast)."""
        d = self.features_dev()
        B, cap = 2 * P, self.cap
        kp, sc = np.zeros((B, cap, 2), np.float32), np.zeros((B, cap), np.float32)
        de, cnt = np.zeros((B, 256, cap), np.float32), np.zeros(B, np.int32)
        for dst, src in ((kp, d["kpts"]), (sc, d["scores"]), (de, d["desc"]), (cnt, d["counts"])):
            self.ctx.check(self.ctx.lib.dimb_read_dev(self.ctx.h, dst.ctypes.data, src, dst.nbytes), "dimb_read_dev")
        return [{"keypoints": kp[b, :cnt[b]].copy(), "scores": sc[b, :cnt[b]].copy(), "descriptors": de[b, :, :cnt[b]].copy()}
                for b in range(B)]

    def __del__(self):
        try:
            self.ctx.lib.dimb_pipe_destroy(self.h)
        except Exception:
            pass
