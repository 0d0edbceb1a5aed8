"""Python host mirror of the C-ABI in include/percepnet_b200.h (ctypes; no compute here).

``Engine`` is the batched-streams counterpart of the reference's ``DenoiseState``
(/root/reference/src/rnnoise.h:49-60): ``Engine.process`` advances S independent 48 kHz streams by F
hops of 480 samples, i.e. S x F calls of ``rnnoise_process_frame``.  The CUDA library is mandatory:
there is no CPU path, and a missing ``libpercepnet_b200.so`` raises at import of the symbol table.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from .weights import PackedModel

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libpercepnet_b200.so")

FRAME = 480
NN_FP32, NN_TENSOR, POSTFILTER, KEEP_TAPS, TRAIN_DATA, CONV_WIDE = 0, 1, 2, 4, 8, 16
RECORD = 138
TAPS = {"features": (0, np.float32, 70), "pitch": (1, np.int32, 4), "pitchf": (2, np.float32, 2),
        "X": (3, np.float32, 800), "P": (4, np.float32, 800), "Ex": (5, np.float32, 34), "gr": (6, np.float32, 68),
        "g_used": (13, np.float32, 34)}

_lib = None


class PnbError(RuntimeError):
    pass


def load_library() -> C.CDLL:
    """dlopen the in-tree CUDA library; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise PnbError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                       "(nvcc, sm_100a).  There is no CPU implementation to fall back to.")
    L = C.CDLL(LIB_PATH)
    vp, sz, i = C.c_void_p, C.c_size_t, C.c_int
    L.pnb_create.argtypes = [C.POINTER(vp), i, i, vp, C.c_uint, i]
    L.pnb_model_load_blob.argtypes = [C.c_char_p, C.POINTER(vp)]
    L.pnb_model_free.argtypes = [vp]
    L.pnb_model_free.restype = None
    L.pnb_train_records_host.argtypes = [vp, vp, sz, vp, sz, i, vp, sz]
    L.pnb_submit_train_records.argtypes = [vp, vp, sz, vp, sz, i, vp, sz]
    L.pnb_train_records_device.argtypes = [vp, vp, sz, vp, sz, i, vp, sz, vp]
    L.pnb_destroy.argtypes = [vp]
    L.pnb_destroy.restype = None
    L.pnb_reset.argtypes = [vp]
    L.pnb_process_host_f32.argtypes = [vp, vp, sz, vp, sz, i, vp]
    L.pnb_process_host_i16.argtypes = [vp, vp, sz, vp, sz, i, vp]
    L.pnb_process_device_f32.argtypes = [vp, vp, sz, vp, sz, i, vp, vp]
    L.pnb_process_device_i16.argtypes = [vp, vp, sz, vp, sz, i, vp, vp]
    L.pnb_submit_host_f32.argtypes = [vp, vp, sz, vp, sz, i]
    L.pnb_submit_host_i16.argtypes = [vp, vp, sz, vp, sz, i]
    L.pnb_submit_device_f32.argtypes = [vp, vp, sz, vp, sz, i, vp]
    L.pnb_submit_device_i16.argtypes = [vp, vp, sz, vp, sz, i, vp]
    L.pnb_flush.argtypes = [vp, vp]
    L.pnb_wait.argtypes = [vp]
    L.pnb_check.argtypes = [vp, vp]
    L.pnb_read_tap.argtypes = [vp, i, vp, sz]
    L.pnb_state_size.restype = sz
    L.pnb_get_state.argtypes = [vp, i, vp, sz]
    L.pnb_set_state.argtypes = [vp, i, vp, sz]
    L.pnb_pitch_only_device.argtypes = [vp, sz, C.c_longlong, vp, vp, vp, vp, vp, vp, vp]
    L.pnb_pitch_only_host.argtypes = [vp, sz, C.c_longlong, vp, vp, vp, vp, vp, vp]
    L.pnb_launch_count.argtypes = [vp]
    L.pnb_launch_count.restype = C.c_longlong
    L.pnb_launches_per_call.argtypes = [vp, i]
    L.pnb_n_streams.argtypes = [vp]
    L.pnb_overlap_info.argtypes = [vp, vp, vp, vp]
    L.pnb_set_overlap.argtypes = [vp, i]
    L.pnb_max_frames.argtypes = [vp]
    L.pnb_profile_enable.argtypes = [vp, i]
    L.pnb_profile_read.argtypes = [vp, vp, vp]
    L.pnb_profile_timeline.argtypes = [vp, vp, vp, vp, i]
    L.pnb_kernel_class_name.argtypes = [i]
    L.pnb_kernel_class_name.restype = C.c_char_p
    L.pnb_last_error.restype = C.c_char_p
    L.pnb_version.restype = C.c_char_p
    _lib = L
    return L


EXPORTS = ("pnb_create", "pnb_destroy", "pnb_reset", "pnb_process_host_f32", "pnb_process_host_i16",
           "pnb_model_load_blob", "pnb_model_load_stream", "pnb_model_free", "pnb_train_records_host", "pnb_train_records_device", "pnb_submit_train_records", "pnb_process_device_f32", "pnb_process_device_i16", "pnb_submit_host_f32", "pnb_submit_host_i16", "pnb_submit_device_f32", "pnb_submit_device_i16", "pnb_flush", "pnb_wait", "pnb_check",
           "pnb_read_tap", "pnb_state_size", "pnb_get_state", "pnb_set_state", "pnb_pitch_only_device", "pnb_pitch_only_host", "pnb_launch_count",
           "pnb_launches_per_call", "pnb_profile_enable", "pnb_profile_read", "pnb_profile_timeline", "pnb_kernel_class_name", "pnb_n_streams", "pnb_overlap_info", "pnb_set_overlap", "pnb_max_frames", "pnb_last_error", "pnb_version")


def pitch_only_device(d_buf: int, stride: int, n_units: int, d_period: int, d_corr: int, d_gain: int, d_lag: int = 0,
                      d_prev_period: int = 0, d_prev_gain: int = 0, stream: int = 0):
    """pnb_pitch_only_device on raw device pointers (BASELINE.json config 5: the pitch analysis alone)."""
    L = load_library()
    rc = L.pnb_pitch_only_device(d_buf, stride, n_units, d_prev_period or None, d_prev_gain or None, d_period, d_corr,
                                 d_gain, d_lag or None, stream or None)
    if rc != 0:
        raise PnbError(f"pnb_pitch_only_device failed ({rc}): {L.pnb_last_error().decode()}")


def pitch_only(bufs: np.ndarray, prev_period=None, prev_gain=None):
    """pnb_pitch_only_host: [n, 1728] float32 pitch buffers -> (period int32 [n], corr [n], gain [n], lag int32 [n])"""
    L = load_library()
    b = np.ascontiguousarray(bufs, np.float32)
    n = b.shape[0]
    T, lag = np.empty(n, np.int32), np.empty(n, np.int32)
    corr, gain = np.empty(n, np.float32), np.empty(n, np.float32)
    pp = None if prev_period is None else np.ascontiguousarray(prev_period, np.int32)
    pg = None if prev_gain is None else np.ascontiguousarray(prev_gain, np.float32)
    rc = L.pnb_pitch_only_host(b.ctypes.data, b.shape[1], n, None if pp is None else pp.ctypes.data,
                               None if pg is None else pg.ctypes.data, T.ctypes.data, corr.ctypes.data, gain.ctypes.data,
                               lag.ctypes.data)
    if rc != 0:
        raise PnbError(f"pnb_pitch_only_host failed ({rc}): {L.pnb_last_error().decode()}")
    return T, corr, gain, lag


class BlobModel:
    """A model read from a binary weight file by the C loader (pnb_model_load_blob)."""

    def __init__(self, path: str):
        self.L = load_library()
        self.ptr = C.c_void_p()
        rc = self.L.pnb_model_load_blob(path.encode(), C.byref(self.ptr))
        if rc != 0:
            raise PnbError(f"pnb_model_load_blob failed ({rc}): {self.L.pnb_last_error().decode()}")

    def free(self):
        if self.ptr:
            self.L.pnb_model_free(self.ptr)
            self.ptr = C.c_void_p()


class Engine:
    def __init__(self, n_streams: int, max_frames: int, model, flags: int = NN_FP32, device: int = 0):
        self.L = load_library()
        self.n_streams, self.max_frames, self.flags, self.device = n_streams, max_frames, flags, device
        self._model = model
        h = C.c_void_p()
        mptr = None if model is None else (model.ptr if isinstance(model, BlobModel) else C.addressof(model.as_c_model()))
        rc = self.L.pnb_create(C.byref(h), n_streams, max_frames, mptr, flags, device)
        if rc != 0:
            raise PnbError(f"pnb_create failed ({rc}): {self.L.pnb_last_error().decode()}")
        self.h = h

    def _ck(self, rc, what):
        if rc != 0:
            raise PnbError(f"{what} failed ({rc}): {self.L.pnb_last_error().decode()}")

    def close(self):
        if getattr(self, "h", None):
            self.L.pnb_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def reset(self):
        self._ck(self.L.pnb_reset(self.h), "pnb_reset")

    def process(self, x: np.ndarray, want_gr: bool = False):
        """x: [S, F*480] float32 (C-API scale) or int16 (CLI wire format) host array.
        Returns (out like x, gr [F, S, 68] or None)."""
        assert x.ndim == 2 and x.shape[0] == self.n_streams and x.shape[1] % FRAME == 0
        F = x.shape[1] // FRAME
        x = np.ascontiguousarray(x)
        out = np.empty_like(x)
        gr = np.empty((F, self.n_streams, 68), np.float32) if want_gr else None
        grp = gr.ctypes.data if want_gr else None
        if x.dtype == np.float32:
            rc = self.L.pnb_process_host_f32(self.h, x.ctypes.data, x.shape[1], out.ctypes.data, x.shape[1], F, grp)
        elif x.dtype == np.int16:
            rc = self.L.pnb_process_host_i16(self.h, x.ctypes.data, x.shape[1], out.ctypes.data, x.shape[1], F, grp)
        else:
            raise TypeError("x must be float32 or int16")
        self._ck(rc, "pnb_process_host")
        return out, gr

    def process_stream_chunks(self, x: np.ndarray, want_gr: bool = False):
        """Any number of hops: walks x in calls of at most max_frames hops."""
        F = x.shape[1] // FRAME
        outs, grs = [], []
        for t0 in range(0, F, self.max_frames):
            t1 = min(F, t0 + self.max_frames)
            o, g = self.process(x[:, t0 * FRAME:t1 * FRAME], want_gr)
            outs.append(o)
            grs.append(g)
        return np.concatenate(outs, axis=1), (np.concatenate(grs, axis=0) if want_gr else None)

    def process_device(self, d_in: int, in_stride: int, d_out: int, out_stride: int, n_frames: int,
                       d_gr: int = 0, stream: int = 0, int16: bool = False):
        f = self.L.pnb_process_device_i16 if int16 else self.L.pnb_process_device_f32
        self._ck(f(self.h, d_in, in_stride, d_out, out_stride, n_frames, d_gr or None, stream or None),
                 "pnb_process_device")

    def submit_device(self, d_in: int, in_stride: int, d_out: int, out_stride: int, n_frames: int, stream: int = 0,
                      int16: bool = False):
        """pnb_submit_device_*: like process_device but not joined into `stream`; flush(stream) or wait() completes it."""
        f = self.L.pnb_submit_device_i16 if int16 else self.L.pnb_submit_device_f32
        self._ck(f(self.h, d_in, in_stride, d_out, out_stride, n_frames, stream or None), "pnb_submit_device")

    def flush(self, stream: int = 0):
        self._ck(self.L.pnb_flush(self.h, stream or None), "pnb_flush")

    def submit(self, x: np.ndarray, out: np.ndarray):
        """Pipelined host call (pnb_submit_host_*): returns at once; x/out ([S, F*480] float32 or int16, ideally
        pinned) must stay alive until wait()."""
        assert x.shape == out.shape and x.dtype == out.dtype and x.shape[0] == self.n_streams
        F = x.shape[1] // FRAME
        f = self.L.pnb_submit_host_f32 if x.dtype == np.float32 else self.L.pnb_submit_host_i16
        self._ck(f(self.h, x.ctypes.data, x.strides[0] // x.itemsize, out.ctypes.data, out.strides[0] // out.itemsize, F),
                 "pnb_submit_host")

    def wait(self):
        self._ck(self.L.pnb_wait(self.h), "pnb_wait")

    def check(self, stream: int = 0):
        """pnb_check: waits for the engine's work (and `stream`) and raises on PNB_ERR_DOMAIN / PNB_ERR_CUDA."""
        self._ck(self.L.pnb_check(self.h, stream or None), "pnb_check")

    def read_tap(self, name: str, n_frames: int) -> np.ndarray:
        code, dt, width = TAPS[name]
        a = np.empty((n_frames, self.n_streams, width), dt)
        self._ck(self.L.pnb_read_tap(self.h, code, a.ctypes.data, a.nbytes), f"pnb_read_tap({name})")
        return a

    def get_state(self, stream: int) -> bytes:
        buf = C.create_string_buffer(self.L.pnb_state_size())
        self._ck(self.L.pnb_get_state(self.h, stream, buf, len(buf)), "pnb_get_state")
        return buf.raw

    def set_state(self, stream: int, blob: bytes):
        self._ck(self.L.pnb_set_state(self.h, stream, blob, len(blob)), "pnb_set_state")

    def read_nn_state(self) -> dict:
        """fp32 network state after the last hop: conv2 output and the five GRU states, [S, width]."""
        out = {}
        for name, code, width in (("c2", 7, 512), ("gru1", 8, 512), ("gru2", 9, 512), ("gru3", 10, 512),
                                  ("gru_gb", 11, 512), ("gru_rb", 12, 128)):
            a = np.empty((self.n_streams, width), np.float32)
            self._ck(self.L.pnb_read_tap(self.h, code, a.ctypes.data, a.nbytes), f"pnb_read_tap({name})")
            out[name] = a
        return out

    def train_records(self, speech: np.ndarray, noisy: np.ndarray) -> np.ndarray:
        """speech, noisy: [n_pairs, F*480] int16 -> [n_pairs, F, 138] float32, the records the reference's train()
        (src/denoise.cpp:600-787) writes for each pair of files.  Engine must have flags=TRAIN_DATA and
        n_streams = 2*n_pairs; state carries over, so call repeatedly with consecutive chunks."""
        N = self.n_streams // 2
        assert speech.shape == noisy.shape and speech.shape[0] == N and speech.shape[1] % FRAME == 0
        assert speech.dtype == np.int16 and noisy.dtype == np.int16
        F = speech.shape[1] // FRAME
        speech, noisy = np.ascontiguousarray(speech), np.ascontiguousarray(noisy)
        rec = np.empty((N, F, RECORD), np.float32)
        self._ck(self.L.pnb_train_records_host(self.h, speech.ctypes.data, speech.shape[1], noisy.ctypes.data,
                                               noisy.shape[1], F, rec.ctypes.data, F * RECORD), "pnb_train_records_host")
        return rec

    def train_records_device(self, d_speech: int, speech_stride: int, d_noisy: int, noisy_stride: int, n_frames: int,
                             d_records: int, records_stride: int, stream: int = 0):
        self._ck(self.L.pnb_train_records_device(self.h, d_speech, speech_stride, d_noisy, noisy_stride, n_frames,
                                                 d_records, records_stride, stream), "pnb_train_records_device")

    def profile(self, on: bool):
        self._ck(self.L.pnb_profile_enable(self.h, 1 if on else 0), "pnb_profile_enable")

    def profile_read(self) -> dict:
        """{kernel class name: (total ms, launches)} since the last read; waits for the device."""
        n = 10
        ms = (C.c_double * n)()
        cnt = (C.c_longlong * n)()
        self._ck(self.L.pnb_profile_read(self.h, ms, cnt), "pnb_profile_read")
        return {self.L.pnb_kernel_class_name(k).decode(): (ms[k], int(cnt[k])) for k in range(n) if cnt[k]}

    def profile_timeline(self, cap: int = 4096):
        """[(class name, start ms, end ms)] of every launch since profiling was enabled / last read."""
        cls = (C.c_int * cap)()
        t0 = (C.c_double * cap)()
        t1 = (C.c_double * cap)()
        n = self.L.pnb_profile_timeline(self.h, cls, t0, t1, cap)
        if n < 0:
            self._ck(n, "pnb_profile_timeline")
        return [(self.L.pnb_kernel_class_name(cls[k]).decode(), t0[k], t1[k]) for k in range(n)]

    def set_overlap(self, on: bool):
        self._ck(self.L.pnb_set_overlap(self.h, 1 if on else 0), "pnb_set_overlap")

    def overlap_info(self) -> dict:
        a, b, c = C.c_int(), C.c_int(), C.c_int()
        self._ck(self.L.pnb_overlap_info(self.h, C.byref(a), C.byref(b), C.byref(c)), "pnb_overlap_info")
        return {"net_sms": a.value, "dsp_sms": b.value, "chunk_hops": c.value}

    @property
    def launches(self) -> int:
        return int(self.L.pnb_launch_count(self.h))

    def launches_per_call(self, n_frames: int) -> int:
        return int(self.L.pnb_launches_per_call(self.h, n_frames))
