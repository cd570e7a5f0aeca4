"""ctypes front-end for the CPU oracle (oracle/nam_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path
from typing import Optional, Sequence

import numpy as np

from . import nam_config
from .nam_config import FlatModel

HERE = Path(__file__).resolve().parent
_LIBS: dict[str, C.CDLL] = {}

f32p = C.POINTER(C.c_float)
f64p = C.POINTER(C.c_double)
i32p = C.POINTER(C.c_int32)


def build(force: bool = False) -> None:
    """Compile libnam_oracle.so / libnam_oracle_fast.so in-tree (gcc, a few seconds)."""
    strict, fast = HERE / "libnam_oracle.so", HERE / "libnam_oracle_fast.so"
    src = HERE / "nam_oracle.c"
    if not force and strict.exists() and fast.exists() and min(strict.stat().st_mtime, fast.stat().st_mtime) >= max(
        src.stat().st_mtime, (HERE / "nam_oracle.h").stat().st_mtime
    ):
        return
    subprocess.run(["make", "-C", str(HERE), "-B", "all"], check=True, capture_output=True)


def build_native_fast(out_dir: str | Path) -> Optional[Path]:
    """-march=native build for CPU-baseline timing on the box that runs the bench."""
    out = Path(out_dir) / "libnam_oracle_native.so"
    try:
        subprocess.run(
            ["make", "-C", str(HERE), "-B", f"FAST_OUT={out}", "FAST_MARCH=native", str(out)],
            check=True,
            capture_output=True,
        )
        return out
    except Exception:
        return None


def _declare(lib: C.CDLL) -> C.CDLL:
    lib.nam_oracle_create.restype = C.c_void_p
    lib.nam_oracle_create.argtypes = [i32p, C.c_int, f32p, C.c_int, f32p, C.c_int, C.c_double, C.c_int, C.c_void_p]
    lib.nam_oracle_destroy.argtypes = [C.c_void_p]
    lib.nam_oracle_last_error.restype = C.c_char_p
    for name in ("in_channels", "out_channels", "prewarm_samples", "weights_consumed"):
        fn = getattr(lib, f"nam_oracle_{name}")
        fn.restype = C.c_int
        fn.argtypes = [C.c_void_p]
    lib.nam_oracle_reset.argtypes = [C.c_void_p, C.c_double, C.c_int, C.c_int]
    lib.nam_oracle_process_f32.argtypes = [C.c_void_p, C.POINTER(f32p), C.POINTER(f32p), C.c_int]
    lib.nam_oracle_process_f64.argtypes = [C.c_void_p, C.POINTER(f64p), C.POINTER(f64p), C.c_int]
    lib.nam_oracle_run_mono_f32.argtypes = [C.c_void_p, f32p, f32p, C.c_long, C.c_int]
    lib.nam_oracle_run_batch_mono_f32.restype = C.c_int
    lib.nam_oracle_run_batch_mono_f32.argtypes = [C.c_void_p, f32p, f32p, C.c_int, C.c_long, C.c_int, C.c_int]
    lib.nam_oracle_conv1d.restype = C.c_int
    lib.nam_oracle_conv1d.argtypes = [C.c_int] * 6 + [f32p, C.c_int, f32p, f32p, C.c_int, C.c_int]
    lib.nam_oracle_conv1x1.restype = C.c_int
    lib.nam_oracle_conv1x1.argtypes = [C.c_int] * 4 + [f32p, C.c_int, f32p, f32p, C.c_int]
    lib.nam_oracle_film.restype = C.c_int
    lib.nam_oracle_film.argtypes = [C.c_int] * 4 + [f32p, C.c_int, f32p, f32p, f32p, C.c_int]
    lib.nam_oracle_activation.restype = C.c_int
    lib.nam_oracle_activation.argtypes = [C.c_int, f32p, C.c_int, C.c_int, f32p, C.c_int, C.c_int]
    lib.nam_oracle_gating.restype = C.c_int
    lib.nam_oracle_gating.argtypes = [C.c_int, C.c_int, f32p, C.c_int, C.c_int, f32p, C.c_int, C.c_int, f32p, f32p, C.c_int]
    lib.nam_oracle_batch_create.restype = C.c_void_p
    lib.nam_oracle_batch_create.argtypes = [C.c_void_p, C.c_int]
    lib.nam_oracle_batch_process.restype = C.c_int
    lib.nam_oracle_batch_process.argtypes = [C.c_void_p, f32p, f32p, C.c_long, C.c_long, C.c_long, C.c_int, C.c_int]
    lib.nam_oracle_batch_destroy.argtypes = [C.c_void_p]
    lib.nam_oracle_layer.restype = C.c_int
    lib.nam_oracle_layer.argtypes = [i32p, C.c_int, f32p, C.c_int, f32p, C.c_int, f32p, f32p, f32p, f32p, C.c_int, C.c_int]
    return lib


def load_lib(kind: str = "strict", path: str | Path | None = None) -> C.CDLL:
    key = str(path) if path else kind
    if key not in _LIBS:
        if path is None:
            build()
            path = HERE / ("libnam_oracle.so" if kind == "strict" else "libnam_oracle_fast.so")
        _LIBS[key] = _declare(C.CDLL(str(path)))
    return _LIBS[key]


def _fp(a: np.ndarray):
    return a.ctypes.data_as(f32p)


class OracleError(RuntimeError):
    pass


class OracleModel:
    """One nam::DSP-like object on the CPU oracle."""

    def __init__(self, flat: FlatModel, fast_tanh: bool = False, lib: C.CDLL | None = None):
        self.lib = lib or load_lib()
        self.flat = flat
        self.fast_tanh = bool(fast_tanh)
        self._h = self._create(flat)

    def _create(self, flat: FlatModel) -> int:
        child = None
        if flat.condition_dsp is not None:
            child = self._create(flat.condition_dsp)
        cfg = np.ascontiguousarray(flat.cfg, dtype=np.int32)
        fpar = np.ascontiguousarray(flat.fparams, dtype=np.float32)
        w = np.ascontiguousarray(flat.weights, dtype=np.float32)
        h = self.lib.nam_oracle_create(
            cfg.ctypes.data_as(i32p), len(cfg), _fp(fpar), len(fpar), _fp(w), len(w), float(flat.sample_rate),
            int(self.fast_tanh), child,
        )
        if not h:
            msg = self.lib.nam_oracle_last_error().decode()
            if child:
                self.lib.nam_oracle_destroy(child)
            raise OracleError(msg)
        return h

    @classmethod
    def from_file(cls, path, fast_tanh: bool = False, lib=None) -> "OracleModel":
        return cls(nam_config.flatten_file(path), fast_tanh=fast_tanh, lib=lib)

    @classmethod
    def from_dict(cls, nam: dict, fast_tanh: bool = False, lib=None) -> "OracleModel":
        return cls(nam_config.flatten(nam), fast_tanh=fast_tanh, lib=lib)

    def close(self):
        if getattr(self, "_h", None):
            self.lib.nam_oracle_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def in_channels(self) -> int:
        return self.lib.nam_oracle_in_channels(self._h)

    @property
    def out_channels(self) -> int:
        return self.lib.nam_oracle_out_channels(self._h)

    @property
    def prewarm_samples(self) -> int:
        return self.lib.nam_oracle_prewarm_samples(self._h)

    @property
    def weights_consumed(self) -> int:
        return self.lib.nam_oracle_weights_consumed(self._h)

    def reset(self, sample_rate: float, max_buffer_size: int, prewarm: bool = True) -> None:
        self.max_buffer_size = int(max_buffer_size)
        self.lib.nam_oracle_reset(self._h, float(sample_rate), int(max_buffer_size), int(prewarm))

    def process(self, x: np.ndarray) -> np.ndarray:
        """x: (in_channels, n) or (n,) float32/float64 -> (out_channels, n) same dtype. n <= max_buffer_size."""
        x = np.asarray(x)
        squeeze = x.ndim == 1
        if squeeze:
            x = x[None, :]
        n = x.shape[1]
        assert x.shape[0] == self.in_channels and n <= self.max_buffer_size
        if x.dtype == np.float64:
            xin = np.ascontiguousarray(x)
            out = np.zeros((self.out_channels, n), dtype=np.float64)
            ip = (f64p * self.in_channels)(*[xin[c].ctypes.data_as(f64p) for c in range(self.in_channels)])
            op = (f64p * self.out_channels)(*[out[c].ctypes.data_as(f64p) for c in range(self.out_channels)])
            self.lib.nam_oracle_process_f64(self._h, ip, op, n)
        else:
            xin = np.ascontiguousarray(x, dtype=np.float32)
            out = np.zeros((self.out_channels, n), dtype=np.float32)
            ip = (f32p * self.in_channels)(*[xin[c].ctypes.data_as(f32p) for c in range(self.in_channels)])
            op = (f32p * self.out_channels)(*[out[c].ctypes.data_as(f32p) for c in range(self.out_channels)])
            self.lib.nam_oracle_process_f32(self._h, ip, op, n)
        return out[0] if squeeze and self.out_channels == 1 else out

    def run(self, x: np.ndarray, block: int) -> np.ndarray:
        """Mono convenience: whole signal in `block`-frame calls (block <= max_buffer_size)."""
        x = np.ascontiguousarray(x, dtype=np.float32)
        out = np.zeros_like(x)
        assert block <= self.max_buffer_size
        self.lib.nam_oracle_run_mono_f32(self._h, _fp(x), _fp(out), len(x), int(block))
        return out

    def run_batch(self, x: np.ndarray, block: int, threads: int | None = None) -> np.ndarray:
        """x: (batch, n) mono streams, each starting from this model's CURRENT state (not advanced)."""
        x = np.ascontiguousarray(x, dtype=np.float32)
        out = np.zeros_like(x)
        threads = threads or os.cpu_count() or 1
        rc = self.lib.nam_oracle_run_batch_mono_f32(self._h, _fp(x), _fp(out), x.shape[0], x.shape[1], int(block), threads)
        if rc != 0:
            raise OracleError(self.lib.nam_oracle_last_error().decode())
        return out


class OracleBatch:
    """`batch` independent instances (clones of a prepared OracleModel, state included) that keep their state
    across process() calls -- one nam::DSP per stream, spread over host threads.  CPU-baseline timing."""

    def __init__(self, proto: OracleModel, batch: int):
        self.lib = proto.lib
        self.batch = int(batch)
        self.max_buffer_size = proto.max_buffer_size
        self._h = self.lib.nam_oracle_batch_create(proto._h, self.batch)
        if not self._h:
            raise OracleError(self.lib.nam_oracle_last_error().decode())

    def process(self, x: np.ndarray, out: np.ndarray | None = None, block: int = 64, threads: int | None = None) -> np.ndarray:
        assert x.dtype == np.float32 and x.ndim == 2 and x.shape[0] == self.batch and x.flags["C_CONTIGUOUS"]
        if out is None:
            out = np.empty_like(x)
        threads = threads or os.cpu_count() or 1
        rc = self.lib.nam_oracle_batch_process(self._h, _fp(x), _fp(out), x.shape[1], x.shape[1], out.shape[1], int(block), threads)
        if rc != 0:
            raise OracleError(self.lib.nam_oracle_last_error().decode())
        return out

    def close(self):
        if getattr(self, "_h", None):
            self.lib.nam_oracle_batch_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ---- module-level helpers (column-major (channels x frames) numpy arrays in "F" order) --------------------------
def _cm(a: np.ndarray) -> np.ndarray:
    """(channels, frames) array -> flat column-major float32 buffer."""
    return np.array(np.asarray(a, dtype=np.float32).T, dtype=np.float32, order="C", copy=True).ravel()


def _from_cm(buf: np.ndarray, channels: int, frames: int) -> np.ndarray:
    return buf.reshape(frames, channels).T.copy()


def _check(rc: int, lib) -> None:
    if rc != 0:
        raise OracleError(lib.nam_oracle_last_error().decode())


def conv1d(x, weights, in_ch, out_ch, kernel, dilation=1, bias=True, groups=1, n_calls=1, lib=None) -> np.ndarray:
    lib = lib or load_lib()
    x = np.asarray(x, dtype=np.float32).reshape(in_ch, -1)
    total = x.shape[1]
    assert total % n_calls == 0
    w = np.ascontiguousarray(weights, dtype=np.float32)
    xin = _cm(x)
    out = np.zeros(out_ch * total, dtype=np.float32)
    _check(lib.nam_oracle_conv1d(in_ch, out_ch, kernel, dilation, int(bias), groups, _fp(w), len(w), _fp(xin), _fp(out),
                                 total // n_calls, n_calls), lib)
    return _from_cm(out, out_ch, total)


def conv1x1(x, weights, in_ch, out_ch, bias=True, groups=1, lib=None) -> np.ndarray:
    lib = lib or load_lib()
    x = np.asarray(x, dtype=np.float32).reshape(in_ch, -1)
    n = x.shape[1]
    w = np.ascontiguousarray(weights, dtype=np.float32)
    xin = _cm(x)
    out = np.zeros(out_ch * n, dtype=np.float32)
    _check(lib.nam_oracle_conv1x1(in_ch, out_ch, int(bias), groups, _fp(w), len(w), _fp(xin), _fp(out), n), lib)
    return _from_cm(out, out_ch, n)


def film(x, cond, weights, cond_dim, input_dim, shift=True, groups=1, lib=None) -> np.ndarray:
    lib = lib or load_lib()
    x = np.asarray(x, dtype=np.float32).reshape(input_dim, -1)
    cond = np.asarray(cond, dtype=np.float32).reshape(cond_dim, -1)
    n = x.shape[1]
    w = np.ascontiguousarray(weights, dtype=np.float32)
    xin, cin = _cm(x), _cm(cond)
    out = np.zeros(input_dim * n, dtype=np.float32)
    _check(lib.nam_oracle_film(cond_dim, input_dim, int(shift), groups, _fp(w), len(w), _fp(xin), _fp(cin), _fp(out), n), lib)
    return _from_cm(out, input_dim, n)


def _act_args(act) -> tuple[int, np.ndarray]:
    cfg: list = []
    fp: list = []
    nam_config._act(act, cfg, fp)
    return cfg[0], np.asarray(fp, dtype=np.float32)


def activation(x, act, fast_tanh=False, lib=None) -> np.ndarray:
    """act: name or {"type":..., params}.  x: (channels, frames) or (n,) (treated as 1 channel)."""
    lib = lib or load_lib()
    x = np.asarray(x, dtype=np.float32)
    shape = x.shape
    x2 = x.reshape(1, -1) if x.ndim == 1 else x
    code, params = _act_args(act)
    buf = _cm(x2)
    _check(lib.nam_oracle_activation(code, _fp(params), len(params), int(fast_tanh), _fp(buf), x2.shape[0], x2.shape[1]), lib)
    return _from_cm(buf, x2.shape[0], x2.shape[1]).reshape(shape)


def gating(x, mode: str, act, sec_act, channels: int, lib=None) -> np.ndarray:
    lib = lib or load_lib()
    x = np.asarray(x, dtype=np.float32).reshape(2 * channels, -1)
    n = x.shape[1]
    c1, p1 = _act_args(act)
    c2, p2 = _act_args(sec_act)
    xin = _cm(x)
    out = np.zeros(channels * n, dtype=np.float32)
    _check(lib.nam_oracle_gating({"gated": 1, "blended": 2}[mode], c1, _fp(p1), len(p1), c2, _fp(p2), len(p2), channels,
                                 _fp(xin), _fp(out), n), lib)
    return _from_cm(out, channels, n)


def layer(x, cond, weights, *, condition_size=1, channels=1, bottleneck=None, kernel_size=1, dilation=1,
          activation="ReLU", gating_mode="none", secondary_activation=None, groups_input=1, groups_input_mixin=1,
          layer1x1=(True, 1), head1x1=(False, None, 1), films=None, fast_tanh=False, lib=None):
    """One Layer (reference make_layer() in tools/test/test_wavenet/test_layer.cpp) from zero history.
    Returns (output_next_layer [channels, n], output_head [rows, n])."""
    lib = lib or load_lib()
    bottleneck = channels if bottleneck is None else bottleneck
    h1_active, h1_out, h1_groups = head1x1
    h1_out = channels if h1_out is None else h1_out
    films = films or {}
    cfg: list = [1, condition_size, channels, bottleneck, 1, 1, 1, 0, groups_input, groups_input_mixin,
                 int(layer1x1[0]), layer1x1[1], int(h1_active), h1_out, h1_groups]
    for key in nam_config.FILM_KEYS:
        a, s, g = films.get(key, (0, 0, 1))
        cfg += [int(a), int(s), int(g)]
    cfg += [kernel_size, dilation, nam_config.GATING[gating_mode]]
    fp: list = []
    nam_config._act(activation, cfg, fp)
    nam_config._act(secondary_activation, cfg, fp)
    x = np.asarray(x, dtype=np.float32).reshape(channels, -1)
    n = x.shape[1]
    cond = np.asarray(cond, dtype=np.float32).reshape(condition_size, -1)
    cfg_a = np.asarray(cfg, dtype=np.int32)
    fp_a = np.asarray(fp, dtype=np.float32)
    w = np.ascontiguousarray(weights, dtype=np.float32)
    head_rows = h1_out if h1_active else bottleneck
    xin, cin = _cm(x), _cm(cond)
    o_next = np.zeros(channels * n, dtype=np.float32)
    o_head = np.zeros(head_rows * n, dtype=np.float32)
    _check(lib.nam_oracle_layer(cfg_a.ctypes.data_as(i32p), len(cfg_a), _fp(fp_a), len(fp_a), _fp(w), len(w), _fp(xin),
                                _fp(cin), _fp(o_next), _fp(o_head), n, int(fast_tanh)), lib)
    return _from_cm(o_next, channels, n), _from_cm(o_head, head_rows, n)
