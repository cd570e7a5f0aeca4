"""Flatten a .nam model description into the three arrays oracle/nam_oracle.c consumes.

TEST INFRASTRUCTURE ONLY (see oracle/nam_oracle.h).  This is an independent restatement of
the reference's JSON config parsers; the product has its own C++ parser
(neuralampmodelercore_b200/csrc/nam_model_spec.cpp) and the two are cross-checked in tests.

Reference behaviour restated here (paths relative to the reference repository):
  * required keys / file validation ......... NAM/nam_file.cpp:9-40
  * version gate ............................. NAM/get_dsp.cpp:18-39,113-128
  * sample_rate (-1 when absent) ............. NAM/get_dsp.cpp:275-281
  * WaveNet config ........................... NAM/wavenet/model.cpp:913-1276
  * activation config ........................ NAM/activations.cpp:55-130
  * LSTM config .............................. NAM/lstm.cpp:171-181
  * Linear config ............................ NAM/linear.cpp:306-316
  * ConvNet config ........................... NAM/convnet.cpp:321-335

cfg (int32) schema, consumed sequentially by nam_oracle.c:
  WaveNet : 1, in_channels, n_arrays, with_head,
            per array: input_size, condition_size, channels, bottleneck, head_size, head_kernel,
                       head_dilation, head_bias, groups_input, groups_input_mixin,
                       layer1x1_active, layer1x1_groups, head1x1_active, head1x1_out, head1x1_groups,
                       8 x (film_active, film_shift, film_groups)   [order FILM_KEYS],
                       n_layers, per layer: kernel, dilation, gating_mode, ACT, SECONDARY_ACT
            if with_head: head_channels, head_out_channels, n_kernel_sizes, kernel_sizes..., ACT
  LSTM    : 2, in_channels, out_channels, num_layers, input_size, hidden_size
  Linear  : 3, in_channels, out_channels, receptive_field, bias
  ConvNet : 4, in_channels, out_channels, channels, n_dilations, dilations..., batchnorm, groups, ACT
  ACT     : type_code, n_params   (the params themselves go to fparams, in order)
fparams (float32): [head_scale (WaveNet only)] then activation parameters in cfg order.
"""
from __future__ import annotations

import json
import re
from dataclasses import dataclass, field
from pathlib import Path
from typing import Any, Optional

import numpy as np

ARCH_WAVENET, ARCH_LSTM, ARCH_LINEAR, ARCH_CONVNET = 1, 2, 3, 4

ACT_CODES = {
    "Tanh": 0,
    "Hardtanh": 1,
    "Fasttanh": 2,
    "ReLU": 3,
    "LeakyReLU": 4,
    "PReLU": 5,
    "Sigmoid": 6,
    "SiLU": 7,
    "Hardswish": 8,
    "LeakyHardtanh": 9,
    "LeakyHardTanh": 9,  # both casings accepted, activations.cpp:77-78
    "Softsign": 10,
}
ACT_IDENTITY = 100

GATING = {"none": 0, "gated": 1, "blended": 2}

FILM_KEYS = [
    "conv_pre_film",
    "conv_post_film",
    "input_mixin_pre_film",
    "input_mixin_post_film",
    "activation_pre_film",
    "activation_post_film",
    "layer1x1_post_film",
    "head1x1_post_film",
]

LATEST_FULLY_SUPPORTED = (0, 7, 0)
EARLIEST_SUPPORTED = (0, 5, 0)


class NamConfigError(RuntimeError):
    pass


def version_support(version: str) -> str:
    """'yes' | 'partial' | 'no'  (NAM/get_dsp.cpp:18-39)."""
    if not re.match(r"^\d+\.\d+\.\d+$", version):
        return "no"
    v = tuple(int(p) for p in version.split("."))
    if v < EARLIEST_SUPPORTED:
        return "no"
    if v[0] > LATEST_FULLY_SUPPORTED[0] or v[1] > LATEST_FULLY_SUPPORTED[1]:
        return "no"
    if v > LATEST_FULLY_SUPPORTED:
        return "partial"
    return "yes"


@dataclass
class FlatModel:
    architecture: str
    cfg: np.ndarray
    fparams: np.ndarray
    weights: np.ndarray
    sample_rate: float
    metadata: dict = field(default_factory=dict)
    condition_dsp: Optional["FlatModel"] = None
    raw_config: dict = field(default_factory=dict)


def _act(j: Any, cfg: list, fp: list) -> None:
    """ActivationConfig::from_json (activations.cpp:55-130) -> (code, n_params) + params."""
    if j is None:
        cfg += [ACT_IDENTITY, 0]
        return
    if isinstance(j, str):
        if j not in ACT_CODES:
            raise NamConfigError(f"Unknown activation type: {j}")
        cfg += [ACT_CODES[j], 0]
        return
    if isinstance(j, dict):
        t = j["type"]
        if t not in ACT_CODES:
            raise NamConfigError(f"Unknown activation type: {t}")
        code = ACT_CODES[t]
        params: list[float] = []
        if t == "PReLU":
            if "negative_slope" in j:
                params = [float(j["negative_slope"])]
            elif "negative_slopes" in j:
                params = [float(x) for x in j["negative_slopes"]]
        elif t == "LeakyReLU":
            params = [float(j.get("negative_slope", 0.01))]
        elif code == 9:
            params = [
                float(j.get("min_val", -1.0)),
                float(j.get("max_val", 1.0)),
                float(j.get("min_slope", 0.01)),
                float(j.get("max_slope", 0.01)),
            ]
        cfg += [code, len(params)]
        fp += params
        return
    raise NamConfigError("Invalid activation config: expected string or object")


def _film(layer_cfg: dict, key: str) -> list[int]:
    # parse_film_params, model.cpp:1190-1201
    if key not in layer_cfg or layer_cfg[key] is False:
        return [0, 0, 1]
    f = layer_cfg[key]
    return [int(bool(f.get("active", True))), int(bool(f.get("shift", True))), int(f.get("groups", 1))]


def _wavenet(config: dict, cfg: list, fp: list) -> None:
    layers = config["layers"]
    with_head = ("head" in config) and config["head"] is not None
    cfg += [ARCH_WAVENET, int(config.get("in_channels", 1)), len(layers), int(with_head)]
    fp.append(float(config["head_scale"]))
    for i, lc in enumerate(layers):
        channels = int(lc["channels"])
        bottleneck = int(lc.get("bottleneck", channels))
        l1_active, l1_groups = 1, 1
        if "layer1x1" in lc:
            l1_active, l1_groups = int(bool(lc["layer1x1"]["active"])), int(lc["layer1x1"]["groups"])
        head_dilation, head_kernel, head_bias = 1, 1, 0
        if lc.get("head") is not None:
            h = lc["head"]
            if not isinstance(h, dict):
                raise NamConfigError(f"Layer array {i}: 'head' must be a JSON object")
            head_size = int(h["out_channels"])
            head_dilation = int(h.get("head_dilation", 1))
            head_kernel = int(h["kernel_size"])
            head_bias = int(bool(h["bias"]))
        elif "head_size" in lc:
            head_size = int(lc["head_size"])
            head_bias = int(bool(lc["head_bias"]))
        else:
            raise NamConfigError(f"Layer array {i}: expected 'head' object or legacy 'head_size' and 'head_bias'")
        if head_kernel < 1:
            raise NamConfigError(f"Layer array {i}: head.kernel_size must be >= 1")
        dilations = [int(d) for d in lc["dilations"]]
        n_layers = len(dilations)
        has_k, has_ks = "kernel_size" in lc, "kernel_sizes" in lc
        if has_k and has_ks:
            raise NamConfigError(f"Layer array {i}: only one of kernel_size or kernel_sizes may be provided")
        if has_ks:
            kernel_sizes = [int(k) for k in lc["kernel_sizes"]]
            if len(kernel_sizes) != n_layers:
                raise NamConfigError(f"Layer array {i}: kernel_sizes array size must match dilations size")
        elif has_k:
            kernel_sizes = [int(lc["kernel_size"])] * n_layers
        else:
            raise NamConfigError(f"Layer array {i}: either kernel_size or kernel_sizes must be provided")
        act_j = lc["activation"]
        acts = list(act_j) if isinstance(act_j, list) else [act_j] * n_layers
        if len(acts) != n_layers:
            raise NamConfigError(f"Layer array {i}: activation array size must match dilations size")
        # gating modes + secondary activations, model.cpp:1062-1176
        if "gating_mode" in lc:
            gm = lc["gating_mode"]
            modes = list(gm) if isinstance(gm, list) else [gm] * n_layers
            if len(modes) != n_layers:
                raise NamConfigError(f"Layer array {i}: gating_mode array size must match dilations size")
            for m in modes:
                if m not in GATING:
                    raise NamConfigError(f"Invalid gating_mode: {m}")
            sec_j = lc.get("secondary_activation", None)
            if isinstance(sec_j, list) and len(sec_j) != n_layers:
                raise NamConfigError(f"Layer array {i}: secondary_activation array size must match dilations size")
            secs = []
            for li, m in enumerate(modes):
                if m == "none":
                    secs.append(None)
                elif "secondary_activation" in lc:
                    secs.append(sec_j[li] if isinstance(sec_j, list) else sec_j)
                else:
                    secs.append("Sigmoid")
        elif "gated" in lc:
            gated = bool(lc["gated"])
            modes = ["gated" if gated else "none"] * n_layers
            secs = ["Sigmoid" if gated else None] * n_layers
        else:
            modes = ["none"] * n_layers
            secs = [None] * n_layers
        h1_active, h1_out, h1_groups = 0, channels, 1
        if "head1x1" in lc:
            h1 = lc["head1x1"]
            h1_active, h1_out, h1_groups = int(bool(h1["active"])), int(h1["out_channels"]), int(h1["groups"])
        films = [_film(lc, k) for k in FILM_KEYS]
        if films[6][0] and not l1_active:
            raise NamConfigError(f"Layer array {i}: layer1x1_post_film cannot be active when layer1x1.active is false")
        cfg += [
            int(lc["input_size"]),
            int(lc["condition_size"]),
            channels,
            bottleneck,
            head_size,
            head_kernel,
            head_dilation,
            head_bias,
            int(lc.get("groups_input", 1)),
            int(lc.get("groups_input_mixin", 1)),
            l1_active,
            l1_groups,
            h1_active,
            h1_out,
            h1_groups,
        ]
        for f in films:
            cfg += f
        cfg.append(n_layers)
        for li in range(n_layers):
            cfg += [kernel_sizes[li], dilations[li], GATING[modes[li]]]
            _act(acts[li], cfg, fp)
            _act(secs[li], cfg, fp)
    if with_head:
        hj = config["head"]
        implied_in = cfg_last_head_size(layers)
        if hj.get("in_channels") is not None and int(hj["in_channels"]) != implied_in:
            raise NamConfigError("WaveNet config: head.in_channels must equal last layer's head_size")
        ks = [int(k) for k in hj["kernel_sizes"]]
        if not ks:
            raise NamConfigError("WaveNet config: head.kernel_sizes must be non-empty")
        cfg += [int(hj["channels"]), int(hj["out_channels"]), len(ks)] + ks
        _act(hj["activation"], cfg, fp)


def cfg_last_head_size(layers: list) -> int:
    lc = layers[-1]
    if lc.get("head") is not None:
        return int(lc["head"]["out_channels"])
    return int(lc["head_size"])


def flatten(nam: dict) -> FlatModel:
    """dict with version/architecture/config/weights[/sample_rate/metadata] -> FlatModel."""
    for key in ("version", "architecture", "config", "weights"):
        if key not in nam:
            raise NamConfigError(f'Invalid .nam: missing required key "{key}".')
    if version_support(nam["version"]) == "no":
        raise NamConfigError(f"Model config is an unsupported version {nam['version']}.")
    arch = nam["architecture"]
    config = nam["config"]
    sr = float(nam["sample_rate"]) if "sample_rate" in nam else -1.0
    cfg: list[int] = []
    fp: list[float] = []
    cond = None
    if arch == "WaveNet":
        if config.get("condition_dsp") is not None:
            cond = flatten(config["condition_dsp"])
        _wavenet(config, cfg, fp)
    elif arch == "LSTM":
        cfg += [
            ARCH_LSTM,
            int(config.get("in_channels", 1)),
            int(config.get("out_channels", 1)),
            int(config["num_layers"]),
            int(config["input_size"]),
            int(config["hidden_size"]),
        ]
    elif arch == "Linear":
        cfg += [
            ARCH_LINEAR,
            int(config.get("in_channels", 1)),
            int(config.get("out_channels", 1)),
            int(config["receptive_field"]),
            int(bool(config["bias"])),
        ]
    elif arch == "ConvNet":
        dil = [int(d) for d in config["dilations"]]
        cfg += [ARCH_CONVNET, int(config.get("in_channels", 1)), int(config.get("out_channels", 1)),
                int(config["channels"]), len(dil)] + dil + [int(bool(config["batchnorm"])), int(config.get("groups", 1))]
        _act(config["activation"], cfg, fp)
    else:
        raise NamConfigError(f"No config parser registered for architecture: {arch}")
    return FlatModel(
        architecture=arch,
        cfg=np.asarray(cfg, dtype=np.int32),
        fparams=np.asarray(fp, dtype=np.float32),
        weights=np.asarray(nam["weights"], dtype=np.float32),
        sample_rate=sr,
        metadata=nam.get("metadata") or {},
        condition_dsp=cond,
        raw_config=config,
    )


def load_nam(path: str | Path) -> dict:
    path = Path(path)
    if not path.exists():
        raise NamConfigError(f"Could not validate .nam file [{path}]: file does not exist.")
    try:
        j = json.loads(path.read_text())
    except json.JSONDecodeError as e:
        raise NamConfigError(f"Could not parse .nam file [{path}]: {e}") from e
    if not isinstance(j, dict):
        raise NamConfigError(f"Invalid .nam file [{path}]: root JSON value must be an object.")
    return j


def flatten_file(path: str | Path) -> FlatModel:
    return flatten(load_nam(path))


def expected_weight_count(nam: dict) -> int:
    """Weight count implied by the config, following the stream order of
    NAM/wavenet/model.cpp:152-181,563-569,661-670 / NAM/lstm.cpp:9-29,70-101 / NAM/linear.cpp:61-81."""
    arch, c = nam["architecture"], nam["config"]
    if arch == "LSTM":
        n, I, H = int(c["num_layers"]), int(c["input_size"]), int(c["hidden_size"])
        out = int(c.get("out_channels", 1))
        total = 0
        for l in range(n):
            i = I if l == 0 else H
            total += 4 * H * (i + H) + 4 * H + 2 * H
        return total + out * H + out
    if arch == "Linear":
        return int(c["receptive_field"]) + int(bool(c["bias"]))
    if arch == "ConvNet":  # convnet.cpp:48-60 (conv, kernel 2, bias iff no batchnorm), :14-37 (4 x dim + eps), :132-153 head
        ch, g, bn = int(c["channels"]), int(c.get("groups", 1)), bool(c["batchnorm"])
        cin, total = int(c.get("in_channels", 1)), 0
        for _d in c["dilations"]:
            total += (cin * ch * 2) // g + (4 * ch + 1 if bn else ch)
            cin = ch
        out = int(c.get("out_channels", 1))
        return total + out * ch + out
    if arch != "WaveNet":
        raise NamConfigError(arch)
    fm = flatten({**nam, "weights": []})
    it = iter(fm.cfg.tolist())
    nxt = lambda: next(it)  # noqa: E731
    assert nxt() == ARCH_WAVENET
    _in_ch, n_arrays, with_head = nxt(), nxt(), nxt()
    total = 0
    last_head = 0

    def skip_act():
        nxt()
        nxt()

    for _ in range(n_arrays):
        (inp, cs, ch, bn, hs, hk, _hd, hb, gi, gm, l1a, l1g, h1a, h1o, h1g) = [nxt() for _ in range(15)]
        films = [(nxt(), nxt(), nxt()) for _ in range(8)]
        nl = nxt()
        total += inp * ch  # rechannel, no bias
        film_dims = None
        for _l in range(nl):
            k, _d, g = nxt(), nxt(), nxt()
            skip_act()
            skip_act()
            z = 2 * bn if g != 0 else bn
            total += (ch * z * k) // gi + z  # conv + bias
            total += (cs * z) // gm  # input mixin
            if l1a:
                total += (bn * ch) // l1g + ch
            if h1a:
                total += (bn * h1o) // h1g + h1o
            film_dims = [ch, z, cs, z, z, bn, ch, h1o]
            for fi, (fa, fs, fg) in enumerate(films):
                if not fa or (fi == 6 and not l1a) or (fi == 7 and not h1a):
                    continue
                od = (2 if fs else 1) * film_dims[fi]
                total += (cs * od) // fg + od
        head_in = h1o if h1a else bn
        total += head_in * hs * hk + (hs if hb else 0)
        last_head = hs
    if with_head:
        hc, ho, nk = nxt(), nxt(), nxt()
        cin = last_head
        for i in range(nk):
            k = nxt()
            cout = ho if i + 1 == nk else hc
            total += cin * cout * k + cout
            cin = cout
    return total + 1  # head_scale
