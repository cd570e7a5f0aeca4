#!/usr/bin/env python
"""Generate the committed fixtures under tests/golden/ (run in the build container, where
/root/reference exists; the GPU box only reads the results).

  models/<name>.npz   the reference's example models as TEST DATA (model config JSON + float32 weight
                      arrays; these are trained parameters, not code).  Re-packed so a test can hand
                      them to the product (JSON text) and to the oracle (flat arrays) without touching
                      /root/reference at run time.
  input_wav.npz       example_audio/input.wav decoded to float32 (24-bit PCM / 2^23, mono, 48 kHz).
  oracle_outputs.npz  outputs of the CPU oracle (oracle/nam_oracle.c, strict build) on that input for
                      each model and tanh regime -- a regression pin of the oracle itself, decimated to
                      keep the repository small: the full first 512 samples, the full 4096 samples
                      around the silence->sine transition, and every 37th sample of the whole signal.
  reference_pins.json known-answer values transcribed from the reference's own unit tests
                      (tools/test/*.cpp, cited per entry) -- the module-level pins of SURVEY.md 8(c).

Usage: python tests/golden/make_golden.py [--reference /root/reference]
"""
from __future__ import annotations

import argparse
import json
import struct
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
HERE = Path(__file__).resolve().parent

MODELS = {
    # fixture name -> file under example_models/
    "wavenet": "wavenet.nam",
    "wavenet_a1_standard": "wavenet_a1_standard.nam",
    "lstm": "lstm.nam",
    "wavenet_condition_dsp": "wavenet_condition_dsp.nam",
    "wavenet_a2_max": "wavenet_a2_max.nam",
    "slimmable_wavenet": "slimmable_wavenet.nam",
}
# sub-models of the SlimmableContainer example (A2.nam): the plain WaveNets inside it
A2_SUBMODELS = {"a2_lite": 0, "a2_full": 1}


def strip_weights(obj, store: list):
    """Replace every "weights": [...] list by {"__weights__": index}; arrays go to `store`."""
    if isinstance(obj, dict):
        out = {}
        for k, v in obj.items():
            if k == "weights" and isinstance(v, list):
                store.append(np.asarray(v, dtype=np.float32))
                out[k] = {"__weights__": len(store) - 1}
            else:
                out[k] = strip_weights(v, store)
        return out
    if isinstance(obj, list):
        return [strip_weights(v, store) for v in obj]
    return obj


def save_model(name: str, nam: dict) -> None:
    store: list = []
    header = strip_weights(nam, store)
    # training metadata is irrelevant to the arithmetic; keep only what the loader reads
    if isinstance(header.get("metadata"), dict):
        header["metadata"] = {k: v for k, v in header["metadata"].items()
                              if k in ("loudness", "input_level_dbu", "output_level_dbu", "gain", "name")}
    arrays = {f"w{i}": a for i, a in enumerate(store)}
    out = HERE / "models" / f"{name}.npz"
    out.parent.mkdir(exist_ok=True)
    np.savez_compressed(out, header=np.frombuffer(json.dumps(header).encode(), dtype=np.uint8), **arrays)
    print(f"wrote {out.relative_to(ROOT)} ({out.stat().st_size} bytes, {sum(len(a) for a in store)} weights)")


def read_wav_pcm(path: Path) -> tuple[np.ndarray, int]:
    """Minimal RIFF/WAVE reader: PCM 16/24/32-bit or IEEE float32, returns mono float32 + sample rate."""
    data = path.read_bytes()
    assert data[:4] == b"RIFF" and data[8:12] == b"WAVE", "not a RIFF/WAVE file"
    pos = 12
    fmt = None
    pcm = None
    while pos + 8 <= len(data):
        cid, size = data[pos:pos + 4], struct.unpack("<I", data[pos + 4:pos + 8])[0]
        body = data[pos + 8:pos + 8 + size]
        if cid == b"fmt ":
            fmt = struct.unpack("<HHIIHH", body[:16])
        elif cid == b"data":
            pcm = body
        pos += 8 + size + (size & 1)
    assert fmt is not None and pcm is not None
    tag, channels, rate, _, _, bits = fmt
    if tag == 3 and bits == 32:
        x = np.frombuffer(pcm, dtype="<f4").astype(np.float32)
    elif bits == 16:
        x = np.frombuffer(pcm, dtype="<i2").astype(np.float32) / 32768.0
    elif bits == 24:
        b = np.frombuffer(pcm, dtype=np.uint8).reshape(-1, 3).astype(np.int32)
        v = b[:, 0] | (b[:, 1] << 8) | (b[:, 2] << 16)
        v = np.where(v & 0x800000, v - 0x1000000, v)
        x = (v.astype(np.float64) / 8388608.0).astype(np.float32)  # / 2^23 (assumption recorded in SURVEY.md 8c)
    elif bits == 32:
        x = (np.frombuffer(pcm, dtype="<i4").astype(np.float64) / 2147483648.0).astype(np.float32)
    else:
        raise ValueError(f"unsupported WAV format tag={tag} bits={bits}")
    if channels > 1:
        x = x.reshape(-1, channels)[:, 0].copy()
    return x, rate


def decimate(y: np.ndarray) -> dict:
    n = len(y)
    return {"head": y[:512].copy(), "transition": y[46000:50096].copy(), "strided": y[::37].copy(), "n": np.int64(n)}


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default="/root/reference")
    args = ap.parse_args()
    ref = Path(args.reference)
    from oracle import oracle  # noqa: E402

    models = {}
    for name, fn in MODELS.items():
        models[name] = json.loads((ref / "example_models" / fn).read_text())
    a2 = json.loads((ref / "example_models" / "A2.nam").read_text())
    for name, idx in A2_SUBMODELS.items():
        models[name] = a2["config"]["submodels"][idx]["model"]
    for name, nam in models.items():
        save_model(name, nam)

    x, rate = read_wav_pcm(ref / "example_audio" / "input.wav")
    assert rate == 48000 and len(x) == 96000
    np.savez_compressed(HERE / "input_wav.npz", x=x, sample_rate=np.int32(rate))
    print(f"wrote tests/golden/input_wav.npz: {len(x)} samples, peak {np.abs(x).max():.6f}")

    outs = {}
    for name in ("wavenet", "wavenet_a1_standard", "lstm", "a2_lite", "a2_full", "wavenet_condition_dsp"):
        for fast in (False, True):
            m = oracle.OracleModel.from_dict(models[name], fast_tanh=fast)
            if m.in_channels != 1 or m.out_channels != 1:
                continue
            m.reset(48000.0, 64)  # render.cpp protocol: Reset(sr, 64) then 64-frame blocks
            y = m.run(x, 64)
            for k, v in decimate(y).items():
                outs[f"{name}.{'fast' if fast else 'exact'}.{k}"] = v
            print(f"oracle {name} fast_tanh={fast}: range [{y.min():.6f}, {y.max():.6f}] y[0]={y[0]:.9g}")
    np.savez_compressed(HERE / "oracle_outputs.npz", **outs)
    print("wrote tests/golden/oracle_outputs.npz")


if __name__ == "__main__":
    main()
