"""Compile the UNMODIFIED reference modules of the hot path to byte-code under oracle/_ref/ — TEST / BASELINE
INFRASTRUCTURE ONLY (never imported by unipose_b200/).

    python oracle/build_ref.py          # needs /root/reference (the build container); no-op elsewhere

The reference is pure Python, so "building" it is `py_compile` of the files where they lie under /root/reference;
only the compiled outputs (sourceless `*.pyc`, importable like `*.so` extension modules) land in oracle/_ref/, which
is git-ignored but travels to the GPU box with the working tree.  No reference source is copied.  With it present,
`bench.py --impl reference` and the `cpu_baseline` leg time the reference's own eager graph on the host cores
(`kind: "reference"`); without it they time the oracle port (`kind: "port"`).

Files: model/unipose.py, model/uniposeLSTM.py, model/modules/{wasp,waspVideo,decoder}.py,
model/modules/backbone/{__init__,resnet}.py, utils/evaluate.py.
"""
from __future__ import annotations

import os
import py_compile
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
OUT = os.path.join(ROOT, "oracle", "_ref")
FILES = [
    "model/unipose.py", "model/uniposeLSTM.py", "model/modules/wasp.py", "model/modules/waspVideo.py",
    "model/modules/decoder.py", "model/modules/backbone/__init__.py", "model/modules/backbone/resnet.py",
    "utils/evaluate.py",
]


def build(force: bool = False) -> bool:
    """Returns True when oracle/_ref holds the compiled reference (built now or earlier)."""
    if not os.path.isdir(os.path.join(REF, "model")):
        return have_ref()
    for rel in FILES:
        src = os.path.join(REF, rel)
        # utils/ is a regular package whose __init__ pulls in matplotlib: evaluate.py is compiled as a top-level module
        dst_rel = "ref_evaluate.pyc" if rel == "utils/evaluate.py" else rel[:-3] + ".pyc"
        dst = os.path.join(OUT, dst_rel)
        if not force and os.path.exists(dst) and os.path.getmtime(dst) >= os.path.getmtime(src):
            continue
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        py_compile.compile(src, cfile=dst, dfile=rel, doraise=True)
    return True


def have_ref() -> bool:
    return all(os.path.exists(os.path.join(OUT, ("ref_evaluate.pyc" if f == "utils/evaluate.py" else f[:-3] + ".pyc")))
               for f in FILES)


def import_reference():
    """(unipose class, uniposeLSTM module, evaluate module) of the compiled reference; raises if absent."""
    if not have_ref():
        raise RuntimeError("oracle/_ref is missing: run `python oracle/build_ref.py` where /root/reference exists")
    if OUT not in sys.path:
        sys.path.insert(0, OUT)
    from model.modules.backbone import resnet
    resnet.model_zoo.load_url = lambda *a, **k: {}      # offline: resnet.py:142 would download the ImageNet weights
    from model.unipose import unipose as RefUnipose
    import model.uniposeLSTM as ref_lstm
    import ref_evaluate
    return RefUnipose, ref_lstm, ref_evaluate


if __name__ == "__main__":
    ok = build(force="--force" in sys.argv)
    print("oracle/_ref:", "ready" if ok else "unavailable (no /root/reference)")
