"""Install the UNMODIFIED reference into the git-ignored ``baseline/_ref/`` (the reference has no setup.py / pyproject, so
``pip install --target`` has nothing to build: this script is that step).  Files are copied byte for byte from
``/root/reference`` where they lie; nothing under ``baseline/_ref`` is ever committed (``.gitignore``), but the directory
travels to the GPU box with the repo snapshot.  Used by

  * ``tests/test_gpu_model_dropin.py``  -- EncodecModel / compress.py with the quantizer swapped,
  * ``bench.py``'s CPU arm (``cpu_baseline.kind == "reference"``) and ``gpu_eager_reference``.

    python scripts/install_reference.py [--src /root/reference]
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import shutil

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEST = os.path.join(ROOT, "baseline", "_ref")
# what the RVQ hot path and its callers need: the quantizer, the model around it, the .ecdc writer
ITEMS = ["quantization", "modules", "model.py", "compress.py", "binary.py", "distrib.py", "utils.py", "LICENSE"]


def install(src: str = "/root/reference", dest: str = DEST) -> dict:
    if not os.path.isdir(src):
        raise FileNotFoundError(f"reference tree {src} not found")
    if os.path.isdir(dest):
        shutil.rmtree(dest)
    os.makedirs(dest)
    manifest = {}
    for item in ITEMS:
        s, d = os.path.join(src, item), os.path.join(dest, item)
        if os.path.isdir(s):
            shutil.copytree(s, d, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
        else:
            shutil.copy2(s, d)
    for base, _, files in os.walk(dest):
        for f in sorted(files):
            p = os.path.join(base, f)
            manifest[os.path.relpath(p, dest)] = hashlib.sha256(open(p, "rb").read()).hexdigest()[:16]
    with open(os.path.join(dest, "MANIFEST.json"), "w") as fh:
        json.dump({"source": src, "files": manifest}, fh, indent=1, sort_keys=True)
    return manifest


def import_reference(dest: str = DEST):
    """Import the installed reference (``quantization``, ``model``, ``compress``, ``binary`` as top-level modules, the way
    the reference's own scripts see them) and return them in a namespace.  ``soundfile`` (needed only by utils.py's
    file helpers, absent from this image) is stubbed."""
    import importlib
    import sys
    import types
    if not os.path.isdir(os.path.join(dest, "quantization")):
        raise FileNotFoundError(f"{dest} is not populated: run scripts/install_reference.py where /root/reference exists")
    if dest not in sys.path:
        sys.path.insert(0, dest)
    for name in ("soundfile",):
        if name not in sys.modules:
            try:
                importlib.import_module(name)
            except Exception:
                sys.modules[name] = types.ModuleType(name)
    ns = types.SimpleNamespace()
    for name in ("distrib", "binary", "quantization", "modules", "utils", "model", "compress"):
        setattr(ns, name, importlib.import_module(name))
    for name in ("distrib", "binary", "quantization", "model", "compress"):
        path = os.path.abspath(getattr(ns, name).__file__)
        assert path.startswith(os.path.abspath(dest)), f"{name} resolved to {path}, not to the installed reference"
    return ns


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--src", default="/root/reference")
    a = ap.parse_args()
    m = install(a.src)
    print(f"installed {len(m)} reference files into {DEST}")
