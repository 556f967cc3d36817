"""Recipe: stage the reference's OWN pure-PyTorch modules for the benchmarked path under ``oracle/_ref/``.

Test / baseline infrastructure only (see oracle/__init__.py).  ``/root/reference`` exists in the build container but
not on the GPU box, and the reference has no build system (nothing to ``pip install``), so this recipe stages the few
files the benchmarked path (BASELINE.json configs[1]: mnist/compose_scores.py) executes -- the expert module and the
schedule -- byte for byte into ``oracle/_ref/`` (git-ignored, so no reference source enters the history; NOT
gpurun-ignored, so it travels with the snapshot like a built ``.so``).  ``bench.py --impl reference`` and the
``cpu_baseline`` / ``gpu_eager_baseline`` legs then time the reference's own ``UNet`` and schedule functions
(``cpu_baseline.kind = "reference"``); when ``oracle/_ref`` is absent they fall back to the oracle port (``"port"``).

    python -m oracle.build_ref          # run by __graft_entry__.build() whenever /root/reference is present
"""
import hashlib
import importlib.util
import json
import os
import shutil
import sys

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
# path in the reference -> what it provides on the benchmarked path
FILES = {
    "mnist/models/unet_small.py": "UNet / ResBlock / SinusoidalPosEmb (the expert of compose_scores.py:19-24)",
    "mnist/schedule.py": "dlog_alphadt / beta / sigma (compose_scores.py:40-43)",
    "shapes/models/unet_small.py": "conditional UNet (shapes/compose_images_ddim.py, compose_images_ito.py experts)",
    "shapes/schedule_2.py": "alpha / sigma / g2 of the shapes samplers",
}


def build(verbose=False):
    """Stage FILES; returns OUT, or None when the reference is not present (GPU box: the staged copy is used as is)."""
    if not os.path.isdir(REF):
        return OUT if os.path.isdir(OUT) else None
    manifest = {}
    for rel in FILES:
        src, dst = os.path.join(REF, rel), os.path.join(OUT, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        manifest[rel] = hashlib.sha256(open(src, "rb").read()).hexdigest()
        if verbose:
            print("staged", rel)
    json.dump(manifest, open(os.path.join(OUT, "MANIFEST.json"), "w"), indent=1)
    return OUT


def _load(rel, name):
    path = os.path.join(OUT, rel)
    if not os.path.exists(path):
        return None
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def load_mnist():
    """(reference UNet class, reference schedule module) from the staged copy, or None when it is absent."""
    unet = _load("mnist/models/unet_small.py", "_cdm_ref_mnist_unet")
    sched = _load("mnist/schedule.py", "_cdm_ref_mnist_schedule")
    if unet is None or sched is None:
        return None
    return unet.UNet, sched


if __name__ == "__main__":
    print(build(verbose=True))
