"""TEST INFRASTRUCTURE ONLY - recipe that stages the UNMODIFIED reference for the GPU box.

The reference (HankYe/Once-for-Both) is pure Python: nothing to compile.  `/root/reference` exists only in the build
container, so this recipe copies its Python sources byte for byte into `oracle/_ref/` (git-ignored: reference sources never
enter the history; NOT gpurun-ignored: the directory travels to the GPU box next to the built .so).  `oracle/ref_shim.py`
imports the reference from there when `/root/reference` is absent.  `__graft_entry__.build()` runs this when the reference is
present; on the GPU box the prebuilt copy is used as it is.

    python oracle/make_ref.py        # -> oracle/_ref/{engine,losses,optim,utils,lr_sched,lr_decay,samplers}.py, models/*.py

A manifest with the sha256 of every copied file is written next to them (`oracle/_ref/MANIFEST.json`) so that a test can
check that what runs on the GPU box is the reference as shipped.
"""
import hashlib
import json
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference"
DST = os.path.join(HERE, "_ref")
FILES = ["engine.py", "losses.py", "optim.py", "utils.py", "lr_sched.py", "lr_decay.py", "samplers.py", "search.py",
         "finetune.py", "datasets.py", "LICENSE",
         "models/__init__.py", "models/base_model.py", "models/layers.py", "models/model.py", "models/pos_embed.py",
         "models/vision_transformer.py"]


def stage(src=SRC, dst=DST):
    if not os.path.isdir(src):
        return None
    manifest = {}
    for rel in FILES:
        s, d = os.path.join(src, rel), os.path.join(dst, rel)
        os.makedirs(os.path.dirname(d), exist_ok=True)
        with open(s, "rb") as f:
            data = f.read()
        manifest[rel] = hashlib.sha256(data).hexdigest()
        if not os.path.exists(d) or open(d, "rb").read() != data:
            if os.path.exists(d):
                os.chmod(d, 0o644)
            shutil.copyfile(s, d)
    with open(os.path.join(dst, "MANIFEST.json"), "w") as f:
        json.dump({"source": src, "sha256": manifest}, f, indent=1, sort_keys=True)
    return dst


if __name__ == "__main__":
    out = stage()
    print(out if out else f"{SRC} not present: nothing staged")
