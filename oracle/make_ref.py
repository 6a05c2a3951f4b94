"""
TEST / BENCHMARK INFRASTRUCTURE -- NOT PRODUCT CODE.

Recipe that makes the UNMODIFIED reference available to ``bench.py --impl reference`` and to the CPU baseline leg on
the GPU box, where /root/reference does not exist: it copies the reference's own Python package files for the hot
path (nasrec/supernet/*.py, nasrec/utils/*.py, nasrec/searcher/tokenizer.py and the shipped best-model configs)
byte for byte into ``oracle/_ref/`` -- an untracked build output (``.gitignore``) that travels with the gpurun snapshot
like the built ``.so`` -- and writes a manifest with the sha256 of every file.  Nothing under oracle/_ref is ever
committed, edited or imported by the product; ``bench.py`` imports it only to time the reference on the host cores
(``cpu_baseline.kind == "reference"``).  Run by ``__graft_entry__.build()`` whenever /root/reference is present.
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference"
DST = os.path.join(HERE, "_ref")
FILES = ["nasrec/supernet/modules.py", "nasrec/supernet/supernet.py", "nasrec/supernet/utils.py",
         "nasrec/utils/config.py", "nasrec/utils/data_pipes.py", "nasrec/utils/io_utils.py",
         "nasrec/utils/lr_schedule.py", "nasrec/utils/train_utils.py", "nasrec/searcher/tokenizer.py"]
CONFIG_DIRS = ["nasrec/configs"]


def vendor(src: str = SRC, dst: str = DST) -> bool:
    if not os.path.isdir(os.path.join(src, "nasrec")):
        return False
    manifest = {}
    for rel in FILES:
        out = os.path.join(dst, rel)
        os.makedirs(os.path.dirname(out), exist_ok=True)
        shutil.copyfile(os.path.join(src, rel), out)
        manifest[rel] = hashlib.sha256(open(out, "rb").read()).hexdigest()
    for d in CONFIG_DIRS:
        for root, _dirs, names in os.walk(os.path.join(src, d)):
            for n in names:
                if n.endswith(".json"):
                    rel = os.path.relpath(os.path.join(root, n), src)
                    out = os.path.join(dst, rel)
                    os.makedirs(os.path.dirname(out), exist_ok=True)
                    shutil.copyfile(os.path.join(src, rel), out)
                    manifest[rel] = hashlib.sha256(open(out, "rb").read()).hexdigest()
    # package markers (the reference is run as scripts from its repo root and has no __init__.py files)
    for pkg in ("nasrec", "nasrec/supernet", "nasrec/utils", "nasrec/searcher"):
        open(os.path.join(dst, pkg, "__init__.py"), "a").close()
    with open(os.path.join(dst, "MANIFEST.json"), "w") as f:
        json.dump({"source": src, "files": manifest}, f, indent=1, sort_keys=True)
    return True


def import_reference():
    """Import the vendored reference (sys.path entry + the two shims of SURVEY.md 8c: fvcore is imported at module
    scope by nasrec/utils/train_utils.py, np.int is used by searcher/tokenizer.py).  Returns the ``nasrec`` package
    or None when oracle/_ref has not been built."""
    import types
    if not os.path.isfile(os.path.join(DST, "nasrec", "supernet", "supernet.py")):
        return None
    if DST not in sys.path:
        sys.path.insert(0, DST)
    if "fvcore" not in sys.modules:
        fv = types.ModuleType("fvcore")
        fvnn = types.ModuleType("fvcore.nn")
        fvnn.FlopCountAnalysis = object
        fv.nn = fvnn
        sys.modules["fvcore"] = fv
        sys.modules["fvcore.nn"] = fvnn
    import numpy as np
    if not hasattr(np, "int"):
        np.int = int
    import nasrec                                   # noqa: F401  (the vendored reference)
    import nasrec.supernet.supernet                 # noqa: F401
    return nasrec


if __name__ == "__main__":
    print("vendored" if vendor() else "no /root/reference here: oracle/_ref left as it is")
