"""Recipe for oracle/_ref/: the reference's OWN hot-path modules, taken from where they lie under /root/reference.

    python oracle/build_ref.py          # also run by __graft_entry__.build()

The reference is pure Python (26 loose files, no setup.py / pyproject, nothing to compile), and its hot path --
`network.py` (model) and `loss.py` (Dice / focal losses) -- imports nothing but torch.  This script copies those two
files UNMODIFIED into the git-ignored oracle/_ref/ (not tracked, not part of the product; it travels to the GPU box with
the gpurun snapshot like a built .so) so that

  * `bench.py --impl reference` and bench.py's `cpu_baseline` / `cudnn_bar` legs time the reference's own code
    (`cpu_baseline.kind == "reference"`) instead of the oracle port, and
  * tests/test_oracle_golden.py can cross-check the oracle port against it when it is present.

When /root/reference is absent (the GPU box) nothing is done: whatever the snapshot brought is used; if oracle/_ref is
missing altogether the callers fall back to the oracle port (`kind == "port"`).  TEST / MEASUREMENT INFRASTRUCTURE ONLY:
nothing under 3d-unet-renal-anatomy-extraction_b200/ imports oracle/.
"""
import hashlib
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("U3D_REFERENCE_DIR", "/root/reference")
DST = os.path.join(HERE, "_ref")
FILES = ("network.py", "loss.py")


def build() -> bool:
    if not all(os.path.isfile(os.path.join(REF, f)) for f in FILES):
        return os.path.isdir(DST)
    os.makedirs(DST, exist_ok=True)
    lines = []
    for f in FILES:
        shutil.copyfile(os.path.join(REF, f), os.path.join(DST, f))
        with open(os.path.join(DST, f), "rb") as fh:
            lines.append(f"{hashlib.sha256(fh.read()).hexdigest()}  {f}")
    with open(os.path.join(DST, "SHA256SUMS"), "w") as fh:
        fh.write("\n".join(lines) + "\n")
    return True


def load():
    """(network, loss) modules of the reference from oracle/_ref, or None when the directory is not there."""
    if not all(os.path.isfile(os.path.join(DST, f)) for f in FILES):
        return None
    import importlib.util
    mods = []
    for f in FILES:
        name = "u3d_reference_" + f[:-3]
        if name in sys.modules:
            mods.append(sys.modules[name])
            continue
        spec = importlib.util.spec_from_file_location(name, os.path.join(DST, f))
        m = importlib.util.module_from_spec(spec)
        sys.modules[name] = m
        spec.loader.exec_module(m)
        mods.append(m)
    return tuple(mods)


if __name__ == "__main__":
    print("oracle/_ref:", "ready" if build() else "not built (no /root/reference here)")
