"""Build recipe of `oracle/_ref/`: the reference's own hot-path modules, byte-compiled FROM WHERE THEY LIE under /root/reference.

    python tools/make_oracle_ref.py            # (build container only; __graft_entry__.build() runs it when the tree is mounted)

The reference is Python source with no packaging, and `/root/reference` does not exist on the GPU box.  This recipe compiles the
modules SURVEY.md §8(c) lists (`compile()` + `marshal`, no source text is copied) into `oracle/_ref/vgqa/...*.bin` — a build
OUTPUT, git-ignored, which travels to the GPU box with the snapshot like the repo's own `.so` (`*.pyc` files do not travel, hence
the suffix).  There `tests/golden/ref_loader.py` imports the code objects through a small meta-path finder, so that bench.py can
time the REFERENCE'S OWN PyTorch modules (`cpu_baseline.kind = "reference"`, and the same
modules in eager bf16 on the B200) instead of the numpy port.  Nothing under `vgqa_b200/` ever imports it.
"""
import marshal
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("VGQA_REFERENCE_ROOT", "/root/reference")
OUT = os.path.join(ROOT, "oracle", "_ref")

# The hot path and what its modules import from the reference package, the Video-Swin extractor whose last stage
# csrc/swin.cu restates, and — for BASELINE configs[0], the FULL VSTGNet.forward on the host cores (tools/full_forward_cpu.py) — the
# rest of vgqa/core, vgqa/utils and vgqa/config.  (vgqa/data, vgqa/inference, tools/, app/ are not needed and not compiled.)
def _files():
    out = []
    for sub in ("core", "utils", "config"):
        for dirpath, _, names in os.walk(os.path.join(REF, "vgqa", sub)):
            for n in sorted(names):
                if n.endswith(".py"):
                    out.append(os.path.relpath(os.path.join(dirpath, n), REF))
    out.append("vgqa/training/evaluator.py")
    return sorted(out)


def build(verbose: bool = True) -> bool:
    if not os.path.isdir(os.path.join(REF, "vgqa", "core", "decoder")):
        if verbose:
            print(f"make_oracle_ref: {REF} is not mounted here — keeping whatever oracle/_ref already holds")
        return False
    if os.path.isdir(OUT):
        shutil.rmtree(OUT)
    FILES = _files()
    for rel in FILES:
        dst = os.path.join(OUT, rel[:-3] + ".bin")
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        with open(os.path.join(REF, rel), "rb") as f:
            code = compile(f.read(), rel, "exec", dont_inherit=True)
        with open(dst, "wb") as f:
            f.write(marshal.dumps(code))
    with open(os.path.join(OUT, "BUILD_INFO"), "w") as f:
        f.write(f"python {sys.version_info[0]}.{sys.version_info[1]}\nbyte-compiled from {REF} by tools/make_oracle_ref.py\n" + "\n".join(FILES) + "\n")
    if verbose:
        print(f"make_oracle_ref: {len(FILES)} modules compiled into {OUT}")
    return True


if __name__ == "__main__":
    build()
