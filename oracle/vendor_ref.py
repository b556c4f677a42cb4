"""Vendor the reference's own module files into oracle/_ref/ (git-ignored, NOT gpurun-ignored).

TEST / BASELINE INFRASTRUCTURE, NOT PRODUCT.  The reference is pure Python; nothing is compiled.  This recipe
copies, byte for byte, the handful of files the graph-block path lives in from where they lie under
/root/reference into oracle/_ref/src/, so that the GPU box (which has no /root/reference) can time the
UNMODIFIED reference modules as the CPU baseline (`bench.py --impl reference`, kind "reference") and as the
eager-PyTorch-on-B200 baseline.  oracle/_ref/ never enters the git history (.gitignore) and nothing under
xggm_b200/ reads it.

    python oracle/vendor_ref.py            # run in the build container; __graft_entry__.build() calls it too

Files (relative to /root/reference/src):
    module/graph_generative_modeling.py  module/gcn.py  module/gin.py  module/gat.py  module/graph_utils.py
    lxrt/modeling.py (GeLU, BertLayerNorm, VisualFeatEncoder)   lxrt/file_utils.py (imported by modeling.py)
    lxrt/optimization.py (BertAdam)      vqa/vqacpv2.py (loss_func / compute_kl_loss are cut out with `ast`;
                                         the module itself is never imported: param.py parses argv)
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = "/root/reference/src"
DST = os.path.join(HERE, "_ref")
FILES = ["module/graph_generative_modeling.py", "module/gcn.py", "module/gin.py", "module/gat.py",
         "module/graph_utils.py", "lxrt/modeling.py", "lxrt/file_utils.py", "lxrt/optimization.py", "vqa/vqacpv2.py"]


def available():
    return os.path.isdir(REF_SRC)


def vendored():
    """True when oracle/_ref holds every file of the recipe."""
    return all(os.path.isfile(os.path.join(DST, "src", f)) for f in FILES)


def vendor(verbose=False):
    if not available():
        return vendored()
    manifest = {}
    for f in FILES:
        src, dst = os.path.join(REF_SRC, f), os.path.join(DST, "src", f)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        manifest[f] = hashlib.sha256(open(dst, "rb").read()).hexdigest()
    with open(os.path.join(DST, "MANIFEST.json"), "w") as fh:
        json.dump({"source": REF_SRC, "sha256": manifest}, fh, indent=1)
    if verbose:
        print(f"vendored {len(FILES)} reference files into {DST}")
    return True


if __name__ == "__main__":
    ok = vendor(verbose=True)
    sys.exit(0 if ok else 1)
