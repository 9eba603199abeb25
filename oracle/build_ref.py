"""TEST / BASELINE INFRASTRUCTURE ONLY -- vendors the reference's own hot-path modules into oracle/_ref/ so that the
UNMODIFIED reference code can be timed next to the CUDA path on the GPU box (`bench.py --impl reference`, kind
"reference") and used as a second checker there.  /root/reference exists only in the build container; oracle/_ref/ is
git-ignored (the reference's sources never enter the history) but not gpurun-ignored, so it travels with the snapshot
like the built .so.  Nothing under segmif_b200/ imports it.

    python -m oracle.build_ref            # copies the files listed below, byte for byte, and writes a manifest

`__graft_entry__.build()` calls this when /root/reference is present."""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref")
SRC = os.environ.get("SEGMIF_REFERENCE_SRC", "/root/reference")
# the modules on the path SURVEY.md 8(a) names (+ the config surface); nothing else of the reference is needed
FILES = ["core/mix_transformer.py", "core/segformer_head.py", "core/model_fusion.py", "core/Entropy.py", "core/loss.py",
         "pytorch_ssim/__init__.py", "lap_loss.py", "datasets/imutils.py", "datasets/voc_fusion3.py"]


def build(src=SRC, dst=DST):
    if not os.path.isfile(os.path.join(src, "core", "model_fusion.py")):
        return None                                           # not in the build container: keep whatever is there
    manifest = {}
    for rel in FILES:
        out = os.path.join(dst, rel)
        os.makedirs(os.path.dirname(out), exist_ok=True)
        shutil.copyfile(os.path.join(src, rel), out)
        manifest[rel] = hashlib.sha256(open(out, "rb").read()).hexdigest()
    with open(os.path.join(dst, "MANIFEST.json"), "w") as f:
        json.dump({"source": "JinyuanLiu-CV/SegMiF (unmodified copies)", "sha256": manifest}, f, indent=1)
    return dst


if __name__ == "__main__":
    print(build() or f"{SRC} not present; nothing copied", file=sys.stderr)
