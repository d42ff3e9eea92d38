"""ORACLE — test / baseline infrastructure only (never imported by the product path).

Recipe that stages the UNMODIFIED reference's hot-path modules next to the oracle so they travel to the GPU box:

    python -m oracle.build_ref          # also run by __graft_entry__.build() when /root/reference is present

The reference (Diamondfan/CTC_pytorch) is pure Python with no setup.py, so there is nothing to compile or pip-install; the
files below are copied byte for byte from where they lie under /root/reference into the git-ignored directory oracle/_ref/
(listed in .gitignore, NOT in .gpurunignore: it ships with the gpurun snapshot exactly like a built .so, and never enters the
history). `bench.py --impl reference` and the `-m gpu` tests import them through oracle/ref_shim.py; a MANIFEST with the
sha256 of every staged file is written so a run can state which bytes it executed.
"""
import hashlib
import json
import os
import shutil

SRC_ROOT = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
DST_ROOT = os.path.join(HERE, "_ref")
FILES = [
    "timit/models/model_ctc.py",     # CTC_Model, BatchRNN, LayerCNN, compute_wer
    "timit/utils/ctcDecoder.py",     # Decoder, GreedyDecoder, BeamDecoder
    "timit/utils/BeamSearch.py",     # ctcBeamSearch
    "timit/utils/NgramLM.py",        # LanguageModel
    "timit/utils/data_loader.py",    # create_input (imported by train_ctc.py)
    "timit/utils/tools.py",          # make_context / skip_feat
    "timit/steps/train_ctc.py",      # run_epoch: the training hot loop itself
]


def staged():
    return os.path.exists(os.path.join(DST_ROOT, "MANIFEST.json"))


def stage(force=False):
    """Copy the files (if the reference tree is present) and write the manifest. Returns the staging root or None."""
    if not os.path.isdir(SRC_ROOT):
        return DST_ROOT if staged() else None
    manifest = {}
    for rel in FILES:
        src, dst = os.path.join(SRC_ROOT, rel), os.path.join(DST_ROOT, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        if force or not os.path.exists(dst) or open(src, "rb").read() != open(dst, "rb").read():
            shutil.copyfile(src, dst)
        manifest[rel] = hashlib.sha256(open(dst, "rb").read()).hexdigest()
    with open(os.path.join(DST_ROOT, "MANIFEST.json"), "w") as fh:
        json.dump({"source": SRC_ROOT, "files": manifest}, fh, indent=1)
    return DST_ROOT


if __name__ == "__main__":
    print(stage(force=True))
