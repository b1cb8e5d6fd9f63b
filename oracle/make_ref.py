#!/usr/bin/env python
"""Recipe for oracle/_ref/: the reference's OWN implementation of the path, as a travelling CPU baseline.

TEST / BENCH INFRASTRUCTURE ONLY (see oracle/einx_oracle.py): nothing under ei-nexus_official_b200/ imports it.

The reference is pure Python, so there is nothing to compile: the leaf modules the path lives in (and the extractor modules around it) are copied
verbatim from the reference checkout into oracle/_ref/ (git-ignored -- the sources never enter this repository's
history -- but shipped to the GPU box with the working tree, like the built libeinx.so).  oracle/ref_arm.py loads
them from there with stub parent packages (core/modules/__init__.py pulls kornia / hydra /
lightning, which are not installed and not needed by these functions).

    python oracle/make_ref.py            # (re)creates oracle/_ref/ from /root/reference (or $EINX_REFERENCE)

Files (reference paths): core/modules/utils/{detector_util,homography,descriptor_util,util}.py,
core/modules/matchers/MNN.py, datasets/representations.py.
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref")
FILES = [
    "core/modules/utils/detector_util.py",
    "core/modules/utils/homography.py",
    "core/modules/utils/descriptor_util.py",
    "core/modules/utils/util.py",
    "core/modules/matchers/MNN.py",
    "datasets/representations.py",
    # the extractor modules around the path (conv backbone + heads in stock PyTorch, Padder / filter glue): the
    # integration test runs their forward() with and without einx.patch_reference()
    "core/modules/event_extractors/EventExtractors.py",
    "core/modules/net/backbone.py",
    "core/modules/net/detector_head.py",
    "core/modules/net/descriptor_head.py",
    "core/modules/net/vgg.py",
    "core/modules/net/pointnet.py",
    "core/modules/net/conv.py",
]


def make_ref(reference: str = None, quiet: bool = False) -> bool:
    """Copy the leaf modules; returns False (and leaves any existing copy alone) when there is no reference checkout."""
    ref = reference or os.environ.get("EINX_REFERENCE", "/root/reference")
    if not all(os.path.isfile(os.path.join(ref, f)) for f in FILES):
        if not quiet:
            print(f"make_ref: no reference checkout at {ref}; keeping {DST} as it is", file=sys.stderr)
        return False
    for f in FILES:
        dst = os.path.join(DST, f)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(ref, f), dst)
    with open(os.path.join(DST, "SOURCE.txt"), "w") as fh:
        fh.write("verbatim copies made by oracle/make_ref.py from the reference checkout; not part of this repository\n")
    return True


if __name__ == "__main__":
    ok = make_ref()
    print(("created " if ok else "not created: ") + DST)
