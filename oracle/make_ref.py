#!/usr/bin/env python
"""Recipe for oracle/_ref/: the UNMODIFIED reference files that its own CPU implementation of the hot path needs
(dlrm_s_pytorch_comm_grad.DLRM_Net + sgd_quantized_gradients_parallel_comm + quantization_supp and their import
closure, 13 files), copied byte for byte from /root/reference so that `bench.py --impl reference` and the
`cpu_baseline` leg can time the reference ITSELF on the GPU box's host cores (where /root/reference does not exist).

    python oracle/make_ref.py            # also run by __graft_entry__.build() when /root/reference is present

Test infrastructure, like everything under oracle/: oracle/_ref/ is git-ignored (never part of the history, never
imported by the product package) but not gpurun-ignored, so it travels with the snapshot.  MANIFEST.json lists the
source path, size and sha256 of every file."""
import hashlib
import json
import os
import shutil
import sys

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref")
FILES = [
    "dlrm_s_pytorch_comm_grad.py", "sgd_quantized_gradients_parallel_comm.py", "extend_distributed.py",
    "dlrm_data_pytorch.py", "data_utils.py", "data_loader_terabyte.py", "mlperf_logger.py",
    "optim/rwsadagrad.py", "quantization_supp/quant_modules_not_quantize_grad.py", "quantization_supp/quant_utils.py",
    "quantization_supp/full_precision_modules.py", "tricks/md_embedding_bag.py", "tricks/qr_embedding_bag.py",
]


def main():
    if not os.path.isdir(REF):
        print(f"{REF} not present: oracle/_ref left as it is", file=sys.stderr)
        return 0
    manifest = {}
    for rel in FILES:
        src, dst = os.path.join(REF, rel), os.path.join(DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        with open(dst, "rb") as f:
            b = f.read()
        manifest[rel] = {"source": src, "bytes": len(b), "sha256": hashlib.sha256(b).hexdigest()}
    for pkg in ("optim", "quantization_supp", "tricks"):               # packages where the reference has an __init__
        init = os.path.join(REF, pkg, "__init__.py")
        if os.path.exists(init):
            shutil.copyfile(init, os.path.join(DST, pkg, "__init__.py"))
    with open(os.path.join(DST, "MANIFEST.json"), "w") as f:
        json.dump(manifest, f, indent=1, sort_keys=True)
    print(f"oracle/_ref: {len(FILES)} files, {sum(v['bytes'] for v in manifest.values())} bytes")
    return 0


if __name__ == "__main__":
    sys.exit(main())
