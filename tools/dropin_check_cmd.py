#!/usr/bin/env python
"""Prints a shell command that runs the UNMODIFIED reference Python layers on the B200 drop-in on a GPU box.

/root/reference exists only in the build container, the GPU only on the gpurun box, and reference
sources must never be copied into this repo.  So the four files the test needs travel INSIDE THE COMMAND
LINE (base64 of a tar.gz, ~20 KB), are unpacked under /tmp on the box, and tests/test_dropin_reference.py
runs against them:

    gpurun --timeout 900 -- "$(python tools/dropin_check_cmd.py)"

The log lands in gpurun_out/r02/dropin_reference_check.log (copied to profiles/ by hand)."""
import base64
import io
import os
import tarfile

REF = "/root/reference"
FILES = ("ray.py", "den.py", "beam.py", "ft2.py")

buf = io.BytesIO()
with tarfile.open(fileobj=buf, mode="w:gz") as tar:
    for f in FILES:
        tar.add(os.path.join(REF, f), arcname=f)
b64 = base64.b64encode(buf.getvalue()).decode()
print("set -x; mkdir -p /tmp/ref_unmodified gpurun_out/r02 && "
      f"echo {b64} | base64 -d | tar xz -C /tmp/ref_unmodified && sha256sum /tmp/ref_unmodified/*.py && "
      "REF_UNMODIFIED_DIR=/tmp/ref_unmodified timeout 800 python -m pytest tests/test_dropin_reference.py -v -rs "
      "-W ignore::DeprecationWarning 2>&1 | tee gpurun_out/r02/dropin_reference_check.log | tail -15")
