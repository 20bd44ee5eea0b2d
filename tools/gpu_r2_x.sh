#!/bin/bash
# config-5 shape on one GPU: where do 20.9 ms per 1.25e7 rays go?
timeout 600 python tools/kernel_sweep.py --nx 201 --nrays 12500000 --L 512 --cells both --only hist64/static/divc/hoist,hist64/refill/divc/hoist,endpoints/static/divc/hoist,endpoints/refill/divc/hoist,extras/static/divc/hoist 2>&1
