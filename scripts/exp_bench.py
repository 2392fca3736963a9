#!/usr/bin/env python
"""bench.py against the A/B library (libcadr_b200_exp.so, built with -DCADR_B200_EXPERIMENTS): the experiment kernels are
selected through CADR_B200_CULL_VARIANT / CADR_B200_SMALL_DIRECT / CADR_B200_DIAG_NOEVAL, which only that library reads.
usage: [ENV=...] python scripts/exp_bench.py <bench.py arguments>"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cadr_b200 import _capi, build  # noqa: E402

_capi.LIB_PATH = build.build_cuda(experiments=True)
import bench  # noqa: E402

sys.exit(bench.main())
