#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 300 python tools/dbg/dbg_refgpu2.py 2>&1 | tail -40
